"""dat_b200 LayerNorm kernels against torch's LayerNorm on the same
inputs: fp32 to 1e-5 relative, bf16 I/O within bf16 rounding; gradients likewise."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("C", [32, 64, 96, 128, 256, 512, 1024])
@pytest.mark.parametrize("mode", ["fp32", "autocast", "bf16_in"])
def test_layernorm_proxy_matches_torch(C, mode):
    from dat_segmentation_b200.layernorm import LayerNormProxy, TorchLayerNormProxy
    torch.manual_seed(C)
    mine, ref = LayerNormProxy(C).cuda(), TorchLayerNormProxy(C).cuda()
    with torch.no_grad():
        mine.norm.weight.uniform_(0.5, 1.5)
        mine.norm.bias.normal_()
    ref.load_state_dict(mine.state_dict())
    B, H, W = 3, 13, 9
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2) * 2 + 0.3   # channels-last strided, like in situ
    if mode == "bf16_in":
        x = x.bfloat16()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    dy = torch.randn(B, C, H, W, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
        ya, yb = mine(xa), ref(xb)
    assert ya.dtype == yb.dtype and ya.shape == yb.shape
    ya.backward(dy.to(ya.dtype))
    yb.backward(dy.to(yb.dtype))
    tol = 1e-5 if mode != "bf16_in" else 1e-2
    assert _rel(ya, yb) < tol
    assert _rel(xa.grad, xb.grad) < (2e-5 if mode != "bf16_in" else 2e-2)
    assert _rel(mine.norm.weight.grad, ref.norm.weight.grad) < 2e-5 if mode != "bf16_in" else 2e-2
    assert _rel(mine.norm.bias.grad, ref.norm.bias.grad) < 2e-5 if mode != "bf16_in" else 2e-2


def test_layernorm_large_rows_deterministic():
    from dat_segmentation_b200.layernorm import LayerNormProxy
    m = LayerNormProxy(256).cuda()
    x = torch.randn(16, 32, 32, 256, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    outs = []
    for _ in range(2):
        m.zero_grad()
        x.grad = None
        y = m(x)
        y.backward(torch.ones_like(y) * 0.5 + y.detach())
        outs.append((y.detach().clone(), x.grad.clone(), m.norm.weight.grad.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))
