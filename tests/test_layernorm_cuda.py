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


@pytest.mark.gpu
@pytest.mark.parametrize("stream_dtype,branch_dtype", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32)])
@pytest.mark.parametrize("C", [64, 256, 512])
def test_residual_add_fused_into_the_norm_matches_the_two_kernel_path(stream_dtype, branch_dtype, C):
    """`x = drop_path(attn) + x; ln = LayerNorm(x)` (dat.py:147-151) in one kernel == scale_residual + LayerNorm fork:
    new stream, normed output and every gradient (branch, stream, gamma, beta)."""
    from dat_segmentation_b200.layernorm import LayerNormProxy
    from dat_segmentation_b200.residual import scale_residual
    torch.manual_seed(C)
    B, H, W = 3, 8, 12
    ln = LayerNormProxy(C).cuda()
    with torch.no_grad():
        ln.norm.weight.uniform_(0.5, 1.5)
        ln.norm.bias.uniform_(-0.5, 0.5)
    scale = torch.tensor([0.0, 1.0 / 0.7, 1.0 / 0.7], device="cuda")
    x0 = torch.randn(B, H, W, C, device="cuda", dtype=stream_dtype).permute(0, 3, 1, 2)
    a0 = torch.randn(B, H, W, C, device="cuda").to(branch_dtype).permute(0, 3, 1, 2)
    gy = torch.randn(B, H, W, C, device="cuda").to(branch_dtype).permute(0, 3, 1, 2)
    gx = torch.randn(B, H, W, C, device="cuda", dtype=stream_dtype).permute(0, 3, 1, 2)
    res = []
    for fused in (True, False):
        x, a = x0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
        ln.zero_grad(set_to_none=True)
        if fused:
            xo, y = ln.forward_residual_fork(a, x, scale, out_dtype=branch_dtype)
        else:
            xo, y = ln.forward_fork(scale_residual(a, x, scale), out_dtype=branch_dtype)
        (xo * gx).sum().backward(retain_graph=True)
        (y.float() * gy.float()).sum().backward()
        res.append([t.detach().float().clone() for t in (xo, y, x.grad, a.grad, ln.norm.weight.grad, ln.norm.bias.grad)])
    for f, u, name in zip(res[0], res[1], ["stream", "norm", "d stream", "d branch", "d gamma", "d beta"]):
        torch.testing.assert_close(f, u, rtol=2e-6, atol=2e-6, msg=lambda m, name=name: f"{name}: {m}")
