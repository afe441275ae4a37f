"""dat_b200 fused residual + stochastic-depth kernel and the LayerNorm residual fork
 against the library expressions they replace."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("adt,xdt", [("bf16", "fp32"), ("bf16", "bf16"), ("fp32", "fp32"), ("bf16", None), ("fp32", None)])
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_scale_residual_matches_library(adt, xdt, layout):
    from dat_segmentation_b200.residual import scale_residual
    torch.manual_seed(0)
    DT = {"bf16": torch.bfloat16, "fp32": torch.float32}
    B, C, H, W = 5, 24, 7, 6

    def mk(dt):
        t = torch.randn(B, H, W, C, device="cuda").to(DT[dt])
        return t.permute(0, 3, 1, 2) if layout == "nhwc" else t.permute(0, 3, 1, 2).contiguous()

    a = mk(adt).requires_grad_(True)
    x = mk(xdt).requires_grad_(True) if xdt else None
    scale = torch.tensor([0.0, 1.0 / 0.7, 1.0 / 0.7, 0.0, 1.0 / 0.7], device="cuda")
    y = scale_residual(a, x, scale)
    ref = a * scale.view(B, 1, 1, 1).to(a.dtype)
    if x is not None:
        ref = ref + x
    assert y.dtype == ref.dtype and y.shape == ref.shape
    tol = 1e-6 if adt == "fp32" and xdt != "bf16" else 1e-2
    assert torch.allclose(y.float(), ref.float(), rtol=tol, atol=tol)
    dy = torch.randn_like(y)
    ga = torch.autograd.grad(y, [a] + ([x] if x is not None else []), dy)
    gr = torch.autograd.grad(ref, [a] + ([x] if x is not None else []), dy)
    for g, r in zip(ga, gr):
        assert g.dtype == r.dtype
        assert torch.allclose(g.float(), r.float(), rtol=tol, atol=tol)


@pytest.mark.parametrize("C", [64, 256])
def test_layernorm_fork_adds_residual_gradient(C):
    from dat_segmentation_b200.layernorm import LayerNormProxy
    torch.manual_seed(1)
    ln = LayerNormProxy(C).cuda()
    with torch.no_grad():
        ln.norm.weight.uniform_(0.5, 1.5)
        ln.norm.bias.normal_()
    x = torch.randn(3, 9, 5, C, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    w1 = torch.randn_like(x)
    w2 = torch.randn_like(x)
    xp, y = ln.forward_fork(x)
    (xp * w1 + y * w2).sum().backward()
    got = (x.grad.clone(), ln.norm.weight.grad.clone(), ln.norm.bias.grad.clone())
    x.grad = ln.norm.weight.grad = ln.norm.bias.grad = None
    (x * w1 + ln(x) * w2).sum().backward()
    ref = (x.grad, ln.norm.weight.grad, ln.norm.bias.grad)
    for g, r in zip(got, ref):
        assert torch.allclose(g, r, rtol=1e-5, atol=1e-5)
    # only the residual output used
    x.grad = None
    xp, y = ln.forward_fork(x)
    (xp * w1).sum().backward()
    assert torch.allclose(x.grad, w1)
