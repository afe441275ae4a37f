"""dat_b200 tensor-core 1x1 convolutions of the MLP blocks against
torch.nn.functional.conv2d in fp32: forward, data, weight and bias gradients.  The kernels run
bf16 / tf32 MMAs with fp32 accumulation, so the bar is the bf16 one of the block (2e-2 relative
to the tensor's max)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


# (cin, cout, B, H, W): the four DAT-T++ stage widths in both directions + ragged pixel counts
SHAPES = [(64, 256, 2, 32, 32), (256, 64, 2, 32, 32), (128, 512, 2, 16, 16), (512, 128, 2, 16, 16),
          (256, 1024, 2, 9, 7), (1024, 256, 2, 9, 7), (512, 2048, 1, 8, 8), (2048, 512, 1, 8, 8),
          (128, 512, 3, 5, 5),
          # DAT-S++ widths (C = 192, 384, 768; 96 x 4 = 384 columns out of a 96-wide input is not tileable)
          (192, 768, 2, 12, 12), (768, 192, 2, 12, 12), (384, 1536, 1, 9, 9), (1536, 384, 1, 9, 9),
          (768, 3072, 1, 8, 8), (3072, 768, 1, 8, 8),
          # many tiles: weight-stationary GEMM mode (fc1 / fc2 of stages 0-1 at full size)
          (64, 256, 2, 160, 160), (256, 64, 2, 160, 160), (128, 512, 1, 200, 200), (512, 128, 1, 200, 200)]


@pytest.mark.parametrize("cin,cout,B,H,W", SHAPES)
@pytest.mark.parametrize("xdt", ["fp32", "bf16"])
def test_pointwise_matches_conv2d(cin, cout, B, H, W, xdt):
    from dat_segmentation_b200.pointwise import PointwiseConvCL, _supported
    assert _supported(B * H * W, cout, cin)
    torch.manual_seed(cin + cout)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = PointwiseConvCL(cin, cout).cuda()
    dtype = torch.float32 if xdt == "fp32" else torch.bfloat16
    x = torch.randn(B, H, W, cin, device="cuda").permute(0, 3, 1, 2).to(dtype)
    dy = torch.randn(B, cout, H, W, device="cuda")
    xa = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ya = m(xa)
    assert ya.dtype == torch.bfloat16 and ya.shape == (B, cout, H, W)
    ya.backward(dy.to(ya.dtype))
    got = (ya.detach().float(), xa.grad.float().clone(), m.weight.grad.clone(), m.bias.grad.clone())
    assert xa.grad.dtype == x.dtype
    m.weight.grad = m.bias.grad = None
    xb = x.float().clone().requires_grad_(True)
    yb = F.conv2d(xb, m.weight, m.bias)
    yb.backward(dy.to(torch.bfloat16).float())
    ref = (yb.detach(), xb.grad, m.weight.grad, m.bias.grad)
    errs = {n: _rel(a, r) for n, a, r in zip(("y", "dx", "dw", "db"), got, ref)}
    print(cin, cout, xdt, {n: f"{e:.2e}" for n, e in errs.items()})
    assert all(e < 2e-2 for e in errs.values()), errs


def test_pointwise_fallbacks_match_library():
    """fp32 (no autocast) execution and untileable widths go through F.conv2d unchanged."""
    from dat_segmentation_b200.pointwise import PointwiseConvCL
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False
    m = PointwiseConvCL(96, 200).cuda()
    x = torch.randn(2, 96, 6, 6, device="cuda")
    assert torch.equal(m(x), F.conv2d(x, m.weight, m.bias))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
        assert torch.equal(y, F.conv2d(x, m.weight, m.bias))


@pytest.mark.parametrize("C,B,H,W", [(64, 3, 16, 16), (256, 2, 32, 32), (512, 2, 8, 12)])
def test_mlp_residual_in_the_gemm_epilogue_matches_the_two_kernel_path(C, B, H, W, monkeypatch):
    """`x = drop_path(mlp(ln)) + x` (dat.py:151-156) with the add fused into the last 1x1 conv's epilogue
    (TransformerMLPWithConv.forward_residual) == MLP + scale_residual: new stream, and the gradients of the MLP input,
    the stream and every MLP parameter."""
    from dat_segmentation_b200.backbone import TransformerMLPWithConv
    from dat_segmentation_b200.residual import scale_residual
    torch.manual_seed(C)
    mlp = TransformerMLPWithConv(C, 4, 0.0, b200_ops=True).cuda().train()
    scale = torch.tensor(([0.0, 1.0 / 0.7, 1.0 / 0.7])[:B], device="cuda")
    ln0 = torch.randn(B, H, W, C, device="cuda").bfloat16().permute(0, 3, 1, 2)
    x0 = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    g = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2)
    res = []
    for fused in (True, False):
        ln, x = ln0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        mlp.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if fused:
                fc2 = mlp.linear2[0]
                y = mlp.forward_residual(ln, x, scale)
            else:
                y = scale_residual(mlp(ln), x, scale)
        assert y.dtype == torch.float32
        (y * g).sum().backward()
        torch.cuda.synchronize()
        res.append([y.detach().clone(), ln.grad.float().clone(), x.grad.clone()] +
                   [p.grad.float().clone() for p in mlp.parameters()])
    names = ["stream", "d mlp input", "d stream"] + [n for n, _ in mlp.named_parameters()]
    for f, u, name in zip(res[0], res[1], names):
        torch.testing.assert_close(f, u, rtol=1e-5, atol=1e-5, msg=lambda m, name=name: f"{name}: {m}")
