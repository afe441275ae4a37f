"""The strided 3 x 3 convolutions of the stem / down-projections (dat.py:213-218,264-274) on the dat_b200 kernels
(im2col + tcgen05 GEMMs + col2im, dat_segmentation_b200/conv.py) against the library convolution: forward under bf16
autocast against F.conv2d on the bf16-rounded operands, gradients against fp32 autograd by relative L2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,cin,cout,H,W,bias", [
    (2, 64, 128, 32, 32, False),     # down-projection 0 -> 1 shape (small map)
    (2, 128, 256, 16, 16, False),
    (1, 256, 512, 16, 16, False),
    (2, 32, 64, 64, 64, True),       # second stem convolution (K = 288 padded to 320)
    (2, 64, 128, 30, 34, True),      # odd output sizes 15 x 17
    (3, 8, 32, 20, 12, True),        # tiny channel count, 32 output channels (CUDA-core weight gradient)
])
def test_conv3x3s2_matches_library(B, cin, cout, H, W, bias):
    from dat_segmentation_b200.conv import Conv3x3s2CL
    torch.manual_seed(cin + cout + H)
    m = Conv3x3s2CL(cin, cout, bias=bias).cuda()
    x = torch.randn(B, H, W, cin, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)     # physically channel-last
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    assert y.dtype == torch.bfloat16 and y.shape == (B, cout, (H - 1) // 2 + 1, (W - 1) // 2 + 1)
    xr = x.detach().clone().requires_grad_(True)
    wr = m.weight.detach().clone().requires_grad_(True)
    br = m.bias.detach().clone().requires_grad_(True) if bias else None
    # forward: same bf16-rounded operands, fp32 accumulation
    y_ref = F.conv2d(xr.bfloat16().float(), wr.bfloat16().float(), br, 2, 1)
    assert (y.float() - y_ref).abs().max().item() < 2e-2 * max(1.0, y_ref.abs().max().item())
    assert _l2(y.float(), y_ref) < 5e-3
    dy = torch.randn_like(y_ref)
    y.backward(dy.bfloat16())
    F.conv2d(xr, wr, br, 2, 1).backward(dy)
    assert _l2(x.grad, xr.grad) < 2e-2
    assert _l2(m.weight.grad, wr.grad) < 2e-2
    if bias:
        assert _l2(m.bias.grad, br.grad) < 2e-2


def test_rgb_stem_convolution_matches_library():
    from dat_segmentation_b200.conv import Conv3x3s2CL
    torch.manual_seed(0)
    m = Conv3x3s2CL(3, 32).cuda()
    x = torch.randn(2, 3, 64, 96, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)
    wr = m.weight.detach().clone().requires_grad_(True)
    br = m.bias.detach().clone().requires_grad_(True)
    y_ref = F.conv2d(x.bfloat16().float(), wr.bfloat16().float(), br, 2, 1)
    assert _l2(y.float(), y_ref) < 5e-3
    dy = torch.randn_like(y_ref)
    y.backward(dy.bfloat16())
    F.conv2d(x, wr, br, 2, 1).backward(dy)
    assert _l2(m.weight.grad, wr.grad) < 2e-2 and _l2(m.bias.grad, br.grad) < 2e-2


def test_gelu_kernel_matches_library():
    from dat_segmentation_b200.conv import GeluCL
    torch.manual_seed(1)
    g = GeluCL()
    x = (torch.randn(2, 16, 12, 32, device="cuda") * 2).permute(0, 3, 1, 2).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y = g(x)
    yr = F.gelu(xr)
    assert y.dtype == torch.float32 and (y - yr).abs().max().item() < 1e-6
    dy = torch.randn_like(yr)
    y.backward(dy)
    yr.backward(dy)
    assert (x.grad - xr.grad).abs().max().item() < 1e-5
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yb = g(x.detach())
    assert yb.dtype == torch.bfloat16                                   # one bf16 rounding of the fp32 value
    assert ((yb.float() - yr.detach()).abs() <= yr.detach().abs() * 2.0 ** -8 + 1e-6).all()


def test_backbone_has_no_library_convolutions_left():
    from dat_segmentation_b200.backbone import build_dat
    from dat_segmentation_b200.conv import Conv3x3s2CL
    m = build_dat()
    convs = [c for c in list(m.patch_proj) + [dp[0] for dp in m.down_projs] if isinstance(c, torch.nn.Conv2d)]
    assert len(convs) == 5 and all(isinstance(c, Conv3x3s2CL) for c in convs)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_to_nchw_contiguous_matches_library_copy(dt):
    from dat_segmentation_b200.conv import to_nchw_contiguous
    torch.manual_seed(2)
    x = torch.randn(3, 20, 13, 40, device="cuda").to(dt).permute(0, 3, 1, 2).requires_grad_(True)   # (B, C, H, W) view
    y = to_nchw_contiguous(x)
    assert y.is_contiguous() and torch.equal(y, x.detach().contiguous())
    dy = torch.randn(3, 40, 20, 13, device="cuda").to(dt)
    y.backward(dy)
    assert torch.equal(x.grad, dy)
