"""dat_b200 channel-last depthwise conv kernels against
torch.nn.functional.conv2d: forward and all gradients, the three fusion modes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _ref(x, w, b, mode, k):
    y = F.conv2d(x, w, b, 1, k // 2, groups=x.shape[1])
    if mode >= 1:
        y = y + x
    if mode == 2:
        y = F.gelu(y)
    return y


@pytest.mark.parametrize("C,k,mode", [(64, 3, 1), (256, 3, 2), (1024, 3, 2), (64, 7, 0), (128, 7, 0), (96, 5, 1), (512, 3, 1)])
@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_dwconv_matches_conv2d(C, k, mode, dt):
    from dat_segmentation_b200.dwconv import dwconv_cl
    torch.manual_seed(C + k)
    B, H, W = 2, 11, 14
    dtype = torch.float32 if dt == "fp32" else torch.bfloat16
    torch.backends.cudnn.allow_tf32 = False
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2).to(dtype)
    w = (torch.randn(C, 1, k, k, device="cuda") / k).requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    dy = torch.randn(B, C, H, W, device="cuda").to(dtype)
    xa = x.clone().requires_grad_(True)
    ya = dwconv_cl(xa, w, b, mode, dtype)
    ya.backward(dy)
    got = (ya.detach(), xa.grad.clone(), w.grad.clone(), b.grad.clone())
    w.grad = b.grad = None
    xb = x.float().clone().requires_grad_(True)          # fp32 reference of the same (rounded) inputs
    yb = _ref(xb, w, b, mode, k)
    yb.backward(dy.float())
    ref = (yb.detach(), xb.grad, w.grad, b.grad)
    tol = 2e-5 if dt == "fp32" else 1.5e-2
    names = ("y", "dx", "dw", "db")
    errs = {n: _rel(a.float(), r) for n, a, r in zip(names, got, ref)}
    print(C, k, mode, dt, {n: f"{e:.2e}" for n, e in errs.items()})
    assert all(e < tol for e in errs.values()), errs


@pytest.mark.parametrize("C,mode,H,W", [(64, 2, 40, 37), (128, 1, 33, 18), (256, 0, 19, 50), (66, 2, 9, 9)])
@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_dwconv3_strips_and_fused_backward(C, mode, H, W, dt):
    """3x3 register-window kernels: several row strips per column, ragged right / bottom edges,
    per-CTA merged partial sums (C = 64) and per-strip ones (C = 66, 256)."""
    from dat_segmentation_b200.dwconv import dwconv_cl
    torch.manual_seed(C + H)
    torch.backends.cudnn.allow_tf32 = False
    B, k = 3, 3
    dtype = torch.float32 if dt == "fp32" else torch.bfloat16
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2).to(dtype)
    w = (torch.randn(C, 1, k, k, device="cuda") / k).requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    dy = torch.randn(B, C, H, W, device="cuda").to(dtype)
    xa = x.clone().requires_grad_(True)
    ya = dwconv_cl(xa, w, b, mode, dtype)
    ya.backward(dy)
    got = (ya.detach(), xa.grad.clone(), w.grad.clone(), b.grad.clone())
    w.grad = b.grad = None
    xb = x.float().clone().requires_grad_(True)
    yb = _ref(xb, w, b, mode, k)
    yb.backward(dy.float())
    ref = (yb.detach(), xb.grad, w.grad, b.grad)
    tol = 2e-5 if dt == "fp32" else 1.5e-2
    errs = {n: _rel(a.float(), r) for n, a, r in zip(("y", "dx", "dw", "db"), got, ref)}
    print(C, mode, H, W, dt, {n: f"{e:.2e}" for n, e in errs.items()})
    assert all(e < tol for e in errs.values()), errs


@pytest.mark.parametrize("C,H,W", [(64, 40, 36), (256, 19, 48), (66, 9, 8), (128, 33, 18)])
@pytest.mark.parametrize("dt", ["fp32", "bf16", "fp32_bf16"])
def test_dwconv7_row_stationary_kernels(C, H, W, dt):
    """7x7 'X' mixer kernels: several row strips, aligned (W % 4 == 0) and ragged widths, merged /
    per-strip partial sums, the fp32-in / bf16-out combination of the autocast backbone."""
    from dat_segmentation_b200.dwconv import dwconv_cl
    torch.manual_seed(C + H)
    torch.backends.cudnn.allow_tf32 = False
    B, k, mode = 3, 7, 0
    xdt = torch.bfloat16 if dt == "bf16" else torch.float32
    ydt = torch.float32 if dt == "fp32" else torch.bfloat16
    x = torch.randn(B, H, W, C, device="cuda").permute(0, 3, 1, 2).to(xdt)
    w = (torch.randn(C, 1, k, k, device="cuda") / k).requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    dy = torch.randn(B, C, H, W, device="cuda").to(ydt)
    xa = x.clone().requires_grad_(True)
    ya = dwconv_cl(xa, w, b, mode, ydt)
    ya.backward(dy)
    got = (ya.detach(), xa.grad.clone(), w.grad.clone(), b.grad.clone())
    assert got[1].dtype == xdt
    w.grad = b.grad = None
    xb = x.float().clone().requires_grad_(True)
    yb = _ref(xb, w, b, mode, k)
    yb.backward(dy.float())
    ref = (yb.detach(), xb.grad, w.grad, b.grad)
    tol = 2e-5 if dt == "fp32" else 1.5e-2
    errs = {n: _rel(a.float(), r) for n, a, r in zip(("y", "dx", "dw", "db"), got, ref)}
    print(C, H, W, dt, {n: f"{e:.2e}" for n, e in errs.items()})
    assert all(e < tol for e in errs.values()), errs
