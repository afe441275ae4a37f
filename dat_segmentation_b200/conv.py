"""The strided 3 x 3 convolutions around the DAT stages - conv stem (`models/backbones/dat.py:213-218`) and
down-projections (`dat.py:264-274`), SURVEY.md section 8f rank 3 - on the dat_b200 kernels: im2col / col2im data
movement (`csrc/conv_im2col.cu`) around the tcgen05 GEMMs of the 1x1 convolutions.

`Conv3x3s2CL` keeps an `nn.Conv2d(cin, cout, 3, 2, 1)`'s parameters (same state-dict keys / shapes, so reference
checkpoints load) and runs under bf16 autocast on CUDA:
    cols = im2col(x) (bf16, K = 9 cin padded to a multiple of 64);  Y = cols W2^T + b;
    dW2 = dY^T cols (+ db);  dcols = dY W2;  dx = col2im(dcols) (gather form, no atomics)
`GeluCL` is the stem's `nn.GELU` as one kernel each way.  Everything else (fp32 execution, CPU tensors, other kernel
sizes / strides) uses the library operator - these modules are a "next" row outside the parity-critical block.
"""
import ctypes as C
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi

__all__ = ["Conv3x3s2CL", "GeluCL", "to_nchw_contiguous"]

_CODE = {torch.float32: _cabi.DAT_F32, torch.bfloat16: _cabi.DAT_BF16}


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _bf16_autocast(x):
    return (x.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and x.dtype in _CODE)


class _ConvFn(torch.autograd.Function):
    """x: (B, H, W, C) channel-last contiguous, or the (B, 3, H, W) fp32 image when `rgb`;  -> y (B * Ho * Wo, Cout) bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, rgb):
        lib = _cabi.lib()
        dev = x.device
        if rgb:
            B, Cin, H, W = x.shape
        else:
            B, H, W, Cin = x.shape
        Cout = weight.shape[0]
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        M, Kp = B * Ho * Wo, lib.dat_conv3x3s2_kp(Cin)
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        with torch.cuda.device(dev):
            st = _stream(dev)
            cols = torch.empty(M, Kp, device=dev, dtype=torch.bfloat16)
            _cabi.check(lib.dat_im2col3x3s2(_ptr(x), _CODE[x.dtype], int(rgb), _ptr(cols), B, H, W, Cin, st), "dat_im2col3x3s2")
            w2 = torch.empty(Cout, Kp, device=dev, dtype=torch.bfloat16)
            _cabi.check(lib.dat_conv_weight_pack(_ptr(w32), _ptr(w2), Cout, Cin, st), "dat_conv_weight_pack")
            y = torch.empty(M, Cout, device=dev, dtype=torch.bfloat16)
            _cabi.check(lib.dat_pointwise_fwd_tc(_ptr(cols), _cabi.DAT_BF16, _ptr(w2), _ptr(b32), _ptr(y), _cabi.DAT_BF16,
                                                 M, Cout, Kp, st), "dat_pointwise_fwd_tc(conv)")
        ctx.save_for_backward(cols, w2)
        ctx.geom = (B, H, W, Cin, Cout, Kp, M, rgb, x.dtype)
        ctx.has_bias, ctx.wdtype = bias is not None, weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        cols, w2 = ctx.saved_tensors
        B, H, W, Cin, Cout, Kp, M, rgb, xdtype = ctx.geom
        dev = cols.device
        dy = dy.to(torch.bfloat16).contiguous()
        with torch.cuda.device(dev):
            st = _stream(dev)
            dw2 = torch.empty(Cout, Kp, device=dev, dtype=torch.float32)
            db = torch.empty(Cout, device=dev, dtype=torch.float32) if ctx.has_bias else None
            nb = lib.dat_pointwise_wgrad_tc_workspace_bytes(M, Cout, Kp)
            if nb > 0:        # tensor-core weight gradient (Cout and Kp multiples of 64)
                ws = torch.empty(nb, device=dev, dtype=torch.uint8)
                _cabi.check(lib.dat_pointwise_wgrad_tc(_ptr(dy), _ptr(cols), _ptr(dw2), _ptr(db), M, Cout, Kp, _ptr(ws), nb, st),
                            "dat_pointwise_wgrad_tc(conv)")
            else:             # the 32-channel stem convolution: CUDA-core kernel
                nb = lib.dat_pointwise_wgrad_workspace_bytes(M, Cout, Kp)
                ws = torch.empty(max(nb, 64), device=dev, dtype=torch.uint8)
                _cabi.check(lib.dat_pointwise_wgrad(_ptr(dy), _cabi.DAT_BF16, _ptr(cols), _cabi.DAT_BF16, _ptr(dw2), _ptr(db),
                                                    M, Cout, Kp, _ptr(ws), ws.numel(), st), "dat_pointwise_wgrad(conv)")
            dw = torch.empty(Cout, Cin, 3, 3, device=dev, dtype=torch.float32)
            _cabi.check(lib.dat_conv_weight_unpack(_ptr(dw2), _ptr(dw), Cout, Cin, st), "dat_conv_weight_unpack")
            dx = None
            if ctx.needs_input_grad[0]:
                if rgb:
                    raise NotImplementedError("gradient with respect to the input image of the RGB stem convolution")
                dcols = torch.empty(M, Kp, device=dev, dtype=torch.bfloat16)
                _cabi.check(lib.dat_pointwise_dgrad_tc(_ptr(dy), _ptr(w2), _ptr(dcols), _cabi.DAT_BF16, M, Cout, Kp, st),
                            "dat_pointwise_dgrad_tc(conv)")
                dx = torch.empty(B, H, W, Cin, device=dev, dtype=xdtype)
                _cabi.check(lib.dat_col2im3x3s2(_ptr(dcols), _ptr(dx), _CODE[xdtype], B, H, W, Cin, st), "dat_col2im3x3s2")
        return dx, dw.to(ctx.wdtype), (db.to(ctx.wdtype) if db is not None else None), None


class Conv3x3s2CL(nn.Conv2d):
    """nn.Conv2d(cin, cout, 3, 2, 1[, bias]) parameters; im2col + tcgen05 GEMMs under bf16 autocast on CUDA."""

    def __init__(self, cin, cout, bias=True):
        super().__init__(cin, cout, 3, 2, 1, bias=bias)

    def _own_path(self, x):
        if os.environ.get("DAT_B200_LIBRARY_CONVS") or not _bf16_autocast(x):
            return False
        B, Cin, H, W = x.shape
        M = B * ((H - 1) // 2 + 1) * ((W - 1) // 2 + 1)
        if self.out_channels % 32 != 0 or M < 64:
            return False
        if Cin == 3:
            return x.dtype == torch.float32 and not x.requires_grad
        return Cin % 8 == 0

    def forward(self, x):
        if not self._own_path(x):
            return F.conv2d(x, self.weight, self.bias, 2, 1)
        B, Cin, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        if Cin == 3:
            y = _ConvFn.apply(x.contiguous(), self.weight, self.bias, True)
        else:
            x_l = x.permute(0, 2, 3, 1)
            if not x_l.is_contiguous():
                x_l = x_l.contiguous()
            y = _ConvFn.apply(x_l, self.weight, self.bias, False)
        return y.reshape(B, Ho, Wo, self.out_channels).permute(0, 3, 1, 2)


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_dtype):
        lib = _cabi.lib()
        dev = x.device
        with torch.cuda.device(dev):
            y = torch.empty(x.shape, device=dev, dtype=out_dtype)
            _cabi.check(lib.dat_gelu_fwd(_ptr(x), _CODE[x.dtype], _ptr(y), _CODE[out_dtype], x.numel(), _stream(dev)), "dat_gelu_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        (x,) = ctx.saved_tensors
        dev = x.device
        if dy.dtype not in _CODE:
            dy = dy.float()
        dy = dy.contiguous()
        with torch.cuda.device(dev):
            dx = torch.empty_like(x)
            _cabi.check(lib.dat_gelu_bwd_mixed(_ptr(dy), _CODE[dy.dtype], _ptr(x), _ptr(dx), _CODE[x.dtype], x.numel(),
                                               _stream(dev)), "dat_gelu_bwd_mixed")
        return dx, None


class GeluCL(nn.GELU):
    """nn.GELU() of the conv stem: one kernel each way; writes bf16 under bf16 autocast (the rounding the following
    convolution applies to its input anyway).  Expects the (physically channel-last) output of LayerNormProxy."""

    def forward(self, x):
        if not (x.is_cuda and x.dtype in _CODE and x.numel() % 4 == 0) or os.environ.get("DAT_B200_LIBRARY_CONVS"):
            return super().forward(x)
        x_l = x.permute(0, 2, 3, 1)
        if not x_l.is_contiguous():
            return super().forward(x)
        out_dtype = torch.bfloat16 if _bf16_autocast(x) else x.dtype
        return _GeluFn.apply(x_l, out_dtype).permute(0, 3, 1, 2)


class _ToNCHWFn(torch.autograd.Function):
    """x_l (B, H, W, C) contiguous -> (B, C, H, W) contiguous."""

    @staticmethod
    def forward(ctx, x_l):
        lib = _cabi.lib()
        B, H, W, Cc = x_l.shape
        dev = x_l.device
        with torch.cuda.device(dev):
            y = torch.empty(B, Cc, H, W, device=dev, dtype=x_l.dtype)
            _cabi.check(lib.dat_transpose_pc(_ptr(x_l), _ptr(y), _CODE[x_l.dtype], B, H * W, Cc, _stream(dev)), "dat_transpose_pc")
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        B, Cc, H, W = dy.shape
        dev = dy.device
        if dy.dtype not in _CODE:
            dy = dy.float()
        dy = dy.contiguous()
        with torch.cuda.device(dev):
            dx = torch.empty(B, H, W, Cc, device=dev, dtype=dy.dtype)
            _cabi.check(lib.dat_transpose_pc(_ptr(dy), _ptr(dx), _CODE[dy.dtype], B, Cc, H * W, _stream(dev)), "dat_transpose_pc")
        return dx


def to_nchw_contiguous(x):
    """`x.contiguous()` for an NCHW-shaped view of channel-last storage (the backbone outputs, dat.py:308-309) as one
    dat_b200 transpose kernel each way; anything else goes to the library copy."""
    if x.is_contiguous():
        return x
    x_l = x.permute(0, 2, 3, 1)
    if not (x.is_cuda and x.dtype in _CODE and x_l.is_contiguous()) or os.environ.get("DAT_B200_LIBRARY_CONVS"):
        return x.contiguous()
    return _ToNCHWFn.apply(x_l)
