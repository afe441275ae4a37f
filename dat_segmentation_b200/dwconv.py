"""Channel-last depthwise convolutions of the DAT backbone on the dat_b200 kernels
(SURVEY.md §8f ranks 2-3): the local perception unit (`dat.py:135-138`), the depthwise 3x3 +
residual + GELU in the middle of `TransformerMLPWithConv` (`dat_blocks.py:338-343`) and the
7x7 'X' mixer (`dat.py:118-121`).

`DepthwiseConvCL` keeps an `nn.Conv2d`'s parameters (same state-dict keys / shapes, so
reference checkpoints load) and runs `y = f(dwconv(x) + b [+ x])` in one kernel; CUDA only.
dtype semantics follow autocast: the convolution result is bf16 under autocast, except the LPU
whose residual add promotes to the dtype of x (fp32 residual stream).
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _cabi
from ._streams import grads_consumed_at_end_only, hold_until_join, serial as _serial, side_stream

__all__ = ["dwconv_cl", "DepthwiseConvCL"]

_CODE = {torch.float32: _cabi.DAT_F32, torch.bfloat16: _cabi.DAT_BF16}
MODE_PLAIN, MODE_RESIDUAL, MODE_RESIDUAL_GELU = 0, 1, 2


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


class _DwConvFn(torch.autograd.Function):
    """x_l (B,H,W,C) contiguous -> y_l (B,H,W,C) of out_dtype."""

    @staticmethod
    def forward(ctx, x_l, weight, bias, mode, out_dtype):
        lib = _cabi.lib()
        B, H, W, Cc = x_l.shape
        k = weight.shape[-1]
        dev = x_l.device
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        with torch.cuda.device(dev):
            y = torch.empty(x_l.shape, device=dev, dtype=out_dtype)
            z = torch.empty(x_l.shape, device=dev, dtype=out_dtype) if mode == MODE_RESIDUAL_GELU else None
            nbytes = lib.dat_dwconv_workspace_bytes(B, H, W, Cc, k)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_dwconv_fwd(_ptr(x_l), _CODE[x_l.dtype], _ptr(w32), _ptr(b32), _ptr(y), _ptr(z),
                                           _CODE[out_dtype], B, H, W, Cc, k, mode, 0, _ptr(ws), nbytes, st),
                        "dat_dwconv_fwd")
        ctx.save_for_backward(x_l, w32, z)
        ctx.mode, ctx.has_bias = mode, bias is not None
        ctx.wdtype = weight.dtype
        ctx.param_refs = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        x_l, w32, z = ctx.saved_tensors
        B, H, W, Cc = x_l.shape
        k = w32.shape[-1]
        dev = x_l.device
        mode = ctx.mode
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if k == 3 and Cc % 2 == 0 and not os.environ.get("DAT_B200_DWCONV_GENERIC"):
                # fused: dz = dy * gelu'(z), dx, dw, db in one pass over dy, z, x
                d_dt = z.dtype if z is not None else (dy.dtype if dy.dtype in _CODE else torch.float32)
                dy = dy.to(d_dt).contiguous()
                nbytes = lib.dat_dwconv_workspace_bytes(B, H, W, Cc, k)
                ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                dx = torch.empty_like(x_l)
                dw = torch.empty_like(w32)
                db = torch.empty(Cc, device=dev, dtype=torch.float32) if ctx.has_bias else None
                _cabi.check(lib.dat_dwconv_bwd(_ptr(x_l), _CODE[x_l.dtype], _ptr(dy), _ptr(z), _CODE[d_dt],
                                               _ptr(w32), _ptr(dx), _ptr(dw), _ptr(db), B, H, W, Cc, k, mode,
                                               _ptr(ws), nbytes, st), "dat_dwconv_bwd")
                return dx, dw.to(ctx.wdtype), (db.to(ctx.wdtype) if db is not None else None), None, None
            if mode == MODE_RESIDUAL_GELU:
                dy = dy.to(z.dtype).contiguous()
                dz = torch.empty_like(z)
                _cabi.check(lib.dat_gelu_bwd(_ptr(dy), _ptr(z), _ptr(dz), _CODE[z.dtype], dz.numel(), st),
                            "dat_gelu_bwd")
            else:
                dz = dy if dy.dtype in _CODE else dy.float()
                dz = dz.contiguous()
            nbytes = lib.dat_dwconv_workspace_bytes(B, H, W, Cc, k)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            # weight / bias gradient: off the critical path, on the side stream (own workspace)
            cur = torch.cuda.current_stream(dev)
            wst = cur if _serial() else side_stream(dev)
            if wst is not cur:
                wst.wait_stream(cur)         # dz was produced on the current stream
            with torch.cuda.stream(wst):
                sw = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                ws2 = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                dw = torch.empty_like(w32)
                db = torch.empty(Cc, device=dev, dtype=torch.float32) if ctx.has_bias else None
                _cabi.check(lib.dat_dwconv_wgrad(_ptr(x_l), _CODE[x_l.dtype], _ptr(dz), _CODE[dz.dtype], _ptr(dw),
                                                 _ptr(db), B, H, W, Cc, k, _ptr(ws2), nbytes, sw), "dat_dwconv_wgrad")
                dw = dw.to(ctx.wdtype)
                db = db.to(ctx.wdtype) if db is not None else None
            if wst is not cur:
                hold_until_join(x_l, dz, ws2)
            dx = torch.empty_like(x_l)
            # data gradient: the same kernel with the flipped filter, + dz for the residual modes
            _cabi.check(lib.dat_dwconv_fwd(_ptr(dz), _CODE[dz.dtype], _ptr(w32), None, _ptr(dx), None,
                                           _CODE[dx.dtype], B, H, W, Cc, k,
                                           MODE_PLAIN if mode == MODE_PLAIN else MODE_RESIDUAL, 1, _ptr(ws),
                                           nbytes, st), "dat_dwconv_fwd(dgrad)")
            if wst is not cur and not grads_consumed_at_end_only(*ctx.param_refs):
                cur.wait_stream(wst)   # dw / db are consumed right after this node: complete them first
                for t in (dw, db):     # allocated on the side stream, read (then freed) on the current one
                    if t is not None:
                        t.record_stream(cur)
        return dx, dw, db, None, None


def dwconv_cl(x, weight, bias, mode, out_dtype=None):
    """x (B,C,H,W) (any strides) -> (B,C,H,W) view of a channel-last result."""
    if not x.is_cuda:
        raise RuntimeError("dwconv_cl (dat_b200) runs on CUDA only")
    if x.dtype not in _CODE:
        raise NotImplementedError(f"dtype {x.dtype} unsupported (float32 / bfloat16)")
    if out_dtype is None:
        out_dtype = x.dtype
    x_l = x.permute(0, 2, 3, 1)
    if not x_l.is_contiguous():
        x_l = x_l.contiguous()
    return _DwConvFn.apply(x_l, weight, bias, mode, out_dtype).permute(0, 3, 1, 2)


class DepthwiseConvCL(nn.Conv2d):
    """nn.Conv2d(C, C, k, 1, k // 2, groups=C) parameters, dat_b200 kernels.  `mode`: 0 plain,
    1 + input (LPU), 2 gelu(. + input) (MLP middle)."""

    def __init__(self, channels, kernel_size, mode=MODE_PLAIN, keep_input_dtype=False):
        super().__init__(channels, channels, kernel_size, 1, kernel_size // 2, groups=channels)
        self.mode, self.keep_input_dtype = mode, keep_input_dtype

    def forward(self, x):
        if torch.is_autocast_enabled("cuda") and not self.keep_input_dtype:
            out_dtype = torch.get_autocast_dtype("cuda")
        else:
            out_dtype = x.dtype
        return dwconv_cl(x, self.weight, self.bias, self.mode, out_dtype)
