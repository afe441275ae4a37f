"""Residual add with stochastic depth in one pass: the
`x = drop_path(branch) + x` of `TransformerStage._inner_forward` (`models/backbones/dat.py:147-156`;
timm `drop_path`: per-sample Bernoulli mask / keep_prob) on the dat_b200 streaming kernel.

    y[b] = x[b] + a[b] * scale[b]          scale = mask / keep_prob  (ones when the path is kept)

dtype semantics are PyTorch's: with a residual the result has `result_type(a, x)` (fp32 stream +
bf16 branch -> fp32), without one ('X' blocks) the dtype of the branch.  Backward: d x = d y
(no kernel), d a = d y * scale[b] in the dtype of a.  CUDA only.
"""
import ctypes as C

import torch

from . import _cabi

__all__ = ["scale_residual", "drop_path_scale"]

_CODE = {torch.float32: _cabi.DAT_F32, torch.bfloat16: _cabi.DAT_BF16}


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _dense(t):
    """t if it is NCHW- or channel-last-contiguous (sample outermost), else a contiguous copy."""
    if t.is_contiguous() or (t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous()):
        return t
    return t.contiguous()


def _launch(a, x, scale, out_dtype):
    lib = _cabi.lib()
    B = a.shape[0]
    per = a.numel() // max(B, 1)
    y = torch.empty_like(a, dtype=out_dtype)       # keeps a's (dense) strides
    with torch.cuda.device(a.device):
        st = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
        _cabi.check(lib.dat_scale_residual(_ptr(a), _CODE[a.dtype], _ptr(x), _CODE[x.dtype] if x is not None else 0,
                                           _ptr(scale), _ptr(y), _CODE[out_dtype], B, per, st), "dat_scale_residual")
    return y


class _ScaleResidualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, scale):
        a = _dense(a)
        if x is not None:
            x = _dense(x)
            if a.stride() != x.stride():           # bring the branch to the layout of the stream
                a = torch.empty_like(x, dtype=a.dtype).copy_(a)
            out_dtype = torch.result_type(a, x)
        else:
            out_dtype = a.dtype
        ctx.save_for_backward(scale)
        ctx.a_dtype, ctx.has_x = a.dtype, x is not None
        return _launch(a, x, scale, out_dtype)

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        if dy.dtype not in _CODE:
            dy = dy.float()
        dy = _dense(dy)
        da = _launch(dy, None, scale, ctx.a_dtype) if ctx.needs_input_grad[0] else None
        return da, (dy if ctx.has_x else None), None


def scale_residual(a, x, scale):
    """a * scale[b] (+ x); a, x (B, ...) CUDA tensors (fp32 / bf16), scale (B,) fp32."""
    if not a.is_cuda:
        raise RuntimeError("scale_residual (dat_b200) runs on CUDA only")
    if a.dtype not in _CODE or (x is not None and x.dtype not in _CODE):
        raise NotImplementedError("scale_residual: float32 / bfloat16 only")
    if (a.numel() // max(a.shape[0], 1)) % 4 != 0:
        raise NotImplementedError("scale_residual: elements per sample must be a multiple of 4")
    return _ScaleResidualFn.apply(a, x, scale)


def drop_path_scale(B, p, training, device):
    """Per-sample scale of stochastic depth: Bernoulli(1 - p) / (1 - p), ones when inactive."""
    if not training or p == 0.0:
        return torch.ones(B, device=device, dtype=torch.float32)
    keep = 1.0 - p
    return torch.empty(B, device=device, dtype=torch.float32).bernoulli_(keep).div_(keep)
