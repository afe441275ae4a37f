"""LayerNormProxy on the dat_b200 kernels (the LayerNorm that feeds every
block, `models/utils/dat_blocks.py:229-240`, used at `models/backbones/dat.py:147,151`).

Same module surface as the reference (`self.norm = nn.LayerNorm(dim)` → state-dict keys
`norm.weight`, `norm.bias`; NCHW in, NCHW view of a channel-last buffer out).  CUDA only;
`TorchLayerNormProxy` is the library-operator twin used by the CPU baseline / CPU tests.
dtype semantics follow `torch.autocast`: LayerNorm computes and returns fp32 under autocast.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi

__all__ = ["LayerNormProxy", "TorchLayerNormProxy"]

_CODE = {torch.float32: _cabi.DAT_F32, torch.bfloat16: _cabi.DAT_BF16}


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


class _LayerNormFn(torch.autograd.Function):
    """x_l (..., C) contiguous → y_l (..., C) of `out_dtype`."""

    @staticmethod
    def forward(ctx, x_l, weight, bias, eps, out_dtype):
        lib = _cabi.lib()
        Cc = x_l.shape[-1]
        rows = x_l.numel() // Cc
        dev = x_l.device
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            y = torch.empty(x_l.shape, device=dev, dtype=out_dtype)
            mean = torch.empty(rows, device=dev, dtype=torch.float32)
            rstd = torch.empty(rows, device=dev, dtype=torch.float32)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_layernorm_fwd(_ptr(x_l), _CODE[x_l.dtype], _ptr(w32), _ptr(b32), _ptr(y),
                                              _CODE[out_dtype], _ptr(mean), _ptr(rstd), rows, Cc,
                                              float(eps), st), "dat_layernorm_fwd")
        ctx.save_for_backward(x_l, w32, mean, rstd)
        ctx.param_dtype = weight.dtype
        return y

    @staticmethod
    def backward(ctx, dy, dres=None):
        lib = _cabi.lib()
        x_l, w32, mean, rstd = ctx.saved_tensors
        Cc = x_l.shape[-1]
        rows = x_l.numel() // Cc
        dev = x_l.device
        if dy.dtype not in _CODE:
            dy = dy.float()
        dy = dy.contiguous()
        if dres is not None:      # gradient of the residual path around the norm: added in the same pass
            dres = dres.to(x_l.dtype).contiguous()
        with torch.cuda.device(dev):
            dx = torch.empty_like(x_l)
            dg = torch.empty(Cc, device=dev, dtype=torch.float32)
            db = torch.empty(Cc, device=dev, dtype=torch.float32)
            nbytes = lib.dat_layernorm_bwd_workspace_bytes(rows, Cc)
            ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_layernorm_bwd(_ptr(dy), _CODE[dy.dtype], _ptr(x_l), _CODE[x_l.dtype],
                                              _ptr(w32), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dres),
                                              _ptr(dg), _ptr(db), rows, Cc, _ptr(ws), nbytes, st),
                        "dat_layernorm_bwd")
        return dx, dg.to(ctx.param_dtype), db.to(ctx.param_dtype), None, None


class _LayerNormForkFn(torch.autograd.Function):
    """x_l -> (x_l, LayerNorm(x_l)): the residual stream and the normed branch input leave one node,
    so the backward adds the residual-path gradient inside the LayerNorm kernel instead of autograd
    running a separate accumulation pass over the stream (dat.py:147-156)."""

    @staticmethod
    def forward(ctx, x_l, weight, bias, eps, out_dtype):
        y = _LayerNormFn.forward(ctx, x_l, weight, bias, eps, out_dtype)
        return x_l, y

    @staticmethod
    def backward(ctx, dres, dy):
        if dy is None:
            return dres, None, None, None, None
        return _LayerNormFn.backward(ctx, dy, dres)


class _ResidualLayerNormForkFn(torch.autograd.Function):
    """(a_l, x_l, scale) -> (x_l + a_l * scale[b], LayerNorm(x_l + a_l * scale[b])): the residual add with stochastic
    depth that precedes a norm (`x = drop_path(attn) + x; ... layer_norms[2d+1](x)`, dat.py:147-151) inside the norm's
    kernel, forward and backward (the branch gradient d a = d x * scale leaves the LayerNorm backward kernel too)."""

    @staticmethod
    def forward(ctx, a_l, x_l, scale, weight, bias, eps, out_dtype):
        lib = _cabi.lib()
        Cc = x_l.shape[-1]
        rows = x_l.numel() // Cc
        dev = x_l.device
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            xout = torch.empty_like(x_l)
            y = torch.empty(x_l.shape, device=dev, dtype=out_dtype)
            mean = torch.empty(rows, device=dev, dtype=torch.float32)
            rstd = torch.empty(rows, device=dev, dtype=torch.float32)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_residual_layernorm_fwd(_ptr(a_l), _ptr(scale), rows // x_l.shape[0], _ptr(x_l),
                                                       _CODE[x_l.dtype], _ptr(w32), _ptr(b32), _ptr(xout), _ptr(y),
                                                       _CODE[out_dtype], _ptr(mean), _ptr(rstd), rows, Cc, float(eps), st),
                        "dat_residual_layernorm_fwd")
        ctx.save_for_backward(xout, w32, mean, rstd, scale)
        ctx.param_dtype, ctx.a_dtype = weight.dtype, a_l.dtype
        return xout, y

    @staticmethod
    def backward(ctx, dres, dy):
        lib = _cabi.lib()
        xout, w32, mean, rstd, scale = ctx.saved_tensors
        Cc = xout.shape[-1]
        rows = xout.numel() // Cc
        dev = xout.device
        if dy is None:                      # the norm's output was not used: only the residual path carries a gradient
            dy = torch.zeros(xout.shape, device=dev, dtype=ctx.a_dtype)
        dy = dy.to(ctx.a_dtype).contiguous()
        if dres is not None:
            dres = dres.to(xout.dtype).contiguous()
        with torch.cuda.device(dev):
            dx = torch.empty_like(xout)
            da = torch.empty(xout.shape, device=dev, dtype=ctx.a_dtype)
            dg = torch.empty(Cc, device=dev, dtype=torch.float32)
            db = torch.empty(Cc, device=dev, dtype=torch.float32)
            nbytes = lib.dat_layernorm_bwd_workspace_bytes(rows, Cc)
            ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_residual_layernorm_bwd(_ptr(dy), _CODE[dy.dtype], _ptr(xout), _CODE[xout.dtype],
                                                       _ptr(w32), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dres), _ptr(da),
                                                       _ptr(scale), rows // xout.shape[0], _ptr(dg), _ptr(db), rows, Cc,
                                                       _ptr(ws), nbytes, st), "dat_residual_layernorm_bwd")
        return da, dx, None, dg.to(ctx.param_dtype), db.to(ctx.param_dtype), None, None


class LayerNormProxy(nn.Module):
    """LayerNorm over the channels of an NCHW tensor on the dat_b200 kernels."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def _prep(self, x, out_dtype=None):
        if not x.is_cuda:
            raise RuntimeError("LayerNormProxy (dat_b200) runs on CUDA only; use TorchLayerNormProxy on CPU")
        if x.dtype not in _CODE:
            raise NotImplementedError(f"dtype {x.dtype} unsupported (float32 / bfloat16)")
        if out_dtype is None:
            out_dtype = torch.float32 if torch.is_autocast_enabled("cuda") else x.dtype
        x_l = x.permute(0, 2, 3, 1)
        if not x_l.is_contiguous():
            x_l = x_l.contiguous()
        return x_l, out_dtype

    def forward(self, x, out_dtype=None):
        """`out_dtype`: dtype of the result; default fp32 under autocast (library semantics).  A caller
        whose only consumer is an autocast convolution may ask for bf16 directly - the same rounding
        the convolution's input cast would apply, without the fp32 round trip through HBM."""
        x_l, out_dtype = self._prep(x, out_dtype)
        y_l = _LayerNormFn.apply(x_l, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype)
        return y_l.permute(0, 3, 1, 2)

    def forward_fork(self, x, out_dtype=None):
        """(x, LayerNorm(x)) for `branch(LN(x)) + x` call sites: use the returned x for the residual."""
        x_l, out_dtype = self._prep(x, out_dtype)
        x_p, y_l = _LayerNormForkFn.apply(x_l, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype)
        return x_p.permute(0, 3, 1, 2), y_l.permute(0, 3, 1, 2)


    def forward_residual_fork(self, a, x, scale, out_dtype=None):
        """(x + a * scale[b], LayerNorm(x + a * scale[b])) in one kernel; a (the branch output) must have the dtype the
        norm writes (`out_dtype`) and the layout of x, else the two-kernel path is taken."""
        from .residual import scale_residual
        x_l, out_dtype = self._prep(x, out_dtype)
        a_l = a.permute(0, 2, 3, 1)
        if (a.dtype != out_dtype or not a_l.is_contiguous() or a.dtype not in _CODE
                or torch.result_type(a, x) != x.dtype):
            return self.forward_fork(scale_residual(a, x, scale), out_dtype)
        x_p, y_l = _ResidualLayerNormForkFn.apply(a_l, x_l, scale, self.norm.weight, self.norm.bias, self.norm.eps, out_dtype)
        return x_p.permute(0, 3, 1, 2), y_l.permute(0, 3, 1, 2)


class TorchLayerNormProxy(nn.Module):
    """Library-operator twin (CPU baseline / CPU tests); identical parameters."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x, out_dtype=None):
        y = self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        return y if out_dtype is None else y.to(out_dtype)

    def forward_fork(self, x, out_dtype=None):
        return x, self.forward(x, out_dtype)

    def forward_residual_fork(self, a, x, scale, out_dtype=None):
        x = x + a * scale.view(-1, 1, 1, 1).to(a.dtype)
        return x, self.forward(x, out_dtype)
