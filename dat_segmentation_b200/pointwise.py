"""1x1 convolutions of the DAT MLP blocks (`TransformerMLPWithConv.linear1/linear2`,
`models/utils/dat_blocks.py:316-348`) on the dat_b200 tcgen05 GEMMs.

`PointwiseConvCL` keeps an `nn.Conv2d(cin, cout, 1)`'s parameters (same state-dict keys) and runs
under bf16 autocast:
  forward   Y = X W^T + b      tf32 MMA straight on an fp32 X (LayerNorm output), bf16 MMA otherwise
  dX = dY W                    the same kernel against a transposed bf16 copy of W
  dW = dY^T X, db = colsum dY  one MN-major tensor-core pass over dY and X, deterministic reductions
Shapes the kernels cannot tile, fp32 (non-autocast) execution and CPU tensors use the library
convolution (`F.conv2d`) — this module is a "next row" outside the parity-critical block.
"""
import ctypes as C
import os as _os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from ._streams import grads_consumed_at_end_only, hold_until_join, serial as _serial, side_stream
from .weights import cached_bf16

__all__ = ["PointwiseConvCL"]

_CODE = {torch.float32: _cabi.DAT_F32, torch.bfloat16: _cabi.DAT_BF16}


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _supported(M, N, K):
    # forward / data-gradient tiling (gemm_tc.cu pick_bn: a tile of 32..256 columns must divide N, resp. K) and
    # the weight-gradient kernel's own predicate (64 x 64 boxes), asked through the C ABI
    return (M >= 64 and N % 32 == 0 and K % 32 == 0
            and _cabi.lib().dat_pointwise_wgrad_tc_workspace_bytes(M, N, K) > 0)


class _PointwiseFn(torch.autograd.Function):
    """x_l (M, K) contiguous (fp32 or bf16) -> y (M, N) bf16.  `w_bf`: the step's bf16 copy of the weight
    (weights.Bf16WeightCache) or None (cast here)."""

    @staticmethod
    def forward(ctx, x_l, weight, bias, w_bf):
        lib = _cabi.lib()
        M, K = x_l.shape
        N = weight.shape[0]
        dev = x_l.device
        w32 = weight.detach().float().reshape(N, K).contiguous()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        with torch.cuda.device(dev):
            y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            if x_l.dtype == torch.float32:
                w_op = w32                                   # tf32 MMA on the fp32 operands
            elif w_bf is not None:
                w_op = w_bf
            else:
                w_op = w_bf = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
                _cabi.check(lib.dat_cast_bf16(_ptr(w32), _ptr(w_op), N * K, _stream(dev)), "dat_cast_bf16")
            _cabi.check(lib.dat_pointwise_fwd_tc(_ptr(x_l), _CODE[x_l.dtype], _ptr(w_op), _ptr(b32), _ptr(y),
                                                 _cabi.DAT_BF16, M, N, K, _stream(dev)), "dat_pointwise_fwd_tc")
        ctx.save_for_backward(x_l, w32)
        ctx.w_bf = w_bf               # bf16 (N, K) copy, re-used by the data gradient (None: made in backward)
        ctx.has_bias, ctx.wdtype, ctx.wshape = bias is not None, weight.dtype, weight.shape
        ctx.param_refs = (weight, bias)   # backward only inspects .grad / hooks (see _streams.grads_consumed_at_end_only)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        x_l, w32 = ctx.saved_tensors[:2]
        M, K = x_l.shape
        N = w32.shape[0]
        dev = x_l.device
        dy = dy.to(torch.bfloat16).contiguous()
        with torch.cuda.device(dev):
            # dW = dY^T X (bf16 operands), on the side stream
            serial = _serial()
            cur = torch.cuda.current_stream(dev)
            wst = cur if serial else side_stream(dev)
            if not serial:
                wst.wait_stream(cur)         # dy (and x) are produced on the current stream
            with torch.cuda.stream(wst):
                sw = _stream(dev)
                if x_l.dtype == torch.float32:
                    xb = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
                    _cabi.check(lib.dat_cast_bf16(_ptr(x_l), _ptr(xb), M * K, sw), "dat_cast_bf16")
                else:
                    xb = x_l
                nbytes = lib.dat_pointwise_wgrad_tc_workspace_bytes(M, N, K)
                ws = torch.empty(max(nbytes, 64), device=dev, dtype=torch.uint8)
                dw = torch.empty(N, K, device=dev, dtype=torch.float32)
                db = torch.empty(N, device=dev, dtype=torch.float32) if ctx.has_bias else None
                # db = column sums of dY ride along in the same tensor-core pass
                _cabi.check(lib.dat_pointwise_wgrad_tc(_ptr(dy), _ptr(xb), _ptr(dw), _ptr(db), M, N, K, _ptr(ws),
                                                       ws.numel(), sw), "dat_pointwise_wgrad_tc")
                dw = dw.reshape(ctx.wshape).to(ctx.wdtype)
                db = db.to(ctx.wdtype) if db is not None else None
            if not serial:
                hold_until_join(dy, x_l, xb, ws)   # what the side stream reads stays alive until the join
            early_join = not serial and not grads_consumed_at_end_only(*ctx.param_refs)
            # dX = dY W on the current stream: the bf16 weight copy is read in place as an MN-major operand;
            # widths without a 64-multiple tile use a K-major product against a transposed copy
            st = _stream(dev)
            dx = torch.empty_like(x_l)
            if K % 64 == 0 and not _os.environ.get("DAT_B200_DGRAD_TRANSPOSE"):
                w_bf = ctx.w_bf
                if w_bf is None:
                    w_bf = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
                    _cabi.check(lib.dat_cast_bf16(_ptr(w32), _ptr(w_bf), N * K, st), "dat_cast_bf16")
                _cabi.check(lib.dat_pointwise_dgrad_tc(_ptr(dy), _ptr(w_bf), _ptr(dx), _CODE[dx.dtype], M, N, K, st),
                            "dat_pointwise_dgrad_tc")
            else:
                wT = torch.empty(K, N, device=dev, dtype=torch.bfloat16)
                _cabi.check(lib.dat_cast_transpose_bf16(_ptr(w32), _ptr(wT), N, K, st), "dat_cast_transpose_bf16")
                _cabi.check(lib.dat_pointwise_fwd_tc(_ptr(dy), _cabi.DAT_BF16, _ptr(wT), None, _ptr(dx),
                                                     _CODE[dx.dtype], M, K, N, st), "dat_pointwise_fwd_tc(dgrad)")
            if early_join:      # dw / db are consumed right after this node (accumulation, hooks): complete them first
                cur.wait_stream(wst)
                for t in (dw, db):      # allocated on the side stream, read (then freed) on the current one
                    if t is not None:
                        t.record_stream(cur)
        return dx, dw, db, None


class _PointwiseResidualFn(torch.autograd.Function):
    """(x_l (M, K) bf16, resid_l (M, N) fp32, scale (B,)) -> resid + scale[b] * bf16(x_l W^T + b): the last 1x1 conv of the
    MLP and the residual add with stochastic depth that follows it (`x = drop_path(mlp(x)) + x`, dat.py:151-156) in ONE
    kernel (the GEMM's epilogue reads the residual stream and writes the new one).  Backward: d resid = dy; the branch
    gradient dy * scale[b] (bf16, the scale_residual kernel) feeds the usual data / weight gradients."""

    @staticmethod
    def forward(ctx, x_l, weight, bias, w_bf, resid_l, scale):
        lib = _cabi.lib()
        M, K = x_l.shape
        N = weight.shape[0]
        dev = x_l.device
        w32 = weight.detach().float().reshape(N, K).contiguous()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        with torch.cuda.device(dev):
            y = torch.empty(M, N, device=dev, dtype=torch.float32)
            if w_bf is None:
                w_bf = torch.empty(N, K, device=dev, dtype=torch.bfloat16)
                _cabi.check(lib.dat_cast_bf16(_ptr(w32), _ptr(w_bf), N * K, _stream(dev)), "dat_cast_bf16")
            _cabi.check(lib.dat_pointwise_fwd_tc_residual(_ptr(x_l), _cabi.DAT_BF16, _ptr(w_bf), _ptr(b32), _ptr(resid_l),
                                                          _ptr(scale), M // scale.numel(), _ptr(y), M, N, K, _stream(dev)),
                        "dat_pointwise_fwd_tc_residual")
        ctx.save_for_backward(x_l, w32, scale)
        ctx.w_bf = w_bf
        ctx.has_bias, ctx.wdtype, ctx.wshape = bias is not None, weight.dtype, weight.shape
        ctx.param_refs = (weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _cabi.lib()
        scale = ctx.saved_tensors[2]
        dy = dy.float().contiguous()
        M, N = dy.shape
        dev = dy.device
        with torch.cuda.device(dev):
            dm = torch.empty(M, N, device=dev, dtype=torch.bfloat16)     # branch gradient dy * scale[b]
            _cabi.check(lib.dat_scale_residual(_ptr(dy), _cabi.DAT_F32, None, 0, _ptr(scale), _ptr(dm), _cabi.DAT_BF16,
                                               scale.numel(), (M // scale.numel()) * N, _stream(dev)), "dat_scale_residual")
        dx, dw, db, _ = _PointwiseFn.backward(ctx, dm)
        return dx, dw, db, None, dy, None


class PointwiseConvCL(nn.Conv2d):
    """nn.Conv2d(cin, cout, 1) parameters; tcgen05 GEMMs under bf16 autocast on CUDA."""

    def __init__(self, cin, cout):
        super().__init__(cin, cout, 1)

    def forward(self, x):
        B, K, H, W = x.shape
        N = self.out_channels
        use_tc = (x.is_cuda and torch.is_autocast_enabled("cuda")
                  and torch.get_autocast_dtype("cuda") == torch.bfloat16 and x.dtype in _CODE
                  and _supported(B * H * W, N, K))
        if not use_tc:
            return F.conv2d(x, self.weight, self.bias)
        x_l = x.permute(0, 2, 3, 1)
        if not x_l.is_contiguous():
            x_l = x_l.contiguous()
        y = _PointwiseFn.apply(x_l.reshape(B * H * W, K), self.weight, self.bias, cached_bf16(self))
        return y.reshape(B, H, W, N).permute(0, 3, 1, 2)

    def residual_fusable(self, x, resid):
        """True when `forward_residual` runs as one kernel: bf16 input, fp32 channel-last residual stream of the output
        shape, 64-column tiles."""
        B, K, H, W = x.shape
        N = self.out_channels
        return (x.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
                and x.dtype == torch.bfloat16 and resid.dtype == torch.float32 and resid.shape == (B, N, H, W)
                and x.permute(0, 2, 3, 1).is_contiguous() and resid.permute(0, 2, 3, 1).is_contiguous()
                and N % 64 == 0 and _supported(B * H * W, N, K) and not _os.environ.get("DAT_B200_NO_GEMM_RESIDUAL"))

    def forward_residual(self, x, resid, scale):
        """resid + scale[b] * conv1x1(x) (the caller checked `residual_fusable`)."""
        B, K, H, W = x.shape
        N = self.out_channels
        y = _PointwiseResidualFn.apply(x.permute(0, 2, 3, 1).reshape(B * H * W, K), self.weight, self.bias,
                                       cached_bf16(self), resid.permute(0, 2, 3, 1).reshape(B * H * W, N), scale)
        return y.reshape(B, H, W, N).permute(0, 3, 1, 2)
