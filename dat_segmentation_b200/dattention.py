"""Drop-in replacement for the reference's deformable-attention block.

Mirrors `DAttentionBaseline` of hehe717/DAT-Segmentation
(`models/utils/dat_blocks.py:19-227`): same 16 positional constructor arguments
(`:21-26`), same parameter names / shapes (state-dict compatible, so DAT++
checkpoints load with `strict=True`), same `forward(x) -> (y, None, None)` contract
(`:138,227`).  All arithmetic runs in the hand-written sm_100a kernels behind the C
ABI of `include/dat_b200.h`; this file only owns parameters, dtype / layout plumbing
and autograd wiring.  There is no PyTorch or CPU fallback: a CPU tensor, a missing
`libdat_b200.so` or an unsupported configuration raises.  The variant branches of the
reference (`use_pe=False`, `no_off`, `dwc_pe`, `fixed_pe`, `log_cpb`; `:57-59,84-99,
156-157,164-167,185-197,221-222`) are implemented with the reference's parameter names.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .weights import cached_bf16

__all__ = ["DAttentionBaseline", "LayerNormProxy"]


class LayerNormProxy(nn.Module):
    """LayerNorm over the channel dim of an NCHW tensor (dat_blocks.py:229-240).
    Inside the block it is only a parameter holder (`conv_offset.1.norm.*`); as a
    standalone module it returns a permuted (physically NHWC) view like the reference."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _fill_struct(struct, names, tensors):
    for n, t in zip(names, tensors):
        setattr(struct, n, t.data_ptr() if t is not None else 0)
    return struct


N_PARAMS = len(_cabi.PARAM_FIELDS)   # 13 common tensors + up to 3 of the position-encoding branch


class _BlockFn(torch.autograd.Function):
    """x_l (B,H,W,C) channel-last contiguous -> y_l (B,H,W,C) act dtype, pos (B,G,Ns,2)."""

    @staticmethod
    def forward(ctx, x_l, meta, *params):
        lib = _cabi.lib()
        B, H, W, Cc = x_l.shape
        act = meta["act_dtype"]
        desc = _cabi.BlockDesc(B, H, W, meta["n_heads"], meta["n_groups"], meta["stride"],
                               meta["ksize"], meta["table_h"], meta["table_w"],
                               float(meta["orf"]),
                               _cabi.DAT_BF16 if x_l.dtype == torch.bfloat16 else _cabi.DAT_F32,
                               _cabi.DAT_BF16 if act == torch.bfloat16 else _cabi.DAT_F32,
                               meta["pe_mode"], int(meta["no_off"]))
        hk, wk = C.c_int32(), C.c_int32()
        _cabi.check(lib.dat_sample_grid(C.byref(desc), C.byref(hk), C.byref(wk)), "dat_sample_grid")
        Ns, G = hk.value * wk.value, meta["n_groups"]
        dev = x_l.device
        assert len(params) == N_PARAMS
        p32 = [_f32c(p) if p is not None else None for p in params]
        with torch.cuda.device(dev):
            e = lambda *s, dt=act: torch.empty(s, device=dev, dtype=dt)
            f32 = torch.float32
            saved = [e(B, H * W, Cc), e(B, G, Ns, Cc // G, dt=f32), e(B, G, Ns, 2, dt=f32),
                     e(B, G, Ns, 2, dt=f32), e(B, Ns, Cc), e(B, Ns, Cc), e(B, Ns, Cc),
                     e(B, H * W, Cc), e(B, meta["n_heads"], H * W, dt=f32)]
            y_l = e(B, H, W, Cc)
            pstruct = _fill_struct(_cabi.BlockParams(), _cabi.PARAM_FIELDS, p32)
            w_bf = meta.get("w_bf16") if act == torch.bfloat16 else None
            if w_bf is not None:      # the step's bf16 operand copies of proj_q / k / v / out (weights.py)
                _fill_struct(pstruct, _cabi.BF16_FIELDS, w_bf)
            sstruct = _fill_struct(_cabi.BlockSaved(), _cabi.SAVED_FIELDS, saved)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            nbytes = lib.dat_block_fwd_workspace_bytes(C.byref(desc))
            ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
            _cabi.check(lib.dat_block_forward(C.byref(desc), C.byref(pstruct), _ptr(x_l), _ptr(y_l),
                                              C.byref(sstruct), _ptr(ws), nbytes, stream),
                        "dat_block_forward")
        ctx.meta = meta
        ctx.w_bf = w_bf
        ctx.desc = desc
        ctx.param_dtypes = [p.dtype if p is not None else None for p in params]
        ctx.present = [p is not None for p in params]
        ctx.save_for_backward(x_l, *[p for p in p32 if p is not None], *saved)
        pos = saved[3]
        ctx.mark_non_differentiable(pos)
        return y_l, pos

    @staticmethod
    def backward(ctx, dy_l, _dpos):
        lib = _cabi.lib()
        tensors = ctx.saved_tensors
        n_present = sum(ctx.present)
        it = iter(tensors[1:1 + n_present])
        x_l, saved = tensors[0], list(tensors[1 + n_present:])
        p32 = [next(it) if here else None for here in ctx.present]
        desc, act = ctx.desc, ctx.meta["act_dtype"]
        dev = x_l.device
        dy_l = dy_l.to(act).contiguous()
        with torch.cuda.device(dev):
            # no_off: the offset network is unused and frozen (dat_blocks.py:57-59) - its gradients stay None
            skip = range(5) if ctx.meta["no_off"] else ()
            grads = [torch.empty_like(p) if p is not None and i not in skip else None for i, p in enumerate(p32)]
            dx = torch.empty(x_l.shape, device=dev, dtype=torch.float32)
            nbytes = lib.dat_block_bwd_workspace_bytes(C.byref(desc))
            ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
            pstruct = _fill_struct(_cabi.BlockParams(), _cabi.PARAM_FIELDS, p32)
            if ctx.w_bf is not None:
                _fill_struct(pstruct, _cabi.BF16_FIELDS, ctx.w_bf)
            gstruct = _fill_struct(_cabi.BlockGrads(), _cabi.PARAM_FIELDS, grads)
            sstruct = _fill_struct(_cabi.BlockSaved(), _cabi.SAVED_FIELDS, saved)
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(lib.dat_block_backward(C.byref(desc), C.byref(pstruct), _ptr(x_l), _ptr(dy_l),
                                               C.byref(sstruct), _ptr(dx), C.byref(gstruct), _ptr(ws),
                                               nbytes, stream), "dat_block_backward")
        if dx.dtype != x_l.dtype:
            dx = dx.to(x_l.dtype)
        grads = [g if g is None or g.dtype == dt else g.to(dt) for g, dt in zip(grads, ctx.param_dtypes)]
        return (dx, None, *grads)


class DAttentionBaseline(nn.Module):
    """B200-native deformable attention; constructor signature of dat_blocks.py:21-26."""

    def __init__(self, q_size, kv_size, n_heads, n_head_channels, n_groups, attn_drop, proj_drop,
                 stride, offset_range_factor, use_pe, dwc_pe, no_off, fixed_pe, ksize, log_cpb,
                 stage_i):
        super().__init__()
        if n_head_channels != _cabi.HEAD_DIM:
            raise NotImplementedError(f"n_head_channels={n_head_channels}: kernels are built for 32 "
                                      "(every DAT++ variant, dat.py:57)")
        if n_heads % n_groups != 0:
            raise ValueError("n_heads must be a multiple of n_groups")
        self.fp16_enabled = False
        self.dwc_pe, self.fixed_pe, self.no_off, self.log_cpb, self.use_pe = dwc_pe, fixed_pe, no_off, log_cpb, use_pe
        self.n_head_channels = n_head_channels
        self.scale = n_head_channels ** -0.5
        self.n_heads = n_heads
        self.q_h, self.q_w = _pair(q_size)
        self.kv_h, self.kv_w = self.q_h // stride, self.q_w // stride   # kv_size ignored, :35-36
        self.nc = n_head_channels * n_heads
        self.n_groups = n_groups
        self.n_group_channels = self.nc // n_groups
        self.n_group_heads = n_heads // n_groups
        self.offset_range_factor = offset_range_factor
        self.ksize, self.stride, self.stage_i = ksize, stride, stage_i
        self.attn_drop_p, self.proj_drop_p = float(attn_drop), float(proj_drop)
        self.return_pos_ref = False  # opt-in: forward returns (y, pos, ref) instead of (y, None, None)
        cg = self.n_group_channels
        pad = ksize // 2 if ksize != stride else 0
        # parameter holders with the reference's names, shapes and default initialisation
        self.conv_offset = nn.Sequential(
            nn.Conv2d(cg, cg, ksize, stride, pad, groups=cg), LayerNormProxy(cg), nn.GELU(),
            nn.Conv2d(cg, 2, 1, 1, 0, bias=False))
        self.proj_q = nn.Conv2d(self.nc, self.nc, 1, 1, 0)
        self.proj_k = nn.Conv2d(self.nc, self.nc, 1, 1, 0)
        self.proj_v = nn.Conv2d(self.nc, self.nc, 1, 1, 0)
        self.proj_out = nn.Conv2d(self.nc, self.nc, 1, 1, 0)
        if no_off:   # dat_blocks.py:57-59
            for prm in self.conv_offset.parameters():
                prm.requires_grad_(False)
        # position-encoding branch, in the reference's precedence order (dat_blocks.py:84-104)
        if self.use_pe and not self.no_off:
            if self.dwc_pe:
                self.pe_mode = _cabi.PE_DWC
                self.rpe_table = nn.Conv2d(self.nc, self.nc, kernel_size=3, stride=1, padding=1, groups=self.nc)
            elif self.fixed_pe:
                self.pe_mode = _cabi.PE_FIXED
                self.rpe_table = nn.Parameter(torch.zeros(n_heads, self.q_h * self.q_w, self.kv_h * self.kv_w))
                nn.init.trunc_normal_(self.rpe_table, std=0.01)
            elif self.log_cpb:
                if self.n_group_heads > 16:
                    raise NotImplementedError("log_cpb: at most 16 heads per group")
                self.pe_mode = _cabi.PE_LOGCPB
                self.rpe_table = nn.Sequential(nn.Linear(2, 32, bias=True), nn.ReLU(inplace=True),
                                               nn.Linear(32, self.n_group_heads, bias=False))
            else:
                self.pe_mode = _cabi.PE_RPE
                self.rpe_table = nn.Parameter(torch.zeros(n_heads, self.q_h * 2 - 1, self.q_w * 2 - 1))
                nn.init.trunc_normal_(self.rpe_table, std=0.01)
        else:
            self.pe_mode = _cabi.PE_NONE
            self.rpe_table = None

    def _params(self):
        """The 16 slots of dat_block_params: 13 common tensors, then rpe_table / pe_b / pe_w2 (None = unused)."""
        co = self.conv_offset
        t = self.rpe_table
        if self.pe_mode == _cabi.PE_DWC:
            pe = (t.weight, t.bias, None)
        elif self.pe_mode == _cabi.PE_LOGCPB:
            pe = (t[0].weight, t[0].bias, t[2].weight)
        elif self.pe_mode == _cabi.PE_NONE:
            pe = (None, None, None)
        else:
            pe = (t, None, None)
        return (co[0].weight, co[0].bias, co[1].norm.weight, co[1].norm.bias, co[3].weight,
                self.proj_q.weight, self.proj_q.bias, self.proj_k.weight, self.proj_k.bias,
                self.proj_v.weight, self.proj_v.bias, self.proj_out.weight, self.proj_out.bias, *pe)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("DAttentionBaseline (dat_b200) runs on CUDA only; there is no CPU path")
        if self.training and self.attn_drop_p > 0:
            # dropout on the softmax probabilities lives inside the fused attention kernel (the reference draws its mask
            # over the materialised (B*h, HW, Ns) tensor, dat_blocks.py:217): not implemented; every shipped config uses 0
            raise NotImplementedError("attn_drop > 0 in training is not implemented")
        B, Cc, H, W = x.shape
        if Cc != self.nc:
            raise ValueError(f"expected {self.nc} channels, got {Cc}")
        if torch.is_autocast_enabled("cuda"):
            act = torch.get_autocast_dtype("cuda")
            if act != torch.bfloat16:
                raise NotImplementedError("autocast dtype must be bfloat16")
        else:
            act = x.dtype
        if act not in (torch.float32, torch.bfloat16) or x.dtype not in (torch.float32, torch.bfloat16):
            raise NotImplementedError(f"dtype {x.dtype} unsupported (float32 / bfloat16)")
        x_l = x.permute(0, 2, 3, 1)
        if not x_l.is_contiguous():   # in situ x is already physically NHWC (dat.py:147)
            x_l = x_l.contiguous()
        has_table = self.pe_mode in (_cabi.PE_RPE, _cabi.PE_FIXED)
        meta = dict(n_heads=self.n_heads, n_groups=self.n_groups, stride=self.stride,
                    ksize=self.ksize, table_h=self.rpe_table.shape[1] if has_table else 1,
                    table_w=self.rpe_table.shape[2] if has_table else 1,
                    orf=self.offset_range_factor, act_dtype=act, pe_mode=self.pe_mode, no_off=bool(self.no_off))
        if act == torch.bfloat16:
            w_bf = [cached_bf16(m) for m in (self.proj_q, self.proj_k, self.proj_v, self.proj_out)]
            if all(w is not None for w in w_bf):
                meta["w_bf16"] = w_bf
        y_l, pos = _BlockFn.apply(x_l, meta, *self._params())
        y = y_l.permute(0, 3, 1, 2)
        if self.training and self.proj_drop_p > 0:     # dat_blocks.py:225: library dropout on the block output (same
            y = F.dropout(y, self.proj_drop_p, True)   # distribution as the reference's mask, not the same draw)
        if not self.return_pos_ref:
            return y, None, None
        hk = H // self.stride if self.no_off else (H + 2 * self.conv_offset[0].padding[0] - self.ksize) // self.stride + 1
        wk = pos.shape[2] // hk
        ry = torch.empty(hk, device=x.device)
        rx = torch.empty(wk, device=x.device)
        _cabi.check(_cabi.lib().dat_ref_points(hk, wk, _ptr(ry), _ptr(rx),
                                               C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                    "dat_ref_points")
        ref = torch.stack(torch.meshgrid(ry, rx, indexing="ij"), -1)[None].expand(B * self.n_groups, -1, -1, -1)
        return y, pos.reshape(B * self.n_groups, hk, wk, 2), ref
