"""Host-side DAT / DAT++ backbone that calls the B200 deformable-attention block.

Row "next" of the scope table: the caller of the hot path
(`models/backbones/dat.py:34-312` in the reference).  Only the deformable-attention
block is hand-written CUDA; everything else in this file is plain PyTorch plumbing
(library convolutions / LayerNorm), kept so that `bench.py` can measure the headline
metric (DAT-T++ images/s fwd+bwd @512²) on a GPU box where the reference tree does not
exist.  Module and parameter names follow the reference, so a reference DAT++ state
dict loads with `strict=True`.
"""
from typing import Callable, Sequence

import os

import torch
import torch.nn as nn

from .conv import Conv3x3s2CL, GeluCL, to_nchw_contiguous
from .dattention import DAttentionBaseline, _pair
from .dwconv import DepthwiseConvCL, MODE_PLAIN, MODE_RESIDUAL, MODE_RESIDUAL_GELU
from .layernorm import LayerNormProxy, TorchLayerNormProxy
from .pointwise import PointwiseConvCL
from .residual import scale_residual
from .weights import Bf16WeightCache

__all__ = ["DAT", "TransformerStage", "DAT_TINY_PP", "DAT_SMALL_PP", "DAT_BASE_PP", "build_dat", "LayerNormProxy",
           "TorchLayerNormProxy"]

# DAT-T++ backbone hyper-parameters (configs/dat/upn_tiny_160k_dp03_lr6.py:9-32)
DAT_TINY_PP = dict(
    dim_stem=64, dims=[64, 128, 256, 512], depths=[2, 4, 18, 2],
    stage_spec=[["X", "D"], ["X", "D"] * 2, ["X", "D"] * 9, ["D", "D"]],
    heads=[2, 4, 8, 16], groups=[1, 2, 4, 8], use_pes=[True] * 4, strides=[8, 4, 2, 1],
    offset_range_factor=[-1, -1, -1, -1], use_dwc_mlps=[True] * 4, use_lpus=[True] * 4,
    use_conv_patches=True, ksizes=[9, 7, 5, 3], nat_ksizes=[7, 7, 7, 7], drop_path_rate=0.3,
    use_checkpoint=False)


# DAT-S++ / DAT-B++ (BASELINE.json configs[2], configs[3]).  Their config files are NOT in the reference tree
# (only the tiny one is): these are the upstream DAT++ hyper-parameters, "assumed, unpinned by the reference"
#  - same structure as the tiny config, wider / deeper.
DAT_SMALL_PP = dict(DAT_TINY_PP, dim_stem=96, dims=[96, 192, 384, 768], depths=[2, 4, 18, 2],
                    heads=[3, 6, 12, 24], groups=[1, 2, 3, 6], drop_path_rate=0.4)
DAT_BASE_PP = dict(DAT_TINY_PP, dim_stem=128, dims=[128, 256, 512, 1024], depths=[2, 4, 18, 2],
                   heads=[4, 8, 16, 32], groups=[2, 4, 8, 16], drop_path_rate=0.6)


class DropPath(nn.Module):
    """Stochastic depth, per sample, kept-path rescaled by 1/(1-p) (timm semantics)."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = 1.0 - self.p
        mask = torch.empty(x.shape[0], 1, 1, 1, device=x.device, dtype=x.dtype).bernoulli_(keep)
        return x * mask / keep


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))

    def forward(self, x):
        return x * self.gamma.view(1, -1, 1, 1)


class TransformerMLP(nn.Module):
    """Linear-GELU-Linear over channels (dat_blocks.py:244-265)."""

    def __init__(self, channels, expansion, drop):
        super().__init__()
        self.chunk = nn.Sequential()
        self.chunk.add_module("linear1", nn.Linear(channels, channels * expansion))
        self.chunk.add_module("act", nn.GELU())
        self.chunk.add_module("drop1", nn.Dropout(drop))
        self.chunk.add_module("linear2", nn.Linear(channels * expansion, channels))
        self.chunk.add_module("drop2", nn.Dropout(drop))

    def forward(self, x):
        return self.chunk(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class TransformerMLPWithConv(nn.Module):
    """1x1 -> (+ depthwise 3x3) -> GELU -> 1x1 (dat_blocks.py:316-348)."""

    def __init__(self, channels, expansion, drop, b200_ops=False):
        super().__init__()
        hidden = channels * expansion
        self.b200_ops = b200_ops
        pw = PointwiseConvCL if b200_ops else (lambda cin, cout: nn.Conv2d(cin, cout, 1))
        self.linear1 = nn.Sequential(pw(channels, hidden))
        self.drop1 = nn.Dropout(drop)
        self.act = nn.GELU()
        self.linear2 = nn.Sequential(pw(hidden, channels))
        self.drop2 = nn.Dropout(drop)
        # b200_ops: gelu(x + dwconv(x) + b) is one fused channel-last kernel
        self.dwc = (DepthwiseConvCL(hidden, 3, MODE_RESIDUAL_GELU) if b200_ops
                    else nn.Conv2d(hidden, hidden, 3, 1, 1, groups=hidden))

    def forward(self, x):
        x = self.drop1(self.linear1(x))
        x = self.dwc(x) if self.b200_ops else self.act(x + self.dwc(x))
        return self.drop2(self.linear2(x))

    def forward_residual(self, x, resid, scale):
        """resid + scale[b] * mlp(x): with the dat_b200 ops the residual add with stochastic depth rides in the epilogue
        of the last 1x1 conv (one kernel less and no bf16 round trip of the branch); else the separate add."""
        h = self.drop1(self.linear1(x))
        h = self.dwc(h) if self.b200_ops else self.act(h + self.dwc(h))
        fc2 = self.linear2[0]
        if (self.b200_ops and not (self.training and self.drop2.p > 0) and hasattr(fc2, "residual_fusable")
                and fc2.residual_fusable(h, resid)):
            return fc2.forward_residual(h, resid, scale)
        return scale_residual(self.drop2(self.linear2(h)), resid, scale)


class TransformerStage(nn.Module):
    """One resolution stage (dat.py:34-165).  Spec letters: 'D' deformable attention,
    'X' depthwise conv mixer.  `attn_cls` lets tests / the CPU baseline substitute the
    block implementation (reference class, oracle port) without touching the wiring."""

    def __init__(self, fmap_size, window_size, dim_in, dim_embed, depths, stage_spec, n_groups,
                 use_pe, heads, stride, offset_range_factor, dwc_pe, no_off, fixed_pe, attn_drop,
                 proj_drop, expansion, drop, drop_path_rate, use_dwc_mlp, ksize, layer_scale_value,
                 use_lpu, log_cpb, stage_i, use_checkpoint, attn_cls: Callable = DAttentionBaseline,
                 norm_cls: Callable = LayerNormProxy, b200_ops: bool = False):
        super().__init__()
        self.b200_ops = b200_ops
        self.fused_residual = b200_ops and hasattr(norm_cls, "forward_fork")
        fmap_size = _pair(fmap_size)
        self.depths, self.stage_spec = depths, list(stage_spec)
        self.use_lpu, self.use_checkpoint = use_lpu, use_checkpoint
        hc = dim_embed // heads
        assert dim_embed == heads * hc
        self.proj = nn.Conv2d(dim_in, dim_embed, 1) if dim_in != dim_embed else nn.Identity()
        self.ln_cnvnxt = nn.ModuleDict(
            {str(d): norm_cls(dim_embed) for d in range(depths) if stage_spec[d] == "X"})
        self.layer_norms = nn.ModuleList(
            [norm_cls(dim_embed) if stage_spec[d // 2] != "X" else nn.Identity()
             for d in range(2 * depths)])
        mlp_cls = TransformerMLPWithConv if use_dwc_mlp else TransformerMLP
        mlp_kw = dict(b200_ops=b200_ops) if use_dwc_mlp else {}
        self.mlps = nn.ModuleList([mlp_cls(dim_embed, expansion, drop, **mlp_kw) for _ in range(depths)])
        self.attns = nn.ModuleList()
        self.drop_path = nn.ModuleList()
        self.layer_scales = nn.ModuleList(
            [LayerScale(dim_embed, layer_scale_value) if layer_scale_value > 0.0 else nn.Identity()
             for _ in range(2 * depths)])
        self.local_perception_units = nn.ModuleList(
            [(DepthwiseConvCL(dim_embed, 3, MODE_RESIDUAL, keep_input_dtype=True) if b200_ops
              else nn.Conv2d(dim_embed, dim_embed, 3, 1, 1, groups=dim_embed)) if use_lpu else nn.Identity()
             for _ in range(depths)])
        for d in range(depths):
            if stage_spec[d] == "D":
                self.attns.append(attn_cls(fmap_size, fmap_size, heads, hc, n_groups, attn_drop,
                                           proj_drop, stride, offset_range_factor, use_pe, dwc_pe,
                                           no_off, fixed_pe, ksize, log_cpb, stage_i))
            elif stage_spec[d] == "X":
                self.attns.append(DepthwiseConvCL(dim_embed, window_size, MODE_PLAIN) if b200_ops else
                                  nn.Conv2d(dim_embed, dim_embed, window_size, padding=window_size // 2,
                                            groups=dim_embed))
            else:
                raise NotImplementedError(f"Spec: {stage_spec[d]} is not supported.")
            self.drop_path.append(DropPath(drop_path_rate[d]) if drop_path_rate[d] > 0.0 else nn.Identity())

    def _inner_forward_b200(self, x):
        """Same dataflow as `_inner_forward` with the dat_b200 fusions: LayerNorm + residual fork in
        one autograd node, `drop_path(branch) + x` in one kernel.  (layer_scale > 0 is applied by
        the library multiply before the fused add.)"""
        x = self.proj(x)
        # the MLP's first 1x1 conv casts its input to the autocast dtype: let the LayerNorm write it
        mlp_in = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else None
        # stochastic-depth scales of the whole stage in one draw: row i = i-th drop_path call
        # (independent per call and per sample, as in the reference), mask / keep_prob
        B = x.shape[0]
        ps = []
        for d in range(self.depths):
            ps += [getattr(self.drop_path[d], "p", 0.0)] * (1 if self.stage_spec[d] == "X" else 2)
        fixed = getattr(self, "_fixed_scales", None)
        if self.training and fixed is not None:        # tests: pre-drawn masks (fix_drop_path_scales)
            assert fixed.shape == (len(ps), B), (fixed.shape, len(ps), B)
            scales = fixed
        elif self.training and any(q > 0.0 for q in ps):
            keep = self._keep_probs(ps, x.device)
            scales = (torch.rand(len(ps), B, device=x.device) < keep).float() / keep
        else:
            scales = torch.ones(len(ps), B, device=x.device)
        si = 0
        for d in range(self.depths):
            p = getattr(self.drop_path[d], "p", 0.0)
            if self.use_lpu:
                x = self.local_perception_units[d](x)
            if self.stage_spec[d] == "X":   # note: no residual around mixer+MLP (dat.py:140-144)
                x = self.attns[d](self.layer_norms[2 * d](x))
                m = self.mlps[d](self.ln_cnvnxt[str(d)](x, out_dtype=mlp_in))
                x = scale_residual(m, None, scales[si]) if p > 0.0 and self.training else m
                si += 1
            else:
                x, ln = self.layer_norms[2 * d].forward_fork(x)
                a, _, _ = self.attns[d](ln)
                # `x = drop_path(attn) + x` and the norm that follows it in one kernel (forward and backward)
                x, ln = self.layer_norms[2 * d + 1].forward_residual_fork(self.layer_scales[2 * d](a), x, scales[si],
                                                                          out_dtype=mlp_in)
                if isinstance(self.layer_scales[2 * d + 1], nn.Identity) and hasattr(self.mlps[d], "forward_residual"):
                    x = self.mlps[d].forward_residual(ln, x, scales[si + 1])
                else:
                    m = self.mlps[d](ln)
                    x = scale_residual(self.layer_scales[2 * d + 1](m), x, scales[si + 1])
                si += 2
        return x

    def drop_path_calls(self):
        """Drop probability of every drop_path call of one forward, in call order ('X': one call, 'D': two)."""
        ps = []
        for d in range(self.depths):
            ps += [getattr(self.drop_path[d], "p", 0.0)] * (1 if self.stage_spec[d] == "X" else 2)
        return ps

    def fix_drop_path_scales(self, batch, generator=None, device=None):
        """Testing hook: draw the stochastic-depth scales (mask / keep_prob, one row per drop_path call) once and
        re-use them in every training forward, so that two runs (eager / graph replay / another implementation)
        see the same masks.  `batch=None` switches back to a fresh draw per forward.  Returns the (calls, batch)
        tensor."""
        if batch is None:
            self._fixed_scales = None
            return None
        ps = self.drop_path_calls()
        keep = torch.tensor([1.0 - q for q in ps]).view(-1, 1)
        scales = (torch.rand(len(ps), batch, generator=generator) < keep).float() / keep
        self._fixed_scales = scales.to(device) if device is not None else scales
        return self._fixed_scales

    def _keep_probs(self, ps, device):
        key = (tuple(ps), str(device))
        if getattr(self, "_keep_cache", (None, None))[0] != key:
            self._keep_cache = (key, torch.tensor([1.0 - q for q in ps], device=device).view(-1, 1))
        return self._keep_cache[1]

    def _inner_forward(self, x):
        if self.fused_residual and x.is_cuda:
            return self._inner_forward_b200(x)
        x = self.proj(x)
        for d in range(self.depths):
            if self.use_lpu:   # b200: conv + bias + residual in one channel-last kernel
                x = (self.local_perception_units[d](x) if self.b200_ops
                     else self.local_perception_units[d](x.contiguous()) + x)
            if self.stage_spec[d] == "X":   # note: no residual around mixer+MLP (dat.py:140-144)
                x = self.attns[d](self.layer_norms[2 * d](x))
                x = self.drop_path[d](self.mlps[d](self.ln_cnvnxt[str(d)](x)))
            else:
                a, _, _ = self.attns[d](self.layer_norms[2 * d](x))
                x = self.drop_path[d](self.layer_scales[2 * d](a)) + x
                m = self.mlps[d](self.layer_norms[2 * d + 1](x))
                x = self.drop_path[d](self.layer_scales[2 * d + 1](m)) + x
        return x

    def forward(self, x):
        if self.training and x.requires_grad and self.use_checkpoint:
            return torch.utils.checkpoint.checkpoint(self._inner_forward, x, use_reentrant=False)
        return self._inner_forward(x)


class DAT(nn.Module):
    """DAT / DAT++ backbone (dat.py:167-312): conv stem, 4 stages, 3 down-projections,
    per-stage output norms; returns the 4 feature maps."""

    def __init__(self, img_size=224, patch_size=4, num_classes=1000, expansion=4, dim_stem=96,
                 dims=(96, 192, 384, 768), depths=(2, 2, 6, 2), heads=(3, 6, 12, 24),
                 heads_q=(6, 12, 24, 48), window_sizes=(7, 7, 7, 7), drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, strides=(-1, -1, -1, -1),
                 offset_range_factor=(1, 2, 3, 4), local_orf=(-1,) * 4, local_kv_sizes=(-1,) * 4,
                 offset_pes=(False,) * 4,
                 stage_spec=(("L", "D"), ("L", "D"), ("L", "D") * 3, ("L", "D")),
                 groups=(-1, -1, 3, 6), use_pes=(False,) * 4, dwc_pes=(False,) * 4,
                 sr_ratios=(8, 4, 2, 1), lower_lr_kvs=None, fixed_pes=(False,) * 4,
                 no_offs=(False,) * 4, ns_per_pts=(4,) * 4, use_dwc_mlps=(False,) * 4,
                 use_conv_patches=False, ksizes=(9, 7, 5, 3), ksize_qnas=(3,) * 4, nqs=(2,) * 4,
                 qna_activation="exp", deform_groups=(0,) * 4, nat_ksizes=(3,) * 4,
                 layer_scale_values=(-1,) * 4, use_lpus=(False,) * 4, use_cmt_mlps=(False,) * 4,
                 log_cpb=(False,) * 4, out_indices=(0, 1, 2, 3), use_checkpoint=True,
                 init_cfg=None, attn_cls: Callable = DAttentionBaseline,
                 norm_cls: Callable = LayerNormProxy, b200_ops: bool = False, **kwargs):
        super().__init__()
        if any(use_cmt_mlps):
            raise NotImplementedError("use_cmt_mlps (BatchNorm MLP variant) is not implemented")
        self.out_indices = out_indices
        half = dim_stem // 2
        # b200_ops: the stride-2 3x3 convolutions (stem, down-projections) run as im2col + tcgen05 GEMMs (conv.py)
        own_convs = b200_ops and use_conv_patches and patch_size == 4
        conv3 = (lambda cin, cout, bias=True: Conv3x3s2CL(cin, cout, bias=bias)) if own_convs else \
                (lambda cin, cout, bias=True: nn.Conv2d(cin, cout, 3, 2, 1, bias=bias))
        if use_conv_patches:
            self.patch_proj = nn.Sequential(
                conv3(3, half) if patch_size == 4 else nn.Conv2d(3, half, 3, patch_size // 2, 1), norm_cls(half),
                GeluCL() if own_convs else nn.GELU(),
                conv3(half, dim_stem) if patch_size == 4 else nn.Conv2d(half, dim_stem, 3, patch_size // 2, 1),
                norm_cls(dim_stem))
        else:
            self.patch_proj = nn.Sequential(nn.Conv2d(3, dim_stem, patch_size, patch_size, 0),
                                            norm_cls(dim_stem))
        fmap = img_size // patch_size
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.stages = nn.ModuleList()
        self.norms = nn.ModuleList()
        for i in range(4):
            dim_in = dim_stem if i == 0 else dims[i - 1] * 2
            lo, hi = sum(depths[:i]), sum(depths[:i + 1])
            self.stages.append(TransformerStage(
                fmap, window_sizes[i], dim_in, dims[i], depths[i], stage_spec[i], groups[i],
                use_pes[i], heads[i], strides[i], offset_range_factor[i], dwc_pes[i], no_offs[i],
                fixed_pes[i], attn_drop_rate, drop_rate, expansion, drop_rate, dpr[lo:hi],
                use_dwc_mlps[i], ksizes[i], layer_scale_values[i], use_lpus[i], log_cpb[i], i,
                use_checkpoint, attn_cls=attn_cls, norm_cls=norm_cls, b200_ops=b200_ops))
            self.norms.append(norm_cls(dims[i]) if i in out_indices else nn.Identity())
            fmap //= 2
        self.down_projs = nn.ModuleList()
        for i in range(3):
            conv = (conv3(dims[i], dims[i + 1], bias=False) if use_conv_patches
                    else nn.Conv2d(dims[i], dims[i + 1], 2, 2, 0, bias=False))
            self.down_projs.append(nn.Sequential(conv, norm_cls(dims[i + 1])))
        if b200_ops and os.environ.get("DAT_B200_NCHW_CONVS") is None:
            # The dat_b200 kernels are channel-last; keep the (library) stem / down-projection convolutions in
            # channels_last too, so their outputs feed the LayerNorm kernels without NCHW <-> NHWC copies.
            # Shapes and state-dict contents are unchanged (only the parameters' strides differ).
            for m in list(self.patch_proj) + [dp[0] for dp in self.down_projs]:
                if isinstance(m, nn.Conv2d) and not isinstance(m, Conv3x3s2CL):
                    m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)

        # bf16 operand copies of every 1x1-conv weight, cast once per forward in one launch (weights.py)
        self._bf16_weights = None
        if b200_ops:
            ents = []
            for m in self.modules():
                if isinstance(m, PointwiseConvCL):
                    ents.append((m, "weight"))
                elif isinstance(m, DAttentionBaseline):
                    ents += [(c, "weight") for c in (m.proj_q, m.proj_k, m.proj_v, m.proj_out)]
            self._bf16_weights = Bf16WeightCache(ents)

    def forward(self, x):
        if (self._bf16_weights is not None and x.is_cuda and torch.is_autocast_enabled("cuda")
                and torch.get_autocast_dtype("cuda") == torch.bfloat16 and not os.environ.get("DAT_B200_NO_WEIGHT_CACHE")):
            self._bf16_weights.refresh()
        x = self.patch_proj(x)
        outs = []
        for i in range(4):
            x = self.stages[i](x)
            outs.append(to_nchw_contiguous(self.norms[i](x)))
            if i < 3:
                x = self.down_projs[i](x)
        return outs


def build_dat(cfg: dict = None, attn_cls: Callable = DAttentionBaseline, norm_cls: Callable = None,
              b200_ops: bool = None, **override) -> DAT:
    """DAT(**cfg) as `models/builder.py:93-102` does (init_cfg / type keys dropped).
    `norm_cls` defaults to the dat_b200 LayerNorm kernels when the block is the dat_b200 block, and to
    the library LayerNorm when another block implementation (reference / oracle, CPU) is plugged in."""
    kw = dict(DAT_TINY_PP if cfg is None else cfg)
    kw.update(override)
    kw.pop("type", None)
    kw.pop("init_cfg", None)
    if norm_cls is None:
        norm_cls = LayerNormProxy if attn_cls is DAttentionBaseline else TorchLayerNormProxy
    if b200_ops is None:   # fused channel-last depthwise convs (LPU, MLP middle, 'X' mixer)
        b200_ops = attn_cls is DAttentionBaseline
    return DAT(attn_cls=attn_cls, norm_cls=norm_cls, b200_ops=b200_ops, **kw)
