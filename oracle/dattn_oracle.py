"""TEST INFRASTRUCTURE — CPU oracle for the DAT deformable-attention block.

This file is a *restatement* of the reference algorithm
(`/root/reference/models/utils/dat_blocks.py:138-227`, class `DAttentionBaseline`)
used only as the checker by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.  Nothing under
`dat_segmentation_b200/` imports it; the product path is CUDA only.

Parity status: PINNED against outputs of the unmodified reference run in the
build container (`tests/golden/make_golden.py` → `tests/golden/*.pt`,
`tests/test_oracle_golden.py`).  The reference itself ships no tests or golden
vectors, and its arithmetic lives in PyTorch/ATen
(torch 2.11.0 here), so the pin is "reference module on CPU fp32, this torch".

Two forms are provided:

* `forward_explicit` / `backward_explicit` — channel-last, every step spelled
  out (tap indices, bilinear weights, softmax, analytic gradients).  This is the
  executable spec of each CUDA kernel boundary.
* `forward_libops` — the same block expressed with the library operators the
  reference dispatches to (conv2d / layer_norm / grid_sample / einsum), used as
  the CPU timing baseline (`cpu_baseline.kind == "port"`).

Layout conventions (explicit form): activations are channel-last,
`x: (B, H, W, C)`; `pos: (B, G, Hk, Wk, 2)` holds (y, x) in [-1, 1] grid units.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

PARAM_KEYS = (
    "conv_offset.0.weight", "conv_offset.0.bias",
    "conv_offset.1.norm.weight", "conv_offset.1.norm.bias",
    "conv_offset.3.weight",
    "proj_q.weight", "proj_q.bias", "proj_k.weight", "proj_k.bias",
    "proj_v.weight", "proj_v.bias", "proj_out.weight", "proj_out.bias",
    "rpe_table",
)


@dataclass(frozen=True)
class BlockCfg:
    """Hyper-parameters of one block (dat_blocks.py:21-50)."""
    q_h: int
    q_w: int
    n_heads: int
    n_head_channels: int
    n_groups: int
    stride: int
    ksize: int
    offset_range_factor: float
    # variant branches (dat_blocks.py:57-59,84-99,156-157,164-167,185-197,221-222); defaults = shipped configs
    use_pe: bool = True
    dwc_pe: bool = False
    no_off: bool = False
    fixed_pe: bool = False
    log_cpb: bool = False

    @property
    def pe_mode(self):
        """'none' | 'dwc' | 'fixed' | 'log_cpb' | 'rpe' in the reference's precedence order (:84-104, :183-214)."""
        if not self.use_pe or self.no_off:
            return "none"
        return "dwc" if self.dwc_pe else "fixed" if self.fixed_pe else "log_cpb" if self.log_cpb else "rpe"

    @property
    def nc(self):
        return self.n_heads * self.n_head_channels

    @property
    def cg(self):
        return self.nc // self.n_groups

    @property
    def hg(self):
        return self.n_heads // self.n_groups

    @property
    def pad(self):  # dat_blocks.py:50
        return self.ksize // 2 if self.ksize != self.stride else 0

    @property
    def table_hw(self):  # dat_blocks.py:101-103
        return 2 * self.q_h - 1, 2 * self.q_w - 1

    def sample_grid(self, H, W):
        hk = (H + 2 * self.pad - self.ksize) // self.stride + 1
        wk = (W + 2 * self.pad - self.ksize) // self.stride + 1
        return hk, wk


def init_params(cfg: BlockCfg, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random parameters with the reference's shapes (dat_blocks.py:52-104).
    Values are N(0, σ) with fan-in scaling — not the reference's init, which does
    not matter for parity tests (weights are always passed explicitly)."""
    g = torch.Generator().manual_seed(seed)
    C, Cg, k = cfg.nc, cfg.cg, cfg.ksize
    th, tw = cfg.table_hw

    def rn(*shape, std):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * std).to(dtype)

    p = {
        "conv_offset.0.weight": rn(Cg, 1, k, k, std=1.0 / k),
        "conv_offset.0.bias": rn(Cg, std=0.1),
        "conv_offset.1.norm.weight": 1.0 + rn(Cg, std=0.1),
        "conv_offset.1.norm.bias": rn(Cg, std=0.1),
        "conv_offset.3.weight": rn(2, Cg, 1, 1, std=0.5 / math.sqrt(Cg)),
        "rpe_table": rn(cfg.n_heads, th, tw, std=0.5),
    }
    for name in ("proj_q", "proj_k", "proj_v", "proj_out"):
        p[f"{name}.weight"] = rn(C, C, 1, 1, std=1.0 / math.sqrt(C))
        p[f"{name}.bias"] = rn(C, std=0.1)
    return p


# ----------------------------------------------------------------------------
# grids (dat_blocks.py:108-136)
# ----------------------------------------------------------------------------

def ref_points(hk: int, wk: int, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """Reference sample centres: ((i + 0.5) / (Hk - 1)) * 2 - 1, each op rounded
    on its own (dat_blocks.py:111-118: linspace(0.5, Hk-0.5, Hk) then
    .div_().mul_().sub_()).  linspace(0.5, Hk-0.5, Hk)[i] == i + 0.5 exactly for
    the sizes in scope (checked in tests)."""
    ry = (torch.arange(hk, dtype=dtype) + 0.5).div(hk - 1.0).mul(2.0).sub(1.0)
    rx = (torch.arange(wk, dtype=dtype) + 0.5).div(wk - 1.0).mul(2.0).sub(1.0)
    return ry, rx


def query_grid(H: int, W: int, dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """(r / (H - 1)) * 2 - 1 (dat_blocks.py:126-133)."""
    qy = torch.arange(H, dtype=dtype).div(H - 1.0).mul(2.0).sub(1.0)
    qx = torch.arange(W, dtype=dtype).div(W - 1.0).mul(2.0).sub(1.0)
    return qy, qx


# ----------------------------------------------------------------------------
# bilinear sampling, align_corners=True, zeros padding
# (torch/include/ATen/native/GridSampler.h:27-36, 205-235;
#  call sites dat_blocks.py:169-172 and :206-210)
# ----------------------------------------------------------------------------

def bilinear_taps(gx: Tensor, gy: Tensor, W: int, H: int):
    """Tap indices and weights for grid coords (gx, gy) in [-1, 1] units.
    Returns x0, y0 (int64, the north-west tap), weights (nw, ne, sw, se) and
    validity masks for the four taps.  Formula order follows ATen:
    ix = ((gx + 1) / 2) * (W - 1); x0 = floor(ix)."""
    ix = ((gx + 1.0) / 2.0) * (W - 1)
    iy = ((gy + 1.0) / 2.0) * (H - 1)
    x0f, y0f = torch.floor(ix), torch.floor(iy)
    x1f, y1f = x0f + 1.0, y0f + 1.0
    wx1, wx0 = ix - x0f, x1f - ix
    wy1, wy0 = iy - y0f, y1f - iy
    x0, y0 = x0f.long(), y0f.long()
    vx0 = (x0 >= 0) & (x0 <= W - 1)
    vx1 = (x0 + 1 >= 0) & (x0 + 1 <= W - 1)
    vy0 = (y0 >= 0) & (y0 <= H - 1)
    vy1 = (y0 + 1 >= 0) & (y0 + 1 <= H - 1)
    weights = (wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1)
    valid = (vx0 & vy0, vx1 & vy0, vx0 & vy1, vx1 & vy1)
    return x0, y0, weights, valid, (wx0, wx1, wy0, wy1)


_TAP_DXY = ((0, 0), (1, 0), (0, 1), (1, 1))  # (dx, dy) for nw, ne, sw, se


# ----------------------------------------------------------------------------
# explicit forward
# ----------------------------------------------------------------------------

def _w2d(p, name):
    w = p[name]
    return w.reshape(w.shape[0], w.shape[1])


def offset_net_explicit(q: Tensor, p: Dict[str, Tensor], cfg: BlockCfg):
    """Offset network on q (B,H,W,C) → raw offsets (B,G,Hk,Wk,2) plus saved
    intermediates (dat_blocks.py:51-56, :144-145).  Same weights for all groups."""
    B, H, W, C = q.shape
    G, Cg, k, s, pd = cfg.n_groups, cfg.cg, cfg.ksize, cfg.stride, cfg.pad
    hk, wk = cfg.sample_grid(H, W)
    qg = q.reshape(B, H, W, G, Cg)
    qp = F.pad(qg, (0, 0, 0, 0, pd, pd, pd, pd))  # pad W and H
    w0 = p["conv_offset.0.weight"].reshape(Cg, k, k)
    t = p["conv_offset.0.bias"].reshape(1, 1, 1, 1, Cg).expand(B, hk, wk, G, Cg).clone()
    for u in range(k):
        for v in range(k):
            win = qp[:, u:u + s * (hk - 1) + 1:s, v:v + s * (wk - 1) + 1:s]
            t = t + win * w0[:, u, v]
    mu = t.mean(-1, keepdim=True)
    var = ((t - mu) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    that = (t - mu) * rstd
    ln = that * p["conv_offset.1.norm.weight"] + p["conv_offset.1.norm.bias"]
    act = 0.5 * ln * (1.0 + torch.erf(ln / math.sqrt(2.0)))
    w3 = p["conv_offset.3.weight"].reshape(2, Cg)
    off = torch.einsum("bijgc,pc->bgijp", act, w3)
    saved = dict(t=t, that=that, rstd=rstd, ln=ln, act=act)
    return off, saved


def positions_explicit(off_raw: Tensor, cfg: BlockCfg):
    """tanh·range·orf (orf >= 0) or clamp (orf < 0); + reference points
    (dat_blocks.py:149-162).  off_raw (B,G,Hk,Wk,2) → pos, pre-clamp value."""
    hk, wk = off_raw.shape[2], off_raw.shape[3]
    ry, rx = ref_points(hk, wk, off_raw.dtype)
    ref = torch.stack(torch.meshgrid(ry, rx, indexing="ij"), -1)
    orf = cfg.offset_range_factor
    if orf >= 0:
        rng = torch.tensor([1.0 / (hk - 1.0), 1.0 / (wk - 1.0)], dtype=torch.float32)
        off = torch.tanh(off_raw).mul(rng.to(off_raw.dtype)).mul(orf)
        pre = off + ref
        pos = pre
    else:
        pre = off_raw + ref
        pos = pre.clamp(-1.0, 1.0)
    return pos, pre, ref


def sample_features_explicit(x: Tensor, pos: Tensor, cfg: BlockCfg):
    """x (B,H,W,C), pos (B,G,Hk,Wk,2)(y,x) → xs (B,Ns,C) and the integer taps
    (dat_blocks.py:169-172)."""
    B, H, W, C = x.shape
    G, Cg = cfg.n_groups, cfg.cg
    Ns = pos.shape[2] * pos.shape[3]
    py, px = pos[..., 0].reshape(B, G, Ns), pos[..., 1].reshape(B, G, Ns)
    x0, y0, wts, valid, _ = bilinear_taps(px, py, W, H)
    xg = x.reshape(B, H * W, G, Cg).permute(0, 2, 1, 3)  # B,G,HW,Cg
    xs = torch.zeros(B, G, Ns, Cg, dtype=x.dtype)
    for (dx, dy), w, ok in zip(_TAP_DXY, wts, valid):
        idx = ((y0 + dy).clamp(0, H - 1) * W + (x0 + dx).clamp(0, W - 1))
        tap = torch.gather(xg, 2, idx[..., None].expand(B, G, Ns, Cg))
        xs = xs + tap * (w * ok)[..., None]
    xs = xs.permute(0, 2, 1, 3).reshape(B, Ns, C)
    return xs, (x0, y0)


def rpe_bias_explicit(pos: Tensor, table: Tensor, H: int, W: int, cfg: BlockCfg):
    """Interpolated relative-position bias (dat_blocks.py:198-212).
    pos (B,G,Hk,Wk,2), table (h,Th,Tw) → bias (B,h,HW,Ns)."""
    B, G = pos.shape[0], pos.shape[1]
    Ns = pos.shape[2] * pos.shape[3]
    h, hg = cfg.n_heads, cfg.hg
    th, tw = table.shape[1], table.shape[2]
    qy, qx = query_grid(H, W, pos.dtype)
    gy = torch.repeat_interleave(qy, W)  # (HW,) row-major m = r*W + c
    gx = qx.repeat(H)
    py, px = pos[..., 0].reshape(B, G, 1, Ns), pos[..., 1].reshape(B, G, 1, Ns)
    dy = (gy.reshape(1, 1, -1, 1) - py) * 0.5
    dx = (gx.reshape(1, 1, -1, 1) - px) * 0.5
    x0, y0, wts, valid, _ = bilinear_taps(dx, dy, tw, th)  # (B,G,HW,Ns)
    tab = table.reshape(G, hg, th * tw)
    bias = torch.zeros(B, G, hg, H * W, Ns, dtype=pos.dtype)
    for (ddx, ddy), w, ok in zip(_TAP_DXY, wts, valid):
        idx = ((y0 + ddy).clamp(0, th - 1) * tw + (x0 + ddx).clamp(0, tw - 1))  # B,G,HW,Ns
        idx_e = idx[:, :, None].expand(B, G, hg, H * W, Ns).reshape(B, G, hg, -1)
        val = torch.gather(tab[None].expand(B, G, hg, th * tw), 3, idx_e).reshape(B, G, hg, H * W, Ns)
        bias = bias + val * (w * ok)[:, :, None]
    return bias.reshape(B, h, H * W, Ns)


def rpe_table_grad_separable(ds: Tensor, pos: Tensor, H: int, W: int, cfg: BlockCfg) -> Tensor:
    """d rpe_table from dS = dL/d(score) in the separable form the CUDA kernel `rpe_table_grad_kernel`
    (csrc/rpe_table_grad.cu) evaluates with tensor-core GEMMs: for every sample n

        dT_n = A_n^T . dS_n . B_n,   A_n[r][y] = hat(iy(r, n) - y),  B_n[c][x] = hat(ix(c, n) - x),
        hat(d) = max(0, 1 - |d|),   ix = ((d_x + 1) / 2)(Tw - 1),  d_x = (q_grid_x[c] - pos_x[n]) / 2

    (the bilinear weights of `F.grid_sample(..., align_corners=True)` with zero padding are exactly these hat
    products restricted to the table; dat_blocks.py:198-212).  ds (B,h,HW,Ns), pos (B,G,Hk,Wk,2) -> (h,Th,Tw)."""
    B, G = pos.shape[0], pos.shape[1]
    Ns = pos.shape[2] * pos.shape[3]
    h, hg = cfg.n_heads, cfg.hg
    th, tw = cfg.table_hw
    qy, qx = query_grid(H, W, pos.dtype)
    py, px = pos[..., 0].reshape(B, G, Ns), pos[..., 1].reshape(B, G, Ns)
    iy = ((qy.reshape(1, 1, H, 1) - py.reshape(B, G, 1, Ns)) * 0.5 + 1) / 2 * (th - 1)     # (B,G,H,Ns)
    ix = ((qx.reshape(1, 1, W, 1) - px.reshape(B, G, 1, Ns)) * 0.5 + 1) / 2 * (tw - 1)     # (B,G,W,Ns)
    ys = torch.arange(th, dtype=pos.dtype).reshape(1, 1, 1, 1, th)
    xs = torch.arange(tw, dtype=pos.dtype).reshape(1, 1, 1, 1, tw)
    A = (1 - (iy[..., None] - ys).abs()).clamp_min(0)          # (B,G,H,Ns,Th)
    Bm = (1 - (ix[..., None] - xs).abs()).clamp_min(0)         # (B,G,W,Ns,Tw)
    dsr = ds.reshape(B, G, hg, H, W, Ns)
    # dT[g,j,y,x] = sum_{b,n,r,c} A[b,g,r,n,y] dS[b,g,j,r,c,n] B[b,g,c,n,x]
    e = torch.einsum("bgjrcn,bgcnx->bgjrnx", dsr, Bm)
    dt = torch.einsum("bgrny,bgjrnx->gjyx", A, e)
    return dt.reshape(h, th, tw)


def merge_sample_chunks(o_parts, lse_parts):
    """Merge of per-chunk attention results (the split-KV path of csrc/attention_tc.cu for Ns > 256):
    chunk c holds O_c = softmax(S_c) V_c and lse_c = logsumexp(S_c) over its own samples; then
    O = sum_c w_c O_c / sum_c w_c with w_c = exp(lse_c - max_c lse_c), lse = max + log sum_c w_c.
    o_parts: list of (..., HW, d); lse_parts: list of (..., HW)."""
    lse = torch.stack(lse_parts, 0)
    mx = lse.max(0).values
    w = (lse - mx).exp()
    den = w.sum(0)
    o = sum(wc[..., None] * oc for wc, oc in zip(w, o_parts)) / den[..., None]
    return o, mx + den.log()


def forward_explicit(x: Tensor, p: Dict[str, Tensor], cfg: BlockCfg,
                     pos_override: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Whole block, channel-last.  x (B,H,W,C) → dict with y (B,H,W,C) and every
    intermediate (dat_blocks.py:138-227)."""
    B, H, W, C = x.shape
    h, hc = cfg.n_heads, cfg.n_head_channels
    q = x @ _w2d(p, "proj_q.weight").T + p["proj_q.bias"]
    off_raw, saved = offset_net_explicit(q, p, cfg)
    pos, pre, ref = positions_explicit(off_raw, cfg)
    if pos_override is not None:
        pos = pos_override
    xs, taps = sample_features_explicit(x, pos, cfg)
    Ns = xs.shape[1]
    k = xs @ _w2d(p, "proj_k.weight").T + p["proj_k.bias"]
    v = xs @ _w2d(p, "proj_v.weight").T + p["proj_v.bias"]
    qh = q.reshape(B, H * W, h, hc).permute(0, 2, 1, 3)
    kh = k.reshape(B, Ns, h, hc).permute(0, 2, 1, 3)
    vh = v.reshape(B, Ns, h, hc).permute(0, 2, 1, 3)
    scale = hc ** -0.5
    s = (qh @ kh.transpose(-1, -2)) * scale
    bias = rpe_bias_explicit(pos, p["rpe_table"], H, W, cfg)
    s = s + bias
    m = s.max(-1, keepdim=True).values
    e = torch.exp(s - m)
    l = e.sum(-1, keepdim=True)
    prob = e / l
    lse = (m + torch.log(l)).squeeze(-1)
    o = (prob @ vh).permute(0, 2, 1, 3).reshape(B, H, W, C)
    y = o @ _w2d(p, "proj_out.weight").T + p["proj_out.bias"]
    out = dict(q=q, off_raw=off_raw, pos=pos, pre=pre, ref=ref, xs=xs, k=k, v=v,
               bias=bias, prob=prob, lse=lse, o=o, y=y, tap_x0=taps[0], tap_y0=taps[1])
    out.update({f"off_{k_}": v_ for k_, v_ in saved.items()})
    return out


# ----------------------------------------------------------------------------
# explicit (analytic) backward — the spec of the CUDA backward kernels
# ----------------------------------------------------------------------------

def backward_explicit(x: Tensor, p: Dict[str, Tensor], cfg: BlockCfg, dy: Tensor,
                      fw: Optional[Dict[str, Tensor]] = None):
    """Gradients of sum(y * dy) w.r.t. x and all 14 parameters, written out by
    hand (no autograd).  Returns (dx (B,H,W,C), {param_key: grad}, extras)."""
    if fw is None:
        fw = forward_explicit(x, p, cfg)
    B, H, W, C = x.shape
    h, hc, G, Cg, hg = cfg.n_heads, cfg.n_head_channels, cfg.n_groups, cfg.cg, cfg.hg
    k_, s_, pd = cfg.ksize, cfg.stride, cfg.pad
    hk, wk = cfg.sample_grid(H, W)
    Ns, HW = hk * wk, H * W
    scale = hc ** -0.5
    g: Dict[str, Tensor] = {}
    X2 = x.reshape(B * HW, C)

    # proj_out
    dy2 = dy.reshape(B * HW, C)
    o2 = fw["o"].reshape(B * HW, C)
    g["proj_out.weight"] = (dy2.T @ o2).reshape(C, C, 1, 1)
    g["proj_out.bias"] = dy2.sum(0)
    do = (dy2 @ _w2d(p, "proj_out.weight")).reshape(B, HW, h, hc).permute(0, 2, 1, 3)

    # attention core
    qh = fw["q"].reshape(B, HW, h, hc).permute(0, 2, 1, 3)
    kh = fw["k"].reshape(B, Ns, h, hc).permute(0, 2, 1, 3)
    vh = fw["v"].reshape(B, Ns, h, hc).permute(0, 2, 1, 3)
    oh = fw["o"].reshape(B, HW, h, hc).permute(0, 2, 1, 3)
    prob = fw["prob"]
    delta = (do * oh).sum(-1, keepdim=True)
    dv = prob.transpose(-1, -2) @ do
    dp = do @ vh.transpose(-1, -2)
    ds = prob * (dp - delta)  # (B,h,HW,Ns) == d bias
    dq = (ds @ kh) * scale
    dk = (ds.transpose(-1, -2) @ qh) * scale
    dq = dq.permute(0, 2, 1, 3).reshape(B, H, W, C)
    dk = dk.permute(0, 2, 1, 3).reshape(B * Ns, C)
    dv = dv.permute(0, 2, 1, 3).reshape(B * Ns, C)

    # rpe bias backward: d table (scatter) and d pos (through the displacement)
    pos = fw["pos"]
    table = p["rpe_table"]
    th, tw = table.shape[1], table.shape[2]
    qy, qx = query_grid(H, W, x.dtype)
    gy = torch.repeat_interleave(qy, W).reshape(1, 1, HW, 1)
    gx = qx.repeat(H).reshape(1, 1, HW, 1)
    py, px = pos[..., 0].reshape(B, G, 1, Ns), pos[..., 1].reshape(B, G, 1, Ns)
    ddy, ddx = (gy - py) * 0.5, (gx - px) * 0.5
    x0, y0, wts, valid, (wx0, wx1, wy0, wy1) = bilinear_taps(ddx, ddy, tw, th)
    dsg = ds.reshape(B, G, hg, HW, Ns)
    dtab = torch.zeros(G, hg, th * tw, dtype=x.dtype)
    tab = table.reshape(G, hg, th * tw)
    tapval = []
    for (ex, ey), w, ok in zip(_TAP_DXY, wts, valid):
        idx = ((y0 + ey).clamp(0, th - 1) * tw + (x0 + ex).clamp(0, tw - 1))
        contrib = dsg * (w * ok)[:, :, None]  # B,G,hg,HW,Ns
        idx_e = idx[:, :, None].expand(B, G, hg, HW, Ns)
        dtab.scatter_add_(2, idx_e.permute(1, 2, 0, 3, 4).reshape(G, hg, -1),
                          contrib.permute(1, 2, 0, 3, 4).reshape(G, hg, -1))
        val = torch.gather(tab[None].expand(B, G, hg, th * tw), 3,
                           idx_e.reshape(B, G, hg, -1)).reshape(B, G, hg, HW, Ns)
        tapval.append(val * ok[:, :, None])
    g["rpe_table"] = dtab.reshape(h, th, tw)
    t_nw, t_ne, t_sw, t_se = tapval
    dbias_dix = (t_ne - t_nw) * wy0[:, :, None] + (t_se - t_sw) * wy1[:, :, None]
    dbias_diy = (t_sw - t_nw) * wx0[:, :, None] + (t_se - t_ne) * wx1[:, :, None]
    # ix = ((d+1)/2)(Tw-1), d = (grid - pos)/2  →  d ix / d pos_x = -(Tw-1)/4
    dpos_x = -(tw - 1) / 4.0 * (dsg * dbias_dix).sum(dim=(2, 3))  # B,G,Ns
    dpos_y = -(th - 1) / 4.0 * (dsg * dbias_diy).sum(dim=(2, 3))

    # proj_k / proj_v
    xs2 = fw["xs"].reshape(B * Ns, C)
    g["proj_k.weight"] = (dk.T @ xs2).reshape(C, C, 1, 1)
    g["proj_k.bias"] = dk.sum(0)
    g["proj_v.weight"] = (dv.T @ xs2).reshape(C, C, 1, 1)
    g["proj_v.bias"] = dv.sum(0)
    dxs = (dk @ _w2d(p, "proj_k.weight") + dv @ _w2d(p, "proj_v.weight")).reshape(B, Ns, G, Cg)
    dxs = dxs.permute(0, 2, 1, 3)  # B,G,Ns,Cg

    # feature sampling backward: scatter into dx, and d pos
    fpy, fpx = pos[..., 0].reshape(B, G, Ns), pos[..., 1].reshape(B, G, Ns)
    fx0, fy0, fw_, fvalid, (fwx0, fwx1, fwy0, fwy1) = bilinear_taps(fpx, fpy, W, H)
    xg = x.reshape(B, HW, G, Cg).permute(0, 2, 1, 3)
    dxg = torch.zeros(B, G, HW, Cg, dtype=x.dtype)
    ftap = []
    for (ex, ey), w, ok in zip(_TAP_DXY, fw_, fvalid):
        idx = ((fy0 + ey).clamp(0, H - 1) * W + (fx0 + ex).clamp(0, W - 1))
        idx_e = idx[..., None].expand(B, G, Ns, Cg)
        dxg.scatter_add_(2, idx_e, dxs * (w * ok)[..., None])
        ftap.append(torch.gather(xg, 2, idx_e) * ok[..., None])
    f_nw, f_ne, f_sw, f_se = ftap
    dxs_dix = (f_ne - f_nw) * fwy0[..., None] + (f_se - f_sw) * fwy1[..., None]
    dxs_diy = (f_sw - f_nw) * fwx0[..., None] + (f_se - f_ne) * fwx1[..., None]
    dpos_x = dpos_x + (W - 1) / 2.0 * (dxs * dxs_dix).sum(-1)
    dpos_y = dpos_y + (H - 1) / 2.0 * (dxs * dxs_diy).sum(-1)
    dx_sample = dxg.permute(0, 2, 1, 3).reshape(B, H, W, C)
    dpos = torch.stack((dpos_y, dpos_x), -1).reshape(B, G, hk, wk, 2)

    # clamp / tanh backward → d raw offsets
    orf = cfg.offset_range_factor
    if orf >= 0:
        rng = torch.tensor([1.0 / (hk - 1.0), 1.0 / (wk - 1.0)], dtype=torch.float32).to(x.dtype)
        th_ = torch.tanh(fw["off_raw"])
        doff = dpos * rng * orf * (1.0 - th_ * th_)
    else:
        pre = fw["pre"]
        doff = dpos * ((pre >= -1.0) & (pre <= 1.0)).to(x.dtype)

    # offset net backward
    act, ln, that, rstd = fw["off_act"], fw["off_ln"], fw["off_that"], fw["off_rstd"]
    w3 = p["conv_offset.3.weight"].reshape(2, Cg)
    doff_l = doff.permute(0, 2, 3, 1, 4)  # B,hk,wk,G,2
    g["conv_offset.3.weight"] = torch.einsum("bijgp,bijgc->pc", doff_l, act).reshape(2, Cg, 1, 1)
    dact = doff_l @ w3  # B,hk,wk,G,Cg
    cdf = 0.5 * (1.0 + torch.erf(ln / math.sqrt(2.0)))
    pdf = torch.exp(-0.5 * ln * ln) / math.sqrt(2.0 * math.pi)
    dln = dact * (cdf + ln * pdf)
    g["conv_offset.1.norm.weight"] = (dln * that).sum(dim=(0, 1, 2, 3))
    g["conv_offset.1.norm.bias"] = dln.sum(dim=(0, 1, 2, 3))
    dthat = dln * p["conv_offset.1.norm.weight"]
    dt = rstd * (dthat - dthat.mean(-1, keepdim=True) - that * (dthat * that).mean(-1, keepdim=True))
    g["conv_offset.0.bias"] = dt.sum(dim=(0, 1, 2, 3))
    qg = fw["q"].reshape(B, H, W, G, Cg)
    qp = F.pad(qg, (0, 0, 0, 0, pd, pd, pd, pd))
    dqp = torch.zeros_like(qp)
    w0 = p["conv_offset.0.weight"].reshape(Cg, k_, k_)
    dw0 = torch.zeros(Cg, k_, k_, dtype=x.dtype)
    for u in range(k_):
        for v in range(k_):
            sl = (slice(None), slice(u, u + s_ * (hk - 1) + 1, s_), slice(v, v + s_ * (wk - 1) + 1, s_))
            dw0[:, u, v] = (dt * qp[sl]).sum(dim=(0, 1, 2, 3))
            dqp[sl] += dt * w0[:, u, v]
    g["conv_offset.0.weight"] = dw0.reshape(Cg, 1, k_, k_)
    dq = dq + dqp[:, pd:pd + H, pd:pd + W].reshape(B, H, W, C)

    # proj_q
    dq2 = dq.reshape(B * HW, C)
    g["proj_q.weight"] = (dq2.T @ X2).reshape(C, C, 1, 1)
    g["proj_q.bias"] = dq2.sum(0)
    dx = (dq2 @ _w2d(p, "proj_q.weight")).reshape(B, H, W, C) + dx_sample
    extras = dict(do=do, ds=ds, dq=dq, dk=dk.reshape(B, Ns, C), dv=dv.reshape(B, Ns, C),
                  dxs=dxs.permute(0, 2, 1, 3).reshape(B, Ns, C), dpos=dpos, doff=doff, dt=dt,
                  dx_sample=dx_sample)
    return dx, g, extras


# ----------------------------------------------------------------------------
# library-operator form (CPU timing baseline; op sequence of dat_blocks.py:143-227)
# ----------------------------------------------------------------------------

def forward_libops(x_nchw: Tensor, p: Dict[str, Tensor], cfg: BlockCfg) -> Tensor:
    """x (B,C,H,W) → y (B,C,H,W) through the same library operators the reference
    module dispatches to.  Differentiable (autograd) — used for fwd+bwd timing."""
    B, C, H, W = x_nchw.shape
    G, Cg, h, hc, hg = cfg.n_groups, cfg.cg, cfg.n_heads, cfg.n_head_channels, cfg.hg
    q = F.conv2d(x_nchw, p["proj_q.weight"], p["proj_q.bias"])
    t = F.conv2d(q.reshape(B * G, Cg, H, W), p["conv_offset.0.weight"], p["conv_offset.0.bias"],
                 stride=cfg.stride, padding=cfg.pad, groups=Cg)
    t = F.layer_norm(t.permute(0, 2, 3, 1), (Cg,), p["conv_offset.1.norm.weight"],
                     p["conv_offset.1.norm.bias"], 1e-5).permute(0, 3, 1, 2)
    off = F.conv2d(F.gelu(t), p["conv_offset.3.weight"]).contiguous()
    hk, wk = off.shape[2], off.shape[3]
    Ns = hk * wk
    orf = cfg.offset_range_factor
    if orf >= 0:
        rng = torch.tensor([1.0 / (hk - 1.0), 1.0 / (wk - 1.0)], device=off.device).reshape(1, 2, 1, 1)  # fp32, :150
        off = off.tanh().mul(rng).mul(orf)
    off = off.permute(0, 2, 3, 1)
    ry, rx = ref_points(hk, wk, x_nchw.dtype)
    ref = torch.stack(torch.meshgrid(ry, rx, indexing="ij"), -1)[None].to(off.device)   # (device: GPU yardstick runs)
    pos = off + ref
    if orf < 0:
        pos = pos.clamp(-1.0, 1.0)
    if cfg.no_off:   # :164-167 — offsets ignored, keys/values from the average-pooled map
        xs = F.avg_pool2d(x_nchw, kernel_size=cfg.stride, stride=cfg.stride)
        hk, wk = xs.shape[2], xs.shape[3]
        Ns = hk * wk
        xs = xs.reshape(B, C, 1, Ns)
    else:
        xs = F.grid_sample(x_nchw.reshape(B * G, Cg, H, W), pos.flip(-1), mode="bilinear",
                           align_corners=True).reshape(B, C, 1, Ns)
    k = F.conv2d(xs, p["proj_k.weight"], p["proj_k.bias"]).reshape(B * h, hc, Ns)
    v = F.conv2d(xs, p["proj_v.weight"], p["proj_v.bias"]).reshape(B * h, hc, Ns)
    attn = torch.einsum("bcm,bcn->bmn", q.reshape(B * h, hc, H * W), k).mul(hc ** -0.5)
    mode = cfg.pe_mode
    lepe = None
    if mode in ("rpe", "log_cpb"):
        qy, qx = query_grid(H, W, x_nchw.dtype)
        qg = torch.stack(torch.meshgrid(qy, qx, indexing="ij"), -1).reshape(1, H * W, 1, 2).to(pos.device)
    if mode == "rpe":        # :198-214
        disp = (qg - pos.reshape(B * G, 1, Ns, 2)).mul(0.5)
        th, tw = p["rpe_table"].shape[1:]
        tab = p["rpe_table"].reshape(1, G, hg, th, tw).expand(B, G, hg, th, tw).reshape(B * G, hg, th, tw)
        bias = F.grid_sample(tab, disp.flip(-1), mode="bilinear", align_corners=True)
        attn = attn + bias.reshape(B * h, H * W, Ns)
    elif mode == "dwc":      # :185-186,221-222 — depthwise 3x3 on q, added to the attention output
        lepe = F.conv2d(q, p["rpe_table.weight"], p["rpe_table.bias"], padding=1, groups=C)
    elif mode == "fixed":    # :187-191 — dense table, bilinearly resized to (HW, Ns)
        bias = F.interpolate(p["rpe_table"][None].expand(B, -1, -1, -1), size=(H * W, Ns), mode="bilinear",
                             align_corners=True)
        attn = attn + bias.reshape(B * h, H * W, Ns)
    elif mode == "log_cpb":  # :192-197 — Swin-V2 style MLP on the log-scaled displacement
        disp = (qg - pos.reshape(B * G, 1, Ns, 2)).mul(4.0)
        disp = torch.sign(disp) * torch.log2(torch.abs(disp) + 1.0) / math.log2(8.0)
        hid = F.relu(F.linear(disp, p["rpe_table.0.weight"], p["rpe_table.0.bias"]))
        bias = F.linear(hid, p["rpe_table.2.weight"])                 # (B*G, HW, Ns, hg)
        attn = attn + bias.permute(0, 3, 1, 2).reshape(B * h, H * W, Ns)
    attn = F.softmax(attn, dim=2)
    out = torch.einsum("bmn,bcn->bcm", attn, v).reshape(B, C, H, W)
    if lepe is not None:
        out = out + lepe
    return F.conv2d(out, p["proj_out.weight"], p["proj_out.bias"])


class _LNHolder(torch.nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.norm = torch.nn.LayerNorm(dim)


class OracleDAttention(torch.nn.Module):
    """nn.Module wrapper over `forward_libops` with the reference's constructor
    signature (dat_blocks.py:21-26) and state-dict keys, so it can stand in for
    the block inside a backbone (CPU baseline timing, backbone-level parity)."""

    def __init__(self, q_size, kv_size, n_heads, n_head_channels, n_groups, attn_drop, proj_drop,
                 stride, offset_range_factor, use_pe, dwc_pe, no_off, fixed_pe, ksize, log_cpb,
                 stage_i):
        super().__init__()
        nn = torch.nn
        q_size = tuple(q_size) if isinstance(q_size, (tuple, list)) else (q_size, q_size)
        self.cfg = BlockCfg(q_size[0], q_size[1], n_heads, n_head_channels, n_groups, stride, ksize,
                            offset_range_factor, use_pe=bool(use_pe), dwc_pe=bool(dwc_pe), no_off=bool(no_off),
                            fixed_pe=bool(fixed_pe), log_cpb=bool(log_cpb))
        c, cg = self.cfg.nc, self.cfg.cg
        self.conv_offset = nn.Sequential(nn.Conv2d(cg, cg, ksize, stride, self.cfg.pad, groups=cg),
                                         _LNHolder(cg), nn.GELU(), nn.Conv2d(cg, 2, 1, bias=False))
        if no_off:
            for prm in self.conv_offset.parameters():
                prm.requires_grad_(False)
        self.proj_q, self.proj_k = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 1)
        self.proj_v, self.proj_out = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 1)
        mode = self.cfg.pe_mode            # same parameter names / shapes as dat_blocks.py:84-104
        if mode == "dwc":
            self.rpe_table = nn.Conv2d(c, c, 3, 1, 1, groups=c)
        elif mode == "fixed":
            kv_h, kv_w = q_size[0] // stride, q_size[1] // stride
            self.rpe_table = nn.Parameter(torch.zeros(n_heads, q_size[0] * q_size[1], kv_h * kv_w))
            nn.init.trunc_normal_(self.rpe_table, std=0.01)
        elif mode == "log_cpb":
            self.rpe_table = nn.Sequential(nn.Linear(2, 32, bias=True), nn.ReLU(inplace=True),
                                           nn.Linear(32, self.cfg.hg, bias=False))
        elif mode == "rpe":
            th, tw = self.cfg.table_hw
            self.rpe_table = nn.Parameter(torch.zeros(n_heads, th, tw))
            nn.init.trunc_normal_(self.rpe_table, std=0.01)
        else:
            self.rpe_table = None

    def forward(self, x):
        params = dict(self.named_parameters())
        return forward_libops(x, params, self.cfg), None, None
