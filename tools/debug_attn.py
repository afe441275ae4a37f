"""Localises differences between the tensor-core and the CUDA-core attention forward (debug aid)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi
from oracle import dattn_oracle as orc

lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(H, W, heads, groups, stride, ksize, qs, pos_range, B=2, tab_scale=0.5):
    cfg = orc.BlockCfg(qs, qs, heads, 32, groups, stride, ksize, -1)
    d = _cabi.BlockDesc(B, H, W, heads, groups, stride, ksize, 2 * qs - 1, 2 * qs - 1, -1.0, 0, 1)
    hk, wk = cfg.sample_grid(H, W)
    Ns, Cc = hk * wk, heads * 32
    g = torch.Generator().manual_seed(H * W + heads)
    q = torch.randn(B, H * W, Cc, generator=g).bfloat16().cuda()
    k = torch.randn(B, Ns, Cc, generator=g).bfloat16().cuda()
    v = torch.randn(B, Ns, Cc, generator=g).bfloat16().cuda()
    pos = ((torch.rand(B, groups, Ns, 2, generator=g) * 2 - 1) * pos_range).cuda()
    tab = (torch.randn(heads, 2 * qs - 1, 2 * qs - 1, generator=g) * tab_scale).cuda()
    nb = lib.dat_attention_fwd_workspace_bytes(C.byref(d))
    ws = torch.empty(max(nb, 64), dtype=torch.uint8, device="cuda")
    res = {}
    for impl in (0, 1):
        o = torch.zeros_like(q)
        lse = torch.zeros(B, heads, H * W, device="cuda")
        _cabi.check(lib.dat_attention_fwd(C.byref(d), p(q), p(k), p(v), p(pos), p(tab), p(o), p(lse), p(ws), nb, impl, st()), "attn")
        torch.cuda.synchronize()
        res[impl] = (o.float().cpu(), lse.cpu())
    eo = (res[0][0] - res[1][0]).abs().reshape(B, H * W, heads, 32)
    el = (res[0][1] - res[1][1]).abs()
    print(f"case H={H} W={W} heads={heads} G={groups} s={stride} Ns={Ns} pos_range={pos_range}: max|do| {eo.max():.3e} max|dlse| {el.max():.3e}")
    if eo.max() > 5e-3:
        print("  per batch:", [f"{eo[b].max():.2e}" for b in range(B)])
        print("  per head :", [f"{eo[:, :, h].max():.2e}" for h in range(heads)])
        nt = (H * W + 127) // 128
        print("  per tile :", [f"{eo[:, t * 128:(t + 1) * 128].max():.2e}" for t in range(nt)])
        rowerr = eo.amax(dim=(0, 2, 3))
        bad = (rowerr > 5e-3).nonzero().flatten().tolist()
        print("  bad query rows:", len(bad), bad[:40])
        print("  lse err per tile:", [f"{el[:, :, t * 128:(t + 1) * 128].max():.2e}" for t in range(nt)])


if __name__ == "__main__":
    run(16, 64, 8, 4, 1, 3, 7, 1.1)
    run(16, 64, 8, 4, 1, 3, 7, 1.0)
    run(16, 16, 8, 4, 1, 3, 7, 1.1)      # Ns = 256, out-of-range samples
    run(16, 16, 8, 4, 1, 3, 7, 1.0)
    run(32, 32, 8, 4, 2, 5, 14, 1.1)
    run(16, 64, 8, 4, 1, 3, 7, 1.1, tab_scale=0.0)
    run(48, 32, 2, 1, 2, 5, 14, 1.0)


def probe(H=16, W=16, heads=8, groups=4, stride=1, ksize=3, qs=7, B=1):
    """lse with all the softmax mass on one sample n0 = S[m, n0] + bias[m, n0]: compares the bias of single samples."""
    cfg = orc.BlockCfg(qs, qs, heads, 32, groups, stride, ksize, -1)
    d = _cabi.BlockDesc(B, H, W, heads, groups, stride, ksize, 2 * qs - 1, 2 * qs - 1, -1.0, 0, 1)
    hk, wk = cfg.sample_grid(H, W)
    Ns, Cc = hk * wk, heads * 32
    g = torch.Generator().manual_seed(1)
    tab = (torch.randn(heads, 2 * qs - 1, 2 * qs - 1, generator=g) * 0.5).cuda()
    nb = lib.dat_attention_fwd_workspace_bytes(C.byref(d))
    ws = torch.empty(max(nb, 64), dtype=torch.uint8, device="cuda")
    q = torch.ones(B, H * W, Cc).bfloat16().cuda()
    v = torch.randn(B, Ns, Cc, generator=g).bfloat16().cuda()
    for (py, px) in [(0.3, 0.2), (0.3, 1.05), (0.3, -1.08), (1.05, 0.2), (-1.07, 0.1), (1.04, 1.06), (0.3, 1.0), (0.3, 0.999)]:
        pos = ((torch.rand(B, groups, Ns, 2, generator=g) * 2 - 1) * 0.9)
        n0 = 37
        pos[:, :, n0, 0] = py
        pos[:, :, n0, 1] = px
        pos = pos.cuda()
        k = torch.zeros(B, Ns, Cc)
        k[:, n0] = 10.0
        k = k.bfloat16().cuda()
        res = {}
        for impl in (0, 1):
            o = torch.zeros_like(q)
            lse = torch.zeros(B, heads, H * W, device="cuda")
            _cabi.check(lib.dat_attention_fwd(C.byref(d), p(q), p(k), p(v), p(pos), p(tab), p(o), p(lse), p(ws), nb, impl, st()), "attn")
            torch.cuda.synchronize()
            res[impl] = lse.cpu()
        e = (res[0] - res[1]).abs()
        i = e.argmax().item()
        hh, mm = (i // (H * W)) % heads, i % (H * W)
        print(f"probe pos=({py},{px}): max|dlse| {e.max():.4f} at head {hh} m {mm} (r {mm // W}, c {mm % W}); tc {res[0].flatten()[i]:.4f} simt {res[1].flatten()[i]:.4f}"
              f"; mean|dlse| {e.mean():.5f}")


if __name__ == "__main__":
    probe()
    run(32, 32, 8, 4, 2, 5, 14, 1.0)
    run(32, 32, 8, 4, 2, 5, 14, 0.5)
