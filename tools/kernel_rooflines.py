"""Achieved algorithmic throughput of the dat_b200 kernels at the DAT-T++ bench shapes (batch 16,
512x512 input, bf16 autocast dtypes), each launched alone through the C ABI: 3 rotating buffer sets
(> L2 in total), 30 back-to-back launches between CUDA events.  Prints a markdown table against the
measured peaks of MEASURED_PEAKS.json.   usage: python tools/kernel_rooflines.py > profiles/rNN_kernel_rooflines.md"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
flat = json.dumps(pk)
HBM = next((v for k, v in pk.items() if "hbm" in k.lower() and isinstance(v, (int, float))), 6541.0)
TF = next((v for k, v in pk.items() if "bf16" in k.lower() and isinstance(v, (int, float))), 1658.4)
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
f32, b16 = torch.float32, torch.bfloat16
CODE = {f32: 0, b16: 1}
NSET, REP, B = 3, 30, 16
rows = []


def timeit(fn):
    for i in range(NSET):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REP):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REP * 1e-3       # seconds


def add(name, shape, t, nbytes=None, flops=None):
    gbs = nbytes / t / 1e9 if nbytes else None
    tfs = flops / t / 1e12 if flops else None
    rows.append((name, shape, t * 1e6, gbs, tfs))


STAGES = [(64, 128), (128, 64), (256, 32), (512, 16)]
for s, (Cc, HW) in enumerate(STAGES):
    M = B * HW * HW
    # ---- LayerNorm fwd / bwd (fp32 stream in, bf16 out for the MLP) ----
    xs = [torch.randn(M, Cc, device="cuda") for _ in range(NSET)]
    ys = [torch.empty(M, Cc, device="cuda", dtype=b16) for _ in range(NSET)]
    g, bta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    t = timeit(lambda i: _cabi.check(lib.dat_layernorm_fwd(p(xs[i % NSET]), 0, p(g), p(bta), p(ys[i % NSET]), 1, p(mean),
                                                             p(rstd), M, Cc, 1e-5, st), "ln"))
    add("layernorm_fwd", f"s{s} {M}x{Cc} f32->bf16", t, nbytes=M * Cc * 6)
    dys = [torch.randn(M, Cc, device="cuda").to(b16) for _ in range(NSET)]
    dxs = [torch.empty(M, Cc, device="cuda") for _ in range(NSET)]
    dg, db = torch.empty(Cc, device="cuda"), torch.empty(Cc, device="cuda")
    nb = lib.dat_layernorm_bwd_workspace_bytes(M, Cc)
    ws = torch.empty(max(nb, 64), device="cuda", dtype=torch.uint8)
    t = timeit(lambda i: _cabi.check(lib.dat_layernorm_bwd(p(dys[i % NSET]), 1, p(xs[i % NSET]), 0, p(g), p(mean), p(rstd),
                                                             p(dxs[i % NSET]), p(xs[(i + 1) % NSET]), p(dg), p(db), M, Cc,
                                                             p(ws), nb, st), "lnb"))
    add("layernorm_bwd (+residual grad, +reduce)", f"s{s} {M}x{Cc}", t, nbytes=M * Cc * (2 + 4 + 4 + 4))
    # ---- residual + drop-path ----
    sc = torch.ones(B, device="cuda")
    t = timeit(lambda i: _cabi.check(lib.dat_scale_residual(p(dys[i % NSET]), 1, p(xs[i % NSET]), 0, p(sc), p(dxs[i % NSET]), 0,
                                                              B, HW * HW * Cc, st), "res"))
    add("scale_residual", f"s{s} {M}x{Cc} bf16+f32->f32", t, nbytes=M * Cc * 10)
    del xs, ys, dys, dxs
    # ---- MLP GEMMs (bf16) ----
    Hd = 4 * Cc
    for name, N, K, ydt in (("fc1 fwd", Hd, Cc, b16), ("fc2 fwd", Cc, Hd, b16), ("fc1 dgrad", Cc, Hd, b16)):
        xa = [torch.randn(M, K, device="cuda").to(b16) for _ in range(NSET)]
        ya = [torch.empty(M, N, device="cuda", dtype=ydt) for _ in range(NSET)]
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(b16)
        bias = torch.randn(N, device="cuda")
        t = timeit(lambda i: _cabi.check(lib.dat_pointwise_fwd_tc(p(xa[i % NSET]), 1, p(w), p(bias), p(ya[i % NSET]), CODE[ydt],
                                                                    M, N, K, st), "gemm"))
        add(f"gemm_tc_persistent {name}", f"s{s} M={M} N={N} K={K}", t, nbytes=M * (K + N) * 2, flops=2.0 * M * N * K)
        if name == "fc1 fwd":
            dyw = [torch.randn(M, N, device="cuda").to(b16) for _ in range(NSET)]
            nbw = lib.dat_pointwise_wgrad_tc_workspace_bytes(M, N, K)
            wsw = torch.empty(max(nbw, 64), device="cuda", dtype=torch.uint8)
            dw, dbb = torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
            t = timeit(lambda i: _cabi.check(lib.dat_pointwise_wgrad_tc(p(dyw[i % NSET]), p(xa[i % NSET]), p(dw), p(dbb), M, N, K,
                                                                          p(wsw), nbw, st), "wgrad"))
            add("gemm_tc_wgrad fc1 (+bias grad, +reduce)", f"s{s} M={M} N={N} K={K}", t, nbytes=M * (K + N) * 2, flops=2.0 * M * N * K)
            del dyw
        del xa, ya
    # ---- depthwise 3x3 (MLP middle, bf16) and 7x7 ('X' mixer, f32 -> bf16) ----
    for k, Cd, mode, xdt, ydt, tag in ((3, Hd, 2, b16, b16, "MLP middle"), (7, Cc, 0, f32, b16, "'X' mixer")):
        xa = [torch.randn(B, HW, HW, Cd, device="cuda").to(xdt) for _ in range(NSET)]
        ya = [torch.empty(B, HW, HW, Cd, device="cuda", dtype=ydt) for _ in range(NSET)]
        za = [torch.empty(B, HW, HW, Cd, device="cuda", dtype=ydt) for _ in range(NSET)]
        w = torch.randn(Cd, 1, k, k, device="cuda") / k
        bias = torch.randn(Cd, device="cuda")
        nbd = lib.dat_dwconv_workspace_bytes(B, HW, HW, Cd, k)
        wsd = torch.empty(nbd, device="cuda", dtype=torch.uint8)
        n = B * HW * HW * Cd
        ex, ey = xa[0].element_size(), ya[0].element_size()
        t = timeit(lambda i: _cabi.check(lib.dat_dwconv_fwd(p(xa[i % NSET]), CODE[xdt], p(w), p(bias), p(ya[i % NSET]),
                                                              p(za[i % NSET]), CODE[ydt], B, HW, HW, Cd, k, mode, 0, p(wsd), nbd,
                                                              st), "dw"))
        add(f"dwconv{k}_fwd {tag}", f"s{s} {B}x{HW}x{HW}x{Cd}", t, nbytes=n * (ex + ey * (2 if mode == 2 else 1)))
        if k == 3:
            dxa = [torch.empty_like(xa[0]) for _ in range(NSET)]
            dw, dbb = torch.empty_like(w), torch.empty_like(bias)
            t = timeit(lambda i: _cabi.check(lib.dat_dwconv_bwd(p(xa[i % NSET]), CODE[xdt], p(ya[i % NSET]), p(za[i % NSET]),
                                                                  CODE[ydt], p(w), p(dxa[i % NSET]), p(dw), p(dbb), B, HW, HW, Cd,
                                                                  3, mode, p(wsd), nbd, st), "dwb"))
            add("dwconv3_bwd fused (gelu', dx, dw, db, +reduce)", f"s{s} {B}x{HW}x{HW}x{Cd}", t, nbytes=n * (2 * ey + 2 * ex))
            del dxa
        del xa, ya, za

# ---- the deformable-attention block itself (K1 offset net, K2 gather, K3 attention, whole block) ----
BLOCK_STAGES = [  # DAT-T++ @512^2: (H=W, C, heads, groups, stride, ksize, q_size)
    (128, 64, 2, 1, 8, 9, 56), (64, 128, 4, 2, 4, 7, 28), (32, 256, 8, 4, 2, 5, 14), (16, 512, 16, 8, 1, 3, 7)]
for s, (Hh, Cc, heads, G, stride, ksize, qs) in enumerate(BLOCK_STAGES):
    HWn, Cg = Hh * Hh, Cc // G
    Th = 2 * qs - 1
    d = _cabi.BlockDesc(B, Hh, Hh, heads, G, stride, ksize, Th, Th, -1.0, _cabi.DAT_F32, _cabi.DAT_BF16)
    hk, wk = C.c_int32(), C.c_int32()
    _cabi.check(lib.dat_sample_grid(C.byref(d), C.byref(hk), C.byref(wk)), "grid")
    Ns = hk.value * wk.value
    gen = torch.Generator(device="cuda").manual_seed(s)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=gen)
    prm = [rn(Cg, 1, ksize, ksize) / ksize, rn(Cg) * 0.1, torch.ones(Cg, device="cuda"), torch.zeros(Cg, device="cuda"),
           rn(2, Cg, 1, 1) / Cg ** 0.5] + [t for _ in range(4) for t in (rn(Cc, Cc, 1, 1) / Cc ** 0.5, rn(Cc) * 0.1)] + \
          [rn(heads, Th, Th) * 0.1]
    ps = _cabi.BlockParams()
    for n_, t_ in zip(_cabi.PARAM_FIELDS, prm):
        setattr(ps, n_, t_.data_ptr())
    xa = [rn(B, HWn, Cc) for _ in range(NSET)]                       # block input, fp32 (LayerNorm output)
    qa = [rn(B, HWn, Cc).to(b16) for _ in range(NSET)]
    ka = [rn(B, Ns, Cc).to(b16) for _ in range(NSET)]
    va = [rn(B, Ns, Cc).to(b16) for _ in range(NSET)]
    oa = [torch.empty(B, HWn, Cc, device="cuda", dtype=b16) for _ in range(NSET)]
    xsa = [torch.empty(B, Ns, Cc, device="cuda", dtype=b16) for _ in range(NSET)]
    tdw = torch.empty(B, G, Ns, Cg, device="cuda")
    offr = torch.empty(B, G, Ns, 2, device="cuda")
    posa = [torch.empty(B, G, Ns, 2, device="cuda") for _ in range(NSET)]
    lse = torch.empty(B, heads, HWn, device="cuda")
    # K1: offset network -> pos
    t = timeit(lambda i: _cabi.check(lib.dat_offset_pos_fwd(C.byref(d), C.byref(ps), p(qa[i % NSET]), p(tdw), p(offr),
                                                             p(posa[i % NSET]), st), "k1"))
    add("K1 offset_pos_fwd (dw conv, LN, GELU, 1x1, ref, clamp)", f"s{s} q {B}x{HWn}x{Cc} k{ksize}/s{stride} -> {B}x{G}x{Ns}", t,
        nbytes=B * HWn * Cc * 2 + B * G * Ns * (4 * Cg + 16))
    # K2: bilinear gather
    t = timeit(lambda i: _cabi.check(lib.dat_sample_fwd(C.byref(d), p(xa[i % NSET]), p(posa[i % NSET]), p(xsa[i % NSET]), None,
                                                         st), "k2"))
    add("K2 sample_fwd (4-tap gather, fp32 x)", f"s{s} x {B}x{HWn}x{Cc} -> {B}x{Ns}x{Cc}", t,
        nbytes=4 * B * Ns * Cc * 4 + B * G * Ns * 8 + B * Ns * Cc * 2)
    # K3: attention core
    nws = lib.dat_attention_fwd_workspace_bytes(C.byref(d))
    wsa = torch.empty(max(nws, 64), device="cuda", dtype=torch.uint8)
    t = timeit(lambda i: _cabi.check(lib.dat_attention_fwd(C.byref(d), p(qa[i % NSET]), p(ka[i % NSET]), p(va[i % NSET]),
                                                            p(posa[i % NSET]), p(prm[13]), p(oa[i % NSET]), p(lse), p(wsa), nws, 0,
                                                            st), "k3"))
    add("K3 attention_fwd (QK^T + rpe bias + softmax + PV)", f"s{s} {B}x{heads} heads, {HWn}x{Ns} scores", t,
        nbytes=(2 * B * HWn * Cc + 2 * B * Ns * Cc) * 2 + B * G * Ns * 8 + heads * Th * Th * 4 + B * heads * HWn * 4,
        flops=4.0 * HWn * Ns * Cc * B)
    # whole block forward / backward through dat_block_forward / dat_block_backward
    e_ = lambda *sh, dt=b16: torch.empty(*sh, device="cuda", dtype=dt)
    saved = [e_(B, HWn, Cc), e_(B, G, Ns, Cg, dt=f32), e_(B, G, Ns, 2, dt=f32), e_(B, G, Ns, 2, dt=f32), e_(B, Ns, Cc),
             e_(B, Ns, Cc), e_(B, Ns, Cc), e_(B, HWn, Cc), e_(B, heads, HWn, dt=f32)]
    ss = _cabi.BlockSaved()
    for n_, t_ in zip(_cabi.SAVED_FIELDS, saved):
        setattr(ss, n_, t_.data_ptr())
    nf = lib.dat_block_fwd_workspace_bytes(C.byref(d))
    nbk = lib.dat_block_bwd_workspace_bytes(C.byref(d))
    wsf = torch.empty(max(nf, nbk, 64), device="cuda", dtype=torch.uint8)
    ya = [e_(B, HWn, Cc) for _ in range(NSET)]
    t = timeit(lambda i: _cabi.check(lib.dat_block_forward(C.byref(d), C.byref(ps), p(xa[i % NSET]), p(ya[i % NSET]), C.byref(ss),
                                                            p(wsf), nf, st), "blockf"))
    dense_f = (4.0 * HWn * Cc * Cc + 4.0 * Ns * Cc * Cc + 4.0 * HWn * Ns * Cc) * B
    add("block forward (8 launches)", f"s{s} x {B}x{HWn}x{Cc} f32 -> y bf16", t, nbytes=B * HWn * Cc * (4 + 2) + 4 * Cc * Cc * 4,
        flops=dense_f)
    grads = [torch.empty_like(t_) for t_ in prm]
    gs = _cabi.BlockGrads()
    for n_, t_ in zip(_cabi.PARAM_FIELDS, grads):
        setattr(gs, n_, t_.data_ptr())
    dya = [rn(B, HWn, Cc).to(b16) for _ in range(NSET)]
    dxa = [torch.empty(B, HWn, Cc, device="cuda") for _ in range(NSET)]
    t = timeit(lambda i: _cabi.check(lib.dat_block_backward(C.byref(d), C.byref(ps), p(xa[NSET - 1]), p(dya[i % NSET]),
                                                             C.byref(ss), p(dxa[i % NSET]), C.byref(gs), p(wsf), nbk, st), "blockb"))
    saved_bytes = B * HWn * Cc * 2 * 2 + 3 * B * Ns * Cc * 2 + B * G * Ns * (4 * Cg + 16) + B * heads * HWn * 4
    add("block backward (all gradients)", f"s{s} dy bf16 -> dx f32 + 14 parameter grads", t,
        nbytes=B * HWn * Cc * (2 + 4 + 4) + saved_bytes + 8 * Cc * Cc * 4, flops=2.0 * dense_f + 2.0 * HWn * Ns * Cc * B)
    del xa, qa, ka, va, oa, xsa, saved, ya, dya, dxa

print("# dat_b200 kernels against the measured B200 peaks\n")
print(f"HBM peak {HBM:.0f} GB/s, dense bf16 peak {TF:.1f} TFLOP/s (MEASURED_PEAKS.json).  Each kernel launched alone through the C ABI, "
      f"batch {B}, 3 rotating buffer sets, {REP} launches between CUDA events (tools/kernel_rooflines.py).  "
      "GB/s = algorithmic bytes (every tensor once) / time.\n")
print("| kernel | shape | µs | GB/s | % HBM peak | TFLOP/s | % bf16 peak |\n|---|---|---:|---:|---:|---:|---:|")
for name, shape, us, gbs, tfs in rows:
    print(f"| {name} | {shape} | {us:.1f} | " + (f"{gbs:.0f} | {100 * gbs / HBM:.0f} %" if gbs else "| ") + " | " +
          (f"{tfs:.0f} | {100 * tfs / TF:.0f} %" if tfs else " | ") + " |")
