"""Kernel-level timing of the tcgen05 1x1-conv GEMMs through the C ABI at the DAT-T++ MLP / block
shapes (batch 16, 512x512): 3 rotating buffer sets, 30 launches between CUDA events; prints
achieved TFLOP/s and algorithmic GB/s."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
CODE = {torch.float32: 0, torch.bfloat16: 1}
NSET, REP = 3, 30
f32, b16 = torch.float32, torch.bfloat16
B = 16
CASES = []
for s, (Cc, HW) in enumerate([(64, 128), (128, 64), (256, 32), (512, 16)]):
    M = B * HW * HW
    CASES += [(f"s{s} fc1 fwd", M, 4 * Cc, Cc, f32, b16), (f"s{s} fc2 fwd", M, Cc, 4 * Cc, b16, b16),
              (f"s{s} fc1 dgrad", M, Cc, 4 * Cc, b16, f32), (f"s{s} fc2 dgrad", M, 4 * Cc, Cc, b16, b16),
              (f"s{s} proj fwd", M, Cc, Cc, b16, b16)]


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
    return e0.elapsed_time(e1) / REP


for name, M, N, K, xdt, ydt in CASES:
    sets = [(torch.randn(M, K, device="cuda").to(xdt), torch.empty(M, N, device="cuda", dtype=ydt)) for _ in range(NSET)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(xdt)
    bias = torch.randn(N, device="cuda")

    def fwd(i):
        x, y = sets[i % NSET]
        _cabi.check(lib.dat_pointwise_fwd_tc(p(x), CODE[xdt], p(w), p(bias), p(y), CODE[ydt], M, N, K, st), "gemm")

    t = timeit(fwd)
    byts = M * K * sets[0][0].element_size() + M * N * sets[0][1].element_size()
    print(f"{name:14s} M={M:6d} N={N:5d} K={K:5d} {str(xdt)[6:]:>8s}->{str(ydt)[6:]:<8s} {t*1e3:7.1f} us "
          f"{2*M*N*K/t/1e9:7.1f} TF/s {byts/t/1e6:6.0f} GB/s")
    if xdt == b16 and N % 64 == 0 and K % 64 == 0:
        dy = [torch.randn(M, N, device="cuda").to(b16) for _ in range(NSET)]
        nb = lib.dat_pointwise_wgrad_tc_workspace_bytes(M, N, K)
        ws = torch.empty(max(nb, 64), device="cuda", dtype=torch.uint8)
        dw = torch.empty(N, K, device="cuda")
        dbias = torch.empty(N, device="cuda")

        def wg(i):
            _cabi.check(lib.dat_pointwise_wgrad_tc(p(dy[i % NSET]), p(sets[i % NSET][0]), p(dw), p(dbias), M, N, K, p(ws), nb, st), "wgrad")

        t = timeit(wg)
        print(f"{'   wgrad':14s} M={M:6d} N={N:5d} K={K:5d} {'':18s} {t*1e3:7.1f} us {2*M*N*K/t/1e9:7.1f} TF/s "
              f"{(M*K*2+M*N*2)/t/1e6:6.0f} GB/s")
