"""Where the persistent tcgen05 GEMM spends its time: the same launches with parts switched off through
DAT_B200_GEMM_DBG (1 = no global stores, 2 = weight panel loaded for a CTA's first tile only, 4 = activations loaded
for the first tile only, 8 = no epilogue work; results are wrong in those modes - timing only).
usage: for d in 0 1 8 2 6 14; do DAT_B200_GEMM_DBG=$d python tools/gemm_decompose.py; done"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
NSET, REP = 3, 30
b16 = torch.bfloat16
CASES = [("s0 fc1", 262144, 256, 64), ("s0 fc2", 262144, 64, 256), ("s1 fc1", 65536, 512, 128), ("s1 fc2", 65536, 128, 512),
         ("s2 fc1", 16384, 1024, 256), ("s2 fc2", 16384, 256, 1024), ("s2 proj", 16384, 256, 256),
         ("s3 fc1", 4096, 2048, 512), ("s3 fc2", 4096, 512, 2048)]
out = []
for name, M, N, K in CASES:
    sets = [(torch.randn(M, K, device="cuda").to(b16), torch.empty(M, N, device="cuda", dtype=b16)) for _ in range(NSET)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(b16)
    bias = torch.randn(N, device="cuda")

    def fwd(i):
        x, y = sets[i % NSET]
        _cabi.check(lib.dat_pointwise_fwd_tc(p(x), 1, p(w), p(bias), p(y), 1, M, N, K, st), "gemm")

    for i in range(NSET):
        fwd(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REP):
        fwd(i)
    e1.record()
    torch.cuda.synchronize()
    out.append(f"{name} {e0.elapsed_time(e1) / REP * 1e3:.1f}")
print(f"dbg={os.environ.get('DAT_B200_GEMM_DBG', '0'):>2s} bn={os.environ.get('DAT_B200_GEMM_BN', '-'):>3s} | " + " | ".join(out))
