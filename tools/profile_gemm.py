"""tcgen05 GEMM launches (s2 fc1, s2 fc2, s0 fc2 shapes of the DAT-T++ MLPs, B = 16) between cudaProfilerStart/Stop."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
b16 = torch.bfloat16
for (M, N, K) in [(16384, 1024, 256), (16384, 256, 1024), (262144, 64, 256)]:
    x = torch.randn(M, K, device="cuda").to(b16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(b16)
    bias = torch.randn(N, device="cuda")
    y = torch.empty(M, N, device="cuda", dtype=b16)
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
        _cabi.check(lib.dat_pointwise_fwd_tc(p(x), 1, p(w), p(bias), p(y), 1, M, N, K, st), "gemm")
        if it == 2:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
print("profiled")
