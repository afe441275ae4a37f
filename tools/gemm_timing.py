"""Phase timing of the tcgen05 GEMM (CTA 0) + CUDA-event timing for a few shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dat_segmentation_b200 import _cabi
lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for kind, M, Cc in [("bf16", 128, 256), ("bf16", 4096, 256), ("bf16", 16384, 256), ("tf32", 16384, 256), ("tf32", 262144, 64), ("bf16", 262144, 64), ("tf32", 4096, 512)]:
    X = torch.randn(M, Cc, device="cuda")
    W = torch.randn(Cc, Cc, device="cuda")
    b = torch.randn(Cc, device="cuda")
    Y = torch.empty(M, Cc, device="cuda", dtype=torch.bfloat16)
    if kind == "bf16":
        X, W = X.bfloat16(), W.bfloat16()
    code = 0 if kind == "tf32" else 1
    run = lambda: _cabi.check(lib.dat_pointwise_fwd_tc(p(X), code, p(W), p(b), p(Y), 1, M, Cc, Cc, st), "gemm")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    out = (C.c_uint64 * 8)()
    lib.dat_debug_gemm_timing(out)
    t = [int(v) for v in out]
    ph = [t[i + 1] - t[i] for i in range(5)]
    print(f"{kind} M={M} C={Cc}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us/launch (back-to-back); CTA0 phases ns: setup {ph[0]}, first-stage {ph[1]}, mma-issue {ph[2]}, acc-ready {t[4]-t[3]}, epilogue {t[5]-t[4]}, total {t[5]-t[0]}")
