"""Kernel-level timing of the 3x3 depthwise kernels through the C ABI (no autograd / allocator in
the timed loop): 3 rotating buffer sets (> L2 in total), 30 back-to-back launches between CUDA
events.  Prints achieved algorithmic GB/s.  usage: time_dwconv3_abi.py [th ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
B = 16
CASES = [("mlp s0", 256, 128, 2, torch.bfloat16, torch.bfloat16), ("mlp s1", 512, 64, 2, torch.bfloat16, torch.bfloat16),
         ("mlp s2", 1024, 32, 2, torch.bfloat16, torch.bfloat16), ("mlp s3", 2048, 16, 2, torch.bfloat16, torch.bfloat16),
         ("lpu s0", 64, 128, 1, torch.float32, torch.float32), ("lpu s1", 128, 64, 1, torch.float32, torch.float32),
         ("lpu s2", 256, 32, 1, torch.float32, torch.float32), ("lpu s3", 512, 16, 1, torch.float32, torch.float32)]
CODE = {torch.float32: 0, torch.bfloat16: 1}
NSET, REP = 3, 30
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(th):
    if th:
        os.environ["DAT_B200_DW3_TH"] = str(th)
    else:
        os.environ.pop("DAT_B200_DW3_TH", None)
    print(f"--- rows per strip: {th or 'auto'}")
    for name, Cc, HW, mode, xdt, ydt in CASES:
        sets = []
        for _ in range(NSET):
            x = torch.randn(B, HW, HW, Cc, device="cuda").to(xdt)
            dy = torch.randn(B, HW, HW, Cc, device="cuda").to(ydt)
            sets.append((x, dy, torch.empty_like(dy), torch.empty_like(dy), torch.empty_like(x)))
        w = torch.randn(Cc, 1, 3, 3, device="cuda") / 3
        b = torch.randn(Cc, device="cuda")
        dw, db = torch.empty_like(w), torch.empty_like(b)
        nb = lib.dat_dwconv_workspace_bytes(B, HW, HW, Cc, 3)
        ws = torch.empty(nb, device="cuda", dtype=torch.uint8)

        def fwd(i):
            x, dy, y, z, dx = sets[i % NSET]
            _cabi.check(lib.dat_dwconv_fwd(p(x), CODE[xdt], p(w), p(b), p(y), p(z), CODE[ydt], B, HW, HW, Cc, 3, mode,
                                           0, p(ws), nb, st), "fwd")

        def bwd(i):
            x, dy, y, z, dx = sets[i % NSET]
            _cabi.check(lib.dat_dwconv_bwd(p(x), CODE[xdt], p(dy), p(z), CODE[ydt], p(w), p(dx), p(dw), p(db), B, HW,
                                           HW, Cc, 3, mode, p(ws), nb, st), "bwd")

        res = []
        for fn in (fwd, bwd):
            for i in range(NSET):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(REP):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / REP)
        n = B * HW * HW * Cc
        ex, ey = sets[0][0].element_size(), sets[0][1].element_size()
        fb = n * (ex + ey * (2 if mode == 2 else 1))
        bb = n * (ey * (2 if mode == 2 else 1) + 2 * ex)
        print(f"{name:8s} C={Cc:5d} {HW:3d}^2 fwd {res[0]*1e3:7.1f} us {fb/res[0]/1e6:6.0f} GB/s | "
              f"bwd(+reduce) {res[1]*1e3:7.1f} us {bb/res[1]/1e6:6.0f} GB/s")


for th in ([int(a) for a in sys.argv[1:]] or [0]):
    run(th)
