"""Times the channel-last depthwise-conv ops at the DAT-T++ shapes (batch 16, 512x512 input):
fwd and fwd+bwd per call, with the algorithmic bytes / time = achieved GB/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200.dwconv import dwconv_cl

B = 16
CASES = [("mlp s0", 256, 128, 3, 2, torch.bfloat16, torch.bfloat16), ("mlp s1", 512, 64, 3, 2, torch.bfloat16, torch.bfloat16),
         ("mlp s2", 1024, 32, 3, 2, torch.bfloat16, torch.bfloat16), ("mlp s3", 2048, 16, 3, 2, torch.bfloat16, torch.bfloat16),
         ("lpu s0", 64, 128, 3, 1, torch.float32, torch.float32), ("lpu s2", 256, 32, 3, 1, torch.float32, torch.float32),
         ("X s0", 64, 128, 7, 0, torch.float32, torch.bfloat16), ("X s2", 256, 32, 7, 0, torch.float32, torch.bfloat16)]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name, C, HW, k, mode, xdt, ydt in CASES:
    x = torch.randn(B, HW, HW, C, device="cuda").to(xdt).permute(0, 3, 1, 2).requires_grad_(True)
    w = (torch.randn(C, 1, k, k, device="cuda") / k).requires_grad_(True)
    b = torch.randn(C, device="cuda").requires_grad_(True)
    dy = torch.randn(B, HW, HW, C, device="cuda").to(ydt).permute(0, 3, 1, 2)
    ex, ey = x.element_size(), dy.element_size()
    n = B * HW * HW * C
    fwd_bytes = n * (ex + ey * (2 if mode == 2 else 1))
    bwd_bytes = n * (ey * (2 if mode == 2 else 1) + 2 * ex)

    def f():
        with torch.no_grad():
            dwconv_cl(x, w, b, mode, ydt)

    def fb():
        x.grad = w.grad = b.grad = None
        dwconv_cl(x, w, b, mode, ydt).backward(dy)

    tf, tfb = timeit(f), timeit(fb)
    tb = tfb - tf
    print(f"{name:8s} C={C:5d} {HW:3d}^2 k={k} fwd {tf*1e3:7.1f} us {fwd_bytes/tf/1e6:7.0f} GB/s | "
          f"bwd {tb*1e3:7.1f} us {bwd_bytes/tb/1e6:7.0f} GB/s")
