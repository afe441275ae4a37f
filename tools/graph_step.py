"""Eager vs CUDA-graph replay of one DAT-T++ fwd+bwd step (B=16, 512x512, bf16 autocast)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dat_segmentation_b200.backbone import build_dat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
model = build_dat().cuda().train()
x = torch.randn(B, 3, 512, 512, device="cuda")

def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    loss = sum(o.float().mean() for o in outs)
    loss.backward()
    return loss

def timeit(fn, n=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

for _ in range(3):
    model.zero_grad(set_to_none=True); step()
print("eager  ms/step (events, wall):", timeit(lambda: (model.zero_grad(set_to_none=True), step())))
# CPU-only cost of issuing the step (GPU idle afterwards is not waited for)
torch.cuda.synchronize(); t0 = time.perf_counter(); model.zero_grad(set_to_none=True); step(); t1 = time.perf_counter(); torch.cuda.synchronize()
print("eager  CPU issue time of one step: %.1f ms" % ((t1 - t0) * 1e3))

s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        model.zero_grad(set_to_none=True); step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
model.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    static_loss = step()
print("graph  ms/step (events, wall):", timeit(g.replay), "loss", static_loss.item())
