"""Does the run-to-run mode of the captured step (about 878 vs 894 images/s) belong to the process
or to the graph capture?  Captures the same TrainStep several times in ONE process and times every capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
ts = bench.TrainStep(dev)
out = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    ts.warm_and_capture(3 if rep == 0 else 1)
    for _ in range(3):
        ts.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ts.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out.append(f"{16 / ms * 1e3:.1f}")
    ts.release()
print("captures in one process (images/s):", " ".join(out))
