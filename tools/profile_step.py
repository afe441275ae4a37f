"""One DAT-T++ backbone fwd+bwd step (batch 16, 512x512, bf16 autocast) between
cudaProfilerStart/Stop, after warm-up: the command ncu wraps
(`ncu --profile-from-start off ...`).  Prints nothing that is a bench value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200.backbone import build_dat

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
model = build_dat().cuda().train()
x = torch.randn(B, 3, 512, 512, device="cuda")


def step():
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = model(x)
    sum(o.float().mean() for o in outs).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
