"""Runs the individual hand-written kernels of one block at a DAT-T++ stage shape a few
times (B=16, bf16) — the short command that ncu wraps for --set full captures."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi
from dat_segmentation_b200.dattention import DAttentionBaseline

STAGES = [(128, 2, 1, 8, 9, 56), (64, 4, 2, 4, 7, 28), (32, 8, 4, 2, 5, 14), (16, 16, 8, 1, 3, 7)]
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = sys.argv[2] if len(sys.argv) > 2 else "fwdbwd"
B = 16
H, heads, groups, stride, ksize, qs = STAGES[stage]
torch.manual_seed(0)
m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False, False,
                       False, ksize, False, stage).cuda()
with torch.no_grad():
    m.conv_offset[3].weight.mul_(2.0)
    m.rpe_table.mul_(10.0)
x = torch.randn(B, H, H, heads * 32, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    flush.zero_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)[0]
    if mode == "fwdbwd":
        y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done", stage, mode)
