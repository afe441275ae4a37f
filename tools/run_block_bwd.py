"""One block fwd+bwd at a DAT-T++ stage shape (B = 16, bf16 autocast) - a small target for ncu.
usage: python tools/run_block_bwd.py <stage> [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200.dattention import DAttentionBaseline

STAGES = [(128, 2, 1, 8, 9, 56), (64, 4, 2, 4, 7, 28), (32, 8, 4, 2, 5, 14), (16, 16, 8, 1, 3, 7)]
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, heads, groups, stride, ksize, qs = STAGES[stage]
torch.manual_seed(0)
m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False, False, False, ksize, False,
                       stage).cuda()
x = torch.randn(16, H, H, heads * 32, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
for _ in range(iters):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x)[0]
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("done")
