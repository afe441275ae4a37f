"""Block fwd+bwd at the 512 x 2048 training-crop shapes of DAT-T++ (16 x 64 = 1024 samples at every stage), batch 4,
bf16 autocast: ms per fwd+bwd with the tensor-core attention backward over sample chunks (default) and with
DAT_B200_ATTN_BWD_SIMT_LARGE_NS=1 (the round-1 path: CUDA-core backward for more than 256 samples).
usage: python tools/time_block_large_ns.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200.dattention import DAttentionBaseline

# (H, W, heads, groups, stride, ksize, q_size) of stages 1-3 for a 512 x 2048 input (stage 0: 128 x 512 map)
STAGES = [(64, 256, 4, 2, 4, 7, 28), (32, 128, 8, 4, 2, 5, 14), (16, 64, 16, 8, 1, 3, 7)]
B = 4
out = []
for H, W, heads, groups, stride, ksize, qs in STAGES:
    torch.manual_seed(0)
    m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False, False, False, ksize,
                           False, 0).cuda()
    x = torch.randn(B, H, W, heads * 32, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)[0]
        y.backward(torch.ones_like(y))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    out.append(f"{H}x{W} h{heads}: {e0.elapsed_time(e1) / 10:.3f} ms")
print(("simt-bwd  " if os.environ.get("DAT_B200_ATTN_BWD_SIMT_LARGE_NS") else "tc-bwd    ") + " | ".join(out))
