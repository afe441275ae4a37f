"""Times the attention backward (dS streamed out) and the tensor-core table-gradient kernel inside one block
backward per DAT-T++ stage (B = 16), using the torch profiler's CUDA kernel table."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from dat_segmentation_b200.dattention import DAttentionBaseline

STAGES = [(128, 2, 1, 8, 9, 56), (64, 4, 2, 4, 7, 28), (32, 8, 4, 2, 5, 14), (16, 16, 8, 1, 3, 7)]
for stage, (H, heads, groups, stride, ksize, qs) in enumerate(STAGES):
    torch.manual_seed(0)
    m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False, False, False, ksize,
                           False, stage).cuda()
    x = torch.randn(16, H, H, heads * 32, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)[0]
        y.backward(torch.ones_like(y))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run()
        torch.cuda.synchronize()
    import re
    allk = os.environ.get("ALL_KERNELS") is not None
    rows = [(e.key, e.device_time_total / 5) for e in prof.key_averages()
            if allk or "attn_bwd" in e.key or "rpe_table_grad" in e.key or "attn_fwd" in e.key]
    rows.sort(key=lambda kv: -kv[1])
    name = lambda k: re.sub(r"[(<].*", "", k.replace("(anonymous namespace)::", "").replace("void ", "").replace("dat::", ""))[:40]
    print(f"stage {stage}: " + ", ".join(f"{name(k)} {t:.0f} us" for k, t in rows[:24]), flush=True)
