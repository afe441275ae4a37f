"""Quick per-stage timing of the block (fwd, fwd+bwd) — development aid, CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dat_segmentation_b200.dattention import DAttentionBaseline

STAGES = [(128, 2, 1, 8, 9, 56), (64, 4, 2, 4, 7, 28), (32, 8, 4, 2, 5, 14), (16, 16, 8, 1, 3, 7)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[len(ts) // 2]


for dt in ("bf16", "fp32"):
    for si, (H, heads, groups, stride, ksize, qs) in enumerate(STAGES):
        m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False,
                               False, False, ksize, False, si).cuda()
        x = torch.randn(B, heads * 32, H, H, device="cuda").permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        x.requires_grad_(True)

        def fwd():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == "bf16"), torch.no_grad():
                return m(x)[0]

        def fwdbwd():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == "bf16"):
                y = m(x)[0]
            y.backward(torch.ones_like(y))

        print(json.dumps(dict(dtype=dt, stage=si, B=B, fwd_ms=round(timeit(fwd), 4), fwdbwd_ms=round(timeit(fwdbwd), 4))), flush=True)
