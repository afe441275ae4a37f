"""BASELINE.json configs[4]: DAttention microbench sweep — DAT-T++ stage presets x groups {1,2,4,8} x
offset_range_factor {1,2,3} x feature maps 16^2..128^2, fwd+bwd through the drop-in module under bf16
autocast, batch 16 (working set of the rotating inputs > L2 for the large maps; an L2 flush precedes
every timed pair otherwise).  Prints a markdown table.   usage: python tools/sweep_dattn.py > profiles/rNN_sweep_dattn.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from dat_segmentation_b200.dattention import DAttentionBaseline

PRESETS = {128: (2, 8, 9, 56), 64: (4, 4, 7, 28), 32: (8, 2, 5, 14), 16: (16, 1, 3, 7)}
B, REP = 16, 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("# DAttention sweep (BASELINE.json configs[4]), one B200, batch 16, bf16 autocast, fwd+bwd\n")
print("Each row: the drop-in module (`dat_block_forward` + `dat_block_backward`), median of 10 CUDA-event timings, "
      "L2 flushed before every timed fwd+bwd pair.  Dense GFLOP = 3 x (4 HW C^2 + 4 Ns C^2 + 4 HW Ns C) per image.\n")
print("| map | C | heads | groups | orf | Ns | fwd ms | bwd ms | images/s | dense TFLOP/s |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for hw, (heads, stride, ksize, qs) in PRESETS.items():
    Cc = heads * 32
    for groups in (1, 2, 4, 8):
        if heads % groups:
            continue
        for orf in (1, 2, 3):
            torch.manual_seed(hw + groups + orf)
            m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, orf, True, False, False, False,
                                   ksize, False, 0).cuda()
            x = torch.randn(B, hw, hw, Cc, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)   # NHWC in situ
            dy = torch.randn(B, hw, hw, Cc, device="cuda").permute(0, 3, 1, 2).bfloat16()
            tf, tb = [], []
            for it in range(REP + 3):
                flush.zero_()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y, _, _ = m(x)
                e[1].record()
                y.backward(dy)
                e[2].record()
                torch.cuda.synchronize()
                if it >= 3:
                    tf.append(e[0].elapsed_time(e[1]))
                    tb.append(e[1].elapsed_time(e[2]))
                x.grad = None
            f, b = sorted(tf)[REP // 2], sorted(tb)[REP // 2]
            Ns = (hw // stride) ** 2 if ksize != stride else ((hw - ksize) // stride + 1) ** 2
            gf = 3 * (4.0 * hw * hw * Cc * Cc + 4.0 * Ns * Cc * Cc + 4.0 * hw * hw * Ns * Cc) * B
            print(f"| {hw}x{hw} | {Cc} | {heads} | {groups} | {orf} | {Ns} | {f:.3f} | {b:.3f} | {B / (f + b) * 1e3:.0f} | "
                  f"{gf / ((f + b) * 1e-3) / 1e12:.1f} |", flush=True)
