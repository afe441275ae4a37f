"""BASELINE.json configs[2] / configs[3] on one GPU: the DAT-T++ / DAT-S++ / DAT-B++ backbones (S++ / B++
hyper-parameters assumed from upstream DAT++, not pinned by the reference tree) —
  train:  fwd+bwd at 512x512, bf16 autocast, train mode, CUDA-graph replay (as bench.py), images/s
  infer:  forward only at 512x2048 (the sliding-window / multi-scale evaluation shape of configs[3]), no_grad
Prints a markdown table.   usage: python tools/bench_family.py > profiles/rNN_family.md"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from dat_segmentation_b200.backbone import DAT_BASE_PP, DAT_SMALL_PP, DAT_TINY_PP, build_dat

dev = torch.device("cuda", 0)


def timed(fn, steps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


print("# DAT++ family on one B200 (dat_b200 kernels; conv stem / down-projections in library ops)\n")
print("| backbone | params (M) | train 512x512 b16 fwd+bwd ms | images/s | infer 512x2048 b4 fwd ms | images/s |\n|---|---:|---:|---:|---:|---:|")
for name, cfg in (("DAT-T++", DAT_TINY_PP), ("DAT-S++ (assumed cfg)", DAT_SMALL_PP), ("DAT-B++ (assumed cfg)", DAT_BASE_PP)):
    torch.manual_seed(0)
    model = build_dat(cfg).to(dev).train()
    nparam = sum(p.numel() for p in model.parameters()) / 1e6
    params = [p for p in model.parameters() if p.requires_grad]
    imgs = torch.randn(16, 3, 512, 512, device=dev)

    def fwd_bwd():
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(imgs)
        loss = sum(o.float().mean() for o in outs)
        loss.backward()
        return loss

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            fwd_bwd()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fwd_bwd()
    t_train = timed(graph.replay)
    del graph
    model.eval()
    wide = torch.randn(4, 3, 512, 2048, device=dev)

    def infer():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(wide)

    t_inf = timed(infer)
    print(f"| {name} | {nparam:.1f} | {t_train:.2f} | {16 / t_train * 1e3:.0f} | {t_inf:.2f} | {4 / t_inf * 1e3:.0f} |", flush=True)
    del model, params, imgs, wide
    torch.cuda.empty_cache()
