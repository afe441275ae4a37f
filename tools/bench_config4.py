"""BASELINE.json configs[3] / SURVEY.md 8d config 4: DAT-B++ backbone inference at the sliding-window / multi-scale
evaluation shapes (2048 x 512 and its ratios {0.5, 0.75, 1, 1.25, 1.5, 1.75}, `tools/test.py:143-148` of the reference;
sizes rounded to multiples of 32 as the reference's resize does), batch-sharded over the visible GPUs: every rank runs
its own images, NO collective on the data path (the only communication is the max over ranks of the timings).  The
SemanticFPN neck / head does not exist in the reference tree, so the backbone is what is timed; the DAT-B++
hyper-parameters are the upstream ones ("assumed, unpinned by the reference").  bf16 autocast, eval mode, no_grad.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config4.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from dat_segmentation_b200.backbone import DAT_BASE_PP, DAT_SMALL_PP, DAT_TINY_PP, build_dat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="base", choices=["tiny", "small", "base"])
    ap.add_argument("--batch", type=int, default=4, help="images per GPU per forward")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.pop("NCCL_P2P_DISABLE", None)
        dist.init_process_group("nccl", device_id=dev)
    cfg = {"tiny": DAT_TINY_PP, "small": DAT_SMALL_PP, "base": DAT_BASE_PP}[args.model]
    torch.manual_seed(0)
    model = build_dat(cfg).to(dev).eval()
    rows = []
    for ratio in (0.5, 0.75, 1.0, 1.25, 1.5, 1.75):
        w, h = int(2048 * ratio + 16) // 32 * 32, int(512 * ratio + 16) // 32 * 32
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        imgs = torch.randn(args.batch, 3, h, w, device=dev, generator=g)

        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return model(imgs)

        for _ in range(2):
            infer()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            infer()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append({"ratio": ratio, "size_hw": [h, w], "ms_per_forward": round(t.item(), 3),
                     "images_per_s": round(args.batch * world / (t.item() * 1e-3), 1)})
        del imgs
    if rank == 0:
        print(json.dumps({"metric": f"DAT-{args.model[0].upper()}++ backbone inference images/sec, 2048x512 multi-scale, batch-sharded",
                          "n_gpus": world, "per_gpu_batch": args.batch, "dtype": "bf16", "data": "synthetic",
                          "collective": "none on the data path", "scales": rows,
                          "config": "BASELINE.json configs[3]; DAT-B++ hyper-parameters assumed (upstream DAT++)"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
