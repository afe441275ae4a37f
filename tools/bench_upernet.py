"""BASELINE.json configs[2] / SURVEY.md 8d config 3: UperNet + DAT++ training step on synthetic ADE20K-shaped crops
(512 x 512, 150 classes, ~5 % of the pixels ignored), DistributedDataParallel over the visible GPUs, bf16 autocast,
AdamW with the reference's weight-decay exclusions (parameter names containing `rpe_table` / `norm`,
new_train.py:146-157), loss = CE + 0.4 x auxiliary CE (new_train.py:197-207).  The step is run eagerly (no CUDA
graph): forward, loss, backward (DDP all-reduce overlapped by DDP itself), optimizer step.

    python tools/bench_upernet.py [--model tiny|small|base] [--batch 2] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_upernet.py ...

Prints one JSON line on rank 0.  NOTE (round 1): written after the round's GPU budget was spent - exercised on CPU only
(tests/test_segmentor_host.py covers the model; this driver has not been timed on a B200 yet).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from dat_segmentation_b200.backbone import DAT_BASE_PP, DAT_SMALL_PP, DAT_TINY_PP
from dat_segmentation_b200.segmentor import build_segmentor, segmentation_loss


def param_groups(model, weight_decay):
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if p.requires_grad:
            (no_decay if ("rpe_table" in name or "norm" in name) else decay).append(p)
    return [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="small", choices=["tiny", "small", "base"])
    ap.add_argument("--batch", type=int, default=2, help="per-GPU batch (the reference recipe uses 2)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.pop("NCCL_P2P_DISABLE", None)
        dist.init_process_group("nccl", device_id=dev)
    cfg = {"tiny": DAT_TINY_PP, "small": DAT_SMALL_PP, "base": DAT_BASE_PP}[args.model]
    torch.manual_seed(0)
    model = build_segmentor(cfg).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(param_groups(model, 0.01), lr=6e-5, betas=(0.9, 0.999))
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    imgs = torch.randn(args.batch, 3, 512, 512, device=dev, generator=g)
    masks = torch.randint(0, 150, (args.batch, 512, 512), device=dev, generator=g)
    masks[torch.rand(masks.shape, device=dev, generator=g) < 0.05] = 255

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = ddp(imgs)
        loss = segmentation_loss(tuple(o.float() for o in out) if isinstance(out, tuple) else out.float(), masks)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(3, args.warmup)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item()
        print(json.dumps({"metric": f"UperNet + DAT-{args.model[0].upper()}++ training step images/sec @512x512",
                          "value": round(args.batch * world / (ms * 1e-3), 2), "unit": "images/s", "n_gpus": world,
                          "steps": args.steps, "ms_per_step": round(ms, 3), "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": "BASELINE.json configs[2]", "per_gpu_batch": args.batch,
                                     "params_M": round(n_params / 1e6, 2), "optimizer": "AdamW", "loss": float(loss.detach()),
                                     "cuda_graph": False,
                                     "backbone_cfg": "reference tiny config" if args.model == "tiny" else "assumed (upstream DAT++)"}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
