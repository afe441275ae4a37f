"""Times the gradient all-reduce of bench.py alone (no compute next to it): the 21.05 M-parameter flat fp32 buffer
(84.2 MB) in one call and in bench.py's four buckets, NCCL over NVLink, CUDA events, max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/time_allreduce.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
opts = None
max_ctas = int(os.environ.get("DAT_B200_NCCL_MAX_CTAS", "0"))
if max_ctas > 0:
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = max_ctas
dist.init_process_group("nccl", device_id=dev, pg_options=opts)
N = 21045618
flat = torch.randn(N, device=dev)
buckets = [7544864, 6258280, 6295584, 946890]


def timed(fn, reps=20):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def bucketed():
    off = 0
    for n in buckets:
        dist.all_reduce(flat[off:off + n], op=dist.ReduceOp.AVG)
        off += n


res = {"n_gpus": world, "max_ctas": max_ctas or None,
       "flat_84MB_ms": round(timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG)), 4),
       "four_buckets_ms": round(timed(bucketed), 4),
       "last_bucket_3.8MB_ms": round(timed(lambda: dist.all_reduce(flat[:946890], op=dist.ReduceOp.AVG)), 4)}
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
