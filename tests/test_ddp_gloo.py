"""N > 1 host logic on CPU: world size 2, gloo.  The backbone (with the oracle port of the
block, since the CUDA kernels need a GPU) is batch-sharded under DistributedDataParallel;
the all-reduced gradients must equal the single-process full-batch gradients — i.e. the path
shards by batch with no other exchange step (new_train.py:116)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dat_segmentation_b200.backbone import build_dat
from oracle.dattn_oracle import OracleDAttention

SMALL = dict(dim_stem=32, dims=[32, 64, 128, 256], depths=[1, 1, 1, 1],   # dims must double (dat.py:227)
             stage_spec=[["D"], ["D"], ["D"], ["D"]], heads=[1, 2, 4, 8], groups=[1, 1, 2, 4],
             use_pes=[True] * 4, strides=[4, 2, 1, 1], offset_range_factor=[-1, 2, -1, 1],
             use_dwc_mlps=[True] * 4, use_lpus=[True] * 4, use_conv_patches=True,
             ksizes=[5, 3, 3, 3], drop_path_rate=0.0, use_checkpoint=False, img_size=64)


def _model():
    torch.manual_seed(11)
    return build_dat(SMALL, attn_cls=OracleDAttention)


def _loss(outs):
    return sum(o.square().mean() for o in outs)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, x, ref_grads, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = torch.nn.parallel.DistributedDataParallel(_model())
        shard = x[rank::world]                       # sample b -> rank b mod world
        _loss(model(shard)).backward()               # DDP averages gradients over ranks
        worst = 0.0
        for (name, p), g in zip(model.module.named_parameters(), ref_grads):
            assert p.grad is not None, name          # find_unused_parameters=False: every param gets a grad
            if g.abs().max() < 1e-6:                 # proj_k.bias: analytically zero gradient
                continue
            worst = max(worst, ((p.grad - g).abs().max() / g.abs().max()).item())
        ok[rank] = worst
    finally:
        dist.destroy_process_group()


def test_batch_sharded_gradients_match_single_process():
    world = 2
    x = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(5))
    ref = _model()
    # mean over ranks of per-shard mean losses == full-batch mean loss for equal shards
    _loss(ref(x)).backward()
    ref_grads = [p.grad.clone() for p in ref.parameters()]
    ok = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), x, ref_grads, ok), nprocs=world, join=True)
    assert len(ok) == world
    assert max(ok.values()) < 2e-3, dict(ok)     # fp32 summation order differs between 2x2 and 1x4


def _serial_flags(rank, world, port, out):
    from dat_segmentation_b200 import _streams
    before = _streams.serial()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with_pg = _streams.serial()
    _streams.ALLOW_WITH_PROCESS_GROUP[0] = True
    opted_in = _streams.serial()
    dist.destroy_process_group()
    out.put((before, with_pg, opted_in))


def test_side_stream_is_off_under_a_process_group_unless_opted_in():
    """DistributedDataParallel's hooks read weight gradients before the side stream's end-of-backward join: with an
    initialised process group the 1x1-conv weight gradients stay on the current stream unless the caller (bench.py)
    opts in."""
    os.environ.pop("DAT_B200_SERIAL_WGRAD", None)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_serial_flags, args=(0, 1, port, q))
    p.start()
    before, with_pg, opted_in = q.get(timeout=120)
    p.join(timeout=60)
    assert (before, with_pg, opted_in) == (False, True, False)
