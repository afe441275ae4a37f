#!/usr/bin/env python
"""Headline benchmark: DAT-T++ backbone fwd+bwd images/s @512x512 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: the backbone of dat_segmentation_b200/backbone.py with the 14 deformable-attention
blocks running in the hand-written sm_100a kernels (C ABI, libdat_b200.so), bf16 autocast,
batch 16 per GPU, synthetic ADE20K-shaped input, random-init weights.  `value` = device-timed
images/s with inputs resident in HBM; `e2e` = the same step fed from pinned host memory with a
device->host read of the loss every step.  `roofline` = the dominant hand-written kernel timed
alone with CUDA events (L2 flushed between launches).  `cpu_baseline` / `--impl reference` =
the oracle port of the reference (library-operator form) on the host cores, bounded sample.
Under torchrun (N > 1): one process per GPU, batch-sharded, the flat gradient buffer all-reduced
by one NCCL call per step, max-over-ranks timing.  The fwd+bwd of a step is a CUDA-graph replay.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DAT-T++ backbone fwd+bwd images/sec @512x512"
UNIT = "images/s"
IMG = 512
WORKLOAD = ("DAT-T++ backbone fwd+bwd, 512x512, batch 16 per GPU, bf16 autocast "
            "(BASELINE.json configs[1])")
PER_GPU_BATCH = 16
STAGES = [  # DAT-T++ @512²: (H=W, C, heads, groups, stride, ksize, q_size, n_blocks)
    (128, 64, 2, 1, 8, 9, 56, 1), (64, 128, 4, 2, 4, 7, 28, 2),
    (32, 256, 8, 4, 2, 5, 14, 9), (16, 512, 16, 8, 1, 3, 7, 2)]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.15)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(s[0])) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.samples[0][1])), "reasons": reasons,
                "samples": len(sm)}


def loss_of(outs):
    # loss = sum of the means of the 4 backbone outputs
    return sum(o.float().mean() for o in outs)


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference on the host cores
# --------------------------------------------------------------------------------------

def cpu_port_run(steps, warmup, batch):
    from dat_segmentation_b200.backbone import build_dat
    from oracle.dattn_oracle import OracleDAttention
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = build_dat(attn_cls=OracleDAttention).train()
    imgs = torch.randn(batch, 3, IMG, IMG)

    def step():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            outs = model(imgs)
        loss_of(outs).backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=batch / dt, ms_per_step=dt * 1e3, cores=cores, threads=torch.get_num_threads(),
                sample=f"{steps} steps of batch {batch} @512x512 fwd+bwd, bf16 autocast, {warmup} warm-up")


def gpu_library_run(dev, steps, warmup):
    """Informational GPU yardstick: the oracle port of the backbone - the reference's own operator
    sequence on library kernels (cuDNN convolutions, cuBLAS bmm, ATen grid_sample / softmax / LayerNorm) - on the same
    B200, same batch, bf16 autocast, fwd+bwd, eager.  Not the product path and not the reference arm."""
    from dat_segmentation_b200.backbone import build_dat
    from oracle.dattn_oracle import OracleDAttention
    torch.manual_seed(0)
    model = build_dat(attn_cls=OracleDAttention).to(dev).train()
    imgs = torch.randn(PER_GPU_BATCH, 3, IMG, IMG, device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model(imgs)
        loss_of(outs).backward()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    return {"value": round(PER_GPU_BATCH / (ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(ms, 3),
            "what": "oracle-port backbone on library kernels (cuDNN / cuBLAS / ATen), same GPU, batch 16, bf16 autocast, "
                    "eager fwd+bwd"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_run(max(1, min(args.steps, 4)), max(1, min(args.warmup, 1)), 2)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": PER_GPU_BATCH,
                   "implementation": "oracle port of the backbone (library operators) on the host cores; each step "
                                     "is a bounded sample of the workload: batch 2 instead of 16",
                   "drop_path_rate": 0.3},
        "cpu_baseline": {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# roofline leg: the dominant hand-written kernel, timed alone on its launch stream
# --------------------------------------------------------------------------------------

def _time_launch(dev, launch, flush, reps=10):
    """Median CUDA-event time (ms) of one launch on torch's current stream, L2 flushed before every launch."""
    st = torch.cuda.current_stream(dev)
    for _ in range(3):
        launch()
    times = []
    for _ in range(reps):
        flush.zero_()                       # > L2 (126 MB): next launch starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        launch()
        e1.record(st)
        torch.cuda.synchronize(dev)
        times.append(e0.elapsed_time(e1))
    return sorted(times)[len(times) // 2]


def ncu_traffic():
    """DRAM traffic per launch of the two roofline kernels, from the committed ncu summary of THIS command line
    (`python bench.py --roofline-only` under `ncu --set full`, i.e. the same launches with the same L2 flush before each
    of them: profiles/r02_ncu_traffic.json, written by tools/summarize_ncu_traffic.py).  None when the file is absent."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return {}
    return json.load(open(path))


def kernel_roofline(dev, pk):
    """Roofline of the dominant kernel of the step, `gemm_tc_persistent_kernel` (13 % of the kernel time, 183
    launches: profiles/r02_launches_step.md), at the shape that carries most of its time - the stage-2 MLP fc1
    (M = B*HW = 16384, N = 1024, K = 256, bf16; 18 MLPs per step run it forward and as the fc2 data gradient).
    Algorithmic bytes = X + W + Y once; its arithmetic intensity (202 FLOP/B) is below the measured
    ridge (1658 TF/s / 6541 GB/s = 254 FLOP/B), so the bound is HBM.  `others`: the attention forward (the
    kernel north_star asks tensor utilisation for) timed the same way."""
    from dat_segmentation_b200 import _cabi
    lib = _cabi.lib()
    traffic = ncu_traffic()
    bf = torch.bfloat16
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    g = torch.Generator(device=dev).manual_seed(0)
    # ---- dominant kernel: persistent tcgen05 GEMM, stage-2 fc1 shape ----
    M, N, K = PER_GPU_BATCH * 32 * 32, 1024, 256
    x = torch.randn(M, K, device=dev, generator=g).to(bf)
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(bf)
    bias = torch.randn(N, device=dev, generator=g)
    y = torch.empty(M, N, device=dev, dtype=bf)
    ms = _time_launch(dev, lambda: _cabi.check(lib.dat_pointwise_fwd_tc(p(x), 1, p(w), p(bias), p(y), 1, M, N, K, sp),
                                                "gemm"), flush)
    gbytes = (M * K + N * K + M * N) * 2 + N * 4
    gflops = 2.0 * M * N * K
    roof = {"kernel": "gemm_tc_persistent_kernel (stage-2 MLP fc1: M=16384 N=1024 K=256, bf16)", "bound": "hbm",
            "achieved": round(gbytes / (ms * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": round(gbytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu summary of
            # `bench.py --roofline-only` (same L2 flush before the launch); null when no summary is committed
            "traffic": traffic.get("gemm", {}).get("dram_bytes"), "traffic_source": traffic.get("gemm", {}).get("source"),
            "ms": round(ms, 4), "algorithmic_bytes": gbytes,
            "tflops_at_this_time": round(gflops / (ms * 1e-3) / 1e12, 1), "tensor_frac": round(gflops / (ms * 1e-3) / 1e12 / pk["bf16_tflops"], 4),
            "peak_source": pk["source"]}
    del x, w, y
    # ---- attention forward, stage-2 shape ----
    H, Cc, heads, groups, stride, ksize, qs, _ = STAGES[2]
    B, HW, Ns = PER_GPU_BATCH, H * H, 256
    d = _cabi.BlockDesc(B, H, H, heads, groups, stride, ksize, 2 * qs - 1, 2 * qs - 1, -1.0, _cabi.DAT_F32, _cabi.DAT_BF16)
    q = torch.randn(B, HW, Cc, device=dev, generator=g).to(bf)
    k = torch.randn(B, Ns, Cc, device=dev, generator=g).to(bf)
    v = torch.randn(B, Ns, Cc, device=dev, generator=g).to(bf)
    pos = (torch.rand(B, groups, Ns, 2, device=dev, generator=g) * 2 - 1)
    tab = torch.randn(heads, 2 * qs - 1, 2 * qs - 1, device=dev, generator=g) * 0.1
    o = torch.empty_like(q)
    lse = torch.empty(B, heads, HW, device=dev)
    nws = lib.dat_attention_fwd_workspace_bytes(C.byref(d))
    ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
    ms_a = _time_launch(dev, lambda: _cabi.check(lib.dat_attention_fwd(C.byref(d), p(q), p(k), p(v), p(pos), p(tab), p(o),
                                                                        p(lse), p(ws), nws, 0, sp), "attention_fwd"), flush)
    flops = 4.0 * HW * Ns * Cc * B
    byts = (2 * B * HW * Cc + 2 * B * Ns * Cc) * 2 + B * groups * Ns * 8 + heads * (2 * qs - 1) ** 2 * 4
    roof["others"] = [{
        "kernel": "attn_fwd_tc_kernel (stage-2 shape, B=16, bf16)", "bound": "tensor",
        "achieved": round(flops / (ms_a * 1e-3) / 1e12, 2), "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
        "frac": round(flops / (ms_a * 1e-3) / 1e12 / pk["bf16_tflops"], 5),
        "traffic": traffic.get("attention", {}).get("dram_bytes"), "traffic_source": traffic.get("attention", {}).get("source"),
        "hbm_frac": round(byts / (ms_a * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "ms": round(ms_a, 4),
        "algorithmic_bytes": byts,
        "note": "attn_fwd_tc2_kernel; bound by CUDA-core issue rate + MUFU (bias interpolation + softmax, 23.7 thread "
                "instructions per score vs 128 MMA FLOP): DESIGN.md 3.1 row 7, profiles/r02_ncu_kernels.md"}]
    roof["per_kernel_table"] = "profiles/r02_kernel_rooflines.md"
    return roof


# --------------------------------------------------------------------------------------

class TrainStep:
    """The benchmarked step: DAT-T++ backbone fwd+bwd at 512x512, bf16 autocast, train mode, the whole fwd+bwd
    captured once in a CUDA graph on a high-priority stream (weight- / table-gradient branches on side streams) and
    replayed.  tests/test_bench_step.py drives exactly this object (graph replay vs serial eager, determinism,
    oracle parity), so what is timed is what is tested."""

    def __init__(self, dev, world=1, rank=0, batch=PER_GPU_BATCH, img=IMG, graph=True, cfg=None, seed=0,
                 fixed_drop_path=None, buckets=None):
        from dat_segmentation_b200 import _cabi
        from dat_segmentation_b200.backbone import build_dat
        self.dev, self.world, self.rank, self.batch = dev, world, rank, batch
        self.lib = _cabi.lib()
        torch.manual_seed(seed)
        self.model = build_dat(cfg).to(dev).train()          # drop_path_rate 0.3 as in the shipped config
        if fixed_drop_path is not None:                      # tests: the same stochastic-depth masks in every run
            g = torch.Generator(device="cpu").manual_seed(fixed_drop_path)
            for st in self.model.stages:
                st.fix_drop_path_scales(batch, g, dev)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        # N > 1: the step's gradients are packed into flat fp32 buckets (multi-tensor copies inside the graph) that
        # are all-reduced over NCCL; see `step()`
        self.flat, self.bucket_slices, self.overlap = None, [], False
        if world > 1:
            self._make_buckets(buckets)
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        self.imgs = torch.randn(batch, 3, img, img, device=dev, generator=gen)   # static step input
        self.graph, self.static_loss = None, None
        self.use_graph = graph
        self.launches_per_step = 0

    # Gradient buckets in backward order (stage 3 first): bucket i is all-reduced on a communication stream while
    # the backward of the earlier stages still runs.  Stage 2 (18 of the 28 blocks) is split in two.
    @staticmethod
    def _bucket_key(name):
        import re
        m = re.match(r"(stages|norms)\.(\d)\.", name)
        if m:
            i = int(m.group(2))
            if i == 2 and m.group(1) == "stages":
                d = re.search(r"\.(\d+)\.", name[len("stages.2."):] + ".")
                blk = int(d.group(1)) if d else 0
                # layer_norms / layer_scales are indexed by 2d, 2d+1; every other list by d
                if ".layer_norms." in name or ".layer_scales." in name:
                    blk //= 2
                return (2, 1 if blk >= 9 else 0)
            return (i, 0) if i >= 2 else (1, 0)
        m = re.match(r"down_projs\.(\d)\.", name)
        if m:   # down_projs[i] feeds stage i + 1: its gradient arrives after that stage's backward
            i = int(m.group(1)) + 1
            return (i, 0) if i >= 2 else (1, 0)
        return (1, 0)    # stem, stages 0-1: the tail of the backward

    def _make_buckets(self, buckets):
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.dev)
        groups = {}
        for name, p in self.model.named_parameters():
            if p.requires_grad:
                groups.setdefault(self._bucket_key(name), []).append(p)
        order = sorted(groups, reverse=True)                 # produced first in the backward -> first in the buffer
        off = 0
        self.bucket_params, self.bucket_views = [], []
        for key in order:
            start, views = off, []
            for p in groups[key]:
                views.append(self.flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.bucket_slices.append((start, off))
            self.bucket_params.append(groups[key])
            self.bucket_views.append(views)
        assert off == n
        self.overlap = os.environ.get("DAT_B200_BENCH_OVERLAP", "1") != "0"
        if self.overlap:
            from dat_segmentation_b200 import _streams
            self.comm = torch.cuda.Stream(self.dev)
            self._pending = [0] * len(order)
            for bi, plist in enumerate(self.bucket_params):
                for p in plist:
                    p.register_post_accumulate_grad_hook(lambda _p, bi=bi: self._grad_ready(bi))
            # the hooks below order the communication stream after the weight-gradient side stream themselves
            _streams.HOOKS_SYNC_THEMSELVES[0] = True

    def _grad_ready(self, bi):
        """Post-accumulate-grad hook: when the last gradient of bucket `bi` exists, pack the bucket and all-reduce it
        on the communication stream (ordered after the current stream and the weight-gradient side stream).  Under
        graph capture these become parallel branches of the step's graph."""
        import torch.distributed as dist
        from dat_segmentation_b200 import _streams
        self._pending[bi] += 1
        if self._pending[bi] < len(self.bucket_params[bi]):
            return
        self._pending[bi] = 0
        cur = torch.cuda.current_stream(self.dev)
        self.comm.wait_stream(cur)
        side = _streams.existing_side_stream(self.dev)
        if side is not None:
            self.comm.wait_stream(side)
        with torch.cuda.stream(self.comm):
            torch._foreach_copy_(self.bucket_views[bi], [p.grad for p in self.bucket_params[bi]])
            lo, hi = self.bucket_slices[bi]
            if not os.environ.get("DAT_B200_BENCH_SKIP_AR"):      # debug: cost of the packing / stream structure alone
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.AVG)

    def fwd_bwd(self):
        for p in self.params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = self.model(self.imgs)
        loss = loss_of(outs)
        loss.backward()
        if self.world > 1:
            if self.overlap:
                torch.cuda.current_stream(self.dev).wait_stream(self.comm)   # join: flat holds the averaged gradients
            else:
                for views, plist in zip(self.bucket_views, self.bucket_params):
                    torch._foreach_copy_(views, [p.grad for p in plist])
        return loss

    def warm_and_capture(self, warmup):
        dev = self.dev
        n0 = self.lib.dat_launch_count()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.fwd_bwd()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.launches_per_step = (self.lib.dat_launch_count() - n0) // max(warmup, 1)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            # The critical chain is captured on a high-priority stream; the weight-gradient / table-gradient branches
            # run on default-priority side streams (library-owned and _streams.py), so when both have a kernel ready
            # the block scheduler serves the critical chain first (kernel nodes keep their stream's priority).
            prio = os.environ.get("DAT_B200_BENCH_PRIORITY", "-1")
            cap_stream = torch.cuda.Stream(dev, priority=int(prio)) if prio != "none" else None
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self.static_loss = self.fwd_bwd()

    def step(self):
        import torch.distributed as dist
        loss = self.static_loss
        if self.graph is not None:
            self.graph.replay()
        else:
            loss = self.fwd_bwd()
        if self.world > 1 and not self.overlap:   # the one exchange step of data-parallel training (new_train.py:116)
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        return loss

    def release(self):
        """Destroy the captured graph before the process group (a graph that holds NCCL work keeps the
        communicator busy at teardown)."""
        self.graph = None
        self.static_loss = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dat_b200", choices=["dat_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--roofline-only", action="store_true", help="only the roofline leg (the command ncu wraps)")
    ap.add_argument("--gpu-library-baseline", action="store_true",
                    help="also time the oracle-port backbone (library operators: cuDNN / cuBLAS / ATen) on the same GPU, "
                         "same step, eager - informational yardstick, adds a `gpu_library_baseline` key")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.roofline_only:
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        print(json.dumps(kernel_roofline(dev, peaks())), flush=True)
        return

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.pop("NCCL_P2P_DISABLE", None)   # the reference's launch scripts set it; NVSwitch wants P2P
        # The bucketed all-reduces run next to the backward (TrainStep._grad_ready): cap the CTAs NCCL may take from
        # it.  NVSwitch cost is latency- not link-bound at these sizes, a handful of CTAs carries a bucket.
        pg_opts = None
        max_ctas = int(os.environ.get("DAT_B200_NCCL_MAX_CTAS", "8"))
        if max_ctas > 0 and hasattr(dist, "ProcessGroupNCCL"):
            pg_opts = dist.ProcessGroupNCCL.Options()
            pg_opts.config.max_ctas = max_ctas
        dist.init_process_group("nccl", device_id=dev, pg_options=pg_opts)
        # this script joins the weight-gradient side stream itself (end of backward()) before it packs and all-reduces
        # the gradients, so the side stream stays on although a process group exists (no DDP hooks here)
        from dat_segmentation_b200 import _streams
        _streams.ALLOW_WITH_PROCESS_GROUP[0] = True
    warmup = max(3, args.warmup)

    ts = TrainStep(dev, world, rank, graph=not args.no_graph)
    imgs = ts.imgs
    host = torch.randn(PER_GPU_BATCH, 3, IMG, IMG).pin_memory()
    ts.warm_and_capture(warmup)
    launches_per_step = ts.launches_per_step
    graph = ts.graph
    step = ts.step

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
    for _ in range(warmup):
        step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    launches = launches_per_step * args.steps
    ms = e0.elapsed_time(e1) / args.steps

    # end-to-end: every step copies its batch from pinned host memory and reads the loss back.  The
    # input pipeline is the usual prefetching loader: the H2D copy of batch i + 1 runs on a copy
    # stream into a staging buffer while step i computes; a device-side copy moves it into the
    # graph's static input at the start of the step.  (All copies are inside the timed region.)
    copy_stream = torch.cuda.Stream(dev)
    staging = torch.empty_like(imgs)

    def prefetch():
        with torch.cuda.stream(copy_stream):
            staging.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    sync_all()
    t0 = time.perf_counter()
    ready = prefetch()
    for i in range(args.steps):
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready)
        imgs.copy_(staging, non_blocking=True)
        consumed = torch.cuda.Event()
        consumed.record(cur)
        copy_stream.wait_event(consumed)          # staging may be overwritten once it has been consumed
        if i + 1 < args.steps:
            ready = prefetch()
        _ = step().item()
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.summary() if sampler else None

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        pk = peaks()
        roof = kernel_roofline(dev, pk)
        total = PER_GPU_BATCH * world
        line = {
            "metric": METRIC, "value": round(total / (ms * 1e-3), 2), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "implementation": "deformable-attention blocks, LayerNorms, residual/drop-path, MLP 1x1 convs and "
                                         "depthwise convs, conv stem and down-projections (im2col + tcgen05 GEMMs) in dat_b200 kernels",
                       "per_gpu_batch": PER_GPU_BATCH, "global_batch": total,
                       "parallelism": (f"dp{world} (batch-sharded; NCCL gradient all-reduce in {len(ts.bucket_slices)} buckets "
                                       f"{'overlapped with the backward inside the captured step' if ts.overlap else 'after the step'})"
                                       if world > 1 else "single GPU"),
                       "l2": "working set per step >> 126 MB L2 (no flush needed); roofline leg flushes L2 per launch",
                       "cuda_graph": graph is not None,
                       "drop_path_rate": 0.3},
            "e2e": {"value": round(total / (e2e_ms * 1e-3), 2), "unit": UNIT,
                    "h2d_bytes_per_step": host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        line["roofline_attention"] = roof["others"][0]      # the same entry at top level (drivers that flatten `roofline`)
        if args.gpu_library_baseline and world == 1:
            line["gpu_library_baseline"] = gpu_library_run(dev, max(3, args.steps), 3)
        if not args.no_cpu_baseline and world == 1:
            r = cpu_port_run(3, 1, 2)
            line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": UNIT, "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    ts.release()
    del graph, step
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
