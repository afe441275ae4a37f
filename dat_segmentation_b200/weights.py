"""bf16 operand copies of the 1x1-convolution weights, made ONCE per step in ONE launch.

Under bf16 autocast every 1x1 convolution (proj_q / proj_k / proj_v / proj_out of the deformable-attention blocks,
`dat_blocks.py:61-79`; linear1 / linear2 of the MLPs, `:316-348`) multiplies by a bf16 copy of its fp32 weight.
The forward reads that copy as a K-major tensor-core operand and the data gradient dX = dY W reads THE SAME copy as an
MN-major operand (gemm_tc.cu), so one cast per weight per step is all that is needed: `Bf16WeightCache.refresh()`
casts all of them with a single `dat_cast_bf16_multi` launch at the start of `DAT.forward` (captured as one node of
the step's CUDA graph) instead of one cast per layer in the forward plus one cast-transpose per layer in the backward
(146 launches per DAT-T++ step).

A module uses its cached copy only while it is provably current: same parameter storage and same tensor version as
when the copy was made (an optimizer step, `load_state_dict` or `.to()` invalidates it; the module then casts itself).
"""
import ctypes as C

import torch

from . import _cabi

__all__ = ["Bf16WeightCache", "cached_bf16"]


def _key(p):
    return (p.data_ptr(), p._version, p.device)


def cached_bf16(module, name="weight"):
    """The current bf16 copy of `module.<name>` made by a Bf16WeightCache, or None."""
    ent = getattr(module, "_dat_b200_bf16", None)
    if not ent or name not in ent:
        return None
    view, key = ent[name]
    p = getattr(module, name)
    return view if key == _key(p) else None


class Bf16WeightCache:
    """Owns one flat bf16 buffer with a slice per registered (module, parameter name)."""

    def __init__(self, entries):
        self.entries = [(m, n) for m, n in entries]
        self._sig = None
        self._flat = None
        self._table = None
        self._views = []

    def _build(self, dev):
        params = [getattr(m, n) for m, n in self.entries]
        offs, total = [], 0
        for p in params:
            assert p.numel() % 8 == 0, "1x1-conv weights have C_in * C_out elements, a multiple of 8"
            offs.append(total)
            total += p.numel()
        self._flat = torch.empty(total, device=dev, dtype=torch.bfloat16)
        self._views = [self._flat[o:o + p.numel()].view(p.shape[0], -1) for o, p in zip(offs, params)]
        items = (_cabi.CastItem * len(params))()
        for i, (p, v) in enumerate(zip(params, self._views)):
            items[i].src, items[i].dst, items[i].n = p.data_ptr(), v.data_ptr(), p.numel()
        raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).clone()
        self._table = raw.to(dev)
        self._sig = tuple((p.data_ptr(), p.device) for p in params)

    def refresh(self):
        """Cast every registered weight (fp32 parameters on one CUDA device) into the flat buffer: one launch on the
        current stream.  Returns False (and does nothing) when the parameters are not fp32 CUDA tensors."""
        if not self.entries:
            return False
        params = [getattr(m, n) for m, n in self.entries]
        dev = params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 or p.device != dev or not p.is_contiguous() for p in params):
            return False
        if self._sig != tuple((p.data_ptr(), p.device) for p in params):
            self._build(dev)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(_cabi.lib().dat_cast_bf16_multi(C.c_void_p(self._table.data_ptr()), len(params), st),
                        "dat_cast_bf16_multi")
        for (m, n), p, v in zip(self.entries, params, self._views):
            ent = m.__dict__.setdefault("_dat_b200_bf16", {})
            ent[n] = (v, _key(p))
        return True
