"""TEST INFRASTRUCTURE — not product code.

Imports the UNMODIFIED reference (`/root/reference`) in the build container so
that the oracle restatement can be pinned against it and golden vectors can be
generated (`tests/golden/make_golden.py`).  `/root/reference` does not exist on
the GPU box: nothing under `-m gpu`, `smoke()` or `bench.py` may call this.

The reference imports `timm.models.layers` (dat_blocks.py:17, dat.py:17) which is
not installed; the three names it uses are provided by an in-memory stub
(SURVEY.md §8c).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("DAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "utils", "dat_blocks.py"))


class _DropPath(torch.nn.Module):
    """timm.DropPath behaviour (scale_by_keep=True); identity in eval / p == 0."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


def _install_timm_stub():
    if "timm.models.layers" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    layers.DropPath = _DropPath
    timm.models = models
    models.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.models"] = models
    sys.modules["timm.models.layers"] = layers


def import_reference():
    """Returns (dat_blocks module, dat module) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_timm_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    blocks = importlib.import_module("models.utils.dat_blocks")
    dat = importlib.import_module("models.backbones.dat")
    return blocks, dat
