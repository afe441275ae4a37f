"""One-call swap of the reference's deformable-attention block for the dat_b200 one inside an (unmodified) reference
tree that is importable as `models.*`.

    import dat_segmentation_b200.install as b200
    b200.install()                       # before or after `models.backbones.dat` has been imported
    model = build_model_from_config("configs/dat/upn_tiny_160k_dp03_lr6.py")

`models/backbones/dat.py:18` star-imports the block, so the class object is replaced both in
`models.utils.dat_blocks` and - when it is already loaded - in `models.backbones.dat`; everything else of the reference
(`DAT`, `TransformerStage`, the heads, `models/builder.py:93-102`, `load_checkpoint`) is used as it is.  If mmsegmentation
is installed, the reference backbone class is also registered in its `BACKBONES` registry under the name `DAT`, which is
what the mmseg-style configs (`type='DAT'`) resolve through `tools/train.py`.  `uninstall()` restores the originals.
"""
import importlib
import sys

from .dattention import DAttentionBaseline as _B200Block

_SAVED = {}


def install(register_mmseg: bool = True):
    """Returns the list of module names whose `DAttentionBaseline` now is the dat_b200 class."""
    patched = []
    blocks = importlib.import_module("models.utils.dat_blocks")
    for name in ("models.utils.dat_blocks", "models.backbones.dat"):
        mod = blocks if name == "models.utils.dat_blocks" else sys.modules.get(name)
        if mod is None or not hasattr(mod, "DAttentionBaseline"):
            continue
        if mod.DAttentionBaseline is not _B200Block:
            _SAVED.setdefault(name, mod.DAttentionBaseline)
            mod.DAttentionBaseline = _B200Block
        patched.append(name)
    if register_mmseg:
        try:
            from mmseg.models.builder import BACKBONES          # optional dependency of the reference's tools/train.py
            dat_cls = importlib.import_module("models.backbones.dat").DAT
            if "DAT" not in BACKBONES.module_dict:
                BACKBONES.register_module(name="DAT", module=dat_cls)
            patched.append("mmseg.BACKBONES[DAT]")
        except ImportError:
            pass
    return patched


def uninstall():
    for name, cls in _SAVED.items():
        mod = sys.modules.get(name)
        if mod is not None:
            mod.DAttentionBaseline = cls
    _SAVED.clear()
