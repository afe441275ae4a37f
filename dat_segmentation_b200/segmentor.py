"""Decode heads and the segmentor wrapper around the backbone (BASELINE.json configs[2]:
UperNet + DAT++ training step).

These layers are plain convolution / BatchNorm / ReLU / pooling / bilinear-resize compositions outside the
deformable-attention hot path: they run on library operators and exist so that the full UperNet model of the
reference (`models/heads/uper_head.py`, `models/heads/fcn_head.py`, `models/segmentor.py`, assembled by
`models/builder.py:81-166`) can be built from this package with the same state-dict keys and the same outputs -
checkpoints of the reference load with `strict=True`, `tests/test_segmentor_host.py` compares both on CPU.

    UPerHead    pyramid pooling on the coarsest map, lateral 1x1 + top-down sum + 3x3 smoothing on the others,
                all levels resized to the finest one, concatenated, fused by a 3x3 block, classified per pixel
    FCNHead     auxiliary head: 3x3 block(s) + per-pixel classifier on one feature level
    EncoderDecoder  backbone -> heads -> logits resized to the input; (main, aux) in training mode
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .backbone import DAT_TINY_PP, build_dat

__all__ = ["UPerHead", "FCNHead", "EncoderDecoder", "build_segmentor", "segmentation_loss"]


def _block(cin: int, cout: int, k: int) -> nn.Sequential:
    """conv (no bias) -> BatchNorm -> ReLU, the unit every head layer is made of (keys `0.*`, `1.*`)."""
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class UPerHead(nn.Module):
    """uper_head.py:8-131.  Parameter names: ppm_modules.{i}.{1,2}, ppm_bottleneck, lateral_convs.{i}, fpn_convs.{i},
    fuse_bottleneck, cls_seg."""

    def __init__(self, in_channels: Sequence[int], num_classes: int, *, channels: int = 512,
                 pool_scales: Sequence[int] = (1, 2, 3, 6), dropout_ratio: float = 0.1, align_corners: bool = False):
        super().__init__()
        self.align_corners = align_corners
        self.num_levels = len(in_channels)
        top = in_channels[-1]
        branch = channels // len(pool_scales)
        self.ppm_modules = nn.ModuleList(
            nn.Sequential(nn.AdaptiveAvgPool2d(s), *_block(top, branch, 1)) for s in pool_scales)
        self.ppm_bottleneck = _block(top + len(pool_scales) * branch, channels, 3)
        self.lateral_convs, self.fpn_convs = nn.ModuleList(), nn.ModuleList()
        for c in in_channels[:-1]:        # built level by level: same parameter-initialisation RNG stream as the reference
            self.lateral_convs.append(_block(c, channels, 1))
            self.fpn_convs.append(_block(channels, channels, 3))
        self.fuse_bottleneck = _block(self.num_levels * channels, channels, 3)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else nn.Identity()
        self.cls_seg = nn.Conv2d(channels, num_classes, 1)

    def _resize(self, t, size):
        return F.interpolate(t, size=size, mode="bilinear", align_corners=self.align_corners)

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        if len(feats) != self.num_levels:
            raise ValueError(f"expected {self.num_levels} feature maps, got {len(feats)}")
        coarse = feats[-1]
        pooled = [coarse] + [self._resize(m(coarse), coarse.shape[2:]) for m in self.ppm_modules]
        levels = [lat(f) for lat, f in zip(self.lateral_convs, feats)] + [self.ppm_bottleneck(torch.cat(pooled, 1))]
        for i in range(self.num_levels - 1, 0, -1):          # top-down: add the coarser level, then smooth
            levels[i - 1] = self.fpn_convs[i - 1](levels[i - 1] + self._resize(levels[i], levels[i - 1].shape[2:]))
        size = levels[0].shape[2:]
        fused = torch.cat([levels[0]] + [self._resize(t, size) for t in levels[1:]], 1)
        return self.cls_seg(self.dropout(self.fuse_bottleneck(fused)))


class FCNHead(nn.Module):
    """fcn_head.py:8-49.  Parameter names: convs.{3j, 3j+1}, cls_seg."""

    def __init__(self, in_channels: int, num_classes: int, *, channels: int = 256, num_convs: int = 1,
                 dropout_ratio: float = 0.1, align_corners: bool = False):
        super().__init__()
        self.align_corners = align_corners
        layers, cin = [], in_channels
        for _ in range(num_convs):
            layers += list(_block(cin, channels, 3))
            cin = channels
        self.convs = nn.Sequential(*layers)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else nn.Identity()
        self.cls_seg = nn.Conv2d(channels, num_classes, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.cls_seg(self.dropout(self.convs(x)))


class EncoderDecoder(nn.Module):
    """segmentor.py:10-63: logits at input resolution; `(main, aux)` when training with an auxiliary head (which reads
    the second-coarsest level), the auxiliary logits otherwise kept in `last_aux_logits`."""

    def __init__(self, backbone: nn.Module, decode_head: nn.Module, auxiliary_head: Optional[nn.Module] = None, *,
                 align_corners: bool = False):
        super().__init__()
        self.backbone, self.decode_head, self.auxiliary_head = backbone, decode_head, auxiliary_head
        self.align_corners = align_corners

    def forward(self, x: torch.Tensor):
        feats = self.backbone(x)
        up = lambda t: F.interpolate(t, size=x.shape[2:], mode="bilinear", align_corners=self.align_corners)
        logits = up(self.decode_head(feats))
        if self.auxiliary_head is None:
            return logits
        aux = up(self.auxiliary_head(feats[-2]))
        self.last_aux_logits = aux if self.training else aux.detach()
        return (logits, aux) if self.training else logits


def build_segmentor(backbone_cfg: dict = None, num_classes: int = 150, with_aux: bool = True, **backbone_kw) -> EncoderDecoder:
    """UperNet over a DAT++ backbone as `build_model_from_config` assembles it for `configs/dat/upn_*.py`
    (decode head on all four levels, FCN auxiliary head on level 2)."""
    cfg = dict(DAT_TINY_PP if backbone_cfg is None else backbone_cfg)
    dims = list(cfg["dims"])
    backbone = build_dat(cfg, **backbone_kw)
    aux = FCNHead(dims[2], num_classes) if with_aux else None
    return EncoderDecoder(backbone, UPerHead(dims, num_classes), aux)


def segmentation_loss(outputs, masks: torch.Tensor, aux_weight: float = 0.4, ignore_index: int = 255) -> torch.Tensor:
    """Cross entropy on the main logits plus `aux_weight` times the auxiliary one (new_train.py:197-207)."""
    if isinstance(outputs, tuple):
        main, aux = outputs
        return F.cross_entropy(main, masks, ignore_index=ignore_index) + \
            aux_weight * F.cross_entropy(aux, masks, ignore_index=ignore_index)
    return F.cross_entropy(outputs, masks, ignore_index=ignore_index)
