"""Heads + segmentor wrapper (dat_segmentation_b200/segmentor.py) against the unmodified reference
(`models/heads/*.py`, `models/segmentor.py`) on CPU - container only (needs /root/reference) - plus a reference-free
shape / loss check."""
import importlib

import pytest
import torch

from oracle import dattn_oracle as orc
from oracle.ref_shim import import_reference, reference_available

from dat_segmentation_b200.segmentor import EncoderDecoder, FCNHead, UPerHead, build_segmentor, segmentation_loss

SMALL = dict(dim_stem=32, dims=[32, 64, 128, 256], depths=[1, 1, 1, 1], stage_spec=[["D"], ["D"], ["D"], ["D"]],
             heads=[1, 2, 4, 8], groups=[1, 1, 2, 4], use_pes=[True] * 4, strides=[4, 2, 1, 1],
             offset_range_factor=[-1, 2, -1, 1], use_dwc_mlps=[True] * 4, use_lpus=[True] * 4, use_conv_patches=True,
             ksizes=[5, 3, 3, 3], drop_path_rate=0.0, use_checkpoint=False, img_size=64)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_heads_match_reference_bit_for_bit():
    import_reference()
    ref_uper = importlib.import_module("models.heads.uper_head").UPerHead
    ref_fcn = importlib.import_module("models.heads.fcn_head").FCNHead
    dims = [32, 64, 128, 256]
    torch.manual_seed(4)
    r_u, r_f = ref_uper(in_channels=dims, num_classes=19, channels=64), ref_fcn(in_channels=128, num_classes=19, channels=32)
    torch.manual_seed(4)
    m_u, m_f = UPerHead(dims, 19, channels=64), FCNHead(128, 19, channels=32)
    for r, m in ((r_u, m_u), (r_f, m_f)):
        sr, sm = r.state_dict(), m.state_dict()
        assert list(sr.keys()) == list(sm.keys())
        assert all(torch.equal(sr[k], sm[k]) for k in sr)            # same layers in the same order: same init stream
        m.load_state_dict(sr, strict=True)
        r.eval(), m.eval()
    feats = [torch.randn(2, c, 32 >> i, 40 >> i) for i, c in enumerate(dims)]
    with torch.no_grad():
        assert torch.equal(r_u(feats), m_u(feats))
        assert torch.equal(r_f(feats[2]), m_f(feats[2]))
    # train mode (BatchNorm batch statistics, dropout off): forward and gradients
    r_u.train(), m_u.train()
    r_u.dropout.p = m_u.dropout.p = 0.0
    fr = [f.clone().requires_grad_(True) for f in feats]
    fm = [f.clone().requires_grad_(True) for f in feats]
    r_u(fr).square().mean().backward()
    m_u(fm).square().mean().backward()
    for a, b in zip(fr, fm):
        assert torch.equal(a.grad, b.grad)


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_segmentor_matches_reference_wrapper():
    import_reference()
    ref_seg = importlib.import_module("models.segmentor").EncoderDecoder
    ref_uper = importlib.import_module("models.heads.uper_head").UPerHead
    ref_fcn = importlib.import_module("models.heads.fcn_head").FCNHead
    from dat_segmentation_b200.backbone import build_dat
    torch.manual_seed(6)
    mine = build_segmentor(SMALL, num_classes=21, attn_cls=orc.OracleDAttention)
    ref = ref_seg(build_dat(SMALL, attn_cls=orc.OracleDAttention), ref_uper(in_channels=SMALL["dims"], num_classes=21),
                  ref_fcn(in_channels=SMALL["dims"][2], num_classes=21))
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    ref.load_state_dict(mine.state_dict(), strict=True)
    x = torch.randn(2, 3, 64, 96)
    ref.eval(), mine.eval()
    with torch.no_grad():
        assert torch.equal(ref(x), mine(x))
    assert torch.equal(ref.last_aux_logits, mine.last_aux_logits)


def test_segmentor_training_outputs_and_loss():
    torch.manual_seed(8)
    model = build_segmentor(SMALL, num_classes=7, attn_cls=orc.OracleDAttention).train()
    x = torch.randn(2, 3, 64, 64)
    masks = torch.randint(0, 7, (2, 64, 64))
    masks[:, :4] = 255                                   # ignored pixels
    main, aux = model(x)
    assert main.shape == aux.shape == (2, 7, 64, 64)
    loss = segmentation_loss((main, aux), masks)
    loss.backward()
    assert torch.isfinite(loss)
    assert all(p.grad is not None for p in model.parameters() if p.requires_grad)
    model.eval()
    with torch.no_grad():
        out = model(x)
    assert isinstance(out, torch.Tensor) and out.shape == (2, 7, 64, 64)
