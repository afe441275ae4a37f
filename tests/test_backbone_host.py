"""Host-side DAT backbone (dat_segmentation_b200/backbone.py): wiring and state-dict parity
with the reference DAT (container only: needs /root/reference), and CUDA-vs-oracle parity
of the whole backbone on the GPU box."""
import pytest
import torch

from oracle import dattn_oracle as orc
from oracle.ref_shim import import_reference, reference_available

from dat_segmentation_b200.backbone import DAT_TINY_PP, build_dat


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_backbone_matches_reference_dat_on_cpu():
    blocks, dat = import_reference()
    cfg = dict(DAT_TINY_PP)
    torch.manual_seed(3)
    ref = dat.DAT(**cfg).eval()
    torch.manual_seed(3)
    mine = build_dat(cfg, attn_cls=blocks.DAttentionBaseline).eval()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    assert all(torch.equal(sd_ref[k], sd_mine[k]) for k in sd_ref)     # same init RNG stream
    mine.load_state_dict(sd_ref, strict=True)
    x = torch.randn(1, 3, 128, 128)
    with torch.no_grad():
        o_ref, o_mine = ref(x), mine(x)
    for a, b in zip(o_ref, o_mine):
        assert a.shape == b.shape and torch.equal(a, b)
    # and with the oracle port of the block in place of the reference block
    port = build_dat(cfg, attn_cls=orc.OracleDAttention).eval()
    port.load_state_dict(sd_ref, strict=True)
    with torch.no_grad():
        o_port = port(x)
    for a, b in zip(o_ref, o_port):
        assert _rel(b, a) < 1e-5


def test_backbone_cuda_class_shares_state_dict_with_port():
    """Keys/shapes of the CUDA-backed backbone == the oracle-port backbone (CPU only: no forward)."""
    a = build_dat(attn_cls=orc.OracleDAttention)
    b = build_dat()
    assert {k: v.shape for k, v in a.state_dict().items()} == {k: v.shape for k, v in b.state_dict().items()}
    n_attn = sum(1 for m in b.modules() if type(m).__name__ == "DAttentionBaseline")
    assert n_attn == 14       # DAT-T++: 1 + 2 + 9 + 2 deformable blocks


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(128, 128), (96, 160)])
def test_backbone_cuda_vs_oracle_port(size):
    # the non-attention layers are library convolutions: keep them in true fp32 here
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    gpu = build_dat(drop_path_rate=0.0).cuda()
    with torch.no_grad():
        for m in gpu.modules():
            if type(m).__name__ == "DAttentionBaseline":
                m.conv_offset[3].weight.mul_(2.0)
                m.rpe_table.mul_(10.0)
    cpu = build_dat(drop_path_rate=0.0, attn_cls=orc.OracleDAttention)
    cpu.load_state_dict({k: v.cpu() for k, v in gpu.state_dict().items()}, strict=True)
    x = torch.randn(2, 3, *size)
    xg = x.cuda().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    og, oc = gpu(xg), cpu(xc)
    for a, b in zip(og, oc):
        assert _rel(a.detach().cpu(), b.detach()) < 2e-4     # 14 blocks deep, fp32, library convs differ too
    sum(o.square().mean() for o in og).backward()
    sum(o.square().mean() for o in oc).backward()
    assert _rel(xg.grad.cpu(), xc.grad) < 5e-3
    gp, cp = dict(gpu.named_parameters()), dict(cpu.named_parameters())
    worst = max((_rel(gp[k].grad.cpu(), cp[k].grad), k) for k in gp if cp[k].grad is not None
                and cp[k].grad.abs().max() > 1e-6)
    print("worst param-grad rel err", worst)
    assert worst[0] < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["small", "base"])
def test_backbone_family_cuda_vs_oracle_port(family):
    """DAT-S++ / DAT-B++ widths (BASELINE.json configs[2], [3]; C = 96.. and 128..1024, 3 / 6 / 12 / 24 and up to
    32 heads, groups 1..16) with the depths cut to one 'X' + 'D' pair per stage so the CPU port stays quick:
    fp32 CUDA backbone vs the oracle-port backbone, then the bf16 autocast forward."""
    from dat_segmentation_b200.backbone import DAT_BASE_PP, DAT_SMALL_PP
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(DAT_SMALL_PP if family == "small" else DAT_BASE_PP, depths=[2, 2, 2, 2],
               stage_spec=[["X", "D"], ["X", "D"], ["X", "D"], ["D", "D"]], drop_path_rate=0.0)
    torch.manual_seed(9)
    gpu = build_dat(cfg).cuda()
    with torch.no_grad():
        for m in gpu.modules():
            if type(m).__name__ == "DAttentionBaseline":
                m.conv_offset[3].weight.mul_(2.0)
                m.rpe_table.mul_(10.0)
    cpu = build_dat(cfg, attn_cls=orc.OracleDAttention)
    cpu.load_state_dict({k: v.cpu() for k, v in gpu.state_dict().items()}, strict=True)
    x = torch.randn(1, 3, 128, 160)
    xg = x.cuda().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    og, oc = gpu(xg), cpu(xc)
    for a, b in zip(og, oc):
        assert _rel(a.detach().cpu(), b.detach()) < 2e-4
    sum(o.square().mean() for o in og).backward()
    sum(o.square().mean() for o in oc).backward()
    assert _rel(xg.grad.cpu(), xc.grad) < 5e-3
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ob = gpu(x.cuda())
    for a, b in zip(ob, oc):
        assert _rel(a.float().cpu(), b.detach()) < 6e-2      # bf16 through 8 blocks; outputs are LayerNormed (O(1))


@pytest.mark.gpu
def test_backbone_with_variant_branches_cuda_vs_oracle_port():
    """One variant branch per stage (dwc_pe, log_cpb, fixed_pe with a real resize, no_off) wired through the backbone's
    per-stage flags exactly as `DAT.__init__` passes them (dat.py:99-116): fp32 CUDA backbone vs oracle-port backbone."""
    from dat_segmentation_b200.backbone import DAT_TINY_PP
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(DAT_TINY_PP, img_size=128, depths=[2, 2, 2, 2], stage_spec=[["X", "D"]] * 3 + [["D", "D"]],
               drop_path_rate=0.0, dwc_pes=[True, False, False, False], log_cpb=[False, True, False, False],
               fixed_pes=[False, False, True, False], no_offs=[False, False, False, True])
    torch.manual_seed(13)
    gpu = build_dat(cfg).cuda()
    kinds = [type(m.rpe_table).__name__ for m in gpu.modules() if type(m).__name__ == "DAttentionBaseline"]
    assert kinds == ["Conv2d", "Sequential", "Parameter", "NoneType", "NoneType"], kinds
    cpu = build_dat(cfg, attn_cls=orc.OracleDAttention)
    cpu.load_state_dict({k: v.cpu() for k, v in gpu.state_dict().items()}, strict=True)
    x = torch.randn(2, 3, 128, 160)
    xg = x.cuda().requires_grad_(True)
    xc = x.clone().requires_grad_(True)
    og, oc = gpu(xg), cpu(xc)
    for a, b in zip(og, oc):
        assert _rel(a.detach().cpu(), b.detach()) < 2e-4
    sum(o.square().mean() for o in og).backward()
    sum(o.square().mean() for o in oc).backward()
    assert _rel(xg.grad.cpu(), xc.grad) < 5e-3
    gp, cp = dict(gpu.named_parameters()), dict(cpu.named_parameters())
    assert {k for k, v in gp.items() if v.grad is None} == {k for k, v in cp.items() if v.grad is None}
    worst = max((_rel(gp[k].grad.cpu(), cp[k].grad), k) for k in gp if cp[k].grad is not None
                and cp[k].grad.abs().max() > 1e-6)
    assert worst[0] < 2e-2, worst


@pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
def test_install_swaps_the_block_inside_the_reference_backbone():
    """`dat_segmentation_b200.install.install()` makes the UNMODIFIED reference `DAT` build dat_b200 blocks (positional
    16-argument construction, dat.py:99-116) with an unchanged state dict; `uninstall()` restores the reference class."""
    blocks, dat = import_reference()
    import dat_segmentation_b200.install as b200
    from dat_segmentation_b200.dattention import DAttentionBaseline as B200Block
    ref_cls = blocks.DAttentionBaseline
    torch.manual_seed(3)
    ref = dat.DAT(**DAT_TINY_PP)
    try:
        patched = b200.install()
        assert "models.utils.dat_blocks" in patched and "models.backbones.dat" in patched
        torch.manual_seed(3)
        swapped = dat.DAT(**DAT_TINY_PP)
        attn = [m for m in swapped.modules() if type(m).__name__ == "DAttentionBaseline"]
        assert len(attn) == 14 and all(type(m) is B200Block for m in attn)
        sd_ref, sd_new = ref.state_dict(), swapped.state_dict()
        assert list(sd_ref.keys()) == list(sd_new.keys())
        assert all(torch.equal(sd_ref[k], sd_new[k]) for k in sd_ref)        # same init stream too
        swapped.load_state_dict(sd_ref, strict=True)
        with pytest.raises(RuntimeError):
            swapped(torch.randn(1, 3, 64, 64))                               # CPU tensor: the block has no CPU path
    finally:
        b200.uninstall()
    assert blocks.DAttentionBaseline is ref_cls and dat.DAttentionBaseline is ref_cls


def test_backbone_activation_checkpointing_gives_the_same_gradients():
    """`use_checkpoint=True` (dat.py:161-165: every stage under torch.utils.checkpoint in training) must not change
    outputs or gradients; CPU, oracle port of the block."""
    small = dict(dim_stem=32, dims=[32, 64, 128, 256], depths=[2, 1, 1, 1], stage_spec=[["X", "D"], ["D"], ["D"], ["D"]],
                 heads=[1, 2, 4, 8], groups=[1, 1, 2, 4], use_pes=[True] * 4, strides=[4, 2, 1, 1],
                 offset_range_factor=[-1, 2, -1, 1], use_dwc_mlps=[True] * 4, use_lpus=[True] * 4, use_conv_patches=True,
                 ksizes=[5, 3, 3, 3], nat_ksizes=[7] * 4, drop_path_rate=0.0, img_size=64)
    torch.manual_seed(21)
    a = build_dat(dict(small, use_checkpoint=False), attn_cls=orc.OracleDAttention).train()
    b = build_dat(dict(small, use_checkpoint=True), attn_cls=orc.OracleDAttention).train()
    b.load_state_dict(a.state_dict(), strict=True)
    x = torch.randn(2, 3, 64, 64)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    oa, ob = a(xa), b(xb)
    for u, v in zip(oa, ob):
        assert torch.equal(u, v)
    sum(o.square().mean() for o in oa).backward()
    sum(o.square().mean() for o in ob).backward()
    assert _rel(xb.grad, xa.grad) < 1e-6
    ga, gb = dict(a.named_parameters()), dict(b.named_parameters())
    assert all(_rel(gb[k].grad, ga[k].grad) < 1e-5 for k in ga if ga[k].grad is not None and ga[k].grad.abs().max() > 0)
