"""CPU-side tests: the C-ABI library loads and exports every symbol declared in
include/dat_b200.h; host-only entry points validate arguments; the Python module
mirrors the reference's constructor / state-dict surface.  No GPU compute."""
import ctypes as C
import os
import re

import pytest
import torch

from golden_util import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cab():
    from dat_segmentation_b200 import _cabi, build
    build.build()
    return _cabi


def test_header_symbols_all_exported(cab):
    hdr = open(os.path.join(ROOT, "include", "dat_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dat_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = cab.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dat_b200.h but not exported"
    assert declared == set(cab.exported_symbols())


def test_sample_grid_matches_conv_arithmetic(cab):
    lib = cab.lib()
    for (H, W, s, k) in [(128, 128, 8, 9), (64, 64, 4, 7), (32, 32, 2, 5), (16, 16, 1, 3), (43, 22, 2, 5),
                         (12, 10, 2, 2), (128, 512, 8, 9)]:
        d = cab.BlockDesc(1, H, W, 2, 1, s, k, 27, 27, -1.0, 0, 0)
        hk, wk = C.c_int32(), C.c_int32()
        assert lib.dat_sample_grid(C.byref(d), C.byref(hk), C.byref(wk)) == 0
        pad = k // 2 if k != s else 0
        conv = torch.nn.Conv2d(1, 1, k, s, pad)
        out = conv(torch.zeros(1, 1, H, W))
        assert (hk.value, wk.value) == tuple(out.shape[2:])


def test_bad_descriptor_is_an_error_not_a_crash(cab):
    lib = cab.lib()
    hk, wk = C.c_int32(), C.c_int32()
    d = cab.BlockDesc(1, 32, 32, 6, 4, 2, 5, 27, 27, -1.0, 0, 0)   # heads % groups != 0
    assert lib.dat_sample_grid(C.byref(d), C.byref(hk), C.byref(wk)) == -1
    assert b"n_heads" in lib.dat_last_error()
    d = cab.BlockDesc(1, 8, 8, 2, 1, 8, 9, 27, 27, -1.0, 0, 0)     # Hk = 1 → reference divides by 0
    assert lib.dat_sample_grid(C.byref(d), C.byref(hk), C.byref(wk)) == -1
    assert lib.dat_block_bwd_workspace_bytes(C.byref(d)) == 0


def test_workspace_query(cab):
    d = cab.BlockDesc(16, 32, 32, 8, 4, 2, 5, 27, 27, -1.0, 0, 1)
    n = cab.lib().dat_block_bwd_workspace_bytes(C.byref(d))
    assert 0 < n < (1 << 31)


def test_module_surface_matches_reference():
    from dat_segmentation_b200.dattention import DAttentionBaseline
    cfg, x, dy, rec = load_case("cfg1_stage2")
    m = DAttentionBaseline((14, 14), (14, 14), 8, 32, 4, 0.0, 0.0, 2, -1, True, False, False, False, 5, False, 2)
    sd = m.state_dict()
    assert list(sd.keys()) == list(rec["params"].keys())          # same names, same order
    assert all(sd[k].shape == rec["params"][k].shape for k in sd)
    m.load_state_dict(rec["params"], strict=True)
    names = [n for n, _ in m.named_parameters()]
    assert "rpe_table" in names and any("norm" in n for n in names)  # weight-decay filters (new_train.py:146-157)
    with pytest.raises(RuntimeError):
        m(x)   # CPU tensor: no fallback


def test_default_init_matches_reference_rng_stream():
    """Same seed → same initial weights as the reference constructor (fixtures were
    generated with manual_seed(7) and two rescalings)."""
    from dat_segmentation_b200.dattention import DAttentionBaseline
    cfg, x, dy, rec = load_case("k_eq_s_orf1")
    torch.manual_seed(7)
    m = DAttentionBaseline((12, 12), (12, 12), 2, 32, 2, 0.0, 0.0, 2, 1, True, False, False, False, 2, False, 2)
    with torch.no_grad():
        m.conv_offset[3].weight.mul_(2.0)
        m.rpe_table.mul_(10.0)
    for k, v in m.state_dict().items():
        assert torch.equal(v, rec["params"][k]), k


def _variant_module(name):
    from golden_util import load_variant
    from dat_segmentation_b200.dattention import DAttentionBaseline
    cfg, x, dy, rec = load_variant(name)
    m, f = rec["meta"], rec["meta"]["flags"]
    mod = DAttentionBaseline(m["q_size"], m["q_size"], m["n_heads"], 32, m["n_groups"], 0.0, 0.0, m["stride"], m["orf"],
                             f["use_pe"], f["dwc_pe"], f["no_off"], f["fixed_pe"], m["ksize"], f["log_cpb"], 2)
    return mod, cfg, x, dy, rec


@pytest.mark.parametrize("name", ["no_pe", "no_off", "dwc_pe", "fixed_pe_resize", "log_cpb"])
def test_variant_modules_share_the_reference_state_dict(name):
    """Variant branches (dat_blocks.py:57-59,84-104): same parameter names / shapes / frozen set as the
    reference module, same default initialisation under the same seed."""
    torch.manual_seed(11)
    mod, cfg, x, dy, rec = _variant_module(name)
    sd = mod.state_dict()
    assert list(sd.keys()) == list(rec["params"].keys())
    assert all(sd[k].shape == rec["params"][k].shape for k in sd)
    assert sorted(k for k, v in mod.named_parameters() if not v.requires_grad) == sorted(rec["frozen"])
    with torch.no_grad():
        mod.conv_offset[3].weight.mul_(2.0)
        if isinstance(mod.rpe_table, torch.nn.Parameter):
            mod.rpe_table.mul_(30.0)
    for k, v in mod.state_dict().items():
        assert torch.equal(v, rec["params"][k]), k
    mod.load_state_dict(rec["params"], strict=True)


def test_pointwise_predicate_follows_the_library(cab):
    """`pointwise._supported` asks the C ABI whether the weight-gradient kernel tiles a shape (a Python copy of the
    rule once disagreed with the library for K = 192).  Host-only calls: no GPU needed."""
    from dat_segmentation_b200.pointwise import _supported
    lib = cab.lib()
    for M, N, K in [(16384, 1024, 256), (16384, 256, 1024), (4096, 768, 192), (4096, 192, 768), (4096, 1536, 384),
                    (262144, 64, 256), (1024, 3072, 768)]:
        assert _supported(M, N, K), (M, N, K)
        assert lib.dat_pointwise_wgrad_tc_workspace_bytes(M, N, K) > 0
    for M, N, K in [(4096, 384, 96), (4096, 96, 384), (32, 256, 256), (4096, 100, 64)]:
        assert not _supported(M, N, K), (M, N, K)


def test_variant_descriptors_and_workspace_queries(cab):
    """pe_mode / no_off in dat_block_desc: the sample grid follows avg_pool2d for no_off (floor division, no padding,
    dat_blocks.py:165-167), the dense-bias variants reserve the (B, h, HW, Ns) fp32 tensors the reference materialises,
    a bad pe_mode is an argument error.  Host-only calls."""
    lib = cab.lib()
    hk, wk = C.c_int32(), C.c_int32()
    d = cab.BlockDesc(2, 18, 22, 2, 1, 4, 7, 27, 27, 2.0, 0, 0, cab.PE_RPE, 1)          # no_off
    assert lib.dat_sample_grid(C.byref(d), C.byref(hk), C.byref(wk)) == 0
    assert (hk.value, wk.value) == (18 // 4, 22 // 4)
    base = cab.BlockDesc(2, 16, 16, 4, 2, 2, 5, 27, 27, -1.0, 0, 0, cab.PE_RPE, 0)
    n_rpe_f = lib.dat_block_fwd_workspace_bytes(C.byref(base))
    n_rpe_b = lib.dat_block_bwd_workspace_bytes(C.byref(base))
    dense = 2 * 4 * 256 * 64 * 4                                                       # B h HW Ns fp32
    for mode, fwd_extra, bwd_extra in ((cab.PE_LOGCPB, dense, 2 * dense), (cab.PE_FIXED, dense // 2, dense // 2 + dense)):
        dv = cab.BlockDesc(2, 16, 16, 4, 2, 2, 5, 64, 16, -1.0, 0, 0, mode, 0)
        assert lib.dat_block_fwd_workspace_bytes(C.byref(dv)) >= n_rpe_f + fwd_extra
        assert lib.dat_block_bwd_workspace_bytes(C.byref(dv)) >= n_rpe_b + bwd_extra
    dv = cab.BlockDesc(2, 16, 16, 4, 2, 2, 5, 0, 0, -1.0, 0, 0, cab.PE_NONE, 0)         # no table: sizes ignored
    assert lib.dat_block_bwd_workspace_bytes(C.byref(dv)) > 0
    bad = cab.BlockDesc(2, 16, 16, 4, 2, 2, 5, 27, 27, -1.0, 0, 0, 9, 0)
    assert lib.dat_sample_grid(C.byref(bad), C.byref(hk), C.byref(wk)) == -1
    assert b"pe_mode" in lib.dat_last_error()
