"""Pins the CPU oracle (oracle/dattn_oracle.py) against outputs of the
unmodified reference module (tests/golden/*.pt).  CPU only."""
import pytest
import torch

from golden_util import CASES, load_case, nhwc, rel_err
from oracle import dattn_oracle as orc

SMALL = [n for n in CASES if n != "cfg1_stage2"]


@pytest.mark.parametrize("name", list(CASES))
def test_forward_explicit_matches_reference(name):
    cfg, x, dy, rec = load_case(name)
    fw = orc.forward_explicit(nhwc(x), rec["params"], cfg)
    assert rel_err(fw["y"], nhwc(rec["y"])) < 1e-5
    assert rel_err(fw["pos"], rec["pos_l"]) < 1e-5
    assert rel_err(fw["xs"], rec["xs_l"]) < 2e-4  # pos error x feature gradient
    if "bias_l" in rec:
        assert rel_err(fw["bias"], rec["bias_l"]) < 2e-4


@pytest.mark.parametrize("name", list(CASES))
def test_reference_points_bit_exact(name):
    """ref = pos - offset is not observable; instead check the oracle's ref grid
    against the reference's own op sequence (linspace → div_ → mul_ → sub_,
    dat_blocks.py:111-118) bit for bit, over the sizes in scope."""
    cfg, x, _, rec = load_case(name)
    hk, wk = rec["pos"].shape[1:3]
    for n in (hk, wk, 2, 3, 16, 64, 112, 128, 512):
        ref = torch.linspace(0.5, n - 0.5, n).div_(n - 1.0).mul_(2.0).sub_(1.0)
        got, _ = orc.ref_points(n, n)
        assert torch.equal(ref, got), n


@pytest.mark.parametrize("name", list(CASES))
def test_tap_indices_bit_exact_given_reference_pos(name):
    """With the reference's pos injected, the oracle's sampled features must be the
    reference's (same integer taps, same weights): max-abs tiny, and the taps the
    oracle reports reproduce xs exactly when used for a manual gather."""
    cfg, x, _, rec = load_case(name)
    xs, (x0, y0) = orc.sample_features_explicit(nhwc(x), rec["pos_l"], cfg)
    assert rel_err(xs, rec["xs_l"]) < 2e-6
    H, W = x.shape[2:]
    px = rec["pos_l"][..., 1].reshape(x0.shape)
    ix = ((px + 1) / 2) * (W - 1)
    assert torch.equal(x0, torch.floor(ix).long())


@pytest.mark.parametrize("name", list(CASES))
def test_libops_port_matches_reference(name):
    cfg, x, dy, rec = load_case(name)
    y = orc.forward_libops(x, rec["params"], cfg)
    assert rel_err(y, rec["y"]) < 1e-6


@pytest.mark.parametrize("name", SMALL)
def test_backward_explicit_matches_reference_autograd(name):
    cfg, x, dy, rec = load_case(name)
    dx, grads, _ = orc.backward_explicit(nhwc(x), rec["params"], cfg, nhwc(dy))
    assert rel_err(dx, nhwc(rec["dx"])) < 2e-5
    for key, ref in rec["grads"].items():
        if key == "proj_k.bias":      # analytically zero (softmax shift invariance)
            assert grads[key].abs().max() < 1e-4
            continue
        assert rel_err(grads[key], ref) < 5e-5, key


def test_backward_explicit_fp64_vs_autograd_of_port():
    """fp64: analytic gradients == autograd through the library-op port."""
    cfg, x, dy, rec = load_case("odd_c96_orf3")
    p64 = {k: v.double() for k, v in rec["params"].items()}
    x64, dy64 = x.double(), dy.double()
    leaves = {k: v.clone().requires_grad_(True) for k, v in p64.items()}
    xin = x64.clone().requires_grad_(True)
    y = orc.forward_libops(xin, leaves, cfg)
    y.backward(dy64)
    dx, grads, _ = orc.backward_explicit(nhwc(x64), p64, cfg, nhwc(dy64))
    assert rel_err(dx, nhwc(xin.grad)) < 1e-10
    for key in leaves:
        if key == "proj_k.bias":
            continue
        assert rel_err(grads[key], leaves[key].grad) < 1e-9, key


# ---- variant branches (dat_blocks.py:57-59,84-99,156-157,164-167,185-197,221-222) ----------------

from golden_util import load_variant  # noqa: E402
from make_golden_variants import VARIANTS  # noqa: E402


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_port_matches_reference(name):
    """The library-op port with the variant flags == the unmodified reference: forward, dx and the
    gradient of every trainable parameter (autograd of the port vs the reference's autograd)."""
    cfg, x, dy, rec = load_variant(name)
    leaves = {k: v.clone().requires_grad_(k not in rec["frozen"]) for k, v in rec["params"].items()}
    xin = x.clone().requires_grad_(True)
    y = orc.forward_libops(xin, leaves, cfg)
    assert rel_err(y, rec["y"]) < 1e-6
    y.backward(dy)
    assert rel_err(xin.grad, rec["dx"]) < 1e-5
    assert set(rec["grads"]) == {k for k, v in leaves.items() if v.grad is not None}
    for k, g in rec["grads"].items():
        assert rel_err(leaves[k].grad, g) < 1e-5, k


@pytest.mark.parametrize("name", ["stage3_small", "k_eq_s_orf1", "odd_c96_orf3"])
def test_table_gradient_separable_form_matches_autograd(name):
    """The GEMM form of d rpe_table used by csrc/rpe_table_grad.cu (dT_n = A_n^T dS_n B_n with hat weights) equals
    autograd through the reference's own bias computation (F.grid_sample of the table at the displacements)."""
    cfg, x, _, rec = load_case(name)
    H, W = x.shape[2:]
    pos = rec["pos_l"].double()
    table = rec["params"]["rpe_table"].double().clone().requires_grad_(True)
    bias = orc.rpe_bias_explicit(pos, table, H, W, cfg)                     # (B,h,HW,Ns), differentiable in table
    g = torch.Generator().manual_seed(3)
    ds = torch.randn(bias.shape, generator=g, dtype=torch.float64)
    bias.backward(ds)
    got = orc.rpe_table_grad_separable(ds, pos, H, W, cfg)
    assert rel_err(got, table.grad) < 1e-10


def test_split_kv_merge_equals_full_softmax():
    """The chunk merge of the split-KV attention forward (attn_combine_kernel) is exact: 1024 samples as
    256 + 256 + 256 + 128 + 64 + 64 against one softmax over all of them (fp64)."""
    g = torch.Generator().manual_seed(5)
    S = torch.randn(3, 40, 1024, generator=g, dtype=torch.float64) * 3
    V = torch.randn(3, 1024, 32, generator=g, dtype=torch.float64)
    full = torch.softmax(S, -1) @ V
    o_parts, l_parts, n0 = [], [], 0
    for n in (256, 256, 256, 128, 64, 64):
        Sc, Vc = S[..., n0:n0 + n], V[:, n0:n0 + n]
        o_parts.append(torch.softmax(Sc, -1) @ Vc)
        l_parts.append(torch.logsumexp(Sc, -1))
        n0 += n
    o, lse = orc.merge_sample_chunks(o_parts, l_parts)
    assert rel_err(o, full) < 1e-12
    assert rel_err(lse, torch.logsumexp(S, -1)) < 1e-12
