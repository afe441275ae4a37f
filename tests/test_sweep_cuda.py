"""BASELINE.json configs[4] / SURVEY.md 8d config 5: the DAttention sweep — stage presets of DAT-T++ x
groups {1,2,4,8} (where heads % groups == 0) x offset_range_factor {1,2,3} x feature maps 16^2..128^2 —
CUDA block (through the module / C ABI) against the CPU oracle on the same seeded inputs.
fp32: 1e-5 relative forward, 5e-5 gradients; bf16 autocast: 2e-2 max-abs forward."""
import pytest
import torch

from golden_util import rel_err
from oracle import dattn_oracle as orc

pytestmark = pytest.mark.gpu

PRESETS = {  # map size at a 512^2 input -> (heads, stride, ksize, q_size)   (upn_tiny...py:11-28)
    128: (2, 8, 9, 56), 64: (4, 4, 7, 28), 32: (8, 2, 5, 14), 16: (16, 1, 3, 7)}
SWEEP = [(hw, g, orf) for hw, (h, _, _, _) in PRESETS.items() for g in (1, 2, 4, 8) if h % g == 0
         for orf in (1, 2, 3)]


def _build(hw, groups, orf, seed):
    """The module with its own (reference-identical) default initialisation under a seed, offsets and
    bias table widened as in tests/golden/make_golden.py so OOB taps / tanh saturation are exercised."""
    from dat_segmentation_b200.dattention import DAttentionBaseline
    heads, stride, ksize, qs = PRESETS[hw]
    cfg = orc.BlockCfg(qs, qs, heads, 32, groups, stride, ksize, orf)
    torch.manual_seed(seed)
    m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, orf, True, False,
                           False, False, ksize, False, 0)
    with torch.no_grad():
        m.conv_offset[3].weight.mul_(2.0)
        m.rpe_table.mul_(10.0)
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return cfg, params, m.cuda()


@pytest.mark.parametrize("hw,groups,orf", SWEEP)
def test_sweep_forward(hw, groups, orf):
    cfg, params, m = _build(hw, groups, orf, seed=hw + groups)
    g = torch.Generator().manual_seed(1000 * hw + 10 * groups + orf)
    x = torch.randn(1, cfg.nc, hw, hw, generator=g)
    y_ref = orc.forward_libops(x, params, cfg)
    with torch.no_grad():
        y, _, _ = m(x.cuda())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yb, _, _ = m(x.cuda())
    assert rel_err(y.cpu(), y_ref) < 1e-5
    assert (yb.float().cpu() - y_ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("hw,groups", sorted({(hw, g) for hw, g, _ in SWEEP}))
def test_sweep_backward_fp32(hw, groups):
    """Gradients of x and of all 14 parameters against torch.autograd of the oracle's library-op
    form (tanh offsets, orf = 2: no clamp mask, gradients continuous)."""
    cfg, params, m = _build(hw, groups, 2, seed=3 * hw + groups)
    g = torch.Generator().manual_seed(77 * hw + groups)
    x = torch.randn(1, cfg.nc, hw, hw, generator=g)
    dy = torch.randn(1, cfg.nc, hw, hw, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    orc.forward_libops(xr, pr, cfg).backward(dy)
    xd = x.cuda().requires_grad_(True)
    y, _, _ = m(xd)
    y.backward(dy.cuda())
    report = {"dx": rel_err(xd.grad.cpu(), xr.grad)}
    for k, p in m.named_parameters():
        if k == "proj_k.bias":      # analytically zero (softmax shift invariance): absolute check
            assert p.grad.abs().max().item() < 1e-4
            continue
        report[k] = rel_err(p.grad.cpu(), pr[k].grad)
    bad = {k: v for k, v in report.items() if v > 5e-5}
    assert not bad, bad
