"""GPU parity of the block's variant branches (dat_blocks.py:57-59,84-99,156-157,
164-167,185-197,221-222) against golden vectors generated from the unmodified reference
(tests/golden/make_golden_variants.py): fp32 forward 1e-5 relative, gradients 5e-5, bf16 autocast
forward 2e-2 max-abs."""
import pytest
import torch

from golden_util import load_variant, rel_err
from oracle import dattn_oracle as orc
from make_golden_variants import VARIANTS
from test_abi_and_host import _variant_module

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_forward_backward_fp32(name):
    mod, cfg, x, dy, rec = _variant_module(name)
    mod.load_state_dict(rec["params"], strict=True)
    mod = mod.cuda()
    xd = x.cuda().requires_grad_(True)
    y, _, _ = mod(xd)
    assert rel_err(y.detach().cpu(), rec["y"]) < 1e-5
    y.backward(dy.cuda())
    report = {"dx": rel_err(xd.grad.cpu(), rec["dx"])}
    for k, p in mod.named_parameters():
        if k in rec["frozen"]:
            assert p.grad is None, k
            continue
        if k == "proj_k.bias":      # analytically zero (softmax shift invariance): absolute check
            assert p.grad.abs().max().item() < 1e-4
            continue
        report[k] = rel_err(p.grad.cpu(), rec["grads"][k])
    bad = {k: v for k, v in report.items() if v > 5e-5}
    if bad:
        # A hidden unit of the log-CPB MLP sitting at ReLU's kink flips between two fp32 evaluation orders: the
        # reference's own fp32 gradient is then off from the fp64 value by more than the tolerance.  Allow what
        # the reference itself loses against the fp64 oracle (pinned to it in tests/test_oracle_golden.py).
        p64 = {k: v.double().requires_grad_(k not in rec["frozen"]) for k, v in rec["params"].items()}
        x64 = x.double().requires_grad_(True)
        orc.forward_libops(x64, p64, cfg).backward(dy.double())
        truth = dict({k: v.grad for k, v in p64.items() if v.grad is not None}, dx=x64.grad)
        ref = dict(rec["grads"], dx=rec["dx"])
        bad = {k: v for k, v in bad.items() if v > 5e-5 + 1.5 * rel_err(ref[k], truth[k])}
    assert not bad, bad


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_forward_bf16_autocast(name):
    mod, cfg, x, dy, rec = _variant_module(name)
    mod.load_state_dict(rec["params"], strict=True)
    mod = mod.cuda()
    xd = x.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y, _, _ = mod(xd)
    assert y.dtype == torch.bfloat16
    assert (y.float().cpu() - rec["y"]).abs().max().item() < 2e-2
    y.backward(dy.cuda().bfloat16())      # the bf16 backward of every branch runs and is finite
    assert torch.isfinite(xd.grad).all()
    for k, p in mod.named_parameters():
        if k not in rec["frozen"]:
            assert torch.isfinite(p.grad).all(), k


def test_variant_pos_ref_opt_in_no_off():
    mod, cfg, x, dy, rec = _variant_module("no_off")
    mod.load_state_dict(rec["params"], strict=True)
    mod = mod.cuda()
    mod.return_pos_ref = True
    with torch.no_grad():
        y, pos, ref = mod(x.cuda())
    assert rel_err(y.cpu(), rec["y"]) < 1e-5
