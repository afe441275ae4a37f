"""Helpers shared by the parity tests: load a golden case (generated from the
unmodified reference by tests/golden/make_golden.py) in oracle conventions."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import CASES, HC, golden_inputs  # noqa: E402

from oracle.dattn_oracle import BlockCfg  # noqa: E402


def rel_err(a, b):
    """max|a-b| / max|b| — the 'relative' error of north_star's fp32 tolerance."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load_case(name):
    rec = torch.load(os.path.join(HERE, "golden", f"{name}.pt"), weights_only=False)
    m = rec["meta"]
    cfg = BlockCfg(m["q_size"][0], m["q_size"][1], m["n_heads"], HC, m["n_groups"], m["stride"],
                   m["ksize"], m["orf"])
    x, dy = golden_inputs(name)
    assert abs(x.double().sum().item() - rec["x_sum"]) < 1e-6, "seeded input drifted"
    assert abs(dy.double().sum().item() - rec["dy_sum"]) < 1e-6, "seeded input drifted"
    B, G = m["B"], m["n_groups"]
    hk, wk = rec["pos"].shape[1:3]
    rec["pos_l"] = rec["pos"].reshape(B, G, hk, wk, 2)
    # xs (B*G, Cg, Hk, Wk) -> (B, Ns, C)
    rec["xs_l"] = rec["xs"].reshape(B, G, cfg.cg, hk * wk).permute(0, 3, 1, 2).reshape(B, hk * wk, cfg.nc)
    if "bias" in rec:
        rec["bias_l"] = rec["bias"].reshape(B, cfg.n_heads, m["H"] * m["W"], hk * wk)
    return cfg, x, dy, rec


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def load_variant(name):
    """Golden case of a variant branch (tests/golden/make_golden_variants.py): (cfg, x, dy, record)."""
    from make_golden_variants import VARIANTS, variant_inputs  # noqa: F401
    rec = torch.load(os.path.join(HERE, "golden", "variants.pt"), weights_only=False)[name]
    m = rec["meta"]
    cfg = BlockCfg(m["q_size"][0], m["q_size"][1], m["n_heads"], HC, m["n_groups"], m["stride"], m["ksize"],
                   m["orf"], **m["flags"])
    x, dy = variant_inputs(name)
    assert abs(x.double().sum().item() - rec["x_sum"]) < 1e-6, "seeded input drifted"
    return cfg, x, dy, rec
