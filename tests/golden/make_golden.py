"""Generates tests/golden/*.pt from the UNMODIFIED reference module.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference `DAttentionBaseline` (dat_blocks.py:19-227) is
constructed with its own default initialisation under a fixed seed, run forward
and backward on CPU fp32, and the results are stored.  `pos`, the sampled
features and the rpe bias are captured from the two `F.grid_sample` calls the
reference makes (dat_blocks.py:169 and :206) by wrapping that function while the
forward runs — the reference source is not edited.  Inputs are regenerated from
seeds by `golden_inputs()`; a float64 checksum guards against RNG drift.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# name: (B, H, W, n_heads, n_groups, stride, ksize, orf, q_size, store_big)
CASES = {
    # BASELINE.json configs[0]: DAT-T++ "stage-3" block, dim 256, 32x32 tokens, batch 2
    "cfg1_stage2": (2, 32, 32, 8, 4, 2, 5, -1, (14, 14), False),
    "stage0_small": (1, 40, 24, 2, 1, 8, 9, -1, (56, 56), True),
    "stage1_orf2": (2, 20, 28, 4, 2, 4, 7, 2, (28, 28), True),
    "stage3_small": (1, 7, 9, 8, 8, 1, 3, -1, (7, 7), True),
    "odd_c96_orf3": (2, 11, 13, 3, 1, 2, 5, 3, (14, 14), True),
    "k_eq_s_orf1": (1, 12, 10, 2, 2, 2, 2, 1, (12, 12), True),
    "stage2_orf1": (1, 16, 16, 8, 4, 2, 5, 1, (14, 14), True),
}
HC = 32


def golden_inputs(name):
    """Seeded x (B,C,H,W) and dy for a case (shared by generator and tests)."""
    B, H, W, heads = CASES[name][:4]
    C = heads * HC
    seed = 1000 + sorted(CASES).index(name)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    return x, dy


def main():
    from oracle.ref_shim import import_reference
    blocks, _ = import_reference()
    F = blocks.F
    for name, (B, H, W, heads, groups, stride, ksize, orf, q_size, big) in CASES.items():
        torch.manual_seed(7)
        mod = blocks.DAttentionBaseline(q_size, q_size, heads, HC, groups, 0.0, 0.0, stride, orf,
                                        True, False, False, False, ksize, False, 2)
        # default init leaves offsets small; widen the last offset conv so clamp /
        # out-of-bounds / tanh-saturation paths are all hit (values still seeded).
        with torch.no_grad():
            mod.conv_offset[3].weight.mul_(2.0)
            mod.rpe_table.mul_(10.0)
        x, dy = golden_inputs(name)
        x.requires_grad_(True)
        calls = []
        orig = F.grid_sample

        def spy(input, grid, **kw):
            out = orig(input, grid, **kw)
            calls.append((input.detach(), grid.detach(), out.detach()))
            return out

        F.grid_sample = spy
        try:
            y, _, _ = mod(x)
        finally:
            F.grid_sample = orig
        y.backward(dy)
        (_, grid_f, xs), (_, _, bias) = calls
        rec = {
            "meta": dict(B=B, H=H, W=W, n_heads=heads, n_groups=groups, stride=stride, ksize=ksize,
                         orf=orf, q_size=q_size, torch=torch.__version__),
            "x_sum": x.detach().double().sum().item(),
            "dy_sum": dy.double().sum().item(),
            "params": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "grads": {k: v.grad.clone() for k, v in mod.named_parameters()},
            "y": y.detach().clone(),
            "dx": x.grad.clone(),
            # grid is (x, y) ordered: store pos as (y, x) like the reference's `pos`
            "pos": grid_f.flip(-1).contiguous(),          # (B*G, Hk, Wk, 2)
            "xs": xs.contiguous(),                        # (B*G, Cg, Hk, Wk)
        }
        if big:
            rec["bias"] = bias.contiguous()               # (B*G, hg, HW, Ns)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            yb, _, _ = mod(x.detach())
        rec["y_autocast_bf16"] = yb.clone()
        # the reference's own autocast-bf16 backward, as a gap to its fp32 gradients
        # (max|a-b|/max|b| per tensor): the yardstick for the bf16 backward tests
        mod.zero_grad()
        xb = x.detach().clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yb2, _, _ = mod(xb)
        yb2.backward(dy.bfloat16())

        def gap(a, b):
            return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()

        def gap_l2(a, b):
            return ((a.double() - b.double()).norm() / b.double().norm()).item()

        rec["bf16_grad_gap"] = {"dx": gap(xb.grad, rec["dx"])}
        rec["bf16_grad_gap_l2"] = {"dx": gap_l2(xb.grad, rec["dx"])}
        for k, v in mod.named_parameters():
            rec["bf16_grad_gap"][k] = gap(v.grad, rec["grads"][k])
            rec["bf16_grad_gap_l2"][k] = gap_l2(v.grad, rec["grads"][k])
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(rec, path)
        print(f"{name}: y {tuple(y.shape)} |y|max {y.abs().max():.4f}  "
              f"clamped {(grid_f.abs() == 1).float().mean():.3f}  "
              f"oob {(grid_f.abs() > 1).float().mean():.3f}  "
              f"bf16 maxabs diff {(yb.float() - y).abs().max():.2e}  "
              f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
