"""Generates tests/golden/variants.pt from the UNMODIFIED reference module: the variant branches of
`DAttentionBaseline` that no shipped config uses (dat_blocks.py:57-59,84-99,156-157,164-167,185-197,
221-222) — `use_pe=False`, `no_off`, `dwc_pe`, `fixed_pe` (with and without a real resize), `log_cpb`.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_variants.py

Every case: reference module, default init under a fixed seed (offset conv / bias parameters widened so
the branches matter), CPU fp32 forward + backward; inputs regenerated from seeds by `variant_inputs()`.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

HC = 32
# name: (B, H, W, heads, groups, stride, ksize, orf, q_size, flags)
VARIANTS = {
    "no_pe": (2, 16, 16, 4, 2, 2, 5, -1, (14, 14), dict(use_pe=False)),
    "no_off": (2, 16, 12, 4, 2, 2, 5, -1, (14, 14), dict(no_off=True)),
    "no_off_s4_orf2": (1, 18, 22, 2, 1, 4, 7, 2, (14, 14), dict(no_off=True)),
    "dwc_pe": (2, 14, 18, 4, 4, 2, 5, 2, (14, 14), dict(dwc_pe=True)),
    "fixed_pe_same": (2, 12, 12, 2, 1, 2, 5, -1, (12, 12), dict(fixed_pe=True)),   # table == (HW, Ns): identity resize
    "fixed_pe_resize": (1, 16, 12, 4, 2, 2, 5, 1, (10, 10), dict(fixed_pe=True)),  # (100, 25) -> (192, 48)
    "log_cpb": (2, 16, 16, 4, 2, 2, 5, -1, (14, 14), dict(log_cpb=True)),
    "log_cpb_hg4_orf3": (1, 10, 14, 4, 1, 2, 3, 3, (14, 14), dict(log_cpb=True)),
}


def variant_inputs(name):
    B, H, W, heads = VARIANTS[name][:4]
    g = torch.Generator().manual_seed(5000 + sorted(VARIANTS).index(name))
    x = torch.randn(B, heads * HC, H, W, generator=g)
    dy = torch.randn(B, heads * HC, H, W, generator=g)
    return x, dy


def flags_of(name):
    f = dict(use_pe=True, dwc_pe=False, no_off=False, fixed_pe=False, log_cpb=False)
    f.update(VARIANTS[name][9])
    return f


def main():
    from oracle.ref_shim import import_reference
    blocks, _ = import_reference()
    out = {}
    for name, (B, H, W, heads, groups, stride, ksize, orf, q_size, _) in VARIANTS.items():
        f = flags_of(name)
        torch.manual_seed(11)
        mod = blocks.DAttentionBaseline(q_size, q_size, heads, HC, groups, 0.0, 0.0, stride, orf, f["use_pe"],
                                        f["dwc_pe"], f["no_off"], f["fixed_pe"], ksize, f["log_cpb"], 2)
        with torch.no_grad():
            mod.conv_offset[3].weight.mul_(2.0)
            if isinstance(mod.rpe_table, torch.nn.Parameter):
                mod.rpe_table.mul_(30.0)
        x, dy = variant_inputs(name)
        x.requires_grad_(True)
        y, _, _ = mod(x)
        y.backward(dy)
        out[name] = {
            "meta": dict(B=B, H=H, W=W, n_heads=heads, n_groups=groups, stride=stride, ksize=ksize, orf=orf,
                         q_size=q_size, flags=f, torch=torch.__version__),
            "x_sum": x.detach().double().sum().item(),
            "params": {k: v.detach().clone() for k, v in mod.state_dict().items()},
            "grads": {k: v.grad.clone() for k, v in mod.named_parameters() if v.grad is not None},
            "frozen": [k for k, v in mod.named_parameters() if not v.requires_grad],
            "y": y.detach().clone(), "dx": x.grad.clone(),
        }
        print(name, tuple(y.shape), f"|y|max {y.abs().max():.3f}", "params", list(out[name]["params"])[-3:],
              "frozen", len(out[name]["frozen"]))
    path = os.path.join(HERE, "variants.pt")
    torch.save(out, path)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
