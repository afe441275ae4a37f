"""Phase breakdown of the tensor-core attention backward (dat_debug_attn_bwd_timing) for one block
fwd+bwd at each DAT-T++ stage shape, B = 16, bf16.  The counters are compiled in only with -DDAT_ATTN_BWD_PROFILE:
  DAT_B200_BUILD_TAG=prof DAT_B200_BUILD_DEFS=-DDAT_ATTN_BWD_PROFILE python -m dat_segmentation_b200.build
  DAT_B200_LIB=$PWD/dat_segmentation_b200/libdat_b200_prof.so python tools/attn_bwd_phases.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi
from dat_segmentation_b200.dattention import DAttentionBaseline

lib = _cabi.lib()
STAGES = [(128, 2, 1, 8, 9, 56), (64, 4, 2, 4, 7, 28), (32, 8, 4, 2, 5, 14), (16, 16, 8, 1, 3, 7)]   # tables of the 224-px config
B = 16
out = (C.c_uint64 * 8)()
for stage, (H, heads, groups, stride, ksize, qs) in enumerate(STAGES):
    torch.manual_seed(0)
    m = DAttentionBaseline((qs, qs), (qs, qs), heads, 32, groups, 0.0, 0.0, stride, -1, True, False, False,
                           False, ksize, False, stage).cuda()
    x = torch.randn(B, H, H, heads * 32, device="cuda").permute(0, 3, 1, 2).requires_grad_(True)
    for it in range(3):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x)[0]
        if it == 2:
            torch.cuda.synchronize()
            lib.dat_debug_attn_bwd_timing(out)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
    lib.dat_debug_attn_bwd_timing(out)
    v = list(out)
    n = max(v[6], 1)
    tot = v[0] / n
    names = ["tile loop", "wait S/dP", "score loop", "dpos colsum", "dQ wait+store", "tile setup"]
    print(f"stage {stage}: {n} CTAs, {tot:.0f} cycles per CTA in the tile loop: " +
          ", ".join(f"{nm} {100 * v[i] / max(v[0], 1):.1f}%" for i, nm in enumerate(names) if i > 0))
