"""One forward launch of the 7x7 depthwise kernel at the stage-0 and stage-2 'X'-mixer shapes between
cudaProfilerStart/Stop (the command ncu wraps; prints no bench value)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dat_segmentation_b200 import _cabi

lib = _cabi.lib()
p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
B = 16
for Cc, HW in ((64, 128), (256, 32)):
    x = torch.randn(B, HW, HW, Cc, device="cuda")
    y = torch.empty(B, HW, HW, Cc, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(Cc, 1, 7, 7, device="cuda") / 7
    b = torch.randn(Cc, device="cuda")
    nb = lib.dat_dwconv_workspace_bytes(B, HW, HW, Cc, 7)
    ws = torch.empty(nb, device="cuda", dtype=torch.uint8)
    for it in range(3):
        if it == 2:
            torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
        _cabi.check(lib.dat_dwconv_fwd(p(x), 0, p(w), p(b), p(y), None, 1, B, HW, HW, Cc, 7, 0, 0, p(ws), nb, st), "dw7")
        if it == 2:
            torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("profiled")
