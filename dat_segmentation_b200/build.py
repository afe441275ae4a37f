"""Builds csrc/*.cu into the in-tree C-ABI shared library `libdat_b200.so`
(nvcc, sm_100a only).  No torch involvement: the library depends on the CUDA
runtime alone.  `python -m dat_segmentation_b200.build [--force]`."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
# DAT_B200_BUILD_TAG=x builds a second variant (build_x/, libdat_b200_x.so) next to the product library: A/B timing only
_TAG = os.environ.get("DAT_B200_BUILD_TAG", "")
OBJ = os.path.join(PKG, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(PKG, "libdat_b200" + ("_" + _TAG if _TAG else "") + ".so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# no --use_fast_math: reference points / tap indices must be bit-exact (IEEE div, no ftz)
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("DAT_B200_BUILD_DEFS", "").split()     # e.g. -DDAT_PDL_NO_EARLY_TRIGGER (A/B builds)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG), "include", "dat_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            cmd = [nvcc, *ARCH, *FLAGS, "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out)
    if jobs or force or not os.path.exists(LIB):
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
