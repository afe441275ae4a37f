"""SASS opcode evidence per kernel of libdat_b200.so (cuobjdump -sass; no GPU needed):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, legacy mma.sync -> HMMA,
plus registers per thread from the cubin.   usage: python tools/sass_summary.py > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dat_segmentation_b200", "libdat_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "MUFU", "HFMA2", "SYNCS", "BAR", "ATOM", "RED"]
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        funcs[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                funcs[cur][k] += 1
print("# SASS opcode summary of `libdat_b200.so` (sm_100a), one row per kernel\n")
print("`cuobjdump -sass` instruction counts (static).  UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, "
      "HMMA = legacy mma.sync.\n")
print("| kernel | regs | instr | " + " | ".join(KEYS) + " |\n|---|---:|---:|" + "---:|" * len(KEYS))
for name, c in funcs.items():
    d = demangle(name)
    d = re.sub(r"dat::\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d).replace("void ", "").replace("__nv_bfloat16", "bf16")
    print(f"| `{d[:80]}` | {regs.get(name, '')} | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
