"""`ncu -i X.ncu-rep --page raw --csv` -> one markdown row per profiled launch with the metrics the
design discussion uses.  usage: summarize_ncu_raw.py raw1.csv [raw2.csv ...] > profiles/rNN_ncu_kernels.md"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "µs", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
        ("dram__bytes_write.sum", "DRAM wr MB", 1.0), ("smsp__inst_executed.sum", "warp-inst M", 1e-6),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0)]
UNIT_SCALE = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3}
print("| kernel | grid | " + " | ".join(c[1] for c in COLS) + " | top stalls (warps per issue) |")
print("|---|---|" + "---:|" * len(COLS) + "---|")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("void ", "").replace("<unnamed>::", "")
        name = name.replace("__nv_bfloat16", "bf16")
        vals = []
        for key, _, sc in COLS:
            i = hdr.index(key)
            v = float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0) * sc
            vals.append(f"{v:.1f}")
        stalls = []
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
            if m and m.group(1) != "selected":
                stalls.append((float(r[i].replace(",", "") or 0), m.group(1)))
        top = ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:3])
        print(f"| `{name[:70]}` | {r[hdr.index('Grid Size')]} | " + " | ".join(vals) + f" | {top} |")
