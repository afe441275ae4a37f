"""`ncu -i X.ncu-rep --page raw --csv` of `python bench.py --roofline-only` -> profiles/r02_ncu_traffic.json: DRAM bytes per
launch (dram__bytes_read.sum + dram__bytes_write.sum) of the persistent GEMM and the attention forward at the bench's
roofline shapes, taken with the same L2 flush before the launch as the timing.   usage: summarize_ncu_traffic.py raw.csv"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
out = {}
for r in data:
    name = r[hdr.index("Kernel Name")]
    key = "gemm" if "gemm_tc_persistent" in name else ("attention" if "attn_fwd_tc" in name else None)
    if key is None:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    d = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
    out[key] = {"dram_bytes": int(tot), "kernel": name.split("(")[0][-60:], "ncu_duration": f"{d} {units[hdr.index('gpu__time_duration.sum')]}",
                "source": "ncu --set full --clock-control none of `python bench.py --roofline-only` (L2 flushed before the launch), "
                          "last profiled launch of the kernel"}
json.dump(out, open("profiles/r02_ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
