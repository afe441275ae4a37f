"""ncu launch list (gpu__time_duration.sum, --csv) -> compact per-kernel summary (markdown).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = name.replace("void ", "").replace("<unnamed>::", "")
    tot[name[:90]] += v
    cnt[name[:90]] += 1
T = sum(tot.values())
mine = sum(v for k, v in tot.items() if k.startswith("dat::"))
print(f"source: `{path}` — {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time "
      f"(cold-cache, serialised by ncu: compare shares, not absolutes)\n")
print(f"dat_b200 kernels: {mine / 1e3:.2f} ms = {100 * mine / T:.1f} % of the step\n")
print("| kernel | launches | total µs | share |\n|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"| `{k}` | {cnt[k]} | {v:.0f} | {100 * v / T:.1f} % |")
