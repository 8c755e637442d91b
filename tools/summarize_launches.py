"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary (markdown).
usage: python tools/summarize_launches.py gpurun_out/X_launches.csv profiles/NAME.md "title" """
import collections
import csv
import re
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += v_us
    tot += v_us
    n += 1
with open(dst, "w") as f:
    f.write(f"# {title}\n\n")
    f.write(f"Source: `{src}` (ncu `--metrics gpu__time_duration.sum --clock-control none`; per-launch times are "
            f"cold-cache and serialised -- compare SHARES, not absolutes).  {n} launches, {tot / 1e3:.3f} ms captured.\n\n")
    f.write("| kernel | launches | total ms | share | us / launch |\n|---|---:|---:|---:|---:|\n")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name[:90]}` | {c} | {t / 1e3:.3f} | {100 * t / tot:.1f}% | {t / c:.1f} |\n")
print("wrote", dst)
