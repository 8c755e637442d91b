"""Summarise an `ncu --set full` report (.ncu-rep) as a markdown table of the counters the roofline discussion uses.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.md   (runs `ncu -i ... --page raw --csv` here; no GPU)"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"), ("sm__cycles_active.avg", "SM cycles active (avg)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {title}\n")
print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`), read with `ncu -i ... --page raw --csv`.\n")
print("| metric | " + " | ".join(f"launch {k}" for k in range(len(data))) + " |")
print("|---|" + "---:|" * len(data))
for key, label in WANT:
    if key not in idx:
        continue
    i = idx[key]
    cells = []
    for r in data:
        v = r[i]
        if key == "Kernel Name":
            v = "`" + v.split("(")[0].replace("void ", "")[:60] + "`"
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:,.2f}".rstrip("0").rstrip(".") + (" " + units[i] if units[i] else "")
            except ValueError:
                pass
        cells.append(v)
    print(f"| {label} | " + " | ".join(cells) + " |")
