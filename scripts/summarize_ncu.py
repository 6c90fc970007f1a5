"""`ncu --set full` reports (.ncu-rep) -> one markdown table row per captured launch (read offline with `ncu -i ... --page raw --csv`).
    python scripts/summarize_ncu.py title a.ncu-rep [b.ncu-rep ...] > profiles/rN_x.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %")]
title, files = sys.argv[1], sys.argv[2:]
print(f"# {title}\n")
print("Read from the committed-elsewhere `.ncu-rep` captures with `ncu -i <rep> --page raw --csv` (capture: `ncu --set full --clock-control none "
      "--import-source on`). Times under the profiler are cold-cache; CUDA-event timings are quoted in DESIGN.md.\n")
print("| capture | kernel | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
for f in files:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"| {f} | (unreadable) |")
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        cells = []
        for k, _ in COLS:
            v = d.get(k, "")
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append((v + " " + u.get(k, "")).strip())
        name = d.get("Kernel Name", "").replace("void ", "").split("(")[0]
        print(f"| {f.split('/')[-1]} | `{name}` | " + " | ".join(cells) + " |")
