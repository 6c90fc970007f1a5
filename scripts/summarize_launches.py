"""ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> markdown table: launches, total ms and share per kernel.
    python scripts/summarize_launches.py gpurun_out/x/launches.csv "title" > profiles/rN_launches_x.summary.md"""
import collections
import csv
import re
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit in ("nsecond", "ns") else v / 1e3 if unit in ("usecond", "us") else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += ms
    tot += ms
print(f"# {title}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv` (eager launches; per-launch times are "
      "cold-cache and serialised, so SHARES -- not absolutes -- are comparable with the CUDA-graph step).\n")
print(f"Total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches.\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |")
