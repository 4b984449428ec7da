"""Turn gpurun_out/*.csv launch lists and *.ncu-rep captures into the small text summaries committed under profiles/.
    python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv > profiles/r01_launches_summary.txt
    python scripts/ncu_summary.py full gpurun_out/r01_full.ncu-rep > profiles/r01_full_summary.csv
"""
import collections
import csv
import subprocess
import sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {n} launches, {tot:.1f} us of kernel time (ncu-serialised, cold cache: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} n={v[0]:5d} total={v[1]:10.1f}us avg={v[1] / v[0]:8.2f}us share={v[1] / tot * 100:5.1f}%")
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct"]
    idx = [h.index(w) for w in want if w in h]
    w = csv.writer(sys.stdout)
    w.writerow([h[i] for i in idx])
    w.writerow([r[1][i] for i in idx])
    for row in r[2:]:
        w.writerow([row[i].split("(")[0] if h[i] == "Kernel Name" else row[i] for i in idx])
