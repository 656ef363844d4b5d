"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches,
total device time, share of the profiled region.  Per-launch times under ncu are cold-cache and
serialised, so SHARES are the quantity to read, not absolutes (B200_PROFILING.md).

    python tools/launch_share.py launches.csv [title]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    unit = r.get("Metric Unit", "ns")
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9,
          "s": 1.0, "second": 1.0}.get(unit, 1e-9)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void\s+", "", name).replace("tnr::<unnamed>::", "").replace("unnamed>::", "")
    rows.append((name.strip(), v))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, v in rows:
    agg[n][0] += 1
    agg[n][1] += v
GEMM = ("gemm_dmma", "ozaki", "cutlass", "gemm")
gemm_t = sum(t for n, (c, t) in agg.items() if any(g in n for g in GEMM))
print(f"# {title}\n")
print(f"{len(rows)} launches, {tot * 1e3:.2f} ms summed device time (ncu, serialised); "
      f"DMMA GEMM kernels: {100 * gemm_t / tot:.1f} % of it\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {t * 1e3:.3f} | {100 * t / tot:.1f} % |")
