#!/bin/bash
# per-kernel durations of a config-5 shard run (cell_order_* kernels next to the walk)
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_ar.csv \
    python tools/kbench.py config5_shard > gpurun_out/ncu_r02_ar.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(l for l in open('gpurun_out/launches_r02_ar.csv') if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    k = r["Kernel Name"].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
for k, (n, us) in agg.items(): print("%-60s launches %3d  total %.1f us  mean %.1f us" % (k, n, us, us / n))
PY
