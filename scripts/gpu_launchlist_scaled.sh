#!/usr/bin/env bash
# Launch list (ncu gpu__time_duration) of the scaled subsequence scans: where does a pass spend its time?
set -x
mkdir -p gpurun_out
cat > /tmp/scan_probe.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import wildboar_b200 as wb
wb.set_devices([0])
n, T, ns, m = 2000, 512, 16, 64
Xs = np.cumsum(np.random.default_rng(8).standard_normal((n, T)), axis=-1)
rng = np.random.default_rng(9)
shp = [Xs[rng.integers(0, n), o:o + m].copy() for o in rng.integers(0, T - m, ns)]
metric = sys.argv[1]
d, i = wb.pairwise_subsequence_distance(shp, Xs, metric=metric, metric_params={"r": 0.1}, return_index=True)
print(metric, wb.last_stats())
PY
for metric in ${METRICS:-scaled_adtw scaled_msm}; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$metric.csv python /tmp/scan_probe.py $metric > gpurun_out/ll_$metric.log 2>&1
  tail -1 gpurun_out/ll_$metric.log
done
