#!/usr/bin/env bash
# One gpurun call: ncu of the band-register kernel behind the subsequence scan (msm: stride-1 windows; scaled_msm: materialised,
# interleaved windows).  Sections instead of --set full + sources: the reports have to fit gpurun's 64 MiB return limit.
set -x
mkdir -p gpurun_out
cat > /tmp/scan_probe.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import wildboar_b200 as wb
wb.set_devices([0])
n, T, ns, m = 2000, 512, 4, 64
Xs = np.cumsum(np.random.default_rng(8).standard_normal((n, T)), axis=-1)
rng = np.random.default_rng(9)
shp = [Xs[rng.integers(0, n), o:o + m].copy() for o in rng.integers(0, T - m, ns)]
metric = sys.argv[1]
for _ in range(2):
    d, i = wb.pairwise_subsequence_distance(shp, Xs, metric=metric, metric_params={"r": 0.1}, return_index=True)
print(metric, wb.last_stats())
PY
SECTIONS="--section SpeedOfLight --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats --section LaunchStats --section Occupancy --section InstructionStats"
for metric in ${METRICS:-msm scaled_msm}; do
  timeout 300 ncu $SECTIONS --clock-control none -k regex:k_band -s 5 -c 1 -f -o gpurun_out/band_$metric python /tmp/scan_probe.py $metric > gpurun_out/ncu_band_$metric.log 2>&1
  tail -3 gpurun_out/ncu_band_$metric.log
done
ls -la gpurun_out/*.ncu-rep
