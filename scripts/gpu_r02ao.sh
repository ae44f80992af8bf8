#!/bin/bash
# round 2, call ao: neighbour-set mode (seeded thresholds for k > 1 where only the set of neighbours is consumed); full suite
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python scripts/probe_knn.py
timeout 300 python scripts/probe_cfg4.py | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
} 2>&1 | tee gpurun_out/r02ao.log
