#!/bin/bash
# round 2, call ai: straggler rule of the LB pass re-tuned for seeded thresholds; chunk width
mkdir -p gpurun_out
{
P0="WILDBOAR_CUDA_PIPED_UPLOAD_KB=0"
for sg in "2,63" "1,63" "0,63" "2,127" "1,127" "2,95" "4,63" "1,191"; do
  echo "== seeded, stragglers $sg"; env $P0 WILDBOAR_CUDA_LB_STRAG=$sg timeout 300 python scripts/probe_cfg4.py | tail -1
done
for sg in "0,63" "1,127"; do for c in 3200 6400; do
  echo "== seeded, stragglers $sg chunk $c"; env $P0 WILDBOAR_CUDA_LB_STRAG=$sg WILDBOAR_CUDA_ARGMIN_CHUNK=$c timeout 300 python scripts/probe_cfg4.py | tail -1
done; done
for sg in "2,63" "0,63" "1,127"; do
  echo "== NOT seeded, stragglers $sg"; env $P0 WILDBOAR_CUDA_NO_SEED=1 WILDBOAR_CUDA_LB_STRAG=$sg timeout 300 python scripts/probe_cfg4.py | tail -1
done
for sg in "2,63" "0,63" "1,127"; do
  echo "== piped default, stragglers $sg"; env WILDBOAR_CUDA_LB_STRAG=$sg timeout 300 python scripts/probe_cfg4.py | tail -1
  echo "== 1 query, stragglers $sg"; env WILDBOAR_CUDA_LB_STRAG=$sg timeout 300 python scripts/probe_cfg4.py 1 | tail -1
done
} 2>&1 | tee gpurun_out/r02ai.log
