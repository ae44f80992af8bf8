#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python bench.py --steps 1 --warmup 3 > gpurun_out/r02bm_bench.json 2> gpurun_out/r02bm_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02bm_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d["configs"]["cfg5"]))
PY
} 2>&1 | tee gpurun_out/r02bm.log
