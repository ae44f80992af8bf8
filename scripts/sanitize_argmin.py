"""DEV TOOLING: small argmin calls through every variant of the cascade (run under compute-sanitizer by scripts/gpu_r02bj.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from wildboar_b200.neighbors import KNeighborsClassifier
rw = lambda n, T, s: np.cumsum(np.random.default_rng(s).standard_normal((n, T)), axis=1)
os.environ["WILDBOAR_CUDA_SEED_MIN"] = "256"
os.environ["WILDBOAR_CUDA_ARGMIN_CHUNK"] = "160"
q, refs = rw(37, 131, 81), rw(1003, 131, 82)
for env in ({}, {"WILDBOAR_CUDA_LB_Q": "0"}, {"WILDBOAR_CUDA_LB_Q": "2"}, {"WILDBOAR_CUDA_LB_Q": "8", "WILDBOAR_CUDA_LB_RB": "3"}, {"WILDBOAR_CUDA_LB_BS": "8"},
            {"WILDBOAR_CUDA_PIPED_UPLOAD_KB": "100"}, {"WILDBOAR_CUDA_PIPED_UPLOAD_KB": "100", "WILDBOAR_CUDA_NO_STAGING": "1"}, {"WILDBOAR_CUDA_NO_SEED": "1"}):
    os.environ.update(env)
    for metric, p in (("dtw", {"r": 0.1}), ("adtw", {"r": 0.1, "p": 0.4}), ("ddtw", {"r": 0.1})):
        for k, srt in ((1, False), (3, False), (5, True)):
            wb.argmin_distance(q, refs, k=k, metric=metric, metric_params=p, sorted=srt, return_distance=True)
    print(env, wb.last_stats()["lb_keogh_pruned"], flush=True)
    for key in env: os.environ.pop(key)
clf = KNeighborsClassifier(5, metric="dtw", metric_params={"r": 0.1}).fit(refs, np.arange(len(refs)) % 3)
clf.predict(q)
print("sanitize run complete")
