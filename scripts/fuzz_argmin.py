"""DEV TOOLING: randomised argmin configurations (seeding, neighbour-set mode via sorted=True, pipelined upload, LB-pass variants,
duplicated references / exact matches, odd lengths) against the CPU oracle.  Usage: python scripts/fuzz_argmin.py [n_cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wildboar_b200 as wb
from oracle import oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
wb.set_devices([0])
KNOBS = ("WILDBOAR_CUDA_SEED_MIN", "WILDBOAR_CUDA_ARGMIN_CHUNK", "WILDBOAR_CUDA_PIPED_UPLOAD_KB", "WILDBOAR_CUDA_LB_Q", "WILDBOAR_CUDA_LB_BS",
         "WILDBOAR_CUDA_LB_RB", "WILDBOAR_CUDA_NO_SEED", "WILDBOAR_CUDA_LB_STRAG")
bad = 0
t_start = time.time()
for case in range(n_cases):
    for k_ in KNOBS:
        os.environ.pop(k_, None)
    T = int(rng.choice([2, 3, 5, 8, 9, 16, 31, 64, 100, 131, 200]))
    r = float(rng.choice([0.0, 0.05, 0.1, 0.3, 1.0]))
    nq = int(rng.integers(1, 70)); ny = int(rng.integers(256, 3000))
    k = int(rng.choice([1, 1, 2, 3, 5, 8, 9]))
    metric = str(rng.choice(["dtw", "dtw", "ddtw", "adtw"]))
    params = {"r": r}
    if metric == "adtw":
        params["p"] = float(rng.choice([0.0, 0.2, 1.5]))
    scale = float(rng.choice([1.0, 1e-3, 1e4])); offset = float(rng.choice([0.0, 0.0, 1e6]))
    q = np.cumsum(rng.standard_normal((nq, T)), axis=1) * scale + offset
    refs = np.cumsum(rng.standard_normal((ny, T)), axis=1) * scale + offset
    if rng.random() < 0.6:      # exact matches and duplicated references (ties)
        for _ in range(int(rng.integers(1, 6))):
            a, b = int(rng.integers(0, ny)), int(rng.integers(0, ny))
            if rng.random() < 0.5:
                refs[a] = q[int(rng.integers(0, nq))]
            refs[b] = refs[a]
    env = {"WILDBOAR_CUDA_SEED_MIN": "256"}
    if rng.random() < 0.6: env["WILDBOAR_CUDA_ARGMIN_CHUNK"] = str(int(rng.choice([32, 96, 160, 512, 4096])))
    if rng.random() < 0.4: env["WILDBOAR_CUDA_PIPED_UPLOAD_KB"] = str(int(rng.choice([1, 8, 64, 300])))
    if rng.random() < 0.5: env["WILDBOAR_CUDA_LB_Q"] = str(int(rng.choice([0, 2, 4, 8])))
    if rng.random() < 0.3: env["WILDBOAR_CUDA_LB_BS"] = "8"
    if rng.random() < 0.3: env["WILDBOAR_CUDA_LB_RB"] = str(int(rng.choice([1, 3, 16, 40])))
    if rng.random() < 0.2: env["WILDBOAR_CUDA_NO_SEED"] = "1"
    if rng.random() < 0.3: env["WILDBOAR_CUDA_LB_STRAG"] = str(rng.choice(["0,0", "1,15", "4,31", "2,63"]))
    os.environ.update(env)
    srt = bool(rng.random() < 0.5)
    kk = min(k, ny)
    oi, od = O.argmin(metric, q, refs, k=kk, n_jobs=0, **params)
    if srt:
        order = np.argsort(od, axis=1, kind="stable")
        oi, od = np.take_along_axis(oi, order, axis=1), np.take_along_axis(od, order, axis=1)
    try:
        idx, dist = wb.argmin_distance(q, refs, k=k, metric=metric, metric_params=params, sorted=srt, return_distance=True)
        ok = np.array_equal(idx, oi) and np.array_equal(dist, od)
    except Exception as e:   # noqa: BLE001
        ok = False
        print("EXCEPTION", repr(e))
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(T=T, r=r, nq=nq, ny=ny, k=k, metric=metric, params=params, sorted=srt, scale=scale, offset=offset), env, flush=True)
# adversarial sets: every distance equal (constant / identical series), zeros, queries that ARE references, steps of one ulp
for name, mk in (("identical refs", lambda: (np.cumsum(rng.standard_normal((9, 40)), axis=1), np.tile(np.cumsum(rng.standard_normal((1, 40)), axis=1), (700, 1)))),
                 ("all zeros", lambda: (np.zeros((5, 33)), np.zeros((600, 33)))),
                 ("constant series", lambda: (np.full((7, 64), 3.25), np.repeat(np.arange(500.0)[:, None] % 7, 64, axis=1))),
                 ("queries are references", lambda: (lambda r: (r[::37].copy(), r))(np.cumsum(rng.standard_normal((900, 50)), axis=1))),
                 ("one-ulp neighbours", lambda: (lambda b: (b[:3].copy(), np.stack([np.nextafter(b[i % 3], np.inf if i % 2 else -np.inf) for i in range(400)])))(np.cumsum(rng.standard_normal((3, 48)), axis=1)))):
    for k_ in KNOBS:
        os.environ.pop(k_, None)
    os.environ["WILDBOAR_CUDA_SEED_MIN"] = "256"
    q, refs = mk()
    for metric in ("dtw", "ddtw", "adtw"):
        for k, srt in ((1, False), (3, False), (3, True), (8, True)):
            params = {"r": 0.1} if metric != "adtw" else {"r": 0.1, "p": 0.5}
            oi, od = O.argmin(metric, q, refs, k=k, n_jobs=0, **params)
            if srt:
                order = np.argsort(od, axis=1, kind="stable")
                oi, od = np.take_along_axis(oi, order, axis=1), np.take_along_axis(od, order, axis=1)
            idx, dist = wb.argmin_distance(q, refs, k=k, metric=metric, metric_params=params, sorted=srt, return_distance=True)
            if not (np.array_equal(idx, oi) and np.array_equal(dist, od)):
                bad += 1
                print("MISMATCH adversarial", name, metric, k, srt, flush=True)
print(f"fuzz: {n_cases} cases + adversarial sets, {bad} mismatches, {time.time() - t_start:.0f} s")
sys.exit(1 if bad else 0)
