import numpy as np

METRICS = ["dtw", "wdtw", "ddtw", "adtw", "lcss", "erp", "edr", "msm", "twe", "wddtw", "wlcss"]
NINE = ["dtw", "wdtw", "ddtw", "adtw", "msm", "twe", "erp", "lcss", "edr"]


def random_walks(n, T, seed):
    """Synthetic inputs of SURVEY 8d: cumulative sums of standard normals."""
    return np.cumsum(np.random.default_rng(seed).standard_normal((n, T)), axis=1)


def golden_cases(golden):
    """Yield (case, metric, extra_params, r, prefix) for every golden entry."""
    import ast
    extra = ast.literal_eval(str(golden["meta_extra"]))
    seen = set()
    for key in golden:
        if "|" not in key:
            continue
        case, metric, pi, r = key.split("|")[:4]
        pre = "|".join((case, metric, pi, r))
        if pre in seen:
            continue
        seen.add(pre)
        yield int(case), metric, (extra.get(metric, {}) if pi == "1" else {}), float(r), pre
