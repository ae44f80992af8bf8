"""wildboar_b200: B200-native (sm_100a) elastic distances behind wildboar's distance API.

Public surface (mirrors ``wildboar.distance``; reference: src/wildboar/distance/_distance.py):

    pairwise_distance, paired_distance, argmin_distance, check_metric,
    pairwise_subsequence_distance, paired_subsequence_distance, subsequence_match, paired_subsequence_match,
    distance_profile (elastic subsequence metrics)

The compute path is ``libwbcuda.so`` (hand-written CUDA, include/wb_cuda.h).  There is no CPU
fallback: without the library or without a B200 the calls raise.
"""
from .distance import (  # noqa: F401
    _METRICS,
    argmin_distance,
    check_metric,
    paired_distance,
    pairwise_distance,
)
from .subsequence import (  # noqa: F401
    argmin_subsequence_distance,
    distance_profile,
    paired_subsequence_distance,
    paired_subsequence_match,
    pairwise_subsequence_distance,
    subsequence_match,
)
from ._shim import device_count, get_precision, last_stats, library_path, set_devices, set_precision  # noqa: F401

__all__ = [
    "pairwise_distance", "paired_distance", "argmin_distance", "check_metric",
    "pairwise_subsequence_distance", "paired_subsequence_distance", "subsequence_match", "paired_subsequence_match",
    "distance_profile", "argmin_subsequence_distance",
    "device_count", "set_devices", "set_precision", "get_precision", "last_stats", "library_path",
]
