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
__version__ = "0.2.0"


def get_include():
    """Directory that holds the C-ABI header ``wb_cuda.h`` (an installed package carries its own copy; a source checkout
    uses ``include/`` at the repository root)."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for d in (os.path.join(here, "..", "include"), os.path.join(here, "include")):  # a source checkout's header wins over a build copy
        if os.path.isfile(os.path.join(d, "wb_cuda.h")):
            return os.path.abspath(d)
    raise FileNotFoundError("wb_cuda.h not found")


from ._shim import device_count, get_precision, last_stats, library_path, pinned_copy, set_devices, set_precision  # noqa: F401

__all__ = [
    "pairwise_distance", "paired_distance", "argmin_distance", "check_metric",
    "pairwise_subsequence_distance", "paired_subsequence_distance", "subsequence_match", "paired_subsequence_match",
    "distance_profile", "argmin_subsequence_distance",
    "get_include", "pinned_copy", "device_count", "set_devices", "set_precision", "get_precision", "last_stats", "library_path",
]
