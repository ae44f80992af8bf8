"""Optional: route the elastic metrics of an installed wildboar to the CUDA path.

``patch()`` wraps ``wildboar.distance.{pairwise,paired,argmin}_distance`` and the subsequence family
(``pairwise_/paired_subsequence_distance``, ``subsequence_match``, ``paired_subsequence_match``, ``distance_profile``,
``argmin_subsequence_distance``; also the same names in ``wildboar.distance._distance``) so that calls whose ``metric`` is one
of the elastic metric strings go to ``wildboar_b200``; every other metric (euclidean, callables, ...) keeps using
wildboar's own implementation -- those are different code paths of the reference, not a
fallback for this one.  ``unpatch()`` restores the originals.
"""
import functools
import importlib

from . import distance as _d
from . import subsequence as _s

_ORIG = {}
_NAMES = ("pairwise_distance", "paired_distance", "argmin_distance")
# the subsequence family: routed when the (possibly `scaled_`) metric is one of the elastic subsequence metrics
_SUB_NAMES = ("pairwise_subsequence_distance", "paired_subsequence_distance", "subsequence_match", "paired_subsequence_match",
              "distance_profile", "argmin_subsequence_distance")


def _wrap(orig, ours):
    @functools.wraps(orig)
    def wrapper(*args, **kwargs):
        metric = kwargs.get("metric", "euclidean")
        if isinstance(metric, str) and metric in _d._METRICS:
            return ours(*args, **kwargs)
        return orig(*args, **kwargs)
    wrapper.__wildboar_b200_original__ = orig
    return wrapper


def _wrap_sub(name, orig, ours):
    @functools.wraps(orig)
    def wrapper(*args, **kwargs):
        metric = kwargs.get("metric")  # defaults are euclidean / mass: not elastic
        base = metric[len("scaled_"):] if isinstance(metric, str) and metric.startswith("scaled_") else metric
        if isinstance(base, str) and base in _s._SUBSEQUENCE_METRICS:
            return ours(*args, **kwargs)
        return orig(*args, **kwargs)
    wrapper.__wildboar_b200_original__ = orig
    return wrapper


# callers that bind the three functions by name at import time: KNN / KMeans / KMedoids
# (`distance/_neighbors.py:121-160, 262-283, 320-347`), MDS, silhouette, change-point segmentation, counterfactuals
_CALLER_MODULES = (
    "wildboar.distance._neighbors", "wildboar.distance._manifold", "wildboar.metrics._cluster", "wildboar.segment._base",
    "wildboar.explain.counterfactual._nice", "wildboar.explain.counterfactual._nn", "wildboar.explain.counterfactual._proto",
    # callers of the subsequence family: motif annotation, shapelet-forest counterfactuals, importances
    "wildboar.annotate._motifs", "wildboar.explain.counterfactual._sf", "wildboar.explain._importance",
)


def patch():
    """Install the wrappers; returns the list of patched callables' qualified names.

    Besides ``wildboar.distance`` and ``wildboar.distance._distance`` every already-imported (or known)
    ``wildboar.*`` module that holds a reference to one of the three functions is re-pointed, so the
    estimators built on them (KNeighborsClassifier, KMeans, KMedoids, MDS, silhouette ...) use the CUDA
    path for elastic metrics without any change."""
    import sys
    base = importlib.import_module("wildboar.distance._distance")
    importlib.import_module("wildboar.distance")
    for m in _CALLER_MODULES:
        try:
            importlib.import_module(m)
        except Exception:  # optional parts of wildboar that are not importable here
            pass
    originals = {name: getattr(base, name) for name in _NAMES + _SUB_NAMES}
    originals = {n: getattr(f, "__wildboar_b200_original__", f) for n, f in originals.items()}
    wrappers = {name: _wrap(originals[name], getattr(_d, name)) for name in _NAMES}
    wrappers.update({name: _wrap_sub(name, originals[name], getattr(_s, name)) for name in _SUB_NAMES})
    done = []
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == "wildboar" or modname.startswith("wildboar.")):
            continue
        for name in _NAMES + _SUB_NAMES:
            cur = getattr(mod, name, None)
            if cur is originals[name]:
                _ORIG[(modname, name)] = cur
                setattr(mod, name, wrappers[name])
                done.append(f"{modname}.{name}")
    return done


def unpatch():
    for (modname, name), orig in list(_ORIG.items()):
        setattr(importlib.import_module(modname), name, orig)
        del _ORIG[(modname, name)]
