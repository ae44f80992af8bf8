"""Optional: route the elastic metrics of an installed wildboar to the CUDA path.

``patch()`` wraps ``wildboar.distance.{pairwise,paired,argmin}_distance`` (and the same names in
``wildboar.distance._distance``) so that calls whose ``metric`` is one of the elastic metric
strings go to ``wildboar_b200``; every other metric (euclidean, callables, ...) keeps using
wildboar's own implementation -- those are different code paths of the reference, not a
fallback for this one.  ``unpatch()`` restores the originals.
"""
import functools
import importlib

from . import distance as _d

_ORIG = {}
_NAMES = ("pairwise_distance", "paired_distance", "argmin_distance")


def _wrap(orig, ours):
    @functools.wraps(orig)
    def wrapper(*args, **kwargs):
        metric = kwargs.get("metric", "euclidean")
        if isinstance(metric, str) and metric in _d._METRICS:
            return ours(*args, **kwargs)
        return orig(*args, **kwargs)
    wrapper.__wildboar_b200_original__ = orig
    return wrapper


def patch():
    """Install the wrappers; returns the list of patched callables' qualified names."""
    mods = [importlib.import_module("wildboar.distance"), importlib.import_module("wildboar.distance._distance")]
    done = []
    for mod in mods:
        for name in _NAMES:
            cur = getattr(mod, name)
            if hasattr(cur, "__wildboar_b200_original__"):
                continue
            _ORIG[(mod.__name__, name)] = cur
            setattr(mod, name, _wrap(cur, getattr(_d, name)))
            done.append(f"{mod.__name__}.{name}")
    return done


def unpatch():
    for (modname, name), orig in list(_ORIG.items()):
        setattr(importlib.import_module(modname), name, orig)
        del _ORIG[(modname, name)]
