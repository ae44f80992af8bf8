"""TEST INFRASTRUCTURE: ctypes wrapper around tests/hostsim/_hostsim.so (device engines compiled for the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hostsim.so")
_ROOT = os.path.join(_HERE, "..", "..")


class WbParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("r", "g", "p", "c", "epsilon", "penalty", "stiffness")] + [
        ("engine", C.c_int32), ("precision", C.c_int32)]


_SO_COOP = os.path.join(_HERE, "_hostsim_coop.so")
_CSRC = os.path.join(_ROOT, "wildboar_b200", "csrc")
_UNITS = {
    _SO: ("hostsim.cpp", ("metrics.cuh", "engine_strip.cuh", "engine_rowscan.cuh", "engine_band.cuh", "dispatch.cuh", "prep.hpp")),
    _SO_COOP: ("hostsim_coop.cpp", ("metrics.cuh", "engine_coop.cuh", "dispatch.cuh", "prep.hpp")),
}


def build(force=False):
    """Compile the two host-simulation objects (in parallel) when a source is newer than the object."""
    procs = []
    for so, (main, deps) in _UNITS.items():
        srcs = [os.path.join(_HERE, main)] + [os.path.join(_CSRC, f) for f in deps]
        if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            procs.append(subprocess.Popen(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                                           "-Wno-unknown-pragmas", "-o", so, srcs[0]]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("building the host simulation failed")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.hostsim_pair.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(WbParams), dp, C.c_int64, dp, C.c_int64,
                                      C.c_int, C.c_double, C.c_int, C.c_int, dp, dp]
    return _lib


_lib_coop = None


def lib_coop():
    global _lib_coop
    if _lib_coop is None:
        build()
        _lib_coop = C.CDLL(_SO_COOP)
    return _lib_coop


def pair(engine, W, metric_id, params, x, y, ea=0, min_dist_raw=float("inf"), ns_extra=0, bs=1):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = C.c_double(0)
    mm = C.c_double(0)
    dp = C.POINTER(C.c_double)
    rc = lib().hostsim_pair(engine, W, metric_id, C.byref(params), x.ctypes.data_as(dp), len(x), y.ctypes.data_as(dp),
                            len(y), ea, min_dist_raw, ns_extra, bs, C.byref(out), C.byref(mm))
    return rc, out.value, mm.value


def coop_pair(W, G, metric_id, params, x, y):
    """(rc, distance) of one pair through the cooperative engine, G lanes emulated in lockstep (rc 1: geometry not
    covered by this (W, G))."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = C.c_double(0)
    dp = C.POINTER(C.c_double)
    f = lib_coop().hostsim_coop_pair
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(WbParams), dp, C.c_int64, dp, C.c_int64, dp]
    rc = f(W, G, metric_id, C.byref(params), x.ctypes.data_as(dp), len(x), y.ctypes.data_as(dp), len(y), C.byref(out))
    return rc, out.value


def inc_window_stats(x, m):
    """(mean, std) of every length-m window of the 1-D series x through the device code path compiled for the host."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    nw = len(x) - m + 1
    mean, std = np.empty(nw), np.empty(nw)
    dp = C.POINTER(C.c_double)
    f = lib().hostsim_inc_window_stats
    f.argtypes = [dp, C.c_int64, C.c_int64, dp, dp]
    f.restype = None
    f(x.ctypes.data_as(dp), len(x), m, mean.ctypes.data_as(dp), std.ctypes.data_as(dp))
    return mean, std


def interleave_roundtrip(a):
    """Mismatches after interleaving the rows of `a` 32 at a time and reading them back as the YS = 32 kernels do."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    f = lib().hostsim_interleave_roundtrip
    f.argtypes = [C.POINTER(C.c_double), C.c_longlong, C.c_int]
    f.restype = C.c_longlong
    return int(f(a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0], a.shape[1]))
