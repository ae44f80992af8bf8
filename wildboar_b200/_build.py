"""Build libwbcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libwbcuda.so")
STAMP = LIB + ".srchash"  # hash of the sources the library was built from (a copied tree does not keep modification times)


def _sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] if os.path.isdir(CSRC) else []
    for inc in (os.path.join(_HERE, "..", "include", "wb_cuda.h"), os.path.join(_HERE, "include", "wb_cuda.h")):
        if os.path.isfile(inc):
            srcs.append(inc)
            break
    return [s for s in srcs if os.path.isfile(s)]


def _source_hash():
    h = hashlib.sha256()
    for s in _sources():
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB):
        return True
    if os.path.isfile(STAMP):
        with open(STAMP) as f:
            return f.read().strip() != _source_hash()
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if force and os.path.exists(LIB):
        os.remove(LIB)
    if not needs_build():
        return LIB
    out = subprocess.run(["make", "-B", "-C", CSRC, "../libwbcuda.so"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError("building libwbcuda.so failed (nvcc, sm_100a)")
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return LIB
