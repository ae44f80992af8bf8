"""Build libwbcuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libwbcuda.so")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "wb_cuda.h")]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force=False, verbose=False):
    if force and os.path.exists(LIB):
        os.remove(LIB)
    if not needs_build():
        return LIB
    out = subprocess.run(["make", "-C", CSRC, "../libwbcuda.so"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError("building libwbcuda.so failed (nvcc, sm_100a)")
    return LIB
