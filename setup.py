"""Build of wildboar-b200 (counterpart of the reference's setup.py:102-127, which cythonizes its extensions).

The one native artefact is ``wildboar_b200/libwbcuda.so``: ``wildboar_b200/csrc/wb_cuda.cu`` compiled by nvcc for sm_100a
ONLY (``-gencode arch=compute_100a,code=sm_100a``), with ``-fmad=false`` -- required: every DP value is bit-equal to the
reference only without FMA contraction.  It is a plain C-ABI shared library (include/wb_cuda.h) that the package loads with
ctypes, not a CPython extension module, so the build step is a custom ``build_ext`` that runs the in-tree Makefile (make
rebuilds only when a source is newer) and ships the result, plus the public header, as package data.

    pip install --no-build-isolation .        # or: python setup.py build_ext --inplace

Environment: NVCC (default ``nvcc``), WILDBOAR_B200_SKIP_BUILD=1 to package an already built library as is.
"""
import os
import shutil
import subprocess

from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "wildboar_b200")


class build_cuda(build_ext):
    """Runs wildboar_b200/csrc/Makefile (nvcc, sm_100a) and places libwbcuda.so + wb_cuda.h in the package."""

    def get_ext_filename(self, fullname):
        # a ctypes-loaded library keeps its plain name (no CPython ABI tag)
        return os.path.join(*fullname.split(".")[:-1], "libwbcuda.so")

    def build_extension(self, ext):
        lib = os.path.join(PKG, "libwbcuda.so")
        if os.environ.get("WILDBOAR_B200_SKIP_BUILD") != "1":
            env = dict(os.environ)
            subprocess.check_call(["make", "-C", os.path.join(PKG, "csrc"), "../libwbcuda.so"], env=env)
        if not os.path.isfile(lib):
            raise RuntimeError("wildboar_b200/libwbcuda.so was not built (nvcc for sm_100a is required; there is no CPU build)")
        # the public C header travels with the package (wildboar_b200.get_include())
        inc = os.path.join(PKG, "include")
        os.makedirs(inc, exist_ok=True)
        shutil.copy2(os.path.join(ROOT, "include", "wb_cuda.h"), os.path.join(inc, "wb_cuda.h"))
        dst = self.get_ext_fullpath(ext.name)
        if os.path.abspath(dst) != os.path.abspath(lib):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copy2(lib, dst)
            dinc = os.path.join(os.path.dirname(dst), "include")
            os.makedirs(dinc, exist_ok=True)
            shutil.copy2(os.path.join(ROOT, "include", "wb_cuda.h"), os.path.join(dinc, "wb_cuda.h"))


setup(
    ext_modules=[Extension("wildboar_b200.libwbcuda", sources=["wildboar_b200/csrc/wb_cuda.cu"])],
    cmdclass={"build_ext": build_cuda},
)
