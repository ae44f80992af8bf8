"""CPU: host logic of the product package -- validation, shapes of errors, C-ABI exports."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol(wb):
    hdr = open(os.path.join(ROOT, "include", "wb_cuda.h")).read()
    names = set(re.findall(r"\b(wb_cuda_[a-z0-9_]+)\s*\(", hdr))
    assert {"wb_cuda_pairwise", "wb_cuda_pairwise_self", "wb_cuda_paired", "wb_cuda_argmin",
            "wb_cuda_pairwise_dev", "wb_cuda_last_error", "wb_cuda_device_count", "wb_cuda_fp64_peak"} <= names
    L = ctypes.CDLL(wb.library_path())
    for n in names:
        assert hasattr(L, n), n


def test_shim_argtypes_match_the_header_prototypes(wb):
    """Every ctypes binding of the shim has as many arguments as the C prototype in include/wb_cuda.h, pointer / integer /
    double in the same positions (a mismatch would be silent undefined behaviour at the ABI)."""
    from wildboar_b200 import _shim
    hdr = open(os.path.join(ROOT, "include", "wb_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|void|const char \*)\s*\*?(wb_cuda_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr))
    L = _shim.lib()
    checked = 0
    for name, args in protos.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        params = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        assert len(params) == len(fn.argtypes), (name, len(params), len(fn.argtypes))
        for prm, ct in zip(params, fn.argtypes):
            is_ptr = "*" in prm
            ct_ptr = hasattr(ct, "contents") or ct in (ctypes.c_void_p, ctypes.c_char_p)
            assert is_ptr == ct_ptr, (name, prm, ct)
            if not is_ptr:
                want = (ctypes.c_double if prm.startswith("double") else ctypes.c_int64 if prm.startswith("int64_t")
                        else ctypes.c_size_t if prm.startswith("size_t") else ctypes.c_int)
                assert ct is want, (name, prm, ct)
        checked += 1
    assert checked >= 15


def test_shim_structs_match_the_header(wb):
    """wb_stats / wb_params of the ctypes shim: same members, order and types as the structs of include/wb_cuda.h."""
    from wildboar_b200 import _shim
    hdr = open(os.path.join(ROOT, "include", "wb_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    ctype = {"double": ctypes.c_double, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "int": ctypes.c_int}
    for cname, cls in (("wb_stats", _shim.WbStats), ("wb_params", _shim.WbParams)):
        body = re.search(r"typedef struct " + cname + r"\s*\{(.*?)\}\s*" + cname + r"\s*;", hdr, flags=re.S).group(1)
        members = [(t, n) for t, n in re.findall(r"\b(double|int64_t|int32_t|int)\s+([a-z_0-9]+)\s*;", body)]
        assert [n for _, n in members] == [n for n, _ in cls._fields_], (cname, members, cls._fields_)
        assert [ctype[t] for t, _ in members] == [c for _, c in cls._fields_], cname


def test_no_gpu_means_error_not_fallback(wb):
    if wb.device_count() > 0:
        pytest.skip("a GPU is present")
    x = np.zeros((2, 8))
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        wb.pairwise_distance(x, x.copy(), metric="dtw")


def test_product_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "wildboar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


@pytest.mark.parametrize("metric,params,exc", [
    ("dtw", {"r": -0.1}, ValueError), ("dtw", {"r": 1.5}, ValueError), ("dtw", {"r": "a"}, TypeError),
    ("dtw", {"bogus": 1}, TypeError), ("wdtw", {"g": -1.0}, ValueError), ("lcss", {"epsilon": 0.0}, ValueError),
    ("edr", {"epsilon": -1.0}, ValueError), ("erp", {"g": -0.5}, ValueError), ("msm", {"c": -1.0}, ValueError),
    ("twe", {"penalty": -1.0}, ValueError), ("twe", {"stiffness": 0.0}, ValueError), ("twe", {"edit_penalty": 1.0}, TypeError),
])
def test_metric_param_validation(wb, metric, params, exc):
    # reference: tests/wildboar/distance/test_distance.py:809-848
    x = np.zeros((2, 8))
    with pytest.raises(exc):
        wb.pairwise_distance(x, x.copy(), metric=metric, metric_params=params)


def test_input_validation(wb):
    x = np.zeros((3, 8))
    with pytest.raises(ValueError, match="unsupported metric"):
        wb.pairwise_distance(x, x.copy(), metric="nope")
    bad = x.copy(); bad[1, 2] = np.nan
    with pytest.raises(ValueError, match="contains NaN"):
        wb.pairwise_distance(bad, x, metric="dtw")
    bad[1, 2] = np.inf
    with pytest.raises(ValueError, match="contains infinity"):
        wb.pairwise_distance(bad, x, metric="dtw")
    with pytest.raises(ValueError, match="same number of samples|broadcast"):
        wb.paired_distance(np.zeros((3, 8)), np.zeros((4, 8)), metric="dtw")
    with pytest.raises(ValueError, match="lower bound must be of shape"):
        wb.argmin_distance(x, x.copy(), metric="dtw", lower_bound=np.zeros((2, 2)))
    with pytest.raises(ValueError):
        wb.argmin_distance(x, x.copy(), k=0, metric="dtw")
    with pytest.raises(ValueError, match="dim must be"):
        wb.pairwise_distance(np.zeros((3, 2, 8)), np.zeros((3, 2, 8)), dim=5, metric="dtw")
    assert wb.pairwise_distance(np.zeros(8), metric="dtw") == 0.0  # 1-D self distance (DI:1238-1239)
    with pytest.warns(FutureWarning):
        wb.check_metric("lcss")(threshold=0.5)


def test_metric_objects_pickle(wb):
    import pickle
    for name, cls in wb._METRICS.items():
        m = cls()
        m2 = pickle.loads(pickle.dumps(m))
        assert type(m2) is cls and m2.__reduce__()[1] == m.__reduce__()[1] or name == "edr"


def test_precision_switch(wb, monkeypatch):
    """set_precision / WILDBOAR_CUDA_PRECISION stamp the C-ABI parameter block (no compute here)."""
    from wildboar_b200 import _shim
    p = wb.check_metric("dtw")(r=0.1)._params()
    assert _shim.apply_engine_override(p).precision == 0
    wb.set_precision("fp32")
    try:
        assert _shim.apply_engine_override(p).precision == 1 and wb.get_precision() == "fp32"
    finally:
        wb.set_precision(None)
    monkeypatch.setenv("WILDBOAR_CUDA_PRECISION", "fp32")
    assert _shim.apply_engine_override(p).precision == 1
    monkeypatch.setenv("WILDBOAR_CUDA_PRECISION", "bf16")
    import pytest
    with pytest.raises(ValueError):
        wb.get_precision()
    with pytest.raises(ValueError):
        wb.set_precision("fp16")


def test_package_installs_with_pip_and_binds_every_symbol(wb, tmp_path):
    """pyproject.toml + setup.py (the reference's packaging counterpart: setup.py:102-127, pyproject.toml:54-65):
    `pip install .` builds / ships libwbcuda.so + wb_cuda.h, and the installed copy -- imported from another working
    directory, without the repository on sys.path -- loads the library and binds every symbol of the header."""
    import subprocess
    import sys
    target = tmp_path / "site"
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-build-isolation", "--no-deps", "--no-index", "--quiet",
                        "--target", str(target), ROOT], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    code = (
        "import os, re, ctypes, wildboar_b200 as wb\n"
        "from wildboar_b200 import _shim\n"
        f"assert os.path.realpath(wb.__file__).startswith(os.path.realpath({str(target)!r})), wb.__file__\n"
        "hdr = open(os.path.join(wb.get_include(), 'wb_cuda.h')).read()\n"
        "L = _shim.lib()\n"
        "names = set(re.findall(r'\\b(wb_cuda_[a-z0-9_]+)\\s*\\(', hdr))\n"
        "assert len(names) >= 25 and all(hasattr(L, n) for n in names)\n"
        "assert os.path.dirname(wb.library_path()) == os.path.dirname(os.path.realpath(wb.__file__))\n"
        "print('ok', wb.__version__, len(names))\n"
    )
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env["PYTHONPATH"] = str(target)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_pinned_helpers_degrade_to_ordinary_memory_without_a_device(wb):
    """result_array / pinned_copy ask the library for page-locked memory and use ordinary numpy memory when there is none
    (no GPU here): the values and dtype are what the caller gave, whatever memory they live in."""
    from wildboar_b200 import _shim
    a = np.arange(300000, dtype=np.float64).reshape(300, 1000)
    b = wb.pinned_copy(a[:, ::2])            # non-contiguous input, > 1 MB
    assert b.dtype == np.float64 and b.flags.c_contiguous and np.array_equal(b, a[:, ::2])
    r = _shim.result_array((7, 9))
    assert r.shape == (7, 9) and r.dtype == np.float64
    assert _shim.lib().wb_cuda_host_alloc(0) in (None, 0) or wb.device_count() > 0
