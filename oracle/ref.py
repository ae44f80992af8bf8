"""TEST INFRASTRUCTURE: import the UNMODIFIED reference (wildboar) built into oracle/_ref.

`load()` returns the `wildboar.distance` module or None when oracle/_ref has not been built
(oracle/build_ref.sh needs /root/reference, which exists only in the build container; the
built oracle/_ref itself travels to the GPU box with the repo snapshot).
"""
import importlib
import os
import sys

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.isfile(os.path.join(_REF, "wildboar", "distance", "__init__.py"))


def load():
    if not available():
        return None
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    try:
        return importlib.import_module("wildboar.distance")
    except Exception:  # pragma: no cover - e.g. ABI mismatch on a different image
        return None
