#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED reference (wildboar, Cython) out of tree
# and installs only the resulting importable package into oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun).  Nothing under oracle/ is on the product path.
#
# The reference is a Python/Cython package: its hot path is generated C, so it cannot be
# compiled "from its own few source files" with plain gcc.  We therefore run cythonize +
# build_ext on a scratch COPY under /tmp (the reference tree is read-only) with the
# reference's default flags (-O2, no -march, no -ffast-math => no FMA contraction), and
# copy the built package (py + .so, no .pyx/.c sources) into oracle/_ref/wildboar.
#
# Usage: oracle/build_ref.sh [/root/reference]
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/wildboar" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi
if [ -f "$OUT/wildboar/distance/__init__.py" ] && ls "$OUT"/wildboar/distance/_elastic*.so >/dev/null 2>&1; then
  echo "build_ref: $OUT already built"; exit 0
fi
TMP="$(mktemp -d /tmp/wb_ref_build.XXXXXX)"
cp -r "$REF/src" "$REF/setup.py" "$REF/pyproject.toml" "$REF/README.md" "$REF/LICENSE" "$TMP/"
chmod -R u+w "$TMP"
# setuptools_scm is not installed: provide the version file it would have generated
echo 'version = "0.0.0+oracle"' > "$TMP/src/wildboar/version.py"
( cd "$TMP" && WILDBOAR_BUILD_NTHREADS=8 python setup.py build_ext --inplace -j 8 ) > "$TMP/build.log" 2>&1 \
  || { tail -50 "$TMP/build.log"; exit 1; }
rm -rf "$OUT"; mkdir -p "$OUT"
# copy package: python files + built extension modules only
( cd "$TMP/src" && find wildboar \( -name '*.py' -o -name '*.so' \) -print0 | xargs -0 cp --parents -t "$OUT" )
find "$OUT" -name '*.so' -exec strip --strip-unneeded {} + || true
rm -rf "$TMP"
PYTHONPATH="$OUT" python -c "from wildboar.distance import pairwise_distance; print('build_ref: reference importable from $OUT')"
