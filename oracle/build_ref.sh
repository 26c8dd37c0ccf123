#!/usr/bin/env bash
# Builds the UNMODIFIED reference (seung-lab/connected-components-3d, cc3d 4.x) into oracle/_ref/.
# Sources are read where they lie under /root/reference (never copied into the repo); only generated
# outputs (the cythonized .cpp and the extension .so) land in oracle/_ref/, which is git-ignored.
# The result is imported as the top-level module `fastcc3d` (tests/, bench.py --impl reference).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
PY=${PYTHON:-python3}
SUFFIX=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
# the reference's own test file travels with the build (git-ignored): tests/test_reference_suite.py runs it against the drop-in
if [ -f "$REF/automated_test.py" ]; then cp -f "$REF/automated_test.py" "$OUT/automated_test.py" && chmod u+w "$OUT/automated_test.py"; fi
# the crackle-v0 benchmark volume itself (3 MB data file): input of the GPU crackle decoder test
if [ -f "$REF/benchmarks/connectomics.npy.ckl.gz" ]; then cp -f "$REF/benchmarks/connectomics.npy.ckl.gz" "$OUT/connectomics.npy.ckl.gz" && chmod u+w "$OUT/connectomics.npy.ckl.gz"; fi
if [ -f "$OUT/fastcc3d$SUFFIX" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "oracle/_ref/fastcc3d$SUFFIX already built"; exit 0
fi
if [ ! -d "$REF/cc3d" ]; then
  echo "reference tree $REF not present; cannot build oracle/_ref" >&2; exit 3
fi
$PY -m cython -3 --cplus "$REF/cc3d/fastcc3d.pyx" -o "$OUT/fastcc3d.cpp"
g++ -std=c++17 -O3 -fPIC -shared -w -I "$REF/cc3d" \
  -I "$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')" \
  -I "$($PY -c 'import numpy;print(numpy.get_include())')" \
  "$OUT/fastcc3d.cpp" -o "$OUT/fastcc3d$SUFFIX"
rm -f "$OUT/fastcc3d.cpp"
echo "built $OUT/fastcc3d$SUFFIX"
