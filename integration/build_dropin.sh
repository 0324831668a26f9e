#!/bin/bash
# Builds the drop-in proof artefacts from the reference tree (dev container only):
#   oracle/_ref/dropin_check            C++ driver using the PATCHED reference header
#   oracle/_ref/pyflagstats*.so         the reference's UNCHANGED python/libflagstats.pyx,
#                                       compiled against the patched header
#   oracle/_ref/samtools_caller         plain-C twin of the benchmark's "RAW SAMTOOLS" reader
#                                       (needs nothing from the reference tree)
# The reference sources are copied to a temp dir, patched there and compiled;
# nothing but binaries lands in the repo (oracle/_ref/ is git-ignored and ships
# to the GPU box with gpurun).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(dirname "$HERE")
REF=${REF:-/root/reference}
OUT=$ROOT/oracle/_ref
[ -f "$ROOT/libflagstats_b200/libflagstats_cuda.so" ] || python -m libflagstats_b200.build
mkdir -p "$OUT"
gcc -std=c99 -O2 -Wall -Wextra -I"$ROOT/include" "$HERE/samtools_caller.c" -o "$OUT/samtools_caller" \
    -L"$ROOT/libflagstats_b200" -lflagstats_cuda -Wl,-rpath,\$ORIGIN/../../libflagstats_b200
[ -f "$REF/libflagstats.h" ] || { echo "no reference tree at $REF: keeping prebuilt artefacts"; exit 0; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
cp "$REF/libflagstats.h" "$REF/libalgebra/libalgebra.h" "$REF/python/libflagstats.pyx" "$TMP/"
patch -s -d "$TMP" -p1 < "$HERE/libflagstats_h_cuda.patch"
LINK="-L$ROOT/libflagstats_b200 -lflagstats_cuda -Wl,-rpath,\$ORIGIN/../../libflagstats_b200"
g++ -std=c++11 -O2 -w -DFLAGSTATS_HAVE_CUDA -I"$TMP" -I"$ROOT/include" \
    "$HERE/dropin_check.cpp" -o "$OUT/dropin_check" $LINK
# the Cython wrapper, byte-identical .pyx (python/setup.py:29-35 builds the same extension)
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPINC=$(python -c "import numpy; print(numpy.get_include())")
EXT=$(python -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
(cd "$TMP" && cython -3 --module-name pyflagstats libflagstats.pyx -o pyflagstats.c 2>/dev/null)
gcc -O2 -w -fPIC -shared -DFLAGSTATS_HAVE_CUDA -DNPY_NO_DEPRECATED_API=0 -I"$TMP" -I"$ROOT/include" \
    -I"$PYINC" -I"$NPINC" "$TMP/pyflagstats.c" -o "$OUT/pyflagstats$EXT" $LINK
cmp -s "$REF/python/libflagstats.pyx" "$TMP/libflagstats.pyx" && echo "pyx unchanged"
ls -la "$OUT"
