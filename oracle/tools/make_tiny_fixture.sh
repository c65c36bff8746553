#!/bin/bash
# TEST INFRASTRUCTURE.  Regenerates /tmp/euler (the fixture every reference test hard-codes,
# SURVEY.md section 4) with the reference's OWN converter, run from a scratch copy under /tmp
# (the reference tree is read-only and the tools need libcommon.so / libeuler_util.so beside util.py).
# Only runs where /root/reference exists.  Output dir: ${1:-/tmp/euler}
set -euo pipefail
REF=${REF:-/root/reference}
OUT=${1:-/tmp/euler}
PKG=$(mktemp -d /tmp/euler_tools_pkg.XXXXXX)
mkdir -p "$PKG/euler"
cp -r "$REF/euler/tools" "$PKG/euler/tools"
: > "$PKG/euler/__init__.py"
CXX="g++ -std=c++11 -O2 -fPIC -include cstdint -D_GLIBCXX_USE_CXX11_ABI=0 -I$REF"
$CXX -shared -o "$PKG/euler/tools/libcommon.so" "$REF/euler/common/hash.cc"
$CXX -shared -o "$PKG/euler/tools/libeuler_util.so" "$REF/euler/util/python_api.cc" "$REF/euler/common/hash.cc"
rm -rf "$OUT"; mkdir -p "$OUT"
# 3 args = no index step (json2partindex.py:205 needs py2 `unicode`; Index/ is out of scope)
PYTHONPATH="$PKG" python "$PKG/euler/tools/generate_euler_data.py" "$REF/tools/test_data/graph.json" "$OUT" 2 >/dev/null
rm -rf "$PKG"
find "$OUT" -type f | sort
