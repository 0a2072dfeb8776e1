#!/bin/bash
# Build a bounds-checked variant of the library next to the normal one (device-side asserts on
# every index of the cell-major path) and print its path; run tests against it with
#   LINTSAMPLER_B200_LIB=$(bash scripts/checked_build.sh) python -m pytest tests -m gpu -k cellmajor
set -e
cd "$(dirname "$0")/.."
cp lintsampler_b200/_lib/liblintsampler_b200.so /tmp/lint_normal.so
LINT_EXTRA_NVCC="-DLINT_CHECKED" python -m lintsampler_b200._build > /dev/null
mkdir -p scripts/ubench
cp lintsampler_b200/_lib/liblintsampler_b200.so scripts/ubench/lib_CHECKED.so
python -m lintsampler_b200._build > /dev/null          # back to the normal build
echo "$PWD/scripts/ubench/lib_CHECKED.so"
