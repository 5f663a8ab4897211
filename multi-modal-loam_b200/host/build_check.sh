#!/bin/bash
# Compile and link the C++ adapters against libmmloam_b200.so (no GPU needed to build).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
g++ -O2 -std=c++17 -Wall -o "$HERE/host_check" "$HERE/host_check.cpp" -L"$HERE/.." -lmmloam_b200 \
    -L/usr/local/cuda/lib64 -Wl,-rpath,"$HERE/..":/usr/local/cuda/lib64 -lcudart
echo "built $HERE/host_check"
