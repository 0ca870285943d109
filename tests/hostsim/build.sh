#!/bin/bash
# builds the test-only CPU twin of the kernels (never loaded by the swinvox_b200 package)
set -e
cd "$(dirname "$0")"
g++ -O2 -march=native -fopenmp -std=c++17 -fPIC -shared -DSVX_HOSTSIM \
    svx_hostsim.cpp -x c++ ../../swinvox_b200/csrc/svx_api.cu -o libsvx_hostsim.so
