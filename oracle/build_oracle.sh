#!/bin/bash
# oracle/build_oracle.sh -- builds the self-contained CPU restatement (no /root/reference needed) into
# oracle/libgcr_oracle.so.  Flags follow the reference's CMakeLists.txt:37 (-O3 -ftree-vectorize -funroll-loops)
# plus -ffp-contract=off so FP64 results do not depend on FMA contraction.
set -e
cd "$(dirname "$0")"
g++ -O3 -ftree-vectorize -funroll-loops -ffp-contract=off -std=c++17 -fPIC -shared -fopenmp \
    -o libgcr_oracle.so gcr_oracle_*.cpp
echo "built oracle/libgcr_oracle.so"
