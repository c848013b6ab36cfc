#!/bin/bash
# graph-cut-ransac_b200/host/build_driver.sh -- builds the batched host driver (gcr_driver.cpp) into
# host/_ref/libgcr_driver.so.  The driver's graph-cut step is the REFERENCE's own Boykov-Kolmogorov max-flow
# (energy.h / graph.h / graph.cpp / maxflow.cpp), compiled from the reference tree where it lies: the min-cut stays
# in the reference's host code (north star), nothing of it is copied into this repository.  Only runs where the
# reference tree exists; the .so travels to the GPU box with the snapshot (host/_ref/ is git-ignored).
set -e
cd "$(dirname "$0")"
REF=${GCR_REFERENCE_DIR:-/root/reference}/src/pygcransac/include
if [ ! -f "$REF/energy.h" ]; then
  # no reference tree (the GPU box): the prebuilt driver must have travelled with the snapshot -- fail loudly otherwise, the
  # Python / C++ entry points cannot work without it (the min-cut is the reference's own Boykov-Kolmogorov code, north star)
  if [ -f _ref/libgcr_driver.so ]; then echo "reference tree not present; using the prebuilt host/_ref/libgcr_driver.so"; exit 0; fi
  echo "ERROR: $REF/energy.h not found and no prebuilt host/_ref/libgcr_driver.so: set GCR_REFERENCE_DIR to a checkout of danini/graph-cut-ransac" >&2
  exit 1
fi
if [ ! -f ../libgcr_b200.so ]; then echo "build libgcr_b200.so first (python graph-cut-ransac_b200/build.py)"; exit 1; fi
mkdir -p _ref
g++ -O2 -DNDEBUG -std=c++17 -w -fPIC -shared -ffp-contract=off -I"$REF" -I../../include -I. \
    -o _ref/libgcr_driver.so gcr_driver.cpp \
    -L.. -lgcr_b200 -Wl,-rpath,'$ORIGIN/../..'
echo "built graph-cut-ransac_b200/host/_ref/libgcr_driver.so"
# the reference's C++ driver signatures (gcr_gcransac_python.h) called from a C++ translation unit, for the tests
g++ -O2 -std=c++17 -fPIC -shared -I../../include -I. -o _ref/libgcr_cpp_api_shim.so ../../tests/cpp/cpp_api_shim.cpp \
    -L_ref -lgcr_driver -L.. -lgcr_b200 -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../..'
echo "built graph-cut-ransac_b200/host/_ref/libgcr_cpp_api_shim.so"
