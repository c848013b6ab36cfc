#!/bin/bash
# oracle/build_ref.sh -- compiles the reference's OWN headers, where they lie under /root/reference, through the
# container shims (oracle/shim) into oracle/_ref/libgcr_ref.so.  Only runs where /root/reference exists (the build
# container); the .so travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
# Flags = the reference's CMakeLists.txt:37 (-O3 -ftree-vectorize -funroll-loops) + -flto so the restated residuals
# inline into the reference's scoring loop exactly as its own header-only squaredResidual would.
set -e
cd "$(dirname "$0")"
REF=${GCR_REFERENCE_DIR:-/root/reference}/src/pygcransac/include
if [ ! -f "$REF/GCRANSAC.h" ]; then echo "reference tree not present; skipping oracle/_ref"; exit 0; fi
mkdir -p _ref
g++ -O3 -ftree-vectorize -funroll-loops -ffp-contract=off -flto -std=c++17 -w -fPIC -shared \
    -Ishim -I"$REF" -I. \
    -o _ref/libgcr_ref.so ref_glue.cpp gcr_oracle_*.cpp "$REF/graph.cpp" "$REF/maxflow.cpp"
echo "built oracle/_ref/libgcr_ref.so"
# the same driver with the CUDA scoring adapter as its template argument (links the product library)
PKG=../graph-cut-ransac_b200
if [ -f "$PKG/libgcr_b200.so" ]; then
g++ -O2 -std=c++17 -w -fPIC -shared -ffp-contract=off \
    -Ishim -I"$REF" -I. -I../include -I"$PKG/host" \
    -o _ref/libgcr_ref_cuda.so ref_glue_cuda.cpp gcr_oracle_*.cpp "$REF/graph.cpp" "$REF/maxflow.cpp" \
    -L"$PKG" -lgcr_b200 -Wl,-rpath,'$ORIGIN/../../graph-cut-ransac_b200'
echo "built oracle/_ref/libgcr_ref_cuda.so"
fi
