#!/bin/bash
# oracle/build_ref_solvers.sh -- compiles the reference's OWN solver / estimator headers, where they lie under /root/reference,
# over oracle/shim_solvers (Eigen's API in front of oracle/linalg.h; cv::Mat container) into oracle/_ref/libgcr_ref_solvers.so.
# Only runs where /root/reference exists; the .so travels with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
set -e
cd "$(dirname "$0")"
REF=${GCR_REFERENCE_DIR:-/root/reference}/src/pygcransac/include
if [ ! -f "$REF/estimators/solver_p3p.h" ]; then echo "reference tree not present; skipping oracle/_ref/libgcr_ref_solvers.so"; exit 0; fi
mkdir -p _ref
# relative_pose/*.cpp: the bundle adjustment behind the default F estimator and the E / F post-run polish (separate TUs in the reference too)
g++ -O2 -ffp-contract=off -std=c++17 -w -fPIC -shared -Ishim_solvers -I"$REF" -I"$REF/relative_pose" -I. \
    -o _ref/libgcr_ref_solvers.so ref_solvers_glue.cpp gcr_oracle_*.cpp "$REF/graph.cpp" "$REF/maxflow.cpp" \
    "$REF/relative_pose/bundle.cpp" "$REF/relative_pose/essential.cpp" "$REF/relative_pose/colmap_models.cpp"
echo "built oracle/_ref/libgcr_ref_solvers.so"
