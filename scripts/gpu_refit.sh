#!/bin/bash
python tools/refit_bench.py | grep "^F"
for v in ba_it0 ba_it0_svd6 svd6; do echo "== $v"; GCR_LIB_PATH=$PWD/graph-cut-ransac_b200/variants/$v.so python tools/refit_bench.py | grep "^F ba_frame 1"; done
