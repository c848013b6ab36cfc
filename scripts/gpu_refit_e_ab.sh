#!/bin/bash
# lane-parallel k_refit_e against a reference build of the same ABI in graph-cut-ransac_b200/variants/v0.so (e.g. commit f25a137, one
# thread per set: GCR_LIB_OUT=.../variants/v0.so python build.py --force in a checkout of it): the whole GPU suite, k_refit_e's
# outputs bit for bit (tools/refit_e_dump.py), the E pipeline
mkdir -p gpurun_out
L=$PWD/graph-cut-ransac_b200
LOG=gpurun_out/final.log
: > $LOG
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log >> $LOG
cat $LOG
GCR_LIB_PATH=$L/variants/v0.so python tools/refit_e_dump.py gpurun_out/refit_e_v0.npz > /dev/null 2>> $LOG
python tools/refit_e_dump.py gpurun_out/refit_e_cur.npz > /dev/null 2>> $LOG
python - >> $LOG 2>&1 <<'PY'
import numpy as np
a, b = np.load("gpurun_out/refit_e_cur.npz"), np.load("gpurun_out/refit_e_v0.npz")
print("k_refit_e lanes vs one thread:", {k: ("bit-identical" if np.array_equal(a[k], b[k]) else "DIFFERENT %d" % int((a[k] != b[k]).sum())) for k in b.files})
PY
tail -n 1 $LOG
timeout 60 python bench.py --workload e2k_pairs --steps 6 --warmup 2 --no-cpu > gpurun_out/e2k_pairs_lanes.json 2> gpurun_out/e2k_pairs_lanes.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/e2k_pairs_lanes.json') if l.startswith('{')][-1])
print('e2k_pairs', d['value'], d['config']['step_ms'])" 2>&1 | tee -a $LOG
