#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/refit_bench.py
python scripts/prof_pairs.py 2>&1 | tail -4 | cut -c1-200
for w in f4k_pairs e2k_pairs; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > gpurun_out/fvar.log 2>&1
python - $w <<'PY'
import json, sys
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "%.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refit_e -s 3 -c 1 -o gpurun_out/r2_prof_k6_e -f python scripts/prof_pipeline_kernels.py e > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log
