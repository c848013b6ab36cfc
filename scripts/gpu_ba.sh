#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python scripts/prof_pairs.py 2>&1 | tail -4 | cut -c1-200
for w in f4k_pairs e2k_pairs h2k_pairs cfg1_pairs; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > gpurun_out/fvar.log 2>&1
python - $w <<'PY'
import json, sys
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "%.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
