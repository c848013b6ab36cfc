#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
python scripts/prof_pairs.py 2>&1 | tail -4 | cut -c1-200
for rep in 1 2 3; do
timeout 300 python bench.py --steps 10 --warmup 2 --no-cpu --workload f4k_pairs > gpurun_out/fvar.log 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print("f4k_pairs %.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
