#!/bin/bash
mkdir -p gpurun_out
for mode in LAZY EAGER LAZY EAGER; do
CUDA_MODULE_LOADING=$mode timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu --workload f4k_pairs > gpurun_out/fvar.log 2>&1
python - $mode <<'PY'
import json, sys
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "f4k_pairs %.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
