#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name\|thread\|core(s)\|socket" | head -5
for rep in 1 2 3; do
  timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --workload f4k_pairs > gpurun_out/fvar.log 2>&1
  python - <<'PY'
import json
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print("f4k_pairs %.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
for w in 8 12 24; do
  timeout 300 python bench.py --steps 6 --warmup 2 --no-cpu --workload f4k_pairs --workers $w > gpurun_out/fvar.log 2>&1
  python - $w <<'PY'
import json, sys
for line in open("gpurun_out/fvar.log"):
    if line.startswith("{"):
        d=json.loads(line); print("workers", sys.argv[1], "f4k_pairs %.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
