#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
GCR_BENCH_SLOW_PAIRS=1 timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu --workload f4k_pairs > gpurun_out/slow.log 2> gpurun_out/slow.err
grep "slow pair" gpurun_out/slow.err | head -6
python - <<'PY'
import json
for line in open("gpurun_out/slow.log"):
    if line.startswith("{"):
        d=json.loads(line); print("f4k_pairs %.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
