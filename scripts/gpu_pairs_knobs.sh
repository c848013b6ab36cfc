#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py -q -x -p no:cacheprovider 2>&1 | tail -1
run() { # workload pairs seats engines
  timeout 200 python bench.py --steps 12 --warmup 4 --no-cpu --workload $1 --pairs-api batch --pairs $2 --seats $3 --engines $4 > gpurun_out/knob.log 2>&1
  python - <<PY
import json
v=None
for line in open("gpurun_out/knob.log"):
    if line.startswith("{"): d=json.loads(line); v=d["value"]; st=d["config"]["step_ms"]
print("$1 pairs $2 seats $3 engines $4 -> %s pairs/s  steps %s" % ("%.0f" % v if v else "FAILED", st if v else ""))
PY
}
run cfg1_pairs 512 512 2
run cfg1_pairs 512 512 2
