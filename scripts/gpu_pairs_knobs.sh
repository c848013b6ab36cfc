#!/bin/bash
mkdir -p gpurun_out
run() { # workload pairs seats engines
  timeout 300 python bench.py --steps 8 --warmup 4 --no-cpu --workload $1 --pairs-api batch --pairs $2 --seats $3 --engines $4 > gpurun_out/knob.log 2>&1
  python - <<PY
import json
v=None
for line in open("gpurun_out/knob.log"):
    if line.startswith("{"): d=json.loads(line); v=d["value"]; st=d["config"]["step_ms"]
print("$1 pairs $2 seats $3 engines $4 -> %s pairs/s  steps %s" % ("%.0f" % v if v else "FAILED", st if v else ""))
PY
}
for w in h2k_pairs cfg1_pairs f4k_pairs; do
  run $w 128 128 2
  run $w 512 512 2
  run $w 512 512 4
done
run e2k_pairs 64 64 2
