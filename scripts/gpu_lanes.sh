#!/bin/bash
mkdir -p gpurun_out
for w in f4k h4k e2k; do
  for cap in 5 3; do
    timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --workload $w --lanes-k3-cap $cap > gpurun_out/lanes_${w}_$cap.log 2>&1
    python - $w $cap <<'PY'
import json, sys
l = json.loads(open("gpurun_out/lanes_%s_%s.log" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
t = l["config"].get("multi_stream_resident")
print(sys.argv[1], "lanes cap", sys.argv[2], "resident ms/round %.4f  K3 frac %.4f  e2e ms/round %.4f  multi-stream resident %s" % (
    l["config"]["ms_per_round"], l["roofline"]["frac"], l["e2e"]["ms_per_round"], ("%.4f" % t["ms_per_round"]) if t else None))
PY
  done
done
