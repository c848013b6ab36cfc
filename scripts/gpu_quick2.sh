#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for w in f4k h4k e2k h2k_batch; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --workload $w > gpurun_out/q_$w.log 2>&1
  python - $w <<'PY'
import json, sys
l = json.loads(open("gpurun_out/q_%s.log" % sys.argv[1]).read().strip().splitlines()[-1])
t = l["config"].get("multi_stream_resident")
print(sys.argv[1], "resident ms/round %.4f  K3 frac %.4f  e2e ms/round %.4f  multi-stream resident %s  pairs/s %s %s" % (
    l["config"]["ms_per_round"], l["roofline"]["frac"], l["e2e"]["ms_per_round"], ("%.4f" % t["ms_per_round"]) if t else None, l["config"].get("pairs_per_s"), l["config"].get("pairs_per_s_e2e")))
PY
done
