#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python scripts/prof_pairs.py cfg1 2>&1 | grep "cfg1 sampler 2"
python scripts/prof_pairs.py 2>&1 | tail -4 | cut -c1-200
for w in cfg1_pairs h2k_pairs f4k_pairs e2k_pairs; do
  for rep in 1 2; do
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --workload $w > gpurun_out/pairs_$w.log 2>&1
  python - $w <<'PY'
import json, sys
for line in open("gpurun_out/pairs_%s.log" % sys.argv[1]):
    if line.startswith("{"): print(sys.argv[1], "%.0f pairs/s" % json.loads(line)["value"])
PY
  done
done
