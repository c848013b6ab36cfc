#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
for w in f4k_pairs f4k_pairs h2k_pairs e2k_pairs; do
GCR_BENCH_SLOW_PAIRS=1 timeout 300 python bench.py --steps 12 --warmup 2 --no-cpu --workload $w > gpurun_out/slow_$w.log 2> gpurun_out/slow_$w.err
grep "slow pair" gpurun_out/slow_$w.err | head -3
python - $w <<'PY'
import json, sys
for line in open("gpurun_out/slow_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d=json.loads(line); print(sys.argv[1], "%.0f pairs/s steps ms %s" % (d["value"], d["config"]["step_ms"]))
PY
done
