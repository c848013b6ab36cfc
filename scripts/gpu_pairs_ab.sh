#!/bin/bash
# pipelines: batch API vs thread-pool API, a few seat / engine settings
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
for w in cfg1_pairs h2k_pairs f4k_pairs e2k_pairs; do
  timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --workload $w --pairs-api threads > gpurun_out/pairs_${w}_threads.log 2>&1
  for cfg in "96 2" "128 4" "64 1" "128 2"; do
    set -- $cfg
    GCR_BATCH_STATS=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --workload $w --pairs-api batch --seats $1 --engines $2 > gpurun_out/pairs_${w}_batch_$1_$2.log 2>&1
  done
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/pairs_*.log")):
    v = None
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line); v = d["value"]
    print("%-50s %s" % (f.split("/")[-1], "%.0f pairs/s" % v if v else "FAILED"))
PY
