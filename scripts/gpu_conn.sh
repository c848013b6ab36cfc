#!/bin/bash
mkdir -p gpurun_out
run() { # conn workload extra...
  c=$1; w=$2; shift; shift
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --workload $w "$@" > gpurun_out/conn.log 2>&1
  python - $c $w "$*" <<'PY'
import json, sys
v=None
for line in open("gpurun_out/conn.log"):
    if line.startswith("{"): d=json.loads(line); v=d["value"]
print("connections", sys.argv[1], sys.argv[2], sys.argv[3], "-> %s" % ("%.4g" % v if v else "FAILED"))
PY
}
for c in 8 32; do
  for w in f4k_pairs h2k_pairs cfg1_pairs e2k_pairs; do run $c $w; done
  run $c f4k_pairs --workers 32
  run $c f4k
  run $c h2k_batch
done
