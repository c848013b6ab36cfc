#!/bin/bash
# A/B of library variants on one box: scripts/ab.sh workload variant.so [variant.so ...]  ("base" = the in-tree library)
w=$1; shift
for v in "$@"; do
  for rep in 1 2; do
    if [ "$v" = base ]; then unset GCR_LIB_PATH; else export GCR_LIB_PATH=$PWD/graph-cut-ransac_b200/variants/$v; fi
    timeout 300 python bench.py --steps 8 --warmup 2 --no-cpu --workload $w > gpurun_out/ab.log 2>&1
    python - "$w" "$v" <<'PY'
import json, sys
l = json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2], "ms/round %.4f  K3 ms %.4f  frac %.4f  e2e ms/round %.4f" % (l["config"]["ms_per_round"], l["roofline"]["k3_ms_per_launch"], l["roofline"]["frac"], l["e2e"]["ms_per_round"]))
PY
  done
done
