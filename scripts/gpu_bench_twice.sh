#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
timeout 1200 python bench.py > gpurun_out/bench_default_$rep.log 2> gpurun_out/bench_default_$rep.err
python - $rep <<'PY'
import json, sys
d=json.loads(open('gpurun_out/bench_default_%s.log' % sys.argv[1]).read().strip().splitlines()[-1])
print("f4k value %.4g e2e %.4g" % (d['value'], d['e2e']['value']), {k: round(v['value']) for k, v in d['extra'].items() if k.endswith('_pairs')})
PY
done
timeout 300 python bench.py --workload f4k_pairs --no-cpu --steps 3 --warmup 1 | python -c "import sys,json; print('f4k_pairs alone', json.loads(sys.stdin.read().strip().splitlines()[-1])['value'])"
