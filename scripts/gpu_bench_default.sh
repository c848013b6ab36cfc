#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err
tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print("f4k value %.4g ms/round %.4f frac %.4f e2e %.4g (%.4f ms/round) multi %s" % (d['value'], d['config']['ms_per_round'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_round'], d['config'].get('multi_stream_resident')))
for k,v in d['extra'].items():
    r=v.get('roofline') or {}
    print(k, '%.4g'%v['value'], v['unit'], 'e2e %.4g' % v.get('e2e'), 'ms/round', v.get('ms_per_round'), 'frac', r.get('frac'), 'pairs/s', v.get('pairs_per_s'), v.get('pairs_per_s_e2e'), v.get('pairs_per_s_multi_stream'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), 'wall', v.get('bench_wall_s'))
PY
