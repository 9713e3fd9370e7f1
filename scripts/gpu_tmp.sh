#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench_n1.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/r02t_bench_n1.json').read().strip().splitlines()[-1])
print('MM', d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'single', d['single_step']['us_per_step'])
for c in d['configs']: print(c['workload'] if 'workload' in c else c.get('env'), c.get('value'), c.get('roofline',{}).get('frac'), (c.get('long_launches') or {}).get('frac'))
PY
