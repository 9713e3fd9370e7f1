#!/bin/bash
# round-2 final tree, N GPUs: the driver-shaped bench (all configs)
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02w_bench_n$N.json 2> gpurun_out/r02w_bench_n$N.err
tail -3 gpurun_out/r02w_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02w_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'MM', round(d['value'],1), round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'per_rank_ms', d.get('per_rank_ms'), 'shard_check', d.get('shard_check'))
for c in d['configs']:
    print(c['workload'].split(',')[0], c['workload'].split(',')[1], round(c.get('value'),1), round(c['roofline']['frac'],3))
PY
