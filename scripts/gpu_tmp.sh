#!/bin/bash
N=${1:-8}
cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
e=d['e2e']; print('N', d['n_gpus'], 'value', round(d['value'],1), 'e2e direct', e['value'], 'd2h', e['d2h_bytes_per_step'], 'ok', e['host_obs_equals_device'], 'patch', e['patch_pipe']['value'], 'raw', e['raw_copy']['value'])"
