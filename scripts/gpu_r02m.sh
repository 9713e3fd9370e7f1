#!/bin/bash
# round-2 multi-GPU call: raw D2H ceiling with all ranks copying at once, then the driver-shaped bench at N = 8 and N = 2
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/d2h_ceiling.py > gpurun_out/r02m_d2h_n$N.json 2> gpurun_out/r02m_d2h.err
tail -1 gpurun_out/r02m_d2h_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m_bench_n$N.json 2> gpurun_out/r02m_bench_n$N.err
tail -c 300 gpurun_out/r02m_bench_n$N.json; tail -3 gpurun_out/r02m_bench_n$N.err
if [ "$N" != "2" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err; fi
nproc > gpurun_out/r02m_host.txt; nvidia-smi topo -m >> gpurun_out/r02m_host.txt 2>&1
