#!/bin/bash
# round-2 GPU call B: new GPU tests (codec, HostPipe, zero-copy, device guard, full episodes), bench line, host memory probes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
nproc > gpurun_out/r02b_host.txt; lscpu | head -25 >> gpurun_out/r02b_host.txt; free -g >> gpurun_out/r02b_host.txt
nvidia-smi topo -m >> gpurun_out/r02b_host.txt 2>&1
cat gpurun_out/r02b_pytest.log; tail -c 2500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
