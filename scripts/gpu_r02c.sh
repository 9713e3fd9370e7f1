#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02c_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
cat gpurun_out/r02c_pytest.log; tail -c 1500 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
