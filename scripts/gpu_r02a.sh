#!/bin/bash
# round-2 GPU call A: parity tier, the driver-shaped bench line, launch list + full-set capture of the K=20 MM rollout
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02a_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/r02a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mg_env_kernel.*MMPolicy -s 40 -c 2 -f -o gpurun_out/r02a_mm \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/r02a_ncu_full.log 2>&1
cat gpurun_out/r02a_pytest.log; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
