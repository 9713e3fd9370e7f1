#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pipe_profile.py 4096 > gpurun_out/r02p_pipe.log 2>&1
cat gpurun_out/r02p_pipe.log
