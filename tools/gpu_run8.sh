#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
BWD=4,8 FWD=3 timeout 300 python tools/gpu_sweep.py 2>&1 | tail -1
timeout 600 python tools/edit_loop_bench.py 200 2>&1 | tail -1
