#!/bin/bash
# GPU session: diagnostics vs reference, pytest -m gpu, launch list + full ncu capture of the render kernels.
mkdir -p gpurun_out
timeout 500 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
grep -v '"tb"' gpurun_out/diag.log | cut -c1-1800 | tail -40
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py 4 > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_(fwd|bwd)_warp" -s 4 -c 2 -o gpurun_out/prof_render -f python tools/prof_step.py 3 > gpurun_out/ncu_render.log 2>&1
tail -5 gpurun_out/ncu_render.log
ls -la gpurun_out
