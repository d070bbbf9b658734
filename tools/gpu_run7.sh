#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 24 --warmup 8 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ours.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['stages_ms'], d['gpu_launches'])"; tail -3 gpurun_out/bench_ours.err
