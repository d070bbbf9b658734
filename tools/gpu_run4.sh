#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/make_golden.py > gpurun_out/golden.log 2>&1; tail -5 gpurun_out/golden.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 16 --warmup 4 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
