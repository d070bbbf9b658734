#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time of every kernel of 3 steps, (2) full capture of our kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python tools/prof_step.py 4 > gpurun_out/r01_launches.log 2>&1
tail -2 gpurun_out/r01_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_|emit_|tile_ranges" -s 10 -c 6 -o gpurun_out/r01_kernels -f python tools/prof_step.py 4 > gpurun_out/r01_ncu.log 2>&1
tail -3 gpurun_out/r01_ncu.log
timeout 300 python bench.py --steps 24 --warmup 8 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json | cut -c1-900
timeout 300 python bench.py --impl reference --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
ls -la gpurun_out | tail -8
