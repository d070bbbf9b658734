#!/bin/bash
# Round-end evidence on one B200: full GPU test-suite, ncu launch list + full capture of our kernels, both bench arms.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python tools/prof_step.py 4 > gpurun_out/r01_launches.log 2>&1
tail -1 gpurun_out/r01_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_|emit_|tile_ranges" -s 10 -c 6 -o gpurun_out/r01_kernels -f python tools/prof_step.py 4 > gpurun_out/r01_ncu.log 2>&1
tail -1 gpurun_out/r01_ncu.log
timeout 300 python bench.py --steps 40 --warmup 8 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-300 gpurun_out/bench_ours.json
timeout 300 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
