#!/bin/bash
# Round-end evidence on one B200 (round 2): full GPU test-suite, ncu launch list + full capture of our kernels, both
# bench arms, the config-5 loop A/B. Everything lands in gpurun_out/ (copy the summaries into profiles/ afterwards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu.log; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 400 python bench.py --steps 40 --warmup 8 > gpurun_out/r02_bench_ours.json 2> gpurun_out/r02_bench_ours.err; cut -c1-250 gpurun_out/r02_bench_ours.json
timeout 400 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-250 gpurun_out/r02_bench_reference.json
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-robustness --option binning_variant=0 > gpurun_out/r02_bench_ours_binning0.json 2> /dev/null; cut -c1-200 gpurun_out/r02_bench_ours_binning0.json
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-robustness --no-sharded --option render_bwd_variant=4 > gpurun_out/r02_bench_ours_bwd_variant4.json 2> /dev/null; cut -c1-200 gpurun_out/r02_bench_ours_bwd_variant4.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_launches.csv python tools/prof_step.py 4 > gpurun_out/r02_launches.log 2>&1; tail -1 gpurun_out/r02_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_|tile_sort|tile_prefix|tile_count" -s 9 -c 9 -o gpurun_out/r02_kernels -f python tools/prof_step.py 4 > gpurun_out/r02_ncu.log 2>&1; tail -2 gpurun_out/r02_ncu.log
timeout 600 python tools/edit_loop_bench.py > gpurun_out/r02_edit_loop_c5.json 2> gpurun_out/r02_edit_loop.err; cat gpurun_out/r02_edit_loop_c5.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
