#!/bin/bash
# bench.py on N GPUs of one box exactly as the driver launches it: tools/gpu_multi.sh N  (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 40 --warmup 8 --no-robustness > gpurun_out/r02_bench_${N}gpu.raw 2> gpurun_out/r02_bench_${N}gpu.err
grep '^{"metric' gpurun_out/r02_bench_${N}gpu.raw | tail -1 > gpurun_out/r02_bench_${N}gpu.json
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu.json"))
s = d.get("sharded_c4") or {}
print(d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 4), "| sharded:", s.get("ms_per_step"), s.get("vs_1gpu_plain"), s.get("image_equals_1gpu"), s.get("phase_ms_max_over_ranks"))
PY
