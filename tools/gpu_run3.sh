#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/diag.json'))
for k,v in d.items():
    if k=='timing':
        for kk,vv in v.items(): print(kk, json.dumps(vv))
    elif isinstance(v,dict) and 'error' in v: print(k, v['error'], v['tb'][-600:])
    elif isinstance(v,dict):
        print(k, {kk:vv for kk,vv in v.items() if not kk.startswith('grad_')}, {kk:(round(vv['rel'],9),round(vv['noise'],9)) for kk,vv in v.items() if kk.startswith('grad_')})
PY
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
