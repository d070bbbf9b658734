#!/bin/bash
# A/B of kernel variants through bench.py's stage timings: tools/ab_bwd.sh "opt=val" "opt=val" ...
mkdir -p gpurun_out
for o in "$@"; do
  f=gpurun_out/ab_$(echo $o | tr '=,' '__').json
  timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-robustness --no-sharded --option $o > $f 2>/dev/null
  python - "$o" "$f" <<'PY'
import json, sys
l = [x for x in open(sys.argv[2]) if x.startswith('{"metric')][-1]
d = json.loads(l); s = d.get("stages_ms", {})
print(sys.argv[1], round(d["ms_per_step"], 4), {k: s.get(k) for k in ("preprocess_fwd", "render_fwd", "render_bwd", "preprocess_bwd")})
PY
done
