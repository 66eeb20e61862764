#!/bin/bash
# Second GPU call of session 3: sparse walk with consecutive positions per thread and two blocks per SM at k > 64.
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("failed:", e)
PY
}
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
{
  for spec in c4:62 c3:62 c2:90; do
    w=${spec%%:*}; p=${spec##*:}
    TSXC_SPARSE_PCT=$p timeout 100 python bench.py --workload $w $B > gpurun_out/s3b_${w}_$p.json 2> gpurun_out/s3b_${w}_$p.log
    echo -n "$w pct=$p: "; show gpurun_out/s3b_${w}_$p.json
  done
  ( time timeout 300 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
} 2>&1 | tee gpurun_out/s3_sparse2.txt
