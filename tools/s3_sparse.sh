#!/bin/bash
# GPU call of session 3: full -m gpu suite with the sparse walk in place, then S1 dense vs sparse on c4 / c3 / c2 / c5.
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
  ( time timeout 600 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
  for spec in c4:50 c4:0 c3:0 c3:70 c2:0 c2:90 c5:90; do
    w=${spec%%:*}; p=${spec##*:}
    TSXC_SPARSE_PCT=$p timeout 200 python bench.py --workload $w $B > gpurun_out/s3_${w}_$p.json 2> gpurun_out/s3_${w}_$p.log
    echo -n "$w pct=$p: "; show gpurun_out/s3_${w}_$p.json
  done
} 2>&1 | tee gpurun_out/s3_sparse.txt
