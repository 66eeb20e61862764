#!/bin/bash
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline"]["chunk_cap_keys"], d["gpu_launches"])
except Exception as e:
    print("failed:", e)
PY
}
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
{
  ( time timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -8
  for w in c2 c3 c4; do
    timeout 300 python bench.py --workload $w $B > gpurun_out/c11_$w.json 2> gpurun_out/c11_$w.log
    echo -n "$w: "; show gpurun_out/c11_$w.json
  done
  df -h /tmp | tail -1; free -g | head -2
  timeout 900 python tools/cli_f1f2_bench.py 2>&1 | tee gpurun_out/f1f2.txt
} 2>&1 | tee gpurun_out/s2_call11.txt
