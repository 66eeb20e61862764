#!/bin/bash
# Round 2, GPU call 2: parity of the region-sorted pipeline, then a first timing of config 2.
set -u
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -25
  echo "== c2"
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c2_c2.json 2> gpurun_out/c2_c2.log
  tail -5 gpurun_out/c2_c2.log
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c2_c2.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", d["ms_per_step"], {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline"]["chunk_cap_keys"], d["gpu_launches"])
except Exception as e:
    print("failed:", e)
PY
} 2>&1 | tee gpurun_out/r2_call2.txt
