#!/bin/bash
set -u
mkdir -p gpurun_out
{
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  echo "== c2 full bench line (e2e, cpu baseline, variants)"
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.log
  echo "rc=$?"; tail -12 gpurun_out/c9_bench.log
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c9_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), "e2e", d["e2e"] and (round(d["e2e"]["value"], 2), round(d["e2e"]["ms_per_step"], 1)))
    print("cpu", json.dumps(d["cpu_baseline"])[:1500])
    print("rand8", d["roofline_rand8"]["frac"], d["roofline_rand8"]["insert_kernel_alone"])
except Exception as e:
    print("failed:", e)
PY
  echo "== reference arm"
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/c9_ref.json 2> gpurun_out/c9_ref.log
  cat gpurun_out/c9_ref.json | cut -c1-600; tail -4 gpurun_out/c9_ref.log
} 2>&1 | tee gpurun_out/r2_call9.txt
