#!/bin/bash
set -u
mkdir -p gpurun_out
{
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
  ( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.log
  echo "rc=$?"
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/final2_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), "rand8", round(d["roofline_rand8"]["frac"], 3), "launches", d["gpu_launches"], d["clocks"])
print("variants", {k: (round(v["value"], 2), round(v["ms_per_step"], 1)) for k, v in d["variants"].items()})
PY
} 2>&1 | tee gpurun_out/s2_final2.txt
