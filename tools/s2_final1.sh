#!/bin/bash
# final single-GPU verification: the driver's three commands + the launch list of the final build
set -u
mkdir -p gpurun_out
{
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  ( time timeout 900 python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.log
  echo "rc=$?"; grep -v "cpu baseline" gpurun_out/final_bench.log | tail -16 | cut -c1-250
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2), "roofline", round(d["roofline"]["frac"], 4), "rand8", round(d["roofline_rand8"]["frac"], 3), round(d["roofline_rand8"]["peak"], 1), "insert alone", round(d["roofline_rand8"]["insert_kernel_alone"]["frac"], 3), "launches", d["gpu_launches"], d["clocks"])
print("cpu", json.dumps(d["cpu_baseline"])[:900])
print("variants", {k: (round(v["value"], 2), round(v["ms_per_step"], 1)) for k, v in d["variants"].items()})
PY
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/final_ref.json 2> gpurun_out/final_ref.log
  cut -c1-300 gpurun_out/final_ref.json
  CMD="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > gpurun_out/final_ncu_list.log 2>&1
  echo "ncu rc=$?"
} 2>&1 | tee gpurun_out/s2_final1.txt
