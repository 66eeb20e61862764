#!/bin/bash
# Round 2, GPU call 3: parity, config 2 timing (ballot vs MATCH.ANY ranking), ncu --set full of the five pipeline kernels
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline"]["chunk_cap_keys"], d["gpu_launches"], (d.get("roofline_rand8") or {}).get("k0r"))
except Exception as e:
    print("failed:", e)
PY
}
{
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or sharded or route or large_default" 2>&1 | tail -5
  echo "== c2 (ballot ranking)"
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/c3_c2.json 2> gpurun_out/c3_c2.log
  tail -3 gpurun_out/c3_c2.log; show gpurun_out/c3_c2.json
  echo "== c2 (MATCH.ANY ranking)"
  TSXC_LIB=$PWD/tsxcount_b200/lib/libtsxcuda_match.so timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/c3_c2m.json 2> gpurun_out/c3_c2m.log
  tail -3 gpurun_out/c3_c2m.log; show gpurun_out/c3_c2m.json
  echo "== ncu"
  export TSXC_CHUNK_KEYS=250000000
  CMD="python bench.py --workload c2 --scale 0.0625 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/c3_ncu_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hist_reads|k_part_reads|k_hist_keys|k_part_keys|k_insert_keys" -c 5 -o gpurun_out/r02_pipeline_c2_scaled -f $CMD > gpurun_out/c3_ncu.log 2>&1
  tail -5 gpurun_out/c3_ncu.log
  ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/r2_call3.txt
