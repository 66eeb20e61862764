#!/bin/bash
# session 2, call 3: lean insert kernel (no stats spills), occupancy variants, K0r by mode/region/occupancy, L2 fetch granularity
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
  ( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -6
  echo "== K0r modes"
  timeout 300 python tools/k0modes.py | tee gpurun_out/k0modes.jsonl
  echo "== K0r modes, TSXC_L2_FETCH=32"
  TSXC_L2_FETCH=32 timeout 300 python tools/k0modes.py | grep -E '"region_kib": (8192|1024)' | tee gpurun_out/k0modes_fetch32.jsonl
  for v in "" minb5 minb4 minb8 pf; do
    if [ -n "$v" ]; then export TSXC_LIB=$PWD/tsxcount_b200/lib/libtsxcuda_$v.so; else unset TSXC_LIB; fi
    timeout 300 python bench.py --workload c2 $B > gpurun_out/c3_c2_$v.json 2> gpurun_out/c3_c2_$v.log
    echo -n "c2 variant '$v': "; show gpurun_out/c3_c2_$v.json
  done
  unset TSXC_LIB
  TSXC_L2_FETCH=32 timeout 300 python bench.py --workload c2 $B > gpurun_out/c3_fetch32.json 2> gpurun_out/c3_fetch32.log
  echo -n "c2 L2 fetch 32: "; show gpurun_out/c3_fetch32.json
  TSXC_REGION_LOG2=24 timeout 300 python bench.py --workload c2 $B > gpurun_out/c3_reg24.json 2> gpurun_out/c3_reg24.log
  echo -n "c2 region 2^24: "; show gpurun_out/c3_reg24.json
  TSXC_REGION_LOG2=21 timeout 300 python bench.py --workload c2 $B > gpurun_out/c3_reg21.json 2> gpurun_out/c3_reg21.log
  echo -n "c2 region 2^21: "; show gpurun_out/c3_reg21.json
  for w in c2-fakeseq c5; do
    timeout 300 python bench.py --workload $w $B > gpurun_out/c3_$w.json 2> gpurun_out/c3_$w.log
    echo -n "$w: "; show gpurun_out/c3_$w.json
  done
} 2>&1 | tee gpurun_out/s2_call3.txt
