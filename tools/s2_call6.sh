#!/bin/bash
# session 2, call 6: single-pass partition (1024 bins) with paged bins: parity, then the workloads
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline"]["chunk_cap_keys"], d["gpu_launches"], "K0r", d["roofline_rand8"] and (round(d["roofline_rand8"]["peak"],1), round(d["roofline_rand8"]["frac"],3)))
except Exception as e:
    print("failed:", e)
PY
}
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
{
  ( time timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -15
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
  for w in c2 c2-fakeseq c5 c3 c4; do
    timeout 300 python bench.py --workload $w $B > gpurun_out/c6_$w.json 2> gpurun_out/c6_$w.log
    echo -n "$w: "; show gpurun_out/c6_$w.json; grep -i "error\|Traceback" gpurun_out/c6_$w.log | head -3
  done
  TSXC_NO_PAGING=1 timeout 300 python bench.py --workload c2 $B > gpurun_out/c6_c2_exact.json 2> gpurun_out/c6_c2_exact.log
  echo -n "c2 exact offsets (S0): "; show gpurun_out/c6_c2_exact.json
  TSXC_REGION_LOG2=26 timeout 300 python bench.py --workload c2 $B > gpurun_out/c6_c2_r26.json 2> gpurun_out/c6_c2_r26.log
  echo -n "c2 region 2^26 (still 1024 bins: no effect expected): "; show gpurun_out/c6_c2_r26.json
  TSXC_REGION_LOG2=28 timeout 300 python bench.py --workload c2 $B > gpurun_out/c6_c2_r28.json 2> gpurun_out/c6_c2_r28.log
  echo -n "c2 region 2^28 (512 bins): "; show gpurun_out/c6_c2_r28.json
} 2>&1 | tee gpurun_out/s2_call6.txt
