#!/bin/bash
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
  for w in c2 c2-fakeseq c5 c3; do
    timeout 300 python bench.py --workload $w $B > gpurun_out/c7_$w.json 2> gpurun_out/c7_$w.log
    echo -n "$w: "; show gpurun_out/c7_$w.json; grep -i "error\|Traceback" gpurun_out/c7_$w.log | head -3
  done
  TSXC_NO_PAGING=1 timeout 300 python bench.py --workload c2 $B > gpurun_out/c7_c2_exact.json 2> gpurun_out/c7_c2_exact.log
  echo -n "c2 exact offsets (S0): "; show gpurun_out/c7_c2_exact.json
  echo "== ncu: S1 with 1024 bins (c2 scaled 1/16, 8 MiB regions of the 8 GiB table)"
  export TSXC_REGION_LOG2=23
  CMD="python bench.py --workload c2 --scale 0.0625 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/c7_scaled_plain.log 2>&1; echo "rc=$?"; grep "timed steps" gpurun_out/c7_scaled_plain.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_part_reads|k_insert_keys|k_count_segs" --launch-skip 3 -c 3 -o gpurun_out/r02_single_pass_c2_scaled -f $CMD > gpurun_out/c7_ncu_full.log 2>&1
  echo "rc=$?"; tail -3 gpurun_out/c7_ncu_full.log
} 2>&1 | tee gpurun_out/s2_call7.txt
