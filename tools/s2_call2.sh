#!/bin/bash
# session 2, call 2: parity of the new ranking + prefetch, A/B of the variants, full ncu capture of the pipeline
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline"]["chunk_cap_keys"], d["gpu_launches"])
    for k, v in (d.get("variants") or {}).items():
        print("  ", k, round(v["value"], 2), round(v["ms_per_step"], 1), {a: round(b, 1) for a, b in v["phase_ms"].items()})
except Exception as e:
    print("failed:", e)
PY
}
B="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
{
  ( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -6
  for v in "" nopf nocd; do
    if [ -n "$v" ]; then export TSXC_LIB=$PWD/tsxcount_b200/lib/libtsxcuda_$v.so; else unset TSXC_LIB; fi
    timeout 300 python bench.py --workload c2 $B > gpurun_out/c2_c2_$v.json 2> gpurun_out/c2_c2_$v.log
    echo -n "c2 variant '$v': "; show gpurun_out/c2_c2_$v.json
  done
  unset TSXC_LIB
  for g in 4 8; do
    TSXC_INSERT_GRID=$g timeout 300 python bench.py --workload c2 $B > gpurun_out/c2_grid$g.json 2> gpurun_out/c2_grid$g.log
    echo -n "c2 insert grid $g: "; show gpurun_out/c2_grid$g.json
  done
  for r in 22 24; do
    TSXC_REGION_LOG2=$r timeout 300 python bench.py --workload c2 $B > gpurun_out/c2_reg$r.json 2> gpurun_out/c2_reg$r.log
    echo -n "c2 region 2^$r: "; show gpurun_out/c2_reg$r.json
  done
  for w in c2-fakeseq c3 c4 c5; do
    timeout 300 python bench.py --workload $w $B > gpurun_out/c2_$w.json 2> gpurun_out/c2_$w.log
    echo -n "$w: "; show gpurun_out/c2_$w.json
  done
  echo "== ncu full capture (c2 scaled 1/16, chunk of 250 M keys)"
  export TSXC_CHUNK_KEYS=250000000
  CMD="python bench.py --workload c2 --scale 0.0625 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/c2_scaled_plain.log 2>&1; echo "rc=$?"; grep "timed steps" gpurun_out/c2_scaled_plain.log
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_insert_keys|k_part_reads|k_part_keys|k_hist_keys|k_hist_reads" --launch-skip 176 -c 14 -o gpurun_out/r02_pipeline_c2_scaled -f $CMD > gpurun_out/c2_ncu_full.log 2>&1
  echo "rc=$?"; tail -3 gpurun_out/c2_ncu_full.log
} 2>&1 | tee gpurun_out/s2_call2.txt
