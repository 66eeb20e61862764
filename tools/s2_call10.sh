#!/bin/bash
# session 2, call 10: the full default bench line, the reference arm, ncu launch list (full size) and full captures (1/4 scale)
set -u
mkdir -p gpurun_out
{
  nproc; grep -m1 "model name" /proc/cpuinfo
  echo "== bench (default args)"
  ( time timeout 900 python bench.py ) > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.log
  echo "rc=$?"; grep -v "cpu baseline" gpurun_out/s10_bench.log | tail -25
  echo "== reference arm"
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/s10_ref.json 2> gpurun_out/s10_ref.log
  cut -c1-400 gpurun_out/s10_ref.json; tail -4 gpurun_out/s10_ref.log
  echo "== ncu launch list (full-size c2, 1 warm-up + 1 timed step)"
  CMD="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > gpurun_out/s10_ncu_list.log 2>&1
  echo "rc=$?"
  echo "== ncu full capture (c2 scaled 1/4: 32 GiB table, 512 bins of 64 MiB, two chunks of 1e9 keys = 0.93 touches per sector and pass)"
  export TSXC_REGION_LOG2=26 TSXC_CHUNK_KEYS=1000000000
  CMD="python bench.py --workload c2 --scale 0.25 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/s10_scaled_plain.log 2>&1; echo "rc=$?"; grep "timed steps" gpurun_out/s10_scaled_plain.log
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_part_reads|k_insert_keys|k_count_segs|k_build_slices" -c 4 -o gpurun_out/r02_single_pass_c2_quarter -f $CMD > gpurun_out/s10_ncu_full.log 2>&1
  echo "rc=$?"; tail -3 gpurun_out/s10_ncu_full.log
} 2>&1 | tee gpurun_out/s2_call10.txt
