#!/bin/bash
# round 2, session 2, call 1: GPU tests, the full bench line, the reference arm, ncu launch list + full captures
set -u
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
  nproc; grep -m1 "model name" /proc/cpuinfo
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12
  echo "== bench (default args)"
  ( time timeout 900 python bench.py ) > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.log
  echo "rc=$?"; tail -30 gpurun_out/s2_bench.log
  echo "== reference arm"
  ( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/s2_ref.json 2> gpurun_out/s2_ref.log
  cut -c1-500 gpurun_out/s2_ref.json; tail -4 gpurun_out/s2_ref.log
  echo "== ncu launch list (full-size c2, 1 warm-up + 1 timed step)"
  CMD="python bench.py --workload c2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > gpurun_out/s2_ncu_list.log 2>&1
  echo "rc=$?"; tail -2 gpurun_out/s2_ncu_list.log
  echo "== ncu full capture (c2 scaled 1/16, chunk of 250 M keys = same touches per region pass as full size)"
  export TSXC_CHUNK_KEYS=250000000
  CMD="python bench.py --workload c2 --scale 0.0625 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/s2_scaled_plain.log 2>&1; echo "rc=$?"; grep "timed steps" gpurun_out/s2_scaled_plain.log
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_insert_keys|k_part_reads|k_part_keys|k_hist_keys|k_hist_reads" --launch-skip 51 -c 8 -o gpurun_out/r02_pipeline_c2_scaled -f $CMD > gpurun_out/s2_ncu_full.log 2>&1
  echo "rc=$?"; tail -3 gpurun_out/s2_ncu_full.log
} 2>&1 | tee gpurun_out/s2_call1.txt
