#!/bin/bash
set -u
mkdir -p gpurun_out
{
  export TSXC_CHUNK_KEYS=250000000
  CMD="python bench.py --workload c2 --scale 0.0625 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-variants"
  timeout 300 $CMD > gpurun_out/c7_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_insert_keys" -c 2 -o gpurun_out/r02_insert_pipelined -f $CMD > gpurun_out/c7_ncu.log 2>&1
  grep "timed steps" gpurun_out/c7_plain.log
  tail -3 gpurun_out/c7_ncu.log
} 2>&1 | tee gpurun_out/r2_call7.txt
