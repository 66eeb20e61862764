#!/bin/bash
set -u
mkdir -p gpurun_out
{
  ( time timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q ) 2>&1 | tail -15
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants 2>&1 | grep "timed steps"
} 2>&1 | tee gpurun_out/s2_call9.txt
