#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== K0r: which returning atomic costs what (8 MiB regions)"
  K0_REGIONS_KIB=8192 K0_MODES=2,5,6,7,8,9,10,4,0 K0_BPS=8,6 timeout 300 python tools/k0modes.py | tee gpurun_out/k0modes_atomics.jsonl
} 2>&1 | tee gpurun_out/s2_call4.txt
