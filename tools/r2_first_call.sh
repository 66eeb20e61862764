#!/bin/bash
# First GPU call of the next round (run under gpurun, ~6 GPU-minutes):
#   gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# 1. parity of the experimental static-slab phase A (TSXC_PART_STATIC=1);
# 2. A/B of the two phase-A variants on every workload (device-resident, no e2e / CPU legs);
# 3. launch list + one full ncu capture of the variant that won, on the 1/16-scale config 2.
set -u
mkdir -p gpurun_out
{
  echo "== experimental parity"; TSXC_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k static_variant 2>&1 | tail -5
  for wl in c2 c2-fakeseq c3 c4 c5; do
    for v in 0 1; do
      echo "== $wl TSXC_PART_STATIC=$v"
      TSXC_PART_STATIC=$v timeout 200 python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline \
        > gpurun_out/r2_${wl}_static$v.json 2> gpurun_out/r2_${wl}_static$v.log
      python - "$wl" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_{sys.argv[1]}_static{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()})
except Exception as e:
    print("failed:", e)
PY
    done
  done
} 2>&1 | tee gpurun_out/r2_first_call.txt
