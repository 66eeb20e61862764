#!/bin/bash
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
  echo "== K0r by region size, CAS and RED (MiB regions)"
  K0_REGIONS_KIB=16384,32768,65536,131072,262144,524288 K0_MODES=2,5 K0_BPS=6 timeout 300 python tools/k0modes.py | tee gpurun_out/k0modes_regions.jsonl
  for r in 25 26 27 28 29; do
    TSXC_REGION_LOG2=$r timeout 300 python bench.py --workload c2 $B > gpurun_out/c5_reg$r.json 2> gpurun_out/c5_reg$r.log
    echo -n "c2 region 2^$r: "; show gpurun_out/c5_reg$r.json
  done
} 2>&1 | tee gpurun_out/s2_call5.txt
