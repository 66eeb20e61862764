#!/bin/bash
# Round 2, GPU call 4: timing after 256-bit bucket loads / PTX ballot match / S2a ILP
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
{
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or sharded or route or synthetic" 2>&1 | tail -5
  echo "== c2 + variants"
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c4_c2.json 2> gpurun_out/c4_c2.log
  tail -3 gpurun_out/c4_c2.log; show gpurun_out/c4_c2.json
  for g in 4 8; do
    echo "== c2 insert grid $g/SM"
    TSXC_INSERT_GRID=$g timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/c4_c2_g$g.json 2> gpurun_out/c4_c2_g$g.log
    show gpurun_out/c4_c2_g$g.json
  done
} 2>&1 | tee gpurun_out/r2_call4.txt
