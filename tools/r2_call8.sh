#!/bin/bash
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
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline or sharded or route or synthetic" 2>&1 | tail -3
  for v in "" minb5 minb4; do
    echo "== c2 variant '$v'"
    if [ -n "$v" ]; then export TSXC_LIB=$PWD/tsxcount_b200/lib/libtsxcuda_$v.so; else unset TSXC_LIB; fi
    for g in 6 8; do
      TSXC_INSERT_GRID=$g timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/c8_c2_$v$g.json 2> gpurun_out/c8_c2_$v$g.log
      echo -n "grid $g: "; show gpurun_out/c8_c2_$v$g.json
    done
  done
  unset TSXC_LIB
  echo "== variants"
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c8_c2.json 2> gpurun_out/c8_c2.log
  show gpurun_out/c8_c2.json
} 2>&1 | tee gpurun_out/r2_call8.txt
