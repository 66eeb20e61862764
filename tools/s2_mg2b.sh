#!/bin/bash
set -u
mkdir -p gpurun_out
{
  for c in 5 3 2; do
    export TSXC_ROUTE_COARSE_BITS=$c
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$c bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e --no-variants > gpurun_out/mg2_c$c.json 2> gpurun_out/mg2_c$c.log
    echo -n "coarse bits $c: rc=$? "
    python - $c <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/mg2_c{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), d["roofline"]["phase_ms_rank0"], d["parity"] and d["parity"]["ok"], round(d["nvlink"]["GB_s_per_gpu_per_direction_during_routing"]))
except Exception as e:
    print("failed:", e)
PY
  done
} 2>&1 | tee gpurun_out/s2_mg2b.txt
