#!/bin/bash
# N-GPU bench line (N = number of visible GPUs)
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
{
  echo "GPUs: $N"
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 2 --warmup 1 ) > gpurun_out/mg$N.json 2> gpurun_out/mg$N.log
  echo "rc=$?"; grep -E "bench|Error|error|Traceback" gpurun_out/mg$N.log | tail -12 | cut -c1-300
  python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/mg{n}.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), "dev", round(d["device_ms_per_step"], 1), d["roofline"]["phase_ms_rank0"], "e2e", d["e2e"] and round(d["e2e"]["value"], 2), d["parity"] and d["parity"]["ok"], d["nvlink"]["GB_s_per_gpu_per_direction_during_routing"])
    for k, v in d["variants"].items(): print("  ", k, round(v["value"], 2), round(v["ms_per_step"], 1), v["phase_ms_rank0"])
except Exception as e:
    print("failed:", e)
PY
} 2>&1 | tee gpurun_out/s2_mg$N.txt
