#!/bin/bash
# 2-GPU validation: NCCL parity tests, C++ CLI --gpus, then a short multi-GPU bench
set -u
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=index,name --format=csv,noheader
  ( time timeout 900 python -m pytest tests/test_multigpu_nccl.py tests/test_cli.py -x -q -m gpu -k "nccl or multi_gpu" ) 2>&1 | tail -15
  echo "== bench --gpus 2"
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/mg2.json 2> gpurun_out/mg2.log
  echo "rc=$?"; grep -E "bench|Error|error|Traceback" gpurun_out/mg2.log | tail -25
  python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/mg2.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", round(d["ms_per_step"], 1), "dev", round(d["device_ms_per_step"], 1), d["roofline"]["phase_ms_rank0"], "e2e", d["e2e"] and round(d["e2e"]["value"], 2), d["parity"] and d["parity"]["ok"], d["nvlink"])
    for k, v in d["variants"].items(): print("  ", k, round(v["value"], 2), round(v["ms_per_step"], 1), v["phase_ms_rank0"])
except Exception as e:
    print("failed:", e)
PY
} 2>&1 | tee gpurun_out/s2_mg2.txt
