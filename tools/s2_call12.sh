#!/bin/bash
set -u
mkdir -p gpurun_out
{
  ( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12
  timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants 2>&1 | grep "timed steps"
  echo "== f2 check timing"
  python - <<'PY'
import os, subprocess, sys, time
sys.path.insert(0, "tools")
from cli_ingest_bench import CLI, make_fastq
small = "/tmp/f2_small.fastq"
make_fastq(small, 800_000, seed=7)
base = [CLI, "--input=" + small, "--k=31", "--l=29", "--s=4", "--mode=CUDA", "--readers=4", "--threads=16"]
def run(extra):
    t0 = time.time(); p = subprocess.run(base + extra, capture_output=True, text=True); return time.time() - t0, p
dt0, p = run([])
dt1, p = run(["--dump=" + small + ".31.count"])
size = os.path.getsize(small + ".31.count")
dt2, p = run(["--check"])
print(f"count {dt0:.2f}s; +dump {size/1e9:.2f} GB in {dt1-dt0:.2f}s; +check in {dt2-dt0:.2f}s = {size/max(dt2-dt0,1e-9)/1e9:.2f} GB/s of text, {96e6/max(dt2-dt0,1e-9)/1e6:.1f} M k-mers/s", [l for l in p.stdout.splitlines() if l.startswith(("total errors", "queried (Xor)"))])
PY
} 2>&1 | tee gpurun_out/s2_call12.txt
