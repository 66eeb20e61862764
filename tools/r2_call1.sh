#!/bin/bash
# Round 2, GPU call 1: does the L2 fetch-granularity limit change phase B / K0r?  (one-line experiment)
set -u
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv
  for f in 0 32; do
    echo "== c2 bench TSXC_L2_FETCH=$f"
    if [ $f = 0 ]; then unset TSXC_L2_FETCH; else export TSXC_L2_FETCH=$f; fi
    timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/c1_c2_fetch$f.json 2> gpurun_out/c1_c2_fetch$f.log
    python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/c1_c2_fetch{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["value"], 2), "Gk-mer/s", d["ms_per_step"], {k: round(v, 1) for k, v in d["roofline"]["phase_ms"].items()}, d["roofline_rand8"]["k0"])
except Exception as e:
    print("failed:", e)
PY
    echo "== K0r TSXC_L2_FETCH=$f"
    timeout 300 python - <<'PY'
import ctypes as C, json, os, sys
sys.path.insert(0, os.getcwd())
import tsxcount_b200 as tsx
lib = tsx._lib.load()
hm = tsx.TSXHashMapCUDA(34, 0, 31)
tb = hm.stats()["table_bytes"]
for region_mb in (8, 16, 64):
    rb = region_mb << 20
    sectors = rb // 32
    for density in (0.65, 1.0):
        for mode, name in ((0, "red"), (2, "load+atomic")):
            ops_per_region = int(sectors * density)
            n_regions = tb // rb
            ms = C.c_float(0)
            for _ in range(2):
                tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, tb, rb, ops_per_region, 1024, mode, C.byref(ms)), hm.handle)
            total = ops_per_region * n_regions
            print(json.dumps({"region_mib": region_mb, "touches": density, "mode": name, "g_ops_per_s": round(total / ms.value / 1e6, 2)}), flush=True)
hm.close()
PY
  done
} 2>&1 | tee gpurun_out/r2_call1.txt
