#!/bin/bash
# Round 2, GPU call 5: what bounds phase B?  K0r with a consumed CAS / returning atomic / load only; insert without warp match
set -u
mkdir -p gpurun_out
{
  echo "== K0r modes (8 MiB regions, 0.93 touches/sector, 2048 ops per item)"
  timeout 300 python - <<'PY'
import ctypes as C, json, os, sys
sys.path.insert(0, os.getcwd())
import tsxcount_b200 as tsx
lib = tsx._lib.load()
hm = tsx.TSXHashMapCUDA(34, 0, 31)
tb = hm.stats()["table_bytes"]
rb = 8 << 20
for touches in (0.93, 0.5):
    for mode, name in ((0, "red"), (2, "load+red"), (3, "load+cas(consumed)"), (4, "atomic with return"), (5, "load only")):
        for item in (2048, 512):
            opr = int(rb // 32 * touches)
            ms = C.c_float(0)
            for _ in range(2):
                tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, tb, rb, opr, item, mode, C.byref(ms)), hm.handle)
            print(json.dumps({"touches": touches, "mode": name, "item": item, "g_ops_per_s": round(opr * (tb // rb) / ms.value / 1e6, 2)}), flush=True)
hm.close()
PY
  echo "== c2, no warp aggregation (TSXC_FLAG_NO_WARP_AGG)"
  TSXC_BENCH_FLAGS=2 timeout 300 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-variants 2>&1 >/dev/null | grep "timed steps"
} 2>&1 | tee gpurun_out/r2_call5.txt
