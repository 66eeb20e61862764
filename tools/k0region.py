#!/usr/bin/env python3
"""K0r: L2-blocked region sweep — RMW rate vs (region size, touches per sector)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsxcount_b200 as tsx  # noqa: E402

lib = tsx._lib.load()
hm = tsx.TSXHashMapCUDA(34, 0, 31)
tb = hm.stats()["table_bytes"]
for region_mb in (16, 32, 64, 128):
    rb = region_mb << 20
    sectors = rb // 32
    for density in (0.25, 0.65, 1.0, 2.0):
        for mode, name in ((0, "red"), (2, "load+atomic")):
            for item in (1024, 8192):
                ops_per_region = int(sectors * density)
                n_regions = tb // rb
                ms = C.c_float(0)
                for _ in range(2):
                    tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, tb, rb, ops_per_region, item, mode, C.byref(ms)), hm.handle)
                total = ops_per_region * n_regions
                print(json.dumps({"region_mib": region_mb, "touches_per_sector": density, "mode": name, "ops_per_item": item,
                                  "g_ops_per_s": round(total / ms.value / 1e6, 2), "ms": round(ms.value, 1)}), flush=True)
hm.close()
