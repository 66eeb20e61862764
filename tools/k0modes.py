#!/usr/bin/env python3
"""K0r by mode, region size and occupancy at the benchmark's touches per sector (0.93): which part of the insert's
dependent chain (sector load -> CAS -> wait) costs what.  Output: JSON lines."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsxcount_b200 as tsx  # noqa: E402

lib = tsx._lib.load()
hm = tsx.TSXHashMapCUDA(34, 0, 31)
tb = hm.stats()["table_bytes"]
NAMES = {1: "load", 2: "load+red", 3: "load+cas(late use)", 5: "load+cas+wait", 4: "atomic(returning)", 0: "red",
         6: "load+cas32+wait", 7: "load+add64(returning)+wait", 8: "load+add32(returning)+wait", 9: "load+exch64+wait",
         10: "load+or64(returning)+wait"}
density = 0.93
REGIONS = [int(x) for x in os.environ.get('K0_REGIONS_KIB', '128,1024,8192,16384').split(',')]
MODES = [int(x) for x in os.environ.get('K0_MODES', '1,2,3,5').split(',')]
BPS = [int(x) for x in os.environ.get('K0_BPS', '8,6,5').split(',')]
for region_kb in REGIONS:
    rb = region_kb << 10
    sectors = rb // 32
    for mode in MODES:
        for bps in BPS:
            if mode in (1,) and bps != 8:
                continue
            ops_per_region = int(sectors * density)
            n_regions = tb // rb
            ms = C.c_float(0)
            for _ in range(2):
                tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, tb, rb, ops_per_region, 1024, mode | (bps << 8), C.byref(ms)), hm.handle)
            total = ops_per_region * n_regions
            print(json.dumps({"region_kib": region_kb, "touches_per_sector": density, "mode": NAMES[mode], "blocks_per_sm": bps,
                              "g_ops_per_s": round(total / ms.value / 1e6, 2), "ms": round(ms.value, 1)}), flush=True)
hm.close()
