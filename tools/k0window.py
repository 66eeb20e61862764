#!/usr/bin/env python3
"""K0w sweep: per-block window size vs random-access rate at a fixed (large) aggregate footprint."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsxcount_b200 as tsx  # noqa: E402

lib = tsx._lib.load()
l = int(sys.argv[1]) if len(sys.argv) > 1 else 34
ops = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 31
hm = tsx.TSXHashMapCUDA(l, 0, 31)
tb = hm.stats()["table_bytes"]
for blocks, threads in ((148, 1024), (296, 1024), (592, 512), (1184, 256)):
    for wb in (1 << 21, 1 << 24, 1 << 26, 1 << 27, 1 << 28, 1 << 29, 1 << 30, 1 << 32, tb):
        row = {"blocks": blocks, "threads": threads, "window_mib": wb / 2**20}
        for mode, name in ((0, "red_add"), (3, "load32"), (2, "load+atomic")):
            ms = C.c_float(0)
            for _ in range(2):
                tsx._lib.check(lib.tsxc_k0_windowed(hm.handle, tb, wb, ops, mode, blocks, threads, C.byref(ms)), hm.handle)
            row[name] = round(ops / ms.value / 1e6, 2)
        print(json.dumps(row), flush=True)
hm.close()
