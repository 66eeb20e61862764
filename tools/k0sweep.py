#!/usr/bin/env python3
"""K0 sweep: random-access rate vs footprint and access kind (developer tool)."""
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
names = {0: "red_add", 1: "cas", 2: "load+atomic", 3: "load32", 4: "store8"}
fp = 1 << 27
while fp <= tb:
    row = {"footprint_gib": fp / 2**30}
    for mode in (0, 3, 4, 2):
        ms = C.c_float(0)
        tsx._lib.check(lib.tsxc_k0_random_rmw(hm.handle, fp, ops, mode, C.byref(ms)), hm.handle)
        tsx._lib.check(lib.tsxc_k0_random_rmw(hm.handle, fp, ops, mode, C.byref(ms)), hm.handle)
        row[names[mode]] = round(ops / ms.value / 1e6, 2)
    print(json.dumps(row), flush=True)
    fp *= 2
hm.close()
