#!/usr/bin/env python3
"""Developer micro-benchmark (not the driver contract; see bench.py): device-resident synthetic reads,
fused count kernel timed with CUDA events on the handle's stream, plus the K0 random-RMW roofline.

  python tools/devbench.py --k 31 --l 30 --reads 8000000 --mode 0 --reps 3
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tsxcount_b200 as tsx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--l", type=int, default=30)
    ap.add_argument("--s", type=int, default=0)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--genome", type=int, default=0)
    ap.add_argument("--sub", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--k0", type=int, default=1, help="run the K0 random-RMW microbenchmark")
    ap.add_argument("--k0-ops", type=int, default=1 << 30)
    ap.add_argument("--query", type=int, default=0, help="time the batched lookup kernel (K4)")
    args = ap.parse_args()

    lib = tsx._lib.load()
    import torch
    torch.cuda.init()
    dev = 0
    n_bases = args.reads * args.read_len
    n_words = (n_bases + 31) // 32
    d_packed = torch.empty(n_words + 8, dtype=torch.int64, device="cuda")
    d_off = torch.empty(args.reads + 1, dtype=torch.int64, device="cuda")
    p = tsx.TsxcGenParams(0xC2, args.reads, args.read_len, args.mode, args.genome, args.sub, 0)
    tsx._lib.check(lib.tsxc_gen_reads_device(C.byref(p), 0, args.reads, dev, None, d_packed.data_ptr(), d_off.data_ptr()))
    torch.cuda.synchronize()
    n_kmers = args.reads * max(0, args.read_len - args.k + 1)

    hm = tsx.TSXHashMapCUDA(args.l, args.s, args.k, device=dev, flags=args.flags)
    st0 = hm.stats()
    print(json.dumps({"layout": {k: st0[k] for k in ("entry_words", "value_bits", "quotient_bits", "slots_per_bucket",
                                                      "n_slots", "table_bytes")}}))
    stream = torch.cuda.ExternalStream(lib.tsxc_stream(hm.handle))
    res = []
    for rep in range(args.reps):
        hm.clear()
        hm.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        hm.addReadsDevice(d_packed.data_ptr(), d_off.data_ptr(), args.reads, n_bases)
        e1.record(stream)
        hm.sync()
        ms = e0.elapsed_time(e1)
        st = hm.stats()
        assert st["kmers_added"] == n_kmers, (st["kmers_added"], n_kmers)
        res.append(ms)
        print(json.dumps({"rep": rep, "ms": round(ms, 3), "gkmers_per_s": round(n_kmers / ms / 1e6, 3),
                          "distinct": st["distinct"], "load": round(st["used_slots"] / st["n_slots"], 4),
                          "overflow_entries": st["overflow_entries"], "max_reprobe": st["max_reprobe"],
                          "partition_ms": round(st["partition_ms"], 2), "insert_ms": round(st["insert_ms"], 2),
                          "launches": st["kernel_launches"]}))
    if args.query:
        # K4: batched getKmerCount on device-resident k-mers (random keys: almost all absent, one probe each)
        n_q = min(1 << 27, st["distinct"])
        d_cnt = torch.empty(n_q, dtype=torch.int64, device="cuda")
        rnd = torch.randint(0, 2**62, (n_q * hm.kw,), dtype=torch.int64, device="cuda")
        if hm.kw == 1 and args.k < 32:
            rnd &= (1 << (2 * args.k)) - 1
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tsx._lib.check(lib.tsxc_lookup_device(hm.handle, rnd.data_ptr(), n_q, d_cnt.data_ptr()), hm.handle)
        e1.record(stream)
        hm.sync()
        print(json.dumps({"k4_lookup_random_keys": n_q, "ms": round(e0.elapsed_time(e1), 2),
                          "g_lookups_per_s": round(n_q / e0.elapsed_time(e1) / 1e6, 2)}))
    if args.k0:
        for mode, name in ((0, "red_add"), (1, "cas"), (2, "sector_load+atomic")):
            hm.clear(); hm.sync()
            ms = C.c_float(0)
            tsx._lib.check(lib.tsxc_k0_random_rmw(hm.handle, st0["table_bytes"], args.k0_ops, mode, C.byref(ms)), hm.handle)
            ms2 = C.c_float(0)
            tsx._lib.check(lib.tsxc_k0_random_rmw(hm.handle, st0["table_bytes"], args.k0_ops, mode, C.byref(ms2)), hm.handle)
            print(json.dumps({"k0": name, "table_bytes": st0["table_bytes"], "ops": args.k0_ops,
                              "g_rmw_per_s_first": round(args.k0_ops / ms.value / 1e6, 3),
                              "g_rmw_per_s_second": round(args.k0_ops / ms2.value / 1e6, 3)}))
    hm.close()


if __name__ == "__main__":
    main()
