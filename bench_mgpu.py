"""bench.py --gpus N (N > 1): the multi-rank arm of the benchmark and its untimed oracle-parity preamble.

Bench harness, not product: this file (like bench.py) may use tests/oracle_py.py as the checker; the product package
tsxcount_b200/ never does.  One rank per GPU (torchrun), weak scaling: every rank brings the workload's reads.
"""
import ctypes as C
import math

import torch
import torch.distributed as dist

from tsxcount_b200 import _lib
from tsxcount_b200.multigpu import CudaRouteBackend, ShardedCounter


# ---------------------------------------------------------------------------------------------------------
# oracle parity through the real multi-rank path (untimed; the oracle is the checker, never the thing measured)
# ---------------------------------------------------------------------------------------------------------
def parity_check(wl, rank, world, local_rank, n_reads=20_000, l_small=22, region_log2="17"):
    """Counts n_reads reads per rank of the workload's generator into a small sharded table through the same
    ShardedCounter / NCCL / peer-store path as the benchmark, gathers every shard's dump on rank 0 and compares the
    union with the oracle's count of all ranks' reads as a map.  Returns a dict on rank 0, None elsewhere."""
    import os
    import sys
    import numpy as np
    root = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(root, "tests"))
    lib = _lib.load()
    k, read_len = wl["k"], wl["read_len"]
    shard_bits = int(math.log2(world))
    dev = torch.device("cuda", local_rank)
    n_bases = n_reads * read_len
    n_words = (n_bases + 31) // 32
    d_packed = torch.zeros(n_words + 8, dtype=torch.int64, device=dev)
    d_off = torch.zeros(n_reads + 1, dtype=torch.int64, device=dev)
    gp = _lib.TsxcGenParams(wl["seed"], max(wl["reads"], n_reads) * world, read_len, wl["mode"], wl["genome"], wl["sub"], 0)
    _lib.check(lib.tsxc_gen_reads_device(C.byref(gp), rank * n_reads, n_reads, local_rank, None, d_packed.data_ptr(), d_off.data_ptr()))
    torch.cuda.synchronize()
    old = {v: os.environ.get(v) for v in ("TSXC_REGION_LOG2", "TSXC_SEG_LOG2", "TSXC_CHUNK_KEYS", "TSXC_PAGE_LOG2")}
    os.environ["TSXC_REGION_LOG2"] = region_log2      # small regions: the small shards take the routed pipeline
    os.environ["TSXC_SEG_LOG2"] = "11"                # several segments, several rounds
    os.environ["TSXC_PAGE_LOG2"] = "5"                # 32-key pages: the small receive buffers still get a page pool (fine pass)
    try:
        be = CudaRouteBackend(k, l_small + shard_bits, 0, rank, world, local_rank)
        sc = ShardedCounter(be, rank, world, recv_cap_keys=n_reads * read_len // 2 + (1 << 17))
        sc.add_reads_device(d_packed, d_off, n_reads, n_bases)
        be.sync()
        keys, counts = be.hm.getAllKmers()
        st = be.hm.stats()
        assert world == 1 or st["group_cap_keys"] == 32, "the parity preamble must take the two-level path (page pool) of the benchmark"
        rounds = sc.rounds
        be.close()
    finally:
        for v, x in old.items():
            if x is None:
                os.environ.pop(v, None)
            else:
                os.environ[v] = x
    mine = (keys.tobytes(), counts.tobytes(), int(st["kmers_added"]), int(st["error_flags"]))
    gathered = [None] * world if rank == 0 else None
    if world > 1:
        dist.gather_object(mine, gathered, dst=0)
    else:
        gathered = [mine]
    if rank != 0:
        return None
    import oracle_py as orc
    seqs = []
    for r in range(world):
        seqs += orc.gen_reads(seed=wl["seed"], n_reads=max(wl["reads"], n_reads) * world, read_len=read_len, mode=wl["mode"],
                              genome_len=wl["genome"], sub_rate_q16=wl["sub"], first=r * n_reads, count=n_reads)
    oc = orc.count_seqs(seqs, k)
    kw = be.hm.kw
    got, added, errs = {}, 0, 0
    for kb, cb, a, e in gathered:
        ks = np.frombuffer(kb, dtype=np.uint64).reshape(-1, kw)
        cs = np.frombuffer(cb, dtype=np.uint64)
        for key, c in zip(ks.tolist(), cs.tolist()):
            assert tuple(key) not in got, "a k-mer is stored on two shards"
            got[tuple(key)] = int(c)
        added += a
        errs |= e
    ok = (got == oc.as_dict(kw)) and added == oc.n_total and errs == 0
    return {"checked": True, "ok": bool(ok), "kmers": int(oc.n_total), "distinct": int(oc.n_distinct), "ranks": world,
            "rounds": rounds, "path": "ShardedCounter: route_hist -> all_gather -> route_send (peer stores) -> all_reduce -> "
                                      "route_insert; shard dumps gathered on rank 0 and compared with the oracle as a map"}


# ---------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): one rank per GPU, weak scaling (every rank brings wl["reads"] reads)
# ---------------------------------------------------------------------------------------------------------
def bench_main(args, wl, rank, world, local_rank, log=lambda m: None, extra_workloads=()):
    import json
    import statistics
    import time

    # stdout carries exactly one JSON line (rank 0): NCCL's version / debug lines go to stderr
    import os
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = _lib.load()
    shard_bits = int(math.log2(world))
    assert 1 << shard_bits == world, "the table is sharded by hash bits: N must be a power of two"

    parity = None
    if not args.no_parity:
        parity = parity_check(wl, rank, world, local_rank)
        if rank == 0:
            log(f"parity preamble: {parity}")
            assert parity["ok"], parity

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    be, sc = None, None

    def run_workload(w, want_e2e):
        nonlocal be, sc
        k, l_global = w["k"], w["l"] + shard_bits
        n_reads, read_len = w["reads"], w["read_len"]
        n_bases = n_reads * read_len
        n_words = (n_bases + 31) // 32
        n_kmers = n_reads * max(0, read_len - k + 1)
        d_packed = torch.empty(n_words + 8, dtype=torch.int64, device=dev)
        d_off = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
        gp = _lib.TsxcGenParams(w["seed"], n_reads * world, read_len, w["mode"], w["genome"], w["sub"], 0)
        _lib.check(lib.tsxc_gen_reads_device(C.byref(gp), rank * n_reads, n_reads, local_rank, None, d_packed.data_ptr(), d_off.data_ptr()))
        torch.cuda.synchronize()
        if be is None or (be.hm.k, be.hm.l) != (k, l_global):
            if be is not None:
                be.close()
            be = CudaRouteBackend(k, l_global, 0, rank, world, local_rank, flags=w.get("flags", 0))
            sc = ShardedCounter(be, rank, world)
        layout = be.hm.stats()

        def one_step():
            be.hm.clear()
            be.hm.sync()
            fence()
            t0 = time.perf_counter()
            be.hm.mark(0)
            sc.add_reads_device(d_packed, d_off, n_reads, n_bases)
            be.hm.mark(1)
            fence()
            dt = time.perf_counter() - t0
            dev_ms = be.hm.elapsed_ms(0, 1)
            t = torch.tensor([dt, dev_ms * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0].item()), float(t[1].item())

        sampler = None
        if rank == 0:
            import bench as _bench
            sampler = _bench.ClockSampler(local_rank)
            sampler.start()          # before the warm-up: nvidia-smi needs a few hundred ms for its first sample
        for i in range(args.warmup):
            dt, dms = one_step()
            if rank == 0:
                log(f"{w['name']} warmup {i}: wall {dt * 1e3:.1f} ms, device {dms * 1e3:.1f} ms")
        if sampler:
            sampler.rows.clear()     # keep only the samples of the timed steps
        steps = [one_step() for _ in range(args.steps)]
        clocks = sampler.stop() if sampler else None
        st = be.hm.stats()
        added = torch.tensor([st["kmers_added"], st["distinct"], st["kernel_launches"], st["error_flags"]], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(added, op=dist.ReduceOp.SUM)
        total_added, total_distinct, launches, errs = [int(x) for x in added.tolist()]
        assert total_added == n_kmers * world and errs == 0, (total_added, n_kmers * world, errs)
        T = sum(s[0] for s in steps)                 # wall clock between fences, max over ranks
        T_dev = sum(s[1] for s in steps)             # CUDA events on the handle's stream, max over ranks
        res = {"name": w["name"], "value": args.steps * n_kmers * world / T / 1e9, "ms_per_step": 1e3 * T / args.steps,
               "device_ms_per_step": 1e3 * T_dev / args.steps, "kmers_per_step": n_kmers * world, "distinct": total_distinct,
               "launches": launches, "clocks": clocks, "layout": layout, "rounds_per_step": sc.rounds // max(1, sc.batches),
               # of the LAST timed step (tsxc_clear resets the accounting): CUDA events around each phase's launches
               "phase_ms_rank0": {"hist": st["hist_ms"], "route": st["part1_ms"], "insert": st["insert_ms"]},
               "n_kmers_rank": n_kmers, "read_len": read_len, "k": k, "l_global": l_global,
               "recv_cap_keys": sc.recv_cap}
        # e2e: the rank's reads start in pinned host memory; H2D copy + routed counting + global distinct read-back
        if want_e2e:
            h_packed = torch.empty(n_words + 8, dtype=torch.int64, pin_memory=True)
            h_off = torch.empty(n_reads + 1, dtype=torch.int64, pin_memory=True)
            h_packed.copy_(d_packed)
            h_off.copy_(d_off)
            times = []
            for it in range(1 + args.steps):
                be.hm.clear()
                be.hm.sync()
                fence()
                t0 = time.perf_counter()
                with torch.cuda.stream(be.stream):
                    d_packed.copy_(h_packed, non_blocking=True)
                    d_off.copy_(h_off, non_blocking=True)
                sc.add_reads_device(d_packed, d_off, n_reads, n_bases)
                got = sc.distinct_global()
                fence()
                dt = time.perf_counter() - t0
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                assert got == total_distinct
                if it:
                    times.append(float(t.item()))
            res["e2e"] = {"value": n_kmers * world / statistics.mean(times) / 1e9, "unit": "Gk-mer/s",
                          "h2d_bytes_per_step": world * ((n_words + 8) * 8 + (n_reads + 1) * 8), "d2h_bytes_per_step": world * 8,
                          "timing": "wall clock, barrier + synchronize on both sides, max over ranks; H2D copies queued on the "
                                    "counting stream"}
            del h_packed, h_off
        del d_packed, d_off
        torch.cuda.empty_cache()
        return res

    main = run_workload(wl, not args.no_e2e)
    extras = {}
    for w in extra_workloads:
        r = run_workload(w, False)
        extras[w["name"]] = {"value": r["value"], "unit": "Gk-mer/s", "ms_per_step": r["ms_per_step"],
                             "device_ms_per_step": r["device_ms_per_step"], "kmers_per_step": r["kmers_per_step"],
                             "distinct": r["distinct"], "phase_ms_rank0": r["phase_ms_rank0"], "workload": w["desc"]}

    if rank == 0:
        E = 8 * main["layout"]["entry_words"]
        k, read_len, n_kmers = main["k"], main["read_len"], main["n_kmers_rank"]
        in_b = 0.25 * read_len / max(1, read_len - k + 1)
        import bench as _bench
        peak, peak_src = _bench.read_peaks()
        route_ms = main["phase_ms_rank0"]["route"]
        achieved = n_kmers * (2 * E + in_b) / (main["ms_per_step"] * 1e-3) / 1e9   # per GPU: this rank's k-mers over the step time
        sent = n_kmers * E * (world - 1) // world                         # bytes a rank stores into its peers per step
        line = {
            "metric": "k-mers counted/sec", "value": main["value"], "unit": "Gk-mer/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] + f" x{world} ranks, table hash-sharded over {world} GPUs (config 5 routing: the "
                                                "routing kernel stores into the owners' peer-mapped buffers over NVLink, then insert)",
                       "k": k, "l_global": main["l_global"], "reads_per_gpu": wl["reads"], "kmers_per_step": main["kmers_per_step"],
                       "distinct": main["distinct"], "entry_bytes": E, "table_bytes_per_gpu": main["layout"]["table_bytes"],
                       "exchange": "peer stores inside the routing kernel (CUDA IPC mapped receive buffers); NCCL only for the "
                                   "per-round histogram all-gather and the barrier",
                       "rounds_per_step": main["rounds_per_step"], "recv_cap_keys": main["recv_cap_keys"],
                       "l2": "inputs and table shards far exceed the 126 MB L2; shards re-zeroed between steps",
                       "timing": "wall clock per step between barrier+synchronize fences, max over ranks; zeroing untimed"},
            "device_ms_per_step": main["device_ms_per_step"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "k_hist_reads + k_part_reads (routing, peer stores) + k_insert_keys (per GPU)",
                         "algorithmic_bytes_per_kmer": 2 * E + in_b, "phase_ms_rank0": main["phase_ms_rank0"]},
            "nvlink": {"sent_bytes_per_gpu_per_step": sent,
                       "GB_s_per_gpu_per_direction_during_routing": (sent / (route_ms * 1e-3) / 1e9) if route_ms else None,
                       "routing_share_of_step": route_ms / main["device_ms_per_step"] if route_ms else None,
                       "note": "payload the routing kernel stores into peer memory, over the routing kernel's own device time"},
            "parity": parity, "variants": extras,
            "cpu_baseline": None, "e2e": main.get("e2e"), "gpu_launches": main["launches"], "clocks": main["clocks"],
        }
        print(json.dumps(line), flush=True)
    fence()
    if be is not None:
        be.close()
    dist.destroy_process_group()
    return 0
