"""Hash-sharded counting over the GPUs of one box: one process per GPU, `torch.distributed` for the plumbing.

Data path per chunk of reads on every rank (the reference has no counterpart, it is one process
on one shared table, src/mains/main.cpp:132-218 of mjoppich/tsxCount):

  route     tsxc_route_chunk      extract + hash, bin each k-mer by (owning shard, table region of that shard)
  exchange  all_to_all_single     bins, bin fill counters and spill lists travel to the owning rank (NCCL over
                                  NVLink/NVSwitch); issued on a side stream so that it overlaps the routing of the
                                  next chunk
  insert    tsxc_insert_routed    the owner inserts the received bins region by region

The table is partitioned by the top log2(world) bits of the bucket index of the bijective k-mer hash, so shards
never exchange anything after insertion: distinct counts add up, dumps concatenate.

`Backend` is what the orchestration needs from a device; `CudaBackend` is the product, tests/ substitute a
NumPy stand-in to exercise the scheduling / exchange logic with the gloo backend on CPUs.
"""
import ctypes as C
import math

import torch
import torch.distributed as dist

from . import _lib
from .hashmap import TSXHashMapCUDA


class CudaBackend:
    """Device buffers are torch tensors; kernels run on the handle's own stream."""

    def __init__(self, k, l_global, s, rank, world, device, flags=0, max_chunk_words=0, kmers_per_position=1.0):
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.hm = TSXHashMapCUDA(l_global, s, k, device=device, flags=flags, shard_rank=rank, n_shards=world)
        self.lay = self.hm.routeLayout(max_chunk_words, kmers_per_position)
        self.kw = self.lay.key_words
        self.stream = torch.cuda.ExternalStream(self.hm._lib.tsxc_stream(self.hm.handle), device=self.device)
        self.comm_stream = torch.cuda.Stream(device=self.device)

    # buffers -----------------------------------------------------------------------------------------
    def alloc_u64(self, n):
        return torch.empty(max(int(n), 1), dtype=torch.int64, device=self.device)

    def to_host(self, t):
        return t.cpu()

    # kernels -----------------------------------------------------------------------------------------
    def prepare(self, d_offsets, n_reads, n_bases):
        self.hm.routePrepare(d_offsets.data_ptr(), n_reads, n_bases)

    def route(self, d_packed, n_bases, w0, w1, bins, cursors, spill, spill_n):
        self.hm.routeChunk(self.lay, d_packed.data_ptr(), n_bases, w0, w1, bins.data_ptr(), cursors.data_ptr(),
                           spill.data_ptr(), spill_n.data_ptr())

    def overflowed(self):
        return self.hm.routeOverflowed()

    def insert(self, bins, cursors, n_sources):
        self.hm.insertRouted(self.lay, bins.data_ptr(), cursors.data_ptr(), n_sources)

    def insert_spill(self, records, n):
        self.hm.addHashCountsDevice(records.data_ptr(), n)

    def sync(self):
        self.hm.sync()

    def distinct(self):
        return self.hm.getKmerCount()


class ShardedCounter:
    """Counts the k-mers of this rank's reads into the table sharded over all ranks of `group`."""

    def __init__(self, backend, rank, world, group=None, min_split_words=1024):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        self.min_split_words = min_split_words
        lay = backend.lay
        assert lay.n_shards == world
        self.lay = lay
        G, kw = world, lay.key_words
        # double-buffered send / receive sets: the exchange of chunk c overlaps the routing of chunk c+1
        self.send = [dict(bins=backend.alloc_u64(G * lay.block_words), cursors=backend.alloc_u64(G * lay.bins_per_shard),
                          spill=backend.alloc_u64(G * lay.spill_cap * (kw + 1)), spill_n=backend.alloc_u64(G))
                     for _ in range(2)]
        self.recv = [dict(bins=backend.alloc_u64(G * lay.block_words), cursors=backend.alloc_u64(G * lay.bins_per_shard),
                          spill_n=backend.alloc_u64(G)) for _ in range(2)]
        self.a2a_bytes = 0
        self.chunks = 0
        self.retries = 0

    # -- exchange -----------------------------------------------------------------------------------------
    def _a2a(self, out, inp):
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, group=self.group)
            self.a2a_bytes += inp.numel() * 8 * (self.world - 1) // self.world

    def _exchange_spill(self, b, send_n_host, recv_n_host):
        """Variable-size exchange of the (hash, count) records; sizes are known on the host."""
        rw = self.lay.key_words + 1
        s = self.send[b]
        cap = self.lay.spill_cap
        parts = [s["spill"][o * cap * rw: o * cap * rw + int(send_n_host[o]) * rw] for o in range(self.world)]
        inp = torch.cat(parts) if sum(int(x) for x in send_n_host) else s["spill"][:0]
        n_out = int(sum(recv_n_host)) * rw
        out = self.be.alloc_u64(n_out)[:n_out]
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_to_all_single(out, inp, output_split_sizes=[int(x) * rw for x in recv_n_host],
                                   input_split_sizes=[int(x) * rw for x in send_n_host], group=self.group)
        return out

    def _agree(self, value, op):
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.int64, device=self.send[0]["cursors"].device)
        dist.all_reduce(t, op=op, group=self.group)
        return int(t.item())

    # -- one batch of reads -------------------------------------------------------------------------------
    def add_reads_device(self, d_packed, d_offsets, n_reads, n_bases):
        be, lay = self.be, self.lay
        n_words = (n_bases + 31) // 32
        be.prepare(d_offsets, n_reads, n_bases)
        # every rank must take part in the same number of exchanges
        my_chunks = math.ceil(n_words / lay.chunk_words) if n_words else 0
        n_chunks = self._agree(my_chunks, dist.ReduceOp.MAX)
        ranges = [(min(n_words, c * lay.chunk_words), min(n_words, (c + 1) * lay.chunk_words)) for c in range(n_chunks)]
        i = 0
        use_cuda = isinstance(be, CudaBackend)
        pending = None  # (buffer set, comm-done event, spill records)
        while i < len(ranges) or pending is not None:
            nxt = None
            if i < len(ranges):
                w0, w1 = ranges[i]
                b = self.chunks & 1
                s = self.send[b]
                be.route(d_packed, n_bases, w0, w1, s["bins"], s["cursors"], s["spill"], s["spill_n"])
                over = 1 if be.overflowed() else 0              # waits for the routing kernel (and the previous insert)
                over = self._agree(over, dist.ReduceOp.MAX)     # the split must be collective
                if over and w1 - w0 > self.min_split_words:
                    mid = (w0 + w1) // 2
                    ranges[i:i + 1] = [(w0, mid), (mid, w1)]
                    self.retries += 1
                    continue
                if over:
                    raise RuntimeError(f"spill lists overflow even for a {self.min_split_words}-word chunk")
                nxt = b
                i += 1
                self.chunks += 1
            # drain the previous chunk: its exchange ran while we were routing
            if pending is not None:
                pb, ev, spill_rec, spill_total = pending
                if use_cuda:
                    be.stream.wait_event(ev)
                be.insert(self.recv[pb]["bins"], self.recv[pb]["cursors"], self.world)
                if spill_total:
                    be.insert_spill(spill_rec, spill_total)
                pending = None
            if nxt is not None:
                pending = self._start_exchange(nxt, use_cuda)
        be.sync()

    def _start_exchange(self, b, use_cuda):
        """Queue the exchange of buffer set b.  Only the tiny spill-count collectives are waited for on the host;
        the bins travel asynchronously while the caller routes the next chunk."""
        be = self.be
        s, r = self.send[b], self.recv[b]

        def body():
            self._a2a(r["spill_n"], s["spill_n"])
            send_n = be.to_host(s["spill_n"]).tolist()
            recv_n = be.to_host(r["spill_n"]).tolist()
            any_spill = self._agree(1 if (sum(send_n) or sum(recv_n)) else 0, dist.ReduceOp.MAX)
            self._a2a(r["bins"], s["bins"])
            self._a2a(r["cursors"], s["cursors"])
            rec = self._exchange_spill(b, send_n, recv_n) if any_spill else None
            return rec, (int(sum(recv_n)) if any_spill else 0)

        if use_cuda:
            # routing of set b is complete (overflowed() synchronised); receive set b was drained two chunks ago
            with torch.cuda.stream(be.comm_stream):
                rec, total = body()
                if rec is not None:
                    rec.record_stream(be.stream)
                ev = torch.cuda.Event()
                ev.record(be.comm_stream)
        else:
            rec, total = body()
            ev = None
        return b, ev, rec, total

    def distinct_global(self):
        return self._agree(self.be.distinct(), dist.ReduceOp.SUM)


# ---------------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): one rank per GPU, weak scaling (every rank brings wl["reads"] reads)
# ---------------------------------------------------------------------------------------------------------
def bench_main(args, wl, rank, world, local_rank, log=lambda m: None):
    import json
    import statistics
    import time

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = _lib.load()
    shard_bits = int(math.log2(world))
    assert 1 << shard_bits == world, "the table is sharded by hash bits: N must be a power of two"
    k, l_global = wl["k"], wl["l"] + shard_bits
    n_reads, read_len = wl["reads"], wl["read_len"]
    n_bases = n_reads * read_len
    n_words = (n_bases + 31) // 32
    n_kmers = n_reads * max(0, read_len - k + 1)

    d_packed = torch.empty(n_words + 8, dtype=torch.int64, device=dev)
    d_off = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
    gp = _lib.TsxcGenParams(wl["seed"], n_reads * world, read_len, wl["mode"], wl["genome"], wl["sub"], 0)
    _lib.check(lib.tsxc_gen_reads_device(C.byref(gp), rank * n_reads, n_reads, local_rank, None, d_packed.data_ptr(), d_off.data_ptr()))
    torch.cuda.synchronize()
    # bins are sized for the k-mers a read of this length yields (+2 %); anything beyond goes through the spill path
    be = CudaBackend(k, l_global, 0, rank, world, local_rank,
                     kmers_per_position=min(1.0, 1.02 * max(0, read_len - k + 1) / read_len))
    sc = ShardedCounter(be, rank, world)
    layout = be.hm.stats()

    def fence():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        be.hm.clear()
        be.hm.sync()
        fence()
        t0 = time.perf_counter()
        sc.add_reads_device(d_packed, d_off, n_reads, n_bases)
        fence()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        dt = one_step()
        if rank == 0:
            log(f"warmup {i}: {dt * 1e3:.1f} ms")
    sampler = None
    if rank == 0:
        import bench as _bench
        sampler = _bench.ClockSampler(local_rank)
        sampler.start()
    steps = [one_step() for _ in range(args.steps)]
    clocks = sampler.stop() if sampler else None
    st = be.hm.stats()
    added = torch.tensor([st["kmers_added"], st["distinct"], st["kernel_launches"], st["error_flags"]], dtype=torch.int64, device=dev)
    dist.all_reduce(added, op=dist.ReduceOp.SUM)
    total_added, total_distinct, launches, errs = [int(x) for x in added.tolist()]
    assert total_added == n_kmers * world and errs == 0, (total_added, n_kmers * world, errs)
    T = sum(steps)
    value = args.steps * n_kmers * world / T / 1e9

    # e2e: the rank's reads start in pinned host memory; H2D copy + routed counting + global distinct read-back
    e2e = None
    if not args.no_e2e:
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
            d_packed.copy_(h_packed, non_blocking=True)
            d_off.copy_(h_off, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            sc.add_reads_device(d_packed, d_off, n_reads, n_bases)
            got = sc.distinct_global()
            fence()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert got == total_distinct
            if it:
                times.append(float(t.item()))
        e2e = {"value": n_kmers * world / statistics.mean(times) / 1e9, "unit": "Gk-mer/s",
               "h2d_bytes_per_step": world * ((n_words + 8) * 8 + (n_reads + 1) * 8), "d2h_bytes_per_step": world * 8,
               "timing": "wall clock, barrier + synchronize on both sides, max over ranks"}

    if rank == 0:
        E = 8 * layout["entry_words"]
        in_b = 0.25 * read_len / max(1, read_len - k + 1)
        import bench as _bench
        peak, peak_src = _bench.read_peaks()
        main_ms = (st["partition_ms"] + st["insert_ms"])
        achieved = n_kmers * (2 * E + in_b) / (T / args.steps) / 1e9   # per GPU: this rank's k-mers over the step time
        line = {
            "metric": "k-mers counted/sec", "value": value, "unit": "Gk-mer/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] + f" x{world} ranks, table hash-sharded over {world} GPUs (config 5 routing: "
                                                "bin by owner -> NCCL all-to-all over NVLink -> insert)",
                       "k": k, "l_global": l_global, "reads_per_gpu": n_reads, "kmers_per_step": n_kmers * world,
                       "distinct": total_distinct, "entry_bytes": E, "table_bytes_per_gpu": layout["table_bytes"],
                       "chunks_per_step": sc.chunks // max(1, args.warmup + args.steps + (0 if args.no_e2e else 1 + args.steps)),
                       "a2a_bytes_per_gpu_per_step": sc.a2a_bytes // max(1, args.warmup + args.steps + (0 if args.no_e2e else 1 + args.steps)),
                       "l2": "inputs and table shards far exceed the 126 MB L2; shards re-zeroed between steps",
                       "timing": "wall clock per step between barrier+synchronize fences, max over ranks; zeroing untimed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_partition_reads<ROUTE> + k_insert_partitions (per GPU)",
                         "algorithmic_bytes_per_kmer": 2 * E + in_b,
                         "phase_ms_rank0_total": {"route": st["partition_ms"], "insert": st["insert_ms"]}},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    fence()
    dist.destroy_process_group()
    return 0
