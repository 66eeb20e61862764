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
        self.kmers_per_position = kmers_per_position
        self.fixed_chunk = max_chunk_words
        self.lay = self.hm.routeLayout(max_chunk_words, kmers_per_position)
        self.kw = self.lay.key_words
        self.stream = torch.cuda.ExternalStream(self.hm._lib.tsxc_stream(self.hm.handle), device=self.device)
        self.comm_stream = torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)   # peer copies (copy engines)

    # buffers -----------------------------------------------------------------------------------------
    def buffer_bytes(self, lay):
        """Two send sets + two receive sets of bins, cursors and spill lists for this layout."""
        G = lay.n_shards
        one = G * lay.block_words * 8 + G * lay.bins_per_shard * 8
        return 4 * one + 2 * G * lay.spill_cap * (lay.key_words + 1) * 8

    def candidate_layouts(self):
        """Largest chunk first: the insert pass of a chunk touches every table region once, so bigger chunks mean
        denser, more local passes; the limit is free HBM next to the shard."""
        if self.fixed_chunk:
            return [self.lay]
        kw = self.lay.key_words
        return [self.hm.routeLayout(w // kw, self.kmers_per_position) for w in (1 << 25, 3 << 23, 1 << 24, 1 << 23, 1 << 22)]

    def fits(self, lay, reserve=3 << 30):
        free, _ = torch.cuda.mem_get_info(self.device)
        return self.buffer_bytes(lay) + reserve <= free

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


class RawBuf:
    """A cudaMalloc'ed buffer (exportable through CUDA IPC) with the one tensor method the counter needs."""

    def __init__(self, lib, device, nbytes):
        self.lib, self.device, self.nbytes = lib, device, int(nbytes)
        self.ptr = C.c_void_p()
        _lib.check(lib.tsxc_device_alloc(device, max(self.nbytes, 8), C.byref(self.ptr)))

    def data_ptr(self):
        return self.ptr.value

    def free(self):
        if self.ptr.value:
            self.lib.tsxc_device_free(self.device, self.ptr)
            self.ptr = C.c_void_p()


class PeerExchange:
    """Optional exchange (TSXC_EXCHANGE=peer; the default is NCCL): bin blocks go to their owner with copy-engine
    peer copies over NVLink (no SMs, no staging).  Measured on 8 B200s it is SLOWER than the NCCL all-to-all
    (762 vs 730 ms per step): the copies run at full link speed and take HBM bandwidth from the routing kernel
    (332 vs 294 ms), which is what the exchange overlaps with.  Kept because it is the building block for the
    next step (routing kernel storing straight into peer memory).  Mechanics:
    the owner's receive buffers are CUDA-IPC mapped into every sender; per buffer set two inter-process events
    order `copies landed -> insert` and `buffer drained -> next copies`.  The ordering of the event calls across
    processes comes from the small host-synchronised collectives every chunk performs anyway (see
    ShardedCounter.add_reads_device).  Raises at construction if IPC is unavailable; the caller then falls back to
    NCCL (collectively)."""

    H = 64

    def __init__(self, be, rank, world, group, lay):
        self.be, self.rank, self.world = be, rank, world
        lib, dev = be.hm._lib, be.device.index
        self.lib, self.dev = lib, dev
        self.block_bytes = lay.block_words * 8
        self.cur_bytes = lay.bins_per_shard * 8
        self.recv_bins = [RawBuf(lib, dev, world * self.block_bytes) for _ in range(2)]
        self.recv_cur = [RawBuf(lib, dev, world * self.cur_bytes) for _ in range(2)]
        mine = torch.zeros(8, self.H, dtype=torch.uint8)
        self.ev_sent, self.ev_drained = [], []
        for b in range(2):
            for j, buf in ((0, self.recv_bins[b]), (1, self.recv_cur[b])):
                h = (C.c_ubyte * self.H)()
                _lib.check(lib.tsxc_ipc_export_mem(dev, buf.ptr, h))
                mine[2 * b + j] = torch.tensor(list(h), dtype=torch.uint8)
            for j, store in ((0, self.ev_sent), (1, self.ev_drained)):
                ev, h = C.c_void_p(), (C.c_ubyte * self.H)()
                _lib.check(lib.tsxc_ipc_event_create(dev, C.byref(ev), h))
                store.append(ev)
                mine[4 + 2 * b + j] = torch.tensor(list(h), dtype=torch.uint8)
        allh = torch.empty(world, 8, self.H, dtype=torch.uint8, device=be.device)
        dist.all_gather_into_tensor(allh, mine.to(be.device), group=group)
        allh = allh.cpu()
        self.peer_bins = [[None, None] for _ in range(world)]
        self.peer_cur = [[None, None] for _ in range(world)]
        self.peer_sent = [[None, None] for _ in range(world)]
        self.peer_drained = [[None, None] for _ in range(world)]
        for o in range(world):
            for b in range(2):
                if o == rank:
                    self.peer_bins[o][b], self.peer_cur[o][b] = self.recv_bins[b].ptr, self.recv_cur[b].ptr
                    self.peer_sent[o][b], self.peer_drained[o][b] = self.ev_sent[b], self.ev_drained[b]
                    continue
                for j, store in ((0, self.peer_bins), (1, self.peer_cur)):
                    h = (C.c_ubyte * self.H)(*allh[o, 2 * b + j].tolist())
                    p = C.c_void_p()
                    _lib.check(lib.tsxc_ipc_open_mem(dev, h, C.byref(p)))
                    store[o][b] = p
                for j, store in ((0, self.peer_sent), (1, self.peer_drained)):
                    h = (C.c_ubyte * self.H)(*allh[o, 4 + 2 * b + j].tolist())
                    ev = C.c_void_p()
                    _lib.check(lib.tsxc_ipc_event_open(dev, h, C.byref(ev)))
                    store[o][b] = ev
        self.used = [0, 0]

    def send(self, b, send_bins, send_cursors, comm_stream_ptr):
        """Queue the copies of buffer set b on the comm stream and record `sent`."""
        lib, dev, G = self.lib, self.dev, self.world
        if self.used[b]:
            for o in range(G):   # the owner must have drained what we sent two chunks ago
                _lib.check(lib.tsxc_stream_wait_event(dev, comm_stream_ptr, self.peer_drained[o][b]))
        for i in range(G):
            o = (self.rank + i) % G   # stagger the destinations
            _lib.check(lib.tsxc_copy_async(dev, C.c_void_p(self.peer_bins[o][b].value + self.rank * self.block_bytes),
                                           C.c_void_p(send_bins.data_ptr() + o * self.block_bytes), self.block_bytes,
                                           comm_stream_ptr))
            _lib.check(lib.tsxc_copy_async(dev, C.c_void_p(self.peer_cur[o][b].value + self.rank * self.cur_bytes),
                                           C.c_void_p(send_cursors.data_ptr() + o * self.cur_bytes), self.cur_bytes,
                                           comm_stream_ptr))
        _lib.check(lib.tsxc_event_record(dev, self.ev_sent[b], comm_stream_ptr))
        self.used[b] += 1

    def wait_all_sent(self, b, stream_ptr):
        for o in range(self.world):
            _lib.check(self.lib.tsxc_stream_wait_event(self.dev, stream_ptr, self.peer_sent[o][b]))

    def mark_drained(self, b, stream_ptr):
        _lib.check(self.lib.tsxc_event_record(self.dev, self.ev_drained[b], stream_ptr))


class ShardedCounter:
    """Counts the k-mers of this rank's reads into the table sharded over all ranks of `group`."""

    def __init__(self, backend, rank, world, group=None, min_split_words=1024):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        self.min_split_words = min_split_words
        if hasattr(backend, "candidate_layouts"):
            # every rank must use the same geometry: take the largest chunk that fits on ALL ranks
            cands = backend.candidate_layouts()
            mine = next((i for i, c in enumerate(cands) if backend.fits(c)), len(cands) - 1)
            t = torch.tensor([mine], dtype=torch.int64, device=backend.device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
            backend.lay = cands[int(t.item())]
        lay = backend.lay
        assert lay.n_shards == world
        self.lay = lay
        G, kw = world, lay.key_words
        # double-buffered send / receive sets: the exchange of chunk c overlaps the routing of chunk c+1
        self.send = [dict(bins=backend.alloc_u64(G * lay.block_words), cursors=backend.alloc_u64(G * lay.bins_per_shard),
                          spill=backend.alloc_u64(G * lay.spill_cap * (kw + 1)), spill_n=backend.alloc_u64(G))
                     for _ in range(2)]
        self.recv = [dict(spill_n=backend.alloc_u64(G)) for _ in range(2)]
        self.a2a_bytes = 0
        self.chunks = 0
        self.retries = 0
        self.peer = None
        import os
        if world > 1 and isinstance(backend, CudaBackend) and os.environ.get("TSXC_EXCHANGE", "nccl") == "peer":
            ok = 1
            try:
                peer = PeerExchange(backend, rank, world, group, lay)
            except Exception as e:  # IPC not available on this host
                ok, peer = 0, None
                self.peer_error = str(e)
            if self._agree(ok, dist.ReduceOp.MIN):
                self.peer = peer
                for b in range(2):   # the receive side lives in the IPC-exported buffers
                    self.recv[b]["bins"], self.recv[b]["cursors"] = peer.recv_bins[b], peer.recv_cur[b]
        if self.peer is None:
            for b in range(2):
                self.recv[b]["bins"] = backend.alloc_u64(G * lay.block_words)
                self.recv[b]["cursors"] = backend.alloc_u64(G * lay.bins_per_shard)
        self.exchange = "peer copies (CUDA IPC, copy engines)" if self.peer else ("nccl all_to_all" if world > 1 else "local")

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
                    if self.peer:
                        self.peer.wait_all_sent(pb, be.stream.cuda_stream)
                be.insert(self.recv[pb]["bins"], self.recv[pb]["cursors"], self.world)
                if self.peer:
                    self.peer.mark_drained(pb, be.stream.cuda_stream)
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
            if self.peer:
                # copies + `sent` record first, on their own stream: the collectives below (on the comm stream, so
                # that their host synchronisation does not wait for the copies) are the host-level barrier that
                # orders this record before every receiver's wait and every `drained` record before the next send
                self.peer.send(b, s["bins"], s["cursors"], be.copy_stream.cuda_stream)
                self.a2a_bytes += (self.world - 1) * (self.peer.block_bytes + self.peer.cur_bytes)
            self._a2a(r["spill_n"], s["spill_n"])
            send_n = be.to_host(s["spill_n"]).tolist()
            recv_n = be.to_host(r["spill_n"]).tolist()
            any_spill = self._agree(1 if (sum(send_n) or sum(recv_n)) else 0, dist.ReduceOp.MAX)
            if not self.peer:
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
    be = CudaBackend(k, l_global, 0, rank, world, local_rank, flags=wl.get("flags", 0),
                     kmers_per_position=min(1.0, 1.02 * max(0, read_len - k + 1) / read_len))
    sc = ShardedCounter(be, rank, world)
    log(f"rank {rank}: exchange = {sc.exchange} {getattr(sc, 'peer_error', '')}")
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
        n_calls = max(1, args.warmup + args.steps + (0 if args.no_e2e else 1 + args.steps))
        a2a_step = sc.a2a_bytes // n_calls                               # bytes this rank sent to peers per step
        achieved = n_kmers * (2 * E + in_b) / (T / args.steps) / 1e9   # per GPU: this rank's k-mers over the step time
        line = {
            "metric": "k-mers counted/sec", "value": value, "unit": "Gk-mer/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] + f" x{world} ranks, table hash-sharded over {world} GPUs (config 5 routing: "
                                                "bin by owner -> NCCL all-to-all over NVLink -> insert)",
                       "k": k, "l_global": l_global, "reads_per_gpu": n_reads, "kmers_per_step": n_kmers * world,
                       "distinct": total_distinct, "entry_bytes": E, "table_bytes_per_gpu": layout["table_bytes"],
                       "exchange": sc.exchange, "chunks_per_step": sc.chunks // n_calls,
                       "a2a_bytes_per_gpu_per_step": a2a_step,
                       "l2": "inputs and table shards far exceed the 126 MB L2; shards re-zeroed between steps",
                       "timing": "wall clock per step between barrier+synchronize fences, max over ranks; zeroing untimed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": "k_partition_reads (bins by owner and region) + k_insert_partitions (per GPU)",
                         "algorithmic_bytes_per_kmer": 2 * E + in_b,
                         "phase_ms_rank0_total": {"route": st["partition_ms"], "insert": st["insert_ms"]}},
            "nvlink": {"sent_bytes_per_gpu_per_step": a2a_step, "avg_GB_s_per_gpu_per_direction": a2a_step / (T / args.steps) / 1e9,
                       "note": "payload of the all-to-all averaged over the whole step; the exchange of chunk c runs on a side "
                               "stream while chunk c+1 is routed and chunk c-1 inserted"},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    fence()
    dist.destroy_process_group()
    return 0
