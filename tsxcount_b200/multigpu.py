"""Hash-sharded counting over the GPUs of one box: one process per GPU, `torch.distributed` for the plumbing.

The table is partitioned by the top log2(world) bits of the bucket index of the bijective k-mer hash, so shards
never exchange anything after insertion: distinct counts add up, dumps concatenate.  A batch of reads is counted in
rounds (the reference has no counterpart, it is one process on one shared table,
src/mains/main.cpp:132-218 of mjoppich/tsxCount).  Per round, everything queued on the handle's stream, no host
synchronisation:

  hist     tsxc_route_hist     this rank's exact k-mer counts per routing bin (owner-major)
  gather   all_gather (NCCL)   every rank learns every rank's counts; it also tells a sender that all receive
                               buffers have been drained (a rank enters it only after its previous insert)
  send     tsxc_route_send     offsets from the gathered counts, then the routing kernel (extract + hash + tile sort)
                               stores its runs straight into the owners' peer-mapped receive buffers over NVLink
  barrier  all_reduce (NCCL)   all stores have landed
  insert   tsxc_route_insert   sort what was received by fine table region and insert it

`CudaRouteBackend` is the product; tests/ substitute a NumPy stand-in with the same methods to exercise the
orchestration with the gloo backend on CPUs.
"""
import contextlib
import ctypes as C
import math

import torch
import torch.distributed as dist

from . import _lib
from .hashmap import TSXHashMapCUDA


class CudaRouteBackend:
    """One shard on one GPU.  Kernels and collectives run on the handle's own stream."""

    def __init__(self, k, l_global, s, rank, world, device, flags=0):
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.dev_index = device
        self.hm = TSXHashMapCUDA(l_global, s, k, device=device, flags=flags, shard_rank=rank, n_shards=world)
        self.lib = self.hm._lib
        self.stream = torch.cuda.ExternalStream(self.lib.tsxc_stream(self.hm.handle), device=self.device)
        self.bins = self.hm.routeInfo().bins
        self._opened = []

    def stream_ctx(self):
        return torch.cuda.stream(self.stream)

    def new_i32(self, n):
        return torch.zeros(max(int(n), 1), dtype=torch.int32, device=self.device)

    def recv_buffer(self, cap_keys=0):
        return self.hm.routeRecvBuffer(cap_keys)

    def export_handle(self, ptr):
        h = (C.c_ubyte * _lib.TSXC_IPC_HANDLE_BYTES)()
        _lib.check(self.lib.tsxc_ipc_export_mem(self.dev_index, C.c_void_p(ptr), h))
        return bytes(h)

    def open_handle(self, handle):
        h = (C.c_ubyte * _lib.TSXC_IPC_HANDLE_BYTES)(*handle)
        p = C.c_void_p()
        _lib.check(self.lib.tsxc_ipc_open_mem(self.dev_index, h, C.byref(p)))
        self._opened.append(p)
        return p.value

    def set_peers(self, ptrs, cap_keys):
        self.hm.routeSetPeers(ptrs, cap_keys)

    def begin(self, d_packed, d_offsets, n_reads, n_bases):
        return self.hm.routeBegin(d_packed.data_ptr(), d_offsets.data_ptr(), n_reads, n_bases)

    def hist(self, rnd, hist):
        self.hm.routeHist(rnd, hist.data_ptr())

    def send(self, rnd, hist_all):
        self.hm.routeSend(rnd, hist_all.data_ptr())

    def insert(self):
        self.hm.routeInsert()

    def sync(self):
        self.hm.sync()

    def distinct(self):
        return self.hm.getKmerCount()

    def close(self):
        self.hm.sync()
        for p in self._opened:
            self.lib.tsxc_ipc_close_mem(self.dev_index, p)
        self._opened = []
        self.hm.close()


class ShardedCounter:
    """Counts the k-mers of this rank's reads into the table sharded over all ranks of `group`."""

    def __init__(self, backend, rank, world, group=None, recv_cap_keys=0):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        ptr, cap = backend.recv_buffer(recv_cap_keys)
        if world > 1:
            infos = [None] * world
            dist.all_gather_object(infos, (backend.export_handle(ptr), cap), group=group)
            self.recv_cap = min(c for _, c in infos)
            peers = [ptr if o == rank else backend.open_handle(infos[o][0]) for o in range(world)]
        else:
            self.recv_cap, peers = cap, [ptr]
        backend.set_peers(peers, self.recv_cap)
        self.hist = backend.new_i32(backend.bins)
        self.hist_all = backend.new_i32(backend.bins * world)
        self.flag = backend.new_i32(1)
        self.rounds = 0
        self.batches = 0

    def _agree_max(self, value):
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.int64, device=self.hist.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def add_reads_device(self, d_packed, d_offsets, n_reads, n_bases):
        """Queues the whole batch; returns without waiting for the device (call backend.sync())."""
        be = self.be
        rounds = self._agree_max(be.begin(d_packed, d_offsets, n_reads, n_bases))   # the only host round trip
        ctx = be.stream_ctx() if hasattr(be, "stream_ctx") else contextlib.nullcontext()
        for r in range(rounds):
            be.hist(r, self.hist)
            with ctx:
                if self.world > 1:
                    dist.all_gather_into_tensor(self.hist_all, self.hist, group=self.group)
                else:
                    self.hist_all.copy_(self.hist)
            be.send(r, self.hist_all)
            if self.world > 1:
                with ctx:
                    dist.all_reduce(self.flag, group=self.group)      # barrier on the stream: stores have landed
            be.insert()
        self.rounds += rounds
        self.batches += 1

    def distinct_global(self):
        v = self.be.distinct()
        if self.world == 1:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=self.hist.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())


