"""CPU tests (gloo, world_size 2) of the multi-GPU orchestration in tsxcount_b200/multigpu.py: the round protocol
(histogram -> all_gather -> send -> barrier -> insert), ranks that need different numbers of rounds, the setup
exchange of buffer handles and the final reduction.  The device is replaced by a NumPy stand-in with the same
methods; hashing goes through the product's own bijective hash (tsxc_debug_hash runs on the host).  The stand-in
restates k_route_offsets (tsxcount_b200/csrc/tsx_radix.cuh): where bin (owner, local coarse bin) of source s starts in
the owner's receive buffer, and checks the properties the CUDA path relies on: the positions all ranks compute from
the gathered histograms tile every receive buffer exactly, without overlap, bin-major."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class NumpyRouteBackend:
    """Stand-in for CudaRouteBackend: same methods, a dict as the shard's table.  The peer stores of the routing
    kernel are emulated with an all-to-all of (position, key) pairs."""

    def __init__(self, k, l_global, rank, world, local_bits, chunk_keys):
        import tsxcount_b200 as tsx
        self.tsx, self.lib = tsx, tsx._lib.load()
        self.k, self.rank, self.world = k, rank, world
        st = tsx._lib.TsxcStats()
        assert self.lib.tsxc_debug_layout(k, l_global, 4, 0, world, C.byref(st)) == 0
        self.kw = st.key_words
        self.lbg = 2 * k - st.quotient_bits
        self.shard_bits = int(np.log2(world))
        self.nbl = 1 << local_bits
        self.bins = world * self.nbl
        self.shift1 = self.lbg - self.shard_bits - local_bits
        self.chunk_keys = chunk_keys
        self.table = {}
        self.recv = None
        self.n_recv = 0
        self.sent_rounds = 0

    def new_i32(self, n):
        return torch.zeros(max(int(n), 1), dtype=torch.int32)

    def recv_buffer(self, cap_keys=0):
        self.cap = cap_keys or 1 << 20
        return 0x1000 + self.rank, self.cap

    def export_handle(self, ptr):
        return ptr.to_bytes(8, "little") * 8

    def open_handle(self, handle):
        return int.from_bytes(handle[:8], "little")

    def set_peers(self, ptrs, cap_keys):
        assert ptrs == [0x1000 + r for r in range(self.world)]
        self.cap = cap_keys
        self.recv = np.zeros((self.cap, self.kw), dtype=np.uint64)

    def _hash(self, words):
        key = np.array(words, dtype=np.uint64)
        out = np.zeros(self.kw, dtype=np.uint64)
        assert self.lib.tsxc_debug_hash(self.k, key.ctypes.data, out.ctypes.data) == 0
        return out

    def begin(self, d_packed, d_offsets, n_reads, n_bases):
        packed = d_packed.numpy().view(np.uint64)
        offsets = d_offsets.numpy().view(np.uint64)
        big = 0
        for j, w in enumerate(packed.tolist()):
            big |= w << (64 * j)
        k, kw = self.k, self.kw
        mask = (1 << (2 * k)) - 1
        hashes = []
        for r in range(n_reads):
            b, e = int(offsets[r]), int(offsets[r + 1])
            for g in range(b, e - k + 1):
                v = (big >> (2 * g)) & mask
                hashes.append(self._hash([(v >> (64 * j)) & (2**64 - 1) for j in range(kw)]))
        H = np.array(hashes, dtype=np.uint64).reshape(-1, kw)
        self.chunks = [H[i:i + self.chunk_keys] for i in range(0, len(H), self.chunk_keys)]
        return len(self.chunks) + 1            # an upper bound, like the product: the surplus round sends nothing

    def _digit1(self, H):
        return ((H[:, 0] & np.uint64((1 << self.lbg) - 1)) >> np.uint64(self.shift1)).astype(np.int64)

    def hist(self, rnd, hist):
        h = np.zeros(self.bins, dtype=np.int32)
        if rnd < len(self.chunks):
            h = np.bincount(self._digit1(self.chunks[rnd]), minlength=self.bins).astype(np.int32)
        hist.copy_(torch.from_numpy(h))

    def send(self, rnd, hist_all):
        G, nbl, kw = self.world, self.nbl, self.kw
        Hall = hist_all.numpy().reshape(G, self.bins).astype(np.int64)
        col = Hall.sum(axis=0)
        ex = np.concatenate([[0], np.cumsum(col)])
        before = Hall[:self.rank].sum(axis=0)
        # k_route_offsets: position of (me, bin d) in the owner's buffer
        start = np.array([ex[d] - ex[(d // nbl) * nbl] + before[d] for d in range(self.bins)], dtype=np.int64)
        for o in range(G):
            assert ex[(o + 1) * nbl] - ex[o * nbl] <= self.cap, "receive buffer overflow"
        self.cur_coff = ex[self.rank * nbl:(self.rank + 1) * nbl + 1] - ex[self.rank * nbl]
        chunk = self.chunks[rnd] if rnd < len(self.chunks) else np.zeros((0, kw), dtype=np.uint64)
        d = self._digit1(chunk)
        order = np.argsort(d, kind="stable")
        chunk, d = chunk[order], d[order]
        pos = np.zeros(len(chunk), dtype=np.int64)
        fill = start.copy()
        for i, b in enumerate(d.tolist()):
            pos[i] = fill[b]
            fill[b] += 1
        owner = d // nbl
        send_counts = [int((owner == o).sum()) for o in range(G)]
        recv_counts = [int(Hall[s, self.rank * nbl:(self.rank + 1) * nbl].sum()) for s in range(G)]
        payload = np.concatenate([pos.reshape(-1, 1).astype(np.uint64), chunk], axis=1).astype(np.uint64).view(np.int64)
        inp = torch.from_numpy(np.ascontiguousarray(payload).reshape(-1))
        out = torch.zeros(sum(recv_counts) * (kw + 1), dtype=torch.int64)
        dist.all_to_all_single(out, inp, output_split_sizes=[c * (kw + 1) for c in recv_counts],
                               input_split_sizes=[c * (kw + 1) for c in send_counts])
        got = out.numpy().view(np.uint64).reshape(-1, kw + 1)
        n = int(self.cur_coff[-1])
        assert len(got) == n
        seen = np.zeros(n, dtype=bool)
        for row in got:
            p = int(row[0])
            assert 0 <= p < n and not seen[p], "two senders were given the same position"
            seen[p] = True
            self.recv[p] = row[1:]
        assert seen.all(), "the positions do not tile the receive buffer"
        self.n_recv = n
        self.sent_rounds += 1

    def insert(self):
        lmask = np.uint64(self.nbl - 1)
        for b in range(self.nbl):
            lo, hi = int(self.cur_coff[b]), int(self.cur_coff[b + 1])
            seg = self.recv[lo:hi]
            if len(seg):
                dd = ((seg[:, 0] & np.uint64((1 << self.lbg) - 1)) >> np.uint64(self.shift1))
                assert ((dd & lmask) == b).all() and ((dd >> np.uint64(int(np.log2(self.nbl)))) == self.rank).all(), \
                    "receive buffer is not bin-major"
            for row in seg.tolist():
                self.table[tuple(row)] = self.table.get(tuple(row), 0) + 1
        self.n_recv = 0

    def sync(self):
        pass

    def distinct(self):
        return len(self.table)


def _worker(rank, world, port, k, chunk_keys, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py as orc
        import tsxcount_b200 as tsx
        from tsxcount_b200.multigpu import ShardedCounter
        all_seqs = orc.gen_reads(seed=77, n_reads=24, read_len=90, mode=1)
        mine = all_seqs[rank::world] if rank == 0 else all_seqs[rank::world][:-5]   # ragged: ranks differ in size
        ascii_, offsets = tsx.sequtils.concat_reads(mine)
        packed, seg, _ = tsx.sequtils.pack_reads(ascii_, offsets)
        n_bases = int(seg[-1])
        be = NumpyRouteBackend(k, 16, rank, world, local_bits=2, chunk_keys=chunk_keys)
        sc = ShardedCounter(be, rank, world, recv_cap_keys=4096 + 64 * rank)      # ranks report different capacities
        assert be.cap == 4096                                                     # ... and agree on the smallest
        for _ in range(2):                                                        # two batches: state carries over
            sc.add_reads_device(torch.from_numpy(packed.view(np.int64).copy()), torch.from_numpy(seg.view(np.int64).copy()),
                                len(seg) - 1, n_bases)
        total_distinct = sc.distinct_global()
        q.put((rank, dict(be.table), total_distinct, sc.rounds, be.sent_rounds, len(be.chunks), [bytes(s) for s in mine]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,chunk_keys", [(21, 100000), (21, 150), (40, 97)])
def test_sharded_counter_world2_gloo(k, chunk_keys):
    import oracle_py as orc
    import tsxcount_b200 as tsx
    world = 2
    port = 29500 + (os.getpid() % 2000) + k + chunk_keys % 50
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, chunk_keys, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lib = tsx._lib.load()
    kw = tsx.sequtils.key_words(k)
    merged, seqs = {}, []
    for rank, table, total_distinct, rounds, sent_rounds, n_chunks, mine in res:
        for h, c in table.items():
            assert h not in merged, "k-mer stored by two shards"
            merged[h] = c
        seqs += mine
    oc = orc.count_seqs(seqs, k)
    assert all(r[2] == oc.n_distinct for r in res)                   # all_reduce(SUM) of the shards' distinct counts
    assert res[0][3] == res[1][3] == res[0][4] == res[1][4]          # every rank took part in every round
    assert res[0][3] == 2 * (max(r[5] for r in res) + 1)             # rounds = the maximum over the ranks, per batch
    if chunk_keys < 1000:
        assert res[0][5] != res[1][5] or res[0][5] > 2               # ranks needed different numbers of rounds
    got = {}
    for h, c in merged.items():
        hv = np.array(h, dtype=np.uint64)
        out = np.zeros(kw, dtype=np.uint64)
        assert lib.tsxc_debug_unhash(k, hv.ctypes.data, out.ctypes.data) == 0
        got[tuple(out.tolist())] = c
    assert got == {key: 2 * c for key, c in oc.as_dict(kw).items()}  # two identical batches
