"""CPU tests (gloo, world_size 2) of the multi-GPU orchestration in tsxcount_b200/multigpu.py: chunk scheduling,
collective chunk splitting on spill overflow, the all-to-all plumbing and the final reduction.  The device is
replaced by a NumPy stand-in with the same buffer layout (bins by (owner, region), fill cursors, holes, spill
records); hashing goes through the product's own bijective hash (tsxc_debug_hash runs on the host)."""
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

HOLE = np.uint64(0xFFFFFFFFFFFFFFFF)


class NumpyBackend:
    """Stand-in for CudaBackend: same methods, same buffer layout, a dict as the shard's table."""

    def __init__(self, k, l_global, rank, world, chunk_words, bin_cap, spill_cap, bins_per_shard=2):
        import tsxcount_b200 as tsx
        self.tsx, self.lib = tsx, tsx._lib.load()
        self.k, self.rank, self.world = k, rank, world
        st = tsx._lib.TsxcStats()
        assert self.lib.tsxc_debug_layout(k, l_global, 4, 0, world, C.byref(st)) == 0
        self.kw = st.key_words
        self.lbg = 2 * k - st.quotient_bits
        self.shard_bits = int(np.log2(world))
        self.lbl = self.lbg - self.shard_bits
        self.pb = int(np.log2(bins_per_shard))
        self.lay = types.SimpleNamespace(n_shards=world, bins_per_shard=bins_per_shard, key_words=self.kw,
                                         spill_record_words=self.kw + 1, chunk_words=chunk_words, bin_cap=bin_cap,
                                         block_words=bins_per_shard * bin_cap * self.kw, spill_cap=spill_cap)
        self.table = {}
        self._over = False
        self.ends = None

    def alloc_u64(self, n):
        return torch.zeros(max(int(n), 1), dtype=torch.int64)

    def to_host(self, t):
        return t

    def prepare(self, d_offsets, n_reads, n_bases):
        self.offsets = d_offsets.numpy().astype(np.uint64)

    def _hash(self, words):
        key = np.array(words, dtype=np.uint64)
        out = np.zeros(self.kw, dtype=np.uint64)
        assert self.lib.tsxc_debug_hash(self.k, key.ctypes.data, out.ctypes.data) == 0
        return out

    def route(self, d_packed, n_bases, w0, w1, bins, cursors, spill, spill_n):
        packed = d_packed.numpy().view(np.uint64)
        big = 0
        for j, w in enumerate(packed.tolist()):
            big |= w << (64 * j)
        B, CUR = bins.numpy().view(np.uint64), cursors.numpy().view(np.uint64)
        SP, SPN = spill.numpy().view(np.uint64), spill_n.numpy().view(np.uint64)
        CUR[:] = 0
        SPN[:] = 0
        lay, kw, k = self.lay, self.kw, self.k
        mask = (1 << (2 * k)) - 1
        for r in range(len(self.offsets) - 1):
            b, e = int(self.offsets[r]), int(self.offsets[r + 1])
            for g in range(b, e - k + 1):
                if not (w0 * 32 <= g < w1 * 32):
                    continue
                v = (big >> (2 * g)) & mask
                h = self._hash([(v >> (64 * j)) & (2**64 - 1) for j in range(kw)])
                bg = int(h[0]) & ((1 << self.lbg) - 1)
                p = bg >> (self.lbl - self.pb)                      # (owner, region) bin
                owner = p >> self.pb
                n = int(CUR[p])
                if n < lay.bin_cap and (g % 7):                     # every 7th k-mer takes the spill path
                    B[(p * lay.bin_cap + n) * kw:(p * lay.bin_cap + n + 1) * kw] = h
                    CUR[p] = n + 1
                else:
                    m = int(SPN[owner])
                    if m >= lay.spill_cap:
                        self._over = True
                        continue
                    base = (owner * lay.spill_cap + m) * (kw + 1)
                    SP[base:base + kw] = h
                    SP[base + kw] = 1
                    SPN[owner] = m + 1

    def overflowed(self):
        o, self._over = self._over, False
        return o

    def insert(self, bins, cursors, n_sources):
        B, CUR = bins.numpy().view(np.uint64), cursors.numpy().view(np.uint64)
        lay, kw = self.lay, self.kw
        for p in range(n_sources * lay.bins_per_shard):
            for i in range(min(int(CUR[p]), lay.bin_cap)):
                h = tuple(B[(p * lay.bin_cap + i) * kw:(p * lay.bin_cap + i + 1) * kw].tolist())
                if h[0] == int(HOLE):
                    continue
                self.table[h] = self.table.get(h, 0) + 1

    def insert_spill(self, records, n):
        R = records.numpy().view(np.uint64)
        kw = self.kw
        for i in range(n):
            h = tuple(R[i * (kw + 1):i * (kw + 1) + kw].tolist())
            self.table[h] = self.table.get(h, 0) + int(R[i * (kw + 1) + kw])

    def sync(self):
        pass

    def distinct(self):
        return len(self.table)


def _worker(rank, world, port, k, spill_cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py as orc
        import tsxcount_b200 as tsx
        from tsxcount_b200.multigpu import ShardedCounter
        all_seqs = orc.gen_reads(seed=77, n_reads=24, read_len=90, mode=1)
        mine = all_seqs[rank::world] if rank == 0 else all_seqs[rank::world][:-3]   # ragged: ranks differ in size
        ascii_, offsets = tsx.sequtils.concat_reads(mine)
        packed, seg, _ = tsx.sequtils.pack_reads(ascii_, offsets)
        n_bases = int(seg[-1])
        be = NumpyBackend(k, 16, rank, world, chunk_words=16, bin_cap=4096, spill_cap=spill_cap)
        sc = ShardedCounter(be, rank, world, min_split_words=1)
        sc.add_reads_device(torch.from_numpy(packed.view(np.int64).copy()), torch.from_numpy(seg.view(np.int64).copy()),
                            len(seg) - 1, n_bases)
        total_distinct = sc.distinct_global()
        # every k-mer this rank stores must be owned by it
        for h in be.table:
            assert ((h[0] & ((1 << be.lbg) - 1)) >> be.lbl) == rank
        q.put((rank, dict(be.table), total_distinct, sc.chunks, sc.retries, [bytes(s) for s in mine]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,spill_cap", [(21, 4096), (21, 24), (40, 4096)])
def test_sharded_counter_world2_gloo(k, spill_cap):
    import oracle_py as orc
    import tsxcount_b200 as tsx
    world = 2
    port = 29500 + (os.getpid() % 2000) + k
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, spill_cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lib = tsx._lib.load()
    kw = tsx.sequtils.key_words(k)
    merged, seqs = {}, []
    for rank, table, total_distinct, chunks, retries, mine in res:
        for h, c in table.items():
            assert h not in merged, "k-mer stored by two shards"
            merged[h] = c
        seqs += mine
    oc = orc.count_seqs(seqs, k)
    assert all(r[2] == oc.n_distinct for r in res)                   # all_reduce(SUM) of the shards' distinct counts
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]          # same number of exchanges / splits on all ranks
    if spill_cap < 100:
        assert res[0][4] > 0                                          # the overflow forced collective chunk splits
    got = {}
    for h, c in merged.items():
        hv = np.array(h, dtype=np.uint64)
        out = np.zeros(kw, dtype=np.uint64)
        assert lib.tsxc_debug_unhash(k, hv.ctypes.data, out.ctypes.data) == 0
        got[tuple(out.tolist())] = c
    assert got == oc.as_dict(kw)
