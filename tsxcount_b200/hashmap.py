"""TSXHashMapCUDA — host-side mirror of the reference's TSXHashMap interface over the C ABI.

The reference selects a serialization backend by constructing a TSXHashMap subclass
(src/mains/main.cpp:429-475 of mjoppich/tsxCount) and then calls addKmer per k-mer
(main.cpp:192), getKmerCount(kmer) per checked k-mer (src/mains/testExecution.h:38-50) and
getKmerCount() for the distinct total (main.cpp:222).  This class keeps those names and meanings;
the batch variants (addKmers / addReads / getKmerCounts) are what a GPU needs to be fed.
Errors: the reference throws TSXException for 2k <= l (src/tsxcount/TSXHashMap.h:91-94) and exits
with 42 when the table is full (:340-343); here both surface as TsxcError with .status
TSXC_E_INVALID / TSXC_E_TABLE_FULL.
"""
import ctypes as C

import numpy as np

from . import _lib, sequtils


class TSXHashMapCUDA:
    def __init__(self, l, storage_bits, k, device=0, flags=_lib.TSXC_FLAG_NONE, shard_rank=0, n_shards=1):
        """TSXHashMap(uint8_t iL, uint32_t iStorageBits, uint16_t iK) — TSXHashMap.h:79"""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.k, self.l, self.s = int(k), int(l), int(storage_bits)
        self.device = device
        if n_shards == 1:
            st = self._lib.tsxc_create(self.k, self.l, self.s, device, flags, C.byref(self._h))
        else:
            st = self._lib.tsxc_create_shard(self.k, self.l, self.s, device, flags, shard_rank, n_shards,
                                             C.byref(self._h))
        _lib.check(st, None)
        self.kw = sequtils.key_words(self.k)

    # -- life cycle -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.tsxc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def clear(self):
        _lib.check(self._lib.tsxc_clear(self._h), self._h)

    def sync(self):
        _lib.check(self._lib.tsxc_sync(self._h), self._h)

    def trim(self):
        """Release the insert pipeline's key buffers (re-sized by the next batch)."""
        _lib.check(self._lib.tsxc_trim(self._h), self._h)

    def mark(self, idx):
        _lib.check(self._lib.tsxc_mark(self._h, idx), self._h)

    def elapsed_ms(self, a, b):
        ms = C.c_float(0)
        _lib.check(self._lib.tsxc_mark_elapsed_ms(self._h, a, b, C.byref(ms)), self._h)
        return ms.value

    # -- reference getters (TSXHashMap.h:162-177) -------------------------------------------------
    def getK(self):
        return self.k

    def getMaxElements(self):
        return self.stats()["n_slots"]

    def getUsedPositions(self):
        return self.stats()["used_slots"]

    # -- insert path ------------------------------------------------------------------------------
    def addKmer(self, kmer):
        """bool TSXHashMap::addKmer(UBigInt& kmer) — TSXHashMap.h:182.  kmer: str or uint64[KW]."""
        arr = sequtils.from_sequence(kmer) if isinstance(kmer, str) else np.asarray(kmer, dtype=np.uint64)
        self.addKmers(arr.reshape(1, self.kw))
        return True

    def addKmers(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.kw)
        _lib.check(self._lib.tsxc_add_kmers(self._h, kmers.ctypes.data, kmers.shape[0]), self._h)
        self.sync()  # the numpy buffer may be released as soon as we return

    def addReads(self, packed, offsets, sync=True):
        """createKMers + fromSequence + addKmer for a packed batch — main.cpp:159-192."""
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _lib.check(self._lib.tsxc_add_reads(self._h, packed.ctypes.data, offsets.ctypes.data, len(offsets) - 1), self._h)
        if sync:
            self.sync()
        return packed, offsets  # caller keeps them alive until sync() when sync=False

    def addSequences(self, seqs):
        """seqs: list of bytes/str reads."""
        seqs = [s.encode() if isinstance(s, str) else s for s in seqs]
        ascii_, offsets = sequtils.concat_reads(seqs)
        packed, seg, nbad = sequtils.pack_reads(ascii_, offsets)
        self.addReads(packed, seg)
        return nbad

    def addFastq(self, path, batch_reads=1 << 16):
        seqs = sequtils.read_fastq(path)
        for i in range(0, len(seqs), batch_reads):
            self.addSequences(seqs[i:i + batch_reads])

    def addReadsDevice(self, d_packed, d_offsets, n_reads, n_bases):
        _lib.check(self._lib.tsxc_add_reads_device(self._h, d_packed, d_offsets, n_reads, n_bases), self._h)

    def addKmersDevice(self, d_kmers, n):
        _lib.check(self._lib.tsxc_add_kmers_device(self._h, d_kmers, n), self._h)

    def routeInfo(self):
        info = _lib.TsxcRouteInfo()
        _lib.check(self._lib.tsxc_route_info(self._h, C.byref(info)), self._h)
        return info

    def routeRecvBuffer(self, cap_keys=0):
        """-> (device pointer, capacity in k-mers) of this shard's receive buffer"""
        p, cap = C.c_void_p(), C.c_uint64(0)
        _lib.check(self._lib.tsxc_route_recv_buffer(self._h, cap_keys, C.byref(p), C.byref(cap)), self._h)
        return p.value, cap.value

    def routeSetPeers(self, peer_ptrs, recv_cap_keys):
        arr = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
        _lib.check(self._lib.tsxc_route_set_peers(self._h, arr, recv_cap_keys), self._h)

    def routeBegin(self, d_packed, d_offsets, n_reads, n_bases):
        rounds = C.c_uint32(0)
        _lib.check(self._lib.tsxc_route_begin(self._h, d_packed, d_offsets, n_reads, n_bases, C.byref(rounds)), self._h)
        return rounds.value

    def routeHist(self, rnd, d_hist):
        _lib.check(self._lib.tsxc_route_hist(self._h, rnd, d_hist), self._h)

    def routeSend(self, rnd, d_hist_all):
        _lib.check(self._lib.tsxc_route_send(self._h, rnd, d_hist_all), self._h)

    def routeInsert(self):
        _lib.check(self._lib.tsxc_route_insert(self._h), self._h)

    # -- query path -------------------------------------------------------------------------------
    def getKmerCount(self, kmer=None):
        """No argument: uint64_t TSXHashMap::getKmerCount() (distinct k-mers, TSXHashMap.h:645-648).
        With a k-mer: UBigInt TSXHashMap::getKmerCount(UBigInt&) (TSXHashMap.h:548-638)."""
        if kmer is None:
            out = C.c_uint64(0)
            _lib.check(self._lib.tsxc_distinct(self._h, C.byref(out)), self._h)
            return out.value
        arr = sequtils.from_sequence(kmer) if isinstance(kmer, str) else np.asarray(kmer, dtype=np.uint64)
        return int(self.getKmerCounts(arr.reshape(1, self.kw))[0])

    def getKmerCounts(self, kmers):
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.kw)
        out = np.zeros(kmers.shape[0], dtype=np.uint64)
        _lib.check(self._lib.tsxc_lookup(self._h, kmers.ctypes.data, kmers.shape[0], out.ctypes.data), self._h)
        return out

    def getAllKmers(self):
        """std::vector<UBigInt> TSXHashMap::getAllKmers() (TSXHashMap.h:660-722), with counts."""
        n = self.getKmerCount()
        keys = np.zeros((max(n, 1), self.kw), dtype=np.uint64)
        counts = np.zeros(max(n, 1), dtype=np.uint64)
        got = C.c_uint64(0)
        _lib.check(self._lib.tsxc_dump(self._h, keys.ctypes.data, counts.ctypes.data, max(n, 1), C.byref(got)), self._h)
        return keys[: got.value], counts[: got.value]

    def dump(self, path):
        """KMER<TAB>COUNT lines, the format of count_kmers.py:32-34."""
        _lib.check(self._lib.tsxc_dump_file(self._h, str(path).encode()), self._h)

    def histogram(self, n_bins=256):
        """hist[c] = number of distinct k-mers with count c (c < n_bins-1); hist[n_bins-1] = all with a larger count."""
        out = np.zeros(n_bins, dtype=np.uint64)
        _lib.check(self._lib.tsxc_histogram(self._h, out.ctypes.data, n_bins), self._h)
        return out

    def stats(self):
        st = _lib.TsxcStats()
        _lib.check(self._lib.tsxc_stats(self._h, C.byref(st)), self._h)
        return st.as_dict()

    def print_stats(self, file=None):
        """void TSXHashMap::print_stats() — TSXHashMap.h:390-395 (same three lines)."""
        import sys
        f = file or sys.stderr
        st = self.stats()
        print(f"Used fields: {st['used_slots']}", file=f)
        print(f"Available fields: {float(st['n_slots']):g}", file=f)
        print(f"k={self.k} l={self.l} entry (key+value) bits={64 * st['entry_words']} storage bits={st['value_bits']}", file=f)
