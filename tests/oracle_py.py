"""ctypes wrapper of oracle/_build/liboracle.so (the CPU restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by tsxcount_b200/."""
import ctypes as C
import gzip
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")
KEY_WORDS = 4


class OrcCounts(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("n_total", C.c_uint64), ("n_skipped", C.c_uint64), ("k", C.c_uint),
                ("keys", C.POINTER(C.c_uint64)), ("counts", C.POINTER(C.c_uint64)), ("first", C.POINTER(C.c_uint64))]


class OrcGenParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint64), ("read_len", C.c_uint32), ("mode", C.c_uint32),
                ("genome_len", C.c_uint64), ("sub_rate_q16", C.c_uint32), ("reserved", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.orc_count_reads.restype = C.POINTER(OrcCounts)
        L.orc_count_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint]
        L.orc_count_reads_canonical.restype = C.POINTER(OrcCounts)
        L.orc_count_reads_canonical.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint]
        L.orc_count_fastq.restype = C.POINTER(OrcCounts)
        L.orc_count_fastq.argtypes = [C.c_char_p, C.c_uint]
        L.orc_write_dump.restype = C.c_int
        L.orc_write_dump.argtypes = [C.POINTER(OrcCounts), C.c_char_p]
        L.orc_free.argtypes = [C.POINTER(OrcCounts)]
        L.orc_gen_reads.argtypes = [C.POINTER(OrcGenParams), C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_encode_kmer.restype = C.c_int
        L.orc_encode_kmer.argtypes = [C.c_char_p, C.c_uint, C.c_void_p]
        L.orc_decode_kmer.argtypes = [C.c_void_p, C.c_uint, C.c_char_p]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        _lib = L
    return _lib


class Counts:
    """(k-mer -> count) map of the oracle: keys uint64[n,4] (first-occurrence order), counts uint64[n]."""

    def __init__(self, ptr):
        c = ptr.contents
        n = c.n_distinct
        self.k = c.k
        self.n_distinct, self.n_total, self.n_skipped = n, c.n_total, c.n_skipped
        self.keys = np.ctypeslib.as_array(c.keys, shape=(max(n, 1) * KEY_WORDS,))[: n * KEY_WORDS].reshape(n, KEY_WORDS).copy()
        self.counts = np.ctypeslib.as_array(c.counts, shape=(max(n, 1),))[:n].copy()
        lib().orc_free(ptr)

    def keys_kw(self, kw):
        return np.ascontiguousarray(self.keys[:, :kw])

    def as_dict(self, kw):
        return {tuple(k): int(c) for k, c in zip(self.keys[:, :kw].tolist(), self.counts.tolist())}


def count_reads(ascii_, offsets, k, canonical=False):
    ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    fn = lib().orc_count_reads_canonical if canonical else lib().orc_count_reads
    p = fn(ascii_.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k)
    assert p, "oracle failed"
    return Counts(p)


def count_seqs(seqs, k, canonical=False):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    ascii_ = np.frombuffer(b"".join(seqs), dtype=np.uint8) if len(seqs) else np.zeros(0, np.uint8)
    return count_reads(ascii_, offsets, k, canonical)


def count_fastq(path, k):
    p = lib().orc_count_fastq(str(path).encode(), k)
    assert p, f"oracle could not read {path}"
    return Counts(p)


def write_dump_fastq(path, k, out):
    p = lib().orc_count_fastq(str(path).encode(), k)
    assert p
    rc = lib().orc_write_dump(p, str(out).encode())
    lib().orc_free(p)
    assert rc == 0


def gen_reads(seed, n_reads, read_len, mode=0, genome_len=0, sub_rate_q16=0, first=0, count=None):
    """ASCII reads [first, first+count) of the synthetic generator as a list of bytes."""
    count = n_reads - first if count is None else count
    p = OrcGenParams(seed, n_reads, read_len, mode, genome_len, sub_rate_q16, 0)
    buf = np.zeros(count * read_len, dtype=np.uint8)
    lib().orc_gen_reads(C.byref(p), first, count, buf.ctypes.data)
    raw = buf.tobytes()
    return [raw[i * read_len:(i + 1) * read_len] for i in range(count)]


def golden_path(name, tmpdir):
    """Decompress tests/golden/<name>.gz into tmpdir and return the path."""
    dst = os.path.join(str(tmpdir), name)
    if not os.path.exists(dst):
        with gzip.open(os.path.join(GOLDEN, name + ".gz"), "rb") as fi, open(dst, "wb") as fo:
            fo.write(fi.read())
    return dst
