"""Host-side sequence helpers: the Python mirror of the reference's TSXSeqUtils / FASTXreader for the
feeder side of the boundary (paths relative to mjoppich/tsxCount):

  from_sequence / to_sequence   src/utils/SequenceUtils.h:86-160 / :47-84
  read_fastq                    src/fastxutils/FastXReader.h:62-95,307-385 (4-line records, empty lines
                                skipped, sequence = 2nd line, ".gz" by file name :185-190)
  pack_reads                    tsxc_pack_reads (C++), see tsxcount_b200/csrc/tsx_host_pack.cpp
"""
import ctypes as C
import gzip

import numpy as np

from . import _lib

_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
_LETTER = "ACGT"


def key_words(k):
    n = _lib.load().tsxc_key_words(k)
    if n == 0:
        raise ValueError(f"k={k} out of range [1,128]")
    return n


def from_sequence(seq):
    """k-mer string -> np.uint64[KW]; base i at bits [2i,2i+1], A=0 C=1 G=2 T=3.  Non-ACGT raises:
    the reference inserts random bits there (SequenceUtils.h:126-137), which has no defined result."""
    kw = key_words(len(seq))
    v = 0
    for i, ch in enumerate(seq):
        try:
            v |= _CODE[ch] << (2 * i)
        except KeyError:
            raise ValueError(f"non-ACGT base {ch!r} at {i}") from None
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(kw)], dtype=np.uint64)


def to_sequence(words, k):
    v = 0
    for j, w in enumerate(np.asarray(words, dtype=np.uint64).tolist()):
        v |= int(w) << (64 * j)
    return "".join(_LETTER[(v >> (2 * i)) & 3] for i in range(k))


def kmers_to_array(kmers, k):
    """list of k-mer strings -> np.uint64[n, KW]"""
    kw = key_words(k)
    out = np.zeros((len(kmers), kw), dtype=np.uint64)
    for i, s in enumerate(kmers):
        if len(s) != k:
            raise ValueError(f"k-mer {s!r} has length {len(s)} != k={k}")
        out[i] = from_sequence(s)
    return out


def read_fastq(path, lines_per_record=4):
    """Sequences of a FASTQ (or FASTA with lines_per_record=2) file, FASTXreader semantics."""
    opener = gzip.open if str(path).endswith(".gz") else open
    seqs, group = [], []
    with opener(path, "rb") as f:
        for raw in f:
            line = raw[:-1] if raw.endswith(b"\n") else raw
            if len(line) == 0:
                continue
            group.append(line)
            if len(group) == lines_per_record:
                seqs.append(group[1])
                group = []
    return seqs


def read_fasta(path, piece_len=1 << 20, overlap=0):
    """Multi-line FASTA (mirror of FastxReader's auto-detected FASTA mode, an extension over the reference): a
    record is a '>' header plus all lines up to the next header; bases are upper-cased; a sequence longer than
    piece_len is cut into pieces overlapping by `overlap` (= k-1) bases, which keeps the k-mer multiset."""
    if piece_len <= overlap:
        raise ValueError("piece length must exceed the overlap")
    opener = gzip.open if str(path).endswith(".gz") else open
    records, cur = [], None
    with opener(path, "rb") as f:
        for raw in f:
            line = raw[:-1] if raw.endswith(b"\n") else raw
            if len(line) == 0:
                continue
            if line[:1] == b">":
                if cur:
                    records.append(b"".join(cur))
                cur = []
            elif cur is not None:
                cur.append(bytes(c - 32 if 97 <= c <= 122 else c for c in line))
            else:                      # sequence lines before the first header still form a record
                cur = [bytes(c - 32 if 97 <= c <= 122 else c for c in line)]
    if cur:
        records.append(b"".join(cur))
    pieces = []
    for s in records:
        start = 0
        while len(s) - start > piece_len:
            pieces.append(s[start:start + piece_len])
            start += piece_len - overlap
        if len(s) - start > (overlap if start else 0):
            pieces.append(s[start:])
    return pieces


def concat_reads(seqs):
    """list of bytes -> (ascii uint8 array, offsets uint64[n+1])"""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.uint64, count=len(seqs))
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    ascii_ = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, dtype=np.uint8)
    return ascii_, offsets


def pack_reads(ascii_, offsets):
    """2-bit pack (tsxc_pack_reads).  Returns (packed uint64[], segment offsets uint64[], n_bad_bases).
    Non-ACGT bytes split a read into segments (no k-mer spans them)."""
    lib = _lib.load()
    ascii_ = np.ascontiguousarray(ascii_, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_reads = len(offsets) - 1
    total = int(offsets[-1]) if n_reads >= 0 else 0
    is_bad = ~np.isin(ascii_[:total], np.frombuffer(b"ACGT", dtype=np.uint8))
    n_bad_upper = int(is_bad.sum())
    packed = np.zeros(total // 32 + 2, dtype=np.uint64)
    seg = np.zeros(n_reads + n_bad_upper + 2, dtype=np.uint64)
    nseg = C.c_uint64(0)
    nbad = C.c_uint64(0)
    _lib.check(lib.tsxc_pack_reads(ascii_.ctypes.data, offsets.ctypes.data, n_reads, packed.ctypes.data,
                                   seg.ctypes.data, len(seg), C.byref(nseg), C.byref(nbad)))
    return packed, seg[: nseg.value + 1].copy(), nbad.value
