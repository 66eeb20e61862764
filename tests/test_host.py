"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the bijective
hash round-trips (TSXHashMap::testHashFunction, TSXHashMap.h:724-735 of the reference), the entry layout
selection, the 2-bit packer and the FASTQ reader.  No compute entry point is called without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_py as orc
import tsxcount_b200 as tsx
from tsxcount_b200 import _lib, sequtils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "tsxcount_cuda.h")).read()
    declared = set(re.findall(r"\b(tsxc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tsxc_table", "tsxc_stats_t", "tsxc_gen_params"}
    lib = _lib.load()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert lib.tsxc_abi_version() == 2


def test_no_cpu_fallback():
    lib = _lib.load()
    if lib.tsxc_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(tsx.TsxcError) as e:
        tsx.TSXHashMapCUDA(20, 4, 14)
    assert e.value.status == _lib.TSXC_E_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tsxcount_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/oracle.c has", ""), f


@pytest.mark.parametrize("k", [1, 2, 5, 14, 16, 17, 31, 32, 33, 48, 63, 64, 65, 90, 96, 97, 127, 128])
def test_hash_is_a_bijection_on_2k_bits(k):
    lib = _lib.load()
    kw = sequtils.key_words(k)
    rng = np.random.default_rng(k)
    nbits = 2 * k
    seen = set()
    for trial in range(300):
        v = int.from_bytes(rng.bytes(32), "little") & ((1 << nbits) - 1)
        if trial == 0:
            v = 0
        if trial == 1:
            v = (1 << nbits) - 1
        key = np.array([(v >> (64 * j)) & (2**64 - 1) for j in range(kw)], dtype=np.uint64)
        h = np.zeros(kw, dtype=np.uint64)
        back = np.zeros(kw, dtype=np.uint64)
        assert lib.tsxc_debug_hash(k, key.ctypes.data, h.ctypes.data) == 0
        assert lib.tsxc_debug_unhash(k, h.ctypes.data, back.ctypes.data) == 0
        assert np.array_equal(back, key)
        hv = sum(int(x) << (64 * j) for j, x in enumerate(h.tolist()))
        assert hv < (1 << nbits), "hash leaves the 2k-bit domain"
        seen.add(hv)
    if nbits >= 20:
        assert len(seen) == 300
    if nbits <= 12:  # exhaustive: a permutation
        imgs = set()
        for v in range(1 << nbits):
            key = np.array([v], dtype=np.uint64)
            h = np.zeros(1, dtype=np.uint64)
            lib.tsxc_debug_hash(k, key.ctypes.data, h.ctypes.data)
            imgs.add(int(h[0]))
        assert imgs == set(range(1 << nbits))


def test_hash_low_bits_avalanche():
    """Bucket indices come from the low bits of hash word 0: single-base changes anywhere in the k-mer
    must flip them about half the time (the reference's triangular matrix fails this)."""
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for k in (31, 63, 127):
        kw = sequtils.key_words(k)
        flips = []
        for _ in range(200):
            v = int.from_bytes(rng.bytes(32), "little") & ((1 << (2 * k)) - 1)
            pos = int(rng.integers(0, 2 * k))
            outs = []
            for x in (v, v ^ (1 << pos)):
                key = np.array([(x >> (64 * j)) & (2**64 - 1) for j in range(kw)], dtype=np.uint64)
                h = np.zeros(kw, dtype=np.uint64)
                lib.tsxc_debug_hash(k, key.ctypes.data, h.ctypes.data)
                outs.append(int(h[0]) & 0xFFFFFFFF)
            flips.append(bin(outs[0] ^ outs[1]).count("1"))
        assert 13 < np.mean(flips) < 19, (k, np.mean(flips))


def test_layout_selection():
    lib = _lib.load()

    def layout(k, l, s, flags=0, shards=1):
        st = _lib.TsxcStats()
        rc = lib.tsxc_debug_layout(k, l, s, flags, shards, C.byref(st))
        return rc, st

    rc, st = layout(31, 34, 0)                      # config 2: 8-byte entries, 4 per 32-byte bucket
    assert rc == 0 and (st.key_words, st.entry_words, st.slots_per_bucket) == (1, 1, 4)
    assert st.table_bytes == (1 << 34) * 8 and st.quotient_bits == 62 - 32 and st.value_bits >= 20
    rc, st = layout(14, 26, 4, _lib.TSXC_FLAG_EXACT_S)   # config 1 defaults (main.cpp:409-413)
    assert rc == 0 and st.value_bits == 4 and st.entry_words == 1
    rc, st = layout(63, 33, 0)                      # config 3: 16-byte entries
    assert rc == 0 and (st.key_words, st.entry_words, st.slots_per_bucket) == (2, 2, 2)
    rc, st = layout(127, 32, 0)                     # config 4: 32-byte entries, one per sector
    assert rc == 0 and (st.key_words, st.entry_words, st.slots_per_bucket) == (4, 4, 1)
    assert st.table_bytes == (1 << 32) * 32
    rc, st = layout(31, 37, 0, 0, 8)                # config 5: 8 shards of 2^34 slots
    assert rc == 0 and st.n_slots == 1 << 34 and st.n_shards == 8
    assert layout(14, 28, 4)[0] == _lib.TSXC_E_INVALID      # 2k <= l (TSXHashMap.h:91-94)
    assert layout(129, 20, 4)[0] == _lib.TSXC_E_INVALID
    assert layout(31, 20, 4, 0, 3)[0] == _lib.TSXC_E_UNSUPPORTED  # shard count must be a power of two


def test_pack_reads_matches_oracle_encoding():
    seqs = orc.gen_reads(seed=9, n_reads=50, read_len=77, mode=0) + [b"", b"ACGT", b"A"]
    ascii_, off = sequtils.concat_reads(seqs)
    packed, seg, nbad = sequtils.pack_reads(ascii_, off)
    assert nbad == 0 and np.array_equal(seg, off)
    total = int(off[-1])
    v = 0
    for j, w in enumerate(packed.tolist()):
        v |= w << (64 * j)
    text = b"".join(seqs)
    assert all("ACGT"[(v >> (2 * i)) & 3] == chr(text[i]) for i in range(total))
    assert v >> (2 * total) == 0


def test_pack_reads_splits_at_non_acgt():
    ascii_, off = sequtils.concat_reads([b"ACGTNACG", b"NN", b"acgT", b"GGGG"])
    packed, seg, nbad = sequtils.pack_reads(ascii_, off)
    assert nbad == 1 + 2 + 3
    lens = np.diff(seg).tolist()
    assert [x for x in lens if x] == [4, 3, 1, 4]
    assert sequtils.to_sequence(packed[:1], 12) == "ACGTACGTGGGG"


def test_from_to_sequence_roundtrip():
    for s in ("A", "ACGT", "T" * 32, "ACGTTGCA" * 8 + "A", "G" * 128):
        assert sequtils.to_sequence(sequtils.from_sequence(s), len(s)) == s
    assert int(sequtils.from_sequence("CAAA")[0]) == 1      # first base = least significant digit
    with pytest.raises(ValueError):
        sequtils.from_sequence("ACGN")


def test_read_fastq_reference_semantics(tmp_path):
    # FastXReader.h:363-372: empty lines are skipped anywhere; records are groups of 4 non-empty lines
    p = tmp_path / "x.fastq"
    p.write_bytes(b"@r1\nACGT\n+\n&&&&\n\n@r2\n\nGGCC\n+\n&&&&\n@r3\nTT\n+\n")
    assert sequtils.read_fastq(p) == [b"ACGT", b"GGCC"]
    import gzip
    gz = tmp_path / "y.fastq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(b"@r1\nACGT\n+\n&&&&\n")
    assert sequtils.read_fastq(gz) == [b"ACGT"]


def test_pack_reads_randomized_against_numpy_packer():
    """Independent check of the SWAR / AVX2 packer: random read lengths (every alignment of the 8- and 32-base fast
    paths), non-ACGT bytes at random places, against a straightforward numpy packer with the same N policy."""
    rng = np.random.default_rng(77)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    for trial in range(30):
        n = int(rng.integers(1, 60))
        seqs = []
        for _ in range(n):
            s = alphabet[rng.integers(0, 4, size=int(rng.integers(0, 300)))].copy()
            if len(s) and rng.random() < 0.4:
                bad = rng.integers(0, len(s), size=int(rng.integers(1, 4)))
                s[bad] = rng.choice(np.frombuffer(b"Nacgt@[`BDH", dtype=np.uint8), size=len(bad))
            seqs.append(s.tobytes())
        ascii_, off = sequtils.concat_reads(seqs)
        packed, seg, nbad = sequtils.pack_reads(ascii_, off)
        # reference packer: drop the bad bytes, remember where segments end
        code = np.full(256, 255, dtype=np.uint8)
        code[[65, 67, 71, 84]] = [0, 1, 2, 3]
        c = code[ascii_]
        good = c != 255
        assert nbad == int((~good).sum())
        codes = c[good].astype(np.uint64)
        total = len(codes)
        want = np.zeros(total // 32 + 2, dtype=np.uint64)
        idx = np.arange(total)
        np.bitwise_or.at(want, idx // 32, codes << (2 * (idx % 32)).astype(np.uint64))
        assert int(seg[-1]) == total
        assert np.array_equal(packed[: (total + 31) // 32], want[: (total + 31) // 32])
        # segment boundaries: after every read end and at every bad byte, counted in kept bases; empty segments
        # only come from reads (a bad byte never opens an empty segment)
        kept_before = np.concatenate([[0], np.cumsum(good)])
        bounds = [0]
        for r in range(len(seqs)):
            b, e = int(off[r]), int(off[r + 1])
            start = int(kept_before[b])
            for i in range(b, e):
                if not good[i]:
                    g = int(kept_before[i])
                    if g > start:
                        bounds.append(g)
                    start = g
            bounds.append(int(kept_before[e]))
        assert seg.tolist() == bounds


def test_reference_tree_binding_compiles_against_the_reference_headers():
    """INTEGRATION.md §2 is code, not prose: tsxcount_b200/host/ref_binding/TSXHashMapCUDA_ref.h derives from the
    reference's TSXHashMap (src/tsxcount/TSXHashMap.h:68) and overrides addKmer / getKmerCount; oracle/Makefile
    compiles it against the reference's own headers where the tree is mounted (this container).  No GPU needed."""
    import subprocess
    if not os.path.exists("/root/reference/src/tsxcount/TSXHashMap.h"):
        pytest.skip("reference tree not mounted")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_adapter_check")
    assert os.path.exists(exe)
    out = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, check=True).stdout
    assert "ref-adapter built" in out


# ---- canonical k-mers (opt-in extension) ----------------------------------------------------------
def _revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


@pytest.mark.parametrize("k", [1, 2, 5, 14, 31, 32, 33, 48, 63, 64, 65, 96, 127, 128])
def test_canonical_key_is_the_lexicographic_minimum_of_kmer_and_reverse_complement(k):
    """tsxc_debug_canonical runs the host build of the very functions the kernels use (tsx_hash.cuh: canonical_key):
    compared with the definition on TEXT for random k-mers, palindromes and the extremes."""
    lib = tsx._lib.load()
    rng = np.random.default_rng(k)
    kw = lib.tsxc_key_words(k)
    texts = ["".join(rng.choice(list("ACGT"), size=k)) for _ in range(200)] + ["A" * k, "T" * k, "C" * k, "G" * k]
    if k % 2 == 0:
        half = "".join(rng.choice(list("ACGT"), size=k // 2))
        texts.append(half + _revcomp(half))                      # its own reverse complement
    for s in texts:
        want = min(s, _revcomp(s))
        key = np.zeros(4, dtype=np.uint64)
        key[:kw] = sequtils.from_sequence(s)[:kw]
        out = np.zeros(4, dtype=np.uint64)
        assert lib.tsxc_debug_canonical(k, key.ctypes.data, out.ctypes.data) == 0
        assert sequtils.to_sequence(out[:kw], k) == want, (s, want)


@pytest.mark.parametrize("k,read_lens", [
    (127, [150] * 300), (128, [128, 127, 129, 150, 400, 1, 0, 128] * 20), (63, [70, 64, 63, 62, 150] * 40),
    (31, [40, 31, 30, 33, 150, 2] * 60), (33, [50] * 300), (5, [4, 5, 6, 1000, 3] * 30), (97, [100, 96, 97, 4000, 98] * 12),
])
def test_sparse_walk_of_the_partition_pass_enumerates_exactly_the_kmers(k, read_lens):
    """S1's walk over valid k-mer starts (tsx_radix.cuh: valid_starts / locate_word / select_bit / kmer_from_stream32),
    run on the host through tsxc_debug_sparse_round, yields exactly the forward k-mers of the reads
    (testExecution.h:15-36), round by round, for reads around k and stream ends inside a round."""
    lib = _lib.load()
    rng = np.random.default_rng(k * 1000 + len(read_lens))
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)) for n in read_lens]
    seqs = [s for s in seqs if len(s)]
    ascii_, offsets = sequtils.concat_reads(seqs)
    packed, seg, nbad = sequtils.pack_reads(ascii_, offsets)
    assert nbad == 0
    n_bases = int(seg[-1])
    n_words = (n_bases + 31) // 32
    ends = np.zeros(n_words + 1, dtype=np.uint32)
    for e in seg[1:].tolist():
        if e > 0:
            ends[(e - 1) >> 5] |= np.uint32(1 << ((e - 1) & 31))
    kw = sequtils.key_words(k)
    got = []
    out = np.zeros(16384 * kw, dtype=np.uint64)
    n_out = C.c_uint32(0)
    for rnd in range(0, n_words, 512):
        assert lib.tsxc_debug_sparse_round(k, packed.ctypes.data, ends.ctypes.data, n_words, n_bases, rnd, n_words,
                                           out.ctypes.data, C.byref(n_out)) == 0
        got.append(out[: n_out.value * kw].reshape(-1, kw).copy())
    got = np.concatenate(got) if got else np.zeros((0, kw), dtype=np.uint64)
    want = [s[i:i + k].decode() for s in seqs for i in range(len(s) - k + 1)]
    assert len(got) == len(want)
    want_arr = sequtils.kmers_to_array(want, k) if want else np.zeros((0, kw), dtype=np.uint64)
    # the walk lists the k-mers in stream order, which is the order of the reads
    assert np.array_equal(got, want_arr)
    # a segment that ends inside the round: words at and after w_end start nothing
    if n_words > 3:
        assert lib.tsxc_debug_sparse_round(k, packed.ctypes.data, ends.ctypes.data, n_words, n_bases, 0, 3,
                                           out.ctypes.data, C.byref(n_out)) == 0
        starts = [int(offsets[r]) + i for r, s in enumerate(seqs) for i in range(len(s) - k + 1)]
        assert n_out.value == sum(1 for g in starts if g < 96)


def test_sparse_walk_randomized_k_and_read_lengths():
    """The same check as above over 40 random (k, read-length distribution) pairs, including streams that end in the
    first word of a round and reads shorter than k only."""
    lib = _lib.load()
    rng = np.random.default_rng(20261019)
    for trial in range(40):
        k = int(rng.integers(1, 129))
        lo = int(rng.integers(0, k + 10))
        hi = lo + int(rng.integers(1, 3 * k + 40))
        n_reads = int(rng.integers(1, 400))
        seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(n))) for n in rng.integers(lo, hi, size=n_reads)]
        seqs = [s for s in seqs if len(s)]
        if not seqs:
            continue
        ascii_, offsets = sequtils.concat_reads(seqs)
        packed, seg, nbad = sequtils.pack_reads(ascii_, offsets)
        n_bases = int(seg[-1])
        n_words = (n_bases + 31) // 32
        ends = np.zeros(n_words + 1, dtype=np.uint32)
        for e in seg[1:].tolist():
            ends[(e - 1) >> 5] |= np.uint32(1 << ((e - 1) & 31))
        kw = sequtils.key_words(k)
        out = np.zeros(16384 * kw, dtype=np.uint64)
        n_out = C.c_uint32(0)
        got = []
        for rnd in range(0, n_words, 512):
            assert lib.tsxc_debug_sparse_round(k, packed.ctypes.data, ends.ctypes.data, n_words, n_bases, rnd, n_words,
                                               out.ctypes.data, C.byref(n_out)) == 0
            got.append(out[: n_out.value * kw].reshape(-1, kw).copy())
        got = np.concatenate(got)
        want = [s[i:i + k].decode() for s in seqs for i in range(len(s) - k + 1)]
        assert len(got) == len(want), (trial, k, lo, hi)
        if want:
            assert np.array_equal(got, sequtils.kmers_to_array(want, k)), (trial, k, lo, hi)
