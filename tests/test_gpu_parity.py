"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit-exact.

Mirrors the reference's own checks (paths relative to mjoppich/tsxCount):
  --check            src/mains/main.cpp:224-396   every listed k-mer has the listed count, the number of
                                                  distinct k-mers matches and no extra k-mer exists
  testHashMapOld     src/mains/testExecution.h:363-497   4 k-mers added N, N/2, N/2, N/4 times
  README example     README.md:13-24
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_py as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tsx():
    import tsxcount_b200 as m
    assert m._lib.load().tsxc_device_count() >= 1, "no sm_100 device visible"
    return m


def check_against_oracle(tsx, hm, oc):
    """The reference's --check, strengthened: dump == oracle as a map, lookups agree, distinct agrees."""
    kw = hm.kw
    assert hm.getKmerCount() == oc.n_distinct
    st = hm.stats()
    assert st["kmers_added"] == oc.n_total
    assert st["error_flags"] == 0
    keys, counts = hm.getAllKmers()
    got = {tuple(k): int(c) for k, c in zip(keys.tolist(), counts.tolist())}
    assert len(got) == len(keys), "dump lists a k-mer twice"
    want = oc.as_dict(kw)
    assert got == want
    if oc.n_distinct:
        looked = hm.getKmerCounts(oc.keys_kw(kw))
        assert np.array_equal(looked, oc.counts)


def run_case(tsx, seqs, k, l, s, flags=0):
    oc = orc.count_seqs(seqs, k)
    with tsx.TSXHashMapCUDA(l, s, k, flags=flags) as hm:
        hm.addSequences(seqs)
        check_against_oracle(tsx, hm, oc)
        return hm.stats(), oc


# ---- config 1: the bundled example ----------------------------------------------------------------
def test_c1_bundled_fastq_matches_golden_count_file(tsx, tmp_path):
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    golden = orc.golden_path("c1_bundled_k14.fastq.14.count", tmp_path)
    want = {}
    with open(golden) as f:
        for line in f:
            kmer, cnt = line.rstrip("\n").split("\t")
            want[kmer] = int(cnt)
    assert len(want) == 194697 and sum(want.values()) == 202204
    # reference defaults: k=14, l=26, s=4 (main.cpp:409-413); exact s exercises the overflow entries
    with tsx.TSXHashMapCUDA(26, 4, 14, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        hm.addFastq(fastq)
        assert hm.getKmerCount() == 194697
        st = hm.stats()
        assert st["value_bits"] == 4 and st["overflow_entries"] == 52  # 52 k-mers have count > 15
        out = tmp_path / "dump.count"
        hm.dump(out)
        got = {}
        with open(out) as f:
            for line in f:
                kmer, cnt = line.rstrip("\n").split("\t")
                assert kmer not in got
                got[kmer] = int(cnt)
        assert got == want
        arr = tsx.sequtils.kmers_to_array(list(want.keys()), 14)
        looked = hm.getKmerCounts(arr)
        assert looked.tolist() == list(want.values())
        # absent k-mers answer 0 (getKmerCount stops at the first empty slot, TSXHashMap.h:548-638)
        absent = [k for k in ("ACGTACGTACGTAC", "TTTTTTTTTTTTTT", "GGGGGGGGGGGGGA") if k not in want]
        assert all(hm.getKmerCount(k) == 0 for k in absent)


def test_readme_example(tsx):
    # README.md:13-24: ATCGAGTCAGTA, k=5 -> 8 k-mers
    seq = b"ATCGAGTCAGTA"
    st, oc = run_case(tsx, [seq], 5, 8, 4)
    assert oc.n_total == 8


# ---- synthetic configs at oracle-sized inputs -----------------------------------------------------
CASES = [
    # name, gen kwargs, n_reads, read_len, k, l, s, flags
    ("c2_uniform_k31", dict(mode=0, seed=0xC2), 3000, 150, 31, 20, 0, 0),
    ("c2_fakeseq_k31_exact_s4", dict(mode=1, seed=0xC2), 3000, 150, 31, 20, 4, 1),
    ("c2_fakeseq_k31_wide", dict(mode=1, seed=0xC2), 3000, 150, 31, 20, 0, 0),
    ("k32_uniform", dict(mode=0, seed=7), 2000, 150, 32, 19, 4, 0),
    ("k32_small_table_class12", dict(mode=0, seed=8), 300, 150, 32, 16, 12, 1),
    ("k33_uniform", dict(mode=0, seed=9), 2000, 150, 33, 19, 4, 0),
    ("c3_zipf_k63_exact_s4", dict(mode=2, seed=0xC3, genome_len=1 << 8, sub_rate_q16=655), 4000, 150, 63, 19, 4, 1),
    ("c3_zipf_k63_wide", dict(mode=2, seed=0xC3, genome_len=1 << 8, sub_rate_q16=655), 4000, 150, 63, 19, 0, 0),
    ("k64_uniform", dict(mode=0, seed=10), 2000, 150, 64, 18, 4, 0),
    ("k64_tiny_table_class24", dict(mode=0, seed=11), 40, 150, 64, 12, 4, 0),
    ("k65_uniform", dict(mode=0, seed=12), 2000, 150, 65, 18, 4, 0),
    ("k96_fakeseq", dict(mode=1, seed=13), 2000, 150, 96, 18, 4, 1),
    ("c4_uniform_k127", dict(mode=0, seed=0xC4), 4000, 150, 127, 18, 0, 0),
    ("c4_zipf_k127_exact_s4", dict(mode=2, seed=0xC4, genome_len=1 << 6, sub_rate_q16=300), 3000, 150, 127, 18, 4, 1),
    ("k128_fakeseq", dict(mode=1, seed=14), 1500, 150, 128, 17, 4, 0),
    ("c5_genome_k31", dict(mode=3, seed=0xC5, genome_len=50000, sub_rate_q16=328), 8000, 150, 31, 21, 4, 0),
    ("k14_genome_exact_s2", dict(mode=3, seed=15, genome_len=3000, sub_rate_q16=100), 4000, 100, 14, 16, 2, 1),
    ("k5_tiny", dict(mode=3, seed=16, genome_len=150), 500, 40, 5, 9, 3, 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_synthetic_parity(tsx, case):
    name, gen, n_reads, read_len, k, l, s, flags = case
    seqs = orc.gen_reads(n_reads=n_reads, read_len=read_len, **gen)
    run_case(tsx, seqs, k, l, s, flags)


def test_all_layout_classes_are_exercised(tsx):
    """(KW, W) classes (1,1) (1,2) (2,2) (2,4) (4,4) are all reached by CASES."""
    lib = tsx._lib.load()
    seen = set()
    for name, gen, n_reads, read_len, k, l, s, flags in CASES:
        st = tsx._lib.TsxcStats()
        assert lib.tsxc_debug_layout(k, l, s, flags, 1, C.byref(st)) == 0, name
        seen.add((st.key_words, st.entry_words))
    assert seen == {(1, 1), (1, 2), (2, 2), (2, 4), (4, 4)}


# ---- the reference's dormant known-answer test ------------------------------------------------------
@pytest.mark.parametrize("k,l", [(14, 20), (31, 20), (63, 18), (127, 16)])
def test_known_answer_four_kmers_two_overflow_levels(tsx, k, l):
    # testExecution.h:406-409,423,493-496: values 10067, 2786, 9816, 156 added N, N/2, N/2, N/4 times,
    # N = 2048*4*24, with s = 4: counts far beyond one overflow level of the reference
    N = 2048 * 4 * 24
    vals, reps = [10067, 2786, 9816, 156], [N, N // 2, N // 2, N // 4]
    with tsx.TSXHashMapCUDA(l, 4, k, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        kw = hm.kw
        keys = np.zeros((4, kw), dtype=np.uint64)
        keys[:, 0] = vals
        rng = np.random.default_rng(1)
        stream = np.repeat(np.arange(4), reps)
        rng.shuffle(stream)
        for part in np.array_split(stream, 7):          # several launches, interleaved k-mers
            hm.addKmers(keys[part])
        assert hm.getKmerCounts(keys).tolist() == reps
        assert hm.getKmerCount() == 4
        st = hm.stats()
        assert st["overflow_entries"] == 4 and st["kmers_added"] == sum(reps)
        k2, c2 = hm.getAllKmers()
        assert sorted(zip(k2[:, 0].tolist(), c2.tolist())) == sorted(zip(vals, reps))


# ---- edge cases -------------------------------------------------------------------------------------
def test_empty_and_short_reads(tsx):
    k = 21
    seqs = [b"", b"ACGT", b"A" * 20, b"C" * 21, b"", b"ACGTACGTACGTACGTACGTACGTA", b"G"]
    st, oc = run_case(tsx, seqs, k, 12, 4)
    assert oc.n_total == 1 + 5
    with tsx.TSXHashMapCUDA(12, 4, k) as hm:               # nothing at all
        hm.addSequences([])
        hm.addSequences([b"", b""])
        assert hm.getKmerCount() == 0 and hm.getAllKmers()[0].shape[0] == 0


def test_ragged_long_reads_cross_word_boundaries(tsx):
    rng = np.random.default_rng(3)
    seqs = []
    for ln in [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 1000, 19751, 5, 14, 15]:
        seqs.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=ln)))
    for k in (14, 32, 33, 64, 65, 128):
        run_case(tsx, seqs, k, 17, 4)


def test_homopolymers_and_heavy_hitter_overflow(tsx):
    # one slot receives every increment: run aggregation + warp aggregation + carry into the overflow entry
    seqs = [b"A" * 5000, b"A" * 777, b"C" * 3000, b"ACGT" * 500, b"T" * 64]
    for k, s, flags in ((31, 4, 1), (31, 0, 0), (63, 3, 1), (127, 2, 1)):
        st, oc = run_case(tsx, seqs, k, 14, s, flags)
        assert st["overflow_entries"] >= (1 if flags else 0)


def test_non_acgt_bases_split_reads(tsx):
    # N policy: no k-mer spans a non-ACGT byte; lower case is non-ACGT as in the reference
    seqs = [b"ACGTACGTNACGTACGTACGT", b"NNNN", b"ACGTacgtACGTACGTACGT", b"ACGTACGTACGTN"]
    st, oc = run_case(tsx, seqs, 8, 10, 4)
    assert oc.n_skipped > 0 and oc.n_total == (1 + 5) + 0 + (0 + 5) + 5


def test_invalid_and_full(tsx):
    with pytest.raises(tsx.TsxcError) as e:                # TSXHashMap.h:91-94
        tsx.TSXHashMapCUDA(28, 4, 14)
    assert e.value.status == tsx.TSXC_E_INVALID
    with pytest.raises(tsx.TsxcError) as e:
        tsx.TSXHashMapCUDA(20, 4, 129)
    assert e.value.status == tsx.TSXC_E_INVALID
    # 2^6 slots cannot hold 5000 distinct 20-mers: reference exits 42 (TSXHashMap.h:340-343)
    seqs = orc.gen_reads(seed=5, n_reads=50, read_len=120, mode=0)
    with tsx.TSXHashMapCUDA(6, 4, 20) as hm:
        with pytest.raises(tsx.TsxcError) as e:
            hm.addSequences(seqs)
        assert e.value.status == tsx.TSXC_E_TABLE_FULL


def test_high_load_factor(tsx):
    # ~0.9 load: long probe sequences, still exact
    seqs = orc.gen_reads(seed=21, n_reads=1000, read_len=150, mode=0)  # 120k distinct 31-mers
    st, oc = run_case(tsx, seqs, 31, 17, 4)                             # 131072 slots
    assert oc.n_distinct / st["n_slots"] > 0.85 and st["max_reprobe"] > 3


def test_repeated_batches_are_linear(tsx):
    seqs = orc.gen_reads(seed=22, n_reads=500, read_len=150, mode=1)
    oc = orc.count_seqs(seqs, 31)
    with tsx.TSXHashMapCUDA(18, 4, 31, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        for _ in range(3):
            hm.addSequences(seqs)
        assert hm.getKmerCount() == oc.n_distinct
        assert np.array_equal(hm.getKmerCounts(oc.keys_kw(1)), 3 * oc.counts)
        hm.clear()
        assert hm.getKmerCount() == 0 and hm.getKmerCounts(oc.keys_kw(1)).sum() == 0


def test_no_warp_aggregation_flag_same_result(tsx):
    seqs = orc.gen_reads(seed=23, n_reads=800, read_len=150, mode=1)
    run_case(tsx, seqs, 31, 18, 4, flags=tsx.TSXC_FLAG_NO_WARP_AGG | tsx.TSXC_FLAG_EXACT_S)


# ---- device generator == oracle generator ------------------------------------------------------------
@pytest.mark.parametrize("mode,genome,sub", [(0, 0, 0), (1, 0, 0), (2, 1 << 10, 655), (3, 100000, 328), (0, 0, 500)])
def test_device_generator_matches_oracle(tsx, mode, genome, sub):
    lib = tsx._lib.load()
    n_reads, read_len, first, count = 5000, 150, 1234, 777
    p = tsx.TsxcGenParams(0xABC, n_reads, read_len, mode, genome, sub, 0)
    n_words = (count * read_len + 31) // 32
    d_packed, d_off = C.c_void_p(), C.c_void_p()
    tsx._lib.check(lib.tsxc_device_alloc(0, n_words * 8, C.byref(d_packed)))
    tsx._lib.check(lib.tsxc_device_alloc(0, (count + 1) * 8, C.byref(d_off)))
    try:
        tsx._lib.check(lib.tsxc_gen_reads_device(C.byref(p), first, count, 0, None, d_packed, d_off))
        packed = np.zeros(n_words, dtype=np.uint64)
        off = np.zeros(count + 1, dtype=np.uint64)
        tsx._lib.check(lib.tsxc_memcpy(0, packed.ctypes.data, d_packed, n_words * 8, 2))
        tsx._lib.check(lib.tsxc_memcpy(0, off.ctypes.data, d_off, (count + 1) * 8, 2))
    finally:
        lib.tsxc_device_free(0, d_packed)
        lib.tsxc_device_free(0, d_off)
    seqs = orc.gen_reads(seed=0xABC, n_reads=n_reads, read_len=read_len, mode=mode, genome_len=genome,
                         sub_rate_q16=sub, first=first, count=count)
    ascii_, offsets = tsx.sequtils.concat_reads(seqs)
    want_packed, want_off, nbad = tsx.sequtils.pack_reads(ascii_, offsets)
    assert nbad == 0
    assert np.array_equal(off, want_off)
    assert np.array_equal(packed, want_packed[:n_words])


# ---- hash-sharded table on one GPU: the multi-GPU data path with every "rank" in this process ---------------
def _route_all(tsx, shards, d_packed, d_off, n_reads, n_bases, recv_cap_keys=0):
    """The round protocol of tsxcount_b200/multigpu.py with the collectives done by hand: all ranks own the same
    reads here, so every rank sends every k-mer and each k-mer arrives len(shards) times."""
    import torch
    lib = tsx._lib.load()
    G = len(shards)
    bufs = [hm.routeRecvBuffer(recv_cap_keys) for hm in shards]
    cap = min(c for _, c in bufs)
    for hm in shards:
        hm.routeSetPeers([p for p, _ in bufs], cap)       # same device: the "peer" pointers are plain pointers
    bins = shards[0].routeInfo().bins
    rounds = max(hm.routeBegin(d_packed, d_off, n_reads, n_bases) for hm in shards)
    hist = torch.zeros(G * bins, dtype=torch.int32, device="cuda")
    for r in range(rounds):
        for i, hm in enumerate(shards):
            hm.routeHist(r, hist.data_ptr() + 4 * i * bins)
            hm.sync()
        for hm in shards:
            hm.routeSend(r, hist.data_ptr())
        for hm in shards:
            hm.sync()                                      # the barrier: every rank's stores have landed
        for hm in shards:
            hm.routeInsert()
            hm.sync()
    return rounds, cap


@pytest.mark.parametrize("pool", [False, True], ids=["from_receive_buffer", "fine_pass_into_page_pool"])
@pytest.mark.parametrize("k,n_shards,mode,region", [(31, 2, 1, "12"), (31, 8, 0, "12"), (63, 4, 2, "14"), (127, 2, 1, "12"),
                                                    (31, 4, 3, "30")])
@pytest.mark.parametrize("walk", ["", "101"], ids=["walk_by_density", "sparse_walk"])
def test_sharded_route_and_insert(tsx, k, n_shards, mode, region, pool, walk, monkeypatch):
    monkeypatch.setenv("TSXC_REGION_LOG2", region)    # "30": no fine regions at all, routing by owner only
    if walk:
        monkeypatch.setenv("TSXC_SPARSE_PCT", walk)   # the routing pass enumerates valid k-mer starts only, whatever their density
    monkeypatch.setenv("TSXC_SEG_LOG2", "9")          # many planner segments, several rounds
    if pool:
        # the two-level mode of large shards: groups of the receive buffer take a second, local partition pass into a
        # pool of 8-key pages that holds a third of a round (default pages of one slice would not fit these tiny buffers)
        monkeypatch.setenv("TSXC_PAGE_LOG2", "3")
        monkeypatch.setenv("TSXC_POOL_KEYS", "40000")
    lib = tsx._lib.load()
    seqs = orc.gen_reads(seed=31, n_reads=3000, read_len=150, mode=mode,
                         genome_len=(1 << 8) if mode == 2 else (50_000 if mode == 3 else 0), sub_rate_q16=328 if mode == 3 else 0)
    oc = orc.count_seqs(seqs, k)
    ascii_, offsets = tsx.sequtils.concat_reads(seqs)
    packed, seg, _ = tsx.sequtils.pack_reads(ascii_, offsets)
    n_bases = int(seg[-1])
    shards = [tsx.TSXHashMapCUDA(20, 4, k, shard_rank=r, n_shards=n_shards) for r in range(n_shards)]
    kw = shards[0].kw
    bufs = []

    def dalloc(nbytes):
        p = C.c_void_p()
        tsx._lib.check(lib.tsxc_device_alloc(0, nbytes, C.byref(p)))
        bufs.append(p)
        return p

    try:
        d_packed, d_off = dalloc(packed.nbytes + 64), dalloc(seg.nbytes)
        tsx._lib.check(lib.tsxc_memcpy(0, d_packed, packed.ctypes.data, packed.nbytes, 1))
        tsx._lib.check(lib.tsxc_memcpy(0, d_off, seg.ctypes.data, seg.nbytes, 1))
        # receive buffers far smaller than the batch: several rounds
        rounds, cap = _route_all(tsx, shards, d_packed, d_off, len(seg) - 1, n_bases, recv_cap_keys=oc.n_total // 3)
        assert rounds >= 2
        got = {}
        for hm in shards:
            st = hm.stats()
            assert st["error_flags"] == 0
            if region != "30":
                assert st["radix_digit2_bits"] > 0
                assert st["group_cap_keys"] == 8 if pool else st["group_cap_keys"] in (0, 1024, 512, 256)
            keys, counts = hm.getAllKmers()
            for key, c in zip(keys.tolist(), counts.tolist()):
                assert tuple(key) not in got, "k-mer present in two shards"
                got[tuple(key)] = int(c)
        want = {key: n_shards * c for key, c in oc.as_dict(kw).items()}     # every "rank" sent the same reads
        assert got == want
        assert sum(hm.getKmerCount() for hm in shards) == oc.n_distinct
        assert sum(hm.stats()["kmers_added"] for hm in shards) == n_shards * oc.n_total
        # a shard answers 0 for k-mers it does not own; the per-shard answers sum to the oracle's
        total = sum(hm.getKmerCounts(oc.keys_kw(kw)) for hm in shards)
        assert np.array_equal(total, n_shards * oc.counts)
    finally:
        for hm in shards:
            hm.close()
        for p in bufs:
            lib.tsxc_device_free(0, p)


def test_route_receive_overflow_is_reported_not_dropped(tsx, monkeypatch):
    """Every k-mer of every rank belongs to ONE owner (a single repeated read): a round brings that owner more than
    its receive buffer holds.  All ranks see it in the gathered histograms, skip the round together, nothing is
    inserted or written out of bounds, and the sticky status says so."""
    monkeypatch.setenv("TSXC_SEG_LOG2", "9")
    lib = tsx._lib.load()
    seqs = [b"ACGTTGCAAGGCTTAACCGGATATCGCGATTA" * 4] * 3000
    oc = orc.count_seqs(seqs, 63)
    assert oc.n_distinct <= 32
    ascii_, offsets = tsx.sequtils.concat_reads(seqs)
    packed, seg, _ = tsx.sequtils.pack_reads(ascii_, offsets)
    shards = [tsx.TSXHashMapCUDA(16, 4, 63, shard_rank=r, n_shards=4) for r in range(4)]
    d_packed, d_off = C.c_void_p(), C.c_void_p()
    tsx._lib.check(lib.tsxc_device_alloc(0, packed.nbytes + 64, C.byref(d_packed)))
    tsx._lib.check(lib.tsxc_device_alloc(0, seg.nbytes, C.byref(d_off)))
    try:
        tsx._lib.check(lib.tsxc_memcpy(0, d_packed, packed.ctypes.data, packed.nbytes, 1))
        tsx._lib.check(lib.tsxc_memcpy(0, d_off, seg.ctypes.data, seg.nbytes, 1))
        with pytest.raises(tsx.TsxcError) as e:
            _route_all(tsx, shards, d_packed, d_off, len(seg) - 1, int(seg[-1]), recv_cap_keys=oc.n_total // 4)
        assert e.value.status == tsx.TSXC_E_INVALID and "receive buffer" in str(e.value)
    finally:
        for hm in shards:
            hm.close()
        lib.tsxc_device_free(0, d_packed)
        lib.tsxc_device_free(0, d_off)


# ---- the region-sorted insert pipeline (tsx_radix.cuh) on small tables ----------------------------------------
@pytest.fixture(params=["12", "16", "18"], ids=["exact_1024_bins", "exact_128_bins", "paged_32_bins"])
def small_regions(monkeypatch, request):
    """Tables of 512 MiB and more take the pipeline; shrink the table regions so that small tables do too.  8 MiB
    tables: 4 KiB regions = the full 10-bit digit, 64 KiB regions = 7 bits (both with exact bin offsets from the
    histogram pass: with pages of one slice the key buffer is too small for a page pool), 256 KiB regions = 32 bins of
    a paged pool with 32-key pages.  The key buffer holds 200 000 k-mers, so most tests need two or three chunks."""
    monkeypatch.setenv("TSXC_REGION_LOG2", request.param)
    monkeypatch.setenv("TSXC_SEG_LOG2", "9")
    monkeypatch.setenv("TSXC_CHUNK_KEYS", "200000")
    if request.param == "18":
        monkeypatch.setenv("TSXC_PAGE_LOG2", "5")
    monkeypatch.setenv("TSXC_LOOKUP_SORT_MIN", "1000")      # batched lookups of these tests are sorted by region too


PART_CASES = [c for c in CASES if c[0] in (
    "c2_uniform_k31", "c2_fakeseq_k31_exact_s4", "c2_fakeseq_k31_wide", "k32_small_table_class12", "c3_zipf_k63_exact_s4",
    "k64_uniform", "k96_fakeseq", "c4_uniform_k127", "c4_zipf_k127_exact_s4", "c5_genome_k31")]


@pytest.mark.parametrize("case", PART_CASES, ids=[c[0] for c in PART_CASES])
def test_pipeline_parity(tsx, small_regions, case):
    name, gen, n_reads, read_len, k, l, s, flags = case
    seqs = orc.gen_reads(n_reads=n_reads, read_len=read_len, **gen)
    st, oc = run_case(tsx, seqs, k, l, s, flags)
    assert st["main_kernel_launches"] >= 4, "expected the pipeline (histogram, partition, insert)"


@pytest.mark.parametrize("pct", ["0", "101"], ids=["dense_walk", "sparse_walk"])
@pytest.mark.parametrize("case", [PART_CASES[0], PART_CASES[1], PART_CASES[4], PART_CASES[6], PART_CASES[7]], ids=lambda c: c[0])
def test_pipeline_both_walks_of_the_partition_pass(tsx, small_regions, monkeypatch, case, pct):
    """S1 walks every base position (dense) or only the positions that start a k-mer (sparse: validity masks of a block
    round, scan, i-th valid position); the planner's density of a chunk picks one.  Both forced on the same inputs,
    k = 31 / 63 / 96 / 127, in all three geometries; ragged reads around k with the sparse walk."""
    monkeypatch.setenv("TSXC_SPARSE_PCT", pct)
    name, gen, n_reads, read_len, k, l, s, flags = case
    seqs = orc.gen_reads(n_reads=n_reads, read_len=read_len, **gen)
    st, oc = run_case(tsx, seqs, k, l, s, flags)
    assert st["main_kernel_launches"] >= 4
    if pct == "101" and k == 127:
        rng = np.random.default_rng(127)
        ragged = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(n))) for n in rng.integers(100, 200, size=3000)]
        ragged += [b"", b"ACGT" * 31 + b"AC", b"T" * 127, b"G" * 126]
        run_case(tsx, ragged, k, l, s, flags)


@pytest.mark.parametrize("case", [PART_CASES[0], PART_CASES[4], PART_CASES[7]], ids=lambda c: c[0])
def test_pipeline_many_chunks_exact_offsets(tsx, monkeypatch, case):
    """A key buffer far smaller than the batch: the planner cuts it into several chunks (insert passes); segments of
    one block round make chunk boundaries fall inside reads.  Bins get exact offsets from the histogram pass."""
    monkeypatch.setenv("TSXC_REGION_LOG2", "12")
    monkeypatch.setenv("TSXC_SEG_LOG2", "9")
    monkeypatch.setenv("TSXC_CHUNK_KEYS", "70000")
    name, gen, n_reads, read_len, k, l, s, flags = case
    seqs = orc.gen_reads(n_reads=n_reads, read_len=read_len, **gen)
    st, oc = run_case(tsx, seqs, k, l, s, flags)
    assert st["chunk_cap_keys"] == 70000 and st["group_cap_keys"] == 0          # no page pool
    if oc.n_total > 140000:
        assert st["main_kernel_launches"] >= 2 + 3 * 4


@pytest.mark.parametrize("k,l,chunk,page_log2", [(31, 20, 1_000_000, 4), (63, 19, 600_000, 4), (127, 18, 300_000, 3)])
def test_pipeline_paged_bins_many_pages_many_chunks(tsx, monkeypatch, k, l, chunk, page_log2):
    """The page pool under load: 32 bins, pages of 16 / 8 keys (every tile run crosses pages, most runs take several
    fresh pages at once), a pool that the batch fills several times over (several chunks); reads sampled from a
    small genome so that the table holds them."""
    monkeypatch.setenv("TSXC_REGION_LOG2", "18")
    monkeypatch.setenv("TSXC_SEG_LOG2", "9")
    monkeypatch.setenv("TSXC_PAGE_LOG2", str(page_log2))
    monkeypatch.setenv("TSXC_CHUNK_KEYS", str(chunk))
    seqs = orc.gen_reads(seed=0xBEEF + k, n_reads=30000, read_len=150, mode=3, genome_len=40_000, sub_rate_q16=40)
    st, oc = run_case(tsx, seqs, k, l, 0)
    assert st["group_cap_keys"] == 1 << page_log2                               # the page size: the pool is in use
    assert oc.n_total > 2 * chunk
    assert st["main_kernel_launches"] >= 2 + 2 * 5


def test_pipeline_ragged_reads_and_repeats(tsx, small_regions):
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(n))) for n in rng.integers(1, 400, size=1500)]
    seqs += [b"A" * 3000, b"", b"ACGT" * 300]
    oc = orc.count_seqs(seqs, 31)
    with tsx.TSXHashMapCUDA(20, 4, 31, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        hm.addSequences(seqs)
        hm.addSequences(seqs)                        # second batch hits existing entries
        assert hm.stats()["main_kernel_launches"] >= 8
        assert hm.getKmerCount() == oc.n_distinct
        assert np.array_equal(hm.getKmerCounts(oc.keys_kw(1)), 2 * oc.counts)


@pytest.mark.parametrize("k,l", [(31, 20), (63, 19), (127, 18)])
def test_pipeline_heavy_hitters_need_no_hint(tsx, small_regions, k, l):
    """Zipf-like dictionary reads: a few k-mers make up most of the input.  Bins are sized from exact histograms and
    duplicates are combined in shared memory before they reach the table: no creation flag, same counts."""
    seqs = orc.gen_reads(seed=0xC3, n_reads=4000, read_len=150, mode=2, genome_len=1 << 6, sub_rate_q16=655)
    st, oc = run_case(tsx, seqs, k, l, 0)
    assert st["main_kernel_launches"] >= 4
    assert oc.counts.max() > 100


def test_pipeline_one_kmer_is_the_whole_input(tsx, small_regions):
    """24 distinct 12-mers, 7.8e5 k-mers: every tile, every bin and every insert slice holds the same few keys."""
    seqs = [(b"A" * 13 + b"C" * 13) * 200] * 150
    oc = orc.count_seqs(seqs, 12)
    assert oc.n_distinct == 24
    with tsx.TSXHashMapCUDA(20, 4, 12, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        hm.addSequences(seqs)
        assert hm.stats()["main_kernel_launches"] >= 4
        check_against_oracle(tsx, hm, oc)
        hm.addSequences(seqs[:50])                         # and the table stays usable afterwards
        assert hm.getKmerCount() == 24


@pytest.mark.parametrize("k,l,region,acc_words", [(5, 9, "8", "1024"), (14, 20, "12", "2048"), (31, 20, "12", "1024"),
                                                  (63, 20, "14", "4096")])
def test_host_batches_are_accumulated_into_insert_passes(tsx, monkeypatch, k, l, region, acc_words):
    """Many small tsxc_add_reads calls without a sync in between: the library appends them to a device-side stream
    (word aligned; the padding bits between two batches must never yield a k-mer, not even for k smaller than the
    padding) and counts a buffer when it is full, at a sync, or before anything reads the table."""
    monkeypatch.setenv("TSXC_REGION_LOG2", region)
    monkeypatch.setenv("TSXC_ACC_WORDS", acc_words)
    rng = np.random.default_rng(k)
    # k = 5: reads over {A, C} only (32 possible 5-mers fit the 512-slot table; the padding bits read as poly-A, a k-mer
    # that really occurs here, so a leak would change its count)
    alphabet = np.frombuffer(b"AC" if k < 8 else b"ACGT", dtype=np.uint8)
    seqs = [bytes(rng.choice(alphabet, size=int(n))) for n in rng.integers(1, 300, size=3000)]
    oc = orc.count_seqs(seqs, k)
    keep = []
    with tsx.TSXHashMapCUDA(l, 4, k, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
        for i in range(0, len(seqs), 37):                       # batches end at arbitrary bit positions of a word
            ascii_, offsets = tsx.sequtils.concat_reads(seqs[i:i + 37])
            packed, seg, _ = tsx.sequtils.pack_reads(ascii_, offsets)
            keep.append(hm.addReads(packed, seg, sync=False))
        # no explicit sync: the lookup / dump / distinct calls below must see every batch
        check_against_oracle(tsx, hm, oc)
        assert hm.stats()["main_kernel_launches"] >= 4


def test_sorted_lookup_handles_absent_duplicate_and_out_of_range_queries(tsx, small_regions):
    seqs = orc.gen_reads(seed=11, n_reads=3000, read_len=150, mode=3, genome_len=40_000, sub_rate_q16=328)
    oc = orc.count_seqs(seqs, 31)
    other = orc.count_seqs(orc.gen_reads(seed=12, n_reads=50, read_len=150, mode=0), 31)
    with tsx.TSXHashMapCUDA(20, 0, 31) as hm:
        hm.addSequences(seqs)
        q = np.concatenate([oc.keys_kw(1), other.keys_kw(1), oc.keys_kw(1)[:500], np.full((3, 1), 2**63, dtype=np.uint64)])
        want = np.concatenate([oc.counts, np.zeros(other.n_distinct, dtype=np.uint64), oc.counts[:500], np.zeros(3, dtype=np.uint64)])
        perm = np.random.default_rng(0).permutation(len(q))
        got = hm.getKmerCounts(q[perm])
        assert np.array_equal(got, want[perm])
        before = hm.stats()["kernel_launches"]
        hm.getKmerCounts(q[:2000])
        assert hm.stats()["kernel_launches"] - before >= 7      # hash, partition passes, probe


# ---- opt-in extensions the reference lacks: canonical k-mers, count histogram ------------------------
@pytest.mark.parametrize("k,l", [(5, 9), (14, 16), (31, 20), (32, 20), (33, 19), (63, 19), (64, 18), (65, 18), (127, 18)])
def test_canonical_mode_matches_the_oracle(tsx, k, l):
    """TSXC_FLAG_CANONICAL: counts of min(k-mer, reverse complement); the oracle canonicalises the TEXT.  Lookups accept
    either strand; the dump lists canonical k-mers only."""
    seqs = orc.gen_reads(seed=77 + k, n_reads=2500 if k > 20 else 300, read_len=150, mode=3, genome_len=4000 if k > 8 else 300,
                         sub_rate_q16=200)
    seqs += [s.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1] for s in seqs[:400]]      # reverse strands of some reads
    oc = orc.count_seqs(seqs, k, canonical=True)
    fwd = orc.count_seqs(seqs, k)
    assert oc.n_distinct < fwd.n_distinct and oc.n_total == fwd.n_total
    with tsx.TSXHashMapCUDA(l, 0, k, flags=tsx.TSXC_FLAG_CANONICAL) as hm:
        hm.addSequences(seqs)
        check_against_oracle(tsx, hm, oc)
        # every forward k-mer answers with the count of its canonical form
        want = oc.as_dict(hm.kw)
        lib = tsx._lib.load()
        keys = fwd.keys_kw(hm.kw)[:2000]
        canon = np.zeros((len(keys), 4), dtype=np.uint64)
        for i, key in enumerate(keys):
            src = np.zeros(4, dtype=np.uint64); src[:hm.kw] = key
            assert lib.tsxc_debug_canonical(k, src.ctypes.data, canon[i].ctypes.data) == 0
        got = hm.getKmerCounts(keys)
        assert [int(x) for x in got] == [want[tuple(c[:hm.kw].tolist())] for c in canon]


def test_canonical_mode_through_the_pipeline(tsx, small_regions):
    seqs = orc.gen_reads(seed=5, n_reads=3000, read_len=150, mode=3, genome_len=30_000, sub_rate_q16=300)
    seqs += [s.translate(bytes.maketrans(b"ACGT", b"TGCA"))[::-1] for s in seqs[:1000]]
    oc = orc.count_seqs(seqs, 31, canonical=True)
    with tsx.TSXHashMapCUDA(20, 0, 31, flags=tsx.TSXC_FLAG_CANONICAL) as hm:
        hm.addSequences(seqs)
        assert hm.stats()["main_kernel_launches"] >= 4
        check_against_oracle(tsx, hm, oc)


@pytest.mark.parametrize("k,l,s,flags", [(14, 16, 2, 1), (31, 20, 0, 0), (63, 19, 4, 1), (127, 18, 0, 0)])
def test_count_histogram(tsx, k, l, s, flags):
    """hist[c] = number of distinct k-mers with count c; counts beyond the last bin (and beyond the value field: overflow
    entries) land in the last bin."""
    seqs = orc.gen_reads(seed=3 + k, n_reads=3000, read_len=150, mode=2, genome_len=1 << 7, sub_rate_q16=400)
    oc = orc.count_seqs(seqs, k)
    assert oc.counts.max() > 64
    with tsx.TSXHashMapCUDA(l, s, k, flags=flags) as hm:
        hm.addSequences(seqs)
        for n_bins in (2, 17, 64, 4096):
            want = np.bincount(np.minimum(oc.counts, n_bins - 1).astype(np.int64), minlength=n_bins).astype(np.uint64)
            assert np.array_equal(hm.histogram(n_bins), want), n_bins
        with pytest.raises(tsx.TsxcError):
            hm.histogram(5000)


def test_direct_flag_forces_single_kernel(tsx, small_regions):
    seqs = orc.gen_reads(seed=3, n_reads=2000, read_len=150, mode=0)
    st, oc = run_case(tsx, seqs, 31, 20, 0, flags=4)
    assert st["main_kernel_launches"] == 1


# ---- the committed fixtures the reference binary was run on -----------------------------------------------
def test_reference_pinned_fixtures(tsx, tmp_path):
    import json
    pins = json.load(open(os.path.join(orc.GOLDEN, "ref_binary_pins.json")))
    for case in pins["cases"]:
        name, k = case["name"], case["k"]
        fastq = orc.golden_path(f"{name}.fastq", tmp_path)
        count = orc.golden_path(f"{name}.fastq.{k}.count", tmp_path)
        want = {}
        with open(count) as f:
            for line in f:
                kmer, c = line.rstrip("\n").split("\t")
                want[kmer] = int(c)
        with tsx.TSXHashMapCUDA(case["l"], case["s"], k, flags=tsx.TSXC_FLAG_EXACT_S) as hm:
            hm.addFastq(fastq)
            assert hm.getKmerCount() == case["oracle_distinct"] == len(want)
            keys, counts = hm.getAllKmers()
            got = {tsx.sequtils.to_sequence(kk, k): int(c) for kk, c in zip(keys, counts)}
            assert got == want, name


# ---- the default two-phase configuration at a size the oracle cannot count: size-independent properties ----
def test_large_default_path_properties(tsx):
    """2.4e8 uniform 31-mers into a 4 GiB table (32 regions of 128 MiB, pages of 1024 k-mers: the default geometry).
    Properties: every k-mer is added exactly once (sum of counts), all are distinct (a duplicate
    has probability ~1e-2 at this size), sampled k-mers regenerated by the oracle are present with count 1 and
    absent ones with 0; a second pass doubles every sampled count and leaves the distinct count unchanged."""
    lib = tsx._lib.load()
    n_reads, read_len, k = 2_000_000, 150, 31
    n_bases = n_reads * read_len
    n_words = (n_bases + 31) // 32
    n_kmers = n_reads * (read_len - k + 1)
    d_packed, d_off = C.c_void_p(), C.c_void_p()
    tsx._lib.check(lib.tsxc_device_alloc(0, (n_words + 8) * 8, C.byref(d_packed)))
    tsx._lib.check(lib.tsxc_device_alloc(0, (n_reads + 1) * 8, C.byref(d_off)))
    try:
        gp = tsx.TsxcGenParams(0xBEEF, n_reads, read_len, 0, 0, 0, 0)
        tsx._lib.check(lib.tsxc_gen_reads_device(C.byref(gp), 0, n_reads, 0, None, d_packed, d_off))
        sample = orc.gen_reads(seed=0xBEEF, n_reads=n_reads, read_len=read_len, mode=0, first=777_000, count=300)
        oc = orc.count_seqs(sample, k)
        other = orc.count_seqs(orc.gen_reads(seed=0xBEEF + 1, n_reads=10, read_len=read_len, mode=0), k)
        with tsx.TSXHashMapCUDA(29, 0, k) as hm:
            for rep in (1, 2):
                hm.addReadsDevice(d_packed, d_off, n_reads, n_bases)
                hm.sync()
                st = hm.stats()
                assert st["error_flags"] == 0 and st["kmers_added"] == rep * n_kmers
                assert st["distinct"] == n_kmers
                assert st["main_kernel_launches"] >= 4 * rep          # the pipeline ran
                assert st["radix_digit1_bits"] == 5 and st["radix_digit2_bits"] == 0 and st["group_cap_keys"] == 1024
                assert np.array_equal(hm.getKmerCounts(oc.keys_kw(1)), rep * oc.counts)
                assert hm.getKmerCounts(other.keys_kw(1)).sum() == 0
    finally:
        lib.tsxc_device_free(0, d_packed)
        lib.tsxc_device_free(0, d_off)
