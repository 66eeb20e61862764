"""CPU tests of the oracle itself: it must reproduce every golden vector the reference holds for this
path before anything is compared against it."""
import gzip
import hashlib
import json
import os

import numpy as np

import oracle_py as orc

GOLD = orc.GOLDEN


def test_oracle_reproduces_bundled_count_file_byte_for_byte(tmp_path):
    # data/small_t7.1000.fastq + .14.count of the reference (fixture copies in tests/golden/)
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    out = tmp_path / "o.count"
    orc.write_dump_fastq(fastq, 14, out)
    want = gzip.open(os.path.join(GOLD, "c1_bundled_k14.fastq.14.count.gz"), "rb").read()
    got = open(out, "rb").read()
    assert got == want
    assert got.count(b"\n") == 194697


def test_bundled_numbers():
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        oc = orc.count_fastq(orc.golden_path("c1_bundled_k14.fastq", d), 14)
    assert (oc.n_distinct, oc.n_total, int(oc.counts.max())) == (194697, 202204, 589)
    assert int((oc.counts > 15).sum()) == 52 and int((oc.counts > 255).sum()) == 5


def test_reference_binary_pins_are_green_and_fixtures_unchanged(tmp_path):
    """tests/golden/ref_binary_pins.json is written by oracle/make_ref_pins.py: the UNMODIFIED reference
    binary ran --check against the oracle's dump for each case and reported total errors 0."""
    pins = json.load(open(os.path.join(GOLD, "ref_binary_pins.json")))
    assert len(pins["cases"]) >= 5
    for case in pins["cases"]:
        assert case["pinned"], case["name"]
        for r in case["runs"]:
            if r.get("informational"):
                continue
            assert r["total_errors"] == 0 and r["xor_kmer_count"] == 0
            assert r["reference_kmer_count"] == case["oracle_distinct"] == r["tsxcount_kmer_count"]
        # the committed fixtures are the files the reference binary saw, and the oracle still
        # produces the same dump from them
        name, k = case["name"], case["k"]
        fq = os.path.join(GOLD, f"{name}.fastq.gz")
        if not os.path.exists(fq):
            continue
        fastq = orc.golden_path(f"{name}.fastq", tmp_path)
        assert hashlib.sha256(open(fastq, "rb").read()).hexdigest() == case["fastq_sha256"]
        out = tmp_path / f"{name}.count"
        orc.write_dump_fastq(fastq, k, out)
        assert hashlib.sha256(open(out, "rb").read()).hexdigest() == case["count_sha256"]


def test_fixtureless_reference_pins_at_word_boundary_k(tmp_path):
    """oracle_only_cases of the same file: the unmodified reference binary checked the restatement's dump at k = 28, 32, 34,
    36 (with overflow entries: counts > 15), 48 and 64 (flat).  No fixtures are committed for them; the FASTQ comes from
    the same generator call and must have the recorded hash, and so must the dump the oracle writes today."""
    import subprocess
    pins = json.load(open(os.path.join(GOLD, "ref_binary_pins.json")))
    cases = pins["oracle_only_cases"]
    assert {c["k"] for c in cases} >= {28, 32, 34, 36, 48, 64}
    assert any(c["max_count"] > 15 for c in cases if c["k"] in (28, 32, 34, 36))      # overflow entries were exercised
    tool = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_build", "kmer_oracle")
    for case in cases:
        assert case["pinned"], case["name"]
        for r in case["runs"]:
            if r.get("informational"):
                continue
            assert r["total_errors"] == 0 and r["xor_kmer_count"] == 0
            assert r["reference_kmer_count"] == case["oracle_distinct"] == r["tsxcount_kmer_count"]
        fastq = tmp_path / (case["name"] + ".fastq")
        subprocess.run([tool, "gen"] + [str(x) for x in case["gen"]] + [str(fastq)], check=True)
        assert hashlib.sha256(open(fastq, "rb").read()).hexdigest() == case["fastq_sha256"]
        out = tmp_path / (case["name"] + ".count")
        orc.write_dump_fastq(fastq, case["k"], out)
        assert hashlib.sha256(open(out, "rb").read()).hexdigest() == case["count_sha256"]
        oc = orc.count_fastq(fastq, case["k"])
        assert (oc.n_distinct, oc.n_total, int(oc.counts.max())) == (case["oracle_distinct"], case["oracle_total"], case["max_count"])


def test_readme_example():
    # README.md:13-24 of the reference
    oc = orc.count_seqs([b"ATCGAGTCAGTA"], 5)
    assert oc.n_total == 8 and oc.n_distinct == 8


def test_extraction_semantics():
    # testExecution.h:15-36: len < k -> nothing; len == k -> one k-mer; no k-mer crosses reads
    oc = orc.count_seqs([b"ACG", b"ACGT", b"ACGTA", b""], 4)
    assert oc.n_total == 1 + 2
    d = oc.as_dict(1)
    enc = lambda s: sum("ACGT".index(c) << (2 * i) for i, c in enumerate(s))  # noqa: E731
    assert d == {(enc("ACGT"),): 2, (enc("CGTA"),): 1}


def test_encoding_first_base_is_least_significant():
    # SequenceUtils.h:96-123
    import ctypes as C
    key = (C.c_uint64 * 4)()
    assert orc.lib().orc_encode_kmer(b"CAAA", 4, key) == 0 and key[0] == 1
    assert orc.lib().orc_encode_kmer(b"AAAT", 4, key) == 0 and key[0] == 3 << 6
    assert orc.lib().orc_encode_kmer(b"ACGN", 4, key) == -1
    buf = C.create_string_buffer(5)
    key[0] = 0b11100100
    orc.lib().orc_decode_kmer(key, 4, buf)
    assert buf.value == b"ACGT"


def test_n_policy_skips_spanning_kmers():
    oc = orc.count_seqs([b"ACGTNACGT"], 4)
    assert oc.n_total == 2 and oc.n_skipped == 4


def test_generator_is_deterministic_and_shaped():
    a = orc.gen_reads(seed=1, n_reads=100, read_len=150, mode=1)
    b = orc.gen_reads(seed=1, n_reads=100, read_len=150, mode=1, first=10, count=5)
    assert a[10:15] == b and all(len(s) == 150 and set(s) <= set(b"ACGT") for s in a)
    assert any(b"A" * 25 in s for s in a)          # poly-A tails (generateFakeSequences.py:11-13 style)
    z = orc.gen_reads(seed=2, n_reads=2000, read_len=50, mode=2, genome_len=1 << 10)
    top = max(set(z), key=z.count)
    assert z.count(top) > 100                       # log-uniform ranks: heavy hitters


def test_canonical_counting_against_a_plain_python_counter():
    """orc_count_reads_canonical (string-level reverse complement + memcmp) against collections.Counter over
    min(kmer, revcomp(kmer)) on the README example and on random reads, k on both sides of a word boundary."""
    from collections import Counter
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rng = np.random.default_rng(4)
    seqs = [b"ATCGAGTCAGTA"] + [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(n))) for n in rng.integers(1, 90, size=60)]
    seqs += [b"ACGT" * 10, b"ACGTNACGTTGCA"]
    for k in (3, 5, 31, 33):
        want = Counter()
        for s in seqs:
            for i in range(len(s) - k + 1):
                w = s[i:i + k]
                if b"N" in w:
                    continue
                want[min(w, w.translate(comp)[::-1])] += 1
        oc = orc.count_seqs(seqs, k, canonical=True)
        got = {}
        for key, c in zip(oc.keys, oc.counts):
            text = np.zeros(k + 1, dtype=np.uint8)
            orc.lib().orc_decode_kmer(np.ascontiguousarray(key).ctypes.data, k, text.ctypes.data_as(__import__("ctypes").c_char_p))
            got[bytes(text[:k])] = int(c)
        assert got == dict(want), k
        assert oc.n_total == sum(want.values())
