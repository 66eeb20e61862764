"""The C++ host CLI (tsxcount_b200/host/main.cpp): the reference's option surface plus --mode=CUDA."""
import os
import subprocess

import pytest

import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tsxcount_b200", "bin", "tsxcount")


@pytest.fixture(scope="module", autouse=True)
def _cli_built():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", ROOT, "cli"], check=True, capture_output=True)


def test_cli_keeps_the_reference_option_surface():
    out = subprocess.run([CLI, "--help"], capture_output=True, text=True).stdout
    # src/mains/main.cpp:30-40 of the reference
    for opt in ("--k=", "--s=", "--l=", "--input=", "--check", "--checkabort", "--threads", "--mode"):
        assert opt in out, opt


def test_cli_refuses_cpu_modes_and_has_no_fallback():
    p = subprocess.run([CLI, "--input=x.fastq", "--mode=OMP"], capture_output=True, text=True)
    assert p.returncode == 2 and "--mode=CUDA only" in p.stderr
    import tsxcount_b200 as tsx
    if tsx._lib.load().tsxc_device_count() == 0:
        p = subprocess.run([CLI, "--input=x.fastq", "--mode=CUDA"], capture_output=True, text=True)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr
        # the reference echoes its parameters on these streams (main.cpp:420-427)
        assert "Running with parameters" in p.stdout and "K=14" in p.stderr and "L=26" in p.stderr and "StorageBits=4" in p.stderr


@pytest.mark.gpu
def test_cli_count_check_dump_on_bundled_example(tmp_path):
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    golden = orc.golden_path("c1_bundled_k14.fastq.14.count", tmp_path)
    dump = tmp_path / "out.count"
    # the reference's documented call (README.md:61) with the new mode
    p = subprocess.run([CLI, f"--input={fastq}", "--mode=CUDA", "--check", f"--dump={dump}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Added a total of 194697 different kmers" in p.stdout
    assert "total errors0" in p.stdout
    assert "Reference kmer count: 194697" in p.stdout and "tsxCount kmer count: 194697" in p.stdout
    assert "queried (Xor) kmer count: 0" in p.stdout
    assert "k=14 l=26 entry (key+value) bits=64 storage bits=4" in p.stderr
    assert sorted(open(dump).read().splitlines()) == sorted(open(golden).read().splitlines())


@pytest.mark.gpu
def test_cli_canonical_and_histogram_on_bundled_example(tmp_path):
    """--canonical / --histogram (extensions): the dump equals the oracle's canonical counts of the bundled reads, the
    histogram their count distribution; --check with a large reference file goes through the block parser."""
    import collections
    import tsxcount_b200 as tsx
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    dump, hist = tmp_path / "canon.count", tmp_path / "canon.hist"
    p = subprocess.run([CLI, f"--input={fastq}", "--mode=CUDA", "--canonical", f"--dump={dump}", f"--histogram={hist}", "--threads=4"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    oc = orc.count_seqs(tsx.sequtils.read_fastq(fastq), 14, canonical=True)
    want = {tsx.sequtils.to_sequence(k[:1], 14): int(c) for k, c in zip(oc.keys, oc.counts)}
    got = dict((a, int(b)) for a, b in (line.split("\t") for line in open(dump).read().splitlines()))
    assert got == want
    assert f"Added a total of {oc.n_distinct} different kmers" in p.stdout
    hwant = collections.Counter(min(c, 256) for c in want.values())
    hgot = {(256 if a == "256+" else int(a)): int(b) for a, b in (line.split("\t") for line in open(hist).read().splitlines())}
    assert hgot == dict(hwant)
    # --check against its own canonical dump (194 000 lines, parsed by 4 workers in one block)
    os.replace(dump, str(fastq) + ".14.count")
    p = subprocess.run([CLI, f"--input={fastq}", "--mode=CUDA", "--canonical", "--check", "--threads=4"], capture_output=True, text=True)
    assert p.returncode == 0 and "total errors0" in p.stdout and "queried (Xor) kmer count: 0" in p.stdout, p.stdout[-500:]


@pytest.mark.gpu
def test_cli_check_detects_a_wrong_reference_and_checkabort_exits_200(tmp_path):
    fastq = orc.golden_path("c2_fakeseq_k31.fastq", tmp_path)
    good = orc.golden_path("c2_fakeseq_k31.fastq.31.count", tmp_path)
    lines = open(good).read().splitlines()
    kmer, cnt = lines[5].split("\t")
    lines[5] = f"{kmer}\t{int(cnt) + 1}"
    open(good, "w").write("\n".join(lines) + "\n")
    args = [CLI, f"--input={fastq}", "--k=31", "--l=22", "--s=4", "--mode=CUDA", "--check"]
    p = subprocess.run(args, capture_output=True, text=True)
    assert p.returncode == 1 and "total errors1" in p.stdout and "Should be" in p.stdout
    p = subprocess.run(args + ["--checkabort"], capture_output=True, text=True)
    assert p.returncode == 200                      # src/mains/main.cpp:285-291


@pytest.mark.gpu
def test_cli_table_full_exits_42(tmp_path):
    fastq = orc.golden_path("c2_uniform_k31.fastq", tmp_path)
    p = subprocess.run([CLI, f"--input={fastq}", "--k=31", "--l=8", "--mode=CUDA"], capture_output=True, text=True)
    assert p.returncode == 42                       # src/tsxcount/TSXHashMap.h:340-343


@pytest.mark.gpu
def test_cli_gz_input_and_n_bases(tmp_path):
    import gzip
    seqs = [b"ACGTACGTNACGTACGTACGT", b"ACGTACGTACGTACGTACGTAAAA"]
    fq = tmp_path / "n.fastq.gz"
    with gzip.open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@r%d\n%s\n+\n%s\n\n" % (i, s, b"&" * len(s)))
    dump = tmp_path / "n.count"
    p = subprocess.run([CLI, f"--input={fq}", "--k=8", "--l=10", "--mode=CUDA", f"--dump={dump}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    oc = orc.count_seqs(seqs, 8)
    got = dict(l.split("\t") for l in open(dump).read().splitlines())
    assert len(got) == oc.n_distinct and sum(int(v) for v in got.values()) == oc.n_total
    assert "Non-ACGT bases: 1" in p.stderr


@pytest.mark.gpu
def test_cli_parallel_readers_give_the_same_counts(tmp_path):
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    orc.golden_path("c1_bundled_k14.fastq.14.count", tmp_path)
    assert os.path.getsize(fastq) > 4 << 16
    p = subprocess.run([CLI, f"--input={fastq}", "--mode=CUDA", "--check", "--readers=4", "--threads=4"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Readers=4" in p.stderr
    assert "Added a total of 194697 different kmers" in p.stdout and "total errors0" in p.stdout


# ---- host feeder (FastxReader + tsxc_pack_reads) without a GPU -------------------------------------------
INGEST = os.path.join(ROOT, "tsxcount_b200", "bin", "ingest_check")


def _fnv(h, data):
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _expected_ingest(seqs):
    import numpy as np
    from tsxcount_b200 import sequtils
    h0 = 0xcbf29ce484222325
    hb = _fnv(h0, b"".join(seqs))
    hl = h0
    for s in seqs:
        hl = _fnv(hl, len(s).to_bytes(8, "little"))
    ascii_, off = sequtils.concat_reads(seqs)
    packed, seg, nbad = sequtils.pack_reads(ascii_, off)
    hp = _fnv(h0, packed[: (int(seg[-1]) + 31) // 32].tobytes())
    return len(seqs), sum(len(s) for s in seqs), nbad, hb, hl, hp


def _run_ingest(path, *extra):
    out = subprocess.run([INGEST, str(path), *map(str, extra)], capture_output=True, text=True, check=True).stdout
    kv = dict(item.split("=") for item in out.split())
    return (int(kv["reads"]), int(kv["bases"]), int(kv["bad"]), int(kv["hash_bases"], 16), int(kv["hash_lens"], 16),
            int(kv["hash_packed"], 16))


@pytest.mark.parametrize("batch,block", [(1 << 18, 8 << 20), (7, 64), (1, 17)])
def test_feeder_matches_python_reader_and_packer(tmp_path, batch, block):
    from tsxcount_b200 import sequtils
    fastq = orc.golden_path("c2_fakeseq_k31.fastq", tmp_path)
    seqs = sequtils.read_fastq(fastq)
    got, want = _run_ingest(fastq, batch, block), _expected_ingest(seqs)
    # every batch is packed on its own (bit 0 of word 0), so the packed hash is comparable only for one batch
    assert got[:5] == want[:5] and (batch < len(seqs) or got[5] == want[5])


def test_feeder_edge_cases(tmp_path):
    """Empty lines anywhere, no trailing newline, incomplete last record, a line longer than the block, gzip, N."""
    import gzip
    from tsxcount_b200 import sequtils
    body = b"@r1\nACGT\n+\n&&&&\n\n\n@r2\n\nGGNCC\n+\n&&&&&\n@r3\n" + b"ACGT" * 5000 + b"\n+\n" + b"&" * 20000 + b"\n@r4\nTTTT\n+"
    p = tmp_path / "e.fastq"
    p.write_bytes(body)
    seqs = sequtils.read_fastq(p)
    assert [len(s) for s in seqs] == [4, 5, 20000]             # r4 is incomplete: dropped (FastXReader.h:242)
    for batch, block in ((1 << 18, 8 << 20), (2, 32), (1, 16)):
        got, want = _run_ingest(p, batch, block), _expected_ingest(seqs)
        assert got[:5] == want[:5] and (batch < len(seqs) or got[5] == want[5])
    gz = tmp_path / "e.fastq.gz"
    with gzip.open(gz, "wb") as f:
        f.write(body)
    assert _run_ingest(gz) == _expected_ingest(seqs)
    misnamed = tmp_path / "gz_without_suffix.fastq"               # the gzip magic decides, not the name
    misnamed.write_bytes(gz.read_bytes())
    assert _run_ingest(misnamed) == _expected_ingest(seqs)


@pytest.mark.parametrize("ranges", [2, 3, 7, 64])
def test_feeder_byte_ranges_deliver_every_record_exactly_once(tmp_path, ranges):
    """--readers: N readers on disjoint byte ranges must reproduce the sequential stream, also when quality lines
    begin with '@' or '+', with empty lines between records and with range borders in every kind of line."""
    import random
    from tsxcount_b200 import sequtils
    rnd = random.Random(ranges)
    p = tmp_path / "r.fastq"
    with open(p, "w") as f:
        for i in range(3000):
            n = rnd.randint(1, 90)
            seq = "".join(rnd.choice("ACGT") for _ in range(n))
            qual = "".join(rnd.choice("@+I#>") for _ in range(n))
            f.write(f"@r{i} extra\n{seq}\n+\n{qual}\n")
            if rnd.random() < 0.05:
                f.write("\n")
    seqs = sequtils.read_fastq(p)
    assert len(seqs) == 3000
    want = _expected_ingest(seqs)
    for batch, block in ((1 << 18, 8 << 20), (5, 64)):
        got = _run_ingest(p, batch, block, ranges)
        assert got[:5] == want[:5]
    # threads, one per range: totals only
    out = subprocess.run([INGEST, str(p), "100", "4096", str(ranges), "parallel"], capture_output=True, text=True, check=True).stdout
    kv = dict(item.split("=") for item in out.split())
    assert (int(kv["reads"]), int(kv["bases"])) == want[:2]


def test_feeder_byte_ranges_more_ranges_than_records(tmp_path):
    from tsxcount_b200 import sequtils
    p = tmp_path / "tiny.fastq"
    p.write_bytes(b"@a\nACGT\n+\n@@@@\n@b\nGG\n+\n@+\n")
    seqs = sequtils.read_fastq(p)
    for ranges in (2, 5, 13, 26):
        assert _run_ingest(p, 10, 16, ranges)[:5] == _expected_ingest(seqs)[:5]


def test_feeder_multiline_fasta_split_keeps_the_kmer_multiset(tmp_path):
    """Extension over the reference: multi-line FASTA, soft-masked bases, long sequences cut into pieces that
    overlap by k-1 bases.  The pieces must match the Python mirror and carry exactly the k-mers of the records."""
    import random
    from tsxcount_b200 import sequtils
    rnd = random.Random(11)
    k = 21
    recs = [rnd.randint(1, 7000) for _ in range(12)] + [k - 1, k, 1000, 2000, 1000 + (1000 - (k - 1)), 0]
    whole = []
    p = tmp_path / "g.fa"
    with open(p, "w") as f:
        for i, n in enumerate(recs):
            seq = "".join(rnd.choice("ACGTacgtN" if rnd.random() < 0.02 else "ACGTacgt") for _ in range(n))
            whole.append(seq.upper().encode())
            f.write(f">chr{i} some description\n")
            for j in range(0, n, 60):
                f.write(seq[j:j + 60] + "\n")
            if rnd.random() < 0.3:
                f.write("\n")
    env = dict(os.environ, INGEST_PIECE="1000", INGEST_OVERLAP=str(k - 1))
    want_pieces = sequtils.read_fasta(p, piece_len=1000, overlap=k - 1)
    want = _expected_ingest(want_pieces)
    for batch, block in ((1 << 18, 8 << 20), (3, 64), (1, 100)):
        out = subprocess.run([INGEST, str(p), str(batch), str(block)], capture_output=True, text=True, check=True, env=env).stdout
        kv = dict(item.split("=") for item in out.split())
        got = (int(kv["reads"]), int(kv["bases"]), int(kv["bad"]), int(kv["hash_bases"], 16), int(kv["hash_lens"], 16))
        assert got == want[:5], (batch, block)
    a, b = orc.count_seqs([w for w in whole if w], k), orc.count_seqs(want_pieces, k)
    assert a.n_total == b.n_total and a.n_distinct == b.n_distinct
    assert (a.keys == b.keys).all() and (a.counts == b.counts).all()
    assert max(len(x) for x in want_pieces) == 1000


def _write_bgzf(path, data, block=30000, eof_marker=True):
    """BGZF as bgzip / htslib write it: independent gzip members of <= 64 KiB with a 'BC' extra subfield holding the
    member size - 1, then the empty end-of-file member."""
    import struct
    import zlib
    out = bytearray()
    for i in list(range(0, len(data), block)) + ([None] if eof_marker else []):
        chunk = b"" if i is None else data[i:i + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = co.compress(chunk) + co.flush()
        bsize = 18 + len(payload) + 8
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize - 1)
        out += payload + struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk))
    with open(path, "wb") as f:
        f.write(out)
    return bytes(out)


def test_feeder_inflates_bgzf_in_parallel_and_rejects_damage(tmp_path):
    """A bgzip (BGZF) input is inflated block-parallel and must deliver exactly the stream of the plain file; a flipped
    payload byte (CRC), a file cut inside a block and garbage where a block header should be all make the reader throw.
    An ordinary single-stream .gz still goes through gzread."""
    import gzip
    import random
    rnd = random.Random(5)
    seqs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(1, 400))) for _ in range(6000)]
    text = "".join(f"@r{i}\n{s}\n+\n{'I' * len(s)}\n" for i, s in enumerate(seqs)).encode()
    plain, bgz, gz = tmp_path / "r.fastq", tmp_path / "r.fastq.bgz", tmp_path / "r.fastq.gz"
    plain.write_bytes(text)
    raw = _write_bgzf(bgz, text)
    with gzip.open(gz, "wb") as f:
        f.write(text)

    def ingest(path, threads, block=8 << 20):
        env = dict(os.environ, INGEST_INFLATE_THREADS=str(threads))
        return subprocess.run([INGEST, str(path), "1000", str(block)], capture_output=True, text=True, env=env)

    want = ingest(plain, 1).stdout
    assert "reads=6000" in want
    for threads in (2, 4, 16):
        p = ingest(bgz, threads, block=1 << 16)
        assert p.returncode == 0 and p.stdout == want, p.stderr
        assert "parallel inflate: yes (BGZF)" in p.stderr
    assert ingest(bgz, 1).stdout == want                     # one thread: plain gzread reads BGZF too
    p = ingest(gz, 4)
    assert p.stdout == want and "parallel inflate: no" in p.stderr
    # damage
    flipped = bytearray(raw); flipped[len(raw) // 2] ^= 0x55
    (tmp_path / "flip.bgz").write_bytes(flipped)
    (tmp_path / "cut.bgz").write_bytes(raw[:len(raw) * 2 // 3])
    for name in ("flip.bgz", "cut.bgz"):
        p = ingest(tmp_path / name, 4)
        assert p.returncode != 0 and "damaged or truncated" in p.stderr, (name, p.stderr[-300:])


def test_feeder_rejects_truncated_gzip(tmp_path):
    """A damaged .gz must not look like a (shorter) input: the reader throws, the tool exits non-zero."""
    import gzip
    seqs = orc.gen_reads(seed=9, n_reads=4000, read_len=150, mode=0)
    raw = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"&" * len(s)) for i, s in enumerate(seqs))
    good = tmp_path / "ok.fastq.gz"
    with gzip.open(good, "wb") as f:
        f.write(raw)
    assert _run_ingest(good)[:2] == (4000, 4000 * 150)
    data = good.read_bytes()
    bad = tmp_path / "cut.fastq.gz"
    bad.write_bytes(data[: len(data) // 2])
    p = subprocess.run([INGEST, str(bad)], capture_output=True, text=True)
    assert p.returncode != 0


@pytest.mark.gpu
def test_reference_tree_binding_drives_addkmer_and_getkmercount():
    """The subclass of the reference's own TSXHashMap (host/ref_binding/TSXHashMapCUDA_ref.h), compiled against
    the reference headers in the build container, driven through the base-class pointer with the reference's own
    createKMers / fromSequence / UBigInt: counts come back exactly (oracle/ref_adapter_main.cpp)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_adapter_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_adapter_check was not built (reference tree absent at build time)")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "total errors0" in p.stdout


def test_cli_new_options_are_listed_and_n_policy_is_explicit():
    out = subprocess.run([CLI, "--help"], capture_output=True, text=True).stdout
    for opt in ("--gpus=", "--n-policy=", "--dump=", "--readers=", "--canonical", "--histogram="):
        assert opt in out, opt
    p = subprocess.run([CLI, "--input=x.fastq", "--mode=CUDA", "--n-policy=random"], capture_output=True, text=True)
    assert p.returncode == 2 and "only 'skip' exists" in p.stderr


@pytest.mark.gpu
def test_cli_multi_gpu_count_check_dump(tmp_path):
    """--gpus=N: one host process, the table hash-sharded over N GPUs, NCCL all-gather / barrier and peer stores
    (host/MultiGpuCounter.h).  Small batches so that several super-batches and rounds happen."""
    import tsxcount_b200 as tsx
    n = tsx._lib.load().tsxc_device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    fastq = orc.golden_path("c1_bundled_k14.fastq", tmp_path)
    golden = orc.golden_path("c1_bundled_k14.fastq.14.count", tmp_path)
    dump = tmp_path / "out.count"
    p = subprocess.run([CLI, f"--input={fastq}", "--mode=CUDA", "--gpus=2", "--batch-reads=40", "--check", f"--dump={dump}"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "GPUs=2" in p.stderr
    assert "Added a total of 194697 different kmers" in p.stdout and "total errors0" in p.stdout
    assert "queried (Xor) kmer count: 0" in p.stdout
    assert sorted(open(dump).read().splitlines()) == sorted(open(golden).read().splitlines())
