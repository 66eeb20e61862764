#!/usr/bin/env python3
"""Pin the CPU restatement (oracle.c) against the UNMODIFIED reference binary.  TEST INFRASTRUCTURE ONLY.

Runs in the build container only (needs oracle/_ref/tsxCount, built by `make -C oracle ref`
from /root/reference).  For every case below it
  1. writes a seeded synthetic FASTQ with `kmer_oracle gen` (or takes the bundled fixture),
  2. writes `<fastq>.<k>.count` with `kmer_oracle count` (the restatement's answer),
  3. runs the reference `tsxCount --input=<fastq> --k --l --s --mode=<M> --threads=<T> --check`
     (src/mains/main.cpp:224-396), which compares ITS OWN table against that file k-mer by k-mer
     and asserts that it holds no additional k-mers (XOR of queried start positions, :378-384),
  4. records `total errors`, the three k-mer counts and sha256 of both files.
A case pins the oracle when total errors == 0 and reference/queried/tsxCount counts all equal the
oracle's distinct count and the XOR count is 0.

The FASTQ + count fixtures small enough to commit are copied to tests/golden/ so that the GPU box
(where /root/reference does not exist) can replay them; the verdicts go to
tests/golden/ref_binary_pins.json.
"""
import gzip
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(HERE, "_ref", "tsxCount")
ORACLE = os.path.join(HERE, "_build", "kmer_oracle")
REFDATA = "/root/reference/data/small_t7.1000.fastq"

# name, generator (mode, seed, n_reads, read_len, genome_len, sub_q16) or None for the bundled file,
# k, l, s, [(mode, threads)...].  Domain limits of the reference binary: SURVEY.md §0.5 / §8(c).
CASES = [
    # the reference runs at ~300 k-mers/s/thread at k=31 and slower at k=63 (O(k^2) hash on heap big-ints),
    # so the synthetic pins are a few 10^4 k-mers each
    ("c1_bundled_k14", None, 14, 26, 4, [("SERIAL", 1), ("OMP", 8)]),
    ("c2_uniform_k31", (0, 0xC2, 400, 150, 0, 0), 31, 22, 4, [("SERIAL", 1), ("OMP", 8)]),
    # heavy-hitter poly-A k-mer with an overflow entry: SERIAL pins it; the reference's own OMP mode miscounts
    # that k-mer under contention (total errors 1 at 8 threads) and is recorded for information only
    ("c2_fakeseq_k31", (1, 0xC2, 300, 150, 0, 0), 31, 22, 4, [("SERIAL", 1), ("OMP", 8, "info")]),
    ("c5_genome_k31", (3, 0xC5, 400, 150, 3000, 328), 31, 22, 4, [("OMP", 8)]),
    ("c3_flat_k63", (0, 0xC3, 120, 150, 0, 0), 63, 20, 4, [("OMP", 8)]),
    ("short_reads_k20", (0, 0x51, 1500, 24, 0, 0), 20, 18, 4, [("OMP", 8)]),
]


# Pins of the restatement alone (word-boundary and in-between k; the GPU tests do not replay these, so no fixture is
# committed: tests/test_oracle.py regenerates the FASTQ with the same generator call and compares the hashes).
# Domain of the reference's --check: 2k - l + 2s <= 64 once an overflow entry exists, so k >= 48 stays flat (counts <= 15).
ORACLE_ONLY_CASES = [
    # (OMP at 8 threads lost one increment of a contended k-mer in one run: the reference's own race, recorded for information)
    ("k28_genome", (3, 0x28, 400, 150, 3000, 328), 28, 22, 4, [("SERIAL", 1), ("OMP", 8, "info")]),
    ("k32_genome_overflow", (3, 0x32, 400, 150, 3000, 328), 32, 22, 4, [("SERIAL", 1), ("OMP", 8)]),
    ("k34_genome_overflow", (3, 0x34, 300, 150, 3000, 328), 34, 22, 4, [("OMP", 8)]),
    ("k36_fakeseq", (1, 0x36, 200, 150, 0, 0), 36, 22, 4, [("SERIAL", 1)]),
    ("k48_genome_flat", (3, 0x48, 150, 150, 5000, 200), 48, 20, 4, [("OMP", 8)]),
    ("k64_genome_flat", (3, 0x64, 100, 150, 4000, 100), 64, 20, 4, [("OMP", 8)]),
    ("k64_uniform", (0, 0x65, 120, 150, 0, 0), 64, 20, 4, [("SERIAL", 1)]),
]


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def run_ref(fastq, k, l, s, mode, threads):
    cmd = [REFBIN, f"--input={fastq}", f"--k={k}", f"--l={l}", f"--s={s}", f"--mode={mode}",
           f"--threads={threads}", "--check"]
    for attempt in range(4):  # the reference occasionally segfaults at start-up
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
        if p.returncode == 0:
            break
    out = p.stdout

    def grab(pat):
        m = re.search(pat, out)
        return int(m.group(1)) if m else None

    return {
        "mode": mode, "threads": threads, "returncode": p.returncode,
        "added_distinct": grab(r"Added a total of (\d+) different kmers"),
        "total_errors": grab(r"total errors(\d+)"),
        "reference_kmer_count": grab(r"Reference kmer count: (-?\d+)"),
        "queried_kmer_count": grab(r"queried kmer count: (-?\d+)"),
        "tsxcount_kmer_count": grab(r"tsxCount kmer count: (-?\d+)"),
        "xor_kmer_count": grab(r"queried \(Xor\) kmer count: (-?\d+)"),
    }


def main():
    for need in (REFBIN, ORACLE):
        if not os.path.exists(need):
            sys.exit(f"missing {need}: run `make -C oracle all` in the build container first")
    os.makedirs(GOLD, exist_ok=True)
    pins_path = os.path.join(GOLD, "ref_binary_pins.json")
    oracle_only = "--oracle-only" in sys.argv[1:]     # add / refresh the fixture-less pins, keep the others as they are
    if oracle_only:
        pins = json.load(open(pins_path))
        pins["oracle_only_cases"] = []
    else:
        pins = {"made_by": "oracle/make_ref_pins.py", "reference_binary": "oracle/_ref/tsxCount (unmodified sources)",
                "cases": [], "oracle_only_cases": []}
    todo = [(c, "oracle_only_cases") for c in ORACLE_ONLY_CASES]
    if not oracle_only:
        todo = [(c, "cases") for c in CASES] + todo
    with tempfile.TemporaryDirectory() as tmp:
        for (name, gen, k, l, s, runs), dest in todo:
            fastq = os.path.join(tmp, name + ".fastq")
            if gen is None:
                shutil.copy(REFDATA, fastq)
            else:
                subprocess.run([ORACLE, "gen"] + [str(x) for x in gen] + [fastq], check=True)
            count = f"{fastq}.{k}.count"
            p = subprocess.run([ORACLE, "count", fastq, str(k), count], check=True, capture_output=True, text=True)
            m = re.search(r"distinct=(\d+) total=(\d+)", p.stderr)
            distinct, total = int(m.group(1)), int(m.group(2))
            case = {"name": name, "gen": gen, "k": k, "l": l, "s": s, "oracle_distinct": distinct,
                    "oracle_total": total, "fastq_sha256": sha(fastq), "count_sha256": sha(count), "runs": []}
            if gen is None:
                case["oracle_dump_identical_to_bundled_count"] = (
                    open(count, "rb").read() == open(REFDATA + f".{k}.count", "rb").read())
            for run in runs:
                mode, threads = run[0], run[1]
                r = run_ref(fastq, k, l, s, mode, threads)
                r["informational"] = len(run) > 2
                r["pinned"] = (r["returncode"] == 0 and r["total_errors"] == 0 and r["xor_kmer_count"] == 0
                               and r["reference_kmer_count"] == distinct and r["tsxcount_kmer_count"] == distinct)
                print(name, r)
                case["runs"].append(r)
            case["pinned"] = all(r["pinned"] for r in case["runs"] if not r["informational"])
            if dest == "oracle_only_cases":
                lines = [ln.rstrip("\n").split("\t") for ln in open(count)]
                case["max_count"] = max(int(c) for _, c in lines)
            pins[dest].append(case)
            if dest != "cases":
                continue
            # fixtures: FASTQ + count dump, gzip'ed
            for src in (fastq, count):
                with open(src, "rb") as fi, gzip.GzipFile(os.path.join(GOLD, os.path.basename(src) + ".gz"), "wb",
                                                          mtime=0) as fo:
                    shutil.copyfileobj(fi, fo)
    with open(pins_path, "w") as f:
        json.dump(pins, f, indent=1)
    bad = [c["name"] for c in pins["cases"] + pins["oracle_only_cases"] if not c["pinned"]]
    print("UNPINNED:", bad if bad else "none")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
