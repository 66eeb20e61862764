"""End-to-end CLI ingest measurement: synthetic FASTQ -> tsxcount --mode=CUDA with 1 and N readers.
Usage: python tools/cli_ingest_bench.py [n_reads] [readers] ; prints one line per run."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "tsxcount_b200", "bin", "tsxcount")


def make_fastq(path, n, L=150, seed=1):
    rng = np.random.default_rng(seed)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L), dtype=np.uint8)]
    rec = np.empty((n, 9 + L + 1 + 2 + L + 1), dtype=np.uint8)
    rec[:, :8] = np.frombuffer(b"@read000", dtype=np.uint8)
    rec[:, 8] = 10
    rec[:, 9:9 + L] = seq
    rec[:, 9 + L] = 10
    rec[:, 10 + L] = ord("+")
    rec[:, 11 + L] = 10
    rec[:, 12 + L:12 + 2 * L] = ord("I")
    rec[:, 12 + 2 * L] = 10
    rec.tofile(path)
    return n * L


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    readers = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    path = "/tmp/cli_ingest.fastq"
    bases = make_fastq(path, n)
    for r in readers:
        for rep in range(reps):
            t0 = time.time()
            p = subprocess.run([CLI, f"--input={path}", "--k=31", "--l=30", "--s=4", "--mode=CUDA", f"--readers={r}",
                                f"--threads={max(3, r + 2)}"], capture_output=True, text=True)
            dt = time.time() - t0
            added = [l for l in p.stdout.splitlines() if l.startswith("Added")]
            counted = [l for l in p.stderr.splitlines() if l.startswith("Counted")]
            print(f"readers={r} rep={rep} rc={p.returncode} wall={dt:.2f}s {bases / dt / 1e9:.3f} Gbases/s {added} {counted}", flush=True)
    os.remove(path)


if __name__ == "__main__":
    main()
