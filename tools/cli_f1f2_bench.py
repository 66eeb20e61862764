"""f1 / f2 measurements through the C++ CLI on the GPU box.
f1: FASTQ -> counts end to end on a >= 10 GB file (a 1 GB block of synthetic reads written 11 times), --readers=1/8.
f2: --dump (text formatted on the device) and --check (region-sorted lookups) on a table of ~1e8 distinct 31-mers.
Usage: python tools/cli_f1f2_bench.py [block_reads] [repeat]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cli_ingest_bench import CLI, make_fastq  # noqa: E402


def run(args):
    t0 = time.time()
    p = subprocess.run([CLI] + args, capture_output=True, text=True)
    dt = time.time() - t0
    keep = [l for l in (p.stdout + p.stderr).splitlines() if l.startswith(("Added", "Counted", "total errors", "Dumped", "Kmer count check"))]
    return dt, p.returncode, keep


def main():
    block = int(sys.argv[1]) if len(sys.argv) > 1 else 3_200_000
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    blk = "/tmp/f1_block.fastq"
    big = "/tmp/f1_big.fastq"
    make_fastq(blk, block)
    data = open(blk, "rb").read()
    with open(big, "wb") as f:
        for _ in range(rep):
            f.write(data)
    size = os.path.getsize(big)
    bases = block * 150 * rep
    print(f"f1 file: {size / 1e9:.2f} GB, {block * rep} reads, {bases / 1e9:.2f} Gbases, {block * rep * 120 / 1e9:.2f} G 31-mers", flush=True)
    for readers, threads in ((1, 3), (8, 16), (8, 16), (12, 16)):
        dt, rc, keep = run([f"--input={big}", "--k=31", "--l=31", "--s=4", "--mode=CUDA", f"--readers={readers}", f"--threads={threads}"])
        print(f"f1 readers={readers} threads={threads} rc={rc} wall={dt:.2f}s {size / dt / 1e9:.2f} GB/s of FASTQ, {bases / dt / 1e9:.3f} Gbases/s, "
              f"{block * rep * 120 / dt / 1e9:.3f} Gk-mer/s incl. start-up  {keep}", flush=True)
    os.remove(big)
    # f2: 800 000 reads -> 9.6e7 distinct 31-mers
    small = "/tmp/f2_small.fastq"
    make_fastq(small, 800_000, seed=7)
    base = ["--input=" + small, "--k=31", "--l=29", "--s=4", "--mode=CUDA", "--readers=4", "--threads=8"]
    dt0, rc, keep = run(base)
    print(f"f2 count only: rc={rc} wall={dt0:.2f}s {keep}", flush=True)
    dump = small + ".31.count"
    dt1, rc, keep = run(base + ["--dump=" + dump])
    dsize = os.path.getsize(dump) if os.path.exists(dump) else 0
    print(f"f2 count + dump: rc={rc} wall={dt1:.2f}s dump {dsize / 1e9:.2f} GB in {dt1 - dt0:.2f}s = {dsize / max(dt1 - dt0, 1e-9) / 1e9:.2f} GB/s of text {keep}", flush=True)
    dt2, rc, keep = run(base + ["--check"])
    print(f"f2 count + check: rc={rc} wall={dt2:.2f}s check of {dsize / 1e9:.2f} GB in {dt2 - dt0:.2f}s {keep}", flush=True)
    for pth in (small, dump, blk):
        if os.path.exists(pth):
            os.remove(pth)


if __name__ == "__main__":
    main()
