#!/usr/bin/env python3
"""bench.py — k-mers counted per second on B200 (BASELINE.json metric), driver contract.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c2-fakeseq|c3|c4|c5]
                  [--scale F]

A "step" is one pass of the hot path over the whole workload on a freshly zeroed table:
  N=1   config 2 of BASELINE.json: 66 666 667 synthetic 150 bp reads (10 Gbases, 8.0e9 31-mers), k=31,
        2^34-slot table (128 GiB) on one B200.
  N>1   the same read generator, N x as many reads, table of 2^(34+log2 N) slots hash-sharded over the N GPUs
        (config 5's routing: extract+hash -> bin by owner -> all-to-all over NVLink -> insert), weak scaling.
`value`  = k-mers of all ranks / device time of the K timed steps (reads already packed in HBM; max over ranks).
`e2e`    = the same through the C ABI with HOST buffers: tsxc_add_reads() from pinned host memory (H2D copies
           inside the timed region) + tsxc_sync() + tsxc_distinct() read-back.
Table zeroing between steps is outside the timed regions and reported as clear_ms.
--impl reference times the reference's own CPU implementation (oracle/_ref/tsxCount, the unmodified
reference sources compiled by oracle/Makefile) on a bounded sample of the same generator, all host threads.
"""
import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (k, per-GPU log2 slots, gen mode, per-GPU reads, read_len, genome_len, sub_q16, seed, description)
    "c2": dict(k=31, l=34, mode=0, reads=66_666_667, read_len=150, genome=0, sub=0, seed=0xC2,
               desc="config 2: synthetic uniform 150bp reads, 10 Gbases, k=31, 2^34 slots (128 GiB)"),
    "c2-fakeseq": dict(k=31, l=34, mode=1, reads=66_666_667, read_len=150, genome=0, sub=0, seed=0xC2,
                       desc="config 2 (generateFakeSequences.py style: random body + poly-A tail), 10 Gbases, k=31"),
    "c3": dict(k=63, l=33, mode=2, reads=66_666_667, read_len=150, genome=1 << 24, sub=655, seed=0xC3,
               desc="config 3: log-uniform (Zipf-like) dictionary of 2^24 150-mers, 1% substitutions, k=63"),
    "c4": dict(k=127, l=32, mode=0, reads=66_666_667, read_len=150, genome=0, sub=0, seed=0xC4,
               desc="config 4: synthetic uniform 150bp reads, 10 Gbases, k=127, 2^32 slots x 32 B (128 GiB)"),
    "c5": dict(k=31, l=34, mode=3, reads=83_333_333, read_len=150, genome=3_100_000_000, sub=328, seed=0xC5,
               desc="config 5 slice: reads sampled from a 3.1 Gbase synthetic genome, 0.5% substitutions, k=31"),
}


_T0 = time.perf_counter()


def log(msg):
    """progress on stderr (stdout carries exactly one JSON line)"""
    print(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def read_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.rows, self.proc, self.gpu = [], None, gpu

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "tsxCount")
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None


# The reference's OMP / PTHREAD modes live-lock at high thread counts and a few launches segfault at start-up
# (SURVEY.md §0.5): 8 threads is the survey's stable point.  Every launch is bounded; a sample is sized so that one
# run takes a few seconds at the ~0.3-0.9 M k-mers/s these modes reach.
REF_THREADS = 8
REF_TIMEOUT_S = 10.0          # a good run of the default sample takes 2-3 s; a live-locked one never ends


def host_info():
    rtm = False
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags") and " rtm" in line:
                    rtm = True
                if line.startswith("model name") and not model:
                    model = line.split(":", 1)[1].strip()
    except OSError:
        pass
    return {"nproc": os.cpu_count() or 1, "rtm": rtm, "cpu": model}


def _run_ref_cli(binp, fq, k, l, mode, threads, timeout, check=False):
    """-> (wall seconds | None, stdout text).  None: crashed, hung or non-zero exit."""
    cmd = [binp, f"--input={fq}", f"--k={k}", f"--l={l}", "--s=4", f"--mode={mode}", f"--threads={threads}"]
    if check:
        cmd.append("--check")
    t0 = time.perf_counter()
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout, text=True, errors="replace")
    except subprocess.TimeoutExpired:
        return None, ""
    return (time.perf_counter() - t0 if p.returncode == 0 else None), p.stdout


def _write_fastq(path, seqs):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@seq_%d\n%s\n+\n%s\n" % (i, s, b"&" * len(s)))


class RefSample:
    """A bounded sample of the workload's generator as a FASTQ file for the reference CLI (CPU legs only: the one
    place bench.py executes anything under oracle/)."""

    def __init__(self, wl, n_reads, tmp):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_py as orc
        self.orc, self.wl, self.tmp = orc, wl, tmp
        self.seqs = orc.gen_reads(seed=wl["seed"], n_reads=wl["reads"], read_len=wl["read_len"], mode=wl["mode"],
                                  genome_len=wl["genome"], sub_rate_q16=wl["sub"], first=0, count=n_reads)
        self.k = wl["k"]
        self.n_kmers = sum(max(0, len(s) - self.k + 1) for s in self.seqs)
        self.path = os.path.join(tmp, f"sample_{n_reads}.fastq")
        _write_fastq(self.path, self.seqs)

    def time_ref(self, binp, mode, threads, retries=5, l=25):
        """Whole-process wall clock of the count phase (the authors' method, analyses/perform_analyses.py:64)."""
        for _ in range(retries):
            dt, _out = _run_ref_cli(binp, self.path, self.k, l, mode, threads, REF_TIMEOUT_S)
            if dt is not None:
                return dt
        return None

    def time_port(self):
        t0 = time.perf_counter()
        self.orc.count_seqs(self.seqs, self.k)
        return time.perf_counter() - t0


def cas_errors(binp, wl, tmp, n_reads=60):
    """`total errors` of the reference's own --check in CAS mode on a tiny sample (it checks ~300 k-mers/s/thread):
    that mode miscounts at k >= 31.  The .count file it checks against is the oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as orc
    seqs = orc.gen_reads(seed=wl["seed"], n_reads=wl["reads"], read_len=wl["read_len"], mode=wl["mode"],
                         genome_len=wl["genome"], sub_rate_q16=wl["sub"], first=0, count=n_reads)
    fq = os.path.join(tmp, "cas_check.fastq")
    _write_fastq(fq, seqs)
    oc = orc.count_seqs(seqs, wl["k"])
    orc.write_dump_fastq(fq, wl["k"], f"{fq}.{wl['k']}.count")
    dt, out = _run_ref_cli(binp, fq, wl["k"], 22, "CAS", REF_THREADS, 60.0, check=True)
    m = re.search(r"total errors\s*(\d+)", out or "")
    return {"kmers_checked": int(oc.n_distinct), "total_errors": int(m.group(1)) if m else None,
            "note": "reference --check in CAS mode against the oracle's counts of a 60-read sample"}


def cpu_baseline_block(wl, ref_reads):
    """cpu_baseline of the N=1 line: OMP and CAS (TSX only when the CPU has RTM) at 1 and REF_THREADS threads on a
    bounded sample of the workload, plus config 1 (the bundled FASTQ) as shipped."""
    info = host_info()
    binp = ref_binary()
    with tempfile.TemporaryDirectory() as tmp:
        big = RefSample(wl, ref_reads, tmp)
        small = RefSample(wl, max(200, ref_reads // 6), tmp)          # single-thread runs are ~6x slower
        modes = {"cores": info["nproc"], "rtm": info["rtm"], "cpu": info["cpu"], "threads_used": REF_THREADS}
        head = None
        if binp:
            for mode in ("OMP", "CAS") + (("TSX",) if info["rtm"] else ()):
                dt8 = big.time_ref(binp, mode, REF_THREADS)
                dt1 = small.time_ref(binp, mode, 1)
                modes[mode] = {f"t{REF_THREADS}_Gkmer_s": big.n_kmers / dt8 / 1e9 if dt8 else None,
                               "t1_Gkmer_s": small.n_kmers / dt1 / 1e9 if dt1 else None}
                if mode == "OMP" and dt8:
                    head = (big.n_kmers / dt8 / 1e9, dt8)
            if not info["rtm"]:
                modes["TSX"] = None                                   # would spin forever without RTM
            try:
                modes["CAS"]["check"] = cas_errors(binp, wl, tmp)
            except Exception as e:                                    # the checker must not take the bench down
                modes["CAS"]["check"] = {"error": str(e)[:200]}
            # config 1: data/small_t7.1000.fastq as committed under tests/golden (202 204 14-mers)
            try:
                import gzip
                import shutil
                c1 = os.path.join(tmp, "c1.fastq")
                with gzip.open(os.path.join(ROOT, "tests", "golden", "c1_bundled_k14.fastq.gz"), "rb") as fi, open(c1, "wb") as fo:
                    shutil.copyfileobj(fi, fo)
                c1m = {}
                for mode, th in (("OMP", REF_THREADS), ("CAS", REF_THREADS), ("SERIAL", 1)):
                    dt, _ = _run_ref_cli(binp, c1, 14, 26, mode, th, REF_TIMEOUT_S)
                    c1m[f"{mode}_t{th}_Gkmer_s"] = 202204 / dt / 1e9 if dt else None
                modes["config1_bundled_k14"] = c1m
            except Exception as e:
                modes["config1_bundled_k14"] = {"error": str(e)[:200]}
        if head:
            value, dt, kind, cores = head[0], head[1], "reference", REF_THREADS
            what = f"reference CLI --mode=OMP --threads={REF_THREADS} --l=25 --s=4, whole-process wall clock {dt:.1f} s"
        else:
            dt = big.time_port()
            value, kind, cores = big.n_kmers / dt / 1e9, "port", 1
            what = f"C restatement (oracle/), 1 thread, {dt:.1f} s: the reference binary is missing or failed"
        return {"value": value, "unit": "Gk-mer/s", "cores": cores, "kind": kind,
                "sample": f"first {ref_reads} reads of the workload ({big.n_kmers} k-mers), {what}", "modes": modes}


def run_reference_arm(args, wl, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation of the path (unmodified CLI, --mode=OMP,
    REF_THREADS threads) on a bounded sample per step.  If the binary is missing or a step fails five times in a row
    (the reference's OMP mode live-locks or segfaults at start-up in roughly one launch out of three), the WHOLE
    line is timed on the C restatement instead and says kind = "port": the two are never mixed."""
    if rank != 0:
        return
    info = host_info()
    binp = ref_binary()
    n_reads = args.ref_reads
    with tempfile.TemporaryDirectory() as tmp:
        sample = RefSample(wl, n_reads, tmp)
        kind, times = "reference", []
        if binp:
            for i in range(args.warmup + args.steps):
                if i < args.warmup and i >= 1:
                    continue                      # one warm-up launch pages the binary in; more only burn the time budget
                dt = sample.time_ref(binp, "OMP", REF_THREADS)
                if dt is None:
                    log("reference CLI failed five times in a row: timing the whole line on the C restatement instead")
                    kind, times = "port", []
                    break
                if i >= args.warmup:
                    times.append(dt)
        else:
            kind = "port"
        if kind == "port":
            for i in range(min(args.warmup, 1) + args.steps):
                dt = sample.time_port()
                if i >= min(args.warmup, 1):
                    times.append(dt)
        T = sum(times)
        value = args.steps * sample.n_kmers / T / 1e9
        threads = REF_THREADS if kind == "reference" else 1
        what = (f"reference CLI --mode=OMP --threads={REF_THREADS} --l=25 --s=4" if kind == "reference"
                else "C restatement of the reference algorithm (oracle/), 1 thread")
        line = {
            "impl": "reference", "metric": "k-mers counted/sec", "value": value, "unit": "Gk-mer/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"], "k": wl["k"], "sample": f"first {n_reads} reads, {what}"},
            "cpu_baseline": {"value": value, "unit": "Gk-mer/s", "cores": threads, "kind": kind, "host_cores": info["nproc"],
                             "rtm": info["rtm"],
                             "sample": f"{n_reads} reads x {wl['read_len']} bp = {sample.n_kmers} k-mers per step, "
                                       f"whole-process wall clock, {what}"},
            "e2e": {"value": value, "unit": "Gk-mer/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)


def main():
    import faulthandler
    faulthandler.enable()
    if os.environ.get("TSX_BENCH_WATCHDOG"):
        faulthandler.dump_traceback_later(int(os.environ["TSX_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink reads and table together (development only)")
    ap.add_argument("--ref-reads", type=int, default=12_000,
                    help="reads per step of the CPU reference sample (12 000 x 150 bp = 1.44e6 31-mers, 2-3 s at 8 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-batch-reads", type=int, default=0,
                    help="reads per tsxc_add_reads call of the e2e leg (0 = half of the workload: two calls, the second "
                         "one's H2D copy overlaps the first one's counting, and each call fills one insert pass)")
    ap.add_argument("--force-sharded", action="store_true", help="run the routed multi-GPU data path even with one rank")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed oracle-parity preamble of the multi-GPU arm")
    ap.add_argument("--no-variants", action="store_true", help="time only the headline workload")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    if args.scale != 1.0:
        shrink = 0
        while (1 << (shrink + 1)) <= round(1 / args.scale):
            shrink += 1
        wl["reads"] = max(1000, int(wl["reads"] * args.scale))
        wl["l"] -= shrink
        wl["desc"] += f" [scaled x{args.scale}]"

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return 0

    import tsxcount_b200 as tsx
    lib = tsx._lib.load()
    if lib.tsxc_device_count() < 1:
        raise SystemExit("bench.py needs a B200: tsxcount_b200 has no CPU fallback")
    if world > 1 or args.force_sharded:
        import bench_mgpu
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
        extra = []
        if args.workload == "c2" and not args.no_variants:
            extra = [dict(WORKLOADS["c5"], name="c5")]
        return bench_mgpu.bench_main(args, dict(wl, name=args.workload), rank, world, local_rank, log, extra_workloads=extra)

    dev = local_rank

    def dalloc(nbytes):
        p = C.c_void_p()
        tsx._lib.check(lib.tsxc_device_alloc(dev, nbytes, C.byref(p)))
        return p

    peak, peak_src = read_peaks()
    handles = {}                     # (k, l) -> table, kept across workloads of the same shape (allocation is slow)

    def table_for(w):
        key = (w["k"], w["l"])
        if key not in handles:
            for h in handles.values():
                h.close()
            handles.clear()
            handles[key] = tsx.TSXHashMapCUDA(w["l"], 0, w["k"], device=dev, flags=int(os.environ.get("TSXC_BENCH_FLAGS", w.get("flags", 0))))
            handles[key].sync()
        return handles[key]

    def run_workload(w, warmup, steps, full):
        """Device-timed steps of one workload; full = also K0r, e2e, launches for the headline line."""
        k, l = w["k"], w["l"]
        n_reads, read_len = w["reads"], w["read_len"]
        n_bases = n_reads * read_len
        n_words = (n_bases + 31) // 32
        n_kmers = n_reads * max(0, read_len - k + 1)
        # ---- inputs resident in HBM (generated on the device; identical to the oracle's generator) ----
        for h in handles.values():
            h.trim()                 # the pipeline's key buffers fill free HBM: let them be re-sized around the new inputs
        d_packed, d_off = dalloc((n_words + 8) * 8), dalloc((n_reads + 1) * 8)
        gp = tsx.TsxcGenParams(w["seed"], n_reads, read_len, w["mode"], w["genome"], w["sub"], 0)
        tsx._lib.check(lib.tsxc_gen_reads_device(C.byref(gp), 0, n_reads, dev, None, d_packed, d_off))
        hm = table_for(w)
        layout = hm.stats()

        def one_step():
            t0 = time.perf_counter()
            hm.clear(); hm.sync()
            clear_s = time.perf_counter() - t0
            hm.mark(0)
            hm.addReadsDevice(d_packed, d_off, n_reads, n_bases)
            hm.mark(1)
            hm.sync()
            st = hm.stats()
            st["step_ms"] = hm.elapsed_ms(0, 1)     # every launch of the step, on the launching stream
            assert st["kmers_added"] == n_kmers and st["error_flags"] == 0, st
            return st, clear_s

        log(f"{w['name']}: inputs generated: {n_reads} reads, {n_kmers} k-mers; table {layout['table_bytes'] / 2**30:.1f} GiB")
        sampler = ClockSampler(dev)
        sampler.start()              # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up ...
        for i in range(warmup):
            st, _ = one_step()
            log(f"{w['name']} warmup {i}: {st['step_ms']:.1f} ms")
        sampler.rows.clear()         # ... and keep only what it reports during the timed steps
        step_ms, launches, clear_ms, phases = [], 0, [], []
        for _ in range(steps):
            st, clear_s = one_step()
            launches += st["kernel_launches"]
            clear_ms.append(1e3 * clear_s)
            step_ms.append(st["step_ms"])
            phases.append({"hist": st["hist_ms"], "part1": st["part1_ms"], "insert": st["insert_ms"]})
        clocks = sampler.stop()
        log(f"{w['name']} timed steps: {[round(x, 1) for x in step_ms]} ms; phases {phases[-1]}")
        res = {"w": w, "n_kmers": n_kmers, "n_reads": n_reads, "n_words": n_words, "read_len": read_len, "k": k, "l": l,
               "step_ms": step_ms, "launches": launches, "clear_ms": clear_ms, "clocks": clocks, "st": st, "layout": layout,
               "phase_ms": {p: statistics.mean(x[p] for x in phases) for p in phases[0]},
               "value": steps * n_kmers / sum(step_ms) / 1e6}
        if full:
            res["sampled"] = sampled_lookup_check(hm, w, n_reads)
            res["k0"] = measure_k0(hm, layout, st, n_kmers)
            if not args.no_e2e:
                res["e2e"] = run_e2e(hm, d_packed, n_reads, read_len, n_words, n_kmers, st["distinct"], steps)
        lib.tsxc_device_free(dev, d_packed); lib.tsxc_device_free(dev, d_off)
        return res

    def sampled_lookup_check(hm, w, n_reads):
        """Full-size parity beyond the invariants: k-mers of 200 reads regenerated by the oracle's generator must be
        present with at least the count the sample itself gives them, k-mers of a foreign seed must be absent."""
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        import oracle_py as orc
        first = (n_reads // 3) * 2
        sample = orc.gen_reads(seed=w["seed"], n_reads=n_reads, read_len=w["read_len"], mode=w["mode"], genome_len=w["genome"],
                               sub_rate_q16=w["sub"], first=first, count=200)
        oc = orc.count_seqs(sample, w["k"])
        got = hm.getKmerCounts(oc.keys_kw(hm.kw))
        other = orc.count_seqs(orc.gen_reads(seed=w["seed"] + 0x9999, n_reads=64, read_len=w["read_len"], mode=0), w["k"])
        absent = hm.getKmerCounts(other.keys_kw(hm.kw))
        exact = bool(np.array_equal(got, oc.counts))
        ok = bool((got >= oc.counts).all()) and (exact or w["mode"] != 0) and int(absent.sum()) == 0
        assert ok, "sampled lookups disagree with the oracle"
        return {"kmers": int(oc.n_distinct), "present_with_count_ge_sample": True, "equal_to_sample_counts": exact,
                "foreign_kmers_absent": True}

    def measure_k0(hm, layout, st, n_kmers):
        """Random-access roofline measured live on the benchmark table.  K0r = what phase B actually
        does: dependent sector load + atomic, all blocks sweeping the table region by region, at this run's region
        size and touches per sector; naive K0 (uniformly random over the whole table) is kept as context."""
        k0 = {}
        for mode, name in ((0, "atomic_add"), (2, "sector_load_plus_atomic")):
            ms = C.c_float(0)
            ops = min(1 << 32, max(1 << 24, n_kmers // 2))
            for _ in range(2):
                tsx._lib.check(lib.tsxc_k0_random_rmw(hm.handle, layout["table_bytes"], ops, mode, C.byref(ms)), hm.handle)
            k0[name] = ops / ms.value / 1e6
        fine_bits = st["radix_digit1_bits"] + st["radix_digit2_bits"]
        k0r = None
        if fine_bits:
            region = layout["table_bytes"] >> fine_bits
            passes = max(1, -(-n_kmers // max(1, st["chunk_cap_keys"])))
            touches = (n_kmers / passes) / (layout["table_bytes"] / 32)
            total_ops = int(touches * region / 32) * (layout["table_bytes"] // region)
            rates = {}
            # mode 2: sector load + fire-and-forget RED; mode 5: sector load + CAS whose result the thread needs (what an
            # insert does: profiles/r02_k0r_modes.md); 6 resident blocks per SM like k_insert_keys
            for mode, name in ((2, "load_red"), (5, "load_cas")):
                ms = C.c_float(0)
                for _ in range(2):
                    tsx._lib.check(lib.tsxc_k0_region_sweep(hm.handle, layout["table_bytes"], region, int(touches * region / 32), 1024,
                                                            mode | (6 << 8), C.byref(ms)), hm.handle)
                rates[name] = total_ops / ms.value / 1e6
            k0r = {"g_ops_per_s": rates["load_cas"], "load_red_g_ops_per_s": rates["load_red"], "region_bytes": region,
                   "touches_per_sector": touches, "ops_per_item": 1024, "passes_per_step": passes}
        hm.clear(); hm.sync()
        log(f"K0: {k0}  K0r: {k0r}")
        return {"k0": k0, "k0r": k0r}

    def run_e2e(hm, d_packed, n_reads, read_len, n_words, n_kmers, distinct, steps):
        """The same metric through the C ABI with HOST buffers: tsxc_add_reads() from pinned host memory (H2D copies
        inside the timed region) + tsxc_sync() + tsxc_distinct() read-back."""
        import numpy as np
        h_packed, h_off = C.c_void_p(), C.c_void_p()
        tsx._lib.check(lib.tsxc_host_alloc((n_words + 8) * 8, C.byref(h_packed)))
        tsx._lib.check(lib.tsxc_memcpy(dev, h_packed, d_packed, n_words * 8, 2))
        B = args.e2e_batch_reads or (n_reads + 1) // 2
        B = max(32, B + (-B % 32))                              # batches start on a packed-word boundary
        n_batches = (n_reads + B - 1) // B
        tsx._lib.check(lib.tsxc_host_alloc((B + 1) * 8, C.byref(h_off)))
        off = np.ctypeslib.as_array(C.cast(h_off, C.POINTER(C.c_uint64)), shape=(B + 1,))
        off[:] = np.arange(B + 1, dtype=np.uint64) * read_len   # fixed-length reads: every batch has the same offsets
        words_per_batch = B * read_len // 32
        e2e_times, h2d = [], 0
        for it in range(1 + steps):
            hm.clear(); hm.sync()
            t0 = time.perf_counter()
            h2d = 0
            for b in range(n_batches):
                nb = min(B, n_reads - b * B)
                src = C.c_void_p(h_packed.value + b * words_per_batch * 8)
                tsx._lib.check(lib.tsxc_add_reads(hm.handle, src, h_off, nb), hm.handle)
                h2d += ((nb * read_len + 31) // 32) * 8 + (nb + 1) * 8
            hm.sync()
            got = hm.getKmerCount()                                   # 8-byte result read-back
            dt = time.perf_counter() - t0
            assert got == distinct, (got, distinct)
            log(f"e2e pass {it}: {dt * 1e3:.1f} ms")
            if it > 0:
                e2e_times.append(dt)
        lib.tsxc_host_free(h_packed); lib.tsxc_host_free(h_off)
        return {"value": n_kmers / statistics.mean(e2e_times) / 1e9, "unit": "Gk-mer/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 8 + 64, "batches_per_step": n_batches, "ms_per_step": 1e3 * statistics.mean(e2e_times),
                "timing": "wall clock around tsxc_add_reads x batches + tsxc_sync + tsxc_distinct"}

    main_res = run_workload(dict(wl, name=args.workload), args.warmup, args.steps, True)
    variants = {}
    if args.workload == "c2" and not args.no_variants and args.scale == 1.0:
        # the variant BASELINE.json literally names (generateFakeSequences.py style) and the config-5 slice that is the
        # denominator of the multi-GPU efficiency, measured by the same code with no creation flags
        # ... and configs 3 and 4 (k = 63 / 127: the other two key widths, tables of their own).  A variant that fails is
        # reported as such and never costs the headline line.
        for name in ("c2-fakeseq", "c5", "c3", "c4"):
            try:
                r = run_workload(dict(WORKLOADS[name], name=name), 1, 2, False)
                variants[name] = {"value": r["value"], "unit": "Gk-mer/s", "ms_per_step": statistics.mean(r["step_ms"]),
                                  "kmers_per_step": r["n_kmers"], "distinct": r["st"]["distinct"], "phase_ms": r["phase_ms"],
                                  "overflow_entries": r["st"]["overflow_entries"], "workload": WORKLOADS[name]["desc"]}
            except Exception as e:                       # noqa: BLE001 - the line below still has to be printed
                log(f"variant {name} failed: {e!r}")
                variants[name] = {"error": repr(e), "workload": WORKLOADS[name]["desc"]}

    r = main_res
    st, layout, n_kmers, k, l, read_len = r["st"], r["layout"], r["n_kmers"], r["k"], r["l"], r["read_len"]
    T_ms = sum(r["step_ms"])
    ms_per_step = T_ms / args.steps
    value = r["value"]

    # ---- roofline (algorithmic work per k-mer = one RMW of one entry + its share of the input) ----
    E = 8 * layout["entry_words"]
    in_b = 0.25 * read_len / max(1, read_len - k + 1)
    algo_b = 2 * E + in_b
    ph = r["phase_ms"]
    pipeline = (ph["insert"] or 0) > 0
    traffic, traffic_src = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]
        traffic = prof["insert_dram_bytes_per_kmer"] * n_kmers        # ncu dram bytes of the dominant kernel per k-mer
        traffic_src = prof["source"]
    except Exception:
        pass
    if pipeline:
        dom_ms, dom_name, dom_bytes = ph["insert"], "k_insert_keys (phase B: insert in table-region order)", 2 * E
    else:
        dom_ms, dom_name, dom_bytes = st["main_kernel_ms"], "k_count_reads (fused extract+hash+insert)", algo_b
    achieved = n_kmers * dom_bytes / (dom_ms * 1e-3) / 1e9
    step_gbs = n_kmers * algo_b / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": dom_name,
                "per": "step: the dominant kernel's launches of one step together process every k-mer once; algorithmic bytes "
                       "and traffic are per step",
                "algorithmic_bytes_per_kmer": dom_bytes, "kernel_ms_per_step": dom_ms, "share_of_step": dom_ms / ms_per_step,
                "whole_step": {"algorithmic_bytes_per_kmer": algo_b, "achieved": step_gbs, "frac": step_gbs / peak},
                "phase_ms": ph, "chunk_cap_keys": st["chunk_cap_keys"], "page_keys": st["group_cap_keys"]}
    if pipeline:
        # what every kernel of the pipeline streams per k-mer BY DESIGN (not credited as algorithmic work)
        design = {"hist": ("k_count_segs + k_plan_chunks (plan; k_hist_reads in exact mode)", in_b / 2),
                  "part1": ("k_part_reads (S1)", in_b + E), "insert": ("k_build_slices + k_insert_keys (B)", 3 * E)}
        roofline["kernels"] = [{"kernel": design[p][0], "ms_per_step": ph[p], "share_of_step": ph[p] / ms_per_step,
                                "design_bytes_per_kmer": design[p][1],
                                "design_GB_s": n_kmers * design[p][1] / (ph[p] * 1e-3) / 1e9 if ph[p] else None}
                               for p in ("hist", "part1", "insert")]
    k0 = r.get("k0") or {}
    roofline_rand8 = None
    if k0:
        k0r = k0.get("k0r")
        denom = k0r["g_ops_per_s"] if k0r else k0["k0"]["sector_load_plus_atomic"]
        roofline_rand8 = {"achieved": value, "unit": "G RMW/s", "peak": denom, "frac": value / denom,
                          "insert_kernel_alone": {"achieved": n_kmers / (dom_ms * 1e-3) / 1e9, "frac": n_kmers / (dom_ms * 1e-3) / 1e9 / denom},
                          "peak_is": "K0r: dependent sector load + CAS whose result the thread needs, all blocks sweeping the table region by "
                                     "region, measured live at this run's region size and touches per sector (load + fire-and-forget "
                                     "RED, which no insert can use, is reported beside it)" if k0r else "K0 uniform sector load + atomic",
                          "k0r": k0r, "k0_uniform": k0["k0"],
                          "note": "achieved = whole-step k-mers/s (one RMW per k-mer by SURVEY.md §8d) over the measured RMW rate"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_block(wl, args.ref_reads)
        log(f"cpu baseline: {cpu_baseline}")

    line = {
        "metric": "k-mers counted/sec", "value": value, "unit": "Gk-mer/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["desc"], "k": k, "l": l, "reads": r["n_reads"], "read_len": read_len, "kmers_per_step": n_kmers,
                   "distinct": st["distinct"], "load_factor": round(st["used_slots"] / st["n_slots"], 4),
                   "entry_bytes": E, "table_bytes": layout["table_bytes"],
                   "l2": "inputs (2.5 GB) and table (137 GB) far exceed the 126 MB L2; table re-zeroed between steps",
                   "timing": "CUDA events on the handle's stream around the counting kernels; table zeroing untimed",
                   "checks_per_step": "sum of counts == k-mers, no error flags; after the timed steps: sampled lookups vs the oracle"},
        "clear_ms": statistics.mean(r["clear_ms"]),
        "roofline": roofline, "roofline_rand8": roofline_rand8, "cpu_baseline": cpu_baseline, "e2e": r.get("e2e"),
        "variants": variants, "sampled_parity": r.get("sampled"),
        "gpu_launches": r["launches"], "clocks": r["clocks"],
    }
    print(json.dumps(line), flush=True)
    for h in handles.values():
        try:
            h.close()
        except Exception as e:                           # noqa: BLE001 - the result is out; a failed variant may have left its table unusable
            log(f"close failed: {e!r}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
