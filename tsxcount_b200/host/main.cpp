// tsxcount — command-line front end with the reference's CLI surface and a new --mode=CUDA.
//
// Mirrors src/mains/main.cpp of mjoppich/tsxCount:
//   options           :30-40   --k/-k --s/-s --l/-l --input/-i --check/-c --checkabort/-a --threads/-t --mode/-m
//   defaults          :409-413 k=14 l=26 s=4
//   parameter echo    :420-427 (same lines on the same streams)
//   count phase       :104-222 ("Added a total of N different kmers")
//   --check           :224-396 (<input>.<k>.count, KMER<TAB>COUNT lines; "total errorsN" and the three counts;
//                               --checkabort -> exit(200), :285-291)
//   statistics        :479     print_stats()
// New: --mode=CUDA (the only mode this binary implements — the CPU modes are the reference's own),
//      --dump=FILE (KMER<TAB>COUNT, format of count_kmers.py:32-34), --device=N, --widevalue (use every
//      spare bit of the entry for the value field instead of exactly s bits), --gpus=N (table hash-sharded over N
//      GPUs, NCCL + peer stores: MultiGpuCounter.h), --n-policy=skip (what happens to k-mers spanning a non-ACGT
//      base; the reference substitutes unseeded random bits, SequenceUtils.h:126-137, which cannot be reproduced).
// The count phase streams FASTQ through a reader thread (FastxReader), packer threads (tsxc_pack_reads into
// pinned buffers) and tsxc_add_reads; see countKMers.
#include <argp.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <exception>
#include <fstream>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

#include "FastxReader.h"
#include "MultiGpuCounter.h"
#include "TSXHashMapCUDA.h"

const char* argp_program_version = "tsxCount-b200 1.0";
static char doc[] = "Count k-mers on a B200 (drop-in --mode=CUDA for tsxCount's hash-map insert path).";
static char args_doc[] = "[FILENAME]...";
static struct argp_option options[] = {
    {"k", 'k', "K", 0, "parameter k"},
    {"s", 's', "STORAGE", 0, "parameter s"},
    {"l", 'l', "L", 0, "parameter l"},
    {"input", 'i', "INPUT_FASTA", 0, "input string"},
    {"check", 'c', 0, OPTION_ARG_OPTIONAL, "check counts"},
    {"checkabort", 'a', 0, OPTION_ARG_OPTIONAL, "abort if check count raises error"},
    {"threads", 't', "THREADS", OPTION_ARG_OPTIONAL, "Number of host threads (1 reader + THREADS-1 packers)."},
    {"mode", 'm', "MODE", OPTION_ARG_OPTIONAL, "counting mode (CUDA)"},
    {"dump", 'd', "FILE", 0, "write KMER<TAB>COUNT lines"},
    {"device", 'g', "N", 0, "CUDA device index"},
    {"widevalue", 'w', 0, 0, "widen the value field to every spare entry bit (default: exactly s bits)"},
    {"readers", 'r', "N", 0, "reader threads on disjoint byte ranges of a plain (not gzip) well-formed input (default 1)"},
    {"gpus", 'G', "N", 0, "shard the table over N GPUs of this box (power of two; default 1)"},
    {"batch-reads", 'B', "N", 0, "reads per GPU and super-batch with --gpus (default 4194304)"},
    {"n-policy", 'N', "POLICY", 0, "k-mers spanning a non-ACGT base: skip (the only policy; default)"},
    {"canonical", 'C', 0, 0, "count canonical k-mers: min(k-mer, reverse complement), like jellyfish -C (default: forward k-mers, as the reference)"},
    {"histogram", 'H', "FILE", 0, "write COUNT<TAB>NUMBER_OF_KMERS lines for counts 1..255 and one line '256+' for the rest"},
    {0}};

struct arguments {
    uint16_t k = 14, l = 26, storagebits = 4;   // main.cpp:409-413
    int threads = 3;   // 1 reader + 2 packers: a reader (~1 Gbases/s) is slower than a packer (~2 with AVX2); more input bandwidth: --readers
    std::string input_path, dump_path, mode = "CUDA";
    bool check = false, checkabort = false, wide = false;
    int device = 0;
    int readers = 1;
    int gpus = 1;
    uint64_t batch_reads = 1u << 22;
    std::string n_policy = "skip";
    bool canonical = false;
    std::string histogram_path;
    uint32_t flags() const { return (wide ? TSXC_FLAG_NONE : TSXC_FLAG_EXACT_S) | (canonical ? TSXC_FLAG_CANONICAL : 0u); }
};

static error_t parse_opt(int key, char* arg, struct argp_state* state) {
    arguments* a = (arguments*)state->input;
    switch (key) {
        case 'k': a->k = (uint16_t)atoi(arg); break;
        case 'l': a->l = (uint16_t)atoi(arg); break;
        case 's': a->storagebits = (uint16_t)atoi(arg); break;
        case 't': if (arg) a->threads = atoi(arg); break;
        case 'i': a->input_path = arg ? arg : ""; break;
        case 'c': a->check = true; break;
        case 'a': a->checkabort = true; break;
        case 'm': if (arg) { a->mode = arg; std::transform(a->mode.begin(), a->mode.end(), a->mode.begin(), ::toupper); } break;
        case 'd': a->dump_path = arg ? arg : ""; break;
        case 'g': a->device = atoi(arg); break;
        case 'w': a->wide = true; break;
        case 'r': a->readers = atoi(arg); break;
        case 'G': a->gpus = atoi(arg); break;
        case 'B': a->batch_reads = std::max<uint64_t>(1, std::strtoull(arg, nullptr, 10)); break;
        case 'N': a->n_policy = arg ? arg : ""; break;
        case 'C': a->canonical = true; break;
        case 'H': a->histogram_path = arg ? arg : ""; break;
        case ARGP_KEY_ARG: return 0;
        default: return ARGP_ERR_UNKNOWN;
    }
    return 0;
}
static struct argp argp_parser = {options, parse_opt, args_doc, doc, 0, 0, 0};

namespace {

struct PinnedBatch {
    uint64_t* packed = nullptr; size_t packed_cap = 0;   // words
    uint64_t* offsets = nullptr; size_t off_cap = 0;     // entries
    void ensure(size_t words, size_t offs) {
        if (words > packed_cap) { if (packed) tsxc_host_free(packed); void* p; if (tsxc_host_alloc(words * 8, &p)) throw TSXException("pinned alloc"); packed = (uint64_t*)p; packed_cap = words; }
        if (offs > off_cap) { if (offsets) tsxc_host_free(offsets); void* p; if (tsxc_host_alloc(offs * 8, &p)) throw TSXException("pinned alloc"); offsets = (uint64_t*)p; off_cap = offs; }
    }
    ~PinnedBatch() { if (packed) tsxc_host_free(packed); if (offsets) tsxc_host_free(offsets); }
};

// Count phase.  The reference has one OpenMP producer that reads 40 records at a time and one task per batch that
// packs and inserts k-mer by k-mer (main.cpp:132-206).  Here: reader threads (FastxReader, ~0.9 Gbases/s each; one
// by default, --readers=N splits a plain file into N byte ranges), `threads`-1 packer threads (tsxc_pack_reads into
// pinned buffers, ~2 Gbases/s each with AVX2) and the GPU behind tsxc_add_reads, connected by two bounded queues.  Batches
// may be submitted in any order (counting commutes).
// A pinned buffer is reused only after a tsxc_sync() that started after its submission returned.
void countKMers(TSXHashMapCUDA& map, const arguments& args) {
    struct Raw { std::string bases; std::vector<uint64_t> offsets; size_t n = 0; };
    enum class St { Free, Filling, Submitted };
    struct Slot { PinnedBatch pin; St st = St::Free; uint64_t epoch = 0; };

    const size_t kBatchReads = 1 << 18;
    const int n_packers = std::max(1, std::min(args.threads > 1 ? args.threads - 1 : 1, 16));
    const int n_slots = 2 * n_packers + 2;
    std::vector<Slot> slots(n_slots);
    std::mutex mu;
    std::condition_variable cv_raw_free, cv_raw_ready, cv_slot;
    std::deque<Raw*> raw_free, raw_ready;
    int n_readers = std::max(1, std::min(args.readers, 64));
    uint64_t file_bytes = 0;
    const bool multi_fasta = FastxReader::sniff(args.input_path) == '>';   // extension: multi-line FASTA input
    if (multi_fasta) n_readers = 1;                  // records are whole chromosomes: one sequential reader
    if (n_readers > 1) {
        struct stat sb;
        if (FastxReader::isGzip(args.input_path) || ::stat(args.input_path.c_str(), &sb) != 0 || sb.st_size < (off_t)(n_readers << 16)) {
            n_readers = 1;   // gzip streams cannot be split; tiny files are not worth it
        } else {
            file_bytes = (uint64_t)sb.st_size;
        }
    }
    std::vector<Raw> raws(n_packers + 2 * n_readers);
    for (auto& r : raws) raw_free.push_back(&r);
    int readers_left = n_readers;
    bool reader_done = false;
    std::exception_ptr failure;
    uint64_t epoch = 0, n_reads_total = 0, n_bad_total = 0;
    bool syncing = false;

    auto reader = [&](int idx) {
        try {
            const uint64_t lo = n_readers > 1 ? file_bytes / n_readers * idx : 0;
            const uint64_t hi = n_readers > 1 && idx + 1 < n_readers ? file_bytes / n_readers * (idx + 1) : ~0ULL;
            // a bgzip (BGZF) input is inflated block-parallel by up to 8 threads; plain gzip is one deflate stream
            FastxReader rd(args.input_path, multi_fasta ? 0 : 4, 8u << 20, lo, hi, std::max(1, std::min(args.threads, 8)));
            if (multi_fasta) rd.setFastaSplit(1u << 20, args.k - 1);   // pieces overlap by k-1 bases: same k-mer multiset
            for (;;) {
                Raw* r;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_raw_free.wait(lk, [&] { return !raw_free.empty() || failure; });
                    if (failure) break;
                    r = raw_free.front(); raw_free.pop_front();
                }
                r->n = rd.nextBatch(kBatchReads, r->bases, r->offsets);
                std::lock_guard<std::mutex> lk(mu);
                if (r->n == 0) { raw_free.push_back(r); cv_raw_free.notify_one(); break; }
                raw_ready.push_back(r);
                cv_raw_ready.notify_one();
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            if (!failure) failure = std::current_exception();
        }
        std::lock_guard<std::mutex> lk(mu);
        if (--readers_left == 0) reader_done = true;
        cv_raw_ready.notify_all();
    };
    std::vector<std::thread> readers;
    for (int i = 0; i < n_readers; ++i) readers.emplace_back(reader, i);

    auto acquire_slot = [&]() -> Slot* {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            if (failure) throw TSXException("aborted: another pipeline thread failed");
            for (auto& sl : slots) if (sl.st == St::Free) { sl.st = St::Filling; return &sl; }
            bool any_submitted = false;
            for (auto& sl : slots) any_submitted |= (sl.st == St::Submitted);
            if (any_submitted && !syncing) {
                // everything submitted before this point is finished once the sync returns
                syncing = true;
                const uint64_t e = epoch++;
                lk.unlock();
                try {
                    map.sync();
                } catch (...) {
                    lk.lock();
                    syncing = false;
                    cv_slot.notify_all();
                    throw;
                }
                lk.lock();
                for (auto& sl : slots) if (sl.st == St::Submitted && sl.epoch <= e) sl.st = St::Free;
                syncing = false;
                cv_slot.notify_all();
                continue;
            }
            cv_slot.wait(lk);
        }
    };

    auto packer = [&] {
        try {
            for (;;) {
                Raw* r;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_raw_ready.wait(lk, [&] { return !raw_ready.empty() || reader_done || failure; });
                    if (failure) return;
                    if (raw_ready.empty()) return;   // reader finished and nothing left
                    r = raw_ready.front(); raw_ready.pop_front();
                }
                Slot* sl = acquire_slot();
                PinnedBatch& pb = sl->pin;
                const size_t n = r->n;
                pb.ensure(r->bases.size() / 32 + 2, n + 2);
                uint64_t nseg = 0, nbad = 0;
                int prc = tsxc_pack_reads(r->bases.data(), r->offsets.data(), n, pb.packed, pb.offsets, pb.off_cap, &nseg, &nbad);
                if (prc == TSXC_E_INVALID) {   // non-ACGT bytes split reads into more segments than reads
                    size_t bad_upper = 0;
                    for (char c : r->bases) bad_upper += !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    pb.ensure(r->bases.size() / 32 + 2, n + bad_upper + 2);
                    prc = tsxc_pack_reads(r->bases.data(), r->offsets.data(), n, pb.packed, pb.offsets, pb.off_cap, &nseg, &nbad);
                }
                if (prc != TSXC_OK) throw TSXException("tsxc_pack_reads failed");
                map.addReads(pb.packed, pb.offsets, nseg);
                std::lock_guard<std::mutex> lk(mu);
                sl->st = St::Submitted; sl->epoch = epoch;
                n_reads_total += n; n_bad_total += nbad;
                raw_free.push_back(r);
                cv_raw_free.notify_one();
                cv_slot.notify_all();
            }
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            if (!failure) failure = std::current_exception();
            cv_raw_free.notify_all(); cv_raw_ready.notify_all(); cv_slot.notify_all();
        }
    };
    std::vector<std::thread> packers;
    for (int i = 0; i < n_packers; ++i) packers.emplace_back(packer);
    for (auto& th : readers) th.join();
    for (auto& th : packers) th.join();
    if (failure) std::rethrow_exception(failure);
    map.sync();
    std::cerr << "Reads: " << n_reads_total << std::endl;
    if (n_bad_total)
        std::cerr << "Non-ACGT bases: " << n_bad_total << " (k-mers spanning them are skipped; the reference substitutes random bits)" << std::endl;
    std::cout << "Added a total of " << map.getKmerCount() << " different kmers" << std::endl;   // main.cpp:222
}

// Count phase with --gpus=N: a super-batch gives every GPU one batch of reads; the batches are packed in parallel
// and counted collectively (MultiGpuCounter::addSuperBatch).  Reading is sequential (one FastxReader).
void countKMersMulti(MultiGpuCounter& mg, const arguments& args) {
    const int N = mg.gpus();
    const bool multi_fasta = FastxReader::sniff(args.input_path) == '>';
    FastxReader rd(args.input_path, multi_fasta ? 0 : 4, 8u << 20, 0, ~0ULL, std::max(1, std::min(args.threads, 8)));
    if (multi_fasta) rd.setFastaSplit(1u << 20, args.k - 1);
    struct Raw { std::string bases; std::vector<uint64_t> offsets; size_t n = 0; };
    std::vector<Raw> raw(N);
    std::vector<PinnedBatch> pin(N);
    std::vector<MultiGpuCounter::HostBatch> hb(N);
    uint64_t n_reads_total = 0, n_bad_total = 0;
    for (;;) {
        size_t got_any = 0;
        for (int d = 0; d < N; ++d) {
            // a batch is read in pieces so that the reader's buffers stay small; pieces are appended
            Raw& r = raw[d];
            r.bases.clear(); r.offsets.assign(1, 0); r.n = 0;
            std::string b; std::vector<uint64_t> o;
            while (r.n < args.batch_reads) {
                const size_t n = rd.nextBatch(std::min<uint64_t>(1u << 18, args.batch_reads - r.n), b, o);
                if (!n) break;
                const uint64_t base = r.bases.size();
                r.bases += b;
                for (size_t i = 1; i <= n; ++i) r.offsets.push_back(base + o[i]);
                r.n += n;
            }
            got_any += r.n;
        }
        if (!got_any) break;
        std::vector<std::thread> th;
        std::vector<uint64_t> nseg(N, 0), nbad(N, 0);
        std::exception_ptr failure;
        std::mutex fmu;
        for (int d = 0; d < N; ++d) th.emplace_back([&, d] {
            try {
                Raw& r = raw[d];
                if (!r.n) return;
                pin[d].ensure(r.bases.size() / 32 + 2, r.n + 2);
                int prc = tsxc_pack_reads(r.bases.data(), r.offsets.data(), r.n, pin[d].packed, pin[d].offsets, pin[d].off_cap, &nseg[d], &nbad[d]);
                if (prc == TSXC_E_INVALID) {
                    size_t bad_upper = 0;
                    for (char c : r.bases) bad_upper += !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    pin[d].ensure(r.bases.size() / 32 + 2, r.n + bad_upper + 2);
                    prc = tsxc_pack_reads(r.bases.data(), r.offsets.data(), r.n, pin[d].packed, pin[d].offsets, pin[d].off_cap, &nseg[d], &nbad[d]);
                }
                if (prc != TSXC_OK) throw TSXException("tsxc_pack_reads failed");
            } catch (...) { std::lock_guard<std::mutex> lk(fmu); if (!failure) failure = std::current_exception(); }
        });
        for (auto& t : th) t.join();
        if (failure) std::rethrow_exception(failure);
        for (int d = 0; d < N; ++d) {
            hb[d].packed = pin[d].packed; hb[d].offsets = pin[d].offsets; hb[d].n_reads = raw[d].n ? nseg[d] : 0;
            n_reads_total += raw[d].n; n_bad_total += nbad[d];
        }
        mg.addSuperBatch(hb);
        mg.sync();                                   // the pinned buffers are refilled by the next super-batch
    }
    std::cerr << "Reads: " << n_reads_total << std::endl;
    if (n_bad_total)
        std::cerr << "Non-ACGT bases: " << n_bad_total << " (k-mers spanning them are skipped; the reference substitutes random bits)" << std::endl;
    std::cout << "Added a total of " << mg.getKmerCount() << " different kmers" << std::endl;   // main.cpp:222
}

// --histogram: COUNT<TAB>NUMBER_OF_DISTINCT_KMERS for every count that occurs, counts above 255 in one last line
template <typename Map>
void writeHistogram(Map& map, const std::string& path) {
    const std::vector<uint64_t> h = map.histogram(257);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    for (uint32_t c = 1; c < 256; ++c) if (h[c]) std::fprintf(f, "%u\t%llu\n", c, (unsigned long long)h[c]);
    if (h[256]) std::fprintf(f, "256+\t%llu\n", (unsigned long long)h[256]);
    std::fclose(f);
}

// KW words -> k-mer text (inverse of encode_kmer; only needed for the error messages of --check)
std::string decode_kmer(const uint64_t* w, uint32_t k) {
    std::string s(k, 'A');
    for (uint32_t i = 0; i < k; ++i) s[i] = "ACGT"[(w[(2 * i) >> 6] >> ((2 * i) & 63)) & 3];
    return s;
}

// --check (main.cpp:224-396): every line KMER<TAB>COUNT of <input>.<k>.count must be in the table with that count, and the
// table must hold nothing else.  The reference parses 100 000 lines at a time into a std::map and probes k-mer by k-mer;
// here the file is read in 64 MiB blocks, the lines of a block are parsed by `threads` workers (no per-line allocation)
// and looked up in one batch, which the library sorts by table region (tsxc_lookup, >= 2^18 queries).
template <typename Map>
int checkCounts(Map& map, const arguments& args) {
    const std::string ref = args.input_path + "." + std::to_string(args.k) + ".count";         // main.cpp:226
    std::cout << "Checking kmer counts against manual hashmap ..." << std::endl;
    std::cerr << "Loading reference file: " << ref << std::endl;
    uint64_t total_errors = 0, ref_count = 0, found = 0;
    const uint32_t kw = map.keyWords();
    const uint32_t k = args.k;
    FILE* file = std::fopen(ref.c_str(), "rb");
    if (file) {
        const size_t kBlock = 64u << 20;
        const int n_workers = std::max(1, std::min(args.threads, 32));
        std::vector<char> buf(kBlock + 4096);
        size_t carry = 0;                                   // bytes of an unfinished line at the start of buf
        struct Part { std::vector<uint64_t> keys, want; uint64_t bad = 0; std::vector<std::string> bad_names; };
        std::vector<Part> parts(n_workers);
        std::vector<uint64_t> keys, want, got;
        auto parse = [&](const char* p, const char* end, Part& out) {
            out.keys.clear(); out.want.clear(); out.bad = 0; out.bad_names.clear();
            while (p < end) {
                const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
                const char* le = nl ? nl : end;
                const char* tab = (const char*)memchr(p, '\t', (size_t)(le - p));
                if (tab) {
                    const size_t base = out.keys.size();
                    out.keys.resize(base + kw, 0);
                    bool ok = (size_t)(tab - p) == k;
                    for (uint32_t i = 0; ok && i < k; ++i) {
                        uint64_t c;
                        switch (p[i]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: c = 0; ok = false; }
                        out.keys[base + ((2 * i) >> 6)] |= c << ((2 * i) & 63);
                    }
                    if (ok) {
                        uint64_t cnt = 0;
                        for (const char* q = tab + 1; q < le && *q >= '0' && *q <= '9'; ++q) cnt = cnt * 10 + (uint64_t)(*q - '0');
                        out.want.push_back(cnt);
                    } else {
                        out.keys.resize(base);
                        ++out.bad;
                        if (out.bad_names.size() < 16) out.bad_names.emplace_back(p, (size_t)(tab - p));
                    }
                }
                p = le + 1;
            }
        };
        bool eof = false;
        while (!eof) {
            const size_t got_bytes = std::fread(buf.data() + carry, 1, kBlock - carry, file);
            size_t n = carry + got_bytes;
            eof = got_bytes == 0;
            if (n == 0) break;
            size_t usable = n;
            if (!eof) {                                     // keep the unfinished last line for the next block
                while (usable > 0 && buf[usable - 1] != '\n') --usable;
                if (usable == 0) { if (n >= kBlock) break; carry = n; continue; }   // a line longer than a block: not a count file
            }
            // cut [0, usable) at line boundaries into one piece per worker
            std::vector<size_t> cut(n_workers + 1, usable);
            cut[0] = 0;
            for (int w = 1; w < n_workers; ++w) {
                size_t c = std::max(cut[w - 1], usable / n_workers * w);
                while (c > 0 && c < usable && buf[c - 1] != '\n') ++c;       // a piece starts right after a newline
                cut[w] = std::min(c, usable);
            }
            std::vector<std::thread> th;
            for (int w = 0; w < n_workers; ++w)
                th.emplace_back([&, w] { parse(buf.data() + cut[w], buf.data() + cut[w + 1], parts[w]); });
            for (auto& t : th) t.join();
            keys.clear(); want.clear();
            for (auto& pt : parts) {
                keys.insert(keys.end(), pt.keys.begin(), pt.keys.end());
                want.insert(want.end(), pt.want.begin(), pt.want.end());
                for (auto& nm : pt.bad_names) std::cout << "kmer: ( " << nm << " ) cannot be encoded" << std::endl;
                total_errors += pt.bad; ref_count += pt.bad;
            }
            if (!want.empty()) {
                std::cout << "Going to check " << want.size() << " kmers" << std::endl;
                got.resize(want.size());
                map.getKmerCounts(keys.data(), want.size(), got.data());
                for (size_t i = 0; i < want.size(); ++i) {
                    if (got[i] != 0) ++found;
                    if (got[i] != want[i]) {                                                    // testExecution.h:50-92
                        std::cout << "kmer: ( " << decode_kmer(keys.data() + i * kw, k) << " ): " << got[i] << " Should be " << want[i] << std::endl;
                        ++total_errors;
                        if (args.checkabort) exit(200);                                         // main.cpp:285-291
                    }
                }
                std::cout << "Checked " << want.size() << " kmers" << std::endl;
                ref_count += want.size();
            }
            carry = n - usable;
            if (carry) std::memmove(buf.data(), buf.data() + usable, carry);
            if (eof) break;
        }
        std::fclose(file);
    }
    const uint64_t distinct = map.getKmerCount();
    std::cout << "total errors" << total_errors << std::endl;                                  // main.cpp:367
    std::cout << "Kmer count check completed." << std::endl;
    std::cout << "Reference kmer count: " << ref_count << std::endl;                            // main.cpp:373-375
    std::cout << "queried kmer count: " << found << std::endl;
    std::cout << "tsxCount kmer count: " << distinct << std::endl;
    // the reference XORs the queried start positions with its k-mer-start bitmap (:378-384): non-zero means the
    // table holds k-mers the reference file does not list
    std::cout << "queried (Xor) kmer count: " << (distinct > found ? distinct - found : found - distinct) << std::endl;
    return total_errors == 0 && distinct == found ? 0 : 1;
}

}  // namespace

int main(int argc, char* argv[]) {
    arguments args;
    argp_parse(&argp_parser, argc, argv, 0, NULL, &args);

    std::cout << "Running with parameters " << std::endl;                                       // main.cpp:420-427
    std::cerr << "K=" << (int)args.k << std::endl;
    std::cerr << "L=" << (int)args.l << std::endl;
    std::cerr << "StorageBits=" << (int)args.storagebits << std::endl;
    std::cerr << "Check=" << (args.check ? "Yes" : "No") << std::endl;
    std::cerr << "Input=" << args.input_path << std::endl;
    std::cerr << "Threads=" << args.threads << std::endl;
    if (args.readers > 1) std::cerr << "Readers=" << args.readers << std::endl;
    std::cerr << "Mode=" << args.mode << std::endl;

    if (args.gpus > 1) std::cerr << "GPUs=" << args.gpus << std::endl;
    if (args.n_policy != "skip") {
        std::cerr << "--n-policy=" << args.n_policy << ": only 'skip' exists (no k-mer spans a non-ACGT base); the reference's random "
                  << "substitution (SequenceUtils.h:126-137) is not reproducible and is not offered." << std::endl;
        return 2;
    }
    if (args.mode != "CUDA") {
        std::cerr << "Mode " << args.mode << " is one of the reference's CPU serialization backends; this binary implements "
                  << "--mode=CUDA only (there is no CPU fallback)." << std::endl;
        return 2;
    }
    try {
        if (args.gpus > 1) {
            std::cerr << "Creating TSXHashMap CUDA, hash-sharded over " << args.gpus << " GPUs" << std::endl;
            // receive buffers: room for what one super-batch can bring to a shard (the batches are hash-uniform)
            const uint64_t recv_cap = args.batch_reads * 400 + (1u << 22);
            MultiGpuCounter mg(args.k, args.l, args.storagebits, args.gpus, args.flags(), recv_cap);
            const auto t0 = std::chrono::steady_clock::now();
            countKMersMulti(mg, args);
            const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const tsxc_stats_t st = mg.stats();
            std::cerr << "Counted " << st.kmers_added << " kmers in " << secs << " s (" << (secs > 0 ? st.kmers_added / secs / 1e6 : 0.0)
                      << " M kmers/s incl. parsing)" << std::endl;
            int rc = 0;
            if (args.check) rc = checkCounts(mg, args);
            if (!args.dump_path.empty()) mg.dump(args.dump_path);
            if (!args.histogram_path.empty()) writeHistogram(mg, args.histogram_path);
            std::cerr << "Used fields: " << st.used_slots << std::endl;                         // print_stats, TSXHashMap.h:390-395
            std::cerr << "Available fields: " << (double)st.n_slots << std::endl;
            std::cerr << "adds: " << st.kmers_added << std::endl;
            std::cerr << "overflow entries: " << st.overflow_entries << std::endl;
            return rc;
        }
        std::cerr << "Creating TSXHashMap CUDA" << std::endl;
        TSXHashMapCUDA map((uint8_t)args.l, args.storagebits, args.k, args.device, args.flags());
        const auto t0 = std::chrono::steady_clock::now();
        countKMers(map, args);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const tsxc_stats_t st = map.stats();
        std::cerr << "Counted " << st.kmers_added << " kmers in " << secs << " s (" << (secs > 0 ? st.kmers_added / secs / 1e6 : 0.0)
                  << " M kmers/s incl. parsing)" << std::endl;
        int rc = 0;
        if (args.check) rc = checkCounts(map, args);
        if (!args.dump_path.empty()) map.dump(args.dump_path);
        if (!args.histogram_path.empty()) writeHistogram(map, args.histogram_path);
        map.print_stats();                                                                      // main.cpp:479
        std::cerr << "adds: " << st.kmers_added << std::endl;
        std::cerr << "overflow entries: " << st.overflow_entries << std::endl;
        return rc;
    } catch (const TSXException& e) {
        std::cerr << "TSXException: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
}
