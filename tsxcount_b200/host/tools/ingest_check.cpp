// ingest_check — host-only exerciser of the feeder (FastxReader + tsxc_pack_reads): prints the number of
// reads, bases, non-ACGT bases, an FNV-1a hash of the concatenated sequences and of the packed stream, and the
// throughput.  Used by tests/test_cli.py (no GPU needed) and for ingest measurements.
//   ingest_check <file> [batch_reads] [block_bytes] [ranges] [parallel]
// The format (FASTQ / multi-line FASTA) is taken from the content; INGEST_PIECE / INGEST_OVERLAP set the FASTA split;
// INGEST_INFLATE_THREADS > 1 lets a BGZF (bgzip) input be inflated by that many threads.
// ranges > 1 reads the file as that many byte ranges (FastxReader's range mode, the CLI's --readers), one after the
// other, which must reproduce the sequential stream exactly (same reads/bases/hash_bases/hash_lens); with a 5th
// argument the ranges are read and packed by one thread each and only the totals and the throughput are printed.
#include <chrono>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <atomic>
#include <vector>

#include "FastxReader.h"
#include "tsxcount_cuda.h"

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: ingest_check <file> [batch_reads] [block_bytes]\n"); return 2; }
    const size_t batch = argc > 2 ? std::strtoull(argv[2], nullptr, 0) : (1u << 18);
    const size_t block = argc > 3 ? std::strtoull(argv[3], nullptr, 0) : (8u << 20);
    const int ranges = argc > 4 ? std::max(1, atoi(argv[4])) : 1;
    const bool parallel = argc > 5;
    uint64_t file_bytes = 0;
    if (ranges > 1) { struct stat sb; if (::stat(argv[1], &sb) != 0) return 2; file_bytes = (uint64_t)sb.st_size; }
    auto lo_of = [&](int i) { return ranges > 1 ? file_bytes / ranges * i : 0; };
    auto hi_of = [&](int i) { return ranges > 1 && i + 1 < ranges ? file_bytes / ranges * (i + 1) : ~0ULL; };
    if (parallel) {
        std::atomic<uint64_t> reads{0}, total{0};
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int i = 0; i < ranges; ++i) th.emplace_back([&, i] {
            FastxReader rd(argv[1], 4, block, lo_of(i), hi_of(i));
            std::string bases; std::vector<uint64_t> offsets, packed, seg;
            for (;;) {
                const size_t n = rd.nextBatch(batch, bases, offsets);
                if (!n) break;
                size_t bad_upper = 0;
                packed.resize(bases.size() / 32 + 2);
                seg.resize(n + 2);
                uint64_t nseg = 0, nbad = 0;
                if (tsxc_pack_reads(bases.data(), offsets.data(), n, packed.data(), seg.data(), seg.size(), &nseg, &nbad) == TSXC_E_INVALID) {
                    for (char c : bases) bad_upper += !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    seg.resize(n + bad_upper + 2);
                    tsxc_pack_reads(bases.data(), offsets.data(), n, packed.data(), seg.data(), seg.size(), &nseg, &nbad);
                }
                reads += n; total += bases.size();
            }
        });
        for (auto& t : th) t.join();
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("reads=%llu bases=%llu\n", (unsigned long long)reads.load(), (unsigned long long)total.load());
        std::fprintf(stderr, "%d readers: %.2f s: %.1f Mbases/s\n", ranges, t, total.load() / (t > 0 ? t : 1) / 1e6);
        return 0;
    }
    std::string bases;
    std::vector<uint64_t> offsets, packed, seg;
    uint64_t reads = 0, total = 0, bad_total = 0, segs = 0, h_bases = 0xcbf29ce484222325ULL, h_lens = h_bases, h_packed = h_bases;
    double t_read = 0, t_pack = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int range = 0; range < ranges; ++range) {
    const int inflate_threads = std::getenv("INGEST_INFLATE_THREADS") ? atoi(std::getenv("INGEST_INFLATE_THREADS")) : 1;
    FastxReader reader(argv[1], 0, block, lo_of(range), hi_of(range), inflate_threads);   // 0: FASTQ or multi-line FASTA by content
    if (inflate_threads > 1) std::fprintf(stderr, "parallel inflate: %s\n", reader.parallelInflate() ? "yes (BGZF)" : "no");
    if (const char* e = std::getenv("INGEST_PIECE")) {                   // multi-line FASTA: piece length, overlap (= k-1)
        const char* o = std::getenv("INGEST_OVERLAP");
        reader.setFastaSplit(std::strtoull(e, nullptr, 0), o ? std::strtoull(o, nullptr, 0) : 0);
    }
    for (;;) {
        const auto a = std::chrono::steady_clock::now();
        const size_t n = reader.nextBatch(batch, bases, offsets);
        const auto b = std::chrono::steady_clock::now();
        if (!n) break;
        packed.resize(bases.size() / 32 + 2);
        seg.resize(n + 2);
        uint64_t nseg = 0, nbad = 0;
        int rc = tsxc_pack_reads(bases.data(), offsets.data(), n, packed.data(), seg.data(), seg.size(), &nseg, &nbad);
        if (rc == TSXC_E_INVALID) {   // non-ACGT bytes split reads into more segments than reads: size for the worst case
            size_t bad_upper = 0;
            for (char c : bases) bad_upper += !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
            seg.resize(n + bad_upper + 2);
            rc = tsxc_pack_reads(bases.data(), offsets.data(), n, packed.data(), seg.data(), seg.size(), &nseg, &nbad);
        }
        if (rc != TSXC_OK) return 1;
        const auto c = std::chrono::steady_clock::now();
        t_read += std::chrono::duration<double>(b - a).count();
        t_pack += std::chrono::duration<double>(c - b).count();
        h_bases = fnv(h_bases, bases.data(), bases.size());
        for (size_t i = 0; i < n; ++i) { const uint64_t len = offsets[i + 1] - offsets[i]; h_lens = fnv(h_lens, &len, 8); }
        h_packed = fnv(h_packed, packed.data(), ((seg[nseg] + 31) / 32) * 8);
        reads += n; total += bases.size(); bad_total += nbad; segs += nseg;
    }
    }
    const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("reads=%llu bases=%llu bad=%llu segments=%llu hash_bases=%016llx hash_lens=%016llx hash_packed=%016llx\n",
                (unsigned long long)reads, (unsigned long long)total, (unsigned long long)bad_total, (unsigned long long)segs,
                (unsigned long long)h_bases, (unsigned long long)h_lens, (unsigned long long)h_packed);
    std::fprintf(stderr, "%.2f s (read %.2f s, pack %.2f s): %.1f Mbases/s\n", t, t_read, t_pack, total / (t > 0 ? t : 1) / 1e6);
    return 0;
}
