// tsx_api.cu — C ABI (include/tsxcount_cuda.h) over the sm_100a kernels.  No CPU fallback: every
// compute entry point needs a CUDA device and reports TSXC_E_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/tsxcount_cuda.h"
#include "tsx_gen.cuh"
#include "tsx_kernels.cuh"

using namespace tsx;

namespace {

thread_local std::string g_create_error;

struct Staging {
    uint64_t* d_packed = nullptr;  size_t cap_packed = 0;   // words
    uint64_t* d_offsets = nullptr; size_t cap_offsets = 0;  // entries
    uint32_t* d_ends = nullptr;    size_t cap_ends = 0;     // words
    cudaEvent_t copied = nullptr, done = nullptr;
    bool used = false;
};

}  // namespace

struct tsxc_table {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D staging
    Layout L{};
    TableView tv{};
    uint64_t* d_words = nullptr;
    unsigned long long* d_ctr = nullptr;
    Staging stage[2];
    int next_stage = 0;
    // device-variant scratch (ends bitmap for caller-resident reads)
    uint32_t* d_ends = nullptr; size_t cap_ends = 0;
    // k-mer / count staging for add_kmers / lookup / dump
    uint64_t* d_keys = nullptr; size_t cap_keys = 0;      // words
    uint64_t* d_counts = nullptr; size_t cap_counts = 0;  // entries
    unsigned long long* d_nout = nullptr;
    // region-partitioned insert (phase A bins)
    uint64_t* d_part = nullptr; size_t cap_part = 0;          // words
    uint64_t* d_spill = nullptr; size_t cap_spill = 0;        // words: (hash, count) records of the single-GPU two-phase path
    unsigned long long* d_cursor = nullptr;                   // kMaxParts + 1 (last = ticket)
    unsigned long long* d_subfill = nullptr; size_t cap_subfill = 0;   // static phase A: fill per (block, bin)
    uint32_t pbits = 0;                                        // log2(#regions); 0 = direct path only
    uint32_t region_log2 = 27;
    bool part_static = false;                                  // EXPERIMENTAL phase A variant (TSXC_PART_STATIC=1 at creation)
    // launch accounting (bench.py's gpu_launches / roofline come from here)
    uint64_t n_launches = 0, n_main_launches = 0;
    double main_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_part, ev_ins;  // per-phase pairs of the two-phase path
    double part_ms = 0.0, ins_ms = 0.0;
    cudaEvent_t marks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    std::mutex mu;
};

namespace {

int fail(tsxc_table* t, int code, const std::string& msg) {
    if (t) t->err = msg; else g_create_error = msg;
    return code;
}

// Event pair around a dominant-kernel launch; resolved lazily in collect_main_ms() after a sync.
bool main_begin(tsxc_table* t, cudaStream_t s, std::pair<cudaEvent_t, cudaEvent_t>* ev) {
    if (!t->ev_free.empty()) { *ev = t->ev_free.back(); t->ev_free.pop_back(); }
    else if (cudaEventCreate(&ev->first) != cudaSuccess || cudaEventCreate(&ev->second) != cudaSuccess) return false;
    return cudaEventRecord(ev->first, s) == cudaSuccess;
}
void main_end(tsxc_table* t, cudaStream_t s, const std::pair<cudaEvent_t, cudaEvent_t>& ev, int launches = 1) {
    cudaEventRecord(ev.second, s);
    t->ev_pending.push_back(ev);
    t->n_main_launches += launches;
}
void collect_main_ms(tsxc_table* t) {  // caller has synchronized the stream
    auto drain = [&](std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v, double& acc) {
        for (auto& ev : v) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) acc += ms;
            t->ev_free.push_back(ev);
        }
        v.clear();
    };
    drain(t->ev_pending, t->main_ms);
    drain(t->ev_part, t->part_ms);
    drain(t->ev_ins, t->ins_ms);
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(t, e_ == cudaErrorMemoryAllocation ? TSXC_E_NOMEM : TSXC_E_CUDA,                 \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                             \
    } while (0)

template <typename T>
int ensure(tsxc_table* t, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return TSXC_OK;
    if (*p) { CU(cudaStreamSynchronize(t->stream)); CU(cudaStreamSynchronize(t->copy_stream)); CU(cudaFree(*p)); *p = nullptr; *cap = 0; }
    const size_t want = std::max(need, *cap + *cap / 2);
    CU(cudaMalloc(p, want * sizeof(T)));
    *cap = want;
    return TSXC_OK;
}

int grid_for(const tsxc_table* t, uint64_t work_items, int per_sm = 8) {
    const uint64_t blocks_needed = (work_items + kBlockThreads - 1) / kBlockThreads;
    const uint64_t cap = (uint64_t)t->sms * per_sm;
    return (int)std::max<uint64_t>(1, std::min(blocks_needed, cap));
}

#define TSX_DISPATCH(L, M)                                      \
    do {                                                        \
        if ((L).KW == 1 && (L).W == 1) { M(1, 1); }             \
        else if ((L).KW == 1 && (L).W == 2) { M(1, 2); }        \
        else if ((L).KW == 2 && (L).W == 2) { M(2, 2); }        \
        else if ((L).KW == 2 && (L).W == 4) { M(2, 4); }        \
        else if ((L).KW == 4 && (L).W == 4) { M(4, 4); }        \
        else return fail(t, TSXC_E_UNSUPPORTED, "no kernel for this entry class"); \
    } while (0)

int status_from_flags(tsxc_table* t, uint64_t flags) {
    if (flags & ERR_TABLE_FULL) return fail(t, TSXC_E_TABLE_FULL, "reprobe limit reached: table full (reference: exit(42))");
    if (flags & ERR_SATURATED) return fail(t, TSXC_E_COUNT_SATURATED, "overflow counter saturated");
    if (flags & ERR_SEND_OVERFLOW) return fail(t, TSXC_E_INVALID, "send buffer of a shard overflowed");
    if (flags & ERR_WRONG_SHARD) return fail(t, TSXC_E_INVALID, "k-mer hash routed to the wrong shard");
    return TSXC_OK;
}

// Geometry of phase A for P bins and chunks of chunk_words packed words.
//   tile   : words a block handles between two run-rotation barriers; longer tiles for many bins, where the
//            barrier (8 warps waiting for the slowest) otherwise shows up as 20 % of the stall samples
//   run    : entries per private run; two runs must cover one tile's arrivals of a bin: mean m = 32*tile/P
//            (every position valid), Poisson tail m + 6*sqrt(m)
//   grid   : 8 blocks per SM for few bins, 4 for many (measured: 3 are resident, a finer grid-stride still
//            balances better: 183-186 ms vs 196 ms at 4 and 217 ms at 3 per SM on config 2; every extra block
//            costs 1.5 runs of holes per bin)
//   cap    : bin capacity = mean + 1 % + 8 sigma + the hole tails of every block (2 runs each) + slack
struct PartGeom { uint32_t tile_words, run; int grid, threads; uint64_t cap; };

PartGeom part_geometry(const tsxc_table* t, uint32_t P, uint64_t chunk_words, double kmers_per_position = 1.0) {
    PartGeom g{};
    // many bins: one fat block per SM keeps the write frontier (one partially filled sector per resident
    // (block, bin)) inside L2; few bins: small blocks, finer grid-stride
    g.threads = P > 1024 ? 1024 : (P > 512 ? 512 : kBlockThreads);
    if (const char* e = std::getenv("TSXC_PART_THREADS")) { const int v = std::atoi(e); if (v == 256 || v == 512 || v == 1024) g.threads = v; }
    uint32_t iters = (P >= 4096 && g.threads == kBlockThreads) ? 4 : 2;
    if (const char* e = std::getenv("TSXC_PART_ITERS")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) iters = (uint32_t)v; }
    int blocks_per_sm = g.threads == 1024 ? 2 : (g.threads == 512 ? 4 : 8);
    if (const char* e = std::getenv("TSXC_PART_GRID")) { const int v = std::atoi(e); if (v >= 1 && v <= 16) blocks_per_sm = v; }
    g.tile_words = (g.threads / 32) * 32 * iters;
    const double m = 32.0 * g.tile_words / P;
    g.run = (uint32_t)std::ceil((m + 6.0 * std::sqrt(m)) / 2.0);
    g.run = (g.run + 3) & ~3u;           // whole 32-byte sectors
    if (g.run < 8) g.run = 8;
    if (const char* e = std::getenv("TSXC_PART_RUN")) { const int v = std::atoi(e); if (v >= 4 && v <= 65536) g.run = (uint32_t)v; }
    const uint64_t tiles = (chunk_words + g.tile_words - 1) / g.tile_words;
    g.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)t->sms * blocks_per_sm));
    const uint64_t mean = (uint64_t)(32.0 * chunk_words * kmers_per_position / P) + 1;
    // hash-uniform bins: sigma = sqrt(mean), so 1 % + 8 sigma on top of the mean is ample
    g.cap = mean + mean / 100 + 8 * (uint64_t)std::sqrt((double)mean) + 2ULL * (uint64_t)g.grid * g.run + 2048;
    g.cap = (g.cap + 7) & ~7ULL;
    return g;
}

uint32_t slice_entries_cfg(uint32_t flags) {
    static const uint32_t v = [] {
        const char* e = std::getenv("TSXC_SLICE");
        const int x = e ? std::atoi(e) : 0;
        return (x >= 256 && x <= (1 << 20)) ? (uint32_t)x : 0u;
    }();
    if (v) return v;
    return (flags & TSXC_FLAG_SKEWED) ? 16 * kSliceEntriesDefault : kSliceEntriesDefault;
}

void launch_partition(uint32_t KW, int threads, int grid, cudaStream_t s, const TableView& tv, const PartView& pv,
                      const uint64_t* d_packed, const uint32_t* d_ends, uint64_t w0, uint64_t w1, uint64_t n_words,
                      uint64_t n_bases) {
    // KW = 4 needs > 64 registers per thread: 1024-thread blocks are not launchable for it
    if (KW == 4 && threads > 512) threads = 512;
#define L_(KW_, T_) k_partition_reads<KW_, T_><<<grid, T_, 0, s>>>(tv, pv, d_packed, d_ends, w0, w1, n_words, n_bases)
    if (KW == 1) { if (threads == 1024) L_(1, 1024); else if (threads == 512) L_(1, 512); else L_(1, 256); }
    else if (KW == 2) { if (threads == 1024) L_(2, 1024); else if (threads == 512) L_(2, 512); else L_(2, 256); }
    else { if (threads == 512) L_(4, 512); else L_(4, 256); }
#undef L_
}

void launch_partition_static(uint32_t KW, int threads, int grid, cudaStream_t s, const TableView& tv, const PartView& pv,
                             const uint64_t* d_packed, const uint32_t* d_ends, uint64_t w0, uint64_t w1, uint64_t n_words,
                             uint64_t n_bases) {
    if (KW == 4 && threads > 512) threads = 512;
#define L_(KW_, T_) k_partition_reads_static<KW_, T_><<<grid, T_, 0, s>>>(tv, pv, d_packed, d_ends, w0, w1, n_words, n_bases)
    if (KW == 1) { if (threads == 1024) L_(1, 1024); else if (threads == 512) L_(1, 512); else L_(1, 256); }
    else if (KW == 2) { if (threads == 1024) L_(2, 1024); else if (threads == 512) L_(2, 512); else L_(2, 256); }
    else { if (threads == 512) L_(4, 512); else L_(4, 256); }
#undef L_
}

// EXPERIMENTAL (TSXC_PART_STATIC=1): slab capacity of the static phase A for a chunk of chunk_words words: the
// positions of the block with the most tiles, spread over P bins, + 6 sigma.
uint64_t static_sub_cap(const PartGeom& g, uint32_t P, uint64_t chunk_words) {
    const uint64_t tiles = (chunk_words + g.tile_words - 1) / g.tile_words;
    const uint64_t tiles_per_block = (tiles + g.grid - 1) / g.grid;
    const double mean = 32.0 * (double)tiles_per_block * g.tile_words / P;
    const uint64_t cap = (uint64_t)(mean + 6.0 * std::sqrt(mean)) + 16;
    return (cap + 3) & ~3ULL;
}

// Two-phase path for tables much larger than the per-SM translation reach (see tsx_kernels.cuh).
int launch_count_reads_partitioned(tsxc_table* t, const uint64_t* d_packed, const uint32_t* d_ends, uint64_t n_words,
                                   uint64_t n_bases, cudaStream_t s) {
    const Layout& L = t->L;
    const uint32_t P = 1u << t->pbits;
    // Chunk = the reads binned before one insert pass.  Larger chunks touch every table region more densely per
    // pass, which phase B turns into DRAM row locality and L2 hits (insert 334 / 290 / 261 ms on config 2 for
    // chunks of 2^24 / 2^25 / 2^26 words), so take the largest chunk (up to 2^27 words) whose bins + spill list fit in free HBM.
    static const int chunk_log2_env = [] { const char* e = std::getenv("TSXC_CHUNK_LOG2"); const int v = e ? std::atoi(e) : 0; return (v >= 16 && v <= 30) ? v : 0; }();
    const bool part_static = t->part_static;
    int chunk_log2 = chunk_log2_env ? chunk_log2_env : 27;
    uint64_t chunk_words = 0, cap = 0, spill_cap = 0, sub_cap = 0;
    PartGeom geo{};
    for (;; --chunk_log2) {
        {   // equal chunks no larger than 2^chunk_log2 / KW words
            const uint64_t max_words = (1ULL << chunk_log2) / L.KW;
            const uint64_t n_chunks = (n_words + max_words - 1) / max_words;
            chunk_words = ((n_words + n_chunks - 1) / n_chunks + 31) & ~31ULL;
        }
        geo = part_geometry(t, P, chunk_words);
        cap = geo.cap;
        if (part_static) {   // bins = grid slabs of sub_cap entries each
            sub_cap = static_sub_cap(geo, P, chunk_words);
            cap = (uint64_t)geo.grid * sub_cap;
        }
        // spill list: one record per 16 positions is far more than homopolymer runs and bin tails ever need;
        // inputs that exceed it (a handful of k-mers making up most of a chunk) are redone by the fused kernel
        spill_cap = std::max<uint64_t>(4096, 32 * chunk_words / 16);
        const size_t need_part = (size_t)P * cap * L.KW, need_spill = (size_t)spill_cap * (L.KW + 1);
        if (chunk_log2_env || chunk_log2 <= 20) break;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); break; }
        const size_t extra = (need_part > t->cap_part ? (need_part - t->cap_part) * 8 : 0) +
                             (need_spill > t->cap_spill ? (need_spill - t->cap_spill) * 8 : 0);
        if (extra + (3ULL << 30) <= free_b) break;   // keep 3 GiB free for staging buffers and the caller
    }
    int rc = ensure(t, &t->d_part, &t->cap_part, (size_t)P * cap * L.KW);
    if (rc) return rc;
    if ((rc = ensure(t, &t->d_spill, &t->cap_spill, (size_t)spill_cap * (L.KW + 1)))) return rc;
    if (part_static && (rc = ensure(t, &t->d_subfill, &t->cap_subfill, (size_t)geo.grid * P))) return rc;
    unsigned long long* ticket = t->d_cursor + kMaxParts;
    unsigned long long* spill_n = t->d_cursor + kMaxParts + 1;
    unsigned int* overflow = reinterpret_cast<unsigned int*>(t->d_cursor + kMaxParts + 2);
    PartView pv{};
    pv.buf = t->d_part; pv.cursor = t->d_cursor; pv.cap = cap;
    pv.pshift = L.LBl - t->pbits; pv.pmask = P - 1; pv.P = P; pv.run = geo.run; pv.tile_words = geo.tile_words;
    pv.spill = t->d_spill; pv.spill_n = spill_n; pv.spill_cap = spill_cap; pv.bins_per_shard_log2 = t->pbits;
    pv.overflow = overflow;
    const uint32_t slice_entries = slice_entries_cfg(t->L.flags);
    uint32_t slices = (uint32_t)((cap + slice_entries - 1) / slice_entries);
    uint32_t n_sources = 1;
    const bool agg = !(L.flags & TSXC_FLAG_NO_WARP_AGG);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, s, &ev);
    int main_launches = 0;
    const PartView pv_bins = pv;   // phase A view; phase B's differs in the static variant
    for (uint64_t w0 = 0; w0 < n_words; w0 += chunk_words) {
        const uint64_t w1 = std::min(n_words, w0 + chunk_words);
        CU(cudaMemsetAsync(t->d_cursor, 0, (kMaxParts + 8) * sizeof(unsigned long long), s));
        const int grid_a = (int)std::max<uint64_t>(1, std::min<uint64_t>((w1 - w0 + geo.tile_words - 1) / geo.tile_words, (uint64_t)geo.grid));
        static const int insert_blocks_per_sm = [] { const char* e = std::getenv("TSXC_INSERT_GRID"); const int v = e ? std::atoi(e) : 0; return (v >= 1 && v <= 16) ? v : 8; }();
        const int grid_b = t->sms * insert_blocks_per_sm;
        std::pair<cudaEvent_t, cudaEvent_t> eva, evb;
        // phase A: bins + spill records, nothing inserted
        const bool ta = main_begin(t, s, &eva);
        if (part_static) {
            PartView pa = pv_bins;                       // [block][bin][sub_cap] slabs, fills at d_subfill[block * P + bin]
            pa.cap = sub_cap; pa.cursor = t->d_subfill;
            launch_partition_static(L.KW, geo.threads, grid_a, s, t->tv, pa, d_packed, d_ends, w0, w1, n_words, n_bases);
            pv = pa;                                     // phase B: every block of phase A is a source
            pv.P = (uint32_t)grid_a * P;
            n_sources = (uint32_t)grid_a;
            slices = (uint32_t)((sub_cap + slice_entries - 1) / slice_entries);
        } else {
            launch_partition(L.KW, geo.threads, grid_a, s, t->tv, pv, d_packed, d_ends, w0, w1, n_words, n_bases);
        }
        if (ta) { cudaEventRecord(eva.second, s); t->ev_part.push_back(eva); }
        // phase B: bins, then the spill records; both skip when the chunk overflowed its spill list ...
        const bool tb = main_begin(t, s, &evb);
#define M(KW_, W_)                                                                                                         \
        if (agg) k_insert_partitions<KW_, W_, true><<<grid_b, kBlockThreads, 0, s>>>(t->tv, pv, slices, slice_entries, n_sources, ticket, overflow);  \
        else k_insert_partitions<KW_, W_, false><<<grid_b, kBlockThreads, 0, s>>>(t->tv, pv, slices, slice_entries, n_sources, ticket, overflow);    \
        k_add_hash_counts<KW_, W_><<<t->sms * 2, kBlockThreads, 0, s>>>(t->tv, t->d_spill, spill_cap, spill_n, overflow);  \
        /* ... in which case the fused kernel redoes the whole chunk (it exits at once otherwise) */                       \
        if (agg) k_count_reads<KW_, W_, true><<<t->sms * 8, kBlockThreads, 0, s>>>(t->tv, d_packed, d_ends, w0, w1, n_words, n_bases, overflow); \
        else k_count_reads<KW_, W_, false><<<t->sms * 8, kBlockThreads, 0, s>>>(t->tv, d_packed, d_ends, w0, w1, n_words, n_bases, overflow)
        TSX_DISPATCH(t->L, M);
#undef M
        if (tb) { cudaEventRecord(evb.second, s); t->ev_ins.push_back(evb); }
        t->n_launches += 4;
        main_launches += 4;
    }
    if (timed) main_end(t, s, ev, main_launches);
    CU(cudaGetLastError());
    return TSXC_OK;
}

int launch_count_reads(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint32_t* d_ends,
                       uint64_t n_reads, uint64_t n_bases, cudaStream_t s) {
    if (n_bases == 0 || n_reads == 0) return TSXC_OK;
    const uint64_t n_words = (n_bases + 31) >> 5;
    CU(cudaMemsetAsync(d_ends, 0, n_words * sizeof(uint32_t), s));
    k_mark_ends<<<grid_for(t, n_reads), kBlockThreads, 0, s>>>(d_offsets, n_reads, d_ends);
    t->n_launches++;
    // the bins pay off once every region receives a few thousand k-mers per chunk
    if (t->pbits > 0 && !(t->L.flags & TSXC_FLAG_DIRECT) && n_words >= (16ULL << t->pbits))
        return launch_count_reads_partitioned(t, d_packed, d_ends, n_words, n_bases, s);
    const int grid = grid_for(t, n_words);
    const bool agg = !(t->L.flags & TSXC_FLAG_NO_WARP_AGG);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, s, &ev);
#define M(KW_, W_)                                                                                                    \
    if (agg) k_count_reads<KW_, W_, true><<<grid, kBlockThreads, 0, s>>>(t->tv, d_packed, d_ends, 0, n_words, n_words, n_bases, nullptr);  \
    else k_count_reads<KW_, W_, false><<<grid, kBlockThreads, 0, s>>>(t->tv, d_packed, d_ends, 0, n_words, n_words, n_bases, nullptr)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    if (timed) main_end(t, s, ev);
    CU(cudaGetLastError());
    return TSXC_OK;
}

int create_impl(uint32_t k, uint32_t l, uint32_t s, int device, uint32_t flags, uint32_t rank, uint32_t n_shards,
                tsxc_table** out) {
    tsxc_table* t = nullptr;  // for CU()/fail() before the handle exists
    if (!out) return fail(t, TSXC_E_INVALID, "out == NULL");
    *out = nullptr;
    if (k < 1 || k > TSXC_MAX_K) return fail(t, TSXC_E_INVALID, "k out of range [1,128]");
    if (2 * k <= l) return fail(t, TSXC_E_INVALID, "Invalid lengths for hashmap size and value of k");  // TSXHashMap.h:93
    Layout L;
    if (!make_layout(k, l, s, flags, rank, n_shards, &L)) return fail(t, TSXC_E_UNSUPPORTED, "(k, l, s, shards) fits no entry class");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(t, TSXC_E_CUDA, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(t, TSXC_E_INVALID, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(t, TSXC_E_CUDA, "device is not sm_100 class");
    CU(cudaSetDevice(device));
    tsxc_table* h = new (std::nothrow) tsxc_table();
    if (!h) return fail(t, TSXC_E_NOMEM, "host allocation failed");
    t = h;
    h->device = device;
    h->sms = prop.multiProcessorCount;
    h->L = L;
    auto bail = [&](int code) { std::string m = h->err; tsxc_destroy(h); g_create_error = m; return code; };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return bail(TSXC_E_CUDA);
    }
    if ((e = cudaMalloc(&h->d_words, L.table_bytes)) != cudaSuccess) {
        h->err = std::string("table allocation of ") + std::to_string(L.table_bytes) + " bytes failed: " + cudaGetErrorString(e);
        cudaGetLastError();
        return bail(TSXC_E_NOMEM);
    }
    if ((e = cudaMalloc(&h->d_ctr, CTR_COUNT * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_nout, sizeof(unsigned long long))) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return bail(TSXC_E_NOMEM);
    }
    for (auto& st : h->stage) {
        if ((e = cudaEventCreateWithFlags(&st.copied, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming)) != cudaSuccess) {
            h->err = cudaGetErrorString(e);
            return bail(TSXC_E_CUDA);
        }
    }
    if ((e = cudaMalloc(&h->d_cursor, (kMaxParts + 8) * sizeof(unsigned long long))) != cudaSuccess) {
        h->err = cudaGetErrorString(e);
        return bail(TSXC_E_NOMEM);
    }
    if (const char* env = std::getenv("TSXC_L2_FETCH")) {  // experiment: L2 fetch granularity hint (32/64/128)
        const int v = std::atoi(env);
        if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
    }
    if (const char* env = std::getenv("TSXC_PART_STATIC")) h->part_static = std::atoi(env) == 1;
    if (const char* env = std::getenv("TSXC_REGION_LOG2")) {
        const int v = std::atoi(env);
        if (v >= 16 && v <= 40) h->region_log2 = (uint32_t)v;
    }
    {   // regions of 2^region_log2 bytes (default 128 MiB); a bucket is 32 bytes
        const uint32_t table_log2 = L.LBl + 5;
        h->pbits = table_log2 > h->region_log2 ? table_log2 - h->region_log2 : 0;
        if (h->pbits > 12) h->pbits = 12;   // kMaxParts bins
        if (h->pbits > L.LBl) h->pbits = L.LBl;
    }
    h->tv = make_view(L, h->d_words, h->d_ctr);
    int rc = tsxc_clear(h);
    if (rc != TSXC_OK) return bail(rc);
    *out = h;
    return TSXC_OK;
}

static int add_keys_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n, bool hashed) {
    if (n == 0) return TSXC_OK;
    const int grid = grid_for(t, n);
    const bool agg = !(t->L.flags & TSXC_FLAG_NO_WARP_AGG);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, t->stream, &ev);
#define M(KW_, W_)                                                                                              \
    if (hashed) { if (agg) k_add_kmers<KW_, W_, true, true><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n);  \
                  else k_add_kmers<KW_, W_, true, false><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n); }   \
    else { if (agg) k_add_kmers<KW_, W_, false, true><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n);        \
           else k_add_kmers<KW_, W_, false, false><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n); }
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    if (timed) main_end(t, t->stream, ev);
    CU(cudaGetLastError());
    return TSXC_OK;
}


// Dump in bucket ranges small enough for the staging buffers; `emit` consumes each chunk on the host.
template <typename Emit>
static int dump_chunks(tsxc_table* t, Emit&& emit) {
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaStreamSynchronize(t->stream));
    const Layout& L = t->L;
    const uint64_t chunk_slots = std::min<uint64_t>(L.n_slots, 1ULL << 24);
    const uint64_t chunk_buckets = std::max<uint64_t>(1, chunk_slots / L.SPB);
    int rc;
    if ((rc = ensure(t, &t->d_keys, &t->cap_keys, (size_t)chunk_slots * L.KW))) return rc;
    if ((rc = ensure(t, &t->d_counts, &t->cap_counts, (size_t)chunk_slots))) return rc;
    std::vector<uint64_t> hk, hc;
    for (uint64_t b0 = 0; b0 < L.n_buckets; b0 += chunk_buckets) {
        const uint64_t b1 = std::min(L.n_buckets, b0 + chunk_buckets);
        CU(cudaMemsetAsync(t->d_nout, 0, sizeof(unsigned long long), t->stream));
        const int grid = grid_for(t, (b1 - b0) * L.SPB);
#define M(KW_, W_) k_dump<KW_, W_><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, b0, b1, t->d_keys, t->d_counts, chunk_slots, t->d_nout)
        TSX_DISPATCH(t->L, M);
#undef M
        t->n_launches++;
        CU(cudaGetLastError());
        unsigned long long n = 0;
        CU(cudaMemcpyAsync(&n, t->d_nout, sizeof n, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaStreamSynchronize(t->stream));
        if (n == 0) continue;
        hk.resize((size_t)n * L.KW); hc.resize((size_t)n);
        CU(cudaMemcpy(hk.data(), t->d_keys, n * L.KW * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(hc.data(), t->d_counts, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        rc = emit(hk.data(), hc.data(), (uint64_t)n);
        if (rc) return rc;
    }
    return TSXC_OK;
}


}  // namespace

extern "C" {

uint32_t tsxc_key_words(uint32_t k) { return (k < 1 || k > TSXC_MAX_K) ? 0 : (k <= 32 ? 1 : (k <= 64 ? 2 : 4)); }
int tsxc_abi_version(void) { return TSXC_ABI_VERSION; }

int tsxc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major >= 10) ++ok;
    }
    return ok;
}

const char* tsxc_status_string(int s) {
    switch (s) {
        case TSXC_OK: return "ok";
        case TSXC_E_INVALID: return "invalid argument";
        case TSXC_E_CUDA: return "CUDA error / no usable device";
        case TSXC_E_NOMEM: return "out of memory";
        case TSXC_E_UNSUPPORTED: return "unsupported (k, l, s)";
        case TSXC_E_COUNT_SATURATED: return "overflow counter saturated";
        case TSXC_E_IO: return "I/O error";
        case TSXC_E_TABLE_FULL: return "table full";
        default: return "unknown status";
    }
}

const char* tsxc_last_error(const tsxc_table* t) { return t ? t->err.c_str() : g_create_error.c_str(); }

int tsxc_create(uint32_t k, uint32_t l, uint32_t s, int device, uint32_t flags, tsxc_table** out) {
    return create_impl(k, l, s, device, flags, 0, 1, out);
}
int tsxc_create_shard(uint32_t k, uint32_t l, uint32_t s, int device, uint32_t flags, uint32_t shard_rank,
                      uint32_t n_shards, tsxc_table** out) {
    return create_impl(k, l, s, device, flags, shard_rank, n_shards, out);
}

int tsxc_destroy(tsxc_table* t) {
    if (!t) return TSXC_OK;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    if (t->copy_stream) cudaStreamSynchronize(t->copy_stream);
    for (auto& st : t->stage) {
        cudaFree(st.d_packed); cudaFree(st.d_offsets); cudaFree(st.d_ends);
        if (st.copied) cudaEventDestroy(st.copied);
        if (st.done) cudaEventDestroy(st.done);
    }
    for (auto* v : {&t->ev_pending, &t->ev_part, &t->ev_ins})
        for (auto& ev : *v) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto& ev : t->ev_free) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto& m : t->marks) if (m) cudaEventDestroy(m);
    cudaFree(t->d_part); cudaFree(t->d_spill); cudaFree(t->d_cursor); cudaFree(t->d_subfill);
    cudaFree(t->d_ends); cudaFree(t->d_keys); cudaFree(t->d_counts); cudaFree(t->d_nout);
    cudaFree(t->d_ctr); cudaFree(t->d_words);
    if (t->stream) cudaStreamDestroy(t->stream);
    if (t->copy_stream) cudaStreamDestroy(t->copy_stream);
    delete t;
    return TSXC_OK;
}

int tsxc_clear(tsxc_table* t) {
    if (!t) return TSXC_E_INVALID;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    CU(cudaMemsetAsync(t->d_words, 0, t->L.table_bytes, t->stream));
    CU(cudaMemsetAsync(t->d_ctr, 0, CTR_COUNT * sizeof(unsigned long long), t->stream));
    t->err.clear();
    for (auto* v : {&t->ev_pending, &t->ev_part, &t->ev_ins}) {
        for (auto& ev : *v) t->ev_free.push_back(ev);
        v->clear();
    }
    t->n_launches = t->n_main_launches = 0; t->main_ms = t->part_ms = t->ins_ms = 0.0;
    return TSXC_OK;
}

void* tsxc_stream(tsxc_table* t) { return t ? (void*)t->stream : nullptr; }

int tsxc_mark(tsxc_table* t, int idx) {
    if (!t || idx < 0 || idx >= 8) return fail(t, TSXC_E_INVALID, "bad mark index");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    if (!t->marks[idx]) CU(cudaEventCreate(&t->marks[idx]));
    CU(cudaEventRecord(t->marks[idx], t->stream));
    return TSXC_OK;
}

int tsxc_mark_elapsed_ms(tsxc_table* t, int a, int b, float* ms_out) {
    if (!t || !ms_out || a < 0 || a >= 8 || b < 0 || b >= 8 || !t->marks[a] || !t->marks[b]) return fail(t, TSXC_E_INVALID, "bad mark index");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaEventElapsedTime(ms_out, t->marks[a], t->marks[b]));
    return TSXC_OK;
}

int tsxc_add_reads_device(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint64_t n_reads,
                          uint64_t n_bases) {
    if (!t || (!d_packed && n_bases) || !d_offsets) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const uint64_t n_words = (n_bases + 31) >> 5;
    int rc = ensure(t, &t->d_ends, &t->cap_ends, (size_t)n_words + 8);
    if (rc) return rc;
    return launch_count_reads(t, d_packed, d_offsets, t->d_ends, n_reads, n_bases, t->stream);
}

int tsxc_add_reads(tsxc_table* t, const uint64_t* packed, const uint64_t* offsets, uint64_t n_reads) {
    if (!t || !offsets) return fail(t, TSXC_E_INVALID, "null argument");
    if (n_reads == 0) return TSXC_OK;
    if (offsets[0] != 0) return fail(t, TSXC_E_INVALID, "offsets[0] must be 0");
    const uint64_t n_bases = offsets[n_reads];
    if (n_bases == 0) return TSXC_OK;
    if (!packed) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const uint64_t n_words = (n_bases + 31) >> 5;
    Staging& st = t->stage[t->next_stage];
    t->next_stage ^= 1;
    // the slot may still be read by the kernel of two calls ago
    if (st.used) CU(cudaStreamWaitEvent(t->copy_stream, st.done, 0));
    int rc;
    if ((rc = ensure(t, &st.d_packed, &st.cap_packed, (size_t)n_words + 8))) return rc;
    if ((rc = ensure(t, &st.d_offsets, &st.cap_offsets, (size_t)n_reads + 1))) return rc;
    if ((rc = ensure(t, &st.d_ends, &st.cap_ends, (size_t)n_words + 8))) return rc;
    CU(cudaMemcpyAsync(st.d_packed, packed, n_words * sizeof(uint64_t), cudaMemcpyHostToDevice, t->copy_stream));
    CU(cudaMemcpyAsync(st.d_offsets, offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, t->copy_stream));
    CU(cudaEventRecord(st.copied, t->copy_stream));
    CU(cudaStreamWaitEvent(t->stream, st.copied, 0));
    rc = launch_count_reads(t, st.d_packed, st.d_offsets, st.d_ends, n_reads, n_bases, t->stream);
    if (rc) return rc;
    CU(cudaEventRecord(st.done, t->stream));
    st.used = true;
    return TSXC_OK;
}

int tsxc_add_kmers_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n) {
    if (!t || (!d_kmers && n)) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    return add_keys_device(t, d_kmers, n, false);
}

int tsxc_add_hashes_device(tsxc_table* t, const uint64_t* d_hashes, uint64_t n) {
    if (!t || (!d_hashes && n)) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    return add_keys_device(t, d_hashes, n, true);
}

int tsxc_add_kmers(tsxc_table* t, const uint64_t* kmers, uint64_t n) {
    if (!t || (!kmers && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const size_t words = (size_t)n * t->L.KW;
    // the staging buffer is shared with other calls: order against the compute stream
    CU(cudaStreamSynchronize(t->stream));
    int rc = ensure(t, &t->d_keys, &t->cap_keys, words);
    if (rc) return rc;
    CU(cudaMemcpyAsync(t->d_keys, kmers, words * sizeof(uint64_t), cudaMemcpyHostToDevice, t->stream));
    return add_keys_device(t, t->d_keys, n, false);
}

int tsxc_sync(tsxc_table* t) {
    if (!t) return TSXC_E_INVALID;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaStreamSynchronize(t->stream));
    unsigned long long flags = 0;
    CU(cudaMemcpy(&flags, t->d_ctr + CTR_ERRORS, sizeof flags, cudaMemcpyDeviceToHost));
    return status_from_flags(t, flags);
}

int tsxc_lookup_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n, uint64_t* d_counts_out) {
    if (!t || ((!d_kmers || !d_counts_out) && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const int grid = grid_for(t, n);
#define M(KW_, W_) k_lookup<KW_, W_><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n, d_counts_out)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_lookup(tsxc_table* t, const uint64_t* kmers, uint64_t n, uint64_t* counts_out) {
    if (!t || ((!kmers || !counts_out) && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    {
        std::lock_guard<std::mutex> g(t->mu);
        CU(cudaSetDevice(t->device));
        CU(cudaStreamSynchronize(t->stream));
        int rc;
        if ((rc = ensure(t, &t->d_keys, &t->cap_keys, (size_t)n * t->L.KW))) return rc;
        if ((rc = ensure(t, &t->d_counts, &t->cap_counts, (size_t)n))) return rc;
        CU(cudaMemcpyAsync(t->d_keys, kmers, n * t->L.KW * sizeof(uint64_t), cudaMemcpyHostToDevice, t->stream));
    }
    int rc = tsxc_lookup_device(t, t->d_keys, n, t->d_counts);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaMemcpyAsync(counts_out, t->d_counts, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return TSXC_OK;
}

int tsxc_stats(tsxc_table* t, tsxc_stats_t* out) {
    if (!t || !out) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaStreamSynchronize(t->stream));
    unsigned long long c[CTR_COUNT];
    CU(cudaMemcpy(c, t->d_ctr, sizeof c, cudaMemcpyDeviceToHost));
    std::memset(out, 0, sizeof *out);
    const Layout& L = t->L;
    out->k = L.k; out->l = L.l; out->s = L.s_req;
    out->key_words = L.KW; out->entry_words = L.W; out->value_bits = L.V; out->quotient_bits = L.Q;
    out->reprobe_bits = L.R; out->slots_per_bucket = L.SPB;
    out->n_shards = 1u << L.shard_bits; out->shard_rank = L.shard_rank;
    out->n_slots = L.n_slots; out->table_bytes = L.table_bytes;
    out->distinct = c[CTR_DISTINCT]; out->overflow_entries = c[CTR_OVERFLOW];
    out->used_slots = c[CTR_DISTINCT] + c[CTR_OVERFLOW];
    out->kmers_added = c[CTR_ADDED]; out->max_reprobe = c[CTR_MAXPROBE]; out->error_flags = c[CTR_ERRORS];
    collect_main_ms(t);
    out->kernel_launches = t->n_launches; out->main_kernel_launches = t->n_main_launches; out->main_kernel_ms = t->main_ms;
    out->partition_ms = t->part_ms; out->insert_ms = t->ins_ms;
    return TSXC_OK;
}

int tsxc_distinct(tsxc_table* t, uint64_t* out) {
    tsxc_stats_t s;
    int rc = tsxc_stats(t, &s);
    if (rc == TSXC_OK && out) *out = s.distinct;
    return rc;
}

int tsxc_dump(tsxc_table* t, uint64_t* kmers_out, uint64_t* counts_out, uint64_t capacity, uint64_t* n_out) {
    if (!t || !n_out || ((!kmers_out || !counts_out) && capacity)) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    uint64_t total = 0;
    const uint32_t KW = t->L.KW;
    int rc = dump_chunks(t, [&](const uint64_t* k, const uint64_t* c, uint64_t n) {
        const uint64_t room = total < capacity ? capacity - total : 0;
        const uint64_t take = std::min(room, n);
        if (take) {
            std::memcpy(kmers_out + total * KW, k, take * KW * sizeof(uint64_t));
            std::memcpy(counts_out + total, c, take * sizeof(uint64_t));
        }
        total += n;
        return 0;
    });
    *n_out = total;
    if (rc) return rc;
    if (total > capacity) return fail(t, TSXC_E_INVALID, "dump truncated: capacity too small");
    return TSXC_OK;
}

int tsxc_dump_file(tsxc_table* t, const char* path) {
    if (!t || !path) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(t, TSXC_E_IO, std::string("cannot open ") + path);
    const uint32_t KW = t->L.KW, k = t->L.k;
    std::vector<char> buf;
    int rc = dump_chunks(t, [&](const uint64_t* keys, const uint64_t* cnt, uint64_t n) {
        static const char LUT[4] = {'A', 'C', 'G', 'T'};  // SequenceUtils.h:65-75
        buf.clear();
        buf.reserve((size_t)n * (k + 22));
        char num[24];
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t* kw = keys + i * KW;
            for (uint32_t b = 0; b < k; ++b) buf.push_back(LUT[(kw[(2 * b) >> 6] >> ((2 * b) & 63)) & 3]);
            buf.push_back('\t');
            int len = std::snprintf(num, sizeof num, "%llu\n", (unsigned long long)cnt[i]);
            buf.insert(buf.end(), num, num + len);
        }
        return std::fwrite(buf.data(), 1, buf.size(), f) == buf.size() ? 0 : (int)TSXC_E_IO;
    });
    if (std::fclose(f) != 0 && rc == TSXC_OK) rc = TSXC_E_IO;
    if (rc == TSXC_E_IO) return fail(t, rc, "write failed");
    return rc;
}

int tsxc_route_layout(tsxc_table* t, uint64_t max_chunk_words, uint32_t kmers_per_position_q16, tsxc_route_layout_t* out) {
    if (!t || !out) return fail(t, TSXC_E_INVALID, "null argument");
    const Layout& L = t->L;
    const uint32_t n_shards = 1u << L.shard_bits;
    // regions of 2^region_log2 bytes inside the shard, limited so that all shards' bins fit one partition kernel
    uint32_t pb = t->pbits;
    while (pb > 0 && ((uint64_t)n_shards << pb) > (uint64_t)kMaxParts) --pb;
    if (n_shards > (uint32_t)kMaxParts) return fail(t, TSXC_E_UNSUPPORTED, "too many shards");
    const uint64_t chunk_words = max_chunk_words ? max_chunk_words : (1ULL << 24) / L.KW;
    const uint32_t P = n_shards << pb;
    const double frac = kmers_per_position_q16 ? std::min(1.0, kmers_per_position_q16 / 65536.0) : 1.0;
    const uint64_t cap = part_geometry(t, P, chunk_words, frac).cap;
    std::memset(out, 0, sizeof *out);
    out->n_shards = n_shards; out->bins_per_shard = 1u << pb; out->key_words = L.KW; out->spill_record_words = L.KW + 1;
    out->chunk_words = chunk_words; out->bin_cap = cap; out->block_words = ((uint64_t)1 << pb) * cap * L.KW;
    out->spill_cap = std::max<uint64_t>(4096, 32 * chunk_words / 32 / n_shards * 2);   // per destination: twice its share of one record per 32 positions
    return TSXC_OK;
}

int tsxc_route_prepare(tsxc_table* t, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases) {
    if (!t || !d_offsets) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const uint64_t n_words = (n_bases + 31) >> 5;
    int rc = ensure(t, &t->d_ends, &t->cap_ends, (size_t)n_words + 8);
    if (rc) return rc;
    CU(cudaMemsetAsync(t->d_ends, 0, (n_words + 8) * sizeof(uint32_t), t->stream));
    if (n_reads) {
        k_mark_ends<<<grid_for(t, n_reads), kBlockThreads, 0, t->stream>>>(d_offsets, n_reads, t->d_ends);
        t->n_launches++;
    }
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_route_chunk(tsxc_table* t, const tsxc_route_layout_t* lay, const uint64_t* d_packed, uint64_t n_bases,
                     uint64_t w_begin, uint64_t w_end, uint64_t* d_bins, unsigned long long* d_cursors,
                     uint64_t* d_spill, unsigned long long* d_spill_n) {
    if (!t || !lay || !d_bins || !d_cursors || !d_spill || !d_spill_n || (!d_packed && w_end > w_begin))
        return fail(t, TSXC_E_INVALID, "null argument");
    if (w_end < w_begin || w_end - w_begin > lay->chunk_words) return fail(t, TSXC_E_INVALID, "chunk larger than the layout allows");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const Layout& L = t->L;
    const uint64_t n_words = (n_bases + 31) >> 5;
    const uint32_t P = lay->n_shards * lay->bins_per_shard;
    uint32_t pb = 0;
    while ((1u << pb) < lay->bins_per_shard) ++pb;
    cudaStream_t s = t->stream;
    CU(cudaMemsetAsync(d_cursors, 0, (size_t)P * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(d_spill_n, 0, (size_t)lay->n_shards * sizeof(unsigned long long), s));
    if (w_end == w_begin) return TSXC_OK;
    const PartGeom geo = part_geometry(t, P, lay->chunk_words);
    PartView pv{};
    pv.buf = d_bins; pv.cursor = d_cursors; pv.cap = lay->bin_cap;
    pv.pshift = L.LBl - pb; pv.pmask = P - 1; pv.P = P;
    pv.run = geo.run; pv.tile_words = geo.tile_words;
    pv.spill = d_spill; pv.spill_n = d_spill_n; pv.spill_cap = lay->spill_cap; pv.bins_per_shard_log2 = pb;
    pv.overflow = reinterpret_cast<unsigned int*>(t->d_cursor + kMaxParts + 3);   // read back by tsxc_route_overflowed
    CU(cudaMemsetAsync(pv.overflow, 0, sizeof(unsigned long long), s));
    const int grid_a = (int)std::max<uint64_t>(1, std::min<uint64_t>((w_end - w_begin + geo.tile_words - 1) / geo.tile_words, (uint64_t)geo.grid));
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, s, &ev);
    launch_partition(L.KW, geo.threads, grid_a, s, t->tv, pv, d_packed, t->d_ends, w_begin, w_end, n_words, n_bases);
    t->n_launches++;
    if (timed) { cudaEventRecord(ev.second, s); t->ev_part.push_back(ev); t->n_main_launches++; }
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_route_overflowed(tsxc_table* t, int* overflowed) {
    if (!t || !overflowed) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    CU(cudaStreamSynchronize(t->stream));
    unsigned long long flag = 0;
    CU(cudaMemcpy(&flag, t->d_cursor + kMaxParts + 3, sizeof flag, cudaMemcpyDeviceToHost));
    *overflowed = flag ? 1 : 0;
    if (flag) CU(cudaMemset(t->d_cursor + kMaxParts + 3, 0, sizeof flag));
    return TSXC_OK;
}

int tsxc_insert_routed(tsxc_table* t, const tsxc_route_layout_t* lay, const uint64_t* d_bins,
                       const unsigned long long* d_cursors, uint32_t n_sources) {
    if (!t || !lay || !d_bins || !d_cursors) return fail(t, TSXC_E_INVALID, "null argument");
    if (n_sources == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    cudaStream_t s = t->stream;
    PartView pv{};
    pv.buf = const_cast<uint64_t*>(d_bins); pv.cursor = const_cast<unsigned long long*>(d_cursors);
    pv.cap = lay->bin_cap; pv.P = n_sources * lay->bins_per_shard;
    const uint32_t slice_entries = slice_entries_cfg(t->L.flags);
    const uint32_t slices = (uint32_t)((lay->bin_cap + slice_entries - 1) / slice_entries);
    unsigned long long* ticket = t->d_cursor + kMaxParts;
    CU(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), s));
    const bool agg = !(t->L.flags & TSXC_FLAG_NO_WARP_AGG);
    const int grid_b = t->sms * 8;
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, s, &ev);
#define M(KW_, W_)                                                                                        \
    if (agg) k_insert_partitions<KW_, W_, true><<<grid_b, kBlockThreads, 0, s>>>(t->tv, pv, slices, slice_entries, n_sources, ticket, nullptr); \
    else k_insert_partitions<KW_, W_, false><<<grid_b, kBlockThreads, 0, s>>>(t->tv, pv, slices, slice_entries, n_sources, ticket, nullptr)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    if (timed) { cudaEventRecord(ev.second, s); t->ev_ins.push_back(ev); t->n_main_launches++; }
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_add_hash_counts_device(tsxc_table* t, const uint64_t* d_records, uint64_t n) {
    if (!t || (!d_records && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const int grid = grid_for(t, n);
#define M(KW_, W_) k_add_hash_counts<KW_, W_><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_records, n, nullptr, nullptr)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_gen_reads_device(const tsxc_gen_params* p, uint64_t first, uint64_t count, int device, void* stream,
                          uint64_t* d_packed, uint64_t* d_offsets) {
    tsxc_table* t = nullptr;
    if (!p || !d_packed || !d_offsets || p->read_len == 0) return fail(t, TSXC_E_INVALID, "null argument");
    if (p->mode > 3) return fail(t, TSXC_E_INVALID, "unknown generator mode");
    GenParams gp{};
    gp.seed = p->seed; gp.n_reads = p->n_reads; gp.read_len = p->read_len; gp.mode = p->mode;
    gp.genome_len = p->genome_len; gp.sub_rate_q16 = p->sub_rate_q16;
    if (p->mode == 2) {
        while ((1ULL << gp.dict_log2) < p->genome_len) ++gp.dict_log2;
        if ((1ULL << gp.dict_log2) != p->genome_len || gp.dict_log2 == 0) return fail(t, TSXC_E_INVALID, "dictionary size must be a power of two > 1");
    }
    if (p->mode == 3 && p->genome_len < p->read_len) return fail(t, TSXC_E_INVALID, "genome shorter than a read");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const uint64_t n_words = (count * p->read_len + 31) >> 5;
    const uint64_t blocks = std::max<uint64_t>(1, std::min<uint64_t>((std::max(n_words, count + 1) + 255) / 256, (uint64_t)sms * 16));
    k_gen_reads<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gp, first, count, d_packed, d_offsets);
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_host_alloc(uint64_t bytes, void** out) {
    tsxc_table* t = nullptr;
    if (!out) return TSXC_E_INVALID;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return TSXC_OK;
}
int tsxc_host_free(void* p) {
    tsxc_table* t = nullptr;
    if (p) CU(cudaFreeHost(p));
    return TSXC_OK;
}
int tsxc_device_alloc(int device, uint64_t bytes, void** out) {
    tsxc_table* t = nullptr;
    if (!out) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(out, bytes ? bytes : 1));
    return TSXC_OK;
}
int tsxc_device_free(int device, void* p) {
    tsxc_table* t = nullptr;
    CU(cudaSetDevice(device));
    if (p) CU(cudaFree(p));
    return TSXC_OK;
}
int tsxc_memcpy(int device, void* dst, const void* src, uint64_t bytes, int kind) {
    tsxc_table* t = nullptr;
    if (kind < 1 || kind > 3) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
    CU(cudaMemcpy(dst, src, bytes, k));
    return TSXC_OK;
}

int tsxc_ipc_export_mem(int device, void* dptr, unsigned char* handle_out) {
    tsxc_table* t = nullptr;
    if (!dptr || !handle_out) return TSXC_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == TSXC_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dptr));
    std::memcpy(handle_out, &h, sizeof h);
    return TSXC_OK;
}
int tsxc_ipc_open_mem(int device, const unsigned char* handle, void** dptr_out) {
    tsxc_table* t = nullptr;
    if (!handle || !dptr_out) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return TSXC_OK;
}
int tsxc_ipc_close_mem(int device, void* dptr) {
    tsxc_table* t = nullptr;
    CU(cudaSetDevice(device));
    if (dptr) CU(cudaIpcCloseMemHandle(dptr));
    return TSXC_OK;
}
int tsxc_ipc_event_create(int device, void** event_out, unsigned char* handle_out) {
    tsxc_table* t = nullptr;
    if (!event_out || !handle_out) return TSXC_E_INVALID;
    static_assert(sizeof(cudaIpcEventHandle_t) == TSXC_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(device));
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
    cudaIpcEventHandle_t h;
    CU(cudaIpcGetEventHandle(&h, ev));
    std::memcpy(handle_out, &h, sizeof h);
    *event_out = (void*)ev;
    return TSXC_OK;
}
int tsxc_ipc_event_open(int device, const unsigned char* handle, void** event_out) {
    tsxc_table* t = nullptr;
    if (!handle || !event_out) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    cudaIpcEventHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    cudaEvent_t ev;
    CU(cudaIpcOpenEventHandle(&ev, h));
    *event_out = (void*)ev;
    return TSXC_OK;
}
int tsxc_event_destroy(int device, void* event) {
    tsxc_table* t = nullptr;
    CU(cudaSetDevice(device));
    if (event) CU(cudaEventDestroy((cudaEvent_t)event));
    return TSXC_OK;
}
int tsxc_event_record(int device, void* event, void* stream) {
    tsxc_table* t = nullptr;
    if (!event) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    CU(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return TSXC_OK;
}
int tsxc_stream_wait_event(int device, void* stream, void* event) {
    tsxc_table* t = nullptr;
    if (!event) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    CU(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return TSXC_OK;
}
int tsxc_copy_async(int device, void* dst, const void* src, uint64_t bytes, void* stream) {
    tsxc_table* t = nullptr;
    if ((!dst || !src) && bytes) return TSXC_E_INVALID;
    CU(cudaSetDevice(device));
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return TSXC_OK;
}

int tsxc_k0_random_rmw(tsxc_table* t, uint64_t table_bytes, uint64_t n_ops, int mode, float* ms_out) {
    if (!t || !ms_out) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    uint64_t bytes = std::min<uint64_t>(table_bytes, t->L.table_bytes);
    uint64_t words = 32;
    while (words * 2 * 8 <= bytes) words *= 2;  // power of two
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, t->stream));
    k_k0_random_rmw<<<t->sms * 8, kBlockThreads, 0, t->stream>>>(t->d_words, words - 1, n_ops, mode);
    CU(cudaEventRecord(b, t->stream));
    CU(cudaEventSynchronize(b));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(ms_out, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return TSXC_OK;
}

int tsxc_k0_windowed(tsxc_table* t, uint64_t footprint_bytes, uint64_t window_bytes, uint64_t n_ops, int mode,
                      int blocks, int threads, float* ms_out) {
    if (!t || !ms_out || blocks < 1 || threads < 32 || threads > 1024) return fail(t, TSXC_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    uint64_t fw = 32, ww = 32;
    while (fw * 2 * 8 <= std::min<uint64_t>(footprint_bytes, t->L.table_bytes)) fw *= 2;
    while (ww * 2 * 8 <= window_bytes && ww * 2 <= fw) ww *= 2;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, t->stream));
    k_k0_windowed<<<blocks, threads, 0, t->stream>>>(t->d_words, fw, ww, n_ops / blocks, mode);
    CU(cudaEventRecord(b, t->stream));
    CU(cudaEventSynchronize(b));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(ms_out, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return TSXC_OK;
}

int tsxc_k0_region_sweep(tsxc_table* t, uint64_t footprint_bytes, uint64_t region_bytes, uint64_t ops_per_region,
                         uint32_t ops_per_item, int mode, float* ms_out) {
    if (!t || !ms_out || !ops_per_item) return fail(t, TSXC_E_INVALID, "bad argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    uint64_t fw = 32, rw = 32;
    while (fw * 2 * 8 <= std::min<uint64_t>(footprint_bytes, t->L.table_bytes)) fw *= 2;
    while (rw * 2 * 8 <= region_bytes && rw * 2 <= fw) rw *= 2;
    unsigned long long* ticket = t->d_cursor + kMaxParts;
    CU(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), t->stream));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, t->stream));
    k_k0_region_sweep<<<t->sms * 8, kBlockThreads, 0, t->stream>>>(t->d_words, rw, fw / rw, ops_per_region, ops_per_item, mode, ticket);
    CU(cudaEventRecord(b, t->stream));
    CU(cudaEventSynchronize(b));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(ms_out, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return TSXC_OK;
}

int tsxc_debug_hash(uint32_t k, const uint64_t* key, uint64_t* out) {
    const uint32_t KW = tsxc_key_words(k);
    if (!KW || !key || !out) return TSXC_E_INVALID;
    const HashParams hp = make_hash_params(k);
    if (KW == 1) { Key<1> x{{key[0]}}; auto h = hash_key<1>(x, hp); out[0] = h.w[0]; }
    else if (KW == 2) { Key<2> x{{key[0], key[1]}}; auto h = hash_key<2>(x, hp); out[0] = h.w[0]; out[1] = h.w[1]; }
    else { Key<4> x{{key[0], key[1], key[2], key[3]}}; auto h = hash_key<4>(x, hp); for (int j = 0; j < 4; ++j) out[j] = h.w[j]; }
    return TSXC_OK;
}

int tsxc_debug_unhash(uint32_t k, const uint64_t* hash, uint64_t* out) {
    const uint32_t KW = tsxc_key_words(k);
    if (!KW || !hash || !out) return TSXC_E_INVALID;
    const HashParams hp = make_hash_params(k);
    if (KW == 1) { Key<1> x{{hash[0]}}; auto h = unhash_key<1>(x, hp); out[0] = h.w[0]; }
    else if (KW == 2) { Key<2> x{{hash[0], hash[1]}}; auto h = unhash_key<2>(x, hp); out[0] = h.w[0]; out[1] = h.w[1]; }
    else { Key<4> x{{hash[0], hash[1], hash[2], hash[3]}}; auto h = unhash_key<4>(x, hp); for (int j = 0; j < 4; ++j) out[j] = h.w[j]; }
    return TSXC_OK;
}

int tsxc_debug_layout(uint32_t k, uint32_t l, uint32_t s, uint32_t flags, uint32_t n_shards, tsxc_stats_t* out) {
    if (!out) return TSXC_E_INVALID;
    if (k < 1 || k > TSXC_MAX_K || 2 * k <= l) return TSXC_E_INVALID;
    Layout L;
    if (!make_layout(k, l, s, flags, 0, n_shards ? n_shards : 1, &L)) return TSXC_E_UNSUPPORTED;
    std::memset(out, 0, sizeof *out);
    out->k = k; out->l = l; out->s = s;
    out->key_words = L.KW; out->entry_words = L.W; out->value_bits = L.V; out->quotient_bits = L.Q;
    out->reprobe_bits = L.R; out->slots_per_bucket = L.SPB; out->n_shards = 1u << L.shard_bits;
    out->n_slots = L.n_slots; out->table_bytes = L.table_bytes;
    return TSXC_OK;
}

}  // extern "C"
