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
#include "tsx_radix.cuh"

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

enum { PH_HIST = 0, PH_PART1 = 1, PH_PART2 = 2, PH_INSERT = 3, PH_COUNT = 4 };

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
    // host batches of a large table are accumulated on the device (packed words + read-end bitmap) until they fill an
    // insert pass: the pipeline's cost per pass does not depend on how the caller cut its input into batches
    struct Accum {
        uint64_t* d_packed = nullptr; uint32_t* d_ends = nullptr;
        uint64_t fill_words = 0;
        cudaEvent_t consumed = nullptr; bool busy = false;         // the pipeline may still be reading it
    } acc[2];
    int acc_cur = 0;
    uint64_t acc_cap_words = 0;
    cudaEvent_t acc_copied = nullptr;
    // device-variant scratch (ends bitmap for caller-resident reads)
    uint32_t* d_ends = nullptr; size_t cap_ends = 0;
    // k-mer / count staging for add_kmers / lookup / dump
    uint64_t* d_keys = nullptr; size_t cap_keys = 0;      // words
    uint64_t* d_counts = nullptr; size_t cap_counts = 0;  // entries
    unsigned long long* d_nout = nullptr;
    char* d_text = nullptr; size_t cap_text = 0;            // dump lines formatted on the device
    uint64_t* d_pairs[2] = {nullptr, nullptr}; size_t cap_pairs[2] = {0, 0};   // region-sorted lookups: (hash, index) records
    // region-sorted insert pipeline (tsx_radix.cuh): S0 histogram, S1/S2 radix partition, phase B insert
    RadixGeom rg{};                                            // the insert pipeline: one digit of up to 10 bits
    RadixGeom rg_lookup{};                                     // region-sorted lookups: two digits of up to 8 bits
    PageGeom pg{};                                             // how buffer A is addressed (paged pool / exact offsets)
    int part_grid = 0;                                         // thread blocks of S1 the page pool was planned for
    bool radix_on = false;                                     // tables this large take the pipeline by default
    uint32_t region_log2 = 27;                                 // target size of a table region (bytes, log2)
    uint32_t sparse_pct = 70;                                  // S1 walks valid positions only when fewer than this % of the positions start a k-mer
                                                               // (measured: S1 18 % faster at 59 % (config 3), 3 % slower at 80 % (config 2))
    uint16_t* d_page_bin = nullptr; size_t cap_page_bin = 0;   // paged mode: bin of every pool page
    uint16_t* d_page_len = nullptr; size_t cap_page_len = 0;   //             keys in it
    ulonglong2* d_slices = nullptr; size_t cap_slices = 0;     // phase B work items: (first key, keys)
    RadixCtl* d_ctl = nullptr;
    uint32_t* d_seghist = nullptr; size_t cap_seghist = 0;     // S0: counts per (segment, digit 1)
    uint32_t* d_segtotal = nullptr; size_t cap_segtotal = 0;
    uint64_t* d_segprefix = nullptr; size_t cap_segprefix = 0;
    uint64_t* d_A = nullptr; uint64_t cap_A = 0;               // keys, sorted by digit 1 (multi-GPU: the receive buffer)
    uint64_t* d_B = nullptr; uint64_t cap_B = 0;               // two-level mode: the page pool a group of A is sorted into
    bool cap_A_limited = false;                                // cap_A was set by free memory, not by a batch size
    uint32_t* d_fhist = nullptr;                               // kMaxFine
    unsigned long long* d_fcur = nullptr;                      // kMaxFine
    // multi-GPU routing (tsxc_route_*)
    uint64_t** d_peers = nullptr;                              // n_shards receive buffers as seen from this device
    bool peers_set = false;
    const uint64_t* route_packed = nullptr;
    uint64_t route_n_words = 0, route_n_bases = 0;
    uint32_t route_rounds = 0;
    unsigned long long* d_ticket_k0 = nullptr;                 // K0r microbenchmark
    // launch accounting (bench.py's gpu_launches / roofline come from here)
    uint64_t n_launches = 0, n_main_launches = 0;
    double main_ms = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    struct PhaseEv { int ph; std::pair<cudaEvent_t, cudaEvent_t> ev; };
    std::vector<PhaseEv> ev_phase;                             // per-phase pairs of the pipeline
    double phase_ms[PH_COUNT] = {0.0, 0.0, 0.0, 0.0};
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
// One pipeline phase: RAII-free pair, begin/end around the launches of the phase.
struct PhaseTimer {
    tsxc_table* t; cudaStream_t s; int ph; std::pair<cudaEvent_t, cudaEvent_t> ev; bool on;
    PhaseTimer(tsxc_table* t_, cudaStream_t s_, int ph_) : t(t_), s(s_), ph(ph_) { on = main_begin(t, s, &ev); }
    void end(int launches) {
        if (on) { cudaEventRecord(ev.second, s); t->ev_phase.push_back({ph, ev}); }
        t->n_launches += launches;
        t->n_main_launches += launches;
    }
};
void collect_main_ms(tsxc_table* t) {  // caller has synchronized the stream
    for (auto& ev : t->ev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) t->main_ms += ms;
        t->ev_free.push_back(ev);
    }
    t->ev_pending.clear();
    for (auto& pe : t->ev_phase) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pe.ev.first, pe.ev.second) == cudaSuccess) { t->phase_ms[pe.ph] += ms; t->main_ms += ms; }
        t->ev_free.push_back(pe.ev);
    }
    t->ev_phase.clear();
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(t, e_ == cudaErrorMemoryAllocation ? TSXC_E_NOMEM : TSXC_E_CUDA,                 \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                             \
    } while (0)

// The pipeline's key buffers take whatever HBM the table leaves free; anything else that needs memory later
// (staging slots, lookup buffers) may claim it back: the buffers are re-sized at the next batch.
int release_radix_buffers(tsxc_table* t) {
    if (!t->d_A) return TSXC_OK;
    if (t->peers_set) return TSXC_OK;          // exported to peers: must stay where it is
    CU(cudaStreamSynchronize(t->stream));
    CU(cudaFree(t->d_A)); t->d_A = nullptr;
    if (t->d_B) { CU(cudaFree(t->d_B)); t->d_B = nullptr; t->cap_B = 0; }
    if (t->d_page_bin) { CU(cudaFree(t->d_page_bin)); t->d_page_bin = nullptr; t->cap_page_bin = 0; }
    if (t->d_page_len) { CU(cudaFree(t->d_page_len)); t->d_page_len = nullptr; t->cap_page_len = 0; }
    if (t->d_slices) { CU(cudaFree(t->d_slices)); t->d_slices = nullptr; t->cap_slices = 0; }
    t->cap_A = 0; t->cap_A_limited = false;
    return TSXC_OK;
}

// Grows *p to exactly `need` elements (small buffers grow geometrically so that a sequence of slightly larger
// batches does not reallocate every time; nothing above 64 MiB is over-allocated).
template <typename T>
int ensure(tsxc_table* t, T** p, size_t* cap, size_t need) {
    if (need <= *cap) return TSXC_OK;
    if (*p) { CU(cudaStreamSynchronize(t->stream)); CU(cudaStreamSynchronize(t->copy_stream)); CU(cudaFree(*p)); *p = nullptr; *cap = 0; }
    size_t want = need;
    if (need * sizeof(T) <= (64u << 20)) want = std::max(need, *cap + *cap / 2);
    cudaError_t e = cudaMalloc(p, want * sizeof(T));
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        int rc = release_radix_buffers(t);
        if (rc) return rc;
        want = need;
        e = cudaMalloc(p, want * sizeof(T));
    }
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return fail(t, e == cudaErrorMemoryAllocation ? TSXC_E_NOMEM : TSXC_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
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

#define TSX_DISPATCH_KW(L, M)                                   \
    do {                                                        \
        if ((L).KW == 1) { M(1); }                              \
        else if ((L).KW == 2) { M(2); }                         \
        else { M(4); }                                          \
    } while (0)

int status_from_flags(tsxc_table* t, uint64_t flags) {
    if (flags & ERR_TABLE_FULL) return fail(t, TSXC_E_TABLE_FULL, "reprobe limit reached: table full (reference: exit(42))");
    if (flags & ERR_SATURATED) return fail(t, TSXC_E_COUNT_SATURATED, "overflow counter saturated");
    if (flags & ERR_SEND_OVERFLOW) return fail(t, TSXC_E_INVALID, "receive buffer of a shard too small for a routing round");
    if (flags & ERR_WRONG_SHARD) return fail(t, TSXC_E_INVALID, "k-mer hash routed to the wrong shard");
    if (flags & ERR_PLAN) return fail(t, TSXC_E_INVALID, "chunk planner: a segment of the batch exceeds the key buffer");
    return TSXC_OK;
}

// ---- region-sorted pipeline: geometry, buffers, launch sequence -------------------------------------------------

uint64_t env_u64(const char* name, uint64_t dflt) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    return std::strtoull(e, nullptr, 10);
}

// Digits of the insert pipeline: regions of 2^region_log2 bytes of a shard of 2^LBl buckets (32 bytes each), at most
// kNB1 of them.  One shard: one digit.  Hash-sharded: a routing digit (owner + coarse local bits) and a local fine digit.
RadixGeom make_radix_geom(const Layout& L, uint32_t region_log2, uint32_t seg_log2) {
    RadixGeom g{};
    const uint32_t table_log2 = L.LBl + 5;
    uint32_t fb = table_log2 > region_log2 ? table_log2 - region_log2 : 0;   // region bits inside the shard
    fb = std::min<uint32_t>(fb, L.LBl);
    if (L.shard_bits == 0) {
        g.d1 = std::min<uint32_t>(10, fb);
        g.d2 = 0;
    } else {
        // hash-sharded: the routing pass keeps runs of 128 bytes for NVLink (at most 256 bins over all shards), the
        // receiver adds the remaining bits in a second, local pass (tsx_radix.cuh, two-level mode)
        // coarse local bits of the routing digit: the fewer, the longer the runs that cross NVLink (2^c bins per shard).
        // 3 bits = runs of 64-256 k-mers; a routing digit of exactly 6 bits would fall between the cheap ballot ranking
        // (<= 5 bits) and the conflict-detect ranking (>= 7 bits), so 8 shards take 4.  Measured at 2 GPUs: route 82.8 /
        // 75.2 / 72.4 ms for 5 / 3 / 2 bits, step 437 / 429.5 / 430 ms.
        uint32_t coarse = 3;
        if (L.shard_bits + coarse == 6) coarse = 4;
        coarse = (uint32_t)std::min<uint64_t>(8, env_u64("TSXC_ROUTE_COARSE_BITS", coarse));
        g.d1 = std::min<uint32_t>(std::min<uint32_t>(8, L.shard_bits + coarse), L.shard_bits + fb);
        const uint32_t c = g.d1 - L.shard_bits;
        g.d2 = std::min<uint32_t>(std::min<uint32_t>(10 - c, 8), fb - c);    // the fine pass sorts by at most 8 bits
    }
    g.nb1 = 1u << g.d1; g.nb2 = 1u << g.d2; g.nbl = 1u << (g.d1 - L.shard_bits);
    g.shift1 = L.LBg - g.d1; g.shift2 = L.LBg - g.d1 - g.d2;
    g.owner_shift = g.d1 - L.shard_bits;
    g.seg_log2 = seg_log2;
    return g;
}

// Two digits of at most 8 bits for the region-sorted lookups (fine regions of 2^region_log2 bytes).
RadixGeom make_lookup_geom(const Layout& L, uint32_t region_log2, uint32_t seg_log2) {
    RadixGeom g{};
    const uint32_t table_log2 = L.LBl + 5;
    uint32_t fb = table_log2 > region_log2 ? table_log2 - region_log2 : 0;
    fb = std::min<uint32_t>(fb, std::min<uint32_t>(L.LBl, 16));
    g.d1 = std::min<uint32_t>(8, L.shard_bits + fb);
    g.d2 = std::min<uint32_t>(8, fb - (g.d1 - L.shard_bits));
    g.nb1 = 1u << g.d1; g.nb2 = 1u << g.d2; g.nbl = 1u << (g.d1 - L.shard_bits);
    g.shift1 = L.LBg - g.d1; g.shift2 = L.LBg - g.d1 - g.d2;
    g.owner_shift = g.d1 - L.shard_bits;
    g.seg_log2 = seg_log2;
    return g;
}


uint32_t slice_keys_of(const Layout& L) { return kBlockThreads * (L.KW == 1 ? 4u : (L.KW == 2 ? 2u : 1u)); }

// Sizes buffer A for a batch with at most `positions` k-mers: as much as the batch needs, at most what HBM leaves
// free (a denser chunk is a faster insert pass: more touches per table region).  A single shard addresses A as a
// pool of pages when the pool holds enough of them; otherwise (small buffers, multi-GPU receive buffers) bins get
// exact offsets from a histogram pass.
int radix_reserve(tsxc_table* t, uint64_t positions, bool for_peers = false) {
    const RadixGeom& g = t->rg;
    const uint64_t seg_keys = 32ULL << g.seg_log2;
    uint64_t want = std::max<uint64_t>(((positions + seg_keys - 1) / seg_keys) * seg_keys, 2 * seg_keys);
    const uint64_t env_cap = env_u64("TSXC_CHUNK_KEYS", 0);
    if (env_cap) want = std::min(want, std::max(env_cap, 2 * seg_keys));
    if (t->d_A && (t->cap_A >= want || t->cap_A_limited)) return TSXC_OK;
    if (t->peers_set) return fail(t, TSXC_E_INVALID, "receive buffer is exported to peers and cannot grow");
    int rc = release_radix_buffers(t);
    if (rc) return rc;
    if (!t->d_fhist) {
        CU(cudaMalloc(&t->d_fhist, kMaxFine * sizeof(uint32_t)));
        CU(cudaMalloc(&t->d_fcur, kMaxFine * sizeof(unsigned long long)));
    }
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t reserve = env_u64("TSXC_RESERVE_MB", 4096) << 20;     // staging slots, lookups, the caller
    const uint64_t budget = free_b > reserve ? free_b - reserve : free_b / 2;
    const uint32_t slice = slice_keys_of(t->L);
    // per key: the key itself (+ 1/4 in the page pool of the two-level mode), its share of a slice descriptor and of the
    // page records
    const bool two_level = g.d2 > 0;
    const double bytes_per_key = 8.0 * t->L.KW * (two_level ? 1.25 : 1.0) + 20.0 / slice;
    const uint64_t cap_max = (uint64_t)((double)budget / bytes_per_key);
    uint64_t cap = std::min(want, cap_max);
    if (cap < 2 * seg_keys) return fail(t, TSXC_E_NOMEM, "not enough free device memory for the key buffer of the insert pipeline");
    // thread blocks of the partition kernels: two per SM, never more than the batch has segments
    const uint64_t grid_max = std::min<uint64_t>((uint64_t)t->sms * 2, std::max<uint64_t>(1, (positions + seg_keys - 1) / seg_keys));
    // paged addressing: pages of one slice (fewer keys only for tests); every (thread block, bin) may leave a page partly
    // empty, so the pool has to be much larger than that
    PageGeom pg{};
    uint32_t slice_log2 = 0;
    while ((1u << (slice_log2 + 1)) <= slice) ++slice_log2;
    uint32_t pl = slice_log2;
    if (const uint64_t e = env_u64("TSXC_PAGE_LOG2", 0)) pl = (uint32_t)std::min<uint64_t>(slice_log2, std::max<uint64_t>(e, 2));
    uint64_t cap_b = 0;
    if (two_level) {
        // the pool holds a quarter of the receive buffer; without room for four coarse bins' slack the groups would be
        // tiny: then phase B reads the receive buffer directly (coarse regions: slower, still exact)
        cap_b = std::max<uint64_t>(cap / 4, 4 * seg_keys);
        if (const uint64_t e = env_u64("TSXC_POOL_KEYS", 0)) cap_b = e;
        const uint64_t n_pages = std::min<uint64_t>(cap_b >> pl, 0x7fffffffULL);
        if (n_pages >= 4ULL * grid_max * g.nb2 && !env_u64("TSXC_NO_PAGING", 0)) {
            pg.paged = 1; pg.page_log2 = pl; pg.n_pages = (uint32_t)n_pages;
            cap_b = n_pages << pl;
        } else {
            cap_b = 0;
        }
    } else if (t->L.shard_bits == 0 && !for_peers && !env_u64("TSXC_NO_PAGING", 0)) {
        const uint64_t n_pages = std::min<uint64_t>(cap >> pl, 0x7fffffffULL);
        if (n_pages >= 4ULL * grid_max * g.nbl) {
            pg.paged = 1; pg.page_log2 = pl; pg.n_pages = (uint32_t)n_pages;
            cap = n_pages << pl;
        }
    }
    CU(cudaMalloc(&t->d_A, cap * t->L.KW * sizeof(uint64_t)));
    if (cap_b) {
        cudaError_t e = cudaMalloc(&t->d_B, cap_b * t->L.KW * sizeof(uint64_t));
        if (e != cudaSuccess) { cudaGetLastError(); cudaFree(t->d_A); t->d_A = nullptr; return fail(t, TSXC_E_NOMEM, "page pool allocation failed"); }
    }
    const size_t n_slices_max = pg.paged ? (size_t)pg.n_pages : (size_t)(cap / slice) + g.nbl + 1;
    if ((rc = ensure(t, &t->d_slices, &t->cap_slices, n_slices_max))) return rc;
    if (pg.paged && (rc = ensure(t, &t->d_page_bin, &t->cap_page_bin, (size_t)pg.n_pages))) return rc;
    if (pg.paged && (rc = ensure(t, &t->d_page_len, &t->cap_page_len, (size_t)pg.n_pages))) return rc;
    if (!t->d_A) return fail(t, TSXC_E_NOMEM, "not enough free device memory for the tables of the insert pipeline");
    t->pg = pg;
    t->part_grid = (int)grid_max;
    t->cap_A = cap; t->cap_B = cap_b; t->cap_A_limited = cap < want;
    return TSXC_OK;
}

template <int KW> constexpr size_t part_smem_bytes(bool paged, bool sparse) {
    return ((paged ? sizeof(TileSmem<KW, kNB1, true>) : sizeof(TileSmem<KW, kNB1, false>)) + 15) / 16 * 16 + 256 * sizeof(uint64_t*) +
           (paged ? sizeof(PageState) : 0) + (sparse ? sizeof(SparseStage<KW>) : 0);
}
template <int KW> constexpr size_t fine_smem_bytes() {
    return ((sizeof(TileSmem<KW, kNB, true>) + 15) / 16 * 16 + sizeof(ItemFeed<KW>) + 15) / 16 * 16 + sizeof(PageState);
}

// Phase B over the keys currently described by ctl (cur_coff / slicestart): slice descriptors, then the insert.
// Two-level mode (hash-sharded tables): A is cut into groups, every group goes through the fine partition into the page
// pool B first.
int launch_sort_insert(tsxc_table* t, cudaStream_t s) {
    const RadixGeom& g = t->rg;
    const bool agg = !(t->L.flags & TSXC_FLAG_NO_WARP_AGG);
    static const int insert_blocks_per_sm = [] { const int v = (int)env_u64("TSXC_INSERT_GRID", 0); return (v >= 1 && v <= 16) ? v : 6; }();
    const int grid_b = t->sms * insert_blocks_per_sm;
    const uint32_t slice = slice_keys_of(t->L);
#define INS_(KW_, W_, SRC_)                                                                                        \
    if (agg) k_insert_keys<KW_, W_, true><<<grid_b, kBlockThreads, 0, s>>>(t->tv, t->d_ctl, SRC_, t->d_slices);    \
    else k_insert_keys<KW_, W_, false><<<grid_b, kBlockThreads, 0, s>>>(t->tv, t->d_ctl, SRC_, t->d_slices)
    if (g.d2 == 0 || !t->d_B) {
        PhaseTimer pt(t, s, PH_INSERT);
        PageGeom pg = t->pg;
        if (g.d2) pg.paged = 0;                     // two-level geometry without a pool: straight from the receive buffer
        k_build_slices<<<t->sms * 4, kBlockThreads, 0, s>>>(t->d_ctl, pg, t->d_page_bin, t->d_page_len, g.nbl, slice, t->d_slices);
#define M(KW_, W_) INS_(KW_, W_, t->d_A)
        TSX_DISPATCH(t->L, M);
#undef M
        pt.end(2);
        return TSXC_OK;
    }
    const int grid2 = t->part_grid > 0 ? t->part_grid : t->sms * 2;
    // a group ends when the pool could run dry: the slack of one coarse bin is at most a quarter of the pool by construction
    // and a bin cut by a group boundary pays it twice, so every group but the last gets through at least a quarter of the
    // pool's worth of (keys + slack) and this many group slots always suffice (the last slot checks it)
    const uint64_t pages_total = (t->cap_A >> t->pg.page_log2) + (uint64_t)g.nbl * g.nb2 * grid2;
    const uint32_t n_groups = (uint32_t)(pages_total / (t->pg.n_pages / 4) + 2);
    unsigned long long* err = t->d_ctr + CTR_ERRORS;
    for (uint32_t gi = 0; gi < n_groups; ++gi) {
        {
            PhaseTimer pt(t, s, PH_PART2);
#define M(KW_)                                                                                                             \
            k_plan_group_paged<KW_><<<1, kNB, 0, s>>>(t->d_ctl, gi, n_groups, g.nbl, g.nb2, t->pg, (uint32_t)grid2, err);            \
            k_part_keys_paged<KW_><<<grid2, kRadixThreads, fine_smem_bytes<KW_>(), s>>>(t->tv, g, t->pg, t->d_ctl, t->d_A, t->d_B, \
                                                                                        t->d_page_bin, t->d_page_len, err)
            TSX_DISPATCH_KW(t->L, M);
#undef M
            k_chunk_end_paged<<<1, 1024, 0, s>>>(t->d_ctl);
            pt.end(3);
        }
        PhaseTimer pt(t, s, PH_INSERT);
        k_build_slices<<<t->sms * 4, kBlockThreads, 0, s>>>(t->d_ctl, t->pg, t->d_page_bin, t->d_page_len, g.nbl * g.nb2, slice, t->d_slices);
#define M(KW_, W_) INS_(KW_, W_, t->d_B)
        TSX_DISPATCH(t->L, M);
#undef M
        pt.end(2);
    }
#undef INS_
    return TSXC_OK;
}


// The partition kernels use more than the 48 KB of shared memory a kernel gets by default.
int opt_in_shared_memory(tsxc_table* t) {
#define M(KW_)                                                                                                                     \
    CU(cudaFuncSetAttribute(k_hist_reads<KW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHistSmemBytes));                    \
    CU(cudaFuncSetAttribute(k_part_reads<KW_, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes<KW_>(false, false))); \
    CU(cudaFuncSetAttribute(k_part_reads<KW_, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes<KW_>(true, false)));   \
    CU(cudaFuncSetAttribute(k_part_reads<KW_, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes<KW_>(false, true)));   \
    CU(cudaFuncSetAttribute(k_part_reads<KW_, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem_bytes<KW_>(true, true)));     \
    CU(cudaFuncSetAttribute(k_part_keys_paged<KW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_smem_bytes<KW_>()))
    TSX_DISPATCH_KW(t->L, M);
#undef M
    return TSXC_OK;
}

// S1 of chunk c (exact mode: into A or the owners' buffers; paged mode: into the pool).  The dense and the sparse walk
// are both launched; the planner's numbers of the chunk (on the device) decide which of them does the work.
int launch_part_reads(tsxc_table* t, uint32_t c, const uint64_t* d_packed, const uint32_t* d_ends, uint64_t n_words, uint64_t n_bases,
                      uint64_t seg0, uint64_t* const* d_peers, cudaStream_t s) {
    const RadixGeom& g = t->rg;
    const int grid_p = t->part_grid > 0 ? t->part_grid : t->sms * 2;
    unsigned long long* err = t->d_ctr + CTR_ERRORS;
    const uint32_t pct = t->sparse_pct;
#define PART_(KW_, PAGED_, SPARSE_, PEERS_, PB_, PL_)                                                                                   \
    k_part_reads<KW_, PAGED_, SPARSE_><<<grid_p, kRadixThreads, part_smem_bytes<KW_>(PAGED_, SPARSE_), s>>>(                           \
        t->tv, g, t->pg, t->d_ctl, c, d_packed, d_ends, n_words, n_bases, seg0, t->d_A, PEERS_, PB_, PL_, err, pct)
#define M(KW_)                                                                                                                          \
    if (t->pg.paged && !d_peers) {                                                                                                       \
        PART_(KW_, true, false, nullptr, t->d_page_bin, t->d_page_len);                                                                  \
        PART_(KW_, true, true, nullptr, t->d_page_bin, t->d_page_len);                                                                   \
    } else {                                                                                                                             \
        PART_(KW_, false, false, d_peers, nullptr, nullptr);                                                                             \
        PART_(KW_, false, true, d_peers, nullptr, nullptr);                                                                              \
    }
    TSX_DISPATCH_KW(t->L, M);
#undef M
#undef PART_
    return TSXC_OK;
}

// Planner over segments [seg0, seg0 + n_segs) of the batch.  exact: S0 histogram (bins get exact offsets);
// otherwise only the k-mers per segment are counted (paged mode).
int launch_hist_plan(tsxc_table* t, const uint64_t* d_packed, const uint32_t* d_ends, uint64_t n_words, uint64_t n_bases,
                     uint64_t seg0, uint32_t n_segs, uint64_t cap, bool exact, cudaStream_t s) {
    const RadixGeom& g = t->rg;
    int rc;
    if (exact && (rc = ensure(t, &t->d_seghist, &t->cap_seghist, (size_t)n_segs * g.nb1))) return rc;
    if ((rc = ensure(t, &t->d_segtotal, &t->cap_segtotal, (size_t)n_segs))) return rc;
    if ((rc = ensure(t, &t->d_segprefix, &t->cap_segprefix, (size_t)n_segs + 1))) return rc;
    PhaseTimer pt(t, s, PH_HIST);
    const int grid = (int)std::min<uint64_t>(n_segs, (uint64_t)t->sms * 2);
    if (exact) {
#define M(KW_) k_hist_reads<KW_><<<grid, kRadixThreads, kHistSmemBytes, s>>>(t->tv, g, d_packed, d_ends, n_words, n_bases, seg0, n_segs, t->d_seghist, t->d_segtotal)
        TSX_DISPATCH_KW(t->L, M);
#undef M
    } else {
#define M(KW_) k_count_segs<KW_><<<grid, kRadixThreads, 0, s>>>(d_ends, n_words, n_bases, t->L.k, g.seg_log2, seg0, n_segs, t->d_segtotal)
        TSX_DISPATCH_KW(t->L, M);
#undef M
    }
    k_plan_chunks<<<1, 1024, 0, s>>>(t->d_ctl, exact ? t->d_seghist : nullptr, t->d_segtotal, t->d_segprefix, n_segs, g.nb1, cap,
                                     32ULL << g.seg_log2, t->d_ctr + CTR_ERRORS);
    pt.end(2);
    return TSXC_OK;
}

// How many segments one planner run may cover so that it never needs more than kMaxChunks chunks, and the
// number of chunk slots to launch for n_segs segments.  The planner evens the chunks out, which at worst halves
// them: a chunk holds at least fit/2 whole segments (fit = segments that always fit the key buffer).
void plan_bounds(const RadixGeom& g, uint64_t cap, uint64_t* segs_per_run, uint32_t n_segs, uint32_t* max_chunks) {
    const uint64_t seg_keys = 32ULL << g.seg_log2;
    const uint64_t fit = std::max<uint64_t>(2, cap / seg_keys);
    if (segs_per_run) *segs_per_run = std::max<uint64_t>(1, (uint64_t)(kMaxChunks - 2) * (fit / 2));
    if (max_chunks) *max_chunks = (uint32_t)std::min<uint64_t>(kMaxChunks, (uint64_t)n_segs / (fit / 2) + 2);
}

int launch_count_reads_radix(tsxc_table* t, const uint64_t* d_packed, const uint32_t* d_ends, uint64_t n_words,
                             uint64_t n_bases, cudaStream_t s) {
    const RadixGeom& g = t->rg;
    int rc = radix_reserve(t, n_words * 32);
    if (rc) return rc;
    const bool paged = t->pg.paged != 0;
    // paged: every (thread block of S1, bin) may leave its last page partly empty
    const uint64_t cap = paged ? (uint64_t)(t->pg.n_pages - (uint64_t)t->part_grid * g.nbl) << t->pg.page_log2 : t->cap_A;
    const uint32_t slice = slice_keys_of(t->L);
    const uint64_t seg_words = 1ULL << g.seg_log2;
    const uint64_t n_segs_total = (n_words + seg_words - 1) / seg_words;
    uint64_t segs_per_run = 0;
    plan_bounds(g, cap, &segs_per_run, 0, nullptr);
    for (uint64_t seg0 = 0; seg0 < n_segs_total; seg0 += segs_per_run) {
        const uint32_t n_segs = (uint32_t)std::min<uint64_t>(segs_per_run, n_segs_total - seg0);
        if ((rc = launch_hist_plan(t, d_packed, d_ends, n_words, n_bases, seg0, n_segs, cap, !paged, s))) return rc;
        uint32_t max_chunks = 0;
        plan_bounds(g, cap, nullptr, n_segs, &max_chunks);
        for (uint32_t c = 0; c < max_chunks; ++c) {
            {
                PhaseTimer pt(t, s, PH_PART1);
                if (paged) {
                    k_chunk_begin_paged<<<1, 1024, 0, s>>>(t->d_ctl, c);
                } else {
                    k_chunk_begin<<<1, 1024, 0, s>>>(t->d_ctl, c, slice);
                }
                if ((rc = launch_part_reads(t, c, d_packed, d_ends, n_words, n_bases, seg0, nullptr, s))) return rc;
                if (paged) k_chunk_end_paged<<<1, 1024, 0, s>>>(t->d_ctl);
                pt.end(paged ? 4 : 3);
            }
            if ((rc = launch_sort_insert(t, s))) return rc;
        }
    }
    CU(cudaGetLastError());
    return TSXC_OK;
}

// Counts the k-mers of a packed stream whose read-end bitmap is already in place.
int launch_count_marked(tsxc_table* t, const uint64_t* d_packed, const uint32_t* d_ends, uint64_t n_words, uint64_t n_bases,
                        cudaStream_t s) {
    // the pipeline pays off once every table region receives a few thousand k-mers per pass
    // (a shard of a hash-sharded table is fed through tsxc_route_*; reads handed to it directly take the fused kernel)
    if (t->radix_on && t->L.shard_bits == 0 && !t->peers_set && !(t->L.flags & TSXC_FLAG_DIRECT) && n_words >= (uint64_t)t->rg.nbl * 4)
        return launch_count_reads_radix(t, d_packed, d_ends, n_words, n_bases, s);
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

int launch_count_reads(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint32_t* d_ends,
                       uint64_t n_reads, uint64_t n_bases, cudaStream_t s) {
    if (n_bases == 0 || n_reads == 0) return TSXC_OK;
    const uint64_t n_words = (n_bases + 31) >> 5;
    CU(cudaMemsetAsync(d_ends, 0, n_words * sizeof(uint32_t), s));
    k_mark_ends<<<grid_for(t, n_reads), kBlockThreads, 0, s>>>(d_offsets, n_reads, d_ends, 0);
    t->n_launches++;
    return launch_count_marked(t, d_packed, d_ends, n_words, n_bases, s);
}

// ---- accumulation of host batches (see tsxc_table::Accum) ---------------------------------------------------------
int ensure_acc(tsxc_table* t) {
    if (t->acc[0].d_packed) return TSXC_OK;
    int rc = release_radix_buffers(t);            // plan the memory again with the accumulation buffers in place
    if (rc) return rc;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t reserve = env_u64("TSXC_RESERVE_MB", 4096) << 20;
    const uint64_t budget = free_b > reserve ? free_b - reserve : free_b / 2;
    // per base position: a key in A if it starts a k-mer, 2 bits + 1 end bit in each of the two buffers
    const double per_pos = 8.0 * t->L.KW + 0.1 + 0.75;
    uint64_t pos = (uint64_t)((double)budget / per_pos);
    pos = std::min<uint64_t>(pos, std::max<uint64_t>(1ULL << 22, 4 * t->L.n_slots));   // more than ~4 k-mers per slot per pass is pointless
    uint64_t words = std::min<uint64_t>(pos / 32, 1ULL << 28);
    if (const uint64_t e = env_u64("TSXC_ACC_WORDS", 0)) words = e;
    words = std::max<uint64_t>(words & ~511ULL, 1024);
    for (auto& a : t->acc) {
        CU(cudaMalloc(&a.d_packed, (words + 8) * sizeof(uint64_t)));
        CU(cudaMalloc(&a.d_ends, (words + 8) * sizeof(uint32_t)));
        if (!a.consumed) CU(cudaEventCreateWithFlags(&a.consumed, cudaEventDisableTiming));
        a.fill_words = 0; a.busy = false;
    }
    if (!t->acc_copied) CU(cudaEventCreateWithFlags(&t->acc_copied, cudaEventDisableTiming));
    t->acc_cap_words = words;
    return TSXC_OK;
}

// Hand the accumulated reads to the counting kernels (asynchronous) and switch to the other buffer.
int flush_acc(tsxc_table* t) {
    if (!t->acc_cap_words) return TSXC_OK;
    tsxc_table::Accum& a = t->acc[t->acc_cur];
    if (a.fill_words == 0) return TSXC_OK;
    CU(cudaEventRecord(t->acc_copied, t->copy_stream));
    CU(cudaStreamWaitEvent(t->stream, t->acc_copied, 0));
    // padding between batches and after the last one is marked as read ends, so the stream simply ends at a word boundary
    int rc = launch_count_marked(t, a.d_packed, a.d_ends, a.fill_words, a.fill_words * 32, t->stream);
    if (rc) return rc;
    CU(cudaEventRecord(a.consumed, t->stream));
    a.busy = true;
    a.fill_words = 0;
    t->acc_cur ^= 1;
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
    if (n_shards > (uint32_t)kNB) return fail(t, TSXC_E_UNSUPPORTED, "at most 256 shards");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(t, TSXC_E_CUDA, "no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(t, TSXC_E_INVALID, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(t, TSXC_E_CUDA, "device is not sm_100 class");
    CU(cudaSetDevice(device));
    if (const uint64_t fetch = env_u64("TSXC_L2_FETCH", 0)) {     // experiment: 32 / 64 / 128 bytes per L2 miss (a hint to the driver)
        if (fetch == 32 || fetch == 64 || fetch == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch);
        cudaGetLastError();
    }
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
        (e = cudaMalloc(&h->d_nout, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_ctl, sizeof(RadixCtl))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_ticket_k0, sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaMalloc(&h->d_peers, kNB * sizeof(uint64_t*))) != cudaSuccess) {
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
    // Table regions of 2^region_log2 bytes (default 128 MiB: profiles/r02_pipeline_history.md); tables of 512 MiB and
    // more take the region-sorted pipeline.  TSXC_REGION_LOG2 overrides both (tests force tiny tables through it).
    uint32_t min_table_log2 = 29;
    if (const char* env = std::getenv("TSXC_REGION_LOG2")) {
        const int v = std::atoi(env);
        if (v >= 6 && v <= 40) { h->region_log2 = (uint32_t)v; min_table_log2 = (uint32_t)v + 1; }
    }
    uint32_t seg_log2 = (uint32_t)env_u64("TSXC_SEG_LOG2", kSegWordsLog2Default);
    seg_log2 = std::max<uint32_t>(kSegWordsLog2Min, std::min<uint32_t>(seg_log2, 20));
    h->rg = make_radix_geom(L, h->region_log2, seg_log2);
    h->rg_lookup = make_lookup_geom(L, std::min<uint32_t>(h->region_log2, 24), seg_log2);
    h->radix_on = (L.LBl + 5 >= min_table_log2) && (h->rg.d1 > L.shard_bits);
    h->sparse_pct = (uint32_t)std::min<uint64_t>(env_u64("TSXC_SPARSE_PCT", h->sparse_pct), 101);   // 0: never, 101: always
    { const int arc = opt_in_shared_memory(h); if (arc != TSXC_OK) return bail(arc); }
    h->tv = make_view(L, h->d_words, h->d_ctr);
    CU(cudaMemsetAsync(h->d_ctl, 0, sizeof(RadixCtl), h->stream));
    int rc = tsxc_clear(h);
    if (rc != TSXC_OK) return bail(rc);
    *out = h;
    return TSXC_OK;
}

static int add_keys_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n) {
    if (n == 0) return TSXC_OK;
    const int grid = grid_for(t, n);
    const bool agg = !(t->L.flags & TSXC_FLAG_NO_WARP_AGG);
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    const bool timed = main_begin(t, t->stream, &ev);
#define M(KW_, W_)                                                                                      \
    if (agg) k_add_kmers<KW_, W_, true><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n);      \
    else k_add_kmers<KW_, W_, false><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, d_kmers, n)
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
    { const int frc = flush_acc(t); if (frc) return frc; }
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
    for (auto& ev : t->ev_pending) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto& pe : t->ev_phase) { cudaEventDestroy(pe.ev.first); cudaEventDestroy(pe.ev.second); }
    for (auto& ev : t->ev_free) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (auto& m : t->marks) if (m) cudaEventDestroy(m);
    for (auto& a : t->acc) { cudaFree(a.d_packed); cudaFree(a.d_ends); if (a.consumed) cudaEventDestroy(a.consumed); }
    if (t->acc_copied) cudaEventDestroy(t->acc_copied);
    cudaFree(t->d_A); cudaFree(t->d_B); cudaFree(t->d_page_bin); cudaFree(t->d_page_len); cudaFree(t->d_slices); cudaFree(t->d_ctl); cudaFree(t->d_seghist); cudaFree(t->d_segtotal); cudaFree(t->d_segprefix);
    cudaFree(t->d_fhist); cudaFree(t->d_fcur); cudaFree(t->d_peers); cudaFree(t->d_ticket_k0);
    cudaFree(t->d_ends); cudaFree(t->d_keys); cudaFree(t->d_counts); cudaFree(t->d_nout); cudaFree(t->d_text); cudaFree(t->d_pairs[0]); cudaFree(t->d_pairs[1]);
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
    for (auto& a : t->acc) a.fill_words = 0;          // reads accumulated but not yet counted are dropped with the table
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaMemsetAsync(t->d_words, 0, t->L.table_bytes, t->stream));
    CU(cudaMemsetAsync(t->d_ctr, 0, CTR_COUNT * sizeof(unsigned long long), t->stream));
    t->err.clear();
    for (auto& ev : t->ev_pending) t->ev_free.push_back(ev);
    t->ev_pending.clear();
    for (auto& pe : t->ev_phase) t->ev_free.push_back(pe.ev);
    t->ev_phase.clear();
    t->n_launches = t->n_main_launches = 0; t->main_ms = 0.0;
    for (double& m : t->phase_ms) m = 0.0;
    return TSXC_OK;
}

int tsxc_trim(tsxc_table* t) {
    if (!t) return TSXC_E_INVALID;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    int rc = release_radix_buffers(t);
    if (rc) return rc;
    if (t->acc_cap_words) {
        CU(cudaStreamSynchronize(t->stream));
        CU(cudaStreamSynchronize(t->copy_stream));
        for (auto& a : t->acc) { cudaFree(a.d_packed); cudaFree(a.d_ends); a.d_packed = nullptr; a.d_ends = nullptr; a.fill_words = 0; a.busy = false; }
        t->acc_cap_words = 0;
    }
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
    if (!t || !ms_out || a < 0 || a >= 8 || b < 0 || b >= 8) return fail(t, TSXC_E_INVALID, "bad mark index");
    std::lock_guard<std::mutex> g(t->mu);
    if (!t->marks[a] || !t->marks[b]) return fail(t, TSXC_E_INVALID, "mark not recorded");
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
    int rc;
    if (t->radix_on && !(t->L.flags & TSXC_FLAG_DIRECT)) {
        if ((rc = ensure_acc(t))) return rc;
        if (2 * n_words <= t->acc_cap_words) {
            // accumulate: append the batch to the current buffer (word aligned), mark its read ends and the padding
            if (t->acc[t->acc_cur].fill_words + n_words > t->acc_cap_words && (rc = flush_acc(t))) return rc;
            tsxc_table::Accum& a = t->acc[t->acc_cur];
            if (a.fill_words == 0 && a.busy) { CU(cudaStreamWaitEvent(t->copy_stream, a.consumed, 0)); a.busy = false; }
            // the offsets are only needed by k_mark_ends on the copy stream: stream order protects the slot's reuse
            if ((rc = ensure(t, &st.d_offsets, &st.cap_offsets, (size_t)n_reads + 1))) return rc;
            const uint64_t base = a.fill_words * 32;
            CU(cudaMemcpyAsync(a.d_packed + a.fill_words, packed, n_words * sizeof(uint64_t), cudaMemcpyHostToDevice, t->copy_stream));
            CU(cudaMemcpyAsync(st.d_offsets, offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, t->copy_stream));
            CU(cudaMemsetAsync(a.d_ends + a.fill_words, 0, n_words * sizeof(uint32_t), t->copy_stream));
            k_mark_ends<<<grid_for(t, n_reads), kBlockThreads, 0, t->copy_stream>>>(st.d_offsets, n_reads, a.d_ends, base);
            if (n_bases & 31) k_mark_padding<<<1, 32, 0, t->copy_stream>>>(a.d_ends, base + n_bases, base + n_words * 32);
            t->n_launches += 2;
            a.fill_words += n_words;
            CU(cudaGetLastError());
            return TSXC_OK;
        }
        if ((rc = flush_acc(t))) return rc;        // a batch this large is a pass of its own; keep the order of submission
    }
    // the slot may still be read by the kernel of two calls ago
    if (st.used) CU(cudaStreamWaitEvent(t->copy_stream, st.done, 0));
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
    return add_keys_device(t, d_kmers, n);
}

int tsxc_add_kmers(tsxc_table* t, const uint64_t* kmers, uint64_t n) {
    if (!t || (!kmers && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const size_t words = (size_t)n * t->L.KW;
    { const int frc = flush_acc(t); if (frc) return frc; }
    // the staging buffer is shared with other calls: order against the compute stream
    CU(cudaStreamSynchronize(t->stream));
    int rc = ensure(t, &t->d_keys, &t->cap_keys, words);
    if (rc) return rc;
    CU(cudaMemcpyAsync(t->d_keys, kmers, words * sizeof(uint64_t), cudaMemcpyHostToDevice, t->stream));
    return add_keys_device(t, t->d_keys, n);
}

int tsxc_sync(tsxc_table* t) {
    if (!t) return TSXC_E_INVALID;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaStreamSynchronize(t->stream));
    unsigned long long flags = 0;
    CU(cudaMemcpy(&flags, t->d_ctr + CTR_ERRORS, sizeof flags, cudaMemcpyDeviceToHost));
    collect_main_ms(t);          // recycle the timing event pairs of the batches that have completed
    return status_from_flags(t, flags);
}

static int lookup_device_locked(tsxc_table* t, const uint64_t* d_kmers, uint64_t n, uint64_t* d_counts_out) {
    const RadixGeom& rg = t->rg_lookup;
    cudaStream_t s = t->stream;
    const uint64_t sorted_min = env_u64("TSXC_LOOKUP_SORT_MIN", 1ULL << 18);
    if (t->radix_on && rg.d1 > 0 && t->L.shard_bits == 0 && !(t->L.flags & TSXC_FLAG_DIRECT) && n >= sorted_min && t->d_fhist) {
        // sort the queries by table region (same partition passes as the insert), then probe in that order
        int rc;
        for (int i = 0; i < 2; ++i) if ((rc = ensure(t, &t->d_pairs[i], &t->cap_pairs[i], (size_t)n * 2))) return rc;
        const int grid_q = grid_for(t, n), grid_p = t->sms * 2;
#define M(KW_) k_hash_queries<KW_><<<grid_q, kBlockThreads, 0, s>>>(t->tv, d_kmers, n, t->d_pairs[0])
        TSX_DISPATCH_KW(t->L, M);
#undef M
        RadixGeom g1 = rg;                    // pass 1: the whole array is one bin, split by digit 1
        g1.nbl = 1; g1.d2 = rg.d1; g1.nb2 = rg.nb1; g1.shift2 = rg.shift1;
        k_single_bin<<<1, kNB + 32, 0, s>>>(t->d_ctl, n);
        k_plan_group<2><<<1, kNB, 0, s>>>(t->d_ctl, 0, n, 1, t->d_fhist, g1.nb2);
        k_hist_keys<2><<<grid_p, kRadixThreads, 0, s>>>(t->tv, g1, t->d_ctl, t->d_pairs[0], t->d_fhist);
        k_scan_fine<<<1, 1024, 0, s>>>(t->d_ctl, t->d_fhist, t->d_fcur, g1.nb2);
        k_part_keys<2><<<grid_p, kRadixThreads, 0, s>>>(t->tv, g1, t->d_ctl, t->d_pairs[0], t->d_fcur, t->d_pairs[1]);
        const uint64_t* sorted = t->d_pairs[1];
        t->n_launches += 6;
        if (rg.d2) {                          // pass 2: digit 2 inside every digit-1 bin
            k_coff_from_cursors<<<1, kNB + 32, 0, s>>>(t->d_ctl, t->d_fcur, rg.nb1, n);
            k_plan_group<2><<<1, kNB, 0, s>>>(t->d_ctl, 0, n, rg.nbl, t->d_fhist, rg.nbl * rg.nb2);
            k_hist_keys<2><<<grid_p, kRadixThreads, 0, s>>>(t->tv, rg, t->d_ctl, t->d_pairs[1], t->d_fhist);
            k_scan_fine<<<1, 1024, 0, s>>>(t->d_ctl, t->d_fhist, t->d_fcur, rg.nbl * rg.nb2);
            k_part_keys<2><<<grid_p, kRadixThreads, 0, s>>>(t->tv, rg, t->d_ctl, t->d_pairs[1], t->d_fcur, t->d_pairs[0]);
            sorted = t->d_pairs[0];
            t->n_launches += 5;
        }
#define M(KW_, W_) k_lookup_pairs<KW_, W_><<<grid_q, kBlockThreads, 0, s>>>(t->tv, sorted, n, d_kmers, d_counts_out)
        TSX_DISPATCH(t->L, M);
#undef M
        t->n_launches++;
        CU(cudaGetLastError());
        return TSXC_OK;
    }
    const int grid = grid_for(t, n);
#define M(KW_, W_) k_lookup<KW_, W_><<<grid, kBlockThreads, 0, s>>>(t->tv, d_kmers, n, d_counts_out)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_lookup_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n, uint64_t* d_counts_out) {
    if (!t || ((!d_kmers || !d_counts_out) && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    return lookup_device_locked(t, d_kmers, n, d_counts_out);
}

int tsxc_lookup(tsxc_table* t, const uint64_t* kmers, uint64_t n, uint64_t* counts_out) {
    if (!t || ((!kmers || !counts_out) && n)) return fail(t, TSXC_E_INVALID, "null argument");
    if (n == 0) return TSXC_OK;
    // one critical section: the staging buffers are shared with add_kmers / dump on the same handle
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    CU(cudaStreamSynchronize(t->stream));
    int rc;
    if ((rc = ensure(t, &t->d_keys, &t->cap_keys, (size_t)n * t->L.KW))) return rc;
    if ((rc = ensure(t, &t->d_counts, &t->cap_counts, (size_t)n))) return rc;
    CU(cudaMemcpyAsync(t->d_keys, kmers, n * t->L.KW * sizeof(uint64_t), cudaMemcpyHostToDevice, t->stream));
    if ((rc = lookup_device_locked(t, t->d_keys, n, t->d_counts))) return rc;
    CU(cudaMemcpyAsync(counts_out, t->d_counts, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return TSXC_OK;
}

int tsxc_stats(tsxc_table* t, tsxc_stats_t* out) {
    if (!t || !out) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
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
    out->partition_ms = t->phase_ms[PH_HIST] + t->phase_ms[PH_PART1] + t->phase_ms[PH_PART2]; out->insert_ms = t->phase_ms[PH_INSERT];
    out->hist_ms = t->phase_ms[PH_HIST]; out->part1_ms = t->phase_ms[PH_PART1]; out->part2_ms = t->phase_ms[PH_PART2];
    out->chunk_cap_keys = t->cap_A; out->group_cap_keys = t->pg.paged ? (1ULL << t->pg.page_log2) : 0;
    out->radix_digit1_bits = t->rg.d1; out->radix_digit2_bits = t->rg.d2;
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

// Table scan -> (k-mer, count) -> text, all on the device; the host only copies and writes the bytes.
int tsxc_dump_file(tsxc_table* t, const char* path) {
    if (!t || !path) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    CU(cudaStreamSynchronize(t->copy_stream));
    CU(cudaStreamSynchronize(t->stream));
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(t, TSXC_E_IO, std::string("cannot open ") + path);
    const Layout& L = t->L;
    const uint64_t chunk_slots = std::min<uint64_t>(L.n_slots, 1ULL << 24);
    const uint64_t chunk_buckets = std::max<uint64_t>(1, chunk_slots / L.SPB);
    const size_t line_max = (size_t)L.k + 22;
    int rc;
    auto done = [&](int code) { std::fclose(f); return code; };
    if ((rc = ensure(t, &t->d_keys, &t->cap_keys, (size_t)chunk_slots * L.KW))) return done(rc);
    if ((rc = ensure(t, &t->d_counts, &t->cap_counts, (size_t)chunk_slots))) return done(rc);
    if ((rc = ensure(t, &t->d_text, &t->cap_text, (size_t)chunk_slots * line_max))) return done(rc);
    std::vector<char> host;
    unsigned long long* d_nbytes = t->d_ticket_k0;      // scratch counter
    for (uint64_t b0 = 0; b0 < L.n_buckets; b0 += chunk_buckets) {
        const uint64_t b1 = std::min(L.n_buckets, b0 + chunk_buckets);
        CU(cudaMemsetAsync(t->d_nout, 0, sizeof(unsigned long long), t->stream));
        CU(cudaMemsetAsync(d_nbytes, 0, sizeof(unsigned long long), t->stream));
        const int grid = grid_for(t, (b1 - b0) * L.SPB);
#define M(KW_, W_) k_dump<KW_, W_><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, b0, b1, t->d_keys, t->d_counts, chunk_slots, t->d_nout)
        TSX_DISPATCH(t->L, M);
#undef M
        unsigned long long n = 0, n_bytes = 0;
        CU(cudaMemcpyAsync(&n, t->d_nout, sizeof n, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaStreamSynchronize(t->stream));
        t->n_launches++;
        if (n == 0) continue;
        k_format_dump<<<grid_for(t, n), kBlockThreads, 0, t->stream>>>(t->d_keys, t->d_counts, n, L.k, L.KW, t->d_text, d_nbytes);
        t->n_launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&n_bytes, d_nbytes, sizeof n_bytes, cudaMemcpyDeviceToHost, t->stream));
        CU(cudaStreamSynchronize(t->stream));
        host.resize((size_t)n_bytes);
        CU(cudaMemcpy(host.data(), t->d_text, (size_t)n_bytes, cudaMemcpyDeviceToHost));
        if (std::fwrite(host.data(), 1, host.size(), f) != host.size()) return done(fail(t, TSXC_E_IO, "write failed"));
    }
    if (std::fclose(f) != 0) return fail(t, TSXC_E_IO, "write failed");
    return TSXC_OK;
}

// hist_out[c] = number of distinct k-mers with count c, c < n_bins - 1; hist_out[n_bins - 1] = all with a larger count
int tsxc_histogram(tsxc_table* t, uint64_t* hist_out, uint32_t n_bins) {
    if (!t || !hist_out || n_bins < 2 || n_bins > 4096) return fail(t, TSXC_E_INVALID, "histogram: 2 <= n_bins <= 4096");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    { const int frc = flush_acc(t); if (frc) return frc; }
    CU(cudaStreamSynchronize(t->copy_stream));
    int rc = ensure(t, &t->d_counts, &t->cap_counts, (size_t)n_bins);
    if (rc) return rc;
    CU(cudaMemsetAsync(t->d_counts, 0, n_bins * sizeof(uint64_t), t->stream));
    const int grid = grid_for(t, t->L.n_slots);
#define M(KW_, W_) k_histogram<KW_, W_><<<grid, kBlockThreads, 0, t->stream>>>(t->tv, n_bins, (unsigned long long*)t->d_counts)
    TSX_DISPATCH(t->L, M);
#undef M
    t->n_launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hist_out, t->d_counts, n_bins * sizeof(uint64_t), cudaMemcpyDeviceToHost, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    return TSXC_OK;
}

/* ---- multi-GPU routing -------------------------------------------------------------------------------------- */
int tsxc_route_info(tsxc_table* t, tsxc_route_info_t* out) {
    if (!t || !out) return fail(t, TSXC_E_INVALID, "null argument");
    std::memset(out, 0, sizeof *out);
    out->n_shards = 1u << t->L.shard_bits; out->shard_rank = t->L.shard_rank;
    out->bins = t->rg.nb1; out->bins_per_shard = t->rg.nbl; out->key_words = t->L.KW;
    out->recv_cap_keys = t->cap_A;
    return TSXC_OK;
}

int tsxc_route_recv_buffer(tsxc_table* t, uint64_t cap_keys, void** d_ptr_out, uint64_t* cap_keys_out) {
    if (!t || !d_ptr_out) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    if (t->peers_set) return fail(t, TSXC_E_INVALID, "receive buffer already exported");
    // cap_keys == 0: as much as free memory allows (radix_reserve's policy), else exactly what the caller asks for
    int rc = radix_reserve(t, cap_keys ? cap_keys : ~0ULL >> 8, true);
    if (rc) return rc;
    *d_ptr_out = t->d_A;
    if (cap_keys_out) *cap_keys_out = t->cap_A;
    return TSXC_OK;
}

int tsxc_route_set_peers(tsxc_table* t, void* const* peer_buffers, uint64_t recv_cap_keys) {
    if (!t || !peer_buffers) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    const uint32_t n = 1u << t->L.shard_bits;
    if (!t->d_A || peer_buffers[t->L.shard_rank] != (void*)t->d_A) return fail(t, TSXC_E_INVALID, "peer_buffers[shard_rank] must be this handle's receive buffer");
    if (recv_cap_keys == 0 || recv_cap_keys > t->cap_A) return fail(t, TSXC_E_INVALID, "common receive capacity exceeds this rank's buffer");
    CU(cudaMemcpyAsync(t->d_peers, peer_buffers, n * sizeof(void*), cudaMemcpyHostToDevice, t->stream));
    CU(cudaStreamSynchronize(t->stream));
    t->cap_A = recv_cap_keys;        // every rank plans against the smallest buffer of the group
    t->peers_set = true;
    return TSXC_OK;
}

int tsxc_route_begin(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases,
                     uint32_t* rounds_out) {
    if (!t || !rounds_out || (n_reads && !d_offsets) || (n_bases && !d_packed)) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    if (!t->peers_set) return fail(t, TSXC_E_INVALID, "tsxc_route_set_peers has not been called");
    const uint64_t n_words = (n_bases + 31) >> 5;
    cudaStream_t s = t->stream;
    int rc = ensure(t, &t->d_ends, &t->cap_ends, (size_t)n_words + 8);
    if (rc) return rc;
    CU(cudaMemsetAsync(t->d_ends, 0, (n_words + 8) * sizeof(uint32_t), s));
    if (n_reads) {
        k_mark_ends<<<grid_for(t, n_reads), kBlockThreads, 0, s>>>(d_offsets, n_reads, t->d_ends, 0);
        t->n_launches++;
    }
    const RadixGeom& rg = t->rg;
    const uint64_t seg_words = 1ULL << rg.seg_log2;
    const uint64_t n_segs = (n_words + seg_words - 1) / seg_words;
    // a rank sends chunks of at most 7/8 of the common receive capacity: owners are hash-uniform, so what a rank
    // receives per round concentrates around the mean chunk size; k_route_offsets checks the exact totals
    const uint64_t cap_send = std::max<uint64_t>(2 * (32ULL << rg.seg_log2), t->cap_A - t->cap_A / 8);
    uint64_t segs_per_run = 0;
    plan_bounds(rg, cap_send, &segs_per_run, 0, nullptr);
    if (n_segs > segs_per_run) return fail(t, TSXC_E_INVALID, "batch too large for one routing plan: split it");
    uint32_t max_chunks = 0;
    if (n_segs) {
        if ((rc = launch_hist_plan(t, d_packed, t->d_ends, n_words, n_bases, 0, (uint32_t)n_segs, cap_send, true, s))) return rc;
        plan_bounds(rg, cap_send, nullptr, (uint32_t)n_segs, &max_chunks);
    } else {
        CU(cudaMemsetAsync(t->d_ctl, 0, 16, s));     // n_chunks = 0
    }
    t->route_packed = d_packed; t->route_n_words = n_words; t->route_n_bases = n_bases; t->route_rounds = max_chunks;
    *rounds_out = max_chunks;
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_route_hist(tsxc_table* t, uint32_t round, uint32_t* d_hist_out) {
    if (!t || !d_hist_out) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    k_round_hist<<<1, 1024, 0, t->stream>>>(t->d_ctl, round, t->rg.nb1, d_hist_out);
    t->n_launches++;
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_route_send(tsxc_table* t, uint32_t round, const uint32_t* d_hist_all) {
    if (!t || !d_hist_all) return fail(t, TSXC_E_INVALID, "null argument");
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    if (!t->peers_set) return fail(t, TSXC_E_INVALID, "tsxc_route_set_peers has not been called");
    const RadixGeom& rg = t->rg;
    cudaStream_t s = t->stream;
    PhaseTimer pt(t, s, PH_PART1);
    k_route_offsets<<<1, 1024, 0, s>>>(t->d_ctl, round, d_hist_all, 1u << t->L.shard_bits, t->L.shard_rank, rg.nb1, rg.nbl, t->cap_A,
                                       slice_keys_of(t->L), t->d_ctr + CTR_ERRORS);
    { const int prc = launch_part_reads(t, round, t->route_packed, t->d_ends, t->route_n_words, t->route_n_bases, 0, t->d_peers, s); if (prc) return prc; }
    pt.end(3);
    CU(cudaGetLastError());
    return TSXC_OK;
}

int tsxc_route_insert(tsxc_table* t) {
    if (!t) return TSXC_E_INVALID;
    std::lock_guard<std::mutex> g(t->mu);
    CU(cudaSetDevice(t->device));
    int rc = launch_sort_insert(t, t->stream);
    if (rc) return rc;
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

int tsxc_enable_peer_access(int device, int peer) {
    tsxc_table* t = nullptr;
    CU(cudaSetDevice(device));
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return fail(t, TSXC_E_CUDA, "no peer access between these devices");
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return TSXC_OK; }
    CU(e);
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
    unsigned long long* ticket = t->d_ticket_k0;
    CU(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), t->stream));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    CU(cudaEventRecord(a, t->stream));
    const int blocks_per_sm = (mode >> 8) >= 1 && (mode >> 8) <= 8 ? (mode >> 8) : 8;      // bits 8..: resident blocks per SM (default 8)
    k_k0_region_sweep<<<t->sms * blocks_per_sm, kBlockThreads, 0, t->stream>>>(t->d_words, rw, fw / rw, ops_per_region, ops_per_item, mode & 0xff, ticket);
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

// min(k-mer, reverse complement) as TSXC_FLAG_CANONICAL tables see it (host code path of the same functions)
int tsxc_debug_canonical(uint32_t k, const uint64_t* key, uint64_t* out) {
    const uint32_t KW = tsxc_key_words(k);
    if (!KW || !key || !out) return TSXC_E_INVALID;
    const HashParams hp = make_hash_params(k, true);
    if (KW == 1) { Key<1> x{{key[0]}}; auto c = canonical_key<1>(x, hp); out[0] = c.w[0]; }
    else if (KW == 2) { Key<2> x{{key[0], key[1]}}; auto c = canonical_key<2>(x, hp); out[0] = c.w[0]; out[1] = c.w[1]; }
    else { Key<4> x{{key[0], key[1], key[2], key[3]}}; auto c = canonical_key<4>(x, hp); for (int j = 0; j < 4; ++j) out[j] = c.w[j]; }
    return TSXC_OK;
}

// One block round of S1's sparse walk (SparseStage in tsx_radix.cuh) on the host, through the same helper functions the
// kernel calls: the k-mers that start in stream words [round, round + 512) and before word w_end, in the order the
// kernel's threads would take them.  keys_out: room for 512 * 32 k-mers.
int tsxc_debug_sparse_round(uint32_t k, const uint64_t* packed, const uint32_t* ends, uint64_t n_words, uint64_t n_bases,
                            uint64_t round, uint64_t w_end, uint64_t* keys_out, uint32_t* n_out) {
    const uint32_t KW = tsxc_key_words(k);
    if (!KW || !packed || !ends || !keys_out || !n_out) return TSXC_E_INVALID;
    const HashParams hp = make_hash_params(k);
    const uint32_t NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);
    std::vector<uint32_t> stream(2 * (kRadixThreads + 4 + 1)), vb(kRadixThreads), pre(kRadixThreads);
    auto word = [&](uint64_t w) { return w < n_words ? packed[w] : 0ULL; };
    auto eword = [&](uint64_t w) { return w < n_words ? ends[w] : 0u; };
    for (uint32_t t = 0; t < (uint32_t)kRadixThreads + KW + 1; ++t) {
        const uint64_t v = word(round + t);
        stream[2 * t] = (uint32_t)v; stream[2 * t + 1] = (uint32_t)(v >> 32);
    }
    uint32_t total = 0;
    for (uint32_t t = 0; t < (uint32_t)kRadixThreads; ++t) {
        const uint64_t wi = round + t;
        uint32_t dist_after = 0xffffffffu;                 // first_end_after over the NE following words
        for (uint32_t j = NE; j >= 1; --j) { const uint32_t e = eword(wi + j); if (e) dist_after = 32u * (j - 1) + (uint32_t)__builtin_ctz(e); }
        const uint64_t g0 = wi << 5;
        int omax = -1;
        if (wi < w_end && g0 + k <= n_bases) omax = n_bases - g0 - k < 31 ? (int)(n_bases - g0 - k) : 31;
        vb[t] = valid_starts(eword(wi), dist_after, k, omax);
        pre[t] = total;
        total += popc32(vb[t]);
    }
    // tiles of kRadixThreads * OPT positions, thread t of a tile takes OPT consecutive ones (as k_part_reads does)
    const uint32_t OPT = 8 / KW;
    for (uint32_t i0 = 0; i0 < total; i0 += OPT) {
      SparseCursor cur = sparse_seek(pre.data(), vb.data(), kRadixThreads, i0);
      for (uint32_t i = i0; i < i0 + OPT && i < total; ++i) {
        const uint32_t o = sparse_next(cur, vb.data());
        const uint32_t w = cur.w;
        if (KW == 1) { auto key = kmer_from_stream32<1>(stream.data(), w, o, hp); keys_out[i] = key.w[0]; }
        else if (KW == 2) { auto key = kmer_from_stream32<2>(stream.data(), w, o, hp); keys_out[2 * (uint64_t)i] = key.w[0]; keys_out[2 * (uint64_t)i + 1] = key.w[1]; }
        else { auto key = kmer_from_stream32<4>(stream.data(), w, o, hp); for (int j = 0; j < 4; ++j) keys_out[4 * (uint64_t)i + j] = key.w[j]; }
      }
    }
    *n_out = total;
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
