// tsx_kernels.cuh — the sm_100a kernels of the counting path.
//
//   K1+K2  k_count_reads        packed reads -> forward k-mers -> hash -> insert, fused (tables that fit L2 / TLB reach)
//          (large tables: the region-sorted pipeline of tsx_radix.cuh)
//          k_add_kmers          batched addKmer on explicit k-mers
//   K4     k_lookup             batched getKmerCount(kmer)
//   K5     k_dump               table scan -> (k-mer, count) via the inverse hash
//   K6     multi-GPU routing = S1 of tsx_radix.cuh storing into the owners' peer-mapped buffers
//          k_mark_ends          read offsets -> "last base of a read" bitmap
//   K0     k_k0_random_rmw / k_k0_windowed / k_k0_region_sweep   random-access roofline microbenchmarks
//
// Reference semantics (paths relative to mjoppich/tsxCount):
//   extraction  src/mains/testExecution.h:15-36   every forward substring seq[i:i+k]; none if len < k
//   encoding    src/utils/SequenceUtils.h:86-123  base i -> bits [2i,2i+1]: a k-mer is a contiguous
//               2k-bit window of the 2-bit packed read stream, so extraction is a funnel shift
//   driver      src/mains/main.cpp:159-192        per read: createKMers -> fromSequence -> addKmer
#pragma once

#include <cstdint>

#include "tsx_table.cuh"

namespace tsx {

constexpr int kBlockThreads = 256;

// ---- read-boundary bitmap ---------------------------------------------------------------------
// bit g of `ends` is set iff base g is the last base of a read.  A k-mer starting at g is valid iff
// no end bit lies in [g, g+k-2] (it may end exactly on a boundary) and g+k <= n_bases.
__global__ void __launch_bounds__(kBlockThreads) k_mark_ends(const uint64_t* __restrict__ offsets, uint64_t n_reads,
                                                             uint32_t* __restrict__ ends, uint64_t base) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        if (e > b) atomicOr(ends + ((base + e - 1) >> 5), 1u << ((base + e - 1) & 31));
    }
}

// Positions [from, to) (fewer than 32, inside one word) are padding between two batches of an accumulated stream:
// every one of them is marked as a read end, so no k-mer (k >= 2) starts in or runs through the padding.
__global__ void k_mark_padding(uint32_t* __restrict__ ends, uint64_t from, uint64_t to) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || to <= from) return;
    const uint32_t lo = (uint32_t)(from & 31), n = (uint32_t)(to - from);
    const uint32_t mask = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << lo;
    atomicOr(ends + (from >> 5), mask);
}

// ---- warp-level window loader -----------------------------------------------------------------
// Lane `lane` of a warp owns stream word base+lane; it needs NEXT further words to cover k-mers that
// start in its word.  One coalesced load per lane plus NEXT tail words loaded by the first lanes;
// neighbours come from shuffles instead of re-reading global memory.
template <int NEXT, typename T>
__device__ __forceinline__ void load_window(const T* __restrict__ src, uint64_t base, uint64_t n_words, unsigned lane,
                                            T (&win)[NEXT + 1]) {
    const unsigned full = 0xffffffffu;
    const uint64_t t = base + lane;
    const T cur = t < n_words ? __ldg(src + t) : T(0);
    T tail = T(0);
    if (lane < NEXT) {
        const uint64_t u = base + 32 + lane;
        tail = u < n_words ? __ldg(src + u) : T(0);
    }
    win[0] = cur;
#pragma unroll
    for (int j = 1; j <= NEXT; ++j) {
        const T a = __shfl_down_sync(full, cur, j);
        const T b = __shfl_sync(full, tail, (lane + j) & 31);
        win[j] = (lane + j < 32) ? a : b;
    }
}

// Enumerates the k-mers that start in one 32-base stream word and hands groups of identical k-mers
// to `sink(key, count)` on one lane per group.
//   - thread-local run aggregation: consecutive identical k-mers (homopolymer runs) are merged;
//   - warp pre-aggregation: lanes holding the same k-mer in the same step are merged with
//     __match_any_sync / __reduce_add_sync so a heavy hitter costs one table update per warp step.
template <int KW, bool WARP_AGG, typename Sink>
__device__ __forceinline__ void for_each_kmer_group(const uint64_t (&win)[KW + 1], uint32_t ends_cur, uint32_t dist_after,
                                                    uint64_t word_index, uint64_t n_bases, uint32_t k,
                                                    const HashParams& hp, Sink&& sink) {
    const unsigned full = 0xffffffffu;
    const uint64_t g0 = word_index << 5;
    uint32_t dist = dist_after;  // distance from position g0+32 to the next read end at/after it
    Key<KW> pend;
#pragma unroll
    for (int j = 0; j < KW; ++j) pend.w[j] = 0;
    uint32_t pend_cnt = 0;

#pragma unroll 2
    for (int o = 31; o >= -1; --o) {
        bool valid = false;
        Key<KW> key;
#pragma unroll
        for (int j = 0; j < KW; ++j) key.w[j] = 0;
        if (o >= 0) {
            dist = ((ends_cur >> o) & 1u) ? 0u : (dist == 0xffffffffu ? dist : dist + 1u);
            valid = (dist >= k - 1) && (g0 + (uint64_t)o + k <= n_bases);
            if (valid) {
                const unsigned sh = 2u * (unsigned)o;
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                    key.w[j] = sh ? ((win[j] >> sh) | (win[j + 1] << (64 - sh))) : win[j];
                }
#pragma unroll
                for (int j = 0; j < KW; ++j) key.w[j] &= word_mask<KW>(j, hp);
            }
        }
        // run aggregation: emit the pending k-mer when the new one differs (or at the end, o == -1)
        bool emit = false;
        Key<KW> ekey = pend;
        uint32_t ecnt = pend_cnt;
        if (valid && pend_cnt && key_eq<KW>(key, pend)) {
            ++pend_cnt;
        } else {
            emit = pend_cnt != 0;
            pend = key;
            pend_cnt = valid ? 1u : 0u;
        }
        if (WARP_AGG) {
            const unsigned emask = __ballot_sync(full, emit);
            if (emask == 0) continue;
            if (emit) {
                unsigned peers = __match_any_sync(emask, ekey.w[0]);
#pragma unroll
                for (int j = 1; j < KW; ++j) peers &= __match_any_sync(emask, ekey.w[j]);
                // singleton groups (the common case) must not enter the reduction: with per-group masks
                // REDUX is issued once per distinct mask, i.e. 32 times per step for all-distinct k-mers
                uint32_t total = ecnt;
                if (peers & (peers - 1)) total = __reduce_add_sync(peers, ecnt);
                if ((unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u)) sink(ekey, (uint64_t)total);
            }
        } else {
            if (emit) sink(ekey, (uint64_t)ecnt);
        }
    }
}

// distance from the first position after the lane's word to the next read end (0xffffffff: none in view)
template <int NE>
__device__ __forceinline__ uint32_t first_end_after(const uint32_t (&ewin)[NE + 1]) {
    uint32_t d = 0xffffffffu;
#pragma unroll
    for (int j = NE; j >= 1; --j)
        if (ewin[j]) d = 32u * (uint32_t)(j - 1) + (uint32_t)(__ffs(ewin[j]) - 1);
    return d;
}

// ---- K1+K2 fused: count every k-mer of a packed read batch -------------------------------------
// Words [w_begin, w_end) of the batch.  only_if != nullptr: the launch is the fallback of a two-phase chunk and
// does nothing unless *only_if is set (see launch_count_reads_partitioned).
template <int KW, int W, bool WARP_AGG>
__global__ void __launch_bounds__(kBlockThreads) k_count_reads(const __grid_constant__ TableView tv, const uint64_t* __restrict__ packed,
                                                               const uint32_t* __restrict__ ends, uint64_t w_begin, uint64_t w_end,
                                                               uint64_t n_words, uint64_t n_bases,
                                                               const unsigned int* __restrict__ only_if) {
    constexpr int NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);  // end-bitmap words of look-ahead: ceil((k-1)/32)
    if (only_if && __ldcg(only_if) == 0) return;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    LocalStats st;
    for (uint64_t base = w_begin + warp * 32; base < w_end; base += n_warps * 32) {
        uint64_t win[KW + 1];
        uint32_t ewin[NE + 1];
        load_window<KW, uint64_t>(packed, base, n_words, lane, win);
        load_window<NE, uint32_t>(ends, base, n_words, lane, ewin);
        const uint32_t dist_after = first_end_after<NE>(ewin);
        const uint64_t limit = (base + lane < w_end) ? n_bases : 0;   // lanes past the range emit nothing
        for_each_kmer_group<KW, WARP_AGG>(win, ewin[0], dist_after, base + lane, limit, tv.L.k, tv.hp,
                                          [&](const Key<KW>& key, uint64_t cnt) {
                                              insert_hashed<KW, W>(tv, hash_key<KW>(key, tv.hp), cnt, st);
                                          });
    }
    flush_stats(tv, st);
}

// ---- K2: batched addKmer -----------------------------------------------------------------------
template <int KW, int W, bool WARP_AGG>
__global__ void __launch_bounds__(kBlockThreads) k_add_kmers(const __grid_constant__ TableView tv, const uint64_t* __restrict__ kmers, uint64_t n) {
    const unsigned full = 0xffffffffu;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    LocalStats st;
    const uint64_t n_round = (n + 31) & ~31ULL;  // keep warps converged for the collectives
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        Key<KW> key;
#pragma unroll
        for (int j = 0; j < KW; ++j) key.w[j] = valid ? __ldg(kmers + i * KW + j) : 0ULL;
        const unsigned vmask = __ballot_sync(full, valid);
        if (!valid) continue;
        uint64_t cnt = 1;
        bool lead = true;
        if (WARP_AGG) {
            unsigned peers = __match_any_sync(vmask, key.w[0]);
#pragma unroll
            for (int j = 1; j < KW; ++j) peers &= __match_any_sync(vmask, key.w[j]);
            cnt = (uint64_t)__popc(peers);
            lead = (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u);
        }
        if (lead) {
#pragma unroll
            for (int j = 0; j < KW; ++j) key.w[j] &= word_mask<KW>(j, tv.hp);
            insert_hashed<KW, W>(tv, hash_key<KW>(key, tv.hp), cnt, st);
        }
    }
    flush_stats(tv, st);
}

// ---- K4: batched lookup ------------------------------------------------------------------------
template <int KW, int W>
__global__ void __launch_bounds__(kBlockThreads) k_lookup(const __grid_constant__ TableView tv, const uint64_t* __restrict__ kmers, uint64_t n,
                                                          uint64_t* __restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Key<KW> key;
        bool in_range = true;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            key.w[j] = __ldg(kmers + i * KW + j);
            const uint64_t m = word_mask<KW>(j, tv.hp);
            in_range &= (key.w[j] & ~m) == 0;
        }
        counts[i] = in_range ? lookup_hashed<KW, W>(tv, hash_key<KW>(key, tv.hp)) : 0ULL;
    }
}

// ---- K5: dump ----------------------------------------------------------------------------------
// Scans buckets [b0, b1); every primary entry yields (k-mer, count).  Output positions are claimed
// with one atomicAdd per warp.
template <int KW, int W>
__global__ void __launch_bounds__(kBlockThreads) k_dump(const __grid_constant__ TableView tv, uint64_t b0, uint64_t b1, uint64_t* __restrict__ kmers_out,
                                                        uint64_t* __restrict__ counts_out, uint64_t capacity,
                                                        unsigned long long* __restrict__ n_out) {
    constexpr int SPB = 4 / W;
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_slots = (b1 - b0) * SPB;
    const uint64_t n_round = (n_slots + 31) & ~31ULL;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool found = false;
        Key<KW> key;
        uint64_t cnt = 0;
        if (i < n_slots) {
            const uint64_t bucket = b0 + i / SPB;
            const uint32_t sl = (uint32_t)(i % SPB);
            const uint64_t* ep = tv.words + (bucket << 2) + sl * W;
            uint64_t e[W];
#pragma unroll
            for (int j = 0; j < W; ++j) e[j] = __ldcg(ep + j);
            const uint64_t h = e[W - 1];
            if (h != 0 && !(h & tv.f_ovf)) {
                found = true;
                const Key<KW> H = hash_of_entry<KW, W>(tv, bucket, e);
                key = unhash_key<KW>(H, tv.hp);
                cnt = h >> tv.vshift;
                if (h & tv.f_hasovf) {
                    const uint32_t pi = (uint32_t)(h & tv.rmask);
                    const uint64_t home = (bucket - tri(pi)) & tv.lbl_mask;
                    cnt += overflow_lookup<KW, W>(tv, home, pi, sl) << tv.L.V;
                }
            }
        }
        const unsigned fm = __ballot_sync(full, found);
        if (fm == 0) continue;
        unsigned long long basepos = 0;
        if (lane == (unsigned)(__ffs(fm) - 1)) basepos = atomicAdd(n_out, (unsigned long long)__popc(fm));
        basepos = __shfl_sync(full, basepos, __ffs(fm) - 1);
        if (found) {
            const uint64_t pos = basepos + __popc(fm & ((1u << lane) - 1u));
            if (pos < capacity) {
#pragma unroll
                for (int j = 0; j < KW; ++j) kmers_out[pos * KW + j] = key.w[j];
                counts_out[pos] = cnt;
            }
        }
    }
}

// ---- count histogram: hist[c] = number of distinct k-mers with count c (the last bin takes everything above) --------
// An opt-in extension (the reference has no histogram).  One scan of the table; per-block
// histogram in shared memory (n_bins <= 4096), flushed with one global atomic per non-empty bin.
template <int KW, int W>
__global__ void __launch_bounds__(kBlockThreads) k_histogram(const __grid_constant__ TableView tv, uint32_t n_bins,
                                                             unsigned long long* __restrict__ hist) {
    constexpr int SPB = 4 / W;
    __shared__ unsigned int sh[4096];
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_slots = tv.L.n_buckets * SPB;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += stride) {
        const uint64_t bucket = i / SPB;
        const uint32_t sl = (uint32_t)(i % SPB);
        const uint64_t h = __ldcg(tv.words + (bucket << 2) + sl * W + (W - 1));
        if (h == 0 || (h & tv.f_ovf)) continue;
        uint64_t cnt = h >> tv.vshift;
        if (h & tv.f_hasovf) {
            const uint32_t pi = (uint32_t)(h & tv.rmask);
            const uint64_t home = (bucket - tri(pi)) & tv.lbl_mask;
            cnt += overflow_lookup<KW, W>(tv, home, pi, sl) << tv.L.V;
        }
        atomicAdd(&sh[cnt < n_bins ? (uint32_t)cnt : n_bins - 1], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// ---- K5b: (k-mer, count) pairs -> text lines "KMER<TAB>COUNT\n" on the device ---------------------------------------
// Format of count_kmers.py:32-34 of the reference; bases decoded like TSXSeqUtils::toSequence (SequenceUtils.h:47-84).
// One thread per pair: line length = k + 2 + decimal digits; a block-wide scan places the block's lines back to
// back, one atomicAdd per block claims its range of the output (line order = table order, unspecified anyway).
__global__ void __launch_bounds__(kBlockThreads) k_format_dump(const uint64_t* __restrict__ kmers, const uint64_t* __restrict__ counts,
                                                               uint64_t n, uint32_t k, uint32_t KW, char* __restrict__ text,
                                                               unsigned long long* __restrict__ n_bytes) {
    __shared__ uint32_t wsum[kBlockThreads / 32];
    __shared__ unsigned long long base_s;
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t n_round = (n + kBlockThreads - 1) / kBlockThreads * kBlockThreads;
    for (uint64_t i0 = (uint64_t)blockIdx.x * kBlockThreads; i0 < n_round; i0 += (uint64_t)gridDim.x * kBlockThreads) {
        const uint64_t i = i0 + threadIdx.x;
        uint64_t c = i < n ? counts[i] : 0;
        uint32_t digits = 1;
        for (uint64_t x = c; x >= 10; x /= 10) ++digits;
        const uint32_t len = i < n ? k + 2 + digits : 0;
        uint32_t inc = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(full, inc, d);
            if (lane >= (unsigned)d) inc += y;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kBlockThreads / 32; ++w) { const uint32_t x = wsum[w]; if ((unsigned)w < warp) pre += x; tot += x; }
        if (threadIdx.x == 0) base_s = atomicAdd(n_bytes, (unsigned long long)tot);
        __syncthreads();
        if (i < n) {
            char* out = text + base_s + pre + inc - len;
            const uint64_t* kw = kmers + i * KW;
            for (uint32_t b = 0; b < k; ++b) out[b] = "ACGT"[(kw[(2 * b) >> 6] >> ((2 * b) & 63)) & 3];
            out[k] = '\t';
            for (uint32_t d = digits; d > 0; --d) { out[k + d] = (char)('0' + c % 10); c /= 10; }
            out[k + 1 + digits] = '\n';
        }
        __syncthreads();
    }
}

// ---- K0: random 8-byte RMW roofline ------------------------------------------------------------
__global__ void __launch_bounds__(kBlockThreads) k_k0_random_rmw(uint64_t* __restrict__ words, uint64_t n_words_mask,
                                                                 uint64_t n_ops, int mode) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t sink = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += stride) {
        const uint64_t a = fmix64(i * kC1 + 0x1234567ULL) & n_words_mask;
        if (mode == 0) {
            atomicAdd((unsigned long long*)(words + a), 1ULL);            // RED, fire and forget
        } else if (mode == 1) {
            atomicCAS((unsigned long long*)(words + a), 0ULL, i | 1ULL);  // CAS with return
        } else if (mode == 2) {
            uint64_t w[4];
            load_bucket(words + (a & ~3ULL), w);                          // sector load, then atomic on it
            const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
            atomicAdd((unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL))), 1ULL << 40);
        } else if (mode == 3) {
            uint64_t w[4];
            load_bucket(words + (a & ~3ULL), w);                          // random 32-byte sector reads only
            sink += w[0] ^ w[1] ^ w[2] ^ w[3];
        } else {
            __stcg(words + a, i);                                         // random 8-byte stores only
        }
    }
    if (sink == 0x123456789ULL) words[0] = sink;
}

// K0w: like K0 but every block confines itself to its own window of the footprint (windows of different
// blocks are disjoint while grid*window <= footprint).  Separates address-translation reach from DRAM
// random-access rate: aggregate footprint stays far beyond L2 while each SM touches few pages.
__global__ void k_k0_windowed(uint64_t* __restrict__ words, uint64_t footprint_words, uint64_t window_words,
                              uint64_t ops_per_block, int mode) {
    const uint64_t n_windows = footprint_words / window_words;
    const uint64_t base = ((uint64_t)blockIdx.x % n_windows) * window_words;
    const uint64_t wmask = window_words - 1;
    uint64_t sink = 0;
    for (uint64_t i = threadIdx.x; i < ops_per_block; i += blockDim.x) {
        const uint64_t a = base + (fmix64((i + (uint64_t)blockIdx.x * ops_per_block) * kC1 + 0x1234567ULL) & wmask);
        if (mode == 0) {
            atomicAdd((unsigned long long*)(words + a), 1ULL);
        } else if (mode == 3) {
            uint64_t w[4];
            load_bucket(words + (a & ~3ULL), w);
            sink += w[0] ^ w[1] ^ w[2] ^ w[3];
        } else {
            uint64_t w[4];
            load_bucket(words + (a & ~3ULL), w);
            const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
            atomicAdd((unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL))), 1ULL << 40);
        }
    }
    if (sink == 0x123456789ULL) words[0] = sink;
}

// K0r: all blocks sweep the footprint region by region (work items handed out in order through a ticket), each
// region receiving ops_per_region random RMWs: the L2-blocked variant of the insert (is a second touch of a
// sector cheaper while its region is resident?).
__global__ void __launch_bounds__(kBlockThreads) k_k0_region_sweep(uint64_t* __restrict__ words, uint64_t region_words,
                                                                   uint64_t n_regions, uint64_t ops_per_region,
                                                                   uint32_t ops_per_item, int mode,
                                                                   unsigned long long* __restrict__ ticket) {
    __shared__ unsigned long long item_s;
    const uint64_t items_per_region = (ops_per_region + ops_per_item - 1) / ops_per_item;
    const unsigned long long n_items = n_regions * items_per_region;
    const uint64_t wmask = region_words - 1;
    uint64_t sink = 0;
    while (true) {
        if (threadIdx.x == 0) item_s = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const unsigned long long item = item_s;
        __syncthreads();
        if (item >= n_items) break;
        const uint64_t base = (item / items_per_region) * region_words;
        for (uint32_t i = threadIdx.x; i < ops_per_item; i += blockDim.x) {
            const uint64_t a = base + (fmix64((item * ops_per_item + i) * kC1 + 0x1234567ULL) & wmask);
            if (mode == 0) {                                  // RED only
                atomicAdd((unsigned long long*)(words + a), 1ULL);
            } else if (mode == 2) {                           // sector load -> RED (fire and forget)
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
                atomicAdd((unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL))), 1ULL << 40);
            } else if (mode == 3) {                           // sector load -> CAS whose result is consumed (the insert's claim)
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
                unsigned long long* p = (unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL)));
                const unsigned long long old = atomicCAS(p, w[(a + pick) & 3ULL], w[(a + pick) & 3ULL] + (1ULL << 40));
                sink += old;
            } else if (mode == 5) {                           // sector load -> CAS -> branch on its result: the thread really waits
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
                unsigned long long* p = (unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL)));
                const unsigned long long old = atomicCAS(p, w[(a + pick) & 3ULL], w[(a + pick) & 3ULL] + (1ULL << 40));
                if (old == 0x5a5a5a5a5a5a5a5aULL) break;
            } else if (mode >= 6 && mode <= 10) {             // sector load -> some other returning atomic -> wait
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
                unsigned long long* p = (unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL)));
                unsigned long long old;
                if (mode == 6) old = atomicCAS((unsigned int*)p, (unsigned int)w[(a + pick) & 3ULL], (unsigned int)w[(a + pick) & 3ULL] + 1u);
                else if (mode == 7) old = atomicAdd(p, 1ULL << 40);
                else if (mode == 8) old = atomicAdd((unsigned int*)p, 1u);
                else if (mode == 9) old = atomicExch(p, w[(a + pick) & 3ULL] + 1ULL);
                else old = atomicOr(p, 1ULL << (i & 31u));
                if (old == 0x5a5a5a5a5a5a5a5aULL) break;
            } else if (mode == 4) {                           // returning atomic only
                sink += atomicAdd((unsigned long long*)(words + a), 1ULL << 40);
            } else {                                          // sector load only
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                sink += w[0] ^ w[1] ^ w[2] ^ w[3];
            }
        }
    }
    if (sink == 0x123456789ULL) words[0] = sink;
}

}  // namespace tsx
