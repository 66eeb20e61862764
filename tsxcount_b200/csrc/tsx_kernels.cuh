// tsx_kernels.cuh — the sm_100a kernels of the counting path.
//
//   K1+K2  k_count_reads        packed reads -> forward k-mers -> hash -> insert, fused (small tables, and the
//                               device-side fallback of a two-phase chunk)
//   K1     k_partition_reads    phase A of the two-phase path: extract + hash, bin every k-mer by (owning shard,)
//                               table region; never touches the table
//   K2     k_insert_partitions  phase B: drain the bins region by region, insert
//          k_add_hash_counts    (hash, count) spill records
//          k_add_kmers          batched addKmer on explicit k-mers / on pre-hashed k-mers
//   K4     k_lookup             batched getKmerCount(kmer)
//   K5     k_dump               table scan -> (k-mer, count) via the inverse hash
//   K6     = K1 with bins keyed by (owner, region) on the sender + K2 on the receiver; the exchange is the host's
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
                                                             uint32_t* __restrict__ ends) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        if (e > b) atomicOr(ends + ((e - 1) >> 5), 1u << ((e - 1) & 31));
    }
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
template <int KW, int W, bool HASHED, bool WARP_AGG>
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
            if (!HASHED) {
#pragma unroll
                for (int j = 0; j < KW; ++j) key.w[j] &= word_mask<KW>(j, tv.hp);
                insert_hashed<KW, W>(tv, hash_key<KW>(key, tv.hp), cnt, st);
            } else {
                insert_hashed<KW, W>(tv, key, cnt, st);
            }
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

// ---- partitioned insert (TLB-aware two-phase path) -----------------------------------------------
// Measured on B200 (profiles/): uniformly random 8-byte RMWs over a 128 GiB table run at 9.8 G/s and a
// dependent sector-load + atomic at 4 G/s, but the same accesses confined to a 64 MiB window per thread
// block run at 19-22 G/s with the SAME aggregate footprint: the limiter of the naive scatter is address
// translation (per-SM TLB reach), not DRAM.  So large tables are updated in two phases per chunk of reads:
//   phase A  k_partition_reads   extract + hash, append the hash to the bin of the 64 MiB table region
//                                that owns its home bucket (streaming writes, 8*KW bytes per k-mer);
//   phase B  k_insert_partitions thread blocks drain one bin slice at a time, so every block probes
//                                inside one region (translations stay resident), consecutive blocks
//                                work on consecutive slices of the same region.
// k-mer groups that the warp already aggregated (count >= 2: homopolymer runs, heavy hitters) and k-mers
// whose bin is full bypass the bins and are inserted directly — bins never overflow, nothing is dropped.
struct PartView {
    uint64_t* buf;                 // P * cap * KW words
    unsigned long long* cursor;    // P fill counters (may exceed cap: excess went the direct way)
    uint64_t cap;                  // entries per bin
    uint32_t pshift;               // bin = ((global bucket index) >> pshift) & pmask
    uint32_t pmask;
    uint32_t P;                    // number of bins
    uint32_t run;                  // entries per private run (phase A), a power of two
    // routing mode (multi-GPU send side): bins are grouped by owning shard, bins_per_shard each; k-mers that
    // cannot be binned (pre-aggregated groups, full bins) become (hash, count) records in the owner's spill list
    uint64_t* spill;               // n_shards * spill_cap records of KW+1 words
    unsigned long long* spill_n;   // n_shards counters
    uint64_t spill_cap;
    uint32_t bins_per_shard_log2;
    uint32_t tile_words;           // packed words a block handles between two run-rotation barriers
    unsigned int* overflow;        // set to 1 when a spill list ran out of room: the chunk's bins are incomplete
};

constexpr int kMaxParts = 4096;
constexpr uint64_t kHole = ~0ULL;                  // word 0 of an unused run entry; real hashes equal to it are never binned

// Phase A, single sweep.  Every block owns, per bin, TWO private runs of pv.run entries inside the bin (each
// reserved with one global atomicAdd on the bin cursor): the current one and the next one.  A k-mer takes the
// next free entry with one shared-memory atomicAdd on the bin's fill counter (fill < run -> current run,
// fill < 2*run -> next run).  Runs are only rotated at the block-wide barrier between tiles, so the bases a
// thread reads after its atomicAdd are always the ones its index refers to.  A tile brings ~run/4 k-mers per
// bin, so both runs running out inside one tile is a tail event; those k-mers take single entries straight
// from the global cursor.  Unused tails of the runs a block still owns at the end are filled with holes.
template <int KW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_partition_reads(const __grid_constant__ TableView tv,
                                                                   const __grid_constant__ PartView pv,
                                                                   const uint64_t* __restrict__ packed,
                                                                   const uint32_t* __restrict__ ends, uint64_t w_begin,
                                                                   uint64_t w_end, uint64_t n_words, uint64_t n_bases) {
    constexpr int NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);
    const uint64_t kTileWords = pv.tile_words;
    constexpr unsigned int kNoRun = 0xffffffffu;
    // bases of the current (.x) and the next (.y) run of every bin as one 8-byte word: one LDS per k-mer; the
    // kernel is bound by the shared-memory pipe (ATOMS + LDS + the global store), not by occupancy
    __shared__ uint2 run_base[kMaxParts];
    __shared__ unsigned int run_fill[kMaxParts];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;
    const uint32_t R = pv.run;
    const bool hole_possible = KW > 1 || tv.hp.nbits == 64;   // for 2k < 64 no hash has all 64 bits set

    // k-mers that do not go through a bin (groups the extractor already aggregated, k-mers whose bin is full, a
    // hash equal to the hole marker) become (hash, count) records of the owner's spill list.  The kernel never
    // touches the table: if a spill list runs out of room it raises pv.overflow and the
    // chunk is redone — by the fused kernel on a single GPU, in smaller pieces by the multi-GPU host loop.
    auto cold = [&](const Key<KW>& H, uint64_t cnt) {
        const uint32_t owner = (uint32_t)(((H.w[0] & tv.lbg_mask) >> pv.pshift) & pv.pmask) >> pv.bins_per_shard_log2;
        const unsigned long long at = atomicAdd(pv.spill_n + owner, 1ULL);
        if (at >= pv.spill_cap) { *pv.overflow = 1u; return; }
        uint64_t* dst = pv.spill + ((uint64_t)owner * pv.spill_cap + at) * (KW + 1);
#pragma unroll
        for (int j = 0; j < KW; ++j) dst[j] = H.w[j];
        dst[KW] = cnt;
    };
    auto reserve = [&](uint32_t p) -> unsigned int {   // one run, or kNoRun when the bin is (nearly) full
        const unsigned long long nb = atomicAdd(pv.cursor + p, (unsigned long long)R);
        if (nb + R <= pv.cap) return (unsigned int)nb;
        for (unsigned long long j = nb; j < pv.cap; ++j) __stcg(pv.buf + ((uint64_t)p * pv.cap + j) * KW, kHole);
        return kNoRun;
    };
    auto fill_holes = [&](uint32_t p, unsigned int base, unsigned int from) {
        if (base == kNoRun) return;
        uint64_t* dst = pv.buf + ((uint64_t)p * pv.cap + base) * KW;
        for (unsigned int j = from; j < R; ++j) __stcg(dst + (uint64_t)j * KW, kHole);
    };

    for (uint32_t p = threadIdx.x; p < pv.P; p += blockDim.x) {
        const unsigned int c = reserve(p), n = reserve(p);
        run_base[p] = make_uint2(c, n); run_fill[p] = 0;
    }
    __syncthreads();

    // One k-mer per lane is kept "in flight": its shared-memory atomicAdd is issued when the k-mer is produced,
    // the returned index is consumed (bases read, hash stored) only when the NEXT k-mer of the lane has issued
    // its own atomic, so the ATOMS round trip overlaps a whole extraction + hash step.
    Key<KW> pend_h;
#pragma unroll
    for (int j = 0; j < KW; ++j) pend_h.w[j] = 0;
    uint32_t pend_p = 0;
    unsigned int pend_idx = 0, pend_cur = 0, pend_next = 0;
    bool pend = false;
    auto complete = [&]() {
        if (!pend) return;
        pend = false;
        const unsigned int rb = pend_idx < R ? pend_cur : pend_next;
        uint64_t pos;
        if (pend_idx < 2 * R && rb != kNoRun) pos = (uint64_t)rb + (pend_idx < R ? pend_idx : pend_idx - R);
        else pos = atomicAdd(pv.cursor + pend_p, 1ULL);   // both runs used up in one tile
        if (pos < pv.cap) {
            uint64_t* dst = pv.buf + ((uint64_t)pend_p * pv.cap + pos) * KW;
#pragma unroll
            for (int j = 0; j < KW; ++j) __stcg(dst + j, pend_h.w[j]);
        } else {
            cold(pend_h, 1);                              // bin full: never dropped
        }
    };

    for (uint64_t tile = w_begin + (uint64_t)blockIdx.x * kTileWords; tile < w_end; tile += (uint64_t)gridDim.x * kTileWords) {
        const uint64_t tile_end = tile + kTileWords < w_end ? tile + kTileWords : w_end;
        for (uint64_t base = tile + wib * 32; base < tile_end; base += (THREADS / 32) * 32) {
            uint64_t win[KW + 1];
            uint32_t ewin[NE + 1];
            load_window<KW, uint64_t>(packed, base, n_words, lane, win);
            load_window<NE, uint32_t>(ends, base, n_words, lane, ewin);
            const uint64_t limit = (base + lane < tile_end) ? n_bases : 0;  // lanes past the chunk emit nothing
            for_each_kmer_group<KW, false>(win, ewin[0], first_end_after<NE>(ewin), base + lane, limit, tv.L.k, tv.hp,
                                           [&](const Key<KW>& key, uint64_t cnt) {
                                               const Key<KW> H = hash_key<KW>(key, tv.hp);
                                               if (cnt >= 2 || (hole_possible && H.w[0] == kHole)) { cold(H, cnt); return; }
                                               const uint32_t p = (uint32_t)((H.w[0] & tv.lbg_mask) >> pv.pshift) & pv.pmask;
                                               const unsigned int idx = atomicAdd(&run_fill[p], 1u);
                                               const uint2 rb2 = run_base[p];                          // stable until the barrier
                                               const unsigned int rc = rb2.x, rn = rb2.y;
                                               complete();                 // the previous k-mer of this lane
                                               pend_h = H; pend_p = p; pend_idx = idx; pend_cur = rc; pend_next = rn; pend = true;
                                           });
        }
        complete();   // the runs must not rotate under an index that is still in flight
        __syncthreads();
        // rotate the runs whose current one was used up during this tile; a thread issues its reservations four
        // at a time (independent global atomics), then consumes them
        constexpr int kBatch = 4;
        for (uint32_t p0 = threadIdx.x; p0 < pv.P; p0 += kBatch * THREADS) {
            unsigned long long nb[kBatch];
            bool rot[kBatch];
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                const uint32_t p = p0 + i * THREADS;
                rot[i] = p < pv.P && run_fill[p] >= R;
                nb[i] = rot[i] ? atomicAdd(pv.cursor + p, (unsigned long long)R) : 0ULL;
            }
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
                if (!rot[i]) continue;
                const uint32_t p = p0 + i * THREADS;
                const unsigned int f = run_fill[p];
                run_fill[p] = (f < 2 * R ? f : 2 * R) - R;
                unsigned int fresh = kNoRun;
                if (nb[i] + R <= pv.cap) fresh = (unsigned int)nb[i];
                else for (unsigned long long j = nb[i]; j < pv.cap; ++j) __stcg(pv.buf + ((uint64_t)p * pv.cap + j) * KW, kHole);
                run_base[p] = make_uint2(run_base[p].y, fresh);
            }
        }
        __syncthreads();
    }
    for (uint32_t p = threadIdx.x; p < pv.P; p += blockDim.x) {
        const unsigned int f = run_fill[p];
        fill_holes(p, run_base[p].x, f < R ? f : R);
        fill_holes(p, run_base[p].y, f < R ? 0u : (f < 2 * R ? f - R : R));
    }
}

// Phase A, static variant (EXPERIMENTAL, selected with TSXC_PART_STATIC=1; not validated on hardware yet — see
// DESIGN.md §9 item 2a).  Every block owns ONE slab of pv.cap entries per bin for the whole chunk, laid out
// [block][bin][pv.cap], so a k-mer costs one shared-memory atomicAdd and one store: no run bases to load, no run
// rotation, no barrier between tiles, no holes, and a block's stores stay inside its own P * cap * 8 * KW byte
// window.  The fill counts go to pv.cursor[block * P + bin]; phase B drains the slabs as n_sources = gridDim.x
// sources.  A slab that runs full sends its k-mers to the spill list like a full bin does.
template <int KW, int THREADS>
__global__ void __launch_bounds__(THREADS) k_partition_reads_static(const __grid_constant__ TableView tv,
                                                                          const __grid_constant__ PartView pv,
                                                                          const uint64_t* __restrict__ packed,
                                                                          const uint32_t* __restrict__ ends, uint64_t w_begin,
                                                                          uint64_t w_end, uint64_t n_words, uint64_t n_bases) {
    constexpr int NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);
    const uint64_t kTileWords = pv.tile_words;
    __shared__ unsigned int fill[kMaxParts];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned wib = threadIdx.x >> 5;
    const bool hole_possible = KW > 1 || tv.hp.nbits == 64;   // phase B skips entries whose word 0 is the hole marker
    uint64_t* const slab = pv.buf + (uint64_t)blockIdx.x * pv.P * pv.cap * KW;

    auto cold = [&](const Key<KW>& H, uint64_t cnt) {
        const uint32_t owner = (uint32_t)(((H.w[0] & tv.lbg_mask) >> pv.pshift) & pv.pmask) >> pv.bins_per_shard_log2;
        const unsigned long long at = atomicAdd(pv.spill_n + owner, 1ULL);
        if (at >= pv.spill_cap) { *pv.overflow = 1u; return; }
        uint64_t* dst = pv.spill + ((uint64_t)owner * pv.spill_cap + at) * (KW + 1);
#pragma unroll
        for (int j = 0; j < KW; ++j) dst[j] = H.w[j];
        dst[KW] = cnt;
    };

    for (uint32_t p = threadIdx.x; p < pv.P; p += blockDim.x) fill[p] = 0;
    __syncthreads();

    // one k-mer per lane in flight, as in k_partition_reads: the ATOMS round trip overlaps the next extraction
    Key<KW> pend_h;
#pragma unroll
    for (int j = 0; j < KW; ++j) pend_h.w[j] = 0;
    uint32_t pend_p = 0;
    unsigned int pend_idx = 0;
    bool pend = false;
    auto complete = [&]() {
        if (!pend) return;
        pend = false;
        if ((uint64_t)pend_idx < pv.cap) {
            uint64_t* dst = slab + ((uint64_t)pend_p * pv.cap + pend_idx) * KW;
#pragma unroll
            for (int j = 0; j < KW; ++j) __stcg(dst + j, pend_h.w[j]);
        } else {
            cold(pend_h, 1);
        }
    };

    for (uint64_t tile = w_begin + (uint64_t)blockIdx.x * kTileWords; tile < w_end; tile += (uint64_t)gridDim.x * kTileWords) {
        const uint64_t tile_end = tile + kTileWords < w_end ? tile + kTileWords : w_end;
        for (uint64_t base = tile + wib * 32; base < tile_end; base += (THREADS / 32) * 32) {
            uint64_t win[KW + 1];
            uint32_t ewin[NE + 1];
            load_window<KW, uint64_t>(packed, base, n_words, lane, win);
            load_window<NE, uint32_t>(ends, base, n_words, lane, ewin);
            const uint64_t limit = (base + lane < tile_end) ? n_bases : 0;
            for_each_kmer_group<KW, false>(win, ewin[0], first_end_after<NE>(ewin), base + lane, limit, tv.L.k, tv.hp,
                                           [&](const Key<KW>& key, uint64_t cnt) {
                                               const Key<KW> H = hash_key<KW>(key, tv.hp);
                                               if (cnt >= 2 || (hole_possible && H.w[0] == kHole)) { cold(H, cnt); return; }
                                               const uint32_t p = (uint32_t)((H.w[0] & tv.lbg_mask) >> pv.pshift) & pv.pmask;
                                               const unsigned int idx = atomicAdd(&fill[p], 1u);
                                               complete();                 // the previous k-mer of this lane
                                               pend_h = H; pend_p = p; pend_idx = idx; pend = true;
                                           });
        }
    }
    complete();
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < pv.P; p += blockDim.x) {
        const unsigned int f = fill[p];
        pv.cursor[(uint64_t)blockIdx.x * pv.P + p] = (uint64_t)f < pv.cap ? f : pv.cap;
    }
}

// phase B: work item = (region, source, slice of slice_entries entries); items are numbered region-major and
// handed out in order through `ticket`.  Slices are SMALL (1 K entries) on purpose: the ~1200 resident blocks then work on
// one or two table regions at a time, and the K0r microbenchmark (tools/k0region.py) shows that concentrating
// all SMs on a 16-64 MiB region lifts the dependent load+atomic rate from 21 G/s (8 K-entry items, ~18 regions
// in flight) to 33 G/s at 0.25 touches per sector and to 56-65 G/s at 0.65: neighbouring sectors are requested
// close together in time (DRAM row locality) and sectors touched twice are still in L2.
constexpr uint32_t kSliceEntriesDefault = 1024;

template <int KW, int W, bool WARP_AGG>
__global__ void __launch_bounds__(kBlockThreads) k_insert_partitions(const __grid_constant__ TableView tv,
                                                                     const __grid_constant__ PartView pv,
                                                                     uint32_t slices_per_bin, uint32_t slice_entries,
                                                                     uint32_t n_sources,
                                                                     unsigned long long* __restrict__ ticket,
                                                                     const unsigned int* __restrict__ skip_if) {
    const unsigned full = 0xffffffffu;
    __shared__ unsigned long long item_s;
    if (skip_if && __ldcg(skip_if) != 0) return;     // the chunk overflowed its spill list and is being redone
    LocalStats st;
    const unsigned long long n_items = (unsigned long long)pv.P * slices_per_bin;
    while (true) {
        if (threadIdx.x == 0) item_s = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const unsigned long long item = item_s;
        __syncthreads();
        if (item >= n_items) break;
        // bins are stored source-major (bin = source * regions + region) but drained REGION-major: all sources'
        // bins of one table region are handled back to back, so a region is visited once per chunk
        const uint32_t regions = pv.P / n_sources;
        const unsigned long long per_region = (unsigned long long)n_sources * slices_per_bin;
        const uint32_t region = (uint32_t)(item / per_region);
        const uint32_t source = (uint32_t)((item % per_region) / slices_per_bin);
        const uint32_t p = source * regions + region;
        const uint64_t lo = (uint64_t)(item % slices_per_bin) * slice_entries;
        unsigned long long n = __ldcg(pv.cursor + p);
        if (n > pv.cap) n = pv.cap;
        if (lo >= n) continue;
        const uint64_t hi = lo + slice_entries < n ? lo + slice_entries : n;
        const uint64_t* src = pv.buf + (uint64_t)p * pv.cap * KW;
        for (uint64_t i0 = lo; i0 < hi; i0 += blockDim.x) {
            const uint64_t i = i0 + threadIdx.x;
            const bool valid = i < hi;
            Key<KW> H;
#pragma unroll
            for (int j = 0; j < KW; ++j) H.w[j] = valid ? __ldcs(src + i * KW + j) : 0ULL;
            const bool live = valid && H.w[0] != kHole;
            const unsigned vmask = __ballot_sync(full, live);
            if (!live) continue;
            uint64_t cnt = 1;
            bool lead = true;
            if (WARP_AGG) {
                unsigned peers = __match_any_sync(vmask, H.w[0]);
#pragma unroll
                for (int j = 1; j < KW; ++j) peers &= __match_any_sync(vmask, H.w[j]);
                cnt = (uint64_t)__popc(peers);
                lead = (unsigned)(__ffs(peers) - 1) == (threadIdx.x & 31u);
            }
            if (lead) insert_hashed<KW, W>(tv, H, cnt, st);
        }
    }
    flush_stats(tv, st);
}

// (hash, count) records: the spill lists of the routing path
// n_dev != nullptr: the record count lives on the device (min(*n_dev, n) records are read)
template <int KW, int W>
__global__ void __launch_bounds__(kBlockThreads) k_add_hash_counts(const __grid_constant__ TableView tv,
                                                                   const uint64_t* __restrict__ rec, uint64_t n,
                                                                   const unsigned long long* __restrict__ n_dev,
                                                                   const unsigned int* __restrict__ skip_if) {
    if (skip_if && __ldcg(skip_if) != 0) return;
    if (n_dev) { const unsigned long long m = __ldcg(n_dev); if (m < n) n = m; }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    LocalStats st;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Key<KW> H;
#pragma unroll
        for (int j = 0; j < KW; ++j) H.w[j] = __ldg(rec + i * (KW + 1) + j);
        insert_hashed<KW, W>(tv, H, __ldg(rec + i * (KW + 1) + KW), st);
    }
    flush_stats(tv, st);
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
    while (true) {
        if (threadIdx.x == 0) item_s = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const unsigned long long item = item_s;
        __syncthreads();
        if (item >= n_items) break;
        const uint64_t base = (item / items_per_region) * region_words;
        for (uint32_t i = threadIdx.x; i < ops_per_item; i += blockDim.x) {
            const uint64_t a = base + (fmix64((item * ops_per_item + i) * kC1 + 0x1234567ULL) & wmask);
            if (mode == 0) {
                atomicAdd((unsigned long long*)(words + a), 1ULL);
            } else {
                uint64_t w[4];
                load_bucket(words + (a & ~3ULL), w);
                const uint64_t pick = (w[0] ^ w[1] ^ w[2] ^ w[3]) == 0x5a5a5a5a5a5a5a5aULL ? 1 : 0;
                atomicAdd((unsigned long long*)(words + ((a & ~3ULL) | ((a + pick) & 3ULL))), 1ULL << 40);
            }
        }
    }
}

}  // namespace tsx
