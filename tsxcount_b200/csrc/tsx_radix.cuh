// tsx_radix.cuh — the region-sorted insert pipeline for tables far larger than L2 / TLB reach.
//
// Measured on B200 (profiles/r01_k0_random_access.md, profiles/r02_k0r_modes.md): a dependent sector load + CAS
// runs at 4 G/s when spread uniformly over 128 GiB, at 37 G/s when all SMs work inside one 128 MiB region at a
// time and at 45 G/s (the ceiling of the atomic-return path) for regions of 32 MiB and less.  So k-mer hashes are
// sorted by table region first and inserted in region order afterwards.  One partition pass with up to 1024 bins
// (128 MiB regions of a 128 GiB table) beats two passes with 2^13 bins: the second pass costs 73 ms per 8e9 k-mers,
// the coarser regions cost the insert 30 ms (profiles/r02_pipeline_history.md).
//
//   plan  k_count_segs   k-mers per segment of the read stream (read-end bitmap only, no hashing)
//         k_plan_chunks  device-side planner: cuts the batch into chunks of at most `cap` k-mers
//   S1    k_part_reads   extract + hash every k-mer, block-local counting sort of a tile by digit 1 in shared memory,
//                        coalesced runs into buffer A.  Single GPU: A is a pool of pages; a bin's run takes the
//                        next positions of the bin's virtual address space (one atomicAdd per (tile, bin)) and the
//                        reservation that crosses into a new page allocates it, so bins grow as the data demands:
//                        no histogram pass, no bin can overflow whatever the skew.  Multi-GPU: exact offsets from
//                        an all-gathered histogram (S0 = k_hist_reads), and the runs go straight into the owning
//                        rank's peer-mapped receive buffer
//   B     k_insert_keys  blocks take consecutive slices of the bins through a ticket, so the whole chip probes one or
//                        two table regions at a time; duplicates inside a slice are combined in shared memory
//                        before they reach the table (heavy hitters cost one atomic per slice, not one per k-mer)
//
// Ranking inside a tile uses warp-private counters and optimistic conflict detection (rank_in_warp); no
// shared-memory atomics: at 2 cycles per lane they would cost more than everything else in these kernels together.
// The two-digit machinery (k_hist_keys / k_part_keys) is what region-sorted lookups still use.
//
// Reference semantics: extraction src/mains/testExecution.h:15-36, encoding src/utils/SequenceUtils.h:86-123,
// insert src/tsxcount/TSXHashMapPerf.h:56-205 (paths relative to mjoppich/tsxCount).  Nothing here has a
// counterpart in the reference: it inserts k-mer by k-mer in input order (src/mains/main.cpp:159-192).
#pragma once

#include <cstdint>

#include "tsx_kernels.cuh"
#include "tsx_table.cuh"

namespace tsx {

constexpr int kRadixThreads = 512;
constexpr int kRadixWarps = kRadixThreads / 32;
constexpr int kNB = 256;                    // bins per digit of the two-digit passes (lookups)
constexpr int kNB1 = 1024;                  // bins of S1 (digit 1 is at most 10 bits wide)
constexpr int kCursorStride = 16;           // S1 cursors sit in their own 128-byte lines (unsigned long long units)
constexpr int kSegWordsLog2Default = 15;    // planner granularity: 2^15 packed words = 2^20 base positions
constexpr int kSegWordsLog2Min = 9;         // one block round (16 warps x 32 words)
constexpr int kMaxChunks = 64;
constexpr int kMaxFine = kNB * kNB;
constexpr int kCombSlots = 1024;            // shared-memory combiner of phase B
constexpr int kPageLog2Max = 16;            // keys per page of the paged buffer A, log2 (>= log2 of a tile)

// Digit geometry, derived once per handle on the host.
//   global bucket index bg = H.w[0] & lbg_mask  (LBg bits; its top shard_bits select the owning shard)
//   digit 1 = top d1 bits of bg                 (includes the owner bits: bins of S1 are owner-major)
//   local coarse bin = digit 1 without the owner bits (nbl = 2^(d1 - shard_bits) per shard)
//   digit 2 = the next d2 bits; fine bin (local) = local coarse bin * nb2 + digit 2      (lookups only)
struct RadixGeom {
    uint32_t d1, d2;
    uint32_t nb1, nb2, nbl;
    uint32_t shift1, shift2;     // bg >> shift1 = digit 1 ; (bg >> shift2) & (nb2-1) = digit 2
    uint32_t owner_shift;        // digit 1 >> owner_shift = owning shard
    uint32_t seg_log2;           // packed words per planner segment, log2 (>= kSegWordsLog2Min)
};

struct ChunkDesc {
    uint64_t seg_begin, seg_end;           // segments [seg_begin, seg_end) of the batch
    uint64_t n_keys;
    uint64_t coff[kNB1 + 1];               // exclusive prefix of the digit-1 counts (exact mode only)
};

struct GroupDesc {
    uint64_t a0, a1;                       // keys [a0, a1) of A
    uint32_t n_items, active;
    uint32_t itemstart[kNB + 1];           // tiles of local coarse bin b are items [itemstart[b], itemstart[b+1])
};

// Buffer A of a single shard is a pool of small pages (as many keys as a slice of phase B, 1024 at k <= 32).  In S1
// every (thread block, bin) fills a page of its own and takes the next one from the pool when it is full (one
// atomicAdd per page; the pages a tile run needs beyond the current one are taken together and are contiguous).
// Bins therefore grow as the data demands: no histogram pass, no bin can overflow whatever the skew, and the
// reservation traffic is one returning atomic per 1024 k-mers instead of one per (tile, bin) = per 4 k-mers, which
// at 1024 bins costs 20 ms per 8e9 k-mers (the atomic-return path again, profiles/r02_pipeline_history.md).
// page_bin[p] = bin of page p (0xffff: unused), page_len[p] = keys in it.
struct PageGeom {
    uint32_t paged;              // 0: bins have exact offsets in A, 1: paged pool
    uint32_t page_log2;          // keys per page, log2 (at most a slice of phase B)
    uint32_t n_pages;            // pages in the pool
    uint32_t pad;
};

struct RadixCtl {
    uint32_t n_chunks, chunk_active;
    uint32_t plan_error, recv_overflow;
    unsigned long long ticket[4];          // 0: S2a, 1: S2b, 2: phase B
    unsigned long long n_insert;           // keys phase B has to insert
    unsigned long long page_next;          // paged mode: pages handed out in this chunk
    uint32_t n_slices, pad0;               // work items of phase B
    uint32_t slicestart[kNB1 + 1];         // slices of local bin b are items [slicestart[b], slicestart[b+1])
    uint32_t bin_pages[kNB1];              // paged mode: pages of bin b (S1), then the fill cursor of k_build_slices
    uint64_t cur_coff[kNB1 + 1];           // exclusive prefix of the local bin counts (= offsets inside A in exact mode)
    unsigned long long cursor1[kNB1 * kCursorStride];   // S1 reservation cursors
    GroupDesc group;
    ChunkDesc chunk[kMaxChunks];
};

#if defined(__CUDACC__)

template <int KW> struct RadixCfg {
    static constexpr int OPT = 8 / KW;                       // keys per thread per tile
    static constexpr int TILE = kRadixThreads * OPT;         // 4096 / 2048 / 1024 keys = 32 KB of shared memory
    static constexpr int MINB = KW == 4 ? 1 : 2;
};

// digit 1 = bits [shift1, shift1 + d1) of hash word 0 (the top of the bucket index), digit 2 = the d2 bits below it:
// one funnel shift and one AND each
__device__ __forceinline__ uint32_t digit1_of(const RadixGeom& rg, uint64_t, uint64_t h0) {
    return (uint32_t)(h0 >> rg.shift1) & (rg.nb1 - 1);
}
__device__ __forceinline__ uint32_t digit2_of(const RadixGeom& rg, uint64_t, uint64_t h0) {
    return (uint32_t)(h0 >> rg.shift2) & (rg.nb2 - 1);
}

// ---- one lane's view of the packed read stream ---------------------------------------------------------------
// Lane `lane` of a warp owns stream word base+lane and enumerates the k-mers that START in it, offsets 31..0
// (descending: the distance to the next read end is carried from the right).
template <int KW>
struct KmerLane {
    static constexpr int NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);
    uint32_t a[2 * KW + 2];      // the lane's window of the stream as 32-bit words
    uint32_t ends_cur, dist;
    int omax;                    // k-mers may start at offsets <= omax of this word (-1: none: end of stream, or lane out of range)
    __device__ __forceinline__ void load(const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ends,
                                         uint64_t base, uint64_t n_words, uint64_t w_end, uint64_t n_bases, unsigned lane, uint32_t k) {
        uint64_t win[KW + 1];
        uint32_t ewin[NE + 1];
        load_window<KW, uint64_t>(packed, base, n_words, lane, win);
        load_window<NE, uint32_t>(ends, base, n_words, lane, ewin);
#pragma unroll
        for (int j = 0; j <= KW; ++j) { a[2 * j] = (uint32_t)win[j]; a[2 * j + 1] = (uint32_t)(win[j] >> 32); }
        ends_cur = ewin[0];
        dist = first_end_after<NE>(ewin);
        const uint64_t g0 = (base + lane) << 5;
        omax = -1;
        if (base + lane < w_end && g0 + k <= n_bases) omax = n_bases - g0 - k < 31 ? (int)(n_bases - g0 - k) : 31;
    }
    // HI: o >= 16, i.e. the k-mer starts in the upper half of the 64-bit word (callers walk o downwards in groups that
    // never straddle 16, so the word selection is a compile-time matter and a k-mer costs two funnel shifts per word)
    template <bool HI>
    __device__ __forceinline__ bool kmer_at(int o, uint32_t k, const HashParams& hp, Key<KW>& key) {
        dist = ((ends_cur >> o) & 1u) ? 0u : (dist == 0xffffffffu ? dist : dist + 1u);
        const bool valid = (dist >= k - 1) && (o <= omax);
        const unsigned sb = (2u * (unsigned)o) & 31u;
        constexpr int H = HI ? 1 : 0;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            const uint32_t lo = __funnelshift_r(a[2 * j + H], a[2 * j + 1 + H], sb);
            const uint32_t hi = __funnelshift_r(a[2 * j + 1 + H], a[2 * j + 2 + H], sb);
            key.w[j] = (((uint64_t)hi << 32) | lo) & word_mask<KW>(j, hp);
        }
        return valid;
    }
};

// ---- sparse streams: enumerate only the positions that start a k-mer ---------------------------------------------
// With reads of 150 bases only 24 of 150 positions start a 127-mer.  The dense walk above hashes and ranks every
// position and hands tiles that are 16 % full to tile_partition, whose per-tile cost does not depend on the fill
// (config 4: 204 ms of S1 for 1.6e9 k-mers).  The sparse walk first turns the read-end bitmap of a block round
// (512 stream words) into one validity mask per word, scans the popcounts, and then every thread takes OPT consecutive
// valid positions of the round: the first by binary search over the scanned counts + select of the i-th set bit of the
// word's mask, the following ones by stepping through the masks.
// Tiles are full, and only real k-mers are extracted and hashed.  The helpers are host-callable so that
// tsxc_debug_sparse_round runs the same arithmetic on the CPU (tests/test_host.py).
TSX_HD uint32_t popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
TSX_HD uint32_t funnel_r32(uint32_t lo, uint32_t hi, uint32_t s) {   // bits [s, s+32) of hi:lo, 0 <= s < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31u));
#endif
}

// Mask of the offsets o of one stream word at which a k-mer starts.  ends_cur: the word's read-end bits; dist_after:
// distance from bit 0 of the NEXT word to the first read end at or after it (0xffffffff: none within reach);
// omax: largest offset that still leaves k bases in the stream (-1: none).  The recurrence is KmerLane::kmer_at's.
TSX_HD uint32_t valid_starts(uint32_t ends_cur, uint32_t dist_after, uint32_t k, int omax) {
    uint32_t dist = dist_after, vbits = 0;
#pragma unroll 8
    for (int o = 31; o >= 0; --o) {
        dist = ((ends_cur >> o) & 1u) ? 0u : (dist == 0xffffffffu ? dist : dist + 1u);
        if (dist >= k - 1 && o <= omax) vbits |= 1u << o;
    }
    return vbits;
}

// Position of the r-th (0-based) set bit of m; m has more than r set bits.
TSX_HD uint32_t select_bit(uint32_t m, uint32_t r) {
    uint32_t pos = 0, c;
    c = popc32(m & 0xffffu); if (r >= c) { pos += 16; r -= c; m >>= 16; }
    c = popc32(m & 0xffu);   if (r >= c) { pos += 8;  r -= c; m >>= 8; }
    c = popc32(m & 0xfu);    if (r >= c) { pos += 4;  r -= c; m >>= 4; }
    c = popc32(m & 0x3u);    if (r >= c) { pos += 2;  r -= c; m >>= 2; }
    c = m & 1u;              if (r >= c) { pos += 1; }
    return pos;
}

// pre[0..n) = exclusive scan of the words' popcounts (n a power of two), i < total: the word that holds the i-th valid
// position = the largest w with pre[w] <= i (words without valid positions share their successor's prefix and lose).
TSX_HD uint32_t locate_word(const uint32_t* pre, uint32_t n, uint32_t i) {
    uint32_t w = 0;
    for (uint32_t step = n >> 1; step; step >>= 1)
        if (pre[w + step] <= i) w += step;
    return w;
}

// A thread takes OPT consecutive valid positions: it seeks the first one (binary search + select) and steps to the
// following ones with one find-first-set each.
struct SparseCursor { uint32_t w, m; };      // word, and the word's valid starts at and above the cursor
TSX_HD uint32_t ctz32(uint32_t x) {          // x != 0
#if defined(__CUDA_ARCH__)
    return (uint32_t)(__ffs((int)x) - 1);
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
TSX_HD SparseCursor sparse_seek(const uint32_t* pre, const uint32_t* vb, uint32_t n, uint32_t i) {
    SparseCursor c;
    c.w = locate_word(pre, n, i);
    c.m = vb[c.w] & (0xffffffffu << select_bit(vb[c.w], i - pre[c.w]));
    return c;
}
// offset of the valid position under the cursor (its word is c.w afterwards); the cursor moves on.  The caller knows
// that a valid position is left at or after the cursor.
TSX_HD uint32_t sparse_next(SparseCursor& c, const uint32_t* vb) {
    while (c.m == 0) c.m = vb[++c.w];
    const uint32_t o = ctz32(c.m);
    c.m &= c.m - 1;
    return o;
}

// The k-mer that starts at offset o of word w of a staged piece of the stream (s32: its 32-bit halves, little end first;
// words w .. w+KW are read).  Same funnel shifts as KmerLane::kmer_at with the half chosen at run time.
template <int KW>
TSX_HD Key<KW> kmer_from_stream32(const uint32_t* s32, uint32_t w, uint32_t o, const HashParams& hp) {
    const uint32_t* a = s32 + 2 * w + (o >> 4);
    const uint32_t sb = (2u * o) & 31u;
    Key<KW> key;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const uint32_t lo = funnel_r32(a[2 * j], a[2 * j + 1], sb);
        const uint32_t hi = funnel_r32(a[2 * j + 1], a[2 * j + 2], sb);
        key.w[j] = (((uint64_t)hi << 32) | lo) & word_mask<KW>(j, hp);
    }
    return key;
}

// One block round of the sparse walk in shared memory
template <int KW>
struct SparseStage {
    alignas(16) uint32_t stream[2 * (kRadixThreads + KW + 1)];   // words [round, round + 512 + KW] of the stream
    uint32_t vb[kRadixThreads];                                  // valid_starts of word round + t
    uint32_t pre[kRadixThreads];                                 // exclusive scan of their popcounts
};

// A chunk takes the sparse walk when fewer than pct % of its base positions start a k-mer (both numbers come from the
// planner: k-mers of the chunk, segments of the chunk).  Uniform over the whole launch.
__device__ __forceinline__ bool chunk_is_sparse(const RadixCtl* __restrict__ ctl, uint32_t c, const RadixGeom& rg, uint32_t pct) {
    const uint64_t positions = (ctl->chunk[c].seg_end - ctl->chunk[c].seg_begin) << (rg.seg_log2 + 5);
    return ctl->chunk[c].n_keys * 100ULL < positions * (uint64_t)pct;
}

// ---- warp-private ranking -------------------------------------------------------------------------------------
// Lanes holding the same digit form a group; the group's first lane bumps the warp's own counter by the group
// size, every lane's rank is the old counter value + its position inside the group.  No atomics: the counter row
// belongs to this warp alone.  (MATCH.ANY instead of one ballot per digit bit was measured and dropped: it keeps the
// XU pipe 93 % busy.)
__device__ __forceinline__ unsigned match_digit(bool valid, uint32_t digit, uint32_t bits) {
    const unsigned full = 0xffffffffu;
    // one ballot per digit bit; lanes whose bit is clear take the complement.  Hand-written so that a bit costs four
    // instructions (test -> predicate, vote, two predicated ANDs): these kernels are bound by the integer ALU pipe.
    unsigned peers = __ballot_sync(full, valid);
#pragma unroll
    for (uint32_t b = 0; b < 10; ++b) {
        if (b >= bits) break;
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     ".reg .b32 t, v;\n\t"
                     "and.b32 t, %1, %2;\n\t"
                     "setp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
                     "@p and.b32 %0, %0, v;\n\t"
                     "@!p lop3.b32 %0, %0, v, 0, 0x30;\n\t"     // a & ~b
                     "}"
                     : "+r"(peers) : "r"(digit), "r"(1u << b));
    }
    return valid ? peers : 0u;
}

#ifndef TSX_RANK_CD
#define TSX_RANK_CD 1
#endif
// Digits of 7-8 bits (128-256 bins) rarely collide inside a warp, so two optimistic rounds come first: every
// pending lane writes its lane id to tag[digit], the lane that reads its own id back owns the counter for this
// round, takes the old value as its rank and bumps it.  Lanes still pending after two rounds (three or more lanes
// share a digit: 8 % of the warp steps at 256 bins, nearly always at 64) are ranked by the ballot match.  Ranks are
// unique per (warp, digit), which is all the counting sort needs (keys are a multiset, stability is irrelevant).
// ~30 instructions per step instead of ~85: the partition kernels are bound by the integer pipe.
template <typename CT>
__device__ __forceinline__ uint32_t rank_in_warp(CT* wcnt, uint8_t* wtag_, bool valid, uint32_t digit, unsigned lane, uint32_t bits) {
    volatile uint8_t* wtag = wtag_;      // written and read back by different lanes between two __syncwarp()
    volatile CT* vcnt = wcnt;
    const unsigned full = 0xffffffffu;
    uint32_t rank = 0;
    bool pend = valid;
#if TSX_RANK_CD
    if (bits >= 7) {
#pragma unroll
        for (int round = 0; round < 2; ++round) {
            if (pend) wtag[digit] = (uint8_t)lane;
            __syncwarp();
            if (pend && wtag[digit] == (uint8_t)lane) {
                rank = vcnt[digit];
                vcnt[digit] = (CT)(rank + 1);
                pend = false;
            }
            __syncwarp();
        }
        if (!__any_sync(full, pend)) return rank;
    }
#else
    (void)wtag; (void)vcnt;
#endif
    const unsigned peers = match_digit(pend, digit, bits);
    const int leader = pend ? (__ffs(peers) - 1) : (int)lane;
    uint32_t old = 0;
    if (pend && (unsigned)leader == lane) {
        old = wcnt[digit];
        wcnt[digit] = (CT)(old + __popc(peers));
    }
    old = __shfl_sync(full, old, leader);
    __syncwarp();
    return pend ? old + __popc(peers & ((1u << lane) - 1u)) : rank;
}

// Exclusive scan of v over all threads of the block (any block size that is a multiple of 32, up to 1024).
// scratch: 32 words of shared memory.  Contains two __syncthreads; every thread of the block must call it.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* scratch, uint32_t* total) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(full, inc, d);
        if (lane >= (unsigned)d) inc += y;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    // every warp scans the (at most 32) warp totals again with shuffles instead of walking them
    uint32_t part = lane < nw ? scratch[lane] : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(full, part, d);
        if (lane >= (unsigned)d) part += y;
    }
    const uint32_t upto = __shfl_sync(full, part, warp ? warp - 1 : 0);
    const uint32_t tot = __shfl_sync(full, part, nw - 1);
    __syncthreads();
    if (total) *total = tot;
    return (warp ? upto : 0u) + inc - v;
}

// 64-bit variant for values below 2^40 each and at most 1024 threads: the low 20 bits and the rest are scanned
// separately (their prefix sums fit 32 bits) and recombined, which is exact.
__device__ __forceinline__ uint64_t block_exscan_u64(uint64_t v, uint32_t* scratch, uint64_t* total) {
    uint32_t tlo = 0, thi = 0;
    const uint32_t plo = block_exscan((uint32_t)(v & 0xfffffULL), scratch, &tlo);
    const uint32_t phi = block_exscan((uint32_t)(v >> 20), scratch, &thi);
    if (total) *total = ((uint64_t)thi << 20) + tlo;
    return ((uint64_t)phi << 20) + plo;
}

// Exclusive scan of v over the first 256 threads of the block (the others pass 0).  scratch: 8 words of shared
// memory.  Contains two __syncthreads; every thread of the block must call it.
__device__ __forceinline__ uint32_t block_exscan_256(uint32_t v, uint32_t* scratch, uint32_t* total) {
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(full, inc, d);
        if (lane >= (unsigned)d) inc += y;
    }
    if (warp < 8 && lane == 31) scratch[warp] = inc;
    __syncthreads();
    uint32_t pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = scratch[w];
        if ((unsigned)w < warp) pre += s;
        tot += s;
    }
    __syncthreads();
    if (total) *total = tot;
    return pre + inc - v;
}

// 64-bit variant for values below 2^40 each: the low 23 bits and the rest are scanned separately (their prefix
// sums over 256 entries fit 32 bits) and recombined, which is exact.
__device__ __forceinline__ uint64_t block_exscan_256_u64(uint64_t v, uint32_t* scratch, uint64_t* total) {
    uint32_t tlo = 0, thi = 0;
    const uint32_t plo = block_exscan_256((uint32_t)(v & 0x7fffffULL), scratch, &tlo);
    const uint32_t phi = block_exscan_256((uint32_t)(v >> 23), scratch, &thi);
    if (total) *total = ((uint64_t)thi << 23) + tlo;
    return ((uint64_t)phi << 23) + plo;
}

template <int KW, int NB, bool PAGED = false>
struct TileSmem {
    alignas(16) uint64_t sorted[RadixCfg<KW>::TILE * KW];
    alignas(16) uint16_t cnt[kRadixWarps][NB];   // zeroed with 128-bit stores
    long long gdelta[NB];             // position of a bin's run in its destination minus its start inside the tile
    long long gdelta2[PAGED ? NB : 2];   // paged: the same for the part of the run that lies in freshly taken pages
    uint32_t binstart[NB];
    uint32_t split[PAGED ? NB : 4];      // paged: tile index at which the run continues in the fresh pages
    uint32_t scratch[32];
};

// A thread block's view of the page pool: its current page of every bin.  Lives in shared memory; bin b is only ever
// touched by the thread that owns b in tile_partition (or, for the fine pass, by the owner of the tile-local bin that
// maps to it), and phases are separated by the barriers of tile_partition.
struct PageState {
    uint32_t cur[kNB1];               // the block's current page of the bin (0xffffffff: none yet)
    uint32_t fill[kNB1];              // keys already in it
    uint32_t cnt[kNB1];               // pages this block has taken for the bin
};

struct PagePool {
    PageState* ps;
    RadixCtl* ctl;
    uint16_t* page_bin;
    uint16_t* page_len;
    unsigned long long* err_ctr;
    uint32_t page_log2, n_pages;
    unsigned long long my_keys;       // keys this thread has placed

    __device__ __forceinline__ void init() {
        for (uint32_t b = threadIdx.x; b < kNB1; b += blockDim.x) { ps->cur[b] = 0xffffffffu; ps->fill[b] = 1u << page_log2; ps->cnt[b] = 0u; }
        my_keys = 0;
    }
    // only the part of a run that does not fit the block's current page of the bin needs fresh pages
    __device__ __forceinline__ unsigned long long reserve(uint32_t bin, uint32_t n) {
        const uint32_t room = (1u << page_log2) - ps->fill[bin];
        if (n <= room) return 0ULL;
        const uint32_t m = (n - room + (1u << page_log2) - 1) >> page_log2;
        return atomicAdd(&ctl->page_next, (unsigned long long)m);
    }
    // run of n keys of `bin` whose first key has tile index `start`; p = what reserve() returned.
    // g1: position minus tile index for keys before `split`, g2: the same for the keys in the fresh pages
    __device__ __forceinline__ void place(uint32_t bin, uint32_t n, uint32_t start, unsigned long long p, long long& g1, long long& g2,
                                          uint32_t& split) {
        const uint32_t page_keys = 1u << page_log2;
        const uint32_t c = ps->cur[bin], fill = ps->fill[bin];
        const uint32_t room = page_keys - fill;
        my_keys += n;
        g1 = (long long)(((uint64_t)c << page_log2) + fill) - (long long)start;      // unused when room == 0
        g2 = 0;
        if (n <= room) {
            split = 0xffffffffu;
            ps->fill[bin] = fill + n;
            return;
        }
        const uint32_t rest = n - room, m = (rest + page_keys - 1) >> page_log2;
        if (p + m > n_pages) { atomicOr(err_ctr, (unsigned long long)ERR_PLAN); p = 0; }   // cannot happen: the planner leaves room
        if (c != 0xffffffffu) page_len[c] = (uint16_t)page_keys;                            // the old page is completed by this run
        for (uint32_t q = 0; q < m; ++q) { page_bin[p + q] = (uint16_t)bin; if (q + 1 < m) page_len[p + q] = (uint16_t)page_keys; }
        split = start + room;
        g2 = (long long)(p << page_log2) - (long long)(start + room);
        ps->cur[bin] = (uint32_t)p + m - 1;
        ps->fill[bin] = rest - ((m - 1) << page_log2);
        ps->cnt[bin] += m;
    }
    // the block's last page of every bin is partly filled; its page counts and keys go to the totals
    __device__ __forceinline__ void finish() {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < kNB1; b += blockDim.x) {
            const uint32_t c = ps->cur[b];
            if (c != 0xffffffffu) page_len[c] = (uint16_t)ps->fill[b];
            if (ps->cnt[b]) atomicAdd(&ctl->bin_pages[b], ps->cnt[b]);
        }
        const unsigned full = 0xffffffffu;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) my_keys += __shfl_xor_sync(full, my_keys, sft);
        if ((threadIdx.x & 31u) == 0 && my_keys) atomicAdd(&ctl->n_insert, my_keys);
    }
};

template <int KW, int NB, bool PAGED>
__device__ __forceinline__ void tile_smem_init(TileSmem<KW, NB, PAGED>& sm) {
    uint4* c = reinterpret_cast<uint4*>(&sm.cnt[0][0]);
    for (uint32_t i = threadIdx.x; i < kRadixWarps * NB / 8; i += kRadixThreads) c[i] = make_uint4(0u, 0u, 0u, 0u);
}

// Block-local counting sort of one tile by a digit of at most log2(NB) bits and coalesced write of every bin's run.
//   Hs / vmask : the thread's OPT keys and which of them exist
//   digit      : key word 0 -> digit
//   reserve    : (bin, n) -> position of a run of n > 0 keys of `bin` in the bin's address space (one atomicAdd; the
//                owning thread issues both of its bins' reservations before it looks at either result)
//   place      : (bin, n, start, position) -> fills sm.gdelta[bin] (and gdelta2 / split when PAGED) for the run that
//                starts at tile index `start`
//   dst        : bin -> base pointer of its destination buffer
// On entry sm.cnt is all zero (and visible to the block); on exit it is zero again.
template <int KW, int NB, bool PAGED, typename DigitFn, typename ReserveFn, typename PlaceFn, typename DstFn>
__device__ __forceinline__ void tile_partition(TileSmem<KW, NB, PAGED>& sm, const Key<KW> (&Hs)[RadixCfg<KW>::OPT], uint32_t vmask,
                                               uint32_t bits, DigitFn digit, ReserveFn reserve, PlaceFn place, DstFn dst) {
    constexpr int OPT = RadixCfg<KW>::OPT;
    static_assert(NB % 2 == 0 && NB / 2 <= kRadixThreads, "two bins per thread");
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // the ranking's tag rows live in `sorted`, which is idle until the scatter below (a barrier lies in between)
    uint8_t* tag = reinterpret_cast<uint8_t*>(sm.sorted) + warp * NB;
    static_assert(kRadixWarps * NB <= RadixCfg<KW>::TILE * KW * 8, "tag rows fit the tile buffer");
    uint32_t rk[OPT];
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        const bool v = (vmask >> j) & 1u;
        const uint32_t d = v ? digit(Hs[j].w[0]) : 0u;
        rk[j] = rank_in_warp<uint16_t>(sm.cnt[warp], tag, v, d, lane, bits) | (d << 16);
    }
    __syncthreads();
    // thread t owns bins 2t and 2t+1 (one 32-bit word of every warp's counter row): totals over the warps; the
    // counters become each warp's offset inside the bin
    uint32_t c0 = 0, c1 = 0;
    if (threadIdx.x < NB / 2) {
#pragma unroll
        for (int w = 0; w < kRadixWarps; ++w) {
            uint32_t* cell = reinterpret_cast<uint32_t*>(&sm.cnt[w][0]) + threadIdx.x;
            const uint32_t c = *cell;
            *cell = (c0 & 0xffffu) | (c1 << 16);
            c0 += c & 0xffffu;
            c1 += c >> 16;
        }
    }
    uint32_t n_tile = 0;
    const uint32_t start0 = block_exscan(c0 + c1, sm.scratch, &n_tile);
    if (threadIdx.x < NB / 2) {
        const uint32_t b0 = 2 * threadIdx.x, start1 = start0 + c0;
        sm.binstart[b0] = start0;
        sm.binstart[b0 + 1] = start1;
        const unsigned long long p0 = c0 ? reserve(b0, c0) : 0ULL;
        const unsigned long long p1 = c1 ? reserve(b0 + 1, c1) : 0ULL;
        if (c0) place(b0, c0, start0, p0);
        if (c1) place(b0 + 1, c1, start1, p1);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OPT; ++j) {
        if ((vmask >> j) & 1u) {
            const uint32_t d = rk[j] >> 16;
            const uint32_t pos = sm.binstart[d] + sm.cnt[warp][d] + (rk[j] & 0xffffu);
#pragma unroll
            for (int w = 0; w < KW; ++w) sm.sorted[pos * KW + w] = Hs[j].w[w];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_tile; i += kRadixThreads) {
        uint64_t h[KW];
#pragma unroll
        for (int w = 0; w < KW; ++w) h[w] = sm.sorted[i * KW + w];
        const uint32_t d = digit(h[0]);
        long long g = sm.gdelta[d];
        if constexpr (PAGED) { if (i >= sm.split[d]) g = sm.gdelta2[d]; }
        uint64_t* out = dst(d) + (uint64_t)(g + (long long)i) * KW;
        if constexpr (KW == 1) {
            __stcg(out, h[0]);
        } else {
#pragma unroll
            for (int w = 0; w < KW; w += 2) __stcg(reinterpret_cast<ulonglong2*>(out + w), make_ulonglong2(h[w], h[w + 1]));
        }
    }
    tile_smem_init<KW, NB, PAGED>(sm);
    __syncthreads();
}

// ---- S0 (exact mode): digit-1 histogram per segment of the read stream ---------------------------------------------
// Segments [seg0, seg0 + n_segs) of the batch; seghist / segtotal are indexed relative to seg0.
// Dynamic shared memory: kRadixWarps * kNB1 counters (uint32) + kRadixWarps * kNB1 tag bytes.
constexpr size_t kHistSmemBytes = (size_t)kRadixWarps * kNB1 * 5 + 32 * 4;
template <int KW>
__global__ void __launch_bounds__(kRadixThreads, RadixCfg<KW>::MINB)
k_hist_reads(const __grid_constant__ TableView tv, const __grid_constant__ RadixGeom rg, const uint64_t* __restrict__ packed,
             const uint32_t* __restrict__ ends, uint64_t n_words, uint64_t n_bases, uint64_t seg0, uint32_t n_segs,
             uint32_t* __restrict__ seghist, uint32_t* __restrict__ segtotal) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw);                       // [kRadixWarps][kNB1]
    uint32_t* scratch = cnt + kRadixWarps * kNB1;                                // [32]
    uint8_t* tag = reinterpret_cast<uint8_t*>(scratch + 32);                     // [kRadixWarps][kNB1]
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t seg_words = 1ULL << rg.seg_log2;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        for (uint32_t i = threadIdx.x; i < kRadixWarps * kNB1; i += kRadixThreads) cnt[i] = 0u;
        __syncthreads();
        const uint64_t w0 = (seg0 + seg) << rg.seg_log2;
        const uint64_t w1 = w0 + seg_words < n_words ? w0 + seg_words : n_words;
        for (uint64_t base = w0 + warp * 32; base < w1; base += kRadixWarps * 32) {
            KmerLane<KW> kl;
            kl.load(packed, ends, base, n_words, w1, n_bases, lane, tv.L.k);
#pragma unroll 4
            for (int o = 31; o >= 16; --o) {
                Key<KW> key;
                const bool valid = kl.template kmer_at<true>(o, tv.L.k, tv.hp, key);
                const Key<KW> H = hash_key<KW>(key, tv.hp);
                (void)rank_in_warp<uint32_t>(cnt + warp * kNB1, tag + warp * kNB1, valid, digit1_of(rg, tv.lbg_mask, H.w[0]), lane, rg.d1);
            }
#pragma unroll 4
            for (int o = 15; o >= 0; --o) {
                Key<KW> key;
                const bool valid = kl.template kmer_at<false>(o, tv.L.k, tv.hp, key);
                const Key<KW> H = hash_key<KW>(key, tv.hp);
                (void)rank_in_warp<uint32_t>(cnt + warp * kNB1, tag + warp * kNB1, valid, digit1_of(rg, tv.lbg_mask, H.w[0]), lane, rg.d1);
            }
        }
        __syncthreads();
        uint32_t mine = 0;
        for (uint32_t b = threadIdx.x; b < rg.nb1; b += kRadixThreads) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kRadixWarps; ++w) total += cnt[w * kNB1 + b];
            seghist[(uint64_t)seg * rg.nb1 + b] = total;
            mine += total;
        }
        uint32_t seg_sum = 0;
        (void)block_exscan(mine, scratch, &seg_sum);
        if (threadIdx.x == 0) segtotal[seg] = seg_sum;
    }
}

// ---- paged mode: k-mers per segment from the read-end bitmap alone --------------------------------------------------
// A k-mer starts at base g iff no read ends inside [g, g+k-2] and g+k <= n_bases: the same rule KmerLane applies,
// without touching the bases.
template <int KW>
__global__ void __launch_bounds__(kRadixThreads) k_count_segs(const uint32_t* __restrict__ ends, uint64_t n_words, uint64_t n_bases,
                                                              uint32_t k, uint32_t seg_log2, uint64_t seg0, uint32_t n_segs,
                                                              uint32_t* __restrict__ segtotal) {
    constexpr int NE = KW == 1 ? 1 : (KW == 2 ? 2 : 4);
    __shared__ uint32_t scratch[32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t seg_words = 1ULL << seg_log2;
    for (uint32_t seg = blockIdx.x; seg < n_segs; seg += gridDim.x) {
        const uint64_t w0 = (seg0 + seg) << seg_log2;
        const uint64_t w1 = w0 + seg_words < n_words ? w0 + seg_words : n_words;
        uint32_t mine = 0;
        for (uint64_t base = w0 + warp * 32; base < w1; base += kRadixWarps * 32) {
            uint32_t ewin[NE + 1];
            load_window<NE, uint32_t>(ends, base, n_words, lane, ewin);
            uint32_t dist = first_end_after<NE>(ewin);
            const uint64_t g0 = (base + lane) << 5;
            const uint64_t limit = (base + lane < w1) ? n_bases : 0;
            const uint32_t ends_cur = ewin[0];
#pragma unroll 8
            for (int o = 31; o >= 0; --o) {
                dist = ((ends_cur >> o) & 1u) ? 0u : (dist == 0xffffffffu ? dist : dist + 1u);
                mine += ((dist >= k - 1) && (g0 + (uint64_t)o + k <= limit)) ? 1u : 0u;
            }
        }
        uint32_t seg_sum = 0;
        (void)block_exscan(mine, scratch, &seg_sum);
        if (threadIdx.x == 0) segtotal[seg] = seg_sum;
    }
}

// ---- planner: chunks of at most cap keys (+ exact digit-1 offsets per chunk when seghist is given) ------------------
// segprefix: n_segs + 1 words of scratch.  Chunk c = the longest run of segments after chunk c-1 whose k-mers fit cap.
// One block of 1024 threads.
__global__ void __launch_bounds__(1024) k_plan_chunks(RadixCtl* __restrict__ ctl, const uint32_t* __restrict__ seghist,
                                                      const uint32_t* __restrict__ segtotal, uint64_t* __restrict__ segprefix,
                                                      uint32_t n_segs, uint32_t nb1, uint64_t cap, uint64_t seg_keys,
                                                      unsigned long long* __restrict__ err_ctr) {
    __shared__ uint32_t scratch[32];
    __shared__ uint32_t n_chunks_s;
    // inclusive prefix of the segment totals: segprefix[s] = keys of segments [0, s)
    const uint32_t per = (n_segs + blockDim.x - 1) / blockDim.x;
    const uint32_t i0 = threadIdx.x * per < n_segs ? threadIdx.x * per : n_segs, i1 = i0 + per < n_segs ? i0 + per : n_segs;
    uint64_t mine = 0;
    for (uint32_t i = i0; i < i1; ++i) mine += segtotal[i];
    uint64_t run = block_exscan_u64(mine, scratch, nullptr);
    for (uint32_t i = i0; i < i1; ++i) { segprefix[i] = run; run += segtotal[i]; }
    if (i1 == n_segs && i0 < n_segs) segprefix[n_segs] = run;
    if (n_segs == 0 && threadIdx.x == 0) segprefix[0] = 0;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0, s0 = 0;
        bool err = false;
        // equal passes instead of full ones and a remainder: the fewest chunks that fit, then share the k-mers evenly
        const uint64_t total = segprefix[n_segs];
        const uint64_t n_min = total ? (total + cap - 1) / cap : 1;
        const uint64_t even = (total + n_min - 1) / n_min + seg_keys;
        if (even < cap) cap = even;
        while (s0 < n_segs && !err) {
            // largest s1 > s0 with segprefix[s1] - segprefix[s0] <= cap
            const uint64_t base = segprefix[s0];
            uint32_t lo = s0, hi = n_segs + 1;           // invariant: prefix[lo] - base <= cap, prefix[hi] - base > cap (virtually)
            while (hi - lo > 1) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (segprefix[mid] - base <= cap) lo = mid; else hi = mid;
            }
            if (lo == s0) { err = true; break; }         // one segment alone exceeds cap
            if (c < kMaxChunks) { ctl->chunk[c].seg_begin = s0; ctl->chunk[c].seg_end = lo; ctl->chunk[c].n_keys = segprefix[lo] - base; }
            ++c;
            s0 = lo;
        }
        if (c > kMaxChunks) err = true;
        if (err) { ctl->plan_error = 1u; c = 0; atomicOr(err_ctr, (unsigned long long)ERR_PLAN); }   // nothing is processed; reported at the next sync
        ctl->n_chunks = c;
        n_chunks_s = c;
    }
    __syncthreads();
    if (!seghist) return;
    const uint32_t nc = n_chunks_s;
    for (uint32_t c = 0; c < nc; ++c) {
        const uint64_t s0 = ctl->chunk[c].seg_begin, s1 = ctl->chunk[c].seg_end;
        uint64_t sum = 0;
        if (threadIdx.x < nb1) {
#pragma unroll 8
            for (uint64_t s = s0; s < s1; ++s) sum += seghist[s * nb1 + threadIdx.x];
        }
        uint64_t tot = 0;
        const uint64_t pre = block_exscan_u64(sum, scratch, &tot);
        if (threadIdx.x < kNB1) ctl->chunk[c].coff[threadIdx.x] = pre;
        if (threadIdx.x == 0) ctl->chunk[c].coff[kNB1] = tot;
    }
}

// Slices of phase B from the local bin counts n_b (every thread of a 1024-thread block passes the count of bin
// threadIdx.x, 0 beyond nbl): cur_coff = exclusive prefix of the counts, slicestart = exclusive prefix of the slices.
__device__ __forceinline__ void publish_bins(RadixCtl* __restrict__ ctl, uint64_t n_b, uint32_t slice_keys, uint32_t* scratch) {
    uint64_t tot = 0;
    const uint64_t pre = block_exscan_u64(n_b, scratch, &tot);
    uint32_t n_slices = 0;
    const uint32_t spre = block_exscan((uint32_t)((n_b + slice_keys - 1) / slice_keys), scratch, &n_slices);
    ctl->cur_coff[threadIdx.x] = pre;
    ctl->slicestart[threadIdx.x] = spre;
    if (threadIdx.x == 0) {
        ctl->cur_coff[kNB1] = tot;
        ctl->slicestart[kNB1] = n_slices;
        ctl->n_slices = n_slices;
        ctl->n_insert = tot;
        ctl->ticket[0] = ctl->ticket[1] = ctl->ticket[2] = 0ULL;
    }
}

// per chunk, single GPU, exact mode: arm the S1 cursors; A's bin offsets are the chunk's own
__global__ void __launch_bounds__(1024) k_chunk_begin(RadixCtl* __restrict__ ctl, uint32_t c, uint32_t slice_keys) {
    __shared__ uint32_t scratch[32];
    const bool active = c < ctl->n_chunks;
    const uint64_t off = active ? ctl->chunk[c].coff[threadIdx.x] : 0ULL;
    const uint64_t nxt = active ? ctl->chunk[c].coff[threadIdx.x + 1] : 0ULL;
    ctl->cursor1[threadIdx.x * kCursorStride] = off;
    if (threadIdx.x == 0) ctl->chunk_active = active ? 1u : 0u;
    publish_bins(ctl, nxt - off, slice_keys, scratch);
}

// per chunk, single GPU, paged mode: bins start empty, no page is taken
__global__ void __launch_bounds__(1024) k_chunk_begin_paged(RadixCtl* __restrict__ ctl, uint32_t c) {
    ctl->bin_pages[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        ctl->chunk_active = c < ctl->n_chunks ? 1u : 0u;
        ctl->page_next = 0ULL;
        ctl->n_insert = 0ULL;
        ctl->n_slices = 0u;
    }
}

// ... and after S1: one slice of phase B per page, grouped by bin (bin_pages becomes the fill cursor of k_build_slices)
__global__ void __launch_bounds__(1024) k_chunk_end_paged(RadixCtl* __restrict__ ctl) {
    __shared__ uint32_t scratch[32];
    const uint32_t n_p = ctl->chunk_active ? ctl->bin_pages[threadIdx.x] : 0u;
    uint32_t n_slices = 0;
    const uint32_t spre = block_exscan(n_p, scratch, &n_slices);
    ctl->slicestart[threadIdx.x] = spre;
    ctl->bin_pages[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
        ctl->slicestart[kNB1] = n_slices;
        ctl->n_slices = n_slices;
        if (!ctl->chunk_active) ctl->n_insert = 0ULL;
        ctl->ticket[0] = ctl->ticket[1] = ctl->ticket[2] = 0ULL;
    }
}

// per round, multi-GPU: this rank's digit-1 counts of round c (zeros once its own chunks are used up), the payload of
// the all-gather that precedes k_route_offsets
__global__ void __launch_bounds__(1024) k_round_hist(const RadixCtl* __restrict__ ctl, uint32_t c, uint32_t nb1, uint32_t* __restrict__ out) {
    if (threadIdx.x >= nb1) return;
    uint32_t v = 0;
    if (c < ctl->n_chunks) v = (uint32_t)(ctl->chunk[c].coff[threadIdx.x + 1] - ctl->chunk[c].coff[threadIdx.x]);
    out[threadIdx.x] = v;
}

// per chunk, multi-GPU: hist_all[s * nb1 + d] = k-mers rank s has for digit-1 bin d in this round (all-gathered).
// Bin d = (owner o, local bin lc).  Owner o's receive buffer is laid out bin-major, source-minor, so every local
// bin is one contiguous range there, exactly like buffer A of the single-GPU exact mode:
//   position of (source s, bin d) in o's buffer = sum_{lc' < lc} sum_s' hist[s'][o, lc'] + sum_{s' < s} hist[s'][d]
// Every rank evaluates the capacity check for every owner on the same data, so all ranks skip the round together.
__global__ void __launch_bounds__(1024) k_route_offsets(RadixCtl* __restrict__ ctl, uint32_t c, const uint32_t* __restrict__ hist_all,
                                                        uint32_t n_ranks, uint32_t my_rank, uint32_t nb1, uint32_t nbl, uint64_t cap_recv,
                                                        uint32_t slice_keys, unsigned long long* __restrict__ err_ctr) {
    __shared__ uint32_t scratch[32];
    __shared__ uint64_t ex_s[kNB1 + 1];
    __shared__ uint32_t over_s;
    if (threadIdx.x == 0) over_s = 0u;
    uint64_t col = 0, before = 0;
    if (threadIdx.x < nb1)
        for (uint32_t s = 0; s < n_ranks; ++s) {
            const uint32_t h = hist_all[(uint64_t)s * nb1 + threadIdx.x];
            if (s < my_rank) before += h;
            col += h;
        }
    uint64_t tot = 0;
    const uint64_t ex = block_exscan_u64(col, scratch, &tot);
    ex_s[threadIdx.x] = ex;
    if (threadIdx.x == 0) ex_s[kNB1] = tot;
    __syncthreads();
    if (threadIdx.x < nb1 && (threadIdx.x % nbl) == 0) {       // first bin of an owner: what that owner receives
        const uint32_t last = threadIdx.x + nbl;
        if (ex_s[last] - ex_s[threadIdx.x] > cap_recv) over_s = 1u;
    }
    __syncthreads();
    const bool over = over_s != 0u;
    const bool mine_active = c < ctl->n_chunks;
    if (threadIdx.x < nb1) {
        const uint32_t owner_first = threadIdx.x - (threadIdx.x % nbl);
        ctl->cursor1[threadIdx.x * kCursorStride] = ex - ex_s[owner_first] + before;
    }
    // my own receive layout: local bin lc = global bin my_rank * nbl + lc
    uint64_t n_b = 0;
    if (threadIdx.x < nbl && !over) n_b = ex_s[my_rank * nbl + threadIdx.x + 1] - ex_s[my_rank * nbl + threadIdx.x];
    if (threadIdx.x == 0) {
        ctl->chunk_active = (!over && mine_active) ? 1u : 0u;   // gates S1 (sending)
        if (over) { ctl->recv_overflow = 1u; atomicOr(err_ctr, (unsigned long long)ERR_SEND_OVERFLOW); }
    }
    publish_bins(ctl, n_b, slice_keys, scratch);
}

// ---- S1: extract + hash + tile sort by digit 1 -> buffer A --------------------------------------------------------
// Exact mode (PAGED = false): run of bin d goes to position cursor1[d]++ of dst_of_owner[owner of d] (multi-GPU: the
// owner's receive buffer, peer-mapped over NVLink for the other ranks: the routing kernel IS the exchange) or of A.
// Paged mode: see PageGeom.
// SPARSE: the walk over valid positions only (see SparseStage); both flavours are launched for every chunk and the one
// that does not match the chunk's density returns at once.
// Dynamic shared memory: TileSmem<KW, kNB1, PAGED> + 256 owner pointers (+ PageState) (+ SparseStage<KW>).
template <int KW, bool PAGED, bool SPARSE>
__global__ void __launch_bounds__(kRadixThreads, SPARSE ? 2 : RadixCfg<KW>::MINB)
k_part_reads(const __grid_constant__ TableView tv, const __grid_constant__ RadixGeom rg, const __grid_constant__ PageGeom pg,
             RadixCtl* __restrict__ ctl, uint32_t c, const uint64_t* __restrict__ packed, const uint32_t* __restrict__ ends,
             uint64_t n_words, uint64_t n_bases, uint64_t seg0, uint64_t* __restrict__ A, uint64_t* const* __restrict__ dst_of_owner,
             uint16_t* __restrict__ page_bin, uint16_t* __restrict__ page_len, unsigned long long* __restrict__ err_ctr,
             uint32_t sparse_pct) {
    constexpr int OPT = RadixCfg<KW>::OPT;
    using Smem = TileSmem<KW, kNB1, PAGED>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr size_t kSmemA = (sizeof(Smem) + 15) & ~(size_t)15;
    constexpr size_t kSmemSp = kSmemA + 256 * sizeof(uint64_t*) + (PAGED ? sizeof(PageState) : 0);
    static_assert(kSmemSp % 16 == 0, "SparseStage is 16-byte aligned");
    uint64_t** owner_base = reinterpret_cast<uint64_t**>(smem_raw + kSmemA);
    PagePool pool{reinterpret_cast<PageState*>(smem_raw + kSmemA + 256 * sizeof(uint64_t*)), ctl, page_bin, page_len, err_ctr, pg.page_log2, pg.n_pages, 0ULL};
    if (!ctl->chunk_active) return;
    if (chunk_is_sparse(ctl, c, rg, sparse_pct) != SPARSE) return;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    tile_smem_init<KW, kNB1, PAGED>(sm);
    if (threadIdx.x < 256) owner_base[threadIdx.x] = (dst_of_owner && threadIdx.x < (rg.nb1 >> rg.owner_shift)) ? dst_of_owner[threadIdx.x] : A;
    if constexpr (PAGED) pool.init();
    __syncthreads();
    const uint64_t seg_begin = ctl->chunk[c].seg_begin, seg_end = ctl->chunk[c].seg_end;
    const uint64_t seg_words = 1ULL << rg.seg_log2;
    const uint64_t lbg_mask = tv.lbg_mask;
    auto digit = [&](uint64_t h0) { return digit1_of(rg, lbg_mask, h0); };
    auto reserve = [&](uint32_t bin, uint32_t n) -> unsigned long long {
        if constexpr (!PAGED) return atomicAdd(&ctl->cursor1[bin * kCursorStride], (unsigned long long)n);
        else return pool.reserve(bin, n);
    };
    auto place = [&](uint32_t bin, uint32_t n, uint32_t start, unsigned long long p) {
        if constexpr (!PAGED) sm.gdelta[bin] = (long long)p - (long long)start;
        else pool.place(bin, n, start, p, sm.gdelta[bin], sm.gdelta2[bin], sm.split[bin]);
    };
    auto dst = [&](uint32_t d) { return owner_base[d >> rg.owner_shift]; };
    for (uint64_t seg = seg_begin + blockIdx.x; seg < seg_end; seg += gridDim.x) {
        const uint64_t w0 = (seg0 + seg) << rg.seg_log2;
        const uint64_t w1 = w0 + seg_words < n_words ? w0 + seg_words : n_words;
        for (uint64_t round = w0; round < w1; round += kRadixWarps * 32) {
            if constexpr (SPARSE) {
                constexpr int NE = KmerLane<KW>::NE;
                constexpr uint32_t TILE = RadixCfg<KW>::TILE;
                SparseStage<KW>& sp = *reinterpret_cast<SparseStage<KW>*>(smem_raw + kSmemSp);
                uint64_t* stream64 = reinterpret_cast<uint64_t*>(sp.stream);
                const uint32_t k = tv.L.k;
                const uint64_t wi = round + threadIdx.x;
                stream64[threadIdx.x] = wi < n_words ? __ldg(packed + wi) : 0ULL;
                if (threadIdx.x <= (unsigned)KW) {
                    const uint64_t u = round + kRadixThreads + threadIdx.x;
                    stream64[kRadixThreads + threadIdx.x] = u < n_words ? __ldg(packed + u) : 0ULL;
                }
                uint32_t ewin[NE + 1];
                load_window<NE, uint32_t>(ends, round + warp * 32, n_words, lane, ewin);
                const uint64_t g0 = wi << 5;
                int omax = -1;
                if (wi < w1 && g0 + k <= n_bases) omax = n_bases - g0 - k < 31 ? (int)(n_bases - g0 - k) : 31;
                const uint32_t vbits = valid_starts(ewin[0], first_end_after<NE>(ewin), k, omax);
                uint32_t n_valid = 0;
                const uint32_t before = block_exscan(popc32(vbits), sm.scratch, &n_valid);
                sp.vb[threadIdx.x] = vbits;
                sp.pre[threadIdx.x] = before;
                __syncthreads();
                for (uint32_t t0 = 0; t0 < n_valid; t0 += TILE) {
                    Key<KW> Hs[OPT];
                    uint32_t vmask = 0;
                    const uint32_t i0 = t0 + threadIdx.x * OPT;          // this thread's valid positions [i0, i0 + OPT)
                    SparseCursor cur{0u, 0u};
                    if (i0 < n_valid) cur = sparse_seek(sp.pre, sp.vb, kRadixThreads, i0);
#pragma unroll
                    for (int j = 0; j < OPT; ++j) {
#pragma unroll
                        for (int w = 0; w < KW; ++w) Hs[j].w[w] = 0ULL;
                        if (i0 + j < n_valid) {
                            const uint32_t o = sparse_next(cur, sp.vb);
                            Hs[j] = hash_key<KW>(kmer_from_stream32<KW>(sp.stream, cur.w, o, tv.hp), tv.hp);
                            vmask |= 1u << j;
                        }
                    }
                    tile_partition<KW, kNB1, PAGED>(sm, Hs, vmask, rg.d1, digit, reserve, place, dst);
                }
                __syncthreads();      // the stage is rewritten by the next round
            } else {
                KmerLane<KW> kl;
                kl.load(packed, ends, round + warp * 32, n_words, w1, n_bases, lane, tv.L.k);
#pragma unroll 1
                for (int o0 = 31; o0 >= 0; o0 -= OPT) {
                    Key<KW> Hs[OPT];
                    uint32_t vmask = 0;
                    if (o0 >= 16) {
#pragma unroll
                        for (int j = 0; j < OPT; ++j) {
                            Key<KW> key;
                            if (kl.template kmer_at<true>(o0 - j, tv.L.k, tv.hp, key)) vmask |= 1u << j;
                            Hs[j] = hash_key<KW>(key, tv.hp);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < OPT; ++j) {
                            Key<KW> key;
                            if (kl.template kmer_at<false>(o0 - j, tv.L.k, tv.hp, key)) vmask |= 1u << j;
                            Hs[j] = hash_key<KW>(key, tv.hp);
                        }
                    }
                    tile_partition<KW, kNB1, PAGED>(sm, Hs, vmask, rg.d1, digit, reserve, place, dst);
                }
            }
        }
    }
    if constexpr (PAGED) pool.finish();
}

// ---- per group of A: work items, histogram reset -------------------------------------------------------------------
// Group g = keys [g*capb, (g+1)*capb) of A.  Items are tiles that do not cross a coarse-bin boundary, so all keys
// of a tile share digit 1 and are told apart by digit 2 alone.
template <int KW>
__global__ void __launch_bounds__(kNB) k_plan_group(RadixCtl* __restrict__ ctl, uint32_t g, uint64_t capb, uint32_t nbl,
                                                    uint32_t* __restrict__ fhist, uint32_t n_fine) {
    constexpr uint32_t TILE = RadixCfg<KW>::TILE;
    __shared__ uint32_t scratch[8];
    GroupDesc& G = ctl->group;
    const uint64_t n_keys = ctl->cur_coff[nbl];
    const uint64_t a0 = (uint64_t)g * capb, a1 = a0 + capb < n_keys ? a0 + capb : n_keys;
    const bool active = a0 < n_keys;
    uint32_t tiles = 0;
    if (active && threadIdx.x < nbl) {
        const uint64_t lo0 = ctl->cur_coff[threadIdx.x], hi0 = ctl->cur_coff[threadIdx.x + 1];
        const uint64_t lo = lo0 > a0 ? lo0 : a0, hi = hi0 < a1 ? hi0 : a1;
        if (hi > lo) tiles = (uint32_t)((hi - lo + TILE - 1) / TILE);
    }
    uint32_t n_items = 0;
    const uint32_t pre = block_exscan_256(tiles, scratch, &n_items);
    G.itemstart[threadIdx.x] = pre;
    if (threadIdx.x == kNB - 1) G.itemstart[kNB] = pre + tiles;
    if (threadIdx.x == 0) {
        G.a0 = a0; G.a1 = active ? a1 : a0; G.n_items = active ? n_items : 0u; G.active = active ? 1u : 0u;
        ctl->n_insert = active ? a1 - a0 : 0ULL;
        ctl->ticket[0] = ctl->ticket[1] = ctl->ticket[2] = 0ULL;
    }
    if (active) for (uint32_t i = threadIdx.x; i < n_fine; i += kNB) fhist[i] = 0u;
}

// Item -> (local coarse bin, key range).  The group's tables are copied to shared memory once; thread 0 holds the
// ticket of the NEXT item (its atomicAdd round trip overlaps the current tile), searches and broadcasts.
struct ItemRange { uint32_t bin; uint32_t pad; uint64_t lo, hi; };

template <int KW>
struct ItemFeed {
    uint64_t coff[kNB + 1];
    uint32_t itemstart[kNB + 1];
    uint64_t a0, a1;
    uint32_t n_items;
    ItemRange cur;
};

template <int KW>
__device__ __forceinline__ void item_feed_init(ItemFeed<KW>& f, const RadixCtl* __restrict__ ctl, uint32_t nbl) {
    for (uint32_t i = threadIdx.x; i <= kNB; i += blockDim.x) {
        f.coff[i] = i <= nbl ? ctl->cur_coff[i] : 0ULL;
        f.itemstart[i] = ctl->group.itemstart[i];
    }
    if (threadIdx.x == 0) { f.a0 = ctl->group.a0; f.a1 = ctl->group.a1; f.n_items = ctl->group.n_items; }
    __syncthreads();
}

// `ahead` lives in thread 0's registers: the ticket already taken for the next call.
template <int KW>
__device__ __forceinline__ bool next_item(ItemFeed<KW>& f, RadixCtl* __restrict__ ctl, int ticket_idx, uint32_t nbl,
                                          unsigned long long& ahead) {
    constexpr uint32_t TILE = RadixCfg<KW>::TILE;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long item = ahead;
        if (item >= f.n_items) {
            f.cur.bin = 0xffffffffu;
        } else {
            ahead = atomicAdd(&ctl->ticket[ticket_idx], 1ULL);       // consumed by the next call
            uint32_t lo_b = 0, hi_b = nbl;     // largest b with itemstart[b] <= item (empty bins share their successor's start)
            while (hi_b - lo_b > 1) {
                const uint32_t mid = (lo_b + hi_b) >> 1;
                if (f.itemstart[mid] <= item) lo_b = mid; else hi_b = mid;
            }
            const uint64_t lo0 = f.coff[lo_b], hi0 = f.coff[lo_b + 1];
            const uint64_t lo = (lo0 > f.a0 ? lo0 : f.a0) + (uint64_t)(item - f.itemstart[lo_b]) * TILE;
            const uint64_t hi_lim = hi0 < f.a1 ? hi0 : f.a1;
            f.cur.bin = lo_b; f.cur.lo = lo; f.cur.hi = lo + TILE < hi_lim ? lo + TILE : hi_lim;
        }
    }
    __syncthreads();
    return f.cur.bin != 0xffffffffu;
}

// ---- S2a: digit-2 histogram of the group ---------------------------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(kRadixThreads, 2)
k_hist_keys(const __grid_constant__ TableView tv, const __grid_constant__ RadixGeom rg, RadixCtl* __restrict__ ctl,
            const uint64_t* __restrict__ A, uint32_t* __restrict__ fhist) {
    constexpr int OPT = RadixCfg<KW>::OPT;
    __shared__ uint16_t cnt[kRadixWarps][kNB];
    __shared__ uint8_t tag[kRadixWarps][kNB];
    __shared__ ItemFeed<KW> feed;
    if (!ctl->group.active) return;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < kRadixWarps * kNB / 2; i += kRadixThreads) reinterpret_cast<uint32_t*>(&cnt[0][0])[i] = 0u;
    item_feed_init<KW>(feed, ctl, rg.nbl);
    unsigned long long ahead = threadIdx.x == 0 ? atomicAdd(&ctl->ticket[0], 1ULL) : 0ULL;
    while (next_item<KW>(feed, ctl, 0, rg.nbl, ahead)) {
        const uint64_t lo = feed.cur.lo, hi = feed.cur.hi;
        const uint32_t bin = feed.cur.bin;
        uint64_t h0[OPT];
#pragma unroll
        for (int j = 0; j < OPT; ++j) {          // all loads first: the ranking below is a dependent chain per key
            const uint64_t i = lo + (uint64_t)j * kRadixThreads + threadIdx.x;
            h0[j] = i < hi ? __ldcg(A + i * KW) : 0ULL;
        }
#pragma unroll
        for (int j = 0; j < OPT; ++j) {
            const uint64_t i = lo + (uint64_t)j * kRadixThreads + threadIdx.x;
            (void)rank_in_warp<uint16_t>(cnt[warp], tag[warp], i < hi, digit2_of(rg, tv.lbg_mask, h0[j]), lane, rg.d2);
        }
        __syncthreads();
        if (threadIdx.x < rg.nb2) {
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < kRadixWarps; ++w) { total += cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = 0; }
            if (total) atomicAdd(fhist + bin * rg.nb2 + threadIdx.x, total);
        }
    }
}

// fine-bin offsets inside B: exclusive scan of fhist (n_fine <= 65536 entries, one block of 1024 threads)
__global__ void __launch_bounds__(1024) k_scan_fine(const RadixCtl* __restrict__ ctl, const uint32_t* __restrict__ fhist,
                                                    unsigned long long* __restrict__ fcur, uint32_t n_fine) {
    __shared__ unsigned long long wsum[32];
    if (!ctl->group.active) return;
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t per = (n_fine + 1023) / 1024;
    const uint32_t i0 = threadIdx.x * per;
    unsigned long long sum = 0;
    for (uint32_t i = i0; i < i0 + per && i < n_fine; ++i) sum += fhist[i];
    unsigned long long inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(full, inc, d);
        if (lane >= (unsigned)d) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long pre = 0;
    for (unsigned w = 0; w < warp; ++w) pre += wsum[w];
    unsigned long long run = pre + inc - sum;
    for (uint32_t i = i0; i < i0 + per && i < n_fine; ++i) { fcur[i] = run; run += fhist[i]; }
}

// ---- S2b: tile sort by digit 2: A -> B -----------------------------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(kRadixThreads, RadixCfg<KW>::MINB)
k_part_keys(const __grid_constant__ TableView tv, const __grid_constant__ RadixGeom rg, RadixCtl* __restrict__ ctl,
            const uint64_t* __restrict__ A, unsigned long long* __restrict__ fcur, uint64_t* __restrict__ B) {
    constexpr int OPT = RadixCfg<KW>::OPT;
    __shared__ TileSmem<KW, kNB, false> sm;
    __shared__ ItemFeed<KW> feed;
    if (!ctl->group.active) return;
    tile_smem_init<KW, kNB, false>(sm);
    item_feed_init<KW>(feed, ctl, rg.nbl);
    const uint64_t lbg_mask = tv.lbg_mask;
    auto digit = [&](uint64_t h0) { return digit2_of(rg, lbg_mask, h0); };
    auto dst = [&](uint32_t) { return B; };
    unsigned long long ahead = threadIdx.x == 0 ? atomicAdd(&ctl->ticket[1], 1ULL) : 0ULL;
    while (next_item<KW>(feed, ctl, 1, rg.nbl, ahead)) {
        const uint64_t lo = feed.cur.lo, hi = feed.cur.hi;
        unsigned long long* cur = fcur + (uint64_t)feed.cur.bin * rg.nb2;
        Key<KW> Hs[OPT];
        uint32_t vmask = 0;
#pragma unroll
        for (int j = 0; j < OPT; ++j) {
            const uint64_t i = lo + (uint64_t)j * kRadixThreads + threadIdx.x;
            const bool v = i < hi;
            if (v) vmask |= 1u << j;
#pragma unroll
            for (int w = 0; w < KW; ++w) Hs[j].w[w] = v ? __ldcs(A + i * KW + w) : 0ULL;
        }
        auto reserve = [&](uint32_t bin, uint32_t n) { return atomicAdd(cur + bin, (unsigned long long)n); };
        auto place = [&](uint32_t bin, uint32_t, uint32_t start, unsigned long long p) { sm.gdelta[bin] = (long long)p - (long long)start; };
        tile_partition<KW, kNB, false>(sm, Hs, vmask, rg.d2, digit, reserve, place, dst);
    }
}

// ---- two-level mode (hash-sharded tables): the received keys, sorted by a coarse local digit, get their fine digit ----
// With G shards the routing pass has 256 / G bins per shard (runs of 16 k-mers = 128 bytes cross NVLink; at 4 k-mers
// per run the peer stores ran at half the speed), which leaves regions of 0.5-1 GiB per bin.  So the receive buffer R
// is cut into groups, every group is partitioned once more by the next d2 bits into the page pool (fine bin = coarse
// bin * nb2 + digit 2, at most kNB1 of them per shard) and inserted from there.  The planner walks the coarse bins
// and ends a group where the pool could run out of pages: a (block, fine bin) pair may leave one page partly empty.
template <int KW>
__global__ void __launch_bounds__(kNB) k_plan_group_paged(RadixCtl* __restrict__ ctl, uint32_t g, uint32_t n_groups, uint32_t nbl, uint32_t nb2,
                                                          const __grid_constant__ PageGeom pg, uint32_t grid2,
                                                          unsigned long long* __restrict__ err_ctr) {
    constexpr uint32_t TILE = RadixCfg<KW>::TILE;
    __shared__ uint32_t scratch[8];
    __shared__ uint64_t a0_s, a1_s;
    GroupDesc& G = ctl->group;
    const uint64_t n_keys = ctl->cur_coff[nbl];
    if (threadIdx.x == 0) {
        const uint64_t a0 = g == 0 ? 0ULL : G.a1;
        uint64_t a1 = a0;
        if (a0 < n_keys) {
            uint32_t b = 0;
            while (b + 1 < nbl && ctl->cur_coff[b + 1] <= a0) ++b;
            const uint64_t fixed = (uint64_t)grid2 * nb2;                  // pages the fine bins of one coarse bin may waste
            uint64_t pages_left = pg.n_pages;
            for (; b < nbl; ++b) {
                const uint64_t lo = ctl->cur_coff[b] > a0 ? ctl->cur_coff[b] : a0, hi = ctl->cur_coff[b + 1];
                if (hi <= lo) continue;
                if (pages_left <= fixed) break;
                const uint64_t can = (pages_left - fixed) << pg.page_log2;
                const uint64_t take = hi - lo < can ? hi - lo : can;
                a1 = lo + take;
                pages_left -= fixed + ((take + (1ULL << pg.page_log2) - 1) >> pg.page_log2);
                if (take < hi - lo) break;
            }
            if (a1 == a0) { atomicOr(err_ctr, (unsigned long long)ERR_PLAN); a1 = n_keys; }   // pool smaller than one bin's slack: refused at creation
            if (g + 1 == n_groups && a1 < n_keys) atomicOr(err_ctr, (unsigned long long)ERR_PLAN);   // out of group slots: cannot happen (see the host side)
        }
        a0_s = a0; a1_s = a1;
    }
    __syncthreads();
    const uint64_t a0 = a0_s, a1 = a1_s;
    const bool active = a1 > a0;
    uint32_t tiles = 0;
    if (active && threadIdx.x < nbl) {
        const uint64_t lo0 = ctl->cur_coff[threadIdx.x], hi0 = ctl->cur_coff[threadIdx.x + 1];
        const uint64_t lo = lo0 > a0 ? lo0 : a0, hi = hi0 < a1 ? hi0 : a1;
        if (hi > lo) tiles = (uint32_t)((hi - lo + TILE - 1) / TILE);
    }
    uint32_t n_items = 0;
    const uint32_t pre = block_exscan_256(tiles, scratch, &n_items);
    G.itemstart[threadIdx.x] = pre;
    if (threadIdx.x == kNB - 1) G.itemstart[kNB] = pre + tiles;
    for (uint32_t i = threadIdx.x; i < kNB1; i += kNB) ctl->bin_pages[i] = 0u;
    if (threadIdx.x == 0) {
        G.a0 = a0; G.a1 = a1; G.n_items = active ? n_items : 0u; G.active = active ? 1u : 0u;
        ctl->chunk_active = active ? 1u : 0u;          // k_chunk_end_paged looks at this
        ctl->page_next = 0ULL;
        ctl->n_insert = 0ULL;
        ctl->n_slices = 0u;
        ctl->ticket[0] = ctl->ticket[1] = ctl->ticket[2] = 0ULL;
    }
}

// Dynamic shared memory: TileSmem<KW, kNB, true> + ItemFeed<KW> + PageState.
template <int KW>
__global__ void __launch_bounds__(kRadixThreads, RadixCfg<KW>::MINB)
k_part_keys_paged(const __grid_constant__ TableView tv, const __grid_constant__ RadixGeom rg, const __grid_constant__ PageGeom pg,
                  RadixCtl* __restrict__ ctl, const uint64_t* __restrict__ R, uint64_t* __restrict__ pool_keys,
                  uint16_t* __restrict__ page_bin, uint16_t* __restrict__ page_len, unsigned long long* __restrict__ err_ctr) {
    constexpr int OPT = RadixCfg<KW>::OPT;
    using Smem = TileSmem<KW, kNB, true>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr size_t kSmemA = (sizeof(Smem) + 15) & ~(size_t)15;
    constexpr size_t kSmemB = (kSmemA + sizeof(ItemFeed<KW>) + 15) & ~(size_t)15;
    ItemFeed<KW>& feed = *reinterpret_cast<ItemFeed<KW>*>(smem_raw + kSmemA);
    PagePool pool{reinterpret_cast<PageState*>(smem_raw + kSmemB), ctl, page_bin, page_len, err_ctr, pg.page_log2, pg.n_pages, 0ULL};
    if (!ctl->group.active) return;
    tile_smem_init<KW, kNB, true>(sm);
    pool.init();
    item_feed_init<KW>(feed, ctl, rg.nbl);
    const uint64_t lbg_mask = tv.lbg_mask;
    auto digit = [&](uint64_t h0) { return digit2_of(rg, lbg_mask, h0); };
    auto dst = [&](uint32_t) { return pool_keys; };
    unsigned long long ahead = threadIdx.x == 0 ? atomicAdd(&ctl->ticket[1], 1ULL) : 0ULL;
    while (next_item<KW>(feed, ctl, 1, rg.nbl, ahead)) {
        const uint64_t lo = feed.cur.lo, hi = feed.cur.hi;
        const uint32_t fine0 = feed.cur.bin * rg.nb2;            // all keys of an item share the coarse bin
        Key<KW> Hs[OPT];
        uint32_t vmask = 0;
#pragma unroll
        for (int j = 0; j < OPT; ++j) {
            const uint64_t i = lo + (uint64_t)j * kRadixThreads + threadIdx.x;
            const bool v = i < hi;
            if (v) vmask |= 1u << j;
#pragma unroll
            for (int w = 0; w < KW; ++w) Hs[j].w[w] = v ? __ldcs(R + i * KW + w) : 0ULL;
        }
        auto reserve = [&](uint32_t bin, uint32_t n) { return pool.reserve(fine0 + bin, n); };
        auto place = [&](uint32_t bin, uint32_t n, uint32_t start, unsigned long long p) {
            pool.place(fine0 + bin, n, start, p, sm.gdelta[bin], sm.gdelta2[bin], sm.split[bin]);
        };
        tile_partition<KW, kNB, true>(sm, Hs, vmask, rg.d2, digit, reserve, place, dst);
    }
    pool.finish();
}

// ---- phase B: insert keys in region order ---------------------------------------------------------------------------
// Work item = one slice of consecutive keys of `src` (ticket order = table-region order).
//
// What was measured about this kernel (profiles/r02_k0r_modes.md, profiles/r02_pipeline_history.md,
// profiles/r02_single_pass_c2_quarter_summary.txt):
//  * all blocks sweeping one table region at a time, sector loads alone run at 165-177 G/s, a load followed by a
//    fire-and-forget RED at 70 G/s, a load followed by ANY 64-bit atomic whose result returns to the SM at 45 G/s (37 G/s
//    with 128 MiB regions), whether the thread waits for the result or not and at any occupancy from 1280 to 2048 threads
//    per SM: the insert's claim is bound by the atomic-return path, and this kernel runs at 0.92-0.94 of that rate;
//  * 40 registers (6 resident blocks per SM) beat 48 / 64 / 32 (221 / 239 / 221 vs 202 ms on config 2), provided the
//    per-k-mer statistics stay out of local memory (see insert_hashed<LEAN>);
//  * prefetching the next slice's home buckets with prefetch.global.L2 costs 10 % (221 -> 242 ms);
//  * merging equal keys of a warp with __match_any_sync on every step kept the XU pipe 93 % busy (MATCH.ANY.U64);
//  * issuing several probes and claims per thread before looking at any result made it slower (439 vs 257 ms): at load
//    factors near 0.5 almost every warp has a lane whose home bucket is full, and that lane's re-probe chain then runs
//    with the other 31 lanes idle.
// So: one key per thread and step, and the duplicate handling only where it pays.  A slice in which two neighbouring
// keys are equal (with a k-mer that makes up a few per cent of a region this is all but certain) is "skewed": equal keys
// of a warp are merged with __match_any_sync and every key of the slice goes through a shared-memory combiner: slot word
// = index of the first key that claimed it (+1) in the low 16 bits | 48 fingerprint bits of hash word 0; a later key with
// the same fingerprint compares itself with the claimer's key in `src` and, if equal, just adds to the slot's count.  The
// slots are flushed once per slice, so a k-mer that dominates a region costs one table update per slice instead of one
// per occurrence.  All other slices take every key straight to insert_hashed().
#ifndef TSX_INSERT_MINB
#define TSX_INSERT_MINB 6
#endif

struct CombSmem {
    unsigned long long key[kCombSlots];
    unsigned int cnt[kCombSlots];
    unsigned int used;
};

// A slice in which neighbouring keys are equal.  Out of line and with its own statistics, so that the common path of
// k_insert_keys keeps its registers (inlined, this code cost the kernel ~110 bytes of spills per thread, and spill
// stores go through to L2: 1.3 sectors per k-mer next to the 2 the insert itself needs).
template <int KW, int W>
__device__ __noinline__ void insert_slice_skewed(const TableView& tv, const uint64_t* __restrict__ src, uint64_t lo, uint64_t hi,
                                                 CombSmem& comb) {
    constexpr int R = KW == 1 ? 4 : (KW == 2 ? 2 : 1);
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    LocalStats st;
#pragma unroll 1
    for (int r = 0; r < R; ++r) {
        const uint64_t i = lo + (uint64_t)r * kBlockThreads + threadIdx.x;
        const bool valid = i < hi;
        Key<KW> H;
#pragma unroll
        for (int j = 0; j < KW; ++j) H.w[j] = valid ? __ldcg(src + i * KW + j) : 0ULL;
        uint64_t cnt = 1;
        bool lead = valid;
        const unsigned vm = __ballot_sync(full, valid);
        if (valid) {
            unsigned peers = __match_any_sync(vm, H.w[0]);
#pragma unroll
            for (int j = 1; j < KW; ++j) peers &= __match_any_sync(vm, H.w[j]);
            cnt = (uint64_t)__popc(peers);
            lead = (unsigned)(__ffs(peers) - 1) == lane;
        }
        if (!lead) continue;
        {
            // combine across the block before touching the table: every key of a skewed slice goes through the
            // shared-memory combiner, so copies that are far apart in the slice are merged too (a k-mer that makes up
            // 1 % of a table region reaches the table once per slice instead of ten times; what matters is not the
            // atomics saved but that all blocks work in the same region, i.e. on the same hot entries, at once)
            const uint32_t slot = (uint32_t)(H.w[0] ^ (H.w[0] >> 40)) & (kCombSlots - 1);
            const unsigned long long tagged = (H.w[0] & ~0xffffULL) | (unsigned long long)((i - lo) + 1);
            const unsigned long long oldk = atomicCAS(&comb.key[slot], 0ULL, tagged);
            bool absorbed = (oldk == 0ULL);
            if (!absorbed && ((oldk ^ tagged) & ~0xffffULL) == 0ULL) {
                const uint64_t other = lo + (oldk & 0xffffULL) - 1;
                absorbed = true;
#pragma unroll
                for (int j = 0; j < KW; ++j) absorbed &= (__ldcg(src + other * KW + j) == H.w[j]);
            }
            if (absorbed) {
                atomicAdd(&comb.cnt[slot], (unsigned int)cnt);
                comb.used = 1u;
                continue;
            }
        }
        insert_hashed<KW, W, true>(tv, H, cnt, st);
    }
    __syncthreads();
    if (comb.used) {
        for (uint32_t s = threadIdx.x; s < kCombSlots; s += kBlockThreads) {
            const unsigned long long tagged = comb.key[s];
            if (tagged == 0ULL) continue;
            const uint64_t idx = lo + (tagged & 0xffffULL) - 1;
            Key<KW> H;
#pragma unroll
            for (int j = 0; j < KW; ++j) H.w[j] = __ldcg(src + idx * KW + j);
            insert_hashed<KW, W, true>(tv, H, (uint64_t)comb.cnt[s], st);
            comb.key[s] = 0ULL; comb.cnt[s] = 0u;
        }
        __syncthreads();
        if (threadIdx.x == 0) comb.used = 0u;
    }
    flush_stats(tv, st);
}

// Work items of phase B.  desc[item] = (position of the slice's first key in A, number of keys).
// Exact mode: slice j of local bin b = keys [j * slice_keys, ...) of the bin, which starts at cur_coff[b].
// Paged mode: one slice per page; the pages of a bin take consecutive items in any order (bin_pages is the cursor).
__global__ void __launch_bounds__(kBlockThreads) k_build_slices(RadixCtl* __restrict__ ctl, const __grid_constant__ PageGeom pg,
                                                                const uint16_t* __restrict__ page_bin, const uint16_t* __restrict__ page_len,
                                                                uint32_t nbl, uint32_t slice_keys, ulonglong2* __restrict__ desc) {
    if (pg.paged) {
        const uint64_t n_used = ctl->page_next < pg.n_pages ? ctl->page_next : pg.n_pages;
        for (uint64_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_used; p += (uint64_t)gridDim.x * blockDim.x) {
            const uint32_t b = page_bin[p];
            if (b >= nbl) continue;
            const uint32_t pos = ctl->slicestart[b] + atomicAdd(&ctl->bin_pages[b], 1u);
            desc[pos] = make_ulonglong2(p << pg.page_log2, (unsigned long long)page_len[p]);
        }
        return;
    }
    const uint32_t n_slices = ctl->n_slices;
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n_slices; item += gridDim.x * blockDim.x) {
        uint32_t lo_b = 0, hi_b = nbl;         // largest b with slicestart[b] <= item (empty bins share their successor's start)
        while (hi_b - lo_b > 1) {
            const uint32_t mid = (lo_b + hi_b) >> 1;
            if (ctl->slicestart[mid] <= item) lo_b = mid; else hi_b = mid;
        }
        const uint64_t v = (uint64_t)(item - ctl->slicestart[lo_b]) * slice_keys;
        const uint64_t n_b = ctl->cur_coff[lo_b + 1] - ctl->cur_coff[lo_b];
        const uint64_t len = n_b - v < slice_keys ? n_b - v : slice_keys;
        desc[item] = make_ulonglong2(ctl->cur_coff[lo_b] + v, len);
    }
}

template <int KW> struct InsertCfg {
    static constexpr int R = KW == 1 ? 4 : (KW == 2 ? 2 : 1);          // keys per thread per slice
    static constexpr uint32_t SLICE = kBlockThreads * R;
};

template <int KW, int W, bool WARP_AGG>
__global__ void __launch_bounds__(kBlockThreads, TSX_INSERT_MINB)
k_insert_keys(const __grid_constant__ TableView tv, RadixCtl* __restrict__ ctl, const uint64_t* __restrict__ src,
              const ulonglong2* __restrict__ desc) {
    constexpr int R = InsertCfg<KW>::R;
    const unsigned full = 0xffffffffu;
    __shared__ unsigned long long item_lo_s;
    __shared__ uint32_t item_len_s;
    __shared__ CombSmem comb;
    const unsigned long long n_items = ctl->n_slices;
    if (n_items == 0) return;
    LocalStats st;
    const unsigned lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < kCombSlots; i += kBlockThreads) { comb.key[i] = 0ULL; comb.cnt[i] = 0u; }
    // thread 0 keeps the ticket of the next slice in flight while the block works on the current one
    unsigned long long ahead = 0ULL;
    if (threadIdx.x == 0) {
        comb.used = 0u;
        ahead = atomicAdd(&ctl->ticket[2], 1ULL);
        if (blockIdx.x == 0) atomicAdd(tv.ctr + CTR_ADDED, ctl->n_insert);   // all keys of the launch; insert_hashed<LEAN> takes back what it skips
    }
    while (true) {
        __syncthreads();                                  // everyone has read the previous slice's descriptor
        if (threadIdx.x == 0) {
            // once the reprobe limit was reached anywhere the run is lost (the reference exits with 42): stop early
            const bool full_table = (__ldcg(tv.ctr + CTR_ERRORS) & (unsigned long long)ERR_TABLE_FULL) != 0ULL;
            if (ahead < n_items && !full_table) {
                const ulonglong2 d = __ldg(desc + ahead);
                item_lo_s = d.x; item_len_s = (uint32_t)d.y;
                ahead = atomicAdd(&ctl->ticket[2], 1ULL);
            } else {
                item_len_s = 0u;
            }
        }
        __syncthreads();
        const uint64_t lo = item_lo_s;
        const uint64_t hi = lo + item_len_s;
        if (hi == lo) break;
        // Pass 1 reads word 0 of the thread's keys only to see whether neighbours are equal; pass 2 reads the keys
        // again (L1 / L2 hits) one at a time with the next one in flight.  Holding all R keys across the inserts
        // instead costs 6-14 registers, which at 6 resident blocks per SM means spills.
        bool dup = false;
        if (WARP_AGG) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint64_t i = lo + (uint64_t)r * kBlockThreads + threadIdx.x;
                const uint64_t k0 = i < hi ? __ldg(src + i * KW) : 0ULL;
                const uint64_t nb = __shfl_down_sync(full, k0, 1);
                dup |= (lane < 31u) && (i + 1 < hi) && (nb == k0);
            }
        }
        if (WARP_AGG && __syncthreads_or(dup ? 1 : 0)) {
            insert_slice_skewed<KW, W>(tv, src, lo, hi, comb);
            continue;
        }
        uint64_t i = lo + threadIdx.x;
        Key<KW> H;
#pragma unroll
        for (int j = 0; j < KW; ++j) H.w[j] = i < hi ? __ldcs(src + i * KW + j) : 0ULL;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const uint64_t i_next = i + kBlockThreads;
            Key<KW> Hn;
#pragma unroll
            for (int j = 0; j < KW; ++j) Hn.w[j] = (r + 1 < R && i_next < hi) ? __ldcs(src + i_next * KW + j) : 0ULL;
            if (i < hi) insert_hashed<KW, W, true>(tv, H, 1, st);
            H = Hn;
            i = i_next;
        }
    }
    flush_stats(tv, st);
}

// ---- region-sorted batched lookup -------------------------------------------------------------------------------------
// getKmerCount(kmer) for millions of k-mers against a table far larger than TLB reach: unsorted, the probes run at the
// uniformly-random rate (5 G/s on 128 GiB); sorted by table region with the same two partition passes as the insert
// they run at the region-sweep rate.  Records are (hash word 0, query index) pairs, i.e. keys of two words.
template <int KW>
__global__ void __launch_bounds__(kBlockThreads) k_hash_queries(const __grid_constant__ TableView tv, const uint64_t* __restrict__ kmers,
                                                                uint64_t n, uint64_t* __restrict__ pairs) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Key<KW> key;
        bool in_range = true;
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            key.w[j] = __ldg(kmers + i * KW + j);
            in_range &= (key.w[j] & ~word_mask<KW>(j, tv.hp)) == 0;
        }
        const Key<KW> H = hash_key<KW>(key, tv.hp);
        pairs[2 * i] = H.w[0];
        pairs[2 * i + 1] = i | (in_range ? 0ULL : 1ULL << 63);      // bits beyond 2k set: not a k-mer, count 0
    }
}

template <int KW, int W>
__global__ void __launch_bounds__(kBlockThreads) k_lookup_pairs(const __grid_constant__ TableView tv, const uint64_t* __restrict__ pairs,
                                                                uint64_t n, const uint64_t* __restrict__ kmers,
                                                                uint64_t* __restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const uint64_t tag = __ldcs(pairs + 2 * j + 1);
        const uint64_t idx = tag & ~(1ULL << 63);
        uint64_t c = 0;
        if (!(tag >> 63)) {
            Key<KW> key;
#pragma unroll
            for (int w = 0; w < KW; ++w) key.w[w] = __ldg(kmers + idx * KW + w);
            c = lookup_hashed<KW, W>(tv, hash_key<KW>(key, tv.hp));
        }
        counts[idx] = c;
    }
}

// one coarse "bin" holding all n records (first partition pass of an unsorted array)
__global__ void k_single_bin(RadixCtl* __restrict__ ctl, uint64_t n) {
    if (threadIdx.x > kNB) return;
    ctl->cur_coff[threadIdx.x] = threadIdx.x == 0 ? 0ULL : n;
}

// after a scatter the cursors of the fine bins stand at the bins' ends: they are the coarse offsets of the next pass
__global__ void k_coff_from_cursors(RadixCtl* __restrict__ ctl, const unsigned long long* __restrict__ fcur, uint32_t nb, uint64_t n) {
    const uint32_t b = threadIdx.x;
    if (b > kNB) return;
    ctl->cur_coff[b] = b == 0 ? 0ULL : (b <= nb ? (uint64_t)fcur[b - 1] : n);
}

#endif  // __CUDACC__

}  // namespace tsx
