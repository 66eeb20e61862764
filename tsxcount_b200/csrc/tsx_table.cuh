// tsx_table.cuh — the HBM-resident counting table: entry layout, probe sequence, lock-free
// insert / increment / overflow / lookup device functions.
//
// Replaces (reference paths relative to mjoppich/tsxCount):
//   src/tsxcount/TSXHashMap.h:79-128,1135-1166    bit-packed entries  value[s] | reprobe[L] | func[2k-L]
//   src/tsxcount/TSXHashMap.h:759-778,1046-1054   pos = (key + i(i+1)/2) mod 2^L, i >= 1
//   src/tsxcount/TSXHashMapPerf.h:56-205          addKmer: probe / insert / increment
//   src/tsxcount/TSXHashMapCAS.h:141-232,268-508  byte-wise CAS store with rollback (the lock-free mode)
//   src/tsxcount/TSXHashMapPerf.h:218-289,426-462,699-881  value overflow -> tagged overflow entry
//   src/tsxcount/TSXHashMap.h:548-638,951-1039    getKmerCount + findOverflowCounts
//   src/tsxcount/TSXHashMap.h:107-108,645-648     k-mer-start bitmap (folded into the entry: OVF flag)
//
// B200 layout.  The unit of HBM traffic is the 32-byte sector, so the probe unit is a 32-byte BUCKET:
// 4 / 2 / 1 entries of 1 / 2 / 4 64-bit words.  Probe i (1-based, as in the reference) of a k-mer with
// hash H inspects bucket (H mod 2^LB + i(i+1)/2) mod 2^LB; the entry stores the quotient H >> LB, so the
// k-mer is recoverable from (bucket, i, quotient) exactly like TSXHashMap::getAllKmers does.
//
// Entry = W words, the LAST word is the head:
//   head  [0,R)          reprobe index i (>= 1, so an occupied entry is never all-zero)
//         R              OVF   1 = overflow entry, 0 = primary  (the reference's m_iKmerStarts bit, inverted)
//         R+1            HASOVF  primary owns an overflow entry (lets lookup/dump skip the chain walk)
//         R+2            BUSY  (W=4 only) key words not yet published
//         [R+3,R+3+Qh)   low Qh bits of the quotient
//         [64-V,64)      value, at the TOP of the word: atomicAdd(c << (64-V)) wraps mod 2^V without
//                        touching any other field, and the returned old value tells exactly one thread
//                        that it produced the carry.
//   body  words 0..W-2   remaining Q-Qh quotient bits, little endian
// Overflow entry (head only, body zero): [0,R) i of the primary | OVF=1 | [R+3,R+5) slot of the primary
//   inside its bucket | [R+5,2R+5) j = extra probes beyond i | [2R+5,64) counter.  It lives in the first
//   free slot on the primary's own probe sequence after probe i (reference: TSXHashMapPerf.h:426-462).
//   count = (overflow counter << V) | value   (reference: TSXHashMap.h:600-609).
#pragma once

#include <cstdint>

#include "tsx_hash.cuh"

namespace tsx {

enum : uint32_t {
    ERR_TABLE_FULL = 1u,
    ERR_SATURATED = 2u,
    ERR_SEND_OVERFLOW = 4u,
    ERR_WRONG_SHARD = 8u,
    ERR_PLAN = 16u,
};

enum : int { CTR_DISTINCT = 0, CTR_OVERFLOW = 1, CTR_ADDED = 2, CTR_MAXPROBE = 3, CTR_ERRORS = 4, CTR_COUNT = 8 };

struct Layout {
    uint32_t k, l, s_req, flags;
    uint32_t KW, W, SPB;           // key words, entry words, slots per bucket
    uint32_t LBg, LBl;             // bucket-index bits: global, local to the shard
    uint32_t shard_bits, shard_rank;
    uint32_t Q, Qh;                // quotient bits total / held in the head
    uint32_t V, R;                 // value bits, reprobe bits
    uint32_t max_probe;            // 2^R - 1
    uint32_t cshift;               // overflow counter position (2R+5)
    uint32_t pad;
    uint64_t n_buckets;            // local
    uint64_t n_slots;              // local
    uint64_t table_bytes;
};

struct TableView {
    uint64_t* words;               // n_buckets * 4 words
    unsigned long long* ctr;       // CTR_COUNT counters
    HashParams hp;
    Layout L;
    uint64_t lbg_mask, lbl_mask;
    uint64_t rmask;                // reprobe field
    uint64_t f_ovf, f_hasovf, f_busy;
    uint64_t qh_mask;              // Qh low bits
    uint64_t cmp_mask;             // head bits compared for a primary match (reprobe|OVF|quotient part)
    uint64_t tag_mask;             // head bits compared for an overflow match
    uint32_t vshift;               // 64 - V
    uint32_t qshift;               // R + 3
};

// Chooses the entry class for (k, l, s).  Returns false when nothing fits.
inline bool make_layout(uint32_t k, uint32_t l, uint32_t s, uint32_t flags, uint32_t shard_rank, uint32_t n_shards,
                        Layout* out) {
    if (k < 1 || k > 128 || l < 2 || l > 40) return false;
    if (2 * k <= l) return false;  // TSXHashMap.h:91-94
    uint32_t shard_bits = 0;
    while ((1u << shard_bits) < n_shards) ++shard_bits;
    if ((1u << shard_bits) != n_shards || shard_rank >= n_shards) return false;
    const bool exact = (flags & 1u) && s > 0;
    const uint32_t KW = k <= 32 ? 1 : (k <= 64 ? 2 : 4);
    static const uint32_t Ws[3] = {1, 2, 4};
    static const uint32_t Rs[3] = {6, 7, 8};
    for (int c = 0; c < 3; ++c) {
        const uint32_t W = Ws[c], R = Rs[c];
        // instantiated (KW, W) pairs: (1,1) (1,2) (2,2) (2,4) (4,4)
        if (W < KW || W > 2 * KW) continue;
        const uint32_t spb_log = W == 1 ? 2 : (W == 2 ? 1 : 0);
        if (l < spb_log + shard_bits + 1) continue;
        const uint32_t LBg = l - spb_log;
        if (2 * k <= LBg) continue;
        const uint32_t Q = 2 * k - LBg;
        const uint32_t body_cap = 64 * (W - 1);
        const uint32_t qh_min = Q > body_cap ? Q - body_cap : 0;
        const uint32_t fixed = 3 + R;
        const uint32_t v_need = s > 0 ? s : 4;
        if (qh_min + fixed + v_need > 64) continue;
        uint32_t V;
        if (exact) {
            V = s;
        } else {
            V = 64 - fixed - qh_min;
            if (W > 1 && V > 40) V = 40;
        }
        uint32_t Qh = 64 - fixed - V;
        if (Qh > Q) Qh = Q;
        Layout L{};
        L.k = k; L.l = l; L.s_req = s; L.flags = flags;
        L.KW = KW; L.W = W; L.SPB = 4 / W;
        L.LBg = LBg; L.shard_bits = shard_bits; L.shard_rank = shard_rank; L.LBl = LBg - shard_bits;
        L.Q = Q; L.Qh = Qh; L.V = V; L.R = R;
        L.max_probe = (1u << R) - 1;
        L.cshift = 2 * R + 5;
        L.n_buckets = 1ULL << L.LBl;
        L.n_slots = L.n_buckets * L.SPB;
        L.table_bytes = L.n_buckets * 32ULL;
        *out = L;
        return true;
    }
    return false;
}

inline TableView make_view(const Layout& L, uint64_t* words, unsigned long long* ctr) {
    TableView tv{};
    tv.words = words; tv.ctr = ctr;
    tv.hp = make_hash_params(L.k, (L.flags & 8u) != 0);     // TSXC_FLAG_CANONICAL
    tv.L = L;
    tv.lbg_mask = low_mask(L.LBg);
    tv.lbl_mask = low_mask(L.LBl);
    tv.rmask = low_mask(L.R);
    tv.f_ovf = 1ULL << L.R;
    tv.f_hasovf = 1ULL << (L.R + 1);
    tv.f_busy = 1ULL << (L.R + 2);
    tv.qh_mask = low_mask(L.Qh);
    tv.qshift = L.R + 3;
    tv.vshift = 64 - L.V;
    tv.cmp_mask = tv.rmask | tv.f_ovf | (tv.qh_mask << tv.qshift);
    tv.tag_mask = low_mask(L.cshift) & ~(tv.f_hasovf | tv.f_busy);
    return tv;
}

#if defined(__CUDACC__)

// ---- memory primitives --------------------------------------------------------------------------
// Table words are written by atomics from every SM: reads must come from L2 (ld.global.cg), never a
// stale L1 line.
// One 256-bit load (sm_100: LDG.E.256) instead of two 128-bit ones: a scattered warp-wide load costs the LSU one
// wavefront per lane PER INSTRUCTION.
__device__ __forceinline__ void load_bucket(const uint64_t* b, uint64_t (&w)[4]) {
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(b) : "memory");
}

// 128-bit compare-and-swap (sm_90+): returns the old value in (o0, o1)
__device__ __forceinline__ void cas128(uint64_t* addr, uint64_t c0, uint64_t c1, uint64_t s0, uint64_t s1, uint64_t& o0,
                                       uint64_t& o1) {
    asm volatile(
        "{\n\t"
        ".reg .b128 cmp, swp, old;\n\t"
        "mov.b128 cmp, {%3, %4};\n\t"
        "mov.b128 swp, {%5, %6};\n\t"
        "atom.global.relaxed.gpu.cas.b128 old, [%2], cmp, swp;\n\t"
        "mov.b128 {%0, %1}, old;\n\t"
        "}"
        : "=l"(o0), "=l"(o1)
        : "l"(addr), "l"(c0), "l"(c1), "l"(s0), "l"(s1)
        : "memory");
}

__device__ __forceinline__ uint64_t tri(uint32_t i) { return ((uint64_t)i * (i + 1)) >> 1; }

// Per-thread statistics, reduced per warp at kernel exit.
struct LocalStats {
    uint32_t distinct = 0, overflow = 0, maxprobe = 0, errors = 0;
    uint64_t added = 0;
};

__device__ __forceinline__ void flush_stats(const TableView& tv, const LocalStats& st) {
    const unsigned full = 0xffffffffu;
    uint32_t d = __reduce_add_sync(full, st.distinct);
    uint32_t o = __reduce_add_sync(full, st.overflow);
    uint32_t m = __reduce_max_sync(full, st.maxprobe > 1u ? st.maxprobe : (st.distinct ? 1u : 0u));
    uint32_t e = __reduce_or_sync(full, st.errors);
    uint64_t a = st.added;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(full, a, s);
    if ((threadIdx.x & 31) == 0) {
        if (d) atomicAdd(tv.ctr + CTR_DISTINCT, (unsigned long long)d);
        if (o) atomicAdd(tv.ctr + CTR_OVERFLOW, (unsigned long long)o);
        if (a) atomicAdd(tv.ctr + CTR_ADDED, (unsigned long long)a);
        if (m) atomicMax(tv.ctr + CTR_MAXPROBE, (unsigned long long)m);
        if (e) atomicOr(tv.ctr + CTR_ERRORS, (unsigned long long)e);
    }
}

// The stored form of a hashed k-mer.
template <int KW, int W>
struct Stored {
    uint64_t bucket;      // local home bucket
    uint64_t headkey;     // quotient part of the head, already shifted; reprobe/flags zero
    uint64_t body[W > 1 ? W - 1 : 1];
    bool owned;           // false: the hash belongs to another shard
};

template <int KW, int W>
__device__ __forceinline__ Stored<KW, W> make_stored(const TableView& tv, const Key<KW>& H) {
    Stored<KW, W> s;
    const uint64_t bg = H.w[0] & tv.lbg_mask;
    s.owned = (bg >> tv.L.LBl) == tv.L.shard_rank;
    s.bucket = bg & tv.lbl_mask;
    const Key<KW> q = shr_small<KW>(H, tv.L.LBg);
    s.headkey = (q.w[0] & tv.qh_mask) << tv.qshift;
    const Key<KW> b = shr_small<KW>(q, tv.L.Qh);
#pragma unroll
    for (int j = 0; j < (W > 1 ? W - 1 : 1); ++j) s.body[j] = (j < KW) ? b.w[j] : 0ULL;
    if (W == 1) s.body[0] = 0;
    return s;
}

// Add `amount` to the overflow entry of the primary found at probe i / slot pslot.
// Reference: handleOverflow, TSXHashMapPerf.h:699-881 (probe on at reprobes i+j, tag = (i, j)).
// Out of line (rare) and WITHOUT a LocalStats reference: a reference handed to a non-inlined function forces the
// caller's statistics into local memory for the whole kernel (ncu showed 2 LDL + 3 STL per inserted k-mer, and spill
// stores go through to L2: 1.3 sectors per k-mer).  Returns error flags | 0x100 if a new overflow entry was created.
constexpr uint32_t kOvfCreated = 0x100u;
template <int KW, int W>
__device__ __noinline__ uint32_t overflow_add_raw(const TableView& tv, uint64_t home, uint32_t i, uint32_t pslot, uint64_t amount) {
    constexpr int SPB = 4 / W;
    const uint64_t tag0 = (uint64_t)i | tv.f_ovf | ((uint64_t)pslot << (tv.L.R + 3));
    const uint32_t cbits = 64 - tv.L.cshift;
    if (amount >> cbits) return ERR_SATURATED;
    for (uint32_t j = 1; j <= tv.L.max_probe; ++j) {
        const uint64_t tag = tag0 | ((uint64_t)j << (tv.L.R + 5));
        uint64_t* bp = tv.words + (((home + tri(i + j)) & tv.lbl_mask) << 2);
        uint64_t w[4];
        load_bucket(bp, w);
#pragma unroll
        for (int sl = 0; sl < SPB; ++sl) {
            uint64_t* hp = bp + sl * W + (W - 1);
            uint64_t h = w[sl * W + (W - 1)];
            const bool empty = (h == 0);  // i >= 1 keeps every occupied head non-zero
            if (empty) {
                const uint64_t nh = tag | (amount << tv.L.cshift);
                if (W == 2) {
                    uint64_t o0, o1;
                    cas128(bp + sl * W, 0, 0, 0, nh, o0, o1);
                    if (o0 == 0 && o1 == 0) return kOvfCreated;
                    h = o1;
                } else {
                    const uint64_t old = atomicCAS((unsigned long long*)hp, 0ULL, (unsigned long long)nh);
                    if (old == 0) return kOvfCreated;
                    h = old;
                }
            }
            if ((h & tv.tag_mask) == tag) {
                const uint64_t old = atomicAdd((unsigned long long*)hp, (unsigned long long)(amount << tv.L.cshift));
                return (((old >> tv.L.cshift) + amount) >> cbits) ? (uint32_t)ERR_SATURATED : 0u;
            }
        }
    }
    return ERR_TABLE_FULL;
}

template <int KW, int W>
__device__ __forceinline__ void overflow_add(const TableView& tv, uint64_t home, uint32_t i, uint32_t pslot, uint64_t amount,
                                             LocalStats& st) {
    const uint32_t r = overflow_add_raw<KW, W>(tv, home, i, pslot, amount);
    st.errors |= r & 0xffu;
    st.overflow += r >> 8;
}

// value += count on a matched primary; propagates the carry.  Reference: incrementElement_key_value,
// TSXHashMapPerf.h:218-289 (value all-ones -> 0 and report overflow).
template <int KW, int W>
__device__ __forceinline__ void add_to_primary(const TableView& tv, uint64_t* hp, uint64_t home, uint32_t i, uint32_t pslot,
                                               uint64_t count, LocalStats& st) {
    const uint64_t vmask = low_mask(tv.L.V);
    const uint64_t lowc = count & vmask;
    uint64_t ov = tv.L.V >= 64 ? 0 : (count >> tv.L.V);
    uint64_t seen = tv.f_hasovf;  // assume set unless we read otherwise
    if (lowc) {
        const uint64_t old = atomicAdd((unsigned long long*)hp, (unsigned long long)(lowc << tv.vshift));
        ov += ((old >> tv.vshift) + lowc) >> tv.L.V;
        seen = old & tv.f_hasovf;
    } else if (ov) {
        seen = 0;
    }
    if (ov) {
        overflow_add<KW, W>(tv, home, i, pslot, ov, st);
        if (!seen) atomicOr((unsigned long long*)hp, (unsigned long long)tv.f_hasovf);
    }
}

// Insert-or-increment of one hashed k-mer.  Reference: addKmer, TSXHashMapPerf.h:56-205 /
// TSXHashMapCAS.h:268-508.  Lock-free: an entry's key bits are immutable once written, a thread moves
// past a slot only after it has seen it occupied by a different key, and every thread claims the
// lowest free slot of a bucket, so a k-mer can never be stored twice.
//
// LEAN (k_insert_keys): the statistics that change with every k-mer are reduced to `distinct`, so that the kernel fits
// 40 registers without spilling in its inner loop.  The caller credits all its keys to CTR_ADDED up front and this
// function takes back the ones it does not insert; the reprobe limit is published at once (the kernel polls the flag
// per slice) instead of being remembered per thread.
template <int KW, int W, bool LEAN = false>
__device__ __forceinline__ void insert_hashed(const TableView& tv, const Key<KW>& H, uint64_t count, LocalStats& st) {
    constexpr int SPB = 4 / W;
    const Stored<KW, W> s = make_stored<KW, W>(tv, H);
    if (!s.owned) { st.errors |= ERR_WRONG_SHARD; if (LEAN) st.added -= count; return; }
    if (!LEAN) {
        if (st.errors & ERR_TABLE_FULL) return;    // this thread already hit the reprobe limit: the run is lost (exit 42)
        st.added += count;
    }
    const uint64_t vmask = low_mask(tv.L.V);
    for (uint32_t i = 1; i <= tv.L.max_probe; ++i) {
        uint64_t* bp = tv.words + (((s.bucket + tri(i)) & tv.lbl_mask) << 2);
        const uint64_t pattern = s.headkey | i;
        uint64_t w[4];
        load_bucket(bp, w);
        bool retry_bucket = false;
#pragma unroll
        for (int sl = 0; sl < SPB; ++sl) {
            uint64_t* ep = bp + sl * W;
            uint64_t* hp = ep + (W - 1);
            uint64_t h = w[sl * W + (W - 1)];
            const bool empty = (h == 0);  // i >= 1 keeps every occupied head non-zero
            if (empty) {
                // claim: value = count mod 2^V, the rest goes to the overflow entry
                const uint64_t nh = pattern | ((count & vmask) << tv.vshift);
                bool won;
                if (W == 1) {
                    const uint64_t old = atomicCAS((unsigned long long*)hp, 0ULL, (unsigned long long)nh);
                    won = (old == 0); h = old;
                } else if (W == 2) {
                    uint64_t o0, o1;
                    cas128(ep, 0, 0, s.body[0], nh, o0, o1);
                    won = (o0 == 0 && o1 == 0); h = o1; w[sl * W] = o0;
                } else {
                    const uint64_t old = atomicCAS((unsigned long long*)hp, 0ULL, (unsigned long long)(nh | tv.f_busy));
                    won = (old == 0); h = old;
                    if (won) {
                        // publish the key words, then drop BUSY (claim-then-publish; the reference's
                        // byte-wise CAS with rollback, TSXHashMapCAS.h:196-225, is what this replaces)
#pragma unroll
                        for (int j = 0; j < W - 1; ++j) __stcg(ep + j, s.body[j]);
                        __threadfence();
                        atomicAnd((unsigned long long*)hp, ~(unsigned long long)tv.f_busy);
                    }
                }
                if (won) {
                    st.distinct++;
                    if (i > 1 && i > st.maxprobe) st.maxprobe = i;     // probe 1 is implied by distinct > 0 (flush_stats)
                    const uint64_t ov = tv.L.V >= 64 ? 0 : (count >> tv.L.V);
                    if (ov) {
                        overflow_add<KW, W>(tv, s.bucket, i, sl, ov, st);
                        atomicOr((unsigned long long*)hp, (unsigned long long)tv.f_hasovf);
                    }
                    return;
                }
            }
            if ((h & tv.cmp_mask) == pattern) {
                bool same = true;
                if (W == 4) {
                    if (h & tv.f_busy) { retry_bucket = true; break; }  // key words not published yet
                    // body words in w[] may predate the publication: reload them.  The fence is the acquire side of
                    // the publication (body stores, fence, BUSY cleared): without it these loads could be satisfied
                    // before the head load that showed BUSY == 0 and return the stale zero body.
                    if (tv.L.Q > tv.L.Qh) {
                        __threadfence();
#pragma unroll
                        for (int j = 0; j < W - 1; ++j) same &= (__ldcg(ep + j) == s.body[j]);
                    }
                } else if (W == 2) {
                    same = (w[sl * W] == s.body[0]);
                    // both words are written by one 128-bit CAS; re-read once in case the two
                    // 8-byte halves of our 16-byte load were not observed together
                    if (!same && w[sl * W] == 0) same = (__ldcg(ep) == s.body[0]);
                }
                if (same) {
                    add_to_primary<KW, W>(tv, hp, s.bucket, i, sl, count, st);
                    return;
                }
            }
        }
        if (retry_bucket) { --i; continue; }
    }
    st.errors |= ERR_TABLE_FULL;
    if (LEAN) atomicOr(tv.ctr + CTR_ERRORS, (unsigned long long)ERR_TABLE_FULL);
}

// Reference: findOverflowCounts, TSXHashMap.h:951-1039
template <int KW, int W>
__device__ __forceinline__ uint64_t overflow_lookup(const TableView& tv, uint64_t home, uint32_t i, uint32_t pslot) {
    constexpr int SPB = 4 / W;
    const uint64_t tag0 = (uint64_t)i | tv.f_ovf | ((uint64_t)pslot << (tv.L.R + 3));
    for (uint32_t j = 1; j <= tv.L.max_probe; ++j) {
        const uint64_t tag = tag0 | ((uint64_t)j << (tv.L.R + 5));
        const uint64_t* bp = tv.words + (((home + tri(i + j)) & tv.lbl_mask) << 2);
        uint64_t w[4];
        load_bucket(bp, w);
#pragma unroll
        for (int sl = 0; sl < SPB; ++sl) {
            const uint64_t h = w[sl * W + (W - 1)];
            if (h == 0) return 0;
            if ((h & tv.tag_mask) == tag) return h >> tv.L.cshift;
        }
    }
    return 0;
}

// Reference: getKmerCount(kmer), TSXHashMap.h:548-638 — probe until an empty slot; on a key+reprobe
// match read the value and add the overflow chain: count = (overflow << s) | value (:600-609).
template <int KW, int W>
__device__ __forceinline__ uint64_t lookup_hashed(const TableView& tv, const Key<KW>& H) {
    constexpr int SPB = 4 / W;
    const Stored<KW, W> s = make_stored<KW, W>(tv, H);
    if (!s.owned) return 0;
    for (uint32_t i = 1; i <= tv.L.max_probe; ++i) {
        const uint64_t* bp = tv.words + (((s.bucket + tri(i)) & tv.lbl_mask) << 2);
        const uint64_t pattern = s.headkey | i;
        uint64_t w[4];
        load_bucket(bp, w);
#pragma unroll
        for (int sl = 0; sl < SPB; ++sl) {
            const uint64_t h = w[sl * W + (W - 1)];
            if (h == 0) return 0;
            if ((h & tv.cmp_mask) == pattern) {
                bool same = true;
#pragma unroll
                for (int j = 0; j < W - 1; ++j) same &= (w[sl * W + j] == s.body[j]);
                if (same) {
                    uint64_t c = h >> tv.vshift;
                    if (h & tv.f_hasovf) c += overflow_lookup<KW, W>(tv, s.bucket, i, sl) << tv.L.V;
                    return c;
                }
            }
        }
    }
    return 0;
}

// Rebuild the hash of the primary entry stored at (bucket, slot).  Reference: getAllKmers,
// TSXHashMap.h:660-722: kmer = inv_apply(func || ((pos - i(i+1)/2) mod 2^L)).
template <int KW, int W>
__device__ __forceinline__ Key<KW> hash_of_entry(const TableView& tv, uint64_t bucket, const uint64_t* e) {
    const uint64_t h = e[W - 1];
    const uint32_t i = (uint32_t)(h & tv.rmask);
    Key<KW> q;
#pragma unroll
    for (int j = 0; j < KW; ++j) q.w[j] = (W > 1 && j < W - 1) ? e[j] : 0ULL;
    q = shl_small<KW>(q, tv.L.Qh);
    q.w[0] |= (h >> tv.qshift) & tv.qh_mask;
    const uint64_t home = (bucket - tri(i)) & tv.lbl_mask;
    const uint64_t bg = ((uint64_t)tv.L.shard_rank << tv.L.LBl) | home;
    Key<KW> H = shl_small<KW>(q, tv.L.LBg);
    H.w[0] |= bg;
    return H;
}

#endif  // __CUDACC__

}  // namespace tsx
