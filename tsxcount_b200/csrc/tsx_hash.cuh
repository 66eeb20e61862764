// tsx_hash.cuh — fixed-width k-mer keys and the bijective hash.
//
// Replaces (reference paths relative to mjoppich/tsxCount):
//   src/tsxutils/UBigInt.h:105-1673        heap-allocated byte-array big integers  -> Key<KW> in registers
//   src/tsxcount/BijectiveKMapping.h:202-225  key = A·kmer over GF(2), A random unit upper-triangular
//   src/tsxcount/BijectiveKMapping.h:643-766  inverse matrix (int8 LU)
// The reference hash costs O(k^2) bit operations and makes the slot index a function of the first
// L/2 bases only.  Only bijectivity is observable (TSXHashMap::getAllKmers inverts it,
// TSXHashMap.h:660-722; testHashFunction checks the round trip, :724-735), so this file defines a new
// invertible mixer on exactly 2k bits:
//   KW=1 (2k <= 64)  : murmur3-finaliser shape reduced mod 2^(2k): xorshift(ceil(n/2)) / odd multiply,
//                      every step a bijection on n bits.
//   KW=2,4 (2k<=256) : unbalanced Feistel over 64-bit words with the (bijective) 64-bit finaliser as
//                      round function; the top partial word is masked so the image stays inside 2k bits.
// Word 0 of the hash ends up fully mixed; the bucket index is taken from its low bits.
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define TSX_HD __host__ __device__ __forceinline__
#else
#define TSX_HD inline
#endif

namespace tsx {

template <int KW>
struct Key {
    uint64_t w[KW];
};

template <int KW>
TSX_HD bool key_eq(const Key<KW>& a, const Key<KW>& b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < KW; ++i) e &= (a.w[i] == b.w[i]);
    return e;
}

TSX_HD uint64_t low_mask(unsigned bits) {  // bits in [0,64]
    return bits >= 64 ? ~0ULL : ((1ULL << bits) - 1ULL);
}

// murmur3 fmix64 and its inverse (both bijections on 64 bits)
constexpr uint64_t kM1 = 0xff51afd7ed558ccdULL;
constexpr uint64_t kM2 = 0xc4ceb9fe1a85ec53ULL;
constexpr uint64_t mul_inverse(uint64_t a) {  // a odd; Newton iteration mod 2^64
    uint64_t x = a;
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
constexpr uint64_t kM1Inv = mul_inverse(kM1);
constexpr uint64_t kM2Inv = mul_inverse(kM2);
static_assert(kM1 * kM1Inv == 1ULL && kM2 * kM2Inv == 1ULL, "modular inverses");

TSX_HD uint64_t fmix64(uint64_t h) {
    h ^= h >> 33; h *= kM1;
    h ^= h >> 33; h *= kM2;
    h ^= h >> 33;
    return h;
}
TSX_HD uint64_t unfmix64(uint64_t h) {
    h ^= h >> 33; h *= kM2Inv;
    h ^= h >> 33; h *= kM1Inv;
    h ^= h >> 33;
    return h;
}

constexpr uint64_t kC1 = 0x9E3779B97F4A7C15ULL;
constexpr uint64_t kC2 = 0xD6E8FEB86659FD93ULL;
constexpr uint64_t kC3 = 0xA0761D6478BD642FULL;
constexpr uint64_t kC4 = 0xE7037ED1A0B428DBULL;

// Hash parameters derived from k once on the host.
struct HashParams {
    uint32_t nbits;     // 2k
    uint32_t top_word;  // index of the highest word holding key bits
    uint64_t top_mask;  // valid bits of that word
    uint32_t xs;        // KW=1: xorshift distance ceil(n/2)
    uint32_t canonical; // 1: every k-mer is replaced by min(k-mer, reverse complement) before it is hashed
};

inline HashParams make_hash_params(uint32_t k, bool canonical = false) {
    HashParams p{};
    p.canonical = canonical ? 1u : 0u;
    p.nbits = 2 * k;
    p.top_word = (p.nbits - 1) / 64;
    p.top_mask = low_mask(p.nbits - 64 * p.top_word);
    p.xs = (p.nbits + 1) / 2;
    return p;
}

// ---- KW = 1 -----------------------------------------------------------------------------------
TSX_HD uint64_t hash1(uint64_t x, const HashParams& p) {
    const uint64_t m = p.top_mask;
    x ^= x >> p.xs; x = (x * kM1) & m;
    x ^= x >> p.xs; x = (x * kM2) & m;
    x ^= x >> p.xs;
    return x;
}
TSX_HD uint64_t unhash1(uint64_t x, const HashParams& p) {
    const uint64_t m = p.top_mask;
    x ^= x >> p.xs; x = (x * kM2Inv) & m;  // xorshift by >= n/2 is an involution on n bits
    x ^= x >> p.xs; x = (x * kM1Inv) & m;
    x ^= x >> p.xs;
    return x;
}

// ---- generic ----------------------------------------------------------------------------------
template <int KW>
TSX_HD uint64_t word_mask(int j, const HashParams& p) {
    return (uint32_t)j < p.top_word ? ~0ULL : ((uint32_t)j == p.top_word ? p.top_mask : 0ULL);
}

// ---- canonical k-mers (opt-in; the reference counts forward k-mers only and has no reverse-complement code) ----------
// Base i sits at bits [2i, 2i+1], A=0 C=1 G=2 T=3, so the complement of a base is its bitwise NOT and the reverse
// complement is: NOT inside the 2k bits, reverse the order of the 2-bit groups, shift down to bit 0.
// "The lexicographically smaller of the k-mer and its reverse complement" (jellyfish -C, KMC; the tests' checker
// computes it on the text) is the NUMERICALLY smaller of the two encodings although base 0 is the least significant
// digit: lex(f) < lex(r)  <=>  rev(f) < rev(r)  <=>  NOT r < NOT f  <=>  f < r, because rev(f) = NOT r and rev(r) = NOT f.
TSX_HD uint64_t rev2_64(uint64_t x) {   // reverses the order of the 32 2-bit groups of a word
#if defined(__CUDA_ARCH__)
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0f0f0f0f0f0f0f0fULL) | ((x & 0x0f0f0f0f0f0f0f0fULL) << 4);
    x = ((x >> 8) & 0x00ff00ff00ff00ffULL) | ((x & 0x00ff00ff00ff00ffULL) << 8);
    x = ((x >> 16) & 0x0000ffff0000ffffULL) | ((x & 0x0000ffff0000ffffULL) << 16);
    x = (x >> 32) | (x << 32);
#endif
    return ((x & 0x5555555555555555ULL) << 1) | ((x >> 1) & 0x5555555555555555ULL);   // the bit pairs back in order
}

template <int KW>
TSX_HD Key<KW> revcomp_key(const Key<KW>& x, const HashParams& p) {
    Key<KW> r;
#pragma unroll
    for (int j = 0; j < KW; ++j) r.w[j] = rev2_64(~x.w[KW - 1 - j] & word_mask<KW>(KW - 1 - j, p));
    const unsigned s = 64u * KW - p.nbits, ws = s >> 6, bs = s & 63u;
    Key<KW> out;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        uint64_t lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < KW; ++i) {          // r.w[j + ws] and r.w[j + ws + 1] without dynamic indexing
            if ((unsigned)i == (unsigned)j + ws) lo = r.w[i];
            if ((unsigned)i == (unsigned)j + ws + 1) hi = r.w[i];
        }
        out.w[j] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
    return out;
}

template <int KW>
TSX_HD Key<KW> canonical_key(const Key<KW>& x, const HashParams& p) {
    const Key<KW> r = revcomp_key<KW>(x, p);
    bool r_less = false, decided = false;
#pragma unroll
    for (int j = KW - 1; j >= 0; --j) {
        if (!decided && r.w[j] != x.w[j]) { r_less = r.w[j] < x.w[j]; decided = true; }
    }
    return r_less ? r : x;
}

template <int KW>
TSX_HD Key<KW> hash_key(const Key<KW>& x_in, const HashParams& p) {
    Key<KW> h;
    const Key<KW> x = p.canonical ? canonical_key<KW>(x_in, p) : x_in;
    if constexpr (KW == 1) {
        h.w[0] = hash1(x.w[0], p);
    } else {
        // 1. mix word 0;  2. fold it into every higher word;  3. fold the higher words back into word 0
        uint64_t x0 = fmix64(x.w[0]);
        uint64_t chain = kC1;
#pragma unroll
        for (int j = 1; j < KW; ++j) {
            h.w[j] = x.w[j] ^ (fmix64(x0 + (uint64_t)j * kC2) & word_mask<KW>(j, p));
            chain = fmix64(h.w[j] + chain);
        }
        h.w[0] = x0 ^ chain;
        if (p.top_word == 0) h.w[0] = 0;  // unreachable (KW>1 implies nbits>64); keeps the compiler honest
    }
    return h;
}

template <int KW>
TSX_HD Key<KW> unhash_key(const Key<KW>& h, const HashParams& p) {
    Key<KW> x;
    if constexpr (KW == 1) {
        x.w[0] = unhash1(h.w[0], p);
    } else {
        uint64_t chain = kC1;
#pragma unroll
        for (int j = 1; j < KW; ++j) chain = fmix64(h.w[j] + chain);
        const uint64_t x0 = h.w[0] ^ chain;
#pragma unroll
        for (int j = 1; j < KW; ++j) x.w[j] = h.w[j] ^ (fmix64(x0 + (uint64_t)j * kC2) & word_mask<KW>(j, p));
        x.w[0] = unfmix64(x0);
    }
    return x;
}

// ---- multiword shifts -------------------------------------------------------------------------
// (hash >> sh) for 0 <= sh < 64, result keeps KW words
template <int KW>
TSX_HD Key<KW> shr_small(const Key<KW>& a, unsigned sh) {
    Key<KW> r;
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        uint64_t lo = a.w[j] >> sh;
        uint64_t hi = (j + 1 < KW && sh) ? (a.w[j + 1] << (64 - sh)) : 0ULL;
        r.w[j] = lo | hi;
    }
    return r;
}
template <int KW>
TSX_HD Key<KW> shl_small(const Key<KW>& a, unsigned sh) {
    Key<KW> r;
#pragma unroll
    for (int j = KW - 1; j >= 0; --j) {
        uint64_t hi = a.w[j] << sh;
        uint64_t lo = (j > 0 && sh) ? (a.w[j - 1] >> (64 - sh)) : 0ULL;
        r.w[j] = hi | lo;
    }
    return r;
}

}  // namespace tsx
