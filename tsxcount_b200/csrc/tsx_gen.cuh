// tsx_gen.cuh — device generator of the synthetic read sets (DESIGN.md "Synthetic inputs").
// Style of the reference's generateFakeSequences.py:7-18 (random ACGT body + poly-A tail), made
// counter-based and integer-only so the same reads can be produced on the host (oracle/oracle.c has an
// independent restatement used by the tests) and on the device without shipping them over PCIe.
#pragma once

#include <cstdint>

#include "tsx_hash.cuh"

namespace tsx {

struct GenParams {
    uint64_t seed;
    uint64_t n_reads;
    uint32_t read_len;
    uint32_t mode;
    uint64_t genome_len;
    uint32_t sub_rate_q16;
    uint32_t dict_log2;
};

TSX_HD uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

TSX_HD unsigned gen_base(const GenParams& p, uint64_t gkey, uint64_t r, uint32_t q) {
    const uint64_t hr = splitmix(p.seed ^ splitmix(r + 0x1234567ULL));
    const uint32_t len = p.read_len;
    uint64_t start;
    uint32_t body = len, tail = 0;
    switch (p.mode) {
        case 1:
            body = len / 2 + (uint32_t)((hr >> 8) % (len / 3 + 1));
            tail = len / 6 + (uint32_t)((hr >> 40) % (len / 6 + 1));
            start = r * (uint64_t)len;
            break;
        case 2: {
            const unsigned lg = p.dict_log2;
            const unsigned b = (unsigned)((hr >> 48) % (lg + 1));
            const uint64_t rank = (hr & ((1ULL << lg) - 1)) >> (lg - b);
            start = rank * (uint64_t)len;
            break;
        }
        case 3:
            start = hr % (p.genome_len - len + 1);
            break;
        default:
            start = r * (uint64_t)len;
            break;
    }
    const uint64_t g = start + q;
    const uint64_t w = splitmix(gkey + (g >> 5));
    unsigned b = (unsigned)(w >> (2 * (g & 31))) & 3u;
    if (p.mode == 1 && q >= body && q < body + tail) b = 0;
    if (p.sub_rate_q16) {
        const uint64_t hs = splitmix(hr + 0x51ED27ULL * (uint64_t)(q + 1));
        if ((hs & 0xFFFF) < p.sub_rate_q16) b = (b + 1 + (unsigned)((hs >> 16) % 3)) & 3u;
    }
    return b;
}

#if defined(__CUDACC__)
// One thread per packed output word (32 bases); reads are laid end to end, fixed length.
__global__ void __launch_bounds__(256) k_gen_reads(GenParams p, uint64_t first, uint64_t count, uint64_t* __restrict__ packed,
                                                   uint64_t* __restrict__ offsets) {
    const uint64_t n_bases = count * p.read_len;
    const uint64_t n_words = (n_bases + 31) >> 5;
    const uint64_t gkey = splitmix(p.seed ^ 0xD6E8FEB86659FD93ULL);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t wdx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wdx < n_words; wdx += stride) {
        uint64_t out = 0;
        if (p.mode == 0 && p.sub_rate_q16 == 0) {
            // independent uniform reads are one contiguous slice of the base stream
            const uint64_t g = first * p.read_len + (wdx << 5);
            const unsigned sh = 2u * (unsigned)(g & 31);
            const uint64_t a = splitmix(gkey + (g >> 5));
            out = a >> sh;
            if (sh) out |= splitmix(gkey + (g >> 5) + 1) << (64 - sh);
            const uint64_t left = n_bases - (wdx << 5);
            if (left < 32) out &= low_mask(2u * (unsigned)left);
        } else {
            uint64_t g = wdx << 5;
            uint64_t r = g / p.read_len;
            uint32_t q = (uint32_t)(g - r * p.read_len);
            for (int b = 0; b < 32 && g < n_bases; ++b, ++g) {
                out |= (uint64_t)gen_base(p, gkey, first + r, q) << (2 * b);
                if (++q == p.read_len) { q = 0; ++r; }
            }
        }
        packed[wdx] = out;
    }
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= count; r += stride) offsets[r] = r * p.read_len;
}
#endif

}  // namespace tsx
