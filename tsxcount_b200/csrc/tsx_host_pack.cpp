// tsx_host_pack.cpp — host-side 2-bit packing of ASCII reads (feeder side of the boundary).
//
// Replaces TSXSeqUtils::fromSequence (src/utils/SequenceUtils.h:86-160 of mjoppich/tsxCount) as used
// per k-mer in src/mains/main.cpp:173: instead of packing each k-mer separately, the whole read is
// packed once with the same digit order (base i -> bits [2i,2i+1], A=0 C=1 G=2 T=3), which makes every
// k-mer a contiguous bit window of the stream.
//
// N policy: the reference writes two rand()%2 bits for any byte outside upper-case ACGT (:126-137) —
// unseeded and called from OpenMP tasks, hence not reproducible.  Here such a byte ends the current
// segment: each maximal ACGT run is emitted as its own segment, so no k-mer spans it.
//
// Fast path: 8 bases at a time.  For the four valid letters ((c >> 1) ^ (c >> 2)) & 3 is exactly the code
// (A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 -> 3); a SWAR test proves that all 8 bytes are valid letters,
// anything else drops to the byte-wise path.  Where the CPU has AVX2 (checked at run time) 32 bases are packed per
// step into one whole output word.
#include <cstdint>
#include <cstring>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define TSXC_HAVE_AVX2_PATH 1
#endif

#include "../../include/tsxcount_cuda.h"

namespace {
struct Lut {
    int8_t v[256];
    Lut() {
        std::memset(v, -1, sizeof v);
        v[(unsigned char)'A'] = 0; v[(unsigned char)'C'] = 1; v[(unsigned char)'G'] = 2; v[(unsigned char)'T'] = 3;
    }
};
const Lut kLut;

constexpr uint64_t kLo7 = 0x7f7f7f7f7f7f7f7fULL, kHi = 0x8080808080808080ULL;
// 0x80 in every byte of x that is zero
inline uint64_t zero_bytes(uint64_t x) { return ~(((x & kLo7) + kLo7) | x) & kHi; }
inline uint64_t rep(uint8_t b) { return 0x0101010101010101ULL * b; }

// 16 bits = the 2-bit codes of 8 valid letters (byte i -> bits [2i, 2i+1])
inline uint64_t codes8(uint64_t x) {
    uint64_t y = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ULL;   // one code per byte
    y = (y | (y >> 6)) & 0x000f000f000f000fULL;                   // two codes per 16 bits
    y = (y | (y >> 12)) & 0x000000ff000000ffULL;                  // four codes per 32 bits
    return (y | (y >> 24)) & 0xffffULL;                           // eight codes
}
inline bool all_acgt(uint64_t x) {
    const uint64_t ok = zero_bytes(x ^ rep('A')) | zero_bytes(x ^ rep('C')) | zero_bytes(x ^ rep('G')) | zero_bytes(x ^ rep('T'));
    return ok == kHi;
}

#ifdef TSXC_HAVE_AVX2_PATH
// 32 ASCII bases -> 64 bits of codes (base i at bits [2i, 2i+1]); false if any byte is not one of ACGT
__attribute__((target("avx2"))) inline bool codes32_avx2(const char* p, uint64_t* out) {
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('C'))),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('T'))));
    if ((unsigned)_mm256_movemask_epi8(ok) != 0xffffffffu) return false;
    // per byte: ((c >> 1) ^ (c >> 2)) & 3; the 16-bit shifts only leak neighbour bits above bit 5
    const __m256i c = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), _mm256_set1_epi8(3));
    const __m256i n = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0401));       // 16-bit lanes: c0 + 4*c1
    const __m256i b = _mm256_madd_epi16(n, _mm256_set1_epi32(0x00100001));      // 32-bit lanes: n0 + 16*n1 (one byte)
    // byte 0 of every dword -> the low 4 bytes of each 128-bit half
    const __m256i sh = _mm256_shuffle_epi8(b, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                                0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    const uint64_t lo = (uint32_t)_mm256_extract_epi32(sh, 0), hi = (uint32_t)_mm256_extract_epi32(sh, 4);
    *out = lo | (hi << 32);
    return true;
}
const bool kHaveAvx2 = __builtin_cpu_supports("avx2");
#endif
}  // namespace

extern "C" int tsxc_pack_reads(const char* ascii, const uint64_t* offsets, uint64_t n_reads, uint64_t* packed_out,
                               uint64_t* seg_offsets_out, uint64_t seg_capacity, uint64_t* n_segments_out,
                               uint64_t* n_bad_bases_out) {
    if (!offsets || !packed_out || !seg_offsets_out || !n_segments_out || seg_capacity < 1) return TSXC_E_INVALID;
    if (n_reads && !ascii) return TSXC_E_INVALID;
    uint64_t nseg = 0, g = 0, bad = 0;  // g = bases written so far
    uint64_t acc = 0;                   // word under construction (bases g - g%32 .. g-1)
    seg_offsets_out[0] = 0;
    auto put_codes = [&](uint64_t codes, unsigned n) {   // append n <= 8 bases (2n bits)
        const unsigned pos = (unsigned)(g & 31);
        acc |= codes << (2 * pos);
        if (pos + n >= 32) {
            packed_out[g >> 5] = acc;
            const unsigned used = 32 - pos;               // bases that went into the finished word
            acc = used < n ? codes >> (2 * used) : 0;
        }
        g += n;
    };
    auto put_word = [&](uint64_t codes) {                 // append 32 bases (64 bits)
        const unsigned pos = (unsigned)(g & 31);
        if (pos == 0) {
            packed_out[g >> 5] = codes;
        } else {
            packed_out[g >> 5] = acc | (codes << (2 * pos));
            acc = codes >> (64 - 2 * pos);
        }
        g += 32;
    };
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        uint64_t seg_start = g;
        uint64_t i = b;
        while (i < e) {
#ifdef TSXC_HAVE_AVX2_PATH
            if (kHaveAvx2 && i + 32 <= e) {
                uint64_t c64;
                if (codes32_avx2(ascii + i, &c64)) { put_word(c64); i += 32; continue; }
            }
#endif
            if (i + 8 <= e) {
                uint64_t x;
                std::memcpy(&x, ascii + i, 8);
                if (all_acgt(x)) { put_codes(codes8(x), 8); i += 8; continue; }
            }
            const int c = kLut.v[(unsigned char)ascii[i]];
            ++i;
            if (c < 0) {
                ++bad;
                if (g > seg_start) {
                    if (nseg + 1 >= seg_capacity) return TSXC_E_INVALID;
                    seg_offsets_out[++nseg] = g;
                }
                seg_start = g;
                continue;
            }
            put_codes((uint64_t)c, 1);
        }
        // a read always closes a segment, even an empty one (read count == segment count on clean input;
        // empty segments contribute no k-mers)
        if (nseg + 1 >= seg_capacity) return TSXC_E_INVALID;
        seg_offsets_out[++nseg] = g;
    }
    if (g & 31) packed_out[g >> 5] = acc;
    *n_segments_out = nseg;
    if (n_bad_bases_out) *n_bad_bases_out = bad;
    return TSXC_OK;
}
