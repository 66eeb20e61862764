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
#include <cstdint>
#include <cstring>

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
}  // namespace

extern "C" int tsxc_pack_reads(const char* ascii, const uint64_t* offsets, uint64_t n_reads, uint64_t* packed_out,
                               uint64_t* seg_offsets_out, uint64_t seg_capacity, uint64_t* n_segments_out,
                               uint64_t* n_bad_bases_out) {
    if (!offsets || !packed_out || !seg_offsets_out || !n_segments_out || seg_capacity < 1) return TSXC_E_INVALID;
    if (n_reads && !ascii) return TSXC_E_INVALID;
    uint64_t nseg = 0, g = 0, bad = 0;  // g = bases written so far
    uint64_t acc = 0;                   // word under construction
    seg_offsets_out[0] = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint64_t b = offsets[r], e = offsets[r + 1];
        uint64_t seg_start = g;
        for (uint64_t i = b; i < e; ++i) {
            const int c = kLut.v[(unsigned char)ascii[i]];
            if (c < 0) {
                ++bad;
                if (g > seg_start) {
                    if (nseg + 1 >= seg_capacity) return TSXC_E_INVALID;
                    seg_offsets_out[++nseg] = g;
                }
                seg_start = g;
                continue;
            }
            acc |= (uint64_t)c << (2 * (g & 31));
            if ((++g & 31) == 0) { packed_out[(g >> 5) - 1] = acc; acc = 0; }
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
