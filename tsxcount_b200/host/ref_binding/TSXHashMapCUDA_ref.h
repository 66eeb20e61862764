// TSXHashMapCUDA_ref.h — the --mode=CUDA member of the reference's TSXHashMap family, written against the
// reference's OWN headers.  Unlike ../TSXHashMapCUDA.h (a stand-alone class for hosts without
// the reference tree) this file only compiles inside / next to mjoppich/tsxCount:
//     g++ -std=c++14 -I<reference> -I<reference>/src -I<repo>/include ... -ltsxcuda
// It derives from TSXHashMap (src/tsxcount/TSXHashMap.h:68) and overrides the virtual per-k-mer interface the
// reference's driver and --check use: addKmer (:182), getKmerCount(kmer) (:548); getKmerCount() (:645) is shadowed.
// The test-side Makefile builds a small check binary from it (test infrastructure; the reference sources stay where
// they are), tests/test_cli.py runs that binary on the GPU.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include <src/tsxcount/TSXHashMap.h>

#include "tsxcount_cuda.h"

class TSXHashMapCUDA : public TSXHashMap {
public:
    // The base class allocates its own bit-packed host table in the constructor (TSXHashMap.h:79-154); it is not
    // used here, so the base gets the smallest legal one (2^2 slots) and the real 2^iL-slot table lives in HBM.
    TSXHashMapCUDA(uint8_t iL, uint32_t iStorageBits, uint16_t iK, uint8_t iThreads = 1, int device = 0)
        : TSXHashMap(2, iStorageBits, iK), m_kw(tsxc_key_words(iK)) {
        const int rc = tsxc_create(iK, iL, iStorageBits, device, TSXC_FLAG_EXACT_S, &m_h);
        if (rc != TSXC_OK) throw TSXException(std::string(tsxc_status_string(rc)) + ": " + tsxc_last_error(nullptr));
        this->setThreads(iThreads);
    }
    ~TSXHashMapCUDA() override { tsxc_destroy(m_h); }

    // UBigInt (bit i of the k-mer = bit i of the number, SequenceUtils.h:96-123) -> KW little-endian words
    void toWords(TSX::tsx_kmer_t& kmer, uint64_t* w) {
        for (uint32_t j = 0; j < 4; ++j) w[j] = 0;
        const uint32_t bits = 2 * this->getK();
        for (uint32_t i = 0; i < bits && i < kmer.getBitCount(); ++i)
            if (kmer.getBit(i)) w[i >> 6] |= 1ULL << (i & 63);
    }

    bool addKmer(TSX::tsx_kmer_t& kmer, bool verbose = false, bool noPrimaryAddition = false) override {
        (void)verbose; (void)noPrimaryAddition;
        uint64_t w[4];
        toWords(kmer, w);
        check(tsxc_add_kmers(m_h, w, 1));
        check(tsxc_sync(m_h));
        return true;
    }
    UBigInt getKmerCount(TSX::tsx_kmer_t& kmer, bool verbose = false, uint32_t addReprobes = 0) override {
        (void)verbose; (void)addReprobes;
        uint64_t w[4], c = 0;
        toWords(kmer, w);
        check(tsxc_lookup(m_h, w, 1, &c));
        return UBigInt(c, this->getMemoryPool());
    }
    uint64_t getKmerCount() { uint64_t n = 0; check(tsxc_distinct(m_h, &n)); return n; }     // TSXHashMap.h:645-648

    // batch entry points (what countKMers, main.cpp:141-206, should call for this mode)
    void addKmers(std::vector<TSX::tsx_kmer_t>& kmers) {
        std::vector<uint64_t> w(kmers.size() * m_kw);
        uint64_t tmp[4];
        for (size_t i = 0; i < kmers.size(); ++i) {
            toWords(kmers[i], tmp);
            for (uint32_t j = 0; j < m_kw; ++j) w[i * m_kw + j] = tmp[j];
        }
        check(tsxc_add_kmers(m_h, w.data(), kmers.size()));
        check(tsxc_sync(m_h));
    }
    // ASCII reads -> 2-bit pack -> extraction + insert on the device (createKMers + fromSequence + addKmer)
    void addReads(const std::vector<std::string>& seqs) {
        std::string ascii;
        std::vector<uint64_t> off(1, 0);
        for (const auto& s : seqs) { ascii += s; off.push_back(ascii.size()); }
        std::vector<uint64_t> packed(ascii.size() / 32 + 2), seg(seqs.size() + ascii.size() + 2);
        uint64_t n_seg = 0, n_bad = 0;
        check(tsxc_pack_reads(ascii.data(), off.data(), seqs.size(), packed.data(), seg.data(), seg.size(), &n_seg, &n_bad));
        check(tsxc_add_reads(m_h, packed.data(), seg.data(), n_seg));
        check(tsxc_sync(m_h));
    }
    tsxc_table* handle() { return m_h; }

private:
    void check(int rc) {
        if (rc == TSXC_OK) return;
        if (rc == TSXC_E_TABLE_FULL) exit(42);                                                 // TSXHashMap.h:340-343
        throw TSXException(std::string(tsxc_status_string(rc)) + ": " + tsxc_last_error(m_h));
    }
    tsxc_table* m_h = nullptr;
    uint32_t m_kw;
};
