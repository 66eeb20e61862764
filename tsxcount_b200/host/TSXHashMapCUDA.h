// TSXHashMapCUDA.h — C++ adapter with the shape of the reference's TSXHashMap over the C ABI.
//
// The reference picks a serialization backend by constructing a TSXHashMap subclass in
// src/mains/main.cpp:429-475 (TSXHashMapPerf / PThreadPerf / OMPPerf / CAS / TSXPerf ...) and drives it
// through the virtual interface of src/tsxcount/TSXHashMap.h:68.  This class is the --mode=CUDA member of
// that family: same constructor arguments and getter names; the per-k-mer virtual addKmer(UBigInt&) is
// kept for single k-mers, and batch entry points are added because a GPU is fed batches.
// Errors: TSXException (TSXHashMap.h:28-47) for invalid parameters, like the reference (:91-94);
// a full table ends the process with exit(42) like TSXHashMap.h:340-343.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsxcount_cuda.h"

class TSXException : public std::runtime_error {
public:
    explicit TSXException(const std::string& text) : std::runtime_error(text) {}
};

class TSXHashMapCUDA {
public:
    // TSXHashMap(uint8_t iL, uint32_t iStorageBits, uint16_t iK[, uint8_t threads]) — TSXHashMap.h:79
    TSXHashMapCUDA(uint8_t iL, uint32_t iStorageBits, uint16_t iK, int device = 0, uint32_t flags = TSXC_FLAG_EXACT_S)
        : m_iL(iL), m_iStorageBits(iStorageBits), m_iK(iK), m_kw(tsxc_key_words(iK)) {
        const int rc = tsxc_create(iK, iL, iStorageBits, device, flags, &m_h);
        if (rc != TSXC_OK) throw TSXException(std::string(tsxc_status_string(rc)) + ": " + tsxc_last_error(nullptr));
    }
    ~TSXHashMapCUDA() { tsxc_destroy(m_h); }
    TSXHashMapCUDA(const TSXHashMapCUDA&) = delete;
    TSXHashMapCUDA& operator=(const TSXHashMapCUDA&) = delete;

    uint32_t getK() const { return m_iK; }                                  // TSXHashMap.h:172
    uint32_t keyWords() const { return m_kw; }
    uint64_t getMaxElements() { return stats().n_slots; }                   // TSXHashMap.h:162
    uint64_t getUsedPositions() { return stats().used_slots; }              // TSXHashMap.h:177
    tsxc_table* handle() { return m_h; }

    // bool addKmer(TSX::tsx_kmer_t& kmer) — TSXHashMap.h:182 (one k-mer, KW little-endian words)
    bool addKmer(const uint64_t* kmer) { check(tsxc_add_kmers(m_h, kmer, 1)); sync(); return true; }
    void addKmers(const uint64_t* kmers, uint64_t n) { check(tsxc_add_kmers(m_h, kmers, n)); }
    // createKMers + fromSequence + addKmer for a packed batch — main.cpp:159-192
    void addReads(const uint64_t* packed, const uint64_t* offsets, uint64_t n_reads) {
        check(tsxc_add_reads(m_h, packed, offsets, n_reads));
    }
    void sync() { check(tsxc_sync(m_h)); }

    // uint64_t getKmerCount() — TSXHashMap.h:645-648
    uint64_t getKmerCount() { uint64_t n = 0; check(tsxc_distinct(m_h, &n)); return n; }
    // UBigInt getKmerCount(kmer) — TSXHashMap.h:548-638
    uint64_t getKmerCount(const uint64_t* kmer) { uint64_t c = 0; check(tsxc_lookup(m_h, kmer, 1, &c)); return c; }
    void getKmerCounts(const uint64_t* kmers, uint64_t n, uint64_t* counts) { check(tsxc_lookup(m_h, kmers, n, counts)); }
    void dump(const std::string& path) { check(tsxc_dump_file(m_h, path.c_str())); }
    // hist[c] = distinct k-mers with count c (c < n_bins-1), hist[n_bins-1] = all with a larger count
    std::vector<uint64_t> histogram(uint32_t n_bins) { std::vector<uint64_t> h(n_bins); check(tsxc_histogram(m_h, h.data(), n_bins)); return h; }

    tsxc_stats_t stats() { tsxc_stats_t s; check(tsxc_stats(m_h, &s)); return s; }

    // void print_stats() — TSXHashMap.h:390-395
    void print_stats() {
        const tsxc_stats_t s = stats();
        std::cerr << "Used fields: " << s.used_slots << std::endl;
        std::cerr << "Available fields: " << (double)s.n_slots << std::endl;
        std::cerr << "k=" << m_iK << " l=" << (uint32_t)m_iL << " entry (key+value) bits=" << 64 * s.entry_words
                  << " storage bits=" << s.value_bits << std::endl;
    }

private:
    void check(int rc) {
        if (rc == TSXC_OK) return;
        if (rc == TSXC_E_TABLE_FULL) {                                       // TSXHashMap.h:334-343
            const tsxc_stats_t s = [&] { tsxc_stats_t x{}; tsxc_stats(m_h, &x); return x; }();
            std::cerr << "Could not insert kmer: " << tsxc_last_error(m_h) << std::endl;
            std::cerr << "Used fields: " << s.used_slots << std::endl;
            std::cerr << "Available fields: " << (double)s.n_slots << std::endl;
            std::exit(42);
        }
        throw TSXException(std::string(tsxc_status_string(rc)) + ": " + tsxc_last_error(m_h));
    }

    tsxc_table* m_h = nullptr;
    uint8_t m_iL;
    uint32_t m_iStorageBits;
    uint16_t m_iK;
    uint32_t m_kw;
};
