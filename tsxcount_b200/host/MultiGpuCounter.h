// MultiGpuCounter.h — hash-sharded counting over the GPUs of one box from ONE host process (the C++ CLI's --gpus=N).
//
// Same round protocol as tsxcount_b200/multigpu.py (one process per GPU, torch.distributed), with NCCL called
// directly: every GPU holds one shard (tsxc_create_shard); a super-batch gives every GPU one batch of reads; per round
//   tsxc_route_hist   ->  ncclAllGather of the per-bin k-mer counts (ncclGroupStart/End over the devices)
//   tsxc_route_send   ->  the routing kernel stores its runs straight into the owners' receive buffers (peer access)
//   ncclAllReduce     ->  barrier on the streams: all stores have landed
//   tsxc_route_insert ->  every shard sorts what it received by table region and inserts it
// The reference has no counterpart (one process, one shared table: src/mains/main.cpp:132-218 of mjoppich/tsxCount).
#pragma once

#include <nccl.h>

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "tsxcount_cuda.h"

class MultiGpuCounter {
public:
    struct HostBatch { const uint64_t* packed = nullptr; const uint64_t* offsets = nullptr; uint64_t n_reads = 0; };

    MultiGpuCounter(uint32_t k, uint32_t l, uint32_t s, int n_gpus, uint32_t flags, uint64_t recv_cap_keys = 0)
        : m_n(n_gpus), m_kw(tsxc_key_words(k)) {
        if (n_gpus < 2 || (n_gpus & (n_gpus - 1))) throw std::runtime_error("--gpus must be a power of two >= 2 (the table is sharded by hash bits)");
        if (tsxc_device_count() < n_gpus) throw std::runtime_error("fewer sm_100 devices than --gpus");
        m_shard.resize(n_gpus, nullptr); m_comm.resize(n_gpus); m_dev.resize(n_gpus);
        m_packed.resize(n_gpus, nullptr); m_off.resize(n_gpus, nullptr); m_cap_words.resize(n_gpus, 0); m_cap_off.resize(n_gpus, 0);
        m_hist.resize(n_gpus, nullptr); m_hist_all.resize(n_gpus, nullptr); m_flag.resize(n_gpus, nullptr);
        std::vector<int> devs(n_gpus);
        for (int d = 0; d < n_gpus; ++d) { devs[d] = d; m_dev[d] = d; }
        for (int d = 0; d < n_gpus; ++d) {
            check(tsxc_create_shard(k, l, s, d, flags, (uint32_t)d, (uint32_t)n_gpus, &m_shard[d]), nullptr);
            for (int p = 0; p < n_gpus; ++p) if (p != d) check(tsxc_enable_peer_access(d, p), nullptr);
        }
        if (ncclCommInitAll(m_comm.data(), n_gpus, devs.data()) != ncclSuccess) throw std::runtime_error("ncclCommInitAll failed");
        // receive buffers: every device can address every other one's (peer access within one process)
        std::vector<void*> bufs(n_gpus);
        uint64_t cap = ~0ULL;
        for (int d = 0; d < n_gpus; ++d) {
            uint64_t c = 0;
            check(tsxc_route_recv_buffer(m_shard[d], recv_cap_keys, &bufs[d], &c), m_shard[d]);
            cap = std::min(cap, c);
        }
        tsxc_route_info_t info;
        check(tsxc_route_info(m_shard[0], &info), m_shard[0]);
        m_bins = info.bins;
        for (int d = 0; d < n_gpus; ++d) {
            check(tsxc_route_set_peers(m_shard[d], bufs.data(), cap), m_shard[d]);
            void* p;
            check(tsxc_device_alloc(d, m_bins * sizeof(uint32_t), &p), nullptr); m_hist[d] = (uint32_t*)p;
            check(tsxc_device_alloc(d, (uint64_t)m_bins * n_gpus * sizeof(uint32_t), &p), nullptr); m_hist_all[d] = (uint32_t*)p;
            check(tsxc_device_alloc(d, 64, &p), nullptr); m_flag[d] = (uint32_t*)p;
        }
        m_recv_cap = cap;
    }
    ~MultiGpuCounter() {
        for (int d = 0; d < m_n; ++d) {
            if (m_shard[d]) tsxc_sync(m_shard[d]);
            tsxc_device_free(d, m_packed[d]); tsxc_device_free(d, m_off[d]);
            tsxc_device_free(d, m_hist[d]); tsxc_device_free(d, m_hist_all[d]); tsxc_device_free(d, m_flag[d]);
        }
        for (auto& c : m_comm) ncclCommDestroy(c);
        for (auto* t : m_shard) tsxc_destroy(t);
    }
    MultiGpuCounter(const MultiGpuCounter&) = delete;
    MultiGpuCounter& operator=(const MultiGpuCounter&) = delete;

    int gpus() const { return m_n; }
    uint32_t keyWords() const { return m_kw; }
    uint64_t recvCapKeys() const { return m_recv_cap; }

    // One batch per GPU (pinned host memory; empty batches allowed).  Returns after the work is queued; the host
    // buffers may be reused after sync().
    void addSuperBatch(const std::vector<HostBatch>& b) {
        uint32_t rounds = 0;
        for (int d = 0; d < m_n; ++d) {
            const uint64_t n_bases = b[d].n_reads ? b[d].offsets[b[d].n_reads] : 0;
            const uint64_t n_words = (n_bases + 31) / 32;
            void* st = tsxc_stream(m_shard[d]);
            grow(d, n_words + 8, b[d].n_reads + 1);
            if (n_words) check(tsxc_copy_async(d, m_packed[d], b[d].packed, n_words * 8, st), nullptr);
            if (b[d].n_reads) check(tsxc_copy_async(d, m_off[d], b[d].offsets, (b[d].n_reads + 1) * 8, st), nullptr);
            uint32_t r = 0;
            check(tsxc_route_begin(m_shard[d], m_packed[d], m_off[d], b[d].n_reads, n_bases, &r), m_shard[d]);
            rounds = std::max(rounds, r);
        }
        for (uint32_t r = 0; r < rounds; ++r) {
            for (int d = 0; d < m_n; ++d) check(tsxc_route_hist(m_shard[d], r, m_hist[d]), m_shard[d]);
            nccl(ncclGroupStart());
            for (int d = 0; d < m_n; ++d)
                nccl(ncclAllGather(m_hist[d], m_hist_all[d], m_bins, ncclUint32, m_comm[d], (cudaStream_t)tsxc_stream(m_shard[d])));
            nccl(ncclGroupEnd());
            for (int d = 0; d < m_n; ++d) check(tsxc_route_send(m_shard[d], r, m_hist_all[d]), m_shard[d]);
            nccl(ncclGroupStart());                                  // barrier on the streams: every rank's stores have landed
            for (int d = 0; d < m_n; ++d)
                nccl(ncclAllReduce(m_flag[d], m_flag[d], 1, ncclUint32, ncclSum, m_comm[d], (cudaStream_t)tsxc_stream(m_shard[d])));
            nccl(ncclGroupEnd());
            for (int d = 0; d < m_n; ++d) check(tsxc_route_insert(m_shard[d]), m_shard[d]);
        }
    }
    void sync() { for (int d = 0; d < m_n; ++d) check(tsxc_sync(m_shard[d]), m_shard[d]); }

    // distinct counts add up; a shard answers 0 for k-mers it does not own; dumps concatenate
    uint64_t getKmerCount() { uint64_t n = 0; for (auto* t : m_shard) { uint64_t x = 0; check(tsxc_distinct(t, &x), t); n += x; } return n; }
    void getKmerCounts(const uint64_t* kmers, uint64_t n, uint64_t* counts) {
        std::vector<uint64_t> part(n);
        std::fill(counts, counts + n, 0);
        for (auto* t : m_shard) {
            check(tsxc_lookup(t, kmers, n, part.data()), t);
            for (uint64_t i = 0; i < n; ++i) counts[i] += part[i];
        }
    }
    std::vector<uint64_t> histogram(uint32_t n_bins) {          // shards hold disjoint k-mers: histograms add up
        std::vector<uint64_t> h(n_bins, 0), part(n_bins);
        for (auto* t : m_shard) {
            check(tsxc_histogram(t, part.data(), n_bins), t);
            for (uint32_t i = 0; i < n_bins; ++i) h[i] += part[i];
        }
        return h;
    }
    void dump(const std::string& path) {
        for (int d = 0; d < m_n; ++d) {
            const std::string part = path + ".shard" + std::to_string(d);
            check(tsxc_dump_file(m_shard[d], part.c_str()), m_shard[d]);
        }
        FILE* out = std::fopen(path.c_str(), "wb");
        if (!out) throw std::runtime_error("cannot open " + path);
        std::vector<char> buf(1 << 22);
        for (int d = 0; d < m_n; ++d) {
            const std::string part = path + ".shard" + std::to_string(d);
            FILE* in = std::fopen(part.c_str(), "rb");
            if (!in) { std::fclose(out); throw std::runtime_error("cannot open " + part); }
            size_t got;
            while ((got = std::fread(buf.data(), 1, buf.size(), in)) > 0) std::fwrite(buf.data(), 1, got, out);
            std::fclose(in);
            std::remove(part.c_str());
        }
        std::fclose(out);
    }
    tsxc_stats_t stats() {      // sums over the shards (sizes, counters); times of shard 0
        tsxc_stats_t s{};
        for (int d = 0; d < m_n; ++d) {
            tsxc_stats_t x;
            check(tsxc_stats(m_shard[d], &x), m_shard[d]);
            if (d == 0) s = x;
            else {
                s.n_slots += x.n_slots; s.table_bytes += x.table_bytes; s.distinct += x.distinct; s.overflow_entries += x.overflow_entries;
                s.used_slots += x.used_slots; s.kmers_added += x.kmers_added; s.max_reprobe = std::max(s.max_reprobe, x.max_reprobe);
                s.error_flags |= x.error_flags; s.kernel_launches += x.kernel_launches;
            }
        }
        return s;
    }

private:
    void grow(int d, uint64_t words, uint64_t offs) {
        void* p;
        if (words > m_cap_words[d]) {
            check(tsxc_sync(m_shard[d]), m_shard[d]);
            tsxc_device_free(d, m_packed[d]);
            check(tsxc_device_alloc(d, words * 8, &p), nullptr);
            m_packed[d] = (uint64_t*)p; m_cap_words[d] = words;
        }
        if (offs > m_cap_off[d]) {
            check(tsxc_sync(m_shard[d]), m_shard[d]);
            tsxc_device_free(d, m_off[d]);
            check(tsxc_device_alloc(d, offs * 8, &p), nullptr);
            m_off[d] = (uint64_t*)p; m_cap_off[d] = offs;
        }
    }
    static void check(int rc, tsxc_table* t) {
        if (rc == TSXC_OK) return;
        if (rc == TSXC_E_TABLE_FULL) { std::fprintf(stderr, "Could not insert kmer: %s\n", tsxc_last_error(t)); std::exit(42); }   // TSXHashMap.h:340-343
        throw std::runtime_error(std::string(tsxc_status_string(rc)) + ": " + tsxc_last_error(t));
    }
    static void nccl(ncclResult_t r) { if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL: ") + ncclGetErrorString(r)); }

    int m_n;
    uint32_t m_kw, m_bins = 0;
    uint64_t m_recv_cap = 0;
    std::vector<tsxc_table*> m_shard;
    std::vector<ncclComm_t> m_comm;
    std::vector<int> m_dev;
    std::vector<uint64_t*> m_packed, m_off;
    std::vector<uint64_t> m_cap_words, m_cap_off;
    std::vector<uint32_t*> m_hist, m_hist_all, m_flag;
};
