/*
 * tsxcount_cuda.h — C ABI of the B200 (sm_100a) k-mer counting path.
 *
 * This is the drop-in boundary for the reference's hash-map insert path: everything below
 * `pMap->addKmer(...)` in src/mains/main.cpp:192 of mjoppich/tsxCount, i.e. the TSXHashMap class
 * (src/tsxcount/TSXHashMap.h:68) with its serialization backends (TSXHashMapPerf / CAS / OMPPerf /
 * PThreadPerf / TSXPerf) plus the k-mer extraction that feeds it (src/mains/testExecution.h:15-36,
 * src/utils/SequenceUtils.h:86-160).  The reference interface is a C++ virtual class called once
 * per k-mer from OpenMP tasks; a GPU cannot be fed that way, so the boundary is batch-granular.
 * Each entry point names the reference member it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C, opaque handle, `int` status codes, no exceptions across the boundary;
 *   - a k-mer is KW = tsxc_key_words(k) little-endian 64-bit words; base i of the k-mer occupies
 *     bits [2i, 2i+1] with A=0 C=1 G=2 T=3 (SequenceUtils.h:96-123), unused high bits are zero;
 *   - packed reads: one concatenated stream of 2-bit bases in the same digit order (base g of the
 *     stream = bits [2(g%32), 2(g%32)+1] of word g/32) plus n_reads+1 base offsets; a k-mer never
 *     crosses a read boundary (testExecution.h:19-36);
 *   - `_device` variants take device pointers that are already resident on the handle's GPU;
 *     the others take host pointers (pinned memory makes the copies asynchronous);
 *   - work is queued on the handle's stream; buffers passed to an add call must stay valid until
 *     the next tsxc_sync();
 *   - there is NO CPU fallback: every entry point fails with TSXC_E_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef TSXCOUNT_CUDA_H
#define TSXCOUNT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSXC_ABI_VERSION 2
#define TSXC_MAX_K 128

/* ---- status codes ------------------------------------------------------------------------ */
enum {
    TSXC_OK = 0,
    TSXC_E_INVALID = 1,          /* TSXException("Invalid lengths for hashmap size and value of k"),
                                    TSXHashMap.h:91-94, and any other bad argument */
    TSXC_E_CUDA = 2,             /* CUDA runtime / no usable device */
    TSXC_E_NOMEM = 3,            /* table or staging allocation failed */
    TSXC_E_UNSUPPORTED = 4,      /* (k, l, s) does not fit any entry class */
    TSXC_E_COUNT_SATURATED = 5,  /* an overflow counter ran out of bits (>= 2^44 + s per k-mer) */
    TSXC_E_IO = 6,
    TSXC_E_TABLE_FULL = 42       /* reprobe limit reached; the reference does exit(42), TSXHashMap.h:340-343 */
};

/* ---- creation flags ---------------------------------------------------------------------- */
#define TSXC_FLAG_NONE 0u
/* Keep exactly `s` value bits per primary entry like the reference (counts >= 2^s spill into
 * overflow entries, TSXHashMapPerf.h:154-164,699-881).  Without it the value field is widened to
 * every spare bit of the entry word (>= s).  s == 0 always means "widest". */
#define TSXC_FLAG_EXACT_S 1u
/* Disable warp-level pre-aggregation (__match_any_sync) — for measurements only. */
#define TSXC_FLAG_NO_WARP_AGG 2u
/* Always use the single fused extract+insert kernel, never the region-sorted pipeline that large tables take by
 * default (DESIGN.md "Why a sort by table region") — for A/B measurements. */
#define TSXC_FLAG_DIRECT 4u
/* (Round 1 had a caller hint TSXC_FLAG_SKEWED for inputs dominated by a few k-mers.  The pipeline's bins now grow page by
 * page as the data demands and duplicates are combined in shared memory before they reach the table, so there is
 * nothing left to hint and the flag is gone.) */

/* Count canonical k-mers: every k-mer (from reads, tsxc_add_kmers, tsxc_lookup) is replaced by the lexicographically
 * smaller of itself and its reverse complement before it is hashed; dumps list the canonical form.  An opt-in
 * extension: the reference counts forward k-mers only (src/mains/testExecution.h:15-36,
 * no reverse-complement code anywhere in src/). */
#define TSXC_FLAG_CANONICAL 8u

typedef struct tsxc_table tsxc_table; /* opaque */

typedef struct tsxc_stats_t {
    uint32_t k, l, s;             /* as requested */
    uint32_t key_words;           /* KW */
    uint32_t entry_words;         /* 64-bit words per table entry: 1, 2 or 4 */
    uint32_t value_bits;          /* bits of the primary value field actually used */
    uint32_t quotient_bits;       /* stored key bits (2k - bucket-index bits) */
    uint32_t reprobe_bits;
    uint32_t slots_per_bucket;    /* one bucket = one 32-byte sector */
    uint32_t n_shards, shard_rank;
    uint32_t reserved0;
    uint64_t n_slots;             /* slots of THIS shard */
    uint64_t table_bytes;
    uint64_t distinct;            /* primary entries   (TSXHashMap::getKmerCount(), TSXHashMap.h:645-648) */
    uint64_t overflow_entries;    /* overflow entries  (handleOverflow, TSXHashMapPerf.h:699-881) */
    uint64_t used_slots;          /* distinct + overflow_entries  ("Used fields", TSXHashMap.h:390-395) */
    uint64_t kmers_added;         /* sum of all increments accepted */
    uint64_t max_reprobe;         /* longest probe sequence seen by an insert */
    uint64_t error_flags;         /* sticky device-side error bits */
    uint64_t kernel_launches;     /* kernels of this library launched on the handle since create/clear */
    uint64_t main_kernel_launches;/* launches of the dominant (extract+insert / insert) kernels among them */
    double   main_kernel_ms;      /* their summed device time (CUDA events on the handle's stream) */
    double   partition_ms;        /* of which: histogram + both radix partition passes of the region-sorted pipeline */
    double   insert_ms;           /* of which: phase B (insert in table-region order) */
    double   hist_ms;             /* partition_ms split: S0 (extract+hash+digit-1 histogram, chunk planner) */
    double   part1_ms;            /*                     S1 (extract+hash+tile sort by digit 1) */
    double   part2_ms;            /*                     S2 (digit-2 histogram, scan, tile sort by digit 2) */
    uint64_t chunk_cap_keys;      /* capacity of key buffer A: k-mers per insert pass (0 until the pipeline ran) */
    uint64_t group_cap_keys;      /* keys per page of the paged key buffer (0: bins have exact offsets) */
    uint32_t radix_digit1_bits, radix_digit2_bits;
} tsxc_stats_t;

/* Words per k-mer for this k: 1 (k<=32), 2 (k<=64), 4 (k<=128); 0 if k is out of range. */
uint32_t tsxc_key_words(uint32_t k);
int tsxc_abi_version(void);
/* Number of usable sm_100 devices (0 on a CPU-only host; never an error). */
int tsxc_device_count(void);
const char* tsxc_status_string(int status);
/* Message of the last failure on this handle ("" if none); handle may be NULL for create failures. */
const char* tsxc_last_error(const tsxc_table* t);

/* ---- life cycle -------------------------------------------------------------------------- */
/* TSXHashMap::TSXHashMap(uint8_t iL, uint32_t iStorageBits, uint16_t iK) — TSXHashMap.h:79-154, selected in
 * main.cpp:429-475.  2^l slots, s value bits, k-mer length k; requires 2k > l (TSXHashMap.h:91-94).
 * The table is allocated and zeroed in HBM of `device`. */
int tsxc_create(uint32_t k, uint32_t l, uint32_t s, int device, uint32_t flags, tsxc_table** out);
/* One shard of a table that is hash-partitioned over n_shards GPUs (n_shards a power of two <= 2^(l-8)).
 * The shard owns the k-mers whose bucket index has shard_rank in its top log2(n_shards) bits and
 * allocates 2^l / n_shards slots.  (No reference counterpart: the reference is single-process.) */
int tsxc_create_shard(uint32_t k, uint32_t l, uint32_t s, int device, uint32_t flags,
                      uint32_t shard_rank, uint32_t n_shards, tsxc_table** out);
/* ~TSXHashMap — TSXHashMap.h:156-160 */
int tsxc_destroy(tsxc_table* t);
/* Re-zero the table and its counters (no reference counterpart; used between benchmark passes). */
int tsxc_clear(tsxc_table* t);
/* Give back the key buffers of the insert pipeline (they take whatever HBM the table leaves free and are sized
 * again by the next batch).  Waits for queued work.  No effect while the buffers are exported to peer ranks. */
int tsxc_trim(tsxc_table* t);
/* The CUDA stream (cudaStream_t) the handle queues work on, as an opaque pointer. */
void* tsxc_stream(tsxc_table* t);

/* Timing marks: record CUDA event `idx` (0..7) on the handle's stream / elapsed device time between two
 * recorded marks after a tsxc_sync().  This is how bench.py times a step on the launching stream. */
int tsxc_mark(tsxc_table* t, int idx);
int tsxc_mark_elapsed_ms(tsxc_table* t, int idx_from, int idx_to, float* ms_out);

/* ---- insert path ------------------------------------------------------------------------- */
/* createKMers + fromSequence + addKmer for a whole batch of reads — main.cpp:159-192,
 * testExecution.h:15-36, SequenceUtils.h:86-160, TSXHashMapPerf.h:56-205 / TSXHashMapCAS.h:268-508.
 * offsets[r]..offsets[r+1] are the bases of read r in the packed stream (offsets[0] == 0).
 * The host variant copies the batch to the device at once (asynchronously from pinned memory).  For tables large
 * enough to take the region-sorted pipeline the batches are appended to a device-side stream and counted when they
 * fill an insert pass — or at tsxc_sync, or before any call that reads the table — so the caller is free to cut its
 * input into batches of any size. */
int tsxc_add_reads(tsxc_table* t, const uint64_t* packed, const uint64_t* offsets, uint64_t n_reads);
int tsxc_add_reads_device(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint64_t n_reads,
                          uint64_t n_bases);
/* Batched TSXHashMap::addKmer(UBigInt& kmer) — TSXHashMap.h:182; n k-mers of KW words each. */
int tsxc_add_kmers(tsxc_table* t, const uint64_t* kmers, uint64_t n);
int tsxc_add_kmers_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n);
/* Wait for all queued work; returns the sticky device status (TSXC_E_TABLE_FULL, ...). */
int tsxc_sync(tsxc_table* t);

/* ---- query path -------------------------------------------------------------------------- */
/* Batched TSXHashMap::getKmerCount(UBigInt& kmer) incl. findOverflowCounts — TSXHashMap.h:548-638,951-1039.
 * counts_out[i] = count of k-mer i, 0 if absent.  (A sharded handle answers 0 for k-mers it does not own.) */
int tsxc_lookup(tsxc_table* t, const uint64_t* kmers, uint64_t n, uint64_t* counts_out);
int tsxc_lookup_device(tsxc_table* t, const uint64_t* d_kmers, uint64_t n, uint64_t* d_counts_out);
/* uint64_t TSXHashMap::getKmerCount() (popcount of the k-mer-start bitmap) — TSXHashMap.h:645-648 */
int tsxc_distinct(tsxc_table* t, uint64_t* out);
/* TSXHashMap::getAllKmers() (TSXHashMap.h:660-722) extended with counts: writes up to `capacity` (k-mer, count)
 * pairs into host buffers (kmers_out: capacity*KW words) and the number found into n_out.  If n_out > capacity
 * the output is truncated and the call returns TSXC_E_INVALID.  Order is table order (unspecified). */
int tsxc_dump(tsxc_table* t, uint64_t* kmers_out, uint64_t* counts_out, uint64_t capacity, uint64_t* n_out);
/* Writes `KMER<TAB>COUNT\n` per distinct k-mer, the format of count_kmers.py:32-34. */
int tsxc_dump_file(tsxc_table* t, const char* path);
/* TSXHashMap::print_stats() numbers — TSXHashMap.h:390-395 */
int tsxc_stats(tsxc_table* t, tsxc_stats_t* out);
/* Count histogram (extension): hist_out[c] = number of distinct k-mers whose count is c for
 * c < n_bins - 1, hist_out[n_bins - 1] = number with a larger count; hist_out[0] is always 0.  2 <= n_bins <= 4096. */
int tsxc_histogram(tsxc_table* t, uint64_t* hist_out, uint32_t n_bins);

/* ---- multi-GPU routing (hash-sharded table) ------------------------------- */
/* One process per GPU, each holding one shard (tsxc_create_shard).  A batch of reads is counted in rounds; in every
 * round each rank routes one chunk of its reads and inserts what it receives.  Per round, all on the handle's stream:
 *   tsxc_route_hist    this rank's exact k-mer counts per routing bin (owner-major) -> caller's d_hist (bins x uint32)
 *   [caller]           all-gather of d_hist over the ranks (NCCL) -> d_hist_all (n_shards x bins x uint32)
 *   tsxc_route_send    every rank derives from d_hist_all where each of its bins starts in each owner's receive
 *                      buffer; the routing kernel (extract + hash + tile sort) then stores its runs STRAIGHT INTO the
 *                      owners' peer-mapped buffers over NVLink: compute and exchange are one kernel
 *   [caller]           a barrier collective on the stream (all stores have landed)
 *   tsxc_route_insert  sort the received k-mers by fine table region and insert them (same kernels as one GPU)
 *   [the next round's all-gather doubles as "every receive buffer has been drained"]
 * Sizes are exact, so nothing can overflow silently: if a round would exceed a receive buffer, every rank sees it in
 * d_hist_all, all of them skip the round, and tsxc_sync reports TSXC_E_INVALID.
 * (No reference counterpart: the reference is one process on one shared table, src/mains/main.cpp:132-218.) */
typedef struct tsxc_route_info_t {
    uint32_t n_shards, shard_rank;
    uint32_t bins;               /* routing bins = entries of d_hist (n_shards * bins_per_shard) */
    uint32_t bins_per_shard;
    uint32_t key_words;
    uint32_t reserved;
    uint64_t recv_cap_keys;      /* capacity of this rank's receive buffer in k-mers (0 before tsxc_route_recv_buffer) */
} tsxc_route_info_t;
int tsxc_route_info(tsxc_table* t, tsxc_route_info_t* out);
/* Allocate this rank's receive buffer (cap_keys k-mers; 0 = what free HBM allows) and return its device pointer so
 * that the caller can export it to the other ranks (tsxc_ipc_export_mem, or directly within one process). */
int tsxc_route_recv_buffer(tsxc_table* t, uint64_t cap_keys, void** d_ptr_out, uint64_t* cap_keys_out);
/* peer_buffers[o] = shard o's receive buffer as addressable from this device (peer_buffers[shard_rank] = own);
 * recv_cap_keys = the smallest capacity among them (every rank must pass the same value). */
int tsxc_route_set_peers(tsxc_table* t, void* const* peer_buffers, uint64_t recv_cap_keys);
/* Start a batch: read-end bitmap, histogram, chunk plan.  *rounds_out = upper bound of the rounds this rank needs
 * (the caller takes the maximum over the ranks; surplus rounds send nothing). */
int tsxc_route_begin(tsxc_table* t, const uint64_t* d_packed, const uint64_t* d_offsets, uint64_t n_reads, uint64_t n_bases,
                     uint32_t* rounds_out);
int tsxc_route_hist(tsxc_table* t, uint32_t round, uint32_t* d_hist_out);
int tsxc_route_send(tsxc_table* t, uint32_t round, const uint32_t* d_hist_all);
int tsxc_route_insert(tsxc_table* t);

/* ---- peer-memory plumbing for the multi-GPU exchange (one process per GPU) ------------------------ */
/* The owner exports its receive buffer (CUDA IPC), the other ranks open it and pass the mapped pointers to
 * tsxc_route_set_peers.  All handles are 64 opaque bytes (cudaIpcMemHandle_t / cudaIpcEventHandle_t). */
/* Within ONE process that drives several devices (the C++ CLI's --gpus=N) no IPC is needed: enable peer access from
 * `device` to `peer` (cudaDeviceEnablePeerAccess) and pass the other shards' buffer pointers as they are. */
int tsxc_enable_peer_access(int device, int peer);
#define TSXC_IPC_HANDLE_BYTES 64
int tsxc_ipc_export_mem(int device, void* dptr, unsigned char* handle_out);
int tsxc_ipc_open_mem(int device, const unsigned char* handle, void** dptr_out);
int tsxc_ipc_close_mem(int device, void* dptr);
int tsxc_ipc_event_create(int device, void** event_out, unsigned char* handle_out);
int tsxc_ipc_event_open(int device, const unsigned char* handle, void** event_out);
int tsxc_event_destroy(int device, void* event);
int tsxc_event_record(int device, void* event, void* stream);
int tsxc_stream_wait_event(int device, void* stream, void* event);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream): dst may be a peer pointer from tsxc_ipc_open_mem. */
int tsxc_copy_async(int device, void* dst, const void* src, uint64_t bytes, void* stream);

/* ---- host-side packing (replaces TSXSeqUtils::fromSequence on the feeder side) ----------- */
/* 2-bit packs ASCII reads.  Non-ACGT bytes cannot be packed: each maximal ACGT run becomes its own
 * segment, which is exactly "skip every k-mer that spans a non-ACGT base" (the reference substitutes
 * random bits there, SequenceUtils.h:126-137 — not reproducible; see DESIGN.md "N policy").
 * packed_out must hold ceil(total_bases/32) words, seg_offsets_out up to n_reads + n_bad_bases + 1
 * entries (pass capacity in seg_capacity).  Returns the number of segments in n_segments_out. */
int tsxc_pack_reads(const char* ascii, const uint64_t* offsets, uint64_t n_reads, uint64_t* packed_out,
                    uint64_t* seg_offsets_out, uint64_t seg_capacity, uint64_t* n_segments_out,
                    uint64_t* n_bad_bases_out);

/* ---- synthetic inputs (DESIGN.md "Synthetic inputs"; style of generateFakeSequences.py:7-18) */
typedef struct tsxc_gen_params {
    uint64_t seed;
    uint64_t n_reads;
    uint32_t read_len;
    uint32_t mode;         /* 0 uniform, 1 fakeseq (poly-A tail), 2 log-uniform dictionary, 3 genome sample */
    uint64_t genome_len;   /* mode 3: genome bases; mode 2: dictionary entries (power of two) */
    uint32_t sub_rate_q16; /* per-base substitution probability, 1/65536 units */
    uint32_t reserved;
} tsxc_gen_params;
/* Fill d_packed (ceil(count*read_len/32) words) and d_offsets (count+1) on `device` with reads
 * [first, first+count) of the generator, on the given stream (cudaStream_t or NULL). */
int tsxc_gen_reads_device(const tsxc_gen_params* p, uint64_t first, uint64_t count, int device, void* stream,
                          uint64_t* d_packed, uint64_t* d_offsets);

/* ---- memory helpers for hosts without a CUDA runtime of their own ----------------------------- */
/* Page-locked host memory (cudaHostAlloc) so tsxc_add_reads copies asynchronously. */
int tsxc_host_alloc(uint64_t bytes, void** out);
int tsxc_host_free(void* p);
int tsxc_device_alloc(int device, uint64_t bytes, void** out);
int tsxc_device_free(int device, void* p);
/* Synchronous copy; kind: 1 host->device, 2 device->host, 3 device->device. */
int tsxc_memcpy(int device, void* dst, const void* src, uint64_t bytes, int kind);

/* ---- K0: random-access roofline microbenchmark --------------------------- */
/* n_ops random 8-byte RMWs (mode 0: atomicAdd without return; 1: atomicCAS; 2: 32-byte sector load +
 * atomicAdd on one of its words) on a scratch region of table_bytes in the handle's table memory
 * (contents are destroyed — call tsxc_clear afterwards).  Returns elapsed device milliseconds. */
int tsxc_k0_random_rmw(tsxc_table* t, uint64_t table_bytes, uint64_t n_ops, int mode, float* ms_out);

/* K0w: the same with every thread block confined to its own window of the footprint (disjoint windows while
 * blocks*window <= footprint); separates address-translation reach from the DRAM random-access rate.
 * mode 0 atomicAdd, 2 sector load + atomicAdd, 3 sector load only. */
int tsxc_k0_windowed(tsxc_table* t, uint64_t footprint_bytes, uint64_t window_bytes, uint64_t n_ops, int mode,
                     int blocks, int threads, float* ms_out);

/* K0r: all blocks sweep the footprint region by region, ops_per_region random RMWs each (L2-blocked variant).
 * mode (low byte): 0 RED, 1 sector load, 2 sector load + RED, 3 sector load + CAS (result consumed late),
 * 4 returning atomic, 5 sector load + CAS + branch on the result (the thread waits: the insert's claim).
 * mode >> 8: resident 256-thread blocks per SM, 1..8 (0 = 8). */
int tsxc_k0_region_sweep(tsxc_table* t, uint64_t footprint_bytes, uint64_t region_bytes, uint64_t ops_per_region,
                         uint32_t ops_per_item, int mode, float* ms_out);

/* ---- debugging / tests ------------------------------------------------------------------- */
/* The bijective hash and its inverse on the host (round-trip property of TSXHashMap::testHashFunction,
 * TSXHashMap.h:724-735).  key/out are KW words. */
int tsxc_debug_hash(uint32_t k, const uint64_t* key, uint64_t* out);
int tsxc_debug_unhash(uint32_t k, const uint64_t* hash, uint64_t* out);
/* min(k-mer, reverse complement) in the library's encoding (what TSXC_FLAG_CANONICAL tables count); host-side. */
int tsxc_debug_canonical(uint32_t k, const uint64_t* key, uint64_t* out);
/* One block round of the partition pass's walk over valid k-mer starts, on the host through the functions the kernel
 * calls: the k-mers (KW words each) that start in stream words [round, round + 512) and before word w_end of a packed
 * stream with its read-end bitmap (bit g set: base g is the last base of a read).  keys_out: room for 16384 k-mers. */
int tsxc_debug_sparse_round(uint32_t k, const uint64_t* packed, const uint32_t* ends, uint64_t n_words, uint64_t n_bases,
                            uint64_t round, uint64_t w_end, uint64_t* keys_out, uint32_t* n_out);
/* Entry layout chosen for (k, l, s, flags, n_shards) without touching a GPU. */
int tsxc_debug_layout(uint32_t k, uint32_t l, uint32_t s, uint32_t flags, uint32_t n_shards, tsxc_stats_t* out);

#ifdef __cplusplus
}
#endif
#endif /* TSXCOUNT_CUDA_H */
