/*
 * oracle.h — CPU restatement of the tsxCount counting path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker the CUDA path is compared against.  It is never linked into,
 * imported by or executed from the product (tsxcount_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * Parity pin: oracle_count_fastq() reproduces /root/reference/data/small_t7.1000.fastq.14.count
 * byte for byte (tests/test_oracle.py, fixture copied to tests/golden/), and agrees with the
 * unmodified reference binary oracle/_ref/tsxCount (--check, "total errors0") on seeded
 * synthetic inputs at k in {14,31,63} (tests/golden/ref_binary_pins.json, made by
 * oracle/make_ref_pins.py).  k > 64 is UNPINNED by the reference (its binary cannot run
 * there); the restatement follows the semantics proven at smaller k.
 *
 * What is restated (reference file:line, relative to /root/reference):
 *   record parsing   src/fastxutils/FastXReader.h:62-95,307-385   4-line records, empty lines skipped
 *   extraction       src/mains/testExecution.h:15-36              every forward substring seq[i:i+k]
 *   encoding         src/utils/SequenceUtils.h:86-160             A=0 C=1 G=2 T=3, base i -> bits [2i,2i+1]
 *   decoding         src/utils/SequenceUtils.h:47-84
 *   dump format      count_kmers.py:32-34                         KMER<TAB>COUNT, first-occurrence order
 *   check semantics  src/mains/main.cpp:224-396, testExecution.h:38-50
 * The table internals (TSXHashMap*.h) are NOT restated: table bytes are unobservable
 * (time-seeded hash, BijectiveKMapping.h:84,284-303); the contract is the (k-mer -> count) map.
 */
#ifndef TSX_ORACLE_H
#define TSX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_K 128
#define ORC_KEY_WORDS 4 /* every key is held in 4 x u64, little-endian word order, zero padded */

/* Non-ACGT policy.  The reference substitutes rand()%2 bits (SequenceUtils.h:126-137), which is
 * not reproducible; every parity input is therefore N-free.  The restatement offers the
 * deterministic policy the product implements: k-mers that span a non-ACGT character are skipped. */
#define ORC_N_SKIP 0

typedef struct orc_counts {
    uint64_t  n_distinct;
    uint64_t  n_total;     /* sum of counts = sum over reads of max(0, len-k+1) minus skipped */
    uint64_t  n_skipped;   /* k-mers dropped by the N policy */
    unsigned  k;
    uint64_t* keys;        /* n_distinct * ORC_KEY_WORDS, sorted by first occurrence */
    uint64_t* counts;      /* n_distinct */
    uint64_t* first;       /* n_distinct: index (in stream order) of the first occurrence */
} orc_counts;

/* 2-bit code of one base; -1 for anything that is not upper-case A/C/G/T (SequenceUtils.h:96-137). */
int orc_base_code(char c);

/* Encode seq[0:k] (must be all ACGT) into key[ORC_KEY_WORDS] (SequenceUtils.h:86-123). */
int orc_encode_kmer(const char* seq, unsigned k, uint64_t* key);

/* Inverse (SequenceUtils.h:47-84).  out must hold k+1 bytes. */
void orc_decode_kmer(const uint64_t* key, unsigned k, char* out);

/* Count all forward k-mers of n_reads reads given as one concatenated ASCII buffer with
 * n_reads+1 offsets (testExecution.h:15-36 applied per read as main.cpp:168-192 does).
 * Returns NULL on allocation failure or invalid k. */
orc_counts* orc_count_reads(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k);
/* The same with every k-mer replaced by the lexicographically smaller of itself and its reverse complement
 * (extension; the reference has no such mode). */
orc_counts* orc_count_reads_canonical(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k);

/* Parse a FASTQ file the way FASTXreader<FASTQEntry> does and count. */
orc_counts* orc_count_fastq(const char* path, unsigned k);

/* Read the sequences of a FASTQ file into one concatenated buffer (malloc'd) + offsets (malloc'd). */
int orc_read_fastq(const char* path, char** bases, uint64_t** offsets, uint64_t* n_reads);

/* Write KMER\tCOUNT\n lines in first-occurrence order (count_kmers.py:32-34). */
int orc_write_dump(const orc_counts* c, const char* path);

/* Binary search-free lookup helper: returns the count of key (0 when absent). O(log n) after an
 * internal sorted index is built on first use. */
uint64_t orc_lookup(orc_counts* c, const uint64_t* key);

void orc_free(orc_counts* c);

/* ---- synthetic read generators (DESIGN.md "Synthetic inputs"); independent restatement of the
 * product's device generator so both sides can be compared bit for bit. ---- */
typedef struct orc_gen_params {
    uint64_t seed;
    uint64_t n_reads;
    uint32_t read_len;
    uint32_t mode;        /* 0 uniform, 1 fakeseq (poly-A tail), 2 zipf dictionary, 3 genome sample */
    uint64_t genome_len;  /* mode 3: bases in the virtual genome; mode 2: dictionary entries (power of two) */
    uint32_t sub_rate_q16;/* substitution probability in 1/65536 units (modes 2,3) */
    uint32_t reserved;
} orc_gen_params;

/* Generates reads [first, first+count) as ASCII into out (count*read_len bytes, no separators). */
void orc_gen_reads(const orc_gen_params* p, uint64_t first, uint64_t count, char* out);

uint64_t orc_mix64(uint64_t x);

#ifdef __cplusplus
}
#endif
#endif
