/*
 * oracle.c — CPU restatement of the tsxCount counting path.  TEST INFRASTRUCTURE ONLY.
 * See oracle.h for the scope, the parity pins and the reference file:line map.
 * Plain C11, no dependency on the product.
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Encoding.  Reference: src/utils/SequenceUtils.h:86-123 — 'A' -> 00, 'C' -> 01, 'G' -> 10,
 * 'T' -> 11 where bit 2i is the LOW bit of the code and bit 2i+1 the high bit, i.e. base i is
 * the 2-bit little-endian digit i of the integer.  Only upper-case letters are recognised; every
 * other byte takes the "random" branch (:126-137), which the restatement reports as -1.
 * ------------------------------------------------------------------------------------------ */
int orc_base_code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default:  return -1;
    }
}

int orc_encode_kmer(const char* seq, unsigned k, uint64_t* key) {
    if (k == 0 || k > ORC_MAX_K) return -1;
    for (int w = 0; w < ORC_KEY_WORDS; ++w) key[w] = 0;
    for (unsigned i = 0; i < k; ++i) {
        int c = orc_base_code(seq[i]);
        if (c < 0) return -1;
        key[(2 * i) >> 6] |= (uint64_t)c << ((2 * i) & 63);
    }
    return 0;
}

/* Reference: src/utils/SequenceUtils.h:47-84 — peel two bits at a time from the low end. */
void orc_decode_kmer(const uint64_t* key, unsigned k, char* out) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (unsigned i = 0; i < k; ++i) out[i] = L[(key[(2 * i) >> 6] >> ((2 * i) & 63)) & 3];
    out[k] = 0;
}

/* ------------------------------------------------------------------------------------------
 * Counting.  Reference: src/mains/testExecution.h:15-36 (createKMers: every substring
 * seq[i:i+k], i = 0..len-k; nothing when len < k) driven per read by src/mains/main.cpp:168-192;
 * each k-mer is then addKmer'ed, so the observable result is the multiset of forward k-mers.
 * ------------------------------------------------------------------------------------------ */
typedef struct rec { uint64_t key[ORC_KEY_WORDS]; uint64_t idx; } rec;

static int cmp_key(const uint64_t* a, const uint64_t* b) {
    for (int w = ORC_KEY_WORDS - 1; w >= 0; --w) {
        if (a[w] < b[w]) return -1;
        if (a[w] > b[w]) return 1;
    }
    return 0;
}
static int cmp_rec(const void* pa, const void* pb) {
    const rec* a = (const rec*)pa; const rec* b = (const rec*)pb;
    int c = cmp_key(a->key, b->key);
    if (c) return c;
    return (a->idx > b->idx) - (a->idx < b->idx);
}
typedef struct ent { uint64_t key[ORC_KEY_WORDS]; uint64_t count; uint64_t first; } ent;
static int cmp_ent_first(const void* pa, const void* pb) {
    const ent* a = (const ent*)pa; const ent* b = (const ent*)pb;
    return (a->first > b->first) - (a->first < b->first);
}

/* Canonical form of the k-mer starting at s (all ACGT): the lexicographically smaller of the k-mer and its reverse
 * complement, compared as TEXT.  An extension: the reference counts forward k-mers only and has no
 * reverse-complement code; this is the convention of jellyfish -C and KMC.  Deliberately string-level, so that it shares
 * nothing with the bit tricks of the CUDA path. */
static void canonical_text(const char* s, unsigned k, char* out) {
    char rc[ORC_MAX_K];
    for (unsigned i = 0; i < k; ++i) {
        const char c = s[k - 1 - i];
        rc[i] = c == 'A' ? 'T' : (c == 'C' ? 'G' : (c == 'G' ? 'C' : 'A'));
    }
    memcpy(out, memcmp(s, rc, k) <= 0 ? s : rc, k);
}

static orc_counts* count_reads_impl(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k, int canonical);

orc_counts* orc_count_reads(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k) {
    return count_reads_impl(bases, offsets, n_reads, k, 0);
}
orc_counts* orc_count_reads_canonical(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k) {
    return count_reads_impl(bases, offsets, n_reads, k, 1);
}

static orc_counts* count_reads_impl(const char* bases, const uint64_t* offsets, uint64_t n_reads, unsigned k, int canonical) {
    if (k == 0 || k > ORC_MAX_K) return NULL;
    uint64_t cap = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t len = offsets[r + 1] - offsets[r];
        if (len >= k) cap += len - k + 1;
    }
    rec* recs = (rec*)malloc((cap ? cap : 1) * sizeof(rec));
    orc_counts* out = (orc_counts*)calloc(1, sizeof(orc_counts));
    if (!recs || !out) { free(recs); free(out); return NULL; }
    out->k = k;

    uint64_t n = 0, stream_idx = 0, skipped = 0;
    const unsigned nbits = 2 * k;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const char* s = bases + offsets[r];
        uint64_t len = offsets[r + 1] - offsets[r];
        if (len < k) continue;                       /* testExecution.h:19-20 */
        /* rolling window: shift right by one base, insert the new base at digit k-1 */
        uint64_t win[ORC_KEY_WORDS] = {0, 0, 0, 0};
        uint64_t good = 0;                           /* consecutive ACGT bases ending here */
        for (uint64_t i = 0; i < len; ++i) {
            int c = orc_base_code(s[i]);
            for (int w = 0; w < ORC_KEY_WORDS - 1; ++w) win[w] = (win[w] >> 2) | (win[w + 1] << 62);
            win[ORC_KEY_WORDS - 1] >>= 2;
            if (c < 0) { good = 0; c = 0; } else { good++; }
            win[(nbits - 2) >> 6] |= (uint64_t)c << ((nbits - 2) & 63);
            if (i + 1 >= k) {
                if (good >= k) {
                    if (canonical) {
                        char text[ORC_MAX_K];
                        canonical_text(s + i + 1 - k, k, text);
                        orc_encode_kmer(text, k, recs[n].key);
                    } else {
                        memcpy(recs[n].key, win, sizeof win);
                    }
                    recs[n].idx = stream_idx;
                    ++n;
                } else {
                    ++skipped;                       /* ORC_N_SKIP policy (reference: random bits) */
                }
                ++stream_idx;
            }
        }
    }
    qsort(recs, n, sizeof(rec), cmp_rec);

    uint64_t nd = 0;
    for (uint64_t i = 0; i < n; ++i) if (i == 0 || cmp_key(recs[i].key, recs[i - 1].key) != 0) ++nd;
    ent* ents = (ent*)malloc((nd ? nd : 1) * sizeof(ent));
    if (!ents) { free(recs); free(out); return NULL; }
    uint64_t j = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t e = i + 1;
        while (e < n && cmp_key(recs[e].key, recs[i].key) == 0) ++e;
        memcpy(ents[j].key, recs[i].key, sizeof ents[j].key);
        ents[j].count = e - i;
        ents[j].first = recs[i].idx;                 /* ties sorted by idx, so this is the minimum */
        ++j; i = e;
    }
    free(recs);
    /* count_kmers.py:28-34 iterates a Counter: insertion (= first occurrence) order. */
    qsort(ents, nd, sizeof(ent), cmp_ent_first);

    out->n_distinct = nd; out->n_total = n; out->n_skipped = skipped;
    out->keys = (uint64_t*)malloc((nd ? nd : 1) * ORC_KEY_WORDS * sizeof(uint64_t));
    out->counts = (uint64_t*)malloc((nd ? nd : 1) * sizeof(uint64_t));
    out->first = (uint64_t*)malloc((nd ? nd : 1) * sizeof(uint64_t));
    if (!out->keys || !out->counts || !out->first) { free(ents); orc_free(out); return NULL; }
    for (uint64_t i = 0; i < nd; ++i) {
        memcpy(out->keys + i * ORC_KEY_WORDS, ents[i].key, sizeof ents[i].key);
        out->counts[i] = ents[i].count;
        out->first[i] = ents[i].first;
    }
    free(ents);
    return out;
}

/* ------------------------------------------------------------------------------------------
 * FASTQ records.  Reference: src/fastxutils/FastXReader.h:307-385 (plain-file branch :363-372:
 * getline, lines of length 0 are dropped), :221-257 (groups of getLinesRequired()=4 lines,
 * :95), :62-78 (sequence = 2nd line of the group).  A trailing incomplete group is read by the
 * reference with a warning (:233-238) and only whole groups are turned into entries (:242).
 * ------------------------------------------------------------------------------------------ */
int orc_read_fastq(const char* path, char** bases_out, uint64_t** offsets_out, uint64_t* n_reads_out) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    size_t bcap = 1 << 20, blen = 0, ocap = 1024, nreads = 0;
    char* bases = (char*)malloc(bcap);
    uint64_t* offs = (uint64_t*)malloc(ocap * sizeof(uint64_t));
    char* line = NULL; size_t lcap = 0; ssize_t got;
    if (!bases || !offs) { fclose(f); free(bases); free(offs); return -1; }
    offs[0] = 0;
    unsigned in_group = 0;
    char* pending = NULL; size_t pending_len = 0, pending_cap = 0;
    while ((got = getline(&line, &lcap, f)) >= 0) {
        size_t len = (size_t)got;
        if (len && line[len - 1] == '\n') --len;     /* std::getline strips '\n' only (not '\r') */
        if (len == 0) continue;                      /* FastXReader.h:367-368 */
        if (in_group == 1) {                         /* second line of the record = sequence */
            if (len + 1 > pending_cap) { pending_cap = 2 * (len + 1); pending = (char*)realloc(pending, pending_cap); }
            memcpy(pending, line, len); pending_len = len;
        }
        if (++in_group == 4) {                       /* whole record available */
            if (blen + pending_len + 1 > bcap) { while (blen + pending_len + 1 > bcap) bcap *= 2; bases = (char*)realloc(bases, bcap); }
            memcpy(bases + blen, pending, pending_len); blen += pending_len;
            if (nreads + 2 > ocap) { ocap *= 2; offs = (uint64_t*)realloc(offs, ocap * sizeof(uint64_t)); }
            offs[++nreads] = blen;
            in_group = 0;
        }
    }
    free(line); free(pending); fclose(f);
    *bases_out = bases; *offsets_out = offs; *n_reads_out = nreads;
    return 0;
}

orc_counts* orc_count_fastq(const char* path, unsigned k) {
    char* bases; uint64_t* offs; uint64_t n;
    if (orc_read_fastq(path, &bases, &offs, &n) != 0) return NULL;
    orc_counts* c = orc_count_reads(bases, offs, n, k);
    free(bases); free(offs);
    return c;
}

/* count_kmers.py:32-34: fout.write(str(x) + "\t" + str(kmerCounts[x]) + "\n") */
int orc_write_dump(const orc_counts* c, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    char buf[ORC_MAX_K + 1];
    for (uint64_t i = 0; i < c->n_distinct; ++i) {
        orc_decode_kmer(c->keys + i * ORC_KEY_WORDS, c->k, buf);
        fprintf(f, "%s\t%llu\n", buf, (unsigned long long)c->counts[i]);
    }
    return fclose(f);
}

uint64_t orc_lookup(orc_counts* c, const uint64_t* key) {
    /* linear probe over the (small) oracle result; tests use this for absent-key checks only */
    for (uint64_t i = 0; i < c->n_distinct; ++i)
        if (cmp_key(c->keys + i * ORC_KEY_WORDS, key) == 0) return c->counts[i];
    return 0;
}

void orc_free(orc_counts* c) {
    if (!c) return;
    free(c->keys); free(c->counts); free(c->first); free(c);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic reads (no reference counterpart beyond the style of generateFakeSequences.py:7-18:
 * random ACGT body + poly-A tail, quality '&').  Specified in DESIGN.md "Synthetic inputs";
 * integer-only so host and device agree bit for bit.
 * ------------------------------------------------------------------------------------------ */
uint64_t orc_mix64(uint64_t x) {            /* splitmix64 finaliser */
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

/* base g of the virtual genome identified by stream id sid */
static unsigned genome_base(uint64_t seed, uint64_t sid, uint64_t g) {
    uint64_t w = orc_mix64(orc_mix64(seed ^ (sid * 0xD6E8FEB86659FD93ULL)) + (g >> 5));
    return (unsigned)(w >> (2 * (g & 31))) & 3;
}

void orc_gen_reads(const orc_gen_params* p, uint64_t first, uint64_t count, char* out) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    const uint32_t len = p->read_len;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t r = first + i;
        const uint64_t hr = orc_mix64(p->seed ^ orc_mix64(r + 0x1234567ULL));
        uint64_t start;           /* position of the read in the virtual genome (stream 1) */
        uint32_t body = len, tail = 0;
        switch (p->mode) {
            case 1:               /* fakeseq: body U[len/2, len/2+len/3], poly-A tail U[len/6, len/3] */
                body = len / 2 + (uint32_t)((hr >> 8) % (len / 3 + 1));
                tail = len / 6 + (uint32_t)((hr >> 40) % (len / 6 + 1));
                start = r * (uint64_t)len;
                break;
            case 2: {             /* log-uniform (Zipf exponent 1) rank over a power-of-two dictionary */
                unsigned lg = 0; while ((1ULL << lg) < p->genome_len) ++lg;
                unsigned b = (unsigned)((hr >> 48) % (lg + 1));
                uint64_t rank = (hr & ((1ULL << lg) - 1)) >> (lg - b);
                start = rank * (uint64_t)len;
                break;
            }
            case 3:               /* uniform sample of a genome of genome_len bases */
                start = hr % (p->genome_len - len + 1);
                break;
            default:              /* 0: independent uniform reads */
                start = r * (uint64_t)len;
                break;
        }
        for (uint32_t q = 0; q < len; ++q) {
            unsigned b = genome_base(p->seed, 1, start + q);
            if (p->mode == 1 && q >= body && q < body + tail) b = 0;   /* poly-A */
            if (p->sub_rate_q16) {
                uint64_t hs = orc_mix64(hr + 0x51ED27ULL * (q + 1));
                if ((hs & 0xFFFF) < p->sub_rate_q16) b = (b + 1 + (unsigned)((hs >> 16) % 3)) & 3;
            }
            out[i * (uint64_t)len + q] = L[b];
        }
    }
}
