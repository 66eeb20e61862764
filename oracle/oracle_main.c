/*
 * kmer_oracle — command-line front end of the CPU restatement.  TEST INFRASTRUCTURE ONLY.
 *
 *   kmer_oracle count <in.fastq> <k> <out.count>      dump in count_kmers.py:32-34 format
 *   kmer_oracle gen <mode> <seed> <n_reads> <read_len> <genome_len> <sub_q16> <out.fastq>
 *                                                     FASTQ in generateFakeSequences.py:14-18 layout
 *                                                     (@id / seq / + / '&' * len)
 */
#include "oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int usage(void) {
    fprintf(stderr,
            "usage: kmer_oracle count <in.fastq> <k> <out.count>\n"
            "       kmer_oracle gen <mode> <seed> <n_reads> <read_len> <genome_len> <sub_q16> <out.fastq>\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    if (!strcmp(argv[1], "count") && argc == 5) {
        orc_counts* c = orc_count_fastq(argv[2], (unsigned)atoi(argv[3]));
        if (!c) { fprintf(stderr, "count failed\n"); return 1; }
        fprintf(stderr, "distinct=%llu total=%llu skipped=%llu\n", (unsigned long long)c->n_distinct,
                (unsigned long long)c->n_total, (unsigned long long)c->n_skipped);
        int rc = orc_write_dump(c, argv[4]);
        orc_free(c);
        return rc ? 1 : 0;
    }
    if (!strcmp(argv[1], "gen") && argc == 9) {
        orc_gen_params p;
        memset(&p, 0, sizeof p);
        p.mode = (uint32_t)strtoul(argv[2], NULL, 0);
        p.seed = strtoull(argv[3], NULL, 0);
        p.n_reads = strtoull(argv[4], NULL, 0);
        p.read_len = (uint32_t)strtoul(argv[5], NULL, 0);
        p.genome_len = strtoull(argv[6], NULL, 0);
        p.sub_rate_q16 = (uint32_t)strtoul(argv[7], NULL, 0);
        FILE* f = fopen(argv[8], "wb");
        if (!f) { perror("open"); return 1; }
        char* seq = (char*)malloc(p.read_len + 1);
        char* qual = (char*)malloc(p.read_len + 1);
        memset(qual, '&', p.read_len); qual[p.read_len] = 0; seq[p.read_len] = 0;
        for (uint64_t r = 0; r < p.n_reads; ++r) {
            orc_gen_reads(&p, r, 1, seq);
            fprintf(f, "@seq_%llu\n%s\n+\n%s\n", (unsigned long long)r, seq, qual);
        }
        free(seq); free(qual);
        return fclose(f) ? 1 : 0;
    }
    return usage();
}
