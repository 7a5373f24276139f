/* tsta_psa_batch_main.c -- TSTA_psa_batch: the batch front-end of SURVEY.md 8f-2.
 *
 * The reference drivers align record 1 of each input only (psa/psa.c:192-193).  This driver
 * pairs record k of -1 with record k of -2 (plain or gzip FASTA, same reader as TSTA_psa),
 * aligns every pair exactly like TSTA_psa would (longer read first, psa.c:194-199) through
 * tsta_psa_batch, and prints one "maxsorce=%d" line per pair in input order.  With -o the
 * alignments are written as the concatenation of the reference's per-pair output
 * (">1\n<a>\n>2\n<b>\n" per pair -- the reference omits the final newline of its single pair;
 * here every pair ends with one so that the records can be split again).
 *
 *   -G n   spread the pairs round-robin over the first n GPUs of the box (one host thread per
 *          GPU; independent units, no collective: SURVEY.md 8e)
 *   -n     score only (TSTA_psa_notrace behaviour), no -o file
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fasta.h"
#include "tsta_cuda.h"

typedef struct {
    int device, ngpu, trace;
    tsta_params prm;
    const fasta_file *f1, *f2;
    size_t npairs;
    /* per shard (pairs device, device+ngpu, ...) */
    size_t count;
    char *seqs0, *seqs1;
    int64_t *off0, *off1, *ops_off;
    int32_t *len0, *len1, *maxsorce, *corner, *alen;
    uint8_t* ops;
    int rc;
    char err[256];
} shard;

static void* run_shard(void* arg)
{
    shard* s = (shard*)arg;
    s->rc = tsta_set_device(s->device);
    if (s->rc == TSTA_OK)
        s->rc = tsta_psa_batch(&s->prm, (int64_t)s->count, s->seqs0, s->off0, s->len0, s->seqs1, s->off1, s->len1,
                               s->trace, s->maxsorce, s->corner, s->ops, s->ops_off, s->alen, NULL);
    if (s->rc != TSTA_OK) snprintf(s->err, sizeof s->err, "%s", tsta_last_error());
    return NULL;
}

static int build_shard(shard* s)
{
    size_t tot0 = 0, tot1 = 0, totops = 0, k, c = 0;
    for (k = (size_t)s->device; k < s->npairs; k += (size_t)s->ngpu) {
        const size_t a = s->f1->rec[k].len, b = s->f2->rec[k].len;
        tot0 += a > b ? a : b;          /* >= b: ties keep the order */
        tot1 += a > b ? b : (a < b ? a : b);
        totops += a + b;
        c++;
    }
    s->count = c;
    s->seqs0 = (char*)malloc(tot0 + 1); s->seqs1 = (char*)malloc(tot1 + 1);
    s->off0 = (int64_t*)malloc((c + 1) * 8); s->off1 = (int64_t*)malloc((c + 1) * 8); s->ops_off = (int64_t*)malloc((c + 1) * 8);
    s->len0 = (int32_t*)malloc((c + 1) * 4); s->len1 = (int32_t*)malloc((c + 1) * 4);
    s->maxsorce = (int32_t*)malloc((c + 1) * 4); s->corner = (int32_t*)malloc((c + 1) * 4); s->alen = (int32_t*)malloc((c + 1) * 4);
    s->ops = s->trace ? (uint8_t*)malloc(totops + 1) : NULL;
    if (!s->seqs0 || !s->seqs1 || !s->off0 || !s->off1 || !s->ops_off || !s->len0 || !s->len1 || !s->maxsorce ||
        !s->corner || !s->alen || (s->trace && !s->ops)) return -1;
    size_t p0 = 0, p1 = 0, po = 0, i = 0;
    for (k = (size_t)s->device; k < s->npairs; k += (size_t)s->ngpu, i++) {
        const fasta_record *x = &s->f1->rec[k], *y = &s->f2->rec[k];
        if (y->len > x->len) { const fasta_record* t = x; x = y; y = t; }     /* psa.c:194-199 */
        memcpy(s->seqs0 + p0, x->seq, x->len); memcpy(s->seqs1 + p1, y->seq, y->len);
        s->off0[i] = (int64_t)p0; s->off1[i] = (int64_t)p1; s->ops_off[i] = (int64_t)po;
        s->len0[i] = (int32_t)x->len; s->len1[i] = (int32_t)y->len;
        p0 += x->len; p1 += y->len; po += x->len + y->len;
    }
    return 0;
}

static void usage(void)
{
    printf("TSTA_psa_batch -1 reads1.fa -2 reads2.fa [-o out.txt] [-n] [-G gpus] [-M 2 -X -5 -E -2 -O -4]\n"
           "pairs record k of -1 with record k of -2; one maxsorce=%%d line per pair\n");
}

int main(int argc, char* argv[])
{
    tsta_params prm = {2, -5, -2, -4};
    const char *in1 = NULL, *in2 = NULL, *output = NULL;
    int c, ngpu = 1, score_only = 0;
    while ((c = getopt(argc, argv, "M:X:E:O:T:W:1:2:o:G:n")) != -1) {
        switch (c) {
        case 'M': prm.M = atoi(optarg); break;
        case 'X': prm.X = atoi(optarg); break;
        case 'E': prm.E = atoi(optarg); break;
        case 'O': prm.O = atoi(optarg); break;
        case 'T': case 'W': break;
        case '1': in1 = optarg; break;
        case '2': in2 = optarg; break;
        case 'o': output = optarg; break;
        case 'G': ngpu = atoi(optarg); break;
        case 'n': score_only = 1; break;
        default: usage(); return 0;
        }
    }
    if (!in1 || !in2) { printf("input file is not specified\n"); usage(); return 0; }
    const int trace = !score_only && output != NULL;
    fasta_file f1, f2;
    fasta_read(in1, 0, &f1);
    fasta_read(in2, 0, &f2);
    if (f1.count != f2.count) {
        fprintf(stderr, "TSTA_psa_batch: %zu records in %s but %zu in %s\n", f1.count, in1, f2.count, in2);
        return 1;
    }
    const int have = tsta_device_count();
    if (have <= 0) { fprintf(stderr, "TSTA_psa_batch: %s\n", have < 0 ? tsta_last_error() : "no CUDA device"); return 2; }
    if (ngpu < 1) ngpu = 1;
    if (ngpu > have) ngpu = have;
    if ((size_t)ngpu > f1.count && f1.count > 0) ngpu = (int)f1.count;

    shard* sh = (shard*)calloc((size_t)ngpu, sizeof(shard));
    pthread_t* th = (pthread_t*)calloc((size_t)ngpu, sizeof(pthread_t));
    for (int g = 0; g < ngpu; g++) {
        sh[g].device = g; sh[g].ngpu = ngpu; sh[g].trace = trace; sh[g].prm = prm;
        sh[g].f1 = &f1; sh[g].f2 = &f2; sh[g].npairs = f1.count;
        if (build_shard(&sh[g])) { fprintf(stderr, "TSTA_psa_batch: out of memory\n"); return 1; }
    }
    for (int g = 0; g < ngpu; g++) pthread_create(&th[g], NULL, run_shard, &sh[g]);
    for (int g = 0; g < ngpu; g++) pthread_join(th[g], NULL);
    for (int g = 0; g < ngpu; g++)
        if (sh[g].rc != TSTA_OK) { fprintf(stderr, "TSTA_psa_batch: %s\n", sh[g].err); return 2; }

    FILE* fo = NULL;
    if (trace) {
        fo = fopen(output, "w");
        if (!fo) { fprintf(stderr, "cannot open %s\n", output); return 1; }
    }
    for (size_t k = 0; k < f1.count; k++) {
        const shard* s = &sh[k % (size_t)ngpu];
        const size_t i = k / (size_t)ngpu;
        printf("maxsorce=%d\n", s->maxsorce[i]);
        if (fo) {
            /* the alen alignment columns sit at the END of the pair's slot, in reading order:
             * 1 = both bases, 0 = base of the longer read against a gap, 2 = gap against a base of the shorter */
            const fasta_record *x = &f1.rec[k], *y = &f2.rec[k];
            if (y->len > x->len) { const fasta_record* t = x; x = y; y = t; }
            const size_t cap = x->len + y->len, n = (size_t)s->alen[i];
            const uint8_t* ops = s->ops + s->ops_off[i] + (cap - n);
            char* row = (char*)malloc(n + 1);
            size_t p = 0;
            fputs(">1\n", fo);
            for (size_t q = 0; q < n; q++) row[q] = ops[q] != 2 ? x->seq[p++] : '-';
            fwrite(row, 1, n, fo);
            fputs("\n>2\n", fo);
            p = 0;
            for (size_t q = 0; q < n; q++) row[q] = ops[q] != 0 ? y->seq[p++] : '-';
            fwrite(row, 1, n, fo);
            fputc('\n', fo);
            free(row);
        }
    }
    if (fo) fclose(fo);
    fasta_free(&f1);
    fasta_free(&f2);
    return 0;
}
