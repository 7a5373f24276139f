/* msa_driver.c -- see msa_driver.h */
#include "msa_driver.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <time.h>

#include "fasta.h"

/* optional phase timing (TSTA_MSA_TIMING=1), printed to stderr at exit; off by default so the
 * driver's stderr stays identical to the reference's */
static double t_export, t_fill, t_walk, t_fuse, t_sort;
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void print_usage(void)
{
    /* text of reference msa/msa.c:62-73 */
    printf("-M                      the sorce of match [default: 2]\n");
    printf("-X                      the sorce of dismatch [default: -5]\n");
    printf("-E                      the sorce of extend-gap [default: -2]\n");
    printf("-O                      the sorce of open-gap [default: -4]\n");
    printf("-T                      the number of threads [default: 10]\n");
    printf("-W                      the width of block(Multiplication of simd data width) [default: 16]\n");
    printf("-i                      the input sequence(fasta format)\n");
    printf("-o                      the output file [default: output.txt]\n");
    printf("example:\n./TSTA_msa -i seq.fa -o output.txt\n");
}

#define MSA_ETOOMANY (-77)

/* control() without the tile loop (c-t-simd.c:979-1049).  Backends that keep the graph on the device get the
 * nodes that changed (poa_graph_delta) instead of a full export in row order; a graph with a node of more than
 * eight predecessors -- or TSTA_MSA_FULL_EXPORT=1 -- stays on the full export for the rest of the job. */
static poa_shadow g_shadow;
static int g_resident = -1;          /* -1: not decided yet, 0: full export, 1: resident */

static int align_read(poa_graph* g, const char* read, int num, const msa_params* p, const msa_backend* be)
{
    const int len = (int)strlen(read);
    if (g_resident < 0) g_resident = be->fill_resident != NULL && getenv("TSTA_MSA_FULL_EXPORT") == NULL;
    poa_csr csr;
    memset(&csr, 0, sizeof csr);
    double t0 = now_s();
    int resident = g_resident;
    if (resident && poa_graph_delta(g, &g_shadow) != 0) resident = g_resident = 0;
    const int need_csr = !resident || memchr(read, 'N', (size_t)len) != NULL;   /* ('N' reads: the backend wants hneg1) */
    if (need_csr) poa_graph_export(g, p->E, p->O, &csr);
    t_export += now_s() - t0;
    const int nrows = g->len;
    int32_t* lastsorce = (int32_t*)malloc((size_t)(nrows + 1) * sizeof(int32_t));
    poa_step* path = (poa_step*)malloc((size_t)(nrows + len + 1) * sizeof(poa_step));
    t0 = now_s();
    int rc = resident ? be->fill_resident(be->ctx, g, &g_shadow, need_csr ? csr.hneg1 : NULL, read, len, p, lastsorce)
                      : be->fill(be->ctx, &csr, read, len, p, lastsorce);
    t_fill += now_s() - t0;
    if (rc == 0) {
        const int start = poa_graph_pick_start(g, lastsorce);
        t0 = now_s();
        const int nsteps = be->walk(be->ctx, start, path);
        t_walk += now_s() - t0;
        t0 = now_s();
        if (nsteps < 0) rc = nsteps;
        else if (poa_graph_fuse(g, read, num, path, nsteps) != 0) rc = MSA_ETOOMANY;
        t_fuse += now_s() - t0;
    }
    free(path);
    free(lastsorce);
    if (need_csr) poa_csr_free(&csr);
    return rc;
}

int msa_main(int argc, char* argv[], const msa_backend* be)
{
    msa_params prm = {2, -5, -2, -4, 0};
    int bS = 10, c;
    const char *input = NULL, *output = "output.txt";
    while ((c = getopt(argc, argv, "M:X:E:O:T:W:i:o:")) != -1) {
        switch (c) {
        case 'M': prm.M = (signed char)atoi(optarg); break;   /* the reference keeps these in a char (msa.c:21-24) */
        case 'X': prm.X = (signed char)atoi(optarg); break;
        case 'E': prm.E = (signed char)atoi(optarg); break;
        case 'O': prm.O = (signed char)atoi(optarg); break;
        case 'T': break;                                       /* threads: not used on the GPU */
        case 'W': bS = atoi(optarg); break;
        case 'i': input = optarg; break;
        case 'o': output = optarg; break;
        default: print_usage(); return 0;
        }
    }
    if (input == NULL) {
        printf("input file is not specified\n");
        print_usage();
        return 0;
    }
    /* tile edge of the reference build being matched: AVX-512 lanes (64) x -W (msa.c:114-116);
     * TSTA_SIMD_BLOCK=16|32 selects the SSE/AVX2 reference builds' geometry instead */
    int block = 64;
    const char* eb = getenv("TSTA_SIMD_BLOCK");
    if (eb && (atoi(eb) == 16 || atoi(eb) == 32 || atoi(eb) == 64)) block = atoi(eb);
    prm.L = block * (bS > 0 ? bS : 1);

    fasta_file f;
    fasta_read(input, 0, &f);
    const int n = (int)f.count;
    fprintf(stderr, "seq_num: %d\n", n);                        /* msa.c:47 */
    if (n == 0) { fprintf(stderr, "no FASTA record found\n"); return 1; }

    poa_graph g;
    poa_graph_init(&g, f.rec[0].seq, n);
    int rc = 0;
    for (int i = 1; i < n - 1 && rc == 0; i++) {                /* msa.c:130-141 */
        rc = align_read(&g, f.rec[i].seq, i, &prm, be);
        double t0 = now_s();
        poa_graph_sort(&g, 0);
        t_sort += now_s() - t0;
        if (i % 100 == 0) printf("\r[%d/%d]", i, n);
    }
    if (rc == 0) {
        printf("\r[%d/%d]", n, n);
        printf("\n");
        rc = align_read(&g, f.rec[n - 1].seq, n - 1, &prm, be);
        poa_graph_sort(&g, 1);
    }
    if (rc == MSA_ETOOMANY) {
        fprintf(stderr, "TSTA_msa: more than %d different symbols aligned in one column (the reference's mismatch_node[4] limit); "
                        "upper-case the input or remove ambiguity codes\n", POA_MM_MAX + 1);
        return 2;
    }
    if (rc != 0) {
        fprintf(stderr, "TSTA_msa: %s\n", be->error ? be->error(be->ctx) : "alignment failed");
        return 2;
    }
    FILE* res = fopen(output, "w");
    if (!res) { fprintf(stderr, "cannot open %s\n", output); return 1; }
    poa_graph_write(&g, res);
    fclose(res);
    if (getenv("TSTA_MSA_TIMING"))
        fprintf(stderr, "timing: export %.3f fill %.3f walk %.3f fuse %.3f sort %.3f s (rows %d)\n",
                t_export, t_fill, t_walk, t_fuse, t_sort, g.len);
    poa_graph_free(&g);
    poa_shadow_free(&g_shadow);
    fasta_free(&f);
    return 0;
}
