/* tsta_psa_main.c -- TSTA_psa (-DTRACE) and TSTA_psa_notrace on B200.
 *
 * Drop-in for the reference drivers (reference psa/psa.c:539-668, built twice by its
 * Makefile:21-25): same flags, same stdout, same -o file bytes.  Everything between
 * "sequences are in memory" and "maxsorce / aligned strings are known" is one call into the
 * CUDA library (include/tsta_cuda.h); there is no CPU fallback.
 *
 *   -T (threads) and -W (tile width) are accepted and ignored: CUDA streams replace the
 *   thread pool, and the tile edge only changes how the reference schedules work, never the
 *   result.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fasta.h"
#include "tsta_cuda.h"

static void print_usage(void)
{
    /* text of reference psa/psa.c:521-537 */
    printf("-M                      the sorce of match [default: 2]\n");
    printf("-X                      the sorce of dismatch [default: -5]\n");
    printf("-E                      the sorce of extend-gap [default: -2]\n");
    printf("-O                      the sorce of open-gap [default: -4]\n");
    printf("-T                      the number of threads [default: 10]\n");
    printf("-W                      the width of block(Multiplication of simd data width) [default: 16]\n");
    printf("-1                      the input sequence1(fasta)\n");
    printf("-2                      the input sequence2(fasta)\n");
#ifdef TRACE
    printf("-o                      the output file [default: output.txt]\n");
    printf("example:\n./TSTA_psa -1 seq1.fa -2 seq2.fa -o output.txt\n");
#else
    printf("example:\n./TSTA_psa -1 seq1.fa -2 seq2.fa\n");
#endif
}

int main(int argc, char* argv[])
{
    tsta_params prm = {2, -5, -2, -4};
    const char *in1 = NULL, *in2 = NULL, *output = "output.txt";
    int c;
    while ((c = getopt(argc, argv, "M:X:E:O:T:W:1:2:o:")) != -1) {
        switch (c) {
        case 'M': prm.M = atoi(optarg); break;
        case 'X': prm.X = atoi(optarg); break;
        case 'E': prm.E = atoi(optarg); break;
        case 'O': prm.O = atoi(optarg); break;
        case 'T': break;   /* threads: not used on the GPU */
        case 'W': break;   /* tile width: does not change the result */
        case '1': in1 = optarg; break;
        case '2': in2 = optarg; break;
        case 'o': output = optarg; break;
        default: print_usage(); return 0;
        }
    }
    if (in1 == NULL || in2 == NULL) {
        printf("input file is not specified\n");
        print_usage();
        return 0;
    }
    (void)output;

    /* record 1 of each file; the longer read becomes sequence 1 (psa.c:192-199) */
    fasta_file f1, f2;
    fasta_read(in1, 1, &f1);
    fasta_read(in2, 1, &f2);
    if (f1.count == 0 || f2.count == 0) {
        fprintf(stderr, "no FASTA record found\n");
        return 1;
    }
    const fasta_record *s0 = &f1.rec[0], *s1 = &f2.rec[0];
    if (s1->len > s0->len) { const fasta_record* t = s0; s0 = s1; s1 = t; }

    tsta_psa_result res;
#ifdef TRACE
    const int want_trace = 1;
#else
    const int want_trace = 0;
#endif
    int rc = tsta_psa_align(&prm, s0->seq, (int64_t)s0->len, s1->seq, (int64_t)s1->len, want_trace, &res);
    if (rc != TSTA_OK) {
        fprintf(stderr, "TSTA_psa: %s\n", tsta_last_error());
        return 2;
    }
#ifdef TRACE
    FILE* fptr = fopen(output, "w");
    if (!fptr) { fprintf(stderr, "cannot open %s\n", output); return 1; }
    fputs(">1\n", fptr);            /* psa.c:512-515: no newline after the second row */
    fputs(res.a, fptr);
    fputs("\n>2\n", fptr);
    fputs(res.b, fptr);
    fclose(fptr);
#endif
    printf("maxsorce=%d\n", res.maxsorce);
    tsta_psa_result_free(&res);
    fasta_free(&f1);
    fasta_free(&f2);
    return 0;
}
