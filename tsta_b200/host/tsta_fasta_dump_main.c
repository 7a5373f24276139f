/* tsta_fasta_dump -- prints "<index> <length> <fnv1a-64 hex>" per record: lets the CPU test
 * suite check the drivers' FASTA front-end (plain, CRLF, multi-line, gzip) without a GPU. */
#include <stdio.h>
#include <stdlib.h>
#include "fasta.h"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: tsta_fasta_dump file [max_records]\n"); return 2; }
    fasta_file f;
    fasta_read(argv[1], argc > 2 ? (size_t)atol(argv[2]) : 0, &f);
    for (size_t k = 0; k < f.count; k++) {
        unsigned long long h = 1469598103934665603ULL;
        for (size_t i = 0; i < f.rec[k].len; i++) { h ^= (unsigned char)f.rec[k].seq[i]; h *= 1099511628211ULL; }
        printf("%zu %zu %016llx\n", k, f.rec[k].len, h);
    }
    fasta_free(&f);
    return 0;
}
