/* fasta.c -- see fasta.h */
#include "fasta.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static char* slurp(const char* path, size_t* size_out)
{
    FILE* probe = fopen(path, "r");
    if (!probe) {
        fprintf(stderr, "File %s does not exist.\n", path);   /* reference seqio.c:175-183 */
        exit(1);
    }
    fclose(probe);
    gzFile g = gzopen(path, "rb");   /* transparent for plain files, inflates gzip */
    if (!g) {
        fprintf(stderr, "File %s does not exist.\n", path);
        exit(1);
    }
    size_t cap = 1 << 20, n = 0;
    char* buf = (char*)malloc(cap + 1);
    for (;;) {
        if (n == cap) {
            cap *= 2;
            buf = (char*)realloc(buf, cap + 1);
        }
        if (!buf) { fprintf(stderr, "out of memory reading %s\n", path); exit(1); }
        int got = gzread(g, buf + n, (unsigned)((cap - n) > (1u << 30) ? (1u << 30) : (cap - n)));
        if (got <= 0) break;
        n += (size_t)got;
    }
    gzclose(g);
    buf[n] = '\0';
    *size_out = n;
    return buf;
}

int fasta_read(const char* path, size_t max_records, fasta_file* out)
{
    size_t size = 0;
    char* buf = slurp(path, &size);
    out->rec = NULL;
    out->count = 0;
    /* format guess: the first '>' or '@' anywhere decides (seqio.c:286-305) */
    size_t p = 0;
    while (p < size && buf[p] != '>' && buf[p] != '@') p++;
    if (p < size && buf[p] == '@') {
        fprintf(stderr, "Cannot read fasta record from a fastq file.\n");   /* seqio.c:470 */
        exit(1);
    }
    size_t cap = 0;
    while (p < size) {
        /* p is at '>' : skip the header line */
        while (p < size && buf[p] != '\n') p++;
        if (p < size) p++;
        /* sequence lines until a line that starts with '>' */
        size_t scap = 256, slen = 0;
        char* s = (char*)malloc(scap + 1);
        while (p < size && buf[p] != '>') {
            size_t e = p;
            while (e < size && buf[e] != '\n') e++;
            size_t l = e - p;
            if (l > 0 && e < size && buf[e - 1] == '\r') l--;   /* CRLF (only when a '\n' follows) */
            if (slen + l > scap) {
                while (slen + l > scap) scap *= 2;
                s = (char*)realloc(s, scap + 1);
            }
            memcpy(s + slen, buf + p, l);
            slen += l;
            p = e < size ? e + 1 : e;
        }
        s[slen] = '\0';
        if (out->count == cap) {
            cap = cap ? cap * 2 : 8;
            out->rec = (fasta_record*)realloc(out->rec, cap * sizeof(fasta_record));
        }
        out->rec[out->count].seq = s;
        out->rec[out->count].len = slen;
        out->count++;
        if (max_records && out->count >= max_records) break;
    }
    free(buf);
    return 0;
}

void fasta_free(fasta_file* f)
{
    for (size_t k = 0; k < f->count; k++) free(f->rec[k].seq);
    free(f->rec);
    f->rec = NULL;
    f->count = 0;
}
