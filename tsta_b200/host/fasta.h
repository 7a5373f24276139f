/* fasta.h -- minimal FASTA reader for the TSTA_* drivers (plain or gzip input).
 *
 * Mirrors what the reference's readers rely on from seqio (reference seqio.c:463-541,
 * 427-461): a record starts at a '>' that begins a line; the name runs to the first blank,
 * the rest of the header line is a comment; the sequence is every following line up to the
 * next line that begins with '>', with the line terminator ('\n' and a preceding '\r')
 * removed and empty lines skipped.  Bytes are kept as they are (no case folding).  gzip
 * input is detected by its magic bytes (seqio.c:196-211).
 */
#ifndef TSTA_FASTA_H
#define TSTA_FASTA_H
#include <stddef.h>

typedef struct {
    char* seq;     /* NUL-terminated */
    size_t len;
} fasta_record;

typedef struct {
    fasta_record* rec;
    size_t count;
} fasta_file;

/* Reads up to max_records records (0 = all).  Exits like the reference on a missing file
 * ("File %s does not exist.", exit 1) or a FASTQ input.  Returns 0 on success. */
int fasta_read(const char* path, size_t max_records, fasta_file* out);
void fasta_free(fasta_file* f);

#endif
