/* msa_driver.h -- the TSTA_msa command line around an abstract DP backend.
 *
 * Mirrors reference msa/msa.c:75-151 (flags, progress text, read loop) and control()
 * (msa/c-t-simd.c:979-1049) minus its tile loop: per read the graph is exported in row order,
 * the backend fills the DP and walks the traceback, the host fuses the path and re-sorts.
 * The product binds the CUDA library (tsta_msa_main.c); the test suite binds the CPU oracle
 * to the very same driver to pin the oracle against the reference's goldens.
 */
#ifndef TSTA_MSA_DRIVER_H
#define TSTA_MSA_DRIVER_H
#include <stdint.h>
#include "poa_graph.h"

typedef struct {
    int M, X, E, O;
    int L;            /* the reference's tile edge (SIMD lanes x -W); observable through quirk Q1 only */
} msa_params;

typedef struct msa_backend {
    void* ctx;
    /* Fill the DP of `read` against the graph; lastsorce[r] = H(r, len-1) for every row. */
    int (*fill)(void* ctx, const poa_csr* g, const char* read, int len, const msa_params* p, int32_t* lastsorce);
    /* Walk the traceback from (start_row, len-1); path has room for nrows+len steps. */
    int (*walk)(void* ctx, int start_row, poa_step* path);
    const char* (*error)(void* ctx);
    /* Optional (NULL: not offered): the same fill for a backend that keeps the graph resident -- it receives the
     * graph itself (its sort[] order) and the records of the nodes that changed since its last call (`delta`);
     * hneg1 is NULL unless the read contains 'N' (see include/tsta_cuda.h: tsta_poa_fill_resident). */
    int (*fill_resident)(void* ctx, const poa_graph* g, const poa_shadow* delta, const int32_t* hneg1,
                         const char* read, int len, const msa_params* p, int32_t* lastsorce);
} msa_backend;

int msa_main(int argc, char* argv[], const msa_backend* be);

#endif
