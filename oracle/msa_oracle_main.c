/* oracle/msa_oracle_main.c -- TEST INFRASTRUCTURE ONLY.
 * TSTA_msa with the CPU oracle (oracle/poa_oracle.c) as DP backend, through the product's own
 * host driver and graph code.  Comparing its -o file with the real reference's pins BOTH the
 * oracle fill and the host-side graph maintenance (tests/test_oracle_poa.py). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "msa_driver.h"

int tsta_oracle_poa_fill(int nrows, const int32_t* row_ptr, const int32_t* pred_row, const uint8_t* base,
                         const int32_t* hneg1, const char* read, int len, int M, int X, int E, int O, int L,
                         int8_t* source, int8_t* esource, int8_t* fsource, int32_t* lastsorce);
int tsta_oracle_poa_walk(int nrows, const int32_t* row_ptr, const int32_t* pred_row, int len,
                         int8_t* source, const int8_t* esource, const int8_t* fsource, int start_row, void* path);

typedef struct {
    int8_t *source, *esource, *fsource;
    poa_csr const* g;
    int len;
} octx;

static int o_fill(void* c, const poa_csr* g, const char* read, int len, const msa_params* p, int32_t* lastsorce)
{
    octx* o = (octx*)c;
    free(o->source); free(o->esource); free(o->fsource);
    const size_t cells = (size_t)g->nrows * (size_t)len + 1;
    o->source = (int8_t*)malloc(cells); o->esource = (int8_t*)malloc(cells); o->fsource = (int8_t*)malloc(cells);
    o->g = g; o->len = len;
    return tsta_oracle_poa_fill(g->nrows, g->row_ptr, g->pred_row, g->base, g->hneg1, read, len,
                                p->M, p->X, p->E, p->O, p->L, o->source, o->esource, o->fsource, lastsorce);
}
static int o_walk(void* c, int start, poa_step* path)
{
    octx* o = (octx*)c;
    return tsta_oracle_poa_walk(o->g->nrows, o->g->row_ptr, o->g->pred_row, o->len, o->source, o->esource,
                                o->fsource, start, path);
}
static const char* o_err(void* c) { (void)c; return "oracle backend failed (out of memory?)"; }

int main(int argc, char* argv[])
{
    octx o = {0};
    msa_backend be = {&o, o_fill, o_walk, o_err};
    _Static_assert(sizeof(poa_step) == 12, "poa_step layout shared with the oracle");
    return msa_main(argc, argv, &be);
}
