/* tsta_msa_main.c -- TSTA_msa on B200: the reference's command line (msa/msa.c) with the
 * sequence-to-graph DP and its traceback walk running in the CUDA library.  No CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "msa_driver.h"
#include "tsta_cuda.h"

typedef struct { tsta_poa* poa; double fill_ms, walk_ms; double cells, edge_cells; } gctx;

static int g_fill(void* c, const poa_csr* g, const char* read, int len, const msa_params* p, int32_t* lastsorce)
{
    gctx* x = (gctx*)c;
    tsta_params prm = {p->M, p->X, p->E, p->O};
    float ms = 0.f;
    int rc = tsta_poa_fill(x->poa, &prm, p->L, g->nrows, g->row_ptr, g->pred_row, g->base, g->hneg1,
                           read, len, lastsorce, &ms);
    x->fill_ms += ms;
    x->cells += (double)g->nrows * (double)len;
    x->edge_cells += (double)g->row_ptr[g->nrows] * (double)len;   /* sum of in-degrees x read length */
    return rc;
}

static int g_fill_resident(void* c, const poa_graph* g, const poa_shadow* d, const int32_t* hneg1,
                           const char* read, int len, const msa_params* p, int32_t* lastsorce)
{
    gctx* x = (gctx*)c;
    tsta_params prm = {p->M, p->X, p->E, p->O};
    float ms = 0.f;
    _Static_assert(sizeof(poa_rec2) == sizeof(tsta_poa_node2) && sizeof(poa_rec8) == sizeof(tsta_poa_node8), "node record layout");
    int rc = tsta_poa_graph_update(x->poa, g->nnode, d->n2, (const tsta_poa_node2*)d->r2, d->n8, (const tsta_poa_node8*)d->r8);
    if (rc == TSTA_OK)
        rc = tsta_poa_fill_resident(x->poa, &prm, p->L, g->len, g->sort, hneg1, read, len, lastsorce, &ms);
    x->fill_ms += ms;
    x->cells += (double)g->len * (double)len;
    x->edge_cells += (double)d->nedges * (double)len;
    return rc;
}

static int g_walk(void* c, int start, poa_step* path)
{
    gctx* x = (gctx*)c;
    int n = 0;
    _Static_assert(sizeof(poa_step) == sizeof(tsta_poa_step), "path step layout");
    float ms = 0.f;
    int rc = tsta_poa_walk(x->poa, start, (tsta_poa_step*)path, &n, &ms);
    x->walk_ms += ms;
    return rc == TSTA_OK ? n : rc;
}

static const char* g_err(void* c) { (void)c; return tsta_last_error(); }

int main(int argc, char* argv[])
{
    gctx x = {NULL, 0.0, 0.0, 0.0, 0.0};
    msa_backend be = {&x, g_fill, g_walk, g_err, g_fill_resident};
    /* the context is created lazily so that usage / missing-input paths need no GPU */
    int need_gpu = 0;
    for (int i = 1; i < argc; i++) if (argv[i][0] == '-' && argv[i][1] == 'i') need_gpu = 1;
    if (need_gpu && tsta_poa_create(&x.poa) != TSTA_OK) {
        fprintf(stderr, "TSTA_msa: %s\n", tsta_last_error());
        return 2;
    }
    int rc = msa_main(argc, argv, &be);
    if (getenv("TSTA_MSA_TIMING"))
        fprintf(stderr, "kernels: fill %.1f ms walk %.1f ms, %.0f node-cells, %.1f GCUPS (fill kernel), %.0f edge-cells\n",
                x.fill_ms, x.walk_ms, x.cells, x.fill_ms > 0 ? x.cells / x.fill_ms / 1e6 : 0.0, x.edge_cells);
    if (getenv("TSTA_MSA_TIMING") && x.poa)
        fprintf(stderr, "export: %lld bytes of graph tables and reads copied to the device (%s)\n", tsta_poa_h2d_bytes(x.poa),
                getenv("TSTA_MSA_FULL_EXPORT") ? "full export every read" : "resident graph: changed nodes + permutation");
    if (x.poa) tsta_poa_destroy(x.poa);
    return rc;
}
