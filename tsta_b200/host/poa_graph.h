/* poa_graph.h -- host-side partial-order graph for TSTA_msa (plain C, no CUDA).
 *
 * Own index-based design (nodes live in one growable array and refer to each other by id);
 * behaviour follows the reference's pointer-based graph:
 *   build from the first read        msa/c-t-simd.c:72-163   (poa_build_init)
 *   fuse one aligned read            msa/c-t-simd.c:659-977  (node_fuse) -- here driven by the
 *                                    compact PATH the DP/traceback returns instead of by
 *                                    poking into per-node trace rows
 *   topological order                msa/topo.c:179-260      (t_sort/toposort/tp and the final
 *                                    toposort1/tp1/modify heaviest-path pass)
 *   MSA rows + consensus             msa/result.c:6-79       (printf_result/printf_con)
 *
 * Order-sensitive details that define parity are kept: predecessor order (ties pick the
 * lowest index), append order of new nodes in `unsort`, mismatch groups, edge weights.
 */
#ifndef TSTA_POA_GRAPH_H
#define TSTA_POA_GRAPH_H

#include <stdint.h>
#include <stdio.h>

#define POA_MM_MAX 4        /* reference: mismatch_node[4] -- at most five alternatives per alignment column */
#define POA_SENTINEL (-1)   /* virtual predecessor of a head node (reference: `init`, sub == -1) */

typedef struct {
    int* pre;       /* predecessor node ids, in reference order (POA_SENTINEL for a head) */
    int* ew;        /* edge weight per predecessor */
    int in;         /* number of real predecessors (0 for a head; pre[0] is then the sentinel) */
    int pre_cap;
    int* next;      /* successor node ids, append order */
    int out, next_cap;
    int mm[POA_MM_MAX];   /* other members of this node's mismatch group */
    int mm_num;
    unsigned char base;
    int sub;        /* row index in the current topological order */
    uint8_t* passing_seq;   /* [nseq] which reads pass through this node */
    /* scratch of the sorts / consensus */
    int in_temp, passing, node_logo;
    int node_sorce, node_base_len, node_sorce_source;
    int alive;
} poa_node;

typedef struct {
    poa_node* node;
    int nnode, cap;
    int* unsort;    /* node ids, insertion order (rewritten to sorted order after every sort) */
    int* sort;      /* node ids in topological order: sort[row] */
    int len;        /* number of live nodes = number of DP rows */
    int order_cap;
    int last_node_num;
    int nseq;
    /* small-array pool: the pre / ew / next / passing_seq arrays of the nodes (a read adds one
     * provisional node per base every round, so malloc/free of these dominated the fuse step) */
    void* pool_blocks;        /* singly linked 1 MiB blocks */
    char* pool_cur;
    size_t pool_left;
    void* pool_free[28];      /* free lists by log2(size in 8-byte units) */
} poa_graph;

/* One traceback step, produced by the DP (GPU) or the oracle: the cell (row, col) and what
 * its effective `source` code said there (after the forced-continuation rules of
 * c-t-simd.c:767-784 were applied). */
enum { POA_STEP_VERT = 0, POA_STEP_MATCH = 1, POA_STEP_MISMATCH = 2, POA_STEP_HORIZ = 3 };
typedef struct {
    int32_t row;     /* DP row = index into sort[] */
    int32_t col;     /* read position */
    uint8_t kind;    /* POA_STEP_* */
    uint8_t pred;    /* predecessor index inside the row's pre[] (VERT/MATCH/MISMATCH) */
} poa_step;

/* CSR export of the current graph in row order: what the DP needs (c-t-simd.c:165-599 inputs). */
typedef struct {
    int nrows;
    int32_t* row_ptr;     /* [nrows+1] */
    int32_t* pred_row;    /* [nedges] predecessor ROW indices in pre[] order; POA_SENTINEL for heads */
    uint8_t* base;        /* [nrows] */
    uint8_t* is_sink;     /* [nrows] out == 0 */
    int32_t* hneg1;       /* [nrows] H(r,-1): O+E for heads, else max_j H(p_j,-1)+E (c-t-simd.c:187,196-200) */
    int nedges;
} poa_csr;

void poa_graph_init(poa_graph* g, const char* first_read, int nseq);
void poa_graph_free(poa_graph* g);
void poa_graph_export(const poa_graph* g, int E, int O, poa_csr* out);
void poa_csr_free(poa_csr* c);

/* Delta export for a graph that stays resident on the device (include/tsta_cuda.h: tsta_poa_graph_update):
 * the records (base, in-degree, predecessor NODE ids) of the live nodes that differ from what was sent
 * before.  The shadow remembers what the device holds; no bookkeeping inside poa_graph_fuse is needed. */
typedef struct { int32_t id; uint32_t base_in; int32_t pre[2]; } poa_rec2;
typedef struct { int32_t id; uint32_t base_in; int32_t pre[8]; } poa_rec8;
typedef struct {
    int cap;              /* nodes the shadow covers */
    uint32_t* meta;       /* [cap] base | in << 8 as last sent (0xffffffff: never sent) */
    int32_t* pre;         /* [cap][8] */
    poa_rec2* r2; int n2, cap2;   /* changed nodes with at most two predecessors */
    poa_rec8* r8; int n8, cap8;   /* ... with three to eight */
    long long nedges;     /* sum of in-degrees (heads count one) over the live rows */
} poa_shadow;
/* Returns 0, or -1 when a live node has more than eight predecessors (it cannot be resident). */
int poa_graph_delta(const poa_graph* g, poa_shadow* sh);
void poa_shadow_free(poa_shadow* sh);

/* Start row of the traceback (c-t-simd.c:737-750): among the last `last_node_num` sinks in row
 * order, the one with the largest lastsorce, lowest row on ties. lastsorce is indexed by row. */
int poa_graph_pick_start(const poa_graph* g, const int32_t* lastsorce);

/* Fuse read `num` (0-based record index) along `path` (nsteps steps, from the end of the read
 * backwards, as walked).  Mirrors node_fuse().  Returns 0, or -1 when a column would collect more than
 * POA_MM_MAX + 1 distinct symbols (the reference overflows mismatch_node[4] there). */
int poa_graph_fuse(poa_graph* g, const char* read, int num, const poa_step* path, int nsteps);

/* t_sort(): `final` != 0 runs the consensus variant (toposort1 + modify). */
void poa_graph_sort(poa_graph* g, int final);

/* printf_result(): ">k" rows followed by "consensus:". */
void poa_graph_write(poa_graph* g, FILE* out);

#endif
