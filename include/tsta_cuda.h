/*
 * tsta_cuda.h -- C ABI of the B200 (sm_100a) global-alignment hot path.
 *
 * This is the drop-in boundary for bxskdh/TSTA.  The reference has no library or FFI
 * surface (its public interface is three executables, reference Makefile:11-32); the seam
 * this ABI replaces is the region of each main()/control() between "sequences are in
 * memory" and "scores / direction matrices are in memory":
 *
 *   tsta_psa_align / tsta_psa_plan_*   replace reference psa/psa.c:594-643
 *       (state allocation 598-614, blockmatrix_init 219-232, the anti-diagonal tile loop
 *        620-639 that feeds block_alignment()/row() 234-433 to the pthread pool, and
 *        trace() 435-519).  Inputs are what readseq() (179-217) leaves in seq[0], seq[1];
 *        outputs are `ms` (printed at 666) and the two strings trace() writes (512-515).
 *   tsta_psa_batch                      the same seam for many independent pairs
 *       (an extension: the reference reads record 1 only, psa/psa.c:192-193).
 *   tsta_poa_*                          replace reference msa/c-t-simd.c:1003-1043 (control():
 *       per-node trace allocation + the tile loop feeding block_alignment()/
 *       block_line_alignment() 165-657).  Outputs are what node_fuse() (659-977) reads:
 *       lastsorce of every sink and source/esource/fsource at (row, column).
 *
 * Plain C types only; caller-owned inputs; library-owned outputs are released with the
 * matching *_free / *_destroy.  Every entry point returns 0 on success and a negative
 * TSTA_E* code on failure; tsta_last_error() gives the message.  There is NO CPU fallback:
 * without a CUDA device every compute entry point fails with TSTA_ENODEV.
 */
#ifndef TSTA_CUDA_H
#define TSTA_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSTA_OK        0
#define TSTA_EINVAL   -1   /* bad argument / parameters outside the parity domain */
#define TSTA_ENODEV   -2   /* no usable CUDA device                              */
#define TSTA_ENOMEM   -3   /* host or device allocation failed                   */
#define TSTA_ECUDA    -4   /* a CUDA call or kernel failed                       */

/* Scoring parameters: reference globals M, X, E, O (psa/psa.c:114-117, msa/msa.c:21-24).
 * Parity domain: O <= 0, E <= 0, X <= 0, all |.| <= 127. */
typedef struct tsta_params {
    int32_t M;  /* match      (-M, default  2) */
    int32_t X;  /* mismatch   (-X, default -5) */
    int32_t E;  /* gap extend (-E, default -2) */
    int32_t O;  /* gap open   (-O, default -4; 0 = linear gaps) */
} tsta_params;

const char* tsta_last_error(void);
int  tsta_device_count(void);            /* >= 0, or a negative TSTA_E* code */
int  tsta_set_device(int device);        /* device used by the calling thread's later calls */
const char* tsta_version(void);

/* ------------------------------------------------------------------------------------ */
/* Pairwise alignment (TSTA_psa / TSTA_psa_notrace)                                      */
/* ------------------------------------------------------------------------------------ */

typedef struct tsta_psa_result {
    int32_t maxsorce;   /* what the reference prints: max(-100, max H over all cells) (psa.c:590,666) */
    int32_t corner;     /* H(m-1, n-1), the global score (not printed by the reference)              */
    int64_t alen;       /* alignment length; 0 when no traceback was requested                        */
    char*   a;          /* ">1" row: s0 with '-' (NUL-terminated, library-owned) or NULL              */
    char*   b;          /* ">2" row: s1 with '-' (NUL-terminated, library-owned) or NULL              */
} tsta_psa_result;

/* One pair, host buffers in, host strings out.  s0 must be the longer read (the caller
 * applies the swap of psa/psa.c:194-199; tsta_psa_align returns TSTA_EINVAL if n < m).
 * want_trace = 0 behaves like TSTA_psa_notrace, 1 like TSTA_psa. */
int  tsta_psa_align(const tsta_params* p, const char* s0, int64_t n, const char* s1, int64_t m,
                    int want_trace, tsta_psa_result* out);
/* tsta_psa_align keeps one plan per (calling thread, trace mode) so that repeated calls reuse the
 * device buffers (a traced 100 kbp pair keeps ~7 GB of directions).  They are freed when the
 * thread exits, or right away by this call (calling thread's plans only).  The call also returns the
 * buffers, streams and pinned staging memory the calling thread's tsta_psa_batch* calls keep between
 * calls (those are otherwise left to process exit). */
void tsta_psa_release_cache(void);
void tsta_psa_result_free(tsta_psa_result* r);

/* Plan API: device-resident state that can be run repeatedly (what bench.py times).
 *   create  -> allocates nothing big yet
 *   upload  -> H2D copy + encoding of one pair (sizes may change between uploads)
 *   run     -> fill (+ traceback) kernels on the plan's stream; optional CUDA-event timings
 *   fetch   -> D2H of scores (and strings when traced)                                     */
typedef struct tsta_psa_plan tsta_psa_plan;
int  tsta_psa_plan_create(tsta_psa_plan** plan, int want_trace);
int  tsta_psa_plan_upload(tsta_psa_plan* plan, const tsta_params* p,
                          const char* s0, int64_t n, const char* s1, int64_t m);
int  tsta_psa_plan_run(tsta_psa_plan* plan, float* fill_ms, float* trace_ms);
/* run = launch (asynchronous: enqueue the kernels on the plan's stream) + wait (synchronise, timings) */
int  tsta_psa_plan_launch(tsta_psa_plan* plan);
int  tsta_psa_plan_wait(tsta_psa_plan* plan, float* fill_ms, float* trace_ms);
int  tsta_psa_plan_fetch(tsta_psa_plan* plan, tsta_psa_result* out);
/* Debug / parity helpers: last DP row H(m-1, 0..n-1) and the packed direction codes
 * (one byte per cell: back | cF<<2 | cE<<3, row-major m x n). */
int  tsta_psa_plan_lastrow(tsta_psa_plan* plan, int32_t* out_n);
int  tsta_psa_plan_dirs(tsta_psa_plan* plan, uint8_t* out_m_by_n);
/* Number of kernels the last run launched and the lane width it chose (for reporting). */
int  tsta_psa_plan_info(tsta_psa_plan* plan, int* launches, int* cols_per_lane, int* bands);
/* 1 when the last upload selected the packed int16x2 kernels (two cells per register; small
 * alphabets and score ranges, csrc/psa_fill16.cuh), 0 for the int32 kernels. */
int  tsta_psa_plan_packed(tsta_psa_plan* plan);
/* Traced pairs on one GPU use the CHECKPOINTED traceback (csrc/psa_fill16.cuh): a score-only fill that saves
 * the lane registers every 128 wavefront steps, then the traced form of the tiles in a corridor around the
 * path, then the walk; a walk that leaves its corridor is resumed in a new one.  This reports the last run:
 * corridors used (1 = the first one sufficed; 0 = another traceback mode ran) and tiles re-run in total.
 * Environment (experiments / tests): TSTA_PSA_TRACE_MODE=full keeps the direction words of the whole matrix,
 * TSTA_PSA_CORRIDOR=<columns> the corridor's margin (default 2048, at least 1024), TSTA_PSA_TILE_BUDGET_MB. */
int  tsta_psa_plan_trace_stats(tsta_psa_plan* plan, int* corridors, long long* tiles, long long* tiles_of_matrix);
/* diagnostics: with TSTA_PSA_TIMELINE=1 in the environment every band of the last run left 8 globaltimer
 * stamps (ns): entry, first boundary rows in, start of chunks 1..5, end; then stale polls, cycles spent
 * importing the left boundary, number of imports.  out = [bands][16]. */
int  tsta_psa_plan_timeline(tsta_psa_plan* plan, unsigned long long* out);
void tsta_psa_plan_destroy(tsta_psa_plan* plan);

/* One long pair spread over several GPUs as COLUMN SLICES (score-only): GPU g
 * holds columns [col0, col0 + n_slice) of s0 and all of s1, and its last band stores the right
 * boundary column (the reference's V[] / F[], psa/psa.c:355-356) straight into the next GPU's
 * memory over NVLink, as the same tagged 64-bit elements that bands use inside one GPU.  All
 * slices are launched together (one process or thread per GPU); slice g only ever waits for
 * slice g-1.  Every slice but the last must be a multiple of 1536 columns.
 *   upload_slice      like upload, for one slice (has_left / has_right: neighbours exist)
 *   inbox_handle      CUDA IPC handle of this slice's left-boundary inbox (has_left slices)
 *   connect_right     open the right neighbour's inbox handle (another process)
 *   connect_right_local  same, when both plans live in this process
 * Every slice must be run the same number of times (the run counter is the elements' tag).
 * fetch() returns the slice's own maximum (combine with max) and, on the last slice, the corner. */
int  tsta_psa_plan_upload_slice(tsta_psa_plan* plan, const tsta_params* p, const char* s0_slice, int64_t n_slice,
                                int64_t col0, int64_t n_total, const char* s1, int64_t m, int has_left, int has_right);
int  tsta_psa_plan_inbox_handle(tsta_psa_plan* plan, unsigned char handle[64]);
int  tsta_psa_plan_connect_right(tsta_psa_plan* plan, const unsigned char handle[64]);
int  tsta_psa_plan_connect_right_local(tsta_psa_plan* plan, tsta_psa_plan* right);
/* Traceback of a sliced pair (plans created with want_trace = 1): after all slices have run, walk
 * them right to left.  trace_from walks THIS slice's direction matrix from (start_i, its last column)
 * in state start_forced (0 free, 2 forced left, 1 forced up) until it leaves through the left edge
 * (end_i >= 0: continue in the left neighbour with end_i / end_forced) or the top (end_i < 0: done).
 * The last slice starts at (m-1, free).  ops() copies the nops moves of the last trace_from in walk
 * order (1 = both bases, 0 = base of s0 against a gap, 2 = gap against a base of s1); the caller
 * concatenates them right to left and adds the leftovers of psa/psa.c:497-510. */
int  tsta_psa_plan_trace_from(tsta_psa_plan* plan, int64_t start_i, int start_forced,
                              int64_t* end_i, int* end_forced, int64_t* nops, float* ms);
int  tsta_psa_plan_ops(tsta_psa_plan* plan, uint8_t* out, int64_t nops);

/* Many independent pairs.  seqs0/seqs1 are concatenations; off*[k], len*[k] locate pair k
 * (len0[k] >= len1[k], caller swaps).  Outputs are arrays of npairs entries.  When
 * want_trace != 0, pair k owns the slot ops + ops_off[k] of capacity len0[k]+len1[k]; its
 * alen[k] alignment columns are written, in reading order, into the LAST alen[k] bytes of
 * that slot, one byte each (1 = both bases, 0 = s0 base against a gap, 2 = gap against an
 * s1 base).  ops/ops_off/alen may be NULL for score-only.  kernel_ms (may be NULL) receives
 * the CUDA-event time of the fill (+ traceback) kernels alone. */
int  tsta_psa_batch(const tsta_params* p, int64_t npairs,
                    const char* seqs0, const int64_t* off0, const int32_t* len0,
                    const char* seqs1, const int64_t* off1, const int32_t* len1,
                    int want_trace, int32_t* maxsorce, int32_t* corner,
                    uint8_t* ops, const int64_t* ops_off, int32_t* alen,
                    float* kernel_ms);

/* The same with the moves PACKED four to a byte (always with traceback): a quarter of the bytes that
 * travel back over PCIe.  ops_off[k] still counts MOVES and must be a multiple of 4; move p of pair k's
 * slot (the alignment again occupies the last alen[k] of its len0+len1 positions) is bits 2(p & 3) .. +1 of
 * byte (ops_off[k] + p) >> 2 of ops2, same codes (1 / 0 / 2).  ops2 holds at least max(ops_off+len0+len1)/4
 * rounded up bytes. */
int  tsta_psa_batch_packed(const tsta_params* p, int64_t npairs,
                           const char* seqs0, const int64_t* off0, const int32_t* len0,
                           const char* seqs1, const int64_t* off1, const int32_t* len1,
                           int32_t* maxsorce, int32_t* corner,
                           uint8_t* ops2, const int64_t* ops_off, int32_t* alen,
                           float* kernel_ms);

/* The 2-bit wire format for ACGT batches (a store that keeps reads packed, or a caller that packs once and aligns
 * many times): bases0 / bases1 hold four bases per byte, base i of a stream in bits 2(i & 3) .. +1 of byte i >> 2,
 * codes 0..3 = A C G T; off*[k] and len*[k] count BASES.  A quarter of the read bytes cross PCIe; the device expands
 * them.  Results are those of tsta_psa_batch on the expanded reads; with want_trace != 0 the moves come back packed
 * as in tsta_psa_batch_packed (ops_off multiples of 4). */
int  tsta_psa_batch_2bit(const tsta_params* p, int64_t npairs,
                         const uint8_t* bases0, const int64_t* off0, const int32_t* len0,
                         const uint8_t* bases1, const int64_t* off1, const int32_t* len1,
                         int want_trace, int32_t* maxsorce, int32_t* corner,
                         uint8_t* ops2, const int64_t* ops_off, int32_t* alen,
                         float* kernel_ms);

/* ------------------------------------------------------------------------------------ */
/* Sequence-to-graph alignment (TSTA_msa)                                                */
/* ------------------------------------------------------------------------------------ */

/* One traceback step as node_fuse() consumes it (msa/c-t-simd.c:751-956): the DP cell and
 * what its effective `source` said there.  kind: 0 vertical to predecessor `pred`, 1 match and
 * 2 mismatch (diagonal to predecessor `pred`), 3 horizontal (the read base gets a new node). */
typedef struct tsta_poa_step {
    int32_t row;      /* DP row = index into the caller's topological order (p->sort[]) */
    int32_t col;      /* read position */
    uint8_t kind;
    uint8_t pred;     /* index into the row's predecessor list */
    uint16_t reserved;
} tsta_poa_step;

typedef struct tsta_poa tsta_poa;   /* device buffers reused across the reads of one MSA job */
int  tsta_poa_create(tsta_poa** ctx);
void tsta_poa_destroy(tsta_poa* ctx);

/* Fill the DP of one read against the graph (replaces msa/c-t-simd.c:1003-1043).  The graph
 * comes in CSR form in ROW ORDER (= p->sort[] order): row r has predecessors
 * pred_row[row_ptr[r] .. row_ptr[r+1]) -- ROW indices, in the node's pre[] order (ties pick
 * the lowest index); a head (in == 0) has the single entry -1.  base[r] is the node's base,
 * hneg1[r] = H(r,-1) (O+E for heads, else max over predecessors + E: c-t-simd.c:187,196-200).
 * L is the reference's tile edge (SIMD lanes x -W), observable only through the late-head
 * quirk.  lastsorce[r] receives H(r, len-1) for every row (the reference
 * keeps it for sinks, c-t-simd.c:577-578). */
int  tsta_poa_fill(tsta_poa* ctx, const tsta_params* p, int L, int nrows,
                   const int32_t* row_ptr, const int32_t* pred_row, const uint8_t* base, const int32_t* hneg1,
                   const char* read, int len, int32_t* lastsorce, float* kernel_ms);
/* Device-resident graph (the reference re-sorts and re-reads the whole graph for every read:
 * msa/msa.c:132-133, msa/topo.c:243-260; here the graph stays in HBM between reads).  Nodes are
 * identified by stable ids 0..nnodes-1.  After a read has been fused the caller sends only the
 * records of the nodes whose base / predecessor list changed (or that are new): base_in = base |
 * in-degree << 8 (in-degree 0: a head), pre[] = predecessor NODE ids in pre[] order.  Nodes with at
 * most two predecessors travel as 16-byte records, up to eight as 40-byte records; a node with more
 * than eight predecessors cannot be resident (use tsta_poa_fill for that job). */
typedef struct tsta_poa_node2 { int32_t id; uint32_t base_in; int32_t pre[2]; } tsta_poa_node2;
typedef struct tsta_poa_node8 { int32_t id; uint32_t base_in; int32_t pre[8]; } tsta_poa_node8;
int  tsta_poa_graph_update(tsta_poa* ctx, int nnodes, int n_narrow, const tsta_poa_node2* narrow,
                           int n_wide, const tsta_poa_node8* wide);
/* tsta_poa_fill on the resident graph: sort[r] = id of the node in row r of the current topological
 * order (it crosses PCIe as 3 bytes per row; that, the changed records and the read are all that is copied).  Row descriptors, the
 * predecessor table in row indices and the spill slots are built on the device; H(r,-1) is computed
 * by the fill itself, so hneg1 may be NULL unless the read runs on the skewed kernel (it contains
 * 'N'), in which case it is required. */
int  tsta_poa_fill_resident(tsta_poa* ctx, const tsta_params* p, int L, int nrows, const int32_t* sort,
                            const int32_t* hneg1, const char* read, int len, int32_t* lastsorce, float* kernel_ms);
/* Bytes of graph tables and reads this context has copied to the device so far (TSTA_MSA_TIMING). */
long long tsta_poa_h2d_bytes(const tsta_poa* ctx);
/* Walk the traceback of the last fill from (start_row, len-1) with the reference's forced
 * continuation rules (c-t-simd.c:767-769, 776-782); path must hold nrows+len steps. */
int  tsta_poa_walk(tsta_poa* ctx, int start_row, tsta_poa_step* path, int* nsteps, float* kernel_ms);
/* Parity helper: the packed trace of the last fill, nrows x len uint16 (see csrc/poa_fill.cuh). */
int  tsta_poa_debug_trace(tsta_poa* ctx, uint16_t* out_rows_by_len);

/* Page-locked host memory for large inputs/outputs (batches): copies from/to such buffers run
 * at full PCIe rate and asynchronously; ordinary malloc'd buffers work too, only slower. */
void* tsta_host_alloc(size_t bytes);   /* NULL on failure */
void  tsta_host_free(void* p);

/* ------------------------------------------------------------------------------------ */
/* Measurement helpers                                                                   */
/* ------------------------------------------------------------------------------------ */

/* Dependency-free VIADDMNMX loop on every SM: measured int32 DPX lane-ops per second
 * (the roofline denominator of SURVEY.md 8d; B200's DPX rate is not documented). */
int  tsta_measure_dpx_peak(double* lane_ops_per_s, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* TSTA_CUDA_H */
