// psa_chase.cuh -- PARALLEL traceback over the tiles of the checkpointed mode (psa_fill16.cuh).
//
// The walk of reference trace() (psa/psa.c:435-463) is n+m dependent steps; one thread needs ~80
// cycles per step, which made the walk as long as the whole fill.  Here the serial part shrinks
// to ONE table lookup per tile the path crosses (~1300 for a 100 kbp pair):
//
//   1. psa_tile_summary_kernel (one warp per corridor tile, all tiles in parallel): for every cell
//      of the tile and every way the walk can arrive there (freely / by a left move / by an up
//      move -- the 3-state automaton of oracle/psa_oracle.c) the cell OUTSIDE the tile where the
//      walk leaves it, by the recurrence
//          E3(x) = tag(x) == diagonal ? D : tag(x) == left ? L : U        (arrived freely)
//          E2(x) = wf(x) ? L : E3(x)                                       (arrived by a left move)
//          E1(x) = we(x) ? U : E3(x)                                       (arrived by an up move)
//      where D / L / U are E3 / E2 / E1 of the diagonal / left / upper neighbour, or that
//      neighbour itself (encoded) when it lies outside the tile.  The recurrence runs in the same
//      skewed lane order as the fill, so the direction words are read exactly as they were written.
//      Kept: the cells a walk can ENTER a tile through -- its last two steps and its last column.
//   2. psa_chase_kernel (one thread): from the start cell, tile after tile: look up where the walk
//      leaves the tile, note the entry (a "visit"), continue in the tile it enters.
//   3. psa_visit_walk_kernel (one thread per visit): the actual moves inside each visited tile.
//   4. psa_visit_compact_kernel: visits' moves -> one contiguous list in walk order.
//
// A visit into a tile that was not re-run (outside the corridor) ends the chase with status 1 and
// the host continues in a new corridor (tsta_psa.cu), exactly like the serial walker.
#pragma once
#include "common.cuh"

namespace tsta {

template <int C>
struct PsaChaseGeom {
    static constexpr int kWords = C > 4 ? 2 : 1;            // direction words per lane and step
    static constexpr int kBandCols = 64 * C;
    static constexpr int kRowWords = 3 * C * 64;            // entries of the last step: [arrival][column of the lane][virtual lane]
    // a walk enters a tile through its last step (from the tile below), through the last column of every lane in
    // the step before (a diagonal move out of the next lane's first row below), or through the band's last column
    static constexpr int kSumWords = kRowWords + 64 + kTileSteps * 3;   // step 127; step 126: [virtual lane]; last column [step][arrival]
    static constexpr int kVisitMoves = kTileSteps + kBandCols;         // most moves one visit can make
};

struct PsaChaseArgs {
    const uint8_t* dirs;         // tiles' direction words (psa_retrace16_kernel)
    long long tile_bytes;
    uint32_t* summary;           // [slot][kSumWords]
    const PsaTileRow* rows;      // tile table
    int nrows;
    int n, m;
    int start_i, start_j, start_arr, start_nops;
    int4* visits;                // (i, j, arrival, -) of every tile entry, in walk order
    int max_visits;
    int* nvisits;
    int* counts;                 // moves of every visit
    int* offsets;                // scratch of the compaction
    uint8_t* slots;              // [max_visits][kVisitMoves]
    uint8_t* ops;                // all moves, walk order (tags: 3 diagonal, 2 left, 1 up)
    int* out;                    // [0] moves so far, [1] i, [2] j, [3] arrival (0 = free), [5] status
    uint32_t* start_exit;        // [3] where a walk from the START cell leaves its tile, per arrival (written by the summary kernel)
};

// The cell a walk leaves a tile to: offsets from the tile's corner, so that 32 bits are plenty.
__device__ __forceinline__ uint32_t psa_exit_code(long long i, long long j, int arr, long long t0, long long jbase)
{
    return 0x80000000u | ((uint32_t)(i - (t0 - 64) + 1) << 12) | ((uint32_t)(j - jbase + 1) << 2) | (uint32_t)arr;
}

template <int C>
__device__ __forceinline__ unsigned psa_tile_code(const uint32_t* tile, int s, int v, int c)
{
    using G = PsaChaseGeom<C>;
    const uint32_t w = __ldcg(tile + (s * 32 + (v & 31)) * G::kWords + (G::kWords == 2 ? (v >> 5) : 0));
    return (w >> (4 * c + (G::kWords == 1 ? 16 * (v >> 5) : 0))) & 15u;
}

// ---------------------------------------------------------------------------------------
// 1. tile summaries
// ---------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(32) psa_tile_summary_kernel(const PsaChaseArgs a)
{
    using G = PsaChaseGeom<C>;
    const int lane = threadIdx.x;
    const int k = a.nrows - (int)gridDim.y + (int)blockIdx.y;    // top ranges first: the chase starts at the bottom, whose lines are then the freshest in L2
    const PsaTileRow row = a.rows[k];
    const int band = row.bmin + (int)blockIdx.x;
    if (band > row.bmax) return;
    const long long slot = row.slot0 + blockIdx.x;
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(a.dirs + slot * a.tile_bytes);
    uint32_t* sum = a.summary + slot * G::kSumWords;
    const long long t0 = (long long)k * kTileSteps, jbase = (long long)band * G::kBandCols;
    const int src_lane = (lane + 31) & 31;

    // two virtual lanes per thread (h = 0: lane, h = 1: lane + 32), C columns each
    uint32_t U[2][C], Dp[2][C], diag[2] = {0, 0}, out2[2] = {0, 0}, out3[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int v = lane + 32 * h;
        const long long i0 = t0 - v;                       // the lane's first row in this tile: everything above is outside
#pragma unroll
        for (int c = 0; c < C; c++) {
            const long long j = jbase + v * C + c;
            U[h][c] = psa_exit_code(i0 - 1, j, 1, t0, jbase);
            Dp[h][c] = psa_exit_code(i0 - 1, j, 3, t0, jbase);
        }
    }
    uint32_t w_next[G::kWords];
#pragma unroll
    for (int q = 0; q < G::kWords; q++) w_next[q] = __ldcs(tile + lane * G::kWords + q);

    for (int s = 0; s < kTileSteps; s++) {
        uint32_t w[G::kWords];
#pragma unroll
        for (int q = 0; q < G::kWords; q++) w[q] = w_next[q];
        if (s + 1 < kTileSteps) {
#pragma unroll
            for (int q = 0; q < G::kWords; q++) w_next[q] = __ldcs(tile + ((s + 1) * 32 + lane) * G::kWords + q);
        }
        // right edge of the previous step, one virtual lane to the left
        const uint32_t r2lo = __shfl_sync(kFullMask, out2[0], src_lane), r3lo = __shfl_sync(kFullMask, out3[0], src_lane);
        const uint32_t r2hi = __shfl_sync(kFullMask, out2[1], src_lane), r3hi = __shfl_sync(kFullMask, out3[1], src_lane);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int v = lane + 32 * h;
            const long long i = t0 + s - v, j0 = jbase + v * C;
            const uint32_t recv2 = (h == 1 && lane == 0) ? r2lo : (h == 0 ? r2lo : r2hi);
            const uint32_t recv3 = (h == 1 && lane == 0) ? r3lo : (h == 0 ? r3lo : r3hi);
            const bool band_edge = v == 0;
            const bool top = i == 0;                          // row -1 is outside the matrix (s == 0 was set up before the loop)
            uint32_t Lc = (band_edge || s == 0) ? psa_exit_code(i, j0 - 1, 2, t0, jbase) : recv2;
            const uint32_t Din = (band_edge || s <= 1 || top) ? psa_exit_code(i - 1, j0 - 1, 3, t0, jbase) : diag[h];
            diag[h] = recv3;
            const uint32_t codes = G::kWords == 2 ? w[h] : (h ? (w[0] >> 16) : (w[0] & 0xffffu));
            uint32_t dold = Din;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t code = (codes >> (4 * c)) & 15u;
                const uint32_t tag = code & 3u;
                const uint32_t Ui = top ? psa_exit_code(-1, j0 + c, 1, t0, jbase) : U[h][c];
                const uint32_t Di = c == 0 ? Din : (top ? psa_exit_code(-1, j0 + c - 1, 3, t0, jbase) : dold);
                const uint32_t E3 = tag == 3u ? Di : (tag == 2u ? Lc : Ui);
                const uint32_t E2 = (code & 4u) ? Lc : E3;
                const uint32_t E1 = (code & 8u) ? Ui : E3;
                dold = Dp[h][c];
                Dp[h][c] = E3;
                U[h][c] = E1;
                Lc = E2;
                if (s == kTileSteps - 1) {
                    uint32_t* e = sum + c * 64 + v;
                    e[0] = E1; e[C * 64] = E2; e[2 * C * 64] = E3;
                }
                if (s == kTileSteps - 2 && c == C - 1) sum[G::kRowWords + v] = E3;
                if (i == a.start_i && j0 + c == a.start_j) { a.start_exit[0] = E1; a.start_exit[1] = E2; a.start_exit[2] = E3; }
                if (c == C - 1) {
                    out2[h] = E2;
                    out3[h] = E3;
                    if (h == 1 && lane == 31) {
                        uint32_t* e = sum + G::kRowWords + 64 + s * 3;
                        e[0] = E1; e[1] = E2; e[2] = E3;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// walking inside ONE tile (start of a chase; every visit)
// ---------------------------------------------------------------------------------------
struct PsaCell {
    int i, j, arr;      // arrival: 3 free (start / after a diagonal move), 2 by a left move, 1 by an up move
};

// Locates the tile of a cell.  Returns false when its direction words do not exist.
template <int C>
__device__ __forceinline__ bool psa_locate(const PsaChaseArgs& a, int i, int j, int& band, int& k, int& s, int& v, int& c, long long& slot)
{
    using G = PsaChaseGeom<C>;
    band = j / G::kBandCols;
    const int r = j - band * G::kBandCols;
    v = r / C;
    c = r - v * C;
    const int t = i + v;
    k = t / kTileSteps;
    s = t - k * kTileSteps;
    const PsaTileRow row = a.rows[k];
    if (band < row.bmin || band > row.bmax) return false;
    slot = row.slot0 + (band - row.bmin);
    return true;
}

// Walks from `x` (inside the tile at `slot`, position s / v / c) until the walk leaves the tile or the matrix.
// Returns the number of moves (written to `out` when it is not NULL); `x` becomes the cell left to.
// bad is set when a code without a back tag is met (corrupt tile).
template <int C>
__device__ __forceinline__ int psa_tile_walk(const PsaChaseArgs& a, long long slot, int s, int v, int c, PsaCell& x, uint8_t* out, bool& bad)
{
    const uint32_t* tile = reinterpret_cast<const uint32_t*>(a.dirs + slot * a.tile_bytes);
    int n = 0, i = x.i, j = x.j, arr = x.arr;
    while (i >= 0 && j >= 0 && s >= 0 && v >= 0) {
        const unsigned code = psa_tile_code<C>(tile, s, v, c);
        const unsigned tag = code & 3u;
        if (tag == 0u) { bad = true; break; }
        const bool forced = (arr == 2 && (code & 4u)) || (arr == 1 && (code & 8u));
        const int mv = forced ? arr : (int)tag;
        if (out) out[n] = (uint8_t)mv;
        n++;
        if (mv & 2) {                 // left or diagonal: one column to the left
            j--;
            if (--c < 0) { c = C - 1; v--; s--; }
        }
        if (mv & 1) {                 // up or diagonal: one row up
            i--;
            s--;
        }
        arr = mv;
    }
    x.i = i; x.j = j; x.arr = arr;
    return n;
}

// ---------------------------------------------------------------------------------------
// 2. the chase: one lookup per tile
// ---------------------------------------------------------------------------------------
constexpr int kChaseRowsSmem = 2560;       // tile-table rows kept in shared memory (40 KB): pairs of up to ~320 kbp rows

template <int C>
__global__ void __launch_bounds__(32) psa_chase_kernel(const PsaChaseArgs a0)
{
    using G = PsaChaseGeom<C>;
    // every visit is a chain of dependent loads (table row -> summary word): the table row comes from shared memory
    __shared__ PsaTileRow s_rows[kChaseRowsSmem];
    PsaChaseArgs a = a0;
    if (a.nrows <= kChaseRowsSmem) {
        for (int q = threadIdx.x; q < a.nrows; q += 32) s_rows[q] = a0.rows[q];
        __syncwarp();
        a.rows = s_rows;
    }
    if (threadIdx.x != 0) return;
    PsaCell x{a.start_i, a.start_j, a.start_arr ? a.start_arr : 3};
    int nv = 0, status = 0;
    bool first = true;
    while (x.i >= 0 && x.j >= 0) {
        int band, k, s, v, c;
        long long slot;
        if (!psa_locate<C>(a, x.i, x.j, band, k, s, v, c, slot)) { status = 1; break; }
        if (nv >= a.max_visits) { status = 2; break; }
        a.visits[nv++] = make_int4(x.i, x.j, x.arr, 0);
        const uint32_t* sum = a.summary + slot * G::kSumWords;
        uint32_t e;
        if (first) { e = __ldcg(a.start_exit + (x.arr - 1)); first = false; }     // the start cell: kept by the summary kernel
        else if (s == kTileSteps - 1) e = __ldcg(sum + (x.arr - 1) * (C * 64) + c * 64 + v);
        else if (v == 63 && c == C - 1) e = __ldcg(sum + G::kRowWords + 64 + s * 3 + (x.arr - 1));
        else if (s == kTileSteps - 2 && c == C - 1 && x.arr == 3) e = __ldcg(sum + G::kRowWords + v);
        else {
            // not a cell whose exit was kept (the start of the alignment): walk this tile step by step
            bool bad = false;
            psa_tile_walk<C>(a, slot, s, v, c, x, nullptr, bad);
            if (bad) { status = 3; nv--; break; }
            continue;
        }
        if (!(e & 0x80000000u)) { status = 3; nv--; break; }          // (never written: corrupt summary)
        x.i = (int)((e >> 12) & 0x7ffffu) - 1 + (k * kTileSteps - 64);
        x.j = (int)((e >> 2) & 0x3ffu) - 1 + band * G::kBandCols;
        x.arr = (int)(e & 3u);
    }
    *a.nvisits = nv;
    a.out[1] = x.i;
    a.out[2] = x.j;
    a.out[3] = x.arr == 3 ? 0 : x.arr;
    a.out[5] = status;
}

// ---------------------------------------------------------------------------------------
// 3. the moves of every visit
// ---------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(64) psa_visit_walk_kernel(const PsaChaseArgs a)
{
    using G = PsaChaseGeom<C>;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= *a.nvisits) return;
    const int4 vis = a.visits[q];
    PsaCell x{vis.x, vis.y, vis.z};
    int band, k, s, v, c;
    long long slot;
    int n = 0;
    if (psa_locate<C>(a, x.i, x.j, band, k, s, v, c, slot)) {
        bool bad = false;
        n = psa_tile_walk<C>(a, slot, s, v, c, x, a.slots + (size_t)q * G::kVisitMoves, bad);
        if (bad) atomicExch(a.out + 5, 3);
    }
    a.counts[q] = n;
}

// ---------------------------------------------------------------------------------------
// 4. visits' moves -> one list (single CTA: an exclusive scan of the counts, then the copies)
// ---------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(1024) psa_visit_compact_kernel(const PsaChaseArgs a)
{
    using G = PsaChaseGeom<C>;
    __shared__ int s_part[1024];
    const int nv = *a.nvisits, tid = threadIdx.x;
    const int per = (nv + 1023) / 1024;
    int local = 0;
    for (int q = tid * per; q < min(nv, (tid + 1) * per); q++) local += a.counts[q];
    s_part[tid] = local;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                       // inclusive scan
        const int add = tid >= o ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += add;
        __syncthreads();
    }
    int off = s_part[tid] - local;
    for (int q = tid * per; q < min(nv, (tid + 1) * per); q++) { a.offsets[q] = off; off += a.counts[q]; }
    const int total = s_part[1023];
    __syncthreads();
    __threadfence_block();
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < nv; q += 32) {
        const int n = a.counts[q];
        const uint8_t* src = a.slots + (size_t)q * G::kVisitMoves;
        uint8_t* dst = a.ops + a.start_nops + a.offsets[q];
        for (int e = lane; e < n; e += 32) dst[e] = src[e];
    }
    if (tid == 0) a.out[0] = a.start_nops + total;
}

}  // namespace tsta
