// psa_trace.cuh -- traceback over the packed 4-bit direction matrix (one long pair).
//
// Replaces reference trace() (psa/psa.c:435-463, the walk).  The reference rewrites
// `back` ahead of itself to force gap continuation; with the cF/cE bits of the packed code
// the same walk is a 3-state automaton (see oracle/psa_oracle.c: tsta_oracle_psa).
//
// The walk is inherently serial (n+m dependent steps), so one thread walks while the whole
// CTA stages the window of the matrix around the current position into shared memory with
// coalesced 128-bit loads: a step then costs one shared-memory load instead of one HBM
// round trip.  Output: the moves in walk order (end of the alignment first), one byte each
// (1 diagonal, 0 left = gap in s1, 2 up = gap in s0); the host turns them into the two
// '-'-padded strings (reference psa/psa.c:467-510).
#pragma once
#include "common.cuh"

namespace tsta {

constexpr int kTraceRows = 128;        // window height (rows)
constexpr int kTraceBytes = 256;       // window width in bytes (= 512 columns)
constexpr int kTraceThreads = 32 + 256;   // warp 0 walks, 8 warps stage the next window
constexpr int kTraceWindowBytes = kTraceRows * kTraceBytes;   // 32 KB per buffer
constexpr int kTraceApron = 8 * kTraceBytes;                  // readable rows in front of a buffer (a block may reach above the window)
constexpr int kTraceBufStride = kTraceWindowBytes + kTraceApron;
constexpr int kTraceSmemBytes = 2 * kTraceBufStride;
constexpr int kTraceBlkRows = 8;       // the walker keeps 8 rows x 16 columns of codes in registers

struct PsaTraceArgs {
    const uint8_t* dirs;
    long long dirs_stride;   // bytes per row
    int n, m;
    uint8_t* ops;            // [n + m] moves in walk order
    int* out;                // out[0] = number of moves written, out[1] = final i, out[2] = final j, out[3] = final forced state
    int start_i, start_j;    // where the walk starts ((m-1, n-1) for a whole pair; a slice continues its right neighbour's walk)
    int start_forced;        // 0 free, 2 forced left, 1 forced up (tag form)
    int enc;                 // back code: 0 = (1 diagonal, 0 left, 2 up), 1 = (3 diagonal, 2 left, 1 up)
    long long* dbg;          // optional: walker cycles, total cycles, windows, prefetch hits
    // skewed layout (psa_trace_skew_kernel): bytes of one band's [step][lane] region, steps per band, bands
    long long skew_band_bytes;
    long long skew_steps;
    int skew_bands;
    // checkpointed traceback: direction words exist for the tiles of this table only (common.cuh: PsaTileRow);
    // NULL = the whole matrix is there, band after band (skew_band_bytes)
    const PsaTileRow* tile_rows;
    long long tile_bytes;
    int start_nops;          // moves already written by an earlier walk of the same alignment
    // out[5] (status): 0 = the walk reached row -1 or column -1; 1 = it ran into cells without direction words:
    // out[0..3] then describe the entry of the last window it could stage, from where the walk can be resumed
};

// `back` bits of a packed code -> move (1 diagonal, 0 left, 2 up) for either encoding
__device__ __forceinline__ int psa_back_move(int c, int enc) { return enc ? ((0x48 >> (2 * (c & 3))) & 3) : (c & 3); }

// Window placement: rows [R0, R0+128), columns [C0, C0+512) with the cell (i, j) in the last
// rows / the last 32 columns.  C0 is a multiple of 32 columns (16-byte global units).
__device__ __forceinline__ void psa_trace_place(int i, int j, int& R0, int& C0)
{
    R0 = max(0, i + 1 - kTraceRows);
    C0 = max(0, ((j + 32) & ~31) - 2 * kTraceBytes);
}

// Stages one window row-major into shared memory (coalesced 16-byte loads, conflict-free 16-byte
// stores).  All loads of a thread are issued before the first is used, so a window costs about one
// HBM round trip.  Codes are normalised to tag form (3 diagonal, 2 left, 1 up) on the way.
__device__ __forceinline__ void psa_trace_stage(const PsaTraceArgs& a, uint8_t* buf, int R0, int C0, int tid, int nthreads)
{
    constexpr int kUnitsPerRow = kTraceBytes / 16;
    constexpr int kUnits = kTraceRows * kUnitsPerRow;     // 2048
    constexpr int kBatch = 4;
    for (int u0 = tid; u0 < kUnits; u0 += kBatch * nthreads) {
        uint4 v[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; b++) {
            const int u = u0 + b * nthreads;
            const int r = u / kUnitsPerRow, q = u % kUnitsPerRow;
            const int row = R0 + r;
            const long long off = (long long)(C0 >> 1) + q * 16;
            v[b] = make_uint4(0, 0, 0, 0);
            if (u < kUnits && row < a.m && off < a.dirs_stride)
                v[b] = __ldcs(reinterpret_cast<const uint4*>(a.dirs + (long long)row * a.dirs_stride + off));
        }
#pragma unroll
        for (int b = 0; b < kBatch; b++) {
            const int u = u0 + b * nthreads;
            if (u >= kUnits) break;
            uint32_t w[4] = {v[b].x, v[b].y, v[b].z, v[b].w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (!a.enc)   // int32 kernels: back = 1 diagonal / 0 left / 2 up  ->  tags 3 / 2 / 1
                    w[k] = (w[k] & 0xccccccccu) | (~w[k] & 0x22222222u) | (((w[k] >> 1) | w[k]) & 0x11111111u);
                // a tag of 0 never comes out of a fill; turn it into a diagonal so that even a
                // corrupt matrix cannot stall the walk
                w[k] |= (~(w[k] | (w[k] >> 1)) & 0x11111111u) * 3u;
            }
            *reinterpret_cast<uint4*>(buf + (size_t)u * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__device__ __forceinline__ unsigned trace_bsel(unsigned m, unsigned a, unsigned b)   // m ? a : b, bitwise
{
    unsigned d;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(m), "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint2 trace_lds_u2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// One thread walks (3-state automaton of oracle/psa_oracle.c) while the other warps already
// stage the window the path is expected to enter next: the walk is serial by nature
// (n+m dependent steps), so everything else is kept off its critical path.
static __global__ void __launch_bounds__(kTraceThreads) psa_trace_kernel(const PsaTraceArgs a)
{
    extern __shared__ __align__(16) uint8_t trace_smem[];      // two window buffers, each behind an apron
    __shared__ int s_i, s_j, s_forced, s_nops;
    const int tid = threadIdx.x;

    if (tid == 0) {
        s_i = a.start_i;
        s_j = a.start_j;
        s_forced = a.start_forced;
        s_nops = 0;
    }
    __syncthreads();
    int cur = 0, R0, C0;
    long long dbg_walk = 0, dbg_total0 = clock64();
    int dbg_windows = 0, dbg_hits = 0;
    psa_trace_place(s_i, s_j, R0, C0);
    psa_trace_stage(a, trace_smem + kTraceApron, R0, C0, tid, kTraceThreads);
    __syncthreads();

    while (true) {
        const int i0 = s_i, j0 = s_j;
        if (i0 < 0 || j0 < 0) break;
        // where the path is expected to leave this window: through the top, about one column per row
        int NR0, NC0;
        psa_trace_place(R0 - 1, max(0, j0 - (i0 - R0 + 1) + 2 * kTraceBytes / 4), NR0, NC0);
        if (tid >= 32) {
            if (R0 > 0) psa_trace_stage(a, trace_smem + kTraceApron + (cur ^ 1) * kTraceBufStride, NR0, NC0, tid - 32, kTraceThreads - 32);
        } else if (tid == 0) {
            const long long dbg_t0 = clock64();
            dbg_windows++;
            const uint8_t* win = trace_smem + kTraceApron + cur * kTraceBufStride;
            // Moves in tag form: 3 diagonal, 2 left, 1 up; forced state 0 free, 2 / 1 forced left / up.
            // The walk is ONE dependent chain, so the inner loop avoids predicates (a compare
            // feeding a select costs ~13 cycles): the walker keeps 8 rows x 16 columns of codes
            // in registers with the CURRENT row always in r[0] -- an up move shifts the rows
            // down by bit-selects, a left move lowers the shift count -- the forced state is an
            // AND mask and an OR value for the next code, and its successor comes out of a
            // 64-bit shift LUT.
            int i = i0, j = j0, nops = s_nops;
            const unsigned st0 = (unsigned)s_forced;
            unsigned fm = st0 ? 0u : 3u, fv = st0;
            // next (mask << 2 | value) by index cFcE * 4 + move:  free (12) unless the move was a
            // gap and the cell says "keep extending" (left & cF -> 2, up & cE -> 1)
            const unsigned long long kNext = 0xC21CC2CCCC1CCCCCull;
            while (i >= R0 && j >= C0) {
                const int j16 = j & ~15;
                const uint8_t* p = win + (i - R0) * kTraceBytes + ((j16 - C0) >> 1);
                unsigned rl[kTraceBlkRows], rh[kTraceBlkRows];
#pragma unroll
                for (int k = 0; k < kTraceBlkRows; k++) {       // rows i, i-1, ... (the apron absorbs rows above the window)
                    const uint2 v = *reinterpret_cast<const uint2*>(p - k * kTraceBytes);
                    rl[k] = v.x; rh[k] = v.y;
                }
                const int rows = min(kTraceBlkRows, i - R0 + 1);
                int rem = rows, pos = 4 * (j - j16);
                do {
                    const unsigned c = (unsigned)((((unsigned long long)rh[0] << 32) | rl[0]) >> pos);   // low nibble = code
                    const unsigned mvt = (c & fm) | fv;                          // forced, or the cell's own back tag
                    const unsigned q16 = ((c << 3) & 32u) | ((c << 1) & 16u);    // (cF << 1 | cE) * 16
                    const unsigned e = (unsigned)(kNext >> (q16 + 4u * mvt));
                    fm = (e >> 2) & 3u;
                    fv = e & 3u;
                    a.ops[nops++] = (uint8_t)((0x48u >> (2 * mvt)) & 3u);        // 1 diagonal, 0 left, 2 up
                    const unsigned up = mvt & 1u, um = 0u - up;
#pragma unroll
                    for (int k = 0; k + 1 < kTraceBlkRows; k++) {
                        rl[k] = trace_bsel(um, rl[k + 1], rl[k]);
                        rh[k] = trace_bsel(um, rh[k + 1], rh[k]);
                    }
                    rem -= (int)up;
                    pos -= 4 * (int)(mvt >> 1);
                } while (((rem - 1) | pos) >= 0);
                i -= rows - rem;
                j = j16 + (pos >> 2);
            }
            const int forced = (int)fv;
            s_i = i; s_j = j; s_forced = forced; s_nops = nops;
            dbg_walk += clock64() - dbg_t0;
        }
        __syncthreads();
        const int i1 = s_i, j1 = s_j;
        if (i1 < 0 || j1 < 0) break;
        if (R0 > 0 && i1 >= NR0 && i1 < NR0 + kTraceRows && j1 >= NC0 && j1 < NC0 + 2 * kTraceBytes) {
            cur ^= 1; R0 = NR0; C0 = NC0;                       // the prefetched window fits
            dbg_hits++;
        } else {
            psa_trace_place(i1, j1, R0, C0);
            psa_trace_stage(a, trace_smem + kTraceApron + cur * kTraceBufStride, R0, C0, tid, kTraceThreads);
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.out[0] = s_nops;
        a.out[1] = s_i;
        a.out[2] = s_j;
        a.out[3] = s_forced;
        if (a.dbg) {
            a.dbg[0] = dbg_walk; a.dbg[1] = clock64() - dbg_total0; a.dbg[2] = dbg_windows; a.dbg[3] = dbg_hits;
        }
    }
}


// =========================================================================================
// Skewed direction layout (psa_fill16.cuh, TM 2): dirs[band][step][lane] words, row r of
// virtual lane v in the word of step r + v; cell code = back tag | wf << 2 | we << 3, where
// wf / we say "a walk that ENTERS this cell by a left / up move keeps going" (the same
// automaton as above, read at the entered cell instead of the cell that was left).
//
// Staging a window is two phases: (1) the stripe of steps of every band under the window is
// copied as it lies in HBM (contiguous, coalesced 16-byte loads) into a raw shared buffer,
// (2) the window is assembled row-major, 4 bits per cell, from the raw words: one task =
// one row x 8 virtual lanes = 8C columns = C whole 32-bit words, so nothing is read-modify-
// written.  The walker is the only consumer, and it keeps just two rows in registers: the
// current one and the one above (the row two above is re-loaded every step while the step's own chain
// runs; an up move shifts the rows by bit-selects).  Moves leave as one tag byte each.
// =========================================================================================
template <int C>
struct PsaSkewGeom {
    static constexpr int kBandCols = 64 * C;
    static constexpr int kGroupCols = 8 * C;                                  // one assembly task
    static constexpr int kAlign = (kGroupCols % 16 == 0) ? kGroupCols : 2 * kGroupCols;   // window start granularity (columns)
    static constexpr int kWinCols = (2 * kTraceBytes / kAlign) * kAlign;      // 512 or 480
    static constexpr int kWords = C > 4 ? 2 : 1;                              // words per lane and step
    static constexpr int kLineBytes = 128 * kWords;
    static constexpr int kMaxBands = (kBandCols - kAlign + kWinCols + kBandCols - 1) / kBandCols;
    static constexpr int kRawLines = (kTraceRows - 1) * kMaxBands + kWinCols / C;
    static constexpr int kRawBytes = kRawLines * kLineBytes;
    static constexpr int kSmemBytes = 2 * kTraceBufStride + kRawBytes;
};

// Direction words of (band b, step t): one line of 32 lanes, or NULL when they do not exist.
template <int C>
__device__ __forceinline__ const uint4* psa_skew_line(const PsaTraceArgs& a, int b, long long t)
{
    using G = PsaSkewGeom<C>;
    if (b >= a.skew_bands || t >= a.skew_steps) return nullptr;
    if (!a.tile_rows) return reinterpret_cast<const uint4*>(a.dirs + (long long)b * a.skew_band_bytes + t * G::kLineBytes);
    const PsaTileRow r = a.tile_rows[t / kTileSteps];
    if (b < r.bmin || b > r.bmax) return nullptr;
    return reinterpret_cast<const uint4*>(a.dirs + (r.slot0 + (b - r.bmin)) * a.tile_bytes + (t % kTileSteps) * G::kLineBytes);
}

template <int C>
__device__ __forceinline__ void psa_skew_place(int i, int j, int& R0, int& C0)
{
    using G = PsaSkewGeom<C>;
    R0 = max(0, i + 1 - kTraceRows);
    C0 = max(0, (j / G::kAlign + 1) * G::kAlign - G::kWinCols);
}

// Stages rows [R0, R0+128) x columns [C0, C0+kWinCols) into `buf` (row pitch kTraceBytes).
// Called by `nthreads` threads that share named barrier `bar_id`.
template <int C>
__device__ __forceinline__ void psa_skew_stage(const PsaTraceArgs& a, uint8_t* buf, uint8_t* raw, int R0, int C0,
                                               int tid, int nthreads, int bar_id)
{
    using G = PsaSkewGeom<C>;
    const int b0 = C0 / G::kBandCols;
    // ---- phase 1: raw stripes.  Band b of the window needs virtual lanes [vlo, vhi] and steps R0+vlo .. R0+127+vhi
    int line0 = 0;
#pragma unroll
    for (int bb = 0; bb < G::kMaxBands; bb++) {
        const int b = b0 + bb;
        const int cb0 = max(C0, b * G::kBandCols), cb1 = min(C0 + G::kWinCols, (b + 1) * G::kBandCols);
        if (cb0 >= cb1) break;
        const int vlo = (cb0 - b * G::kBandCols) / C, vhi = (cb1 - 1 - b * G::kBandCols) / C;
        const int nlines = kTraceRows + vhi - vlo;
        const long long tlo = (long long)R0 + vlo;
        const int units = nlines * (G::kLineBytes / 16);
        uint4* dst = reinterpret_cast<uint4*>(raw + (size_t)line0 * G::kLineBytes);
        constexpr int kUnitsPerLine = G::kLineBytes / 16;
        // the stripe (at most 128 + 63 steps) touches at most three step ranges: their first lines are looked up
        // once, so that the loads below do not wait for a table entry each
        const long long k0 = tlo / kTileSteps;
        const uint4* first[3];
#pragma unroll
        for (int q = 0; q < 3; q++) first[q] = psa_skew_line<C>(a, b, (k0 + q) * kTileSteps);
        // all loads of a batch are issued before the first is stored: a stripe costs about one HBM round trip
        constexpr int kBatch = 6;
        for (int u0 = tid; u0 < units; u0 += kBatch * nthreads) {
            uint4 v[kBatch];
#pragma unroll
            for (int e = 0; e < kBatch; e++) {
                const int u = u0 + e * nthreads;
                const long long t = tlo + u / kUnitsPerLine;
                const int q = (int)(t / kTileSteps - k0);
                const uint4* src = q == 0 ? first[0] : (q == 1 ? first[1] : first[2]);
                // (cells without direction words read as code 0, which no fill produces: the walker stops on them)
                v[e] = make_uint4(0, 0, 0, 0);
                if (u < units && src && t < a.skew_steps) v[e] = __ldcs(src + (t % kTileSteps) * kUnitsPerLine + u % kUnitsPerLine);
            }
#pragma unroll
            for (int e = 0; e < kBatch; e++) {
                const int u = u0 + e * nthreads;
                if (u < units) dst[u] = v[e];
            }
        }
        line0 += nlines;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
    // ---- phase 2: assembly.  Task (row rho, group gamma): 8 virtual lanes -> C words of the window row
    constexpr int kGroups = G::kWinCols / G::kGroupCols;
    const uint32_t* raw32 = reinterpret_cast<const uint32_t*>(raw);
    for (int task = tid; task < kTraceRows * kGroups; task += nthreads) {
        const int rho = task / kGroups, gamma = task - rho * kGroups;
        const int col = C0 + gamma * G::kGroupCols;
        const int b = col / G::kBandCols;
        const int v0 = (col - b * G::kBandCols) / C;               // multiple of 8
        // raw line of (band b, row rho, virtual lane v): lines of the bands before b, then rho + v - vlo(b)
        int lbase = 0;
#pragma unroll
        for (int bb = 0; bb < G::kMaxBands - 1; bb++) {
            const int bq = b0 + bb;
            if (bq < b) {
                const int cb0 = max(C0, bq * G::kBandCols), cb1 = min(C0 + G::kWinCols, (bq + 1) * G::kBandCols);
                lbase += kTraceRows + (cb1 - 1 - bq * G::kBandCols) / C - (cb0 - bq * G::kBandCols) / C;
            }
        }
        const int vlo = (max(C0, b * G::kBandCols) - b * G::kBandCols) / C;
        unsigned long long acc = 0;
        int fill = 0, outw = 0;
        uint32_t* dstw = reinterpret_cast<uint32_t*>(buf + rho * kTraceBytes + (gamma * G::kGroupCols) / 2);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int v = v0 + e;
            const int line = lbase + rho + v - vlo;
            uint32_t w = raw32[(size_t)line * (32 * G::kWords) + (v & 31) * G::kWords + (G::kWords == 2 ? (v >> 5) : 0)];
            if (G::kWords == 1) w = (v >> 5) ? (w >> 16) : (w & 0xffffu);
            constexpr uint32_t kPiece = 4 * C < 32 ? (1u << ((4 * C) & 31)) - 1u : 0xffffffffu;   // 4C bits
            w &= kPiece;
            acc |= (unsigned long long)w << fill;
            fill += 4 * C;
            if (fill >= 32) {
                dstw[outw++] = (uint32_t)acc;
                acc >>= 32;
                fill -= 32;
            }
        }
    }
}

// The walker's warp has an SM sub-partition to itself: warps 4, 8, 12 (which would share warp 0's scheduler)
// only wait at the barriers, the other nine stage windows.
constexpr int kSkewWarps = 13;
constexpr int kSkewThreads = 32 * kSkewWarps;
constexpr int kSkewStagers = 32 * 9;

template <int C>
static __global__ void __launch_bounds__(kSkewThreads) psa_trace_skew_kernel(const PsaTraceArgs a)
{
    using G = PsaSkewGeom<C>;
    extern __shared__ __align__(16) uint8_t trace_smem[];      // two window buffers (each behind an apron), then the raw stripes
    __shared__ int s_i, s_j, s_arr, s_nops, s_status;
    uint8_t* const raw = trace_smem + 2 * kTraceBufStride;
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const bool stager = warp > 0 && (warp & 3) != 0;
    const int stid = (warp - 1 - (warp >> 2)) * 32 + (tid & 31);   // 0 .. kSkewStagers-1 over the staging warps

    if (tid == 0) {
        s_i = a.start_i;
        s_j = a.start_j;
        s_arr = a.start_forced ? a.start_forced : 3;           // how the walk entered the start cell: 3 free, 2 by a left move, 1 by an up move
        s_nops = a.start_nops;
        s_status = 0;
    }
    __syncthreads();
    int cur = 0, R0, C0;
    long long dbg_walk = 0, dbg_total0 = clock64();
    int dbg_windows = 0, dbg_hits = 0;
    psa_skew_place<C>(s_i, s_j, R0, C0);
    if (stager) psa_skew_stage<C>(a, trace_smem + kTraceApron, raw, R0, C0, stid, kSkewStagers, 1);
    __syncthreads();

    while (true) {
        const int i0 = s_i, j0 = s_j;
        if (i0 < 0 || j0 < 0) break;
        // where the path is expected to leave this window: through the top, about one column per row
        int NR0, NC0;
        psa_skew_place<C>(R0 - 1, max(0, j0 - (i0 - R0 + 1) + G::kWinCols / 4), NR0, NC0);
        if (stager) {
            if (R0 > 0) psa_skew_stage<C>(a, trace_smem + kTraceApron + (cur ^ 1) * kTraceBufStride, raw, NR0, NC0,
                                           stid, kSkewStagers, 1);
        } else if (tid == 0) {
            const long long dbg_t0 = clock64();
            dbg_windows++;
            const uint8_t* win = trace_smem + kTraceApron + cur * kTraceBufStride;
            int i = i0, j = j0;
            unsigned nops = (unsigned)s_nops;
            unsigned mvp = (unsigned)s_arr;                    // previous move (3 diagonal / start, 2 left, 1 up)
            const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
            unsigned sh = 4u - mvp, nf = 6u >> mvp;
            uint8_t* out = a.ops + nops;
            unsigned bad = 0;                                  // sign bit: a code without a back tag was read (no direction words there)
            while (i >= R0 && j >= C0) {
                const int j16 = j & ~15;
                const uint32_t colbase = win_s + (uint32_t)((j16 - C0) >> 1);
                int q = (i - R0) * kTraceBytes;                      // row offset inside the window; an up move subtracts one row
                uint2 cw = trace_lds_u2(colbase + q);
                uint2 u1 = trace_lds_u2(colbase + q - kTraceBytes);  // (rows above R0 are the apron: loaded, never used)
                int pos = 4 * (j - j16);
                const int q0 = q;
                // A step uses up at most one row and one column, so min(rows, columns) steps need no exit test at all:
                // the loop branch does not depend on the moves, and the only chain that spans iterations is
                // code -> forced? -> move -> row select / column shift.  The row two above is loaded at the top and
                // only used by the selects at the bottom, so its latency hides behind that chain.
                int safe = min(q / kTraceBytes, pos >> 2) + 1;
                while (safe > 0) {
#pragma unroll 2
                    for (int s = 0; s < safe; s++) {
                        const uint2 u2 = trace_lds_u2(colbase + q - 2 * kTraceBytes);
                        const unsigned c = (unsigned)((((unsigned long long)cw.y << 32) | cw.x) >> pos);   // low nibble = code
                        // forced: entered by a left move and wf (bit 2) set, or by an up move and we (bit 3) set
                        unsigned f, mv, um;
                        asm volatile("lop3.b32 %0, %1, %2, 1, 0x80;" : "=r"(f) : "r"(c >> sh), "r"(nf));        // (c >> sh) & nf & 1
                        const unsigned c3 = c & 3u;
                        bad |= c3 - 1u;
                        asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(mv) : "r"(f), "r"(mvp - c3), "r"(c3));   // f ? mvp : c3, no predicate
                        asm volatile("bfe.s32 %0, %1, 0, 1;" : "=r"(um) : "r"(mv));                               // 0 - (mv & 1): up / diagonal
                        out[s] = (uint8_t)mv;                        // tags; the host maps them to 1 / 0 / 2
                        cw.x = trace_bsel(um, u1.x, cw.x);
                        cw.y = trace_bsel(um, u1.y, cw.y);
                        u1.x = trace_bsel(um, u2.x, u1.x);
                        u1.y = trace_bsel(um, u2.y, u1.y);
                        q += (int)um * kTraceBytes;
                        asm volatile("mad.lo.s32 %0, %1, -2, %2;" : "=r"(pos) : "r"(mv & 2u), "r"(pos));
                        mvp = mv;
                        sh = 4u - mv;
                        nf = 6u >> mv;
                    }
                    out += safe;
                    safe = ((q | pos) < 0 || (int)bad < 0) ? 0 : min(q / kTraceBytes, pos >> 2) + 1;
                }
                if ((int)bad < 0) break;
                i -= (q0 - q) / kTraceBytes;
                j = j16 + (pos >> 2);
            }
            nops = (unsigned)(out - a.ops);
            if ((int)bad < 0) s_status = 1;                    // (s_i, s_j, s_arr, s_nops keep describing the entry of this window)
            else { s_i = i; s_j = j; s_arr = (int)mvp; s_nops = (int)nops; }
            dbg_walk += clock64() - dbg_t0;
        }
        __syncthreads();
        if (s_status) break;
        const int i1 = s_i, j1 = s_j;
        if (i1 < 0 || j1 < 0) break;
        if (R0 > 0 && i1 >= NR0 && i1 < NR0 + kTraceRows && j1 >= NC0 && j1 < NC0 + G::kWinCols) {
            cur ^= 1; R0 = NR0; C0 = NC0;                       // the prefetched window fits
            dbg_hits++;
        } else {
            psa_skew_place<C>(i1, j1, R0, C0);
            if (stager) psa_skew_stage<C>(a, trace_smem + kTraceApron + cur * kTraceBufStride, raw, R0, C0, stid, kSkewStagers, 1);
        }
        __syncthreads();
    }
    if (tid == 0) {
        a.out[0] = s_nops;
        a.out[1] = s_i;
        a.out[2] = s_j;
        a.out[3] = s_arr == 3 ? 0 : s_arr;
        a.out[5] = s_status;
        if (a.dbg) {
            a.dbg[0] = dbg_walk; a.dbg[1] = clock64() - dbg_total0; a.dbg[2] = dbg_windows; a.dbg[3] = dbg_hits;
        }
    }
}

}  // namespace tsta

namespace tsta {

// Batch traceback: one thread per pair, straight from global memory (millions of
// independent walks hide each other's latency).  The walk runs from the end of the
// alignment to its start, so moves are written from the END of pair k's slot
// (ops + ops_off[k], capacity len0+len1) backwards, INCLUDING the leftover moves of
// psa.c:497-510: the alignment ends up in reading order in the last alen[k] bytes.
struct PsaBatchTraceArgs {
    const uint8_t* dirs;
    const long long* dirs_off;   // [npairs]
    const int* len0;
    const int* len1;
    long long npairs;
    int cols_per_band;           // 32*C of the fill that produced dirs
    uint8_t* ops;
    const long long* ops_off;
    int* alen;
    int enc;                     // see PsaTraceArgs
};

// PACK: four moves per byte (move p of the slot in bits 2(p & 3) of byte (ops_off + p) >> 2; ops_off is a
// multiple of four): a quarter of the bytes that cross PCIe afterwards.
template <bool PACK>
static __global__ void __launch_bounds__(128) psa_batch_trace_kernel(const PsaBatchTraceArgs a)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.npairs) return;
    const int n = a.len0[k], m = a.len1[k];
    const long long stride = (long long)((n + a.cols_per_band - 1) / a.cols_per_band) * (a.cols_per_band / 2);
    const uint8_t* d = a.dirs + a.dirs_off[k];
    uint8_t* ops = a.ops + a.ops_off[k] + (n + m - 1);   // written backwards
    uint8_t* packed = a.ops + (a.ops_off[k] >> 2);
    int i = m - 1, j = n - 1, forced = -1, nops = 0;
    unsigned acc = 0;
    auto put = [&](int mv) {
        if (PACK) {
            const int p = n + m - 1 - nops;               // position in the slot, descending
            acc |= (unsigned)mv << (2 * (p & 3));
            if ((p & 3) == 0) { packed[p >> 2] = (uint8_t)acc; acc = 0; }
        } else ops[-nops] = (uint8_t)mv;
        nops++;
    };
    if (a.enc == 2) {
        // codes of psa_step16's TM 2 / TM 3: bit 2 (wf) / bit 3 (we) are read at the cell a walk ENTERS by a left / up move
        int entered = 1;                                  // how the walk came into the current cell: 1 free, 0 by a left move, 2 by an up move
        while (i >= 0 && j >= 0) {
            const uint8_t byte = __ldcs(d + (long long)i * stride + (j >> 1));
            const int c = (j & 1) ? (byte >> 4) : (byte & 15);
            const int mv = (entered == 0 && (c & 4)) ? 0 : ((entered == 2 && (c & 8)) ? 2 : psa_back_move(c, 1));
            put(mv);
            if (mv == 1) { i--; j--; }
            else if (mv == 0) j--;
            else i--;
            entered = mv;
        }
    } else
    while (i >= 0 && j >= 0) {
        const uint8_t byte = __ldcs(d + (long long)i * stride + (j >> 1));
        const int c = (j & 1) ? (byte >> 4) : (byte & 15);
        const int mv = forced >= 0 ? forced : psa_back_move(c, a.enc);
        put(mv);
        if (mv == 1) { i--; j--; forced = -1; }
        else if (mv == 0) { forced = (c & 4) ? 0 : -1; j--; }
        else { forced = (c & 8) ? 2 : -1; i--; }
    }
    while (j >= 0) { put(0); j--; }
    while (i >= 0) { put(2); i--; }
    if (PACK) {
        const int p = n + m - 1 - nops;                   // next (unwritten) position: flush a partly filled byte
        if ((p & 3) != 3) packed[p >> 2] = (uint8_t)acc;
    }
    a.alen[k] = nops;
}

}  // namespace tsta
