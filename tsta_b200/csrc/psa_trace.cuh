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

static __global__ void __launch_bounds__(128) psa_batch_trace_kernel(const PsaBatchTraceArgs a)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.npairs) return;
    const int n = a.len0[k], m = a.len1[k];
    const long long stride = (long long)((n + a.cols_per_band - 1) / a.cols_per_band) * (a.cols_per_band / 2);
    const uint8_t* d = a.dirs + a.dirs_off[k];
    uint8_t* ops = a.ops + a.ops_off[k] + (n + m - 1);   // written backwards
    int i = m - 1, j = n - 1, forced = -1, nops = 0;
    while (i >= 0 && j >= 0) {
        const uint8_t byte = __ldcs(d + (long long)i * stride + (j >> 1));
        const int c = (j & 1) ? (byte >> 4) : (byte & 15);
        const int mv = forced >= 0 ? forced : psa_back_move(c, a.enc);
        ops[-(nops++)] = (uint8_t)mv;
        if (mv == 1) { i--; j--; forced = -1; }
        else if (mv == 0) { forced = (c & 4) ? 0 : -1; j--; }
        else { forced = (c & 8) ? 2 : -1; i--; }
    }
    while (j >= 0) { ops[-(nops++)] = 0; j--; }
    while (i >= 0) { ops[-(nops++)] = 2; i--; }
    a.alen[k] = nops;
}

}  // namespace tsta
