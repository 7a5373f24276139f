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
constexpr int kTraceThreads = 256;

struct PsaTraceArgs {
    const uint8_t* dirs;
    long long dirs_stride;   // bytes per row
    int n, m;
    uint8_t* ops;            // [n + m] moves in walk order
    int* out;                // out[0] = number of moves written, out[1] = final i, out[2] = final j
    int enc;                 // back code: 0 = (1 diagonal, 0 left, 2 up), 1 = (3 diagonal, 2 left, 1 up)
};

// `back` bits of a packed code -> move (1 diagonal, 0 left, 2 up) for either encoding
__device__ __forceinline__ int psa_back_move(int c, int enc) { return enc ? ((0x48 >> (2 * (c & 3))) & 3) : (c & 3); }

static __global__ void __launch_bounds__(kTraceThreads) psa_trace_kernel(const PsaTraceArgs a)
{
    __shared__ __align__(16) uint8_t tile[kTraceRows * kTraceBytes];
    __shared__ int s_i, s_j, s_forced, s_nops;

    if (threadIdx.x == 0) {
        s_i = a.m - 1;
        s_j = a.n - 1;
        s_forced = -1;
        s_nops = 0;
    }
    __syncthreads();

    while (true) {
        const int i0 = s_i, j0 = s_j;
        if (i0 < 0 || j0 < 0) break;
        // window: rows (i0 - kTraceRows, i0], bytes [wb0, wb0 + kTraceBytes), 16-byte aligned,
        // placed so that the current column is at its right end
        const int jb = j0 >> 1;
        int wb0 = ((jb + 16) & ~15) - kTraceBytes;
        if (wb0 < 0) wb0 = 0;
        constexpr int kUnitsPerRow = kTraceBytes / 16;
        for (int u = threadIdx.x; u < kTraceRows * kUnitsPerRow; u += kTraceThreads) {
            const int r = u / kUnitsPerRow, q = u % kUnitsPerRow;
            const int row = i0 - r;
            const long long off = (long long)wb0 + q * 16;
            if (row >= 0 && off < a.dirs_stride) {
                const uint4 v = __ldcs(reinterpret_cast<const uint4*>(a.dirs + (long long)row * a.dirs_stride + off));
                *reinterpret_cast<uint4*>(tile + r * kTraceBytes + q * 16) = v;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int i = i0, j = j0, forced = s_forced, nops = s_nops;
            const int ilim = i0 - kTraceRows;  // exclusive
            const int jlim = wb0 * 2;          // inclusive lowest column in the window
            while (i > ilim && i >= 0 && j >= jlim) {
                const uint8_t byte = tile[(i0 - i) * kTraceBytes + ((j >> 1) - wb0)];
                const int c = (j & 1) ? (byte >> 4) : (byte & 15);
                const int d = forced >= 0 ? forced : psa_back_move(c, a.enc);
                a.ops[nops++] = (uint8_t)d;
                if (d == 1) { i--; j--; forced = -1; }
                else if (d == 0) { forced = (c & 4) ? 0 : -1; j--; }
                else { forced = (c & 8) ? 2 : -1; i--; }
            }
            s_i = i; s_j = j; s_forced = forced; s_nops = nops;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.out[0] = s_nops;
        a.out[1] = s_i;
        a.out[2] = s_j;
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
