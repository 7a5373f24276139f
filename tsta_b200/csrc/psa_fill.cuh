// psa_fill.cuh -- the pairwise score-matrix fill for ONE long pair on sm_100a.
//
// Replaces the tile wavefront of the reference (psa/psa.c:620-639 scheduling
// block_alignment()/row() psa/psa.c:234-433 onto a pthread pool).
//
// Decomposition
//   * the n columns (s0, the longer read) are cut into BANDS of 32*C columns; one warp
//     owns one band for the whole sweep over the m rows (s1);
//   * inside a warp, lane k owns C consecutive columns (H(i-1,.) and Eg(i,.) live in
//     registers) and is SKEWED one row behind lane k-1: at step t lane k computes row
//     t-k.  The right edge (H, F) of a lane's strip travels to lane k+1 with
//     __shfl_up_sync -- this is the anti-diagonal wavefront at register granularity;
//   * band b hands its right boundary column to band b+1 through a global (L2-resident)
//     array in chunks of 32 rows, published with a release store on a per-band progress
//     counter and consumed with an acquire load.  Bands take their index from an atomic
//     ticket, so a waiting band only ever waits for a band that started earlier --
//     forward progress does not depend on the grid fitting on the chip.  CUDA streams +
//     these counters replace threadPoolAdd() and the busy-wait barrier of psa.c:638.
//
// Arithmetic: absolute int32 Gotoh recurrences (exactly what the reference's int8
// difference encoding computes inside its parameter domain):
//      Eg(i,j) = max(Eg(i-1,j)+E, H(i-1,j)+O+E)     Fg(i,j) = max(Fg(i,j-1)+E, H(i,j-1)+O+E)
//      H(i,j)  = max(H(i-1,j-1)+sub, Eg, Fg)        maxsorce = max over cells of H
// with DPX instructions (__viaddmax_s32 -> VIADDMNMX, __vimax3_s32 -> VIMNMX3) and the
// substitution score fetched with one PRMT from a per-row 8-entry int8 table.
//
// Traceback variant (TRACE): every score is carried DOUBLED and the spare low bit of the
// E/F gap states holds "this gap state was opened, not extended" (the reference's
// fback/eback == +-2, psa/psa.c:301-327).  With a = gap+E and t = H+O+E the next state is
// max(2a, 2t+1) and the reference's traceback rule "keep extending" (psa/psa.c:450-459)
// becomes the single comparison 2a+open >= 2t+1.  Each cell emits a 4-bit code
//      bits 0-1 back (1 diag, 0 left/F, 2 up/E; diag > F > E on ties, psa.c:294-298)
//      bit 2    cF: a left move from this cell forces the next cell to move left
//      bit 3    cE: an up move from this cell forces the next cell to move up
// staged through shared memory and written to HBM row-major with 128-bit stores.
#pragma once
#include "common.cuh"

namespace tsta {

constexpr int kPadCode = 7;      // PRMT mode: column code that never matches a row
constexpr int kChunkRows = 32;   // rows per inter-band hand-off chunk
constexpr int kRing = 64;        // shared-memory ring depth in rows (>= 32 skew + 32 chunk)

struct PsaFillArgs {
    const uint8_t* qcode;   // [n_pad] column (s0) codes, pad columns = kPadCode / 255
    const uint2* rowtab;    // [m] PRMT mode: 8 x int8 substitution scores of row i (scaled)
    const uint8_t* rcode;   // [m] generic mode: row (s1) codes
    int n, m, n_pad, nbands;
    int M, X, E, O;         // unscaled
    int2* bnd;              // [nbands][bnd_stride] right boundary (H|cF, F) of every band
    long long bnd_stride;
    int* progress;          // [nbands] rows published by each band
    unsigned* ticket;       // band index dispenser
    uint8_t* dirs;          // TRACE: [m][dirs_stride] packed 4-bit codes
    long long dirs_stride;  // bytes per row (= n_pad / 2)
    int* out_max;           // atomicMax target (scaled units)
    int* lastrow;           // [n_pad] H(m-1, j) (scaled units)
};

template <int C, bool TRACE, bool PRMT>
__global__ void __launch_bounds__(32) psa_fill_kernel(const PsaFillArgs a)
{
    constexpr int S = TRACE ? 2 : 1;                    // score scale
    constexpr int kDirBytesPerLane = C / 2;             // 4 bits per cell
    constexpr int kDirRowBytes = 32 * kDirBytesPerLane; // bytes of one band row
    static_assert(C == 4 || C == 8, "lane strip width");

    __shared__ int2 s_in[kChunkRows];
    __shared__ int2 s_out[kRing];
    __shared__ __align__(16) uint8_t s_dir[TRACE ? kRing * kDirRowBytes : 16];

    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);

    const int m = a.m;
    const int Es = S * a.E;
    const int OEs = S * (a.O + a.E) + (TRACE ? 1 : 0);  // TRACE: 2(O+E)+1, the "opened" form
    const int Ms = S * a.M, Xs = S * a.X;
    const int j0 = (band * 32 + lane) * C;

    // ---- per-lane column state -----------------------------------------------------------
    int Hup[C];          // H(i-1, j0+c)
    int Eg[C];           // Eg(i, j0+c)   (TRACE: doubled, bit0 = opened)
    uint32_t sel[C];     // PRMT selector (PRMT) or column code (generic)
    int cEf[C];          // TRACE: cE(i, j0+c) as 0/8
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int j = j0 + c;
        const uint32_t code = a.qcode[j];
        sel[c] = PRMT ? (code * 0x1111u | 0x8880u) : code;
        const int hb = a.O + (j + 1) * a.E;             // H(-1, j)        psa.c:226
        Hup[c] = S * hb;
        Eg[c] = S * (hb + a.O + a.E) + (TRACE ? 1 : 0); // Eg(0, j), opened psa.c:222-224
        cEf[c] = 0;
    }
    int diag = S * (j0 == 0 ? 0 : a.O + j0 * a.E);      // H(i-1, j0-1)
    int ms = INT_MIN;
    int Hout = 0, Fout = 0;                              // right edge of the previous step

    const int2* bnd_in = a.bnd + (long long)(band - 1) * a.bnd_stride;
    int2* bnd_out = a.bnd + (long long)band * a.bnd_stride;
    const int nchunks = (m + 31 + kChunkRows - 1) / kChunkRows;

    // software-pipelined per-row substitution table: every lane starts at row 0
    uint2 Tnext = make_uint2(0, 0);
    uint32_t rnext = 0;
    if (m > 0) {
        if (PRMT) Tnext = __ldg(&a.rowtab[0]);
        else rnext = __ldg(&a.rcode[0]);
    }

    for (int ch = 0; ch < nchunks; ch++) {
        const int t0 = ch * kChunkRows;
        if (band > 0 && t0 < m) {
            const int need = min(m, t0 + kChunkRows);
            if (lane == 0) {
                while (ld_acquire_gpu(a.progress + band - 1) < need) __nanosleep(64);
            }
            __syncwarp();
            const int row = t0 + lane;
            if (row < m) s_in[lane] = __ldcg(bnd_in + row);
            __syncwarp();
        }

#pragma unroll 1
        for (int s = 0; s < kChunkRows; s++) {
            const int i = t0 + s - lane;
            int Hin = __shfl_up_sync(kFullMask, Hout, 1);
            int Fin = __shfl_up_sync(kFullMask, Fout, 1);
            if (lane == 0) {
                if (band > 0) {
                    const int2 v = s_in[s];
                    Hin = v.x;
                    Fin = v.y;
                } else {
                    const int hb = a.O + (i + 1) * a.E;                // H(i, -1)     psa.c:228-231
                    Hin = S * hb;
                    Fin = S * (hb + a.O + a.E) + (TRACE ? 1 : 0);      // Fg(i, 0), opened
                }
            }
            if (i >= 0 && i < m) {
                const uint2 T = Tnext;
                const uint32_t rc = rnext;
                if (i + 1 < m) {
                    if (PRMT) Tnext = __ldg(&a.rowtab[i + 1]);
                    else rnext = __ldg(&a.rcode[i + 1]);
                }
                int cFv = 0;
                if (TRACE) {
                    cFv = (Hin & 1) << 2;
                    Hin &= ~1;
                }
                int D = diag;
                diag = Hin;
                int F = Fin;
                uint32_t word = 0;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    const int sub = PRMT ? prmt(T.x, T.y, sel[c]) : (sel[c] == rc ? Ms : Xs);
                    const int hs = D + sub;
                    D = Hup[c];
                    if (!TRACE) {
                        const int H = __vimax3_s32(hs, Eg[c], F);
                        const int t = H + OEs;
                        F = __viaddmax_s32(F, Es, t);
                        Eg[c] = __viaddmax_s32(Eg[c], Es, t);
                        Hup[c] = H;
                    } else {
                        const int Ec = Eg[c] & ~1;
                        const int Fc = F & ~1;
                        const int H = __vimax3_s32(hs, Ec, Fc);
                        int code = (H == Fc) ? 0 : 2;
                        code = (H == hs) ? 1 : code;
                        code |= cFv | cEf[c];
                        word |= (uint32_t)code << (4 * c);
                        const int Tp = H + OEs;                        // 2(H+O+E)+1
                        cFv = (F + Es >= Tp) ? 4 : 0;                  // cF of column j+1
                        cEf[c] = (Eg[c] + Es >= Tp) ? 8 : 0;           // cE of row i+1
                        F = __viaddmax_s32(Fc, Es, Tp);
                        Eg[c] = __viaddmax_s32(Ec, Es, Tp);
                        Hup[c] = H;
                    }
                }
                // running maximum; pad columns (last band) are included on purpose: with
                // X<=0, E<=0, O<=0 (enforced by check_params) a pad cell never exceeds the
                // real cells it derives from, and the reference includes them too (psa.c:333)
#pragma unroll
                for (int c = 0; c < C; c += 2) ms = __vimax3_s32(ms, Hup[c], Hup[c + 1]);
                Hout = Hup[C - 1] | (TRACE ? (cFv >> 2) : 0);
                Fout = F;
                if (lane == 31) s_out[i & (kRing - 1)] = make_int2(Hout, Fout);
                if (TRACE) {
                    uint8_t* dst = s_dir + (i & (kRing - 1)) * kDirRowBytes + lane * kDirBytesPerLane;
                    if (C == 8) *reinterpret_cast<uint32_t*>(dst) = word;
                    else *reinterpret_cast<uint16_t*>(dst) = (uint16_t)word;
                }
            }
        }

        // ---- publish the rows every lane has finished: [lo, hi) ---------------------------
        __syncwarp();
        const int hi = min(m, t0 + 1);
        const int lo = ch == 0 ? 0 : min(m, t0 - kChunkRows + 1);
        if (hi > lo) {
            if (band + 1 < a.nbands) {
                const int row = lo + lane;
                if (row < hi) bnd_out[row] = s_out[row & (kRing - 1)];
            }
            if (TRACE) {
                // 16-byte units: C per row; lane handles units lane, lane+32, ...
                constexpr int kUnitsPerRow = kDirRowBytes / 16;
                const int units = (hi - lo) * kUnitsPerRow;
                for (int u = lane; u < units; u += 32) {
                    const int row = lo + u / kUnitsPerRow;
                    const int part = u % kUnitsPerRow;
                    const uint4 v = *reinterpret_cast<const uint4*>(
                        s_dir + (row & (kRing - 1)) * kDirRowBytes + part * 16);
                    uint4* dst = reinterpret_cast<uint4*>(
                        a.dirs + (long long)row * a.dirs_stride + (long long)band * kDirRowBytes + part * 16);
                    __stcs(dst, v);
                }
            }
            __syncwarp();
            if (lane == 0 && band + 1 < a.nbands) {
                __threadfence();
                st_release_gpu(a.progress + band, hi);
            }
        }
    }

    // after the sweep Hup[] holds the last DP row
#pragma unroll
    for (int c = 0; c < C; c++) a.lastrow[j0 + c] = Hup[c];

    ms = warp_max(ms);
    if (lane == 0 && ms != INT_MIN) atomicMax(a.out_max, ms);
}

}  // namespace tsta
