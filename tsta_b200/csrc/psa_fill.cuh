// psa_fill.cuh -- the pairwise score-matrix fill on sm_100a (one long pair, or a batch).
//
// Replaces the tile wavefront of the reference (psa/psa.c:620-639 scheduling
// block_alignment()/row() psa/psa.c:234-433 onto a pthread pool).
//
// Decomposition
//   * the n columns (s0, the longer read) are cut into BANDS of 32*C columns; one warp
//     sweeps one band over all m rows (s1);
//   * inside a warp, lane k owns C consecutive columns (H(i-1,.) and Eg(i,.) live in
//     registers) and is SKEWED one row behind lane k-1: at step t lane k computes row
//     t-k.  The right edge (H, F) of a lane's strip travels to lane k+1 with
//     __shfl_up_sync -- the anti-diagonal wavefront at register granularity;
//   * PIPE mode (one long pair): every band is its own warp.  Band b hands its right
//     boundary column to band b+1 through a global (L2-resident) array, one 64-bit element
//     per row that carries its own validity tag (data + epoch in a single store), so the
//     hand-off needs no fence and no separate flag; the consumer prefetches a chunk ahead.
//     Bands take their index from an atomic ticket, so a waiting band only ever waits for a
//     band that started earlier -- forward progress does not depend on the grid fitting on
//     the chip.  CUDA streams + these tags replace threadPoolAdd() and the busy-wait
//     barrier of psa.c:638;
//   * sequential mode (batches of independent pairs): one warp walks the bands of its
//     pair one after the other, the boundary column parked in a per-warp scratch column.
//
// Arithmetic: absolute int32 Gotoh recurrences (exactly what the reference's int8
// difference encoding computes inside its parameter domain):
//      Eg(i,j) = max(Eg(i-1,j)+E, H(i-1,j)+O+E)     Fg(i,j) = max(Fg(i,j-1)+E, H(i,j-1)+O+E)
//      H(i,j)  = max(H(i-1,j-1)+sub, Eg, Fg)        maxsorce = max over cells of H
// with DPX instructions (__viaddmax_s32 -> VIADDMNMX, __vimax3_s32 -> VIMNMX3).  Because
// O <= 0, H+O+E inside the F recurrence can be replaced by max(h,Eg)+O+E, which does not
// depend on F: the serial chain along a row is ONE VIADDMNMX per cell.  The substitution
// score comes from one PRMT on a per-row 8-entry int8 table (alphabets of <= 7 symbols) or
// a compare/select (any alphabet).
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
constexpr int kMaxChunkRows = 32; // rows per staging chunk: template parameter CH (8, 16 or 32)
constexpr int kDirRing = 64;     // direction ring depth in rows (>= 32 skew + 32 chunk)
constexpr int kRowRing = 128;    // row-table ring depth (staged one chunk ahead)

// Scoring + alphabet, shared by every kernel of the pairwise path.
struct PsaScoring {
    int M, X, E, O;        // unscaled
    uint8_t lut[256];      // raw byte -> code 0..6 (PRMT mode only)
};

// Per-warp shared-memory workspace.
template <int C, bool TRACE>
struct __align__(16) PsaWarpSmem {
    uint8_t dir[TRACE ? kDirRing * (32 * C / 2) : 16];  // ring of packed direction rows
    uint2 rowring[kRowRing];  // per-row substitution table (PRMT) / raw row byte in .x (generic)
    int2 in[kMaxChunkRows];   // left boundary (H|cF, F) of the current chunk
};

// Boundary column element, 8 bytes, written and read with ONE 64-bit access so that data and
// validity tag can never be observed torn (the NCCL "LL" idea): no fences, no separate flag.
//   low  word: H (TRACE: doubled, bit0 = cF of the next column)
//   high word: bits 0-15 = F - H (always within [2(O+E), 0]), bits 16-31 = epoch tag (!= 0)
__device__ __forceinline__ unsigned long long bnd_pack(int Hout, int Fout, bool trace, unsigned epoch_hi)
{
    const int Hc = trace ? (Hout & ~1) : Hout;
    const unsigned hi = ((unsigned)(Fout - Hc) & 0xffffu) | epoch_hi;   // epoch_hi = epoch << 16
    return ((unsigned long long)hi << 32) | (unsigned)Hout;
}
__device__ __forceinline__ int2 bnd_unpack(unsigned long long v, bool trace)
{
    const int H = (int)(unsigned)v;
    const int d = (int)(short)(unsigned short)(v >> 32);
    return make_int2(H, (trace ? (H & ~1) : H) + d);
}
#ifndef TSTA_BND_LD_SCOPE
#define TSTA_BND_LD_SCOPE "sys"     // the producer may be a peer GPU (column slices)
#endif
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed." TSTA_BND_LD_SCOPE ".global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Shared-memory accesses of the step loop go through explicit 32-bit shared addresses: on
// sm_100 the compiler otherwise re-derives the CTA's shared window (S2UR SR_CgaCtaId) inside
// the loop, which lands on the step's critical path.
__device__ __forceinline__ uint2 lds_u2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y));
}
// Predicated (branch-free) tagged store of one boundary element.
__device__ __forceinline__ void st_relaxed_u64_if(unsigned long long* p, unsigned long long v, bool doit)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.relaxed.gpu.global.b64 [%0], %1;\n\t}"
                 ::"l"(p), "l"(v), "r"((unsigned)doit) : "memory");
}

// Everything one band sweep needs.
struct PsaBand {
    const uint8_t* s0;      // raw columns of this pair
    const uint8_t* s1;      // raw rows of this pair
    int n, m;
    int band, nbands;
    const unsigned long long* bnd_in;   // [m] left boundary column (band > 0)
    unsigned long long* bnd_out;        // [m] right boundary column (written when band+1 < nbands)
    unsigned epoch;         // PIPE: tag that marks boundary elements of THIS run (1..65535)
    uint8_t* dirs;          // TRACE: [m][dirs_stride] packed codes of this pair
    long long dirs_stride;
    int* lastrow;           // H(m-1, j) goes to lastrow[j - lastrow_col0] (scaled units), or NULL
    int lastrow_col0;
    // column slice of a pair that is spread over several GPUs (packed score-only kernels):
    // s0 / n / band indices are local to the slice, col_off is the slice's first global column
    int col_off = 0;
    bool has_left = false;  // band 0 gets its left boundary from bnd_in (the neighbour GPU's stores)
    bool has_right = false; // the last band exports its right boundary to bnd_out (peer memory)
    int* abort_flag = nullptr;          // see PsaPairArgs
    unsigned long long timeout_ns = 0;
    unsigned long long* timeline = nullptr;
    // checkpointed traceback (psa_fill16.cuh): chunk range of this sweep, the checkpoints the score-only
    // pass leaves behind, the checkpoint a retrace sweep resumes from
    int ch_begin = 0, ch_end = INT_MAX;
    uint32_t* ckpt_out = nullptr;        // this band's checkpoints [k][words][32 lanes] (k = step / 128, k >= 1)
    const uint32_t* ckpt_in = nullptr;   // the checkpoint of chunk ch_begin, in the score-only pass's form
    bool bnd_unscaled = false;           // traced sweep over boundary columns a score-only pass wrote
    long long dir_step0 = 0;             // TM 2: `dirs` starts at this step
};

// Fills the 256-entry byte -> substitution-table map (PRMT mode).
__device__ __forceinline__ void psa_build_tab(uint2* tab, const PsaScoring& sc, int S, int tid, int nthreads)
{
    const int Ms = S * sc.M, Xs = S * sc.X;
    for (int b = tid; b < 256; b += nthreads) {
        const uint32_t c = sc.lut[b];
        uint32_t w[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int s = ((uint32_t)(h * 4 + k) == c) ? Ms : Xs;
                v |= (uint32_t)(s & 0xff) << (8 * k);
            }
            w[h] = v;
        }
        tab[b] = make_uint2(w[0], w[1]);
    }
}

// Per-lane register state of a band sweep.
template <int C>
struct PsaLane {
    int Hup[C];          // H(i-1, j0+c)
    int Eg[C];           // Eg(i, j0+c)   (TRACE: doubled, bit0 = opened)
    uint32_t sel[C];     // PRMT selector (PRMT) or raw column byte / 256 for pad (generic)
    int cEf[C];          // TRACE: cE(i, j0+c) as 0/8
    int diag;            // H(i-1, j0-1)
    int Hout, Fout;      // right edge produced by the previous step
    int ms;              // running maximum
    uint2 Tnext;         // substitution table of the next row (software prefetch)
};

// One wavefront step of one lane: row i of the lane's C columns.
//   CHECK = false : steady state, every lane has a valid row (no predication at all)
//   CHECK = true  : ramp-up / ramp-down chunks, lanes outside [0, m) keep their state
// Schedule: the two shuffles and the shared-memory loads are issued first; everything that
// does not depend on the incoming F (substitution, max(h, Eg), +O+E) fills their latency;
// the serial part is one VIADDMNMX per cell (two ops in TRACE mode).
template <int C, bool TRACE, bool PRMT, bool CHECK>
__device__ __forceinline__ void psa_step(PsaLane<C>& L, const int i, const int s, const int m, const int lane,
                                         const int Es, const int OEs, const int Kf, const int Ms, const int Xs,
                                         const uint32_t in_addr, const uint32_t ring_addr, const uint32_t dir_addr,
                                         unsigned long long* bnd_out, const bool store_right, const unsigned epoch_hi)
{
    constexpr int kDirBytesPerLane = C / 2;
    constexpr int kDirRowBytes = 32 * kDirBytesPerLane;
    constexpr int kWords = (C + 7) / 8;

    const int Hs = __shfl_up_sync(kFullMask, L.Hout, 1);
    const int Fs = __shfl_up_sync(kFullMask, L.Fout, 1);
    const uint2 v0 = lds_u2(in_addr + s * 8);                  // broadcast: lane 0's boundary
    const uint2 T = L.Tnext;
    L.Tnext = lds_u2(ring_addr + (((i + 1) & (kRowRing - 1)) << 3));
    const bool active = !CHECK || ((unsigned)i < (unsigned)m);

    if (active) {
        // ---- F-independent pre-work (uses only last step's state) ---------------------------
        int hE[C], g[C], hs[C], Ec[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int sub = PRMT ? prmt(T.x, T.y, L.sel[c]) : (L.sel[c] == T.x ? Ms : Xs);
            const int D = c == 0 ? L.diag : L.Hup[c - 1];
            if (!TRACE) {
                hE[c] = __viaddmax_s32(D, sub, L.Eg[c]);       // max(h, Eg)
            } else {
                hs[c] = D + sub;
                Ec[c] = L.Eg[c] & ~1;
                hE[c] = max(hs[c], Ec[c]);
            }
            g[c] = hE[c] + OEs;                                // max(h,Eg)+O+E  (TRACE: doubled, +1)
        }
        // ---- incoming edge ---------------------------------------------------------------
        int Hin = lane == 0 ? (int)v0.x : Hs;
        int F = lane == 0 ? (int)v0.y : Fs;
        int cFv = 0;
        if (TRACE) {
            cFv = (Hin & 1) << 2;
            Hin &= ~1;
        }
        L.diag = Hin;
        // ---- serial chain + per-cell epilogue ----------------------------------------------
        uint32_t word[kWords];
#pragma unroll
        for (int k = 0; k < kWords; k++) word[k] = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (!TRACE) {
                const int Fc = F;
                F = __viaddmax_s32(Fc, Es, g[c]);              // Fg(i, j+1): the only serial op
                const int H = max(hE[c], Fc);
                L.Eg[c] = __viaddmax_s32(L.Eg[c], Es, H + OEs); // the add issues on the FMA pipe (IMAD)
                L.Hup[c] = H;
            } else {
                const int Fin = F;
                const int Fc = Fin & ~1;
                F = __viaddmax_s32(Fc, Kf, g[c]);              // max(2a, 2t+1) with t folded (Kf = max(2E, 2(O+E)+1))
                const int H = max(hE[c], Fc);
                int code = (H == Fc) ? 0 : 2;
                code = (H == hs[c]) ? 1 : code;
                code |= cFv | L.cEf[c];
                word[c / 8] |= (uint32_t)code << (4 * (c & 7));
                const int Tp = H + OEs;                        // 2(H+O+E)+1
                cFv = (Fin + Es >= Tp) ? 4 : 0;                // cF of column j+1
                L.cEf[c] = (L.Eg[c] + Es >= Tp) ? 8 : 0;       // cE of row i+1
                L.Eg[c] = __viaddmax_s32(Ec[c], Es, Tp);
                L.Hup[c] = H;
            }
        }
        // running maximum; pad columns (last band) are included on purpose: with X<=0, E<=0,
        // O<=0 (enforced by check_params) a pad cell never exceeds the real cells it derives
        // from, and the reference includes them too (psa.c:333)
#pragma unroll
        for (int c = 0; c < C; c += 2) L.ms = __vimax3_s32(L.ms, L.Hup[c], L.Hup[c + 1]);
        L.Hout = L.Hup[C - 1] | (TRACE ? (cFv >> 2) : 0);
        L.Fout = F;
        st_relaxed_u64_if(bnd_out + i, bnd_pack(L.Hout, L.Fout, TRACE, epoch_hi), store_right);
        if (TRACE) {
            const uint32_t dst = dir_addr + (i & (kDirRing - 1)) * kDirRowBytes + lane * kDirBytesPerLane;
            if (C == 16) sts_u2(dst, word[0], word[kWords - 1]);
            else if (C == 8) sts_u32(dst, word[0]);
            else sts_u16(dst, word[0]);
        }
    }
}

// One band of one pair.  Returns the lane-local running maximum (scaled units).
template <int C, bool TRACE, bool PRMT, bool PIPE, int CH>
__device__ __forceinline__ int psa_sweep_band(const PsaBand& b, const PsaScoring& sc, const uint2* tab,
                                              PsaWarpSmem<C, TRACE>& sm, const int lane)
{
    constexpr int S = TRACE ? 2 : 1;                    // score scale
    constexpr int kDirRowBytes = 32 * C / 2;            // bytes of one band row of codes
    static_assert(C == 4 || C == 6 || C == 8 || C == 16, "lane strip width");
    static_assert(!TRACE || C != 6, "traced strips must pack to whole 16-bit units");
    static_assert(CH == 8 || CH == 16 || CH == 32, "chunk rows");
    constexpr int kChunkRows = CH;

    const int m = b.m;
    const int Es = S * sc.E;
    const int OEs = S * (sc.O + sc.E) + (TRACE ? 1 : 0);  // TRACE: 2(O+E)+1, the "opened" form
    const int Kf = max(Es, OEs);
    const int Ms = S * sc.M, Xs = S * sc.X;
    const int j0 = (b.band * 32 + lane) * C;

    PsaLane<C> L;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int j = j0 + c;
        if (PRMT) {
            const uint32_t code = j < b.n ? sc.lut[b.s0[j]] : (uint32_t)kPadCode;
            L.sel[c] = code * 0x1111u | 0x8880u;
        } else {
            L.sel[c] = j < b.n ? (uint32_t)b.s0[j] : 256u;
        }
        const int hb = sc.O + (j + 1) * sc.E;               // H(-1, j)        psa.c:226
        L.Hup[c] = S * hb;
        L.Eg[c] = S * (hb + sc.O + sc.E) + (TRACE ? 1 : 0); // Eg(0, j), opened psa.c:222-224
        L.cEf[c] = 0;
    }
    L.diag = S * (j0 == 0 ? 0 : sc.O + j0 * sc.E);          // H(-1, j0-1)
    L.ms = INT_MIN;
    L.Hout = 0;
    L.Fout = 0;

    const int nchunks = (m + 31 + kChunkRows - 1) / kChunkRows;   // lane 31 finishes row m-1 at step m+30
    const bool export_right = (b.band + 1 < b.nbands) && b.bnd_out != nullptr;
    const unsigned epoch = PIPE ? b.epoch : 1u;
    const unsigned epoch_hi = epoch << 16;
    const bool store_right = export_right && lane == 31;
    const uint32_t in_addr = (uint32_t)__cvta_generic_to_shared(sm.in);
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(sm.rowring);
    const uint32_t dir_addr = (uint32_t)__cvta_generic_to_shared(sm.dir);

    // ---- prologue: row tables of chunk 0, prefetch registers for chunk 1 / boundary chunk 0 ----
    __syncwarp();
    const bool stager = lane < kChunkRows;                 // lanes that stage one row per chunk
    if (stager && lane < m) {
        const uint32_t raw = b.s1[lane];
        sm.rowring[lane] = PRMT ? tab[raw] : make_uint2(raw, 0u);
    }
    uint32_t raw_next = (stager && kChunkRows + lane < m) ? b.s1[kChunkRows + lane] : 0u;  // row of the NEXT chunk
    unsigned long long bnd_next = 0;                                               // boundary of THIS chunk (prefetched)
    if (b.band > 0 && stager && lane < m) bnd_next = ld_relaxed_u64(b.bnd_in + lane);
    __syncwarp();
    L.Tnext = sm.rowring[0];                               // lane 0 starts at row 0; others re-prefetch before row 0

    for (int ch = 0; ch < nchunks; ch++) {
        const int t0 = ch * kChunkRows;
        // ---- stage: next chunk's row tables, this chunk's left boundary ----------------------
        {
            const int mm = stager ? m : 0;                  // non-staging lanes see an empty range
            const int rown = t0 + kChunkRows + lane;       // row of the next chunk
            if (rown < mm) sm.rowring[rown & (kRowRing - 1)] = PRMT ? tab[raw_next] : make_uint2(raw_next, 0u);
            const int rownn = rown + kChunkRows;
            raw_next = rownn < mm ? b.s1[rownn] : 0u;
            const int row = stager ? t0 + lane : m;
            if (b.band > 0) {
                if (PIPE) {
                    // warp-uniform spin (all lanes leave together, so the warp stays converged
                    // for the shuffles of the step loop)
                    while (!__all_sync(kFullMask, row >= m || (unsigned)(bnd_next >> 48) == epoch)) {
                        __nanosleep(20);
                        if (row < m && (unsigned)(bnd_next >> 48) != epoch) bnd_next = ld_relaxed_u64(b.bnd_in + row);
                    }
                }
                if (row < m) sm.in[lane] = bnd_unpack(bnd_next, TRACE);
            } else if (row < m) {
                const int hb = sc.O + (row + 1) * sc.E;     // H(i, -1)   psa.c:228-231; Fg(i,0) opened
                sm.in[lane] = make_int2(S * hb, S * (hb + sc.O + sc.E) + (TRACE ? 1 : 0));
            }
            const int rowp = row + kChunkRows;             // prefetch the next chunk's boundary
            if (b.band > 0 && rowp < mm) bnd_next = ld_relaxed_u64(b.bnd_in + rowp);
        }
        __syncwarp();

        if (t0 >= 31 && t0 + kChunkRows - 1 < m) {
#pragma unroll 2
            for (int s = 0; s < kChunkRows; s++)
                psa_step<C, TRACE, PRMT, false>(L, t0 + s - lane, s, m, lane, Es, OEs, Kf, Ms, Xs, in_addr, ring_addr,
                                                dir_addr, b.bnd_out, store_right, epoch_hi);
        } else {
#pragma unroll 1
            for (int s = 0; s < kChunkRows; s++)
                psa_step<C, TRACE, PRMT, true>(L, t0 + s - lane, s, m, lane, Es, OEs, Kf, Ms, Xs, in_addr, ring_addr,
                                               dir_addr, b.bnd_out, store_right, epoch_hi);
        }

        // ---- flush the direction rows every lane has finished: [lo, hi) ----------------------
        __syncwarp();
        if (TRACE) {
            const int hi = min(m, max(0, t0 + kChunkRows - 31));
            const int lo = min(m, max(0, t0 - 31));
            // 16-byte units: kDirRowBytes/16 per row; lane handles units lane, lane+32, ...
            constexpr int kUnitsPerRow = kDirRowBytes / 16;
            const int units = (hi - lo) * kUnitsPerRow;
            for (int u = lane; u < units; u += 32) {
                const int row = lo + u / kUnitsPerRow;
                const int part = u % kUnitsPerRow;
                const uint4 v = *reinterpret_cast<const uint4*>(
                    sm.dir + (row & (kDirRing - 1)) * kDirRowBytes + part * 16);
                uint4* dst = reinterpret_cast<uint4*>(
                    b.dirs + (long long)row * b.dirs_stride + (long long)b.band * kDirRowBytes + part * 16);
                __stcs(dst, v);
            }
        }
    }

    // after the sweep Hup[] holds the last DP row
    if (b.lastrow) {
#pragma unroll
        for (int c = 0; c < C; c++) b.lastrow[j0 - b.lastrow_col0 + c] = L.Hup[c];
    }
    return L.ms;
}

// ---------------------------------------------------------------------------------------
// one long pair: one warp (= one CTA) per band, pipelined through global flags
// ---------------------------------------------------------------------------------------
struct PsaPairArgs {
    PsaScoring sc;
    const uint8_t* s0;
    const uint8_t* s1;
    int n, m, nbands;
    unsigned long long* bnd;  // [nbands][bnd_stride] packed boundary columns (zeroed at allocation)
    long long bnd_stride;
    unsigned epoch;         // tag of this run's boundary elements (1..65535, never reused before a re-zero)
    unsigned* ticket;
    uint8_t* dirs;
    long long dirs_stride;
    int* out_max;           // atomicMax target (scaled units)
    int* lastrow;           // [n_pad]
    // multi-GPU column slices (psa_fill16.cuh): global column of the slice, the tagged column
    // the left neighbour GPU writes into this GPU's memory, the right neighbour's inbox (peer pointer)
    int col_off;
    const unsigned long long* inbox;
    unsigned long long* outbox;
    // hand-off watchdog (packed kernels): a band that polls its left neighbour longer than timeout_ns
    // raises *abort_flag and stops waiting (results are then garbage and the host reports TSTA_ECUDA)
    int* abort_flag;
    unsigned long long timeout_ns;
    unsigned long long* timeline;   // optional (TSTA_PSA_TIMELINE=1): 8 globaltimer stamps per band, see tools/timeline.py
    uint32_t* ckpt;                 // optional: lane-register checkpoints every 128 steps, [band][k][words][32] (psa_fill16.cuh)
    long long ckpt_band_words;
};

template <int C, bool TRACE, bool PRMT, int CH>
__global__ void __launch_bounds__(32) psa_pair_kernel(const __grid_constant__ PsaPairArgs a)
{
    __shared__ PsaWarpSmem<C, TRACE> sm;
    __shared__ uint2 tab[PRMT ? 256 : 1];
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);
    if (PRMT) psa_build_tab(tab, a.sc, TRACE ? 2 : 1, lane, 32);
    __syncwarp();

    PsaBand b;
    b.s0 = a.s0; b.s1 = a.s1; b.n = a.n; b.m = a.m;
    b.band = band; b.nbands = a.nbands;
    b.bnd_in = a.bnd + (long long)(band - 1) * a.bnd_stride;
    b.bnd_out = a.bnd + (long long)band * a.bnd_stride;
    b.epoch = a.epoch;
    b.dirs = a.dirs; b.dirs_stride = a.dirs_stride;
    b.lastrow = a.lastrow; b.lastrow_col0 = 0;
    int ms = psa_sweep_band<C, TRACE, PRMT, true, CH>(b, a.sc, tab, sm, lane);
    ms = warp_max(ms);
    if (lane == 0) atomicMax(a.out_max, ms);
}

// ---------------------------------------------------------------------------------------
// a batch of independent pairs: persistent warps, each takes whole pairs from a ticket and
// walks that pair's bands sequentially (no inter-warp dependency at all)
// ---------------------------------------------------------------------------------------
struct PsaBatchArgs {
    PsaScoring sc;
    const uint8_t* seqs0;
    const uint8_t* seqs1;
    const long long* off0;  // [npairs]
    const long long* off1;
    const int* len0;        // [npairs] (len0 >= len1)
    const int* len1;
    long long npairs;
    unsigned long long* scratch;  // [total warps][scratch_stride] boundary column parking
    long long scratch_stride;
    unsigned long long* ticket;
    uint8_t* dirs;          // TRACE: packed codes of all pairs
    const long long* dirs_off;  // [npairs] byte offset of pair k's matrix
    int* lastrow;           // [total warps][32*C] last-band last row, for the corner read
    int* maxsorce;          // [npairs] unscaled, seeded with -100
    int* corner;            // [npairs] unscaled
};

template <int C, bool TRACE, bool PRMT, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) psa_batch_kernel(const __grid_constant__ PsaBatchArgs a)
{
    constexpr int S = TRACE ? 2 : 1;
    __shared__ PsaWarpSmem<C, TRACE> sm[WARPS];
    __shared__ uint2 tab[PRMT ? 256 : 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (PRMT) psa_build_tab(tab, a.sc, S, threadIdx.x, 32 * WARPS);
    __syncthreads();
    const long long gw = (long long)blockIdx.x * WARPS + w;
    unsigned long long* scratch = a.scratch + gw * a.scratch_stride;
    int* lastrow = a.lastrow + gw * (32 * C);

    while (true) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(a.ticket, 1ull);
        k = __shfl_sync(kFullMask, k, 0);
        if (k >= (unsigned long long)a.npairs) break;
        const int n = a.len0[k], m = a.len1[k];
        const int nbands = (n + 32 * C - 1) / (32 * C);
        PsaBand b;
        b.s0 = a.seqs0 + a.off0[k]; b.s1 = a.seqs1 + a.off1[k];
        b.n = n; b.m = m; b.nbands = nbands;
        // In place: band b prefetches row r of the column at the start of chunk r/32 - 1 and
        // overwrites it (lane 31, step r+31) during chunk >= r/32, so every row is read before
        // it is rewritten; __syncwarp at the top of each sweep orders band b's stores before
        // band b+1's loads.
        b.bnd_in = scratch; b.bnd_out = scratch;
        b.epoch = 1u;
        b.dirs = TRACE ? a.dirs + a.dirs_off[k] : nullptr;
        b.dirs_stride = (long long)nbands * (32 * C / 2);
        b.lastrow_col0 = (nbands - 1) * 32 * C;
        int ms = INT_MIN;
        if (m > 0) {
            for (int band = 0; band < nbands; band++) {
                b.band = band;
                b.lastrow = (band == nbands - 1) ? lastrow : nullptr;
                ms = max(ms, psa_sweep_band<C, TRACE, PRMT, false, 32>(b, a.sc, tab, sm[w], lane));
            }
        }
        ms = warp_max(ms);
        __syncwarp();
        if (lane == 0) {
            const bool real = n > 0 && m > 0;
            a.maxsorce[k] = real ? max(-100, ms / S) : -100;   // seed: psa.c:110,590
            a.corner[k] = real ? __ldcg(lastrow + (n - 1) - b.lastrow_col0) / S : 0;
        }
        __syncwarp();
    }
}

}  // namespace tsta
