// psa_fill16.cuh -- the pairwise fill with TWO cells per 32-bit register (packed int16 x 2).
//
// Same decomposition as psa_fill.cuh (bands of columns, lanes skewed one row apart, tagged
// boundary hand-off between bands) and the same recurrences (reference psa/psa.c:234-363), but
// every register holds two independent cells, so one DPX instruction (VIADDMNMX.S16x2,
// VIMNMX3.S16x2, VIADD.16x2; one PRMT for both substitution scores) advances two cells:
//
//   * a warp is 64 VIRTUAL lanes: physical lane k carries virtual lane k in the low halves
//     and virtual lane k+32 in the high halves.  Virtual lane v owns C columns and works on
//     row t - v at step t.  The right edge moves to the next virtual lane by ONE rotating
//     shuffle: lane k receives lane k-1 (both halves at once), lane 0 receives lane 31's low
//     half into its high half (virtual lane 31 -> 32) and the band's left boundary into its
//     low half.  A band is therefore 64*C columns wide;
//   * the reference keeps int8 DIFFERENCES against a moving int32 column (psa.c:291,329); here
//     the int16 halves are scores relative to a warp-wide int32 `base` that is re-centred at
//     every chunk of rows.  Neighbouring cells differ by at most M + |O+E| (affine-gap
//     bound), all live cells of a warp lie within 64 rows x 64*C columns, so the span is
//     bounded and the host only selects this kernel when it fits int16 with margin
//     (psa16_fits).  Outside that domain the int32 kernels of psa_fill.cuh are used;
//   * gap states are carried shifted by the open penalty (Et = Eg - (O+E), Ft = Fg - (O+E)),
//     which makes every recurrence a single max(a + const, b):
//         X  = H(i-1,j-1) + sub              VIADD.16x2
//         hE = max(Et + (O+E), X)            VIADDMNMX.S16x2    = max(Eg, h)
//         Ft'= max(Ft + E, hE)               VIADDMNMX.S16x2    the only serial op
//         H  = max(Ft + (O+E), hE)           VIADDMNMX.S16x2
//         Et'= max(Et + E, H)                VIADDMNMX.S16x2
//   * boundary elements carry the ABSOLUTE scores modulo 2^16; the consumer re-centres them on
//     its own base (exact, because the true value is within int16 of that base anyway).
//
// Traceback variant (TRACE): scores are carried times FOUR, which frees two low bits per value.
//   * the three candidates of H are tagged with their tie-break priority (psa.c:294-298: diagonal
//     beats F beats E): H' = max3(4h+3, 4Fg+2, 4Eg+1) -- ONE VIMNMX3.S16x2 -- and the winner's tag
//     is the `back` code (3 diagonal, 2 left, 1 up);
//   * gap states carry "opened, not extended" in bit 1 (the reference's fback/eback == +-2):
//     with a = gap+E and t = H+O+E the next state is max(4a, 4t+2);
//   * the reference's rule "keep extending" (trace(), psa.c:450-459; cF/cE of SURVEY.md 8c) is
//     4a + 2*opened >= 4t+2, which is bit 0 of max(state + 4E + 1, 4t+2): one more VIADDMNMX each
//     for E and F, no compare/select at all;
//   * the 4-bit codes of a lane's C columns are assembled with IMADs (FMA pipe) and leave through
//     a shared-memory ring indexed by STEP (one store per step, no address arithmetic); the flush
//     that writes whole rows to HBM with 128-bit stores does the lane -> row transposition.
//     (TM == 1: batches of pairs.)
//
// Traceback variant of ONE LONG PAIR (TM == 2, "skewed" direction layout): the direction words
// never meet shared memory.  Every step each lane stores its word straight to HBM at
// [band][step][lane] -- 128 contiguous bytes per warp and step -- so row r of virtual lane v sits
// in the word of step r + v; the traceback kernel undoes the skew while it stages the few
// windows the path actually crosses (psa_trace.cuh), which moves the lane -> row transposition
// from 1e10 cells to ~1e8.  The cell code holds what the cell itself knows:
//     bits 0-1  back tag (3 diagonal, 2 left, 1 up)
//     bit 2     wf: a walk that ENTERS this cell by a left move keeps moving left
//               (= cF of the cell to the right)
//     bit 3     we: a walk that enters this cell by an up move keeps moving up (= cE of the cell below)
// so nothing is carried from row to row or lane to lane for the codes, and F takes the textbook
// form max(Fg+E, H+O+E) from the tagged H:  per packed column
//     hs3 = D + sub            Ec1 = (Et & ~3) | 1        Fc2 = Fin | 2
//     Hp  = max3(hs3, Fc2, Ec1)                            VIMNMX3.S16x2, winner's tag = back
//     Hn  = Hp & ~3            T2 = Hn + 4(O+E)+2
//     F'  = max(Fc2 + 4E-2, T2)    En = max(Ec1 + 4E-1, T2)
//     WE  = max(Et + 4E+1, T2)     WF = max(Fin + 4E+1, T2)      bit 0 = we / wf
// = 14.5 ALU-pipe instructions against 17.5 of the TM == 1 form.
//
// Checkpointed traceback (the default for one long pair on one GPU).  A traced cell costs 14.5
// ALU-pipe instructions against 6.5 score-only, and a traceback visits ~2(n+m) of the n*m cells.  So the
// whole matrix is filled SCORE-ONLY (TM 0) while every band saves its lane registers every 128 steps
// (kCkptWords<C> words per lane: H and E state of the lane's C columns, the diagonal input, the right-edge
// hand-off pair, the warp's base) next to the boundary columns the bands exchange anyway.  Any TILE
// = (band, 128 steps) can then be re-run on its own in the traced form (TM 2): psa_retrace16_kernel
// resumes the sweep from the tile's checkpoint, takes the left boundary from the stored column, and
// writes the tile's direction words.  Only the tiles in a corridor around the path are re-run (a few
// per cent of the matrix, all independent -> the whole GPU works on them), and if the walk leaves the
// corridor the host re-centres it and resumes (tsta_psa.cu).  The score-only state converts exactly:
// scores times four; a gap state carried minus (O+E) is "opened" iff it equals the H it was derived from.
//
// Ramp-in without masks (CK == 2): the rows above row 0 that the skewed lanes would "work on"
// during the first 63 steps are made FIXED POINTS of the recurrence -- their substitution table
// holds -128 for every symbol, every lane starts with the right-edge values of row -1 in its
// hand-off registers, and then H(r, j) = H(-1, j), Eg = H(-1, j)+O+E (opened), Fg(r, j+1) = H(-1, j)+E
// reproduce themselves for every r < 0 (O <= 0, E <= 0).  Only the running maximum is masked.  The
// time a band needs until its first boundary rows are out is what every later band waits for
// (measured: the masked start-up cost 2.3 x a steady step and set 80 of the 180 steps of
// effective lag per band, profiles/r02_calibration.md).
#pragma once
#include "psa_fill.cuh"
#include <type_traits>

namespace tsta {

constexpr int kRing16 = 128;          // row-table ring (rows)
constexpr int kStreamRing = 32;       // streamed left boundary: rows of the shared-memory ring
constexpr uint32_t kMin16x2 = 0x80008000u;
constexpr int kDirSteps = 128;        // TRACE: ring of per-step direction words (steps)
constexpr int kCkptChunks = kTileSteps / 32;          // checkpoint every 4 chunks of 32 steps
template <int C> constexpr int kCkptWords = 2 * C + 4;   // per lane: Hup[C], Et[C], diag, Hout, Fout, base

__device__ __forceinline__ uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
__device__ __forceinline__ uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
__device__ __forceinline__ uint32_t dup16(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)v >> 16; }
__device__ __forceinline__ uint32_t prmtu(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// bitwise select: bits of a where m is set, bits of b elsewhere (one LOP3)
__device__ __forceinline__ uint32_t bsel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }
// predicated tagged store of one boundary element given as two 32-bit halves
__device__ __forceinline__ void st_relaxed_pair_if(unsigned long long* p, uint32_t lo, uint32_t hi, bool doit)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 v;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.b64 v, {%1, %2};\n\t"
                 "@p st.relaxed." TSTA_BND_LD_SCOPE ".global.b64 [%0], v;\n\t}"
                 ::"l"(p), "r"(lo), "r"(hi), "r"((unsigned)doit) : "memory");
}
template <class SM>
__device__ __forceinline__ void ring16_put(SM& sm, int row, uint2 v)
{
    const int k = row & (kRing16 - 1);
    sm.rowring[k] = v;
    if (k < kMaxChunkRows) sm.rowring[kRing16 + k] = v;
}

// The row-table ring through a 32-bit shared address (the generic form makes the compiler rebuild the shared window
// base -- S2UR + ULEA -- in front of every access of the chunk-boundary code).
__device__ __forceinline__ void ring16_put_at(uint32_t ring_addr, int row, uint2 v)
{
    const int k = row & (kRing16 - 1);
    sts_u2(ring_addr + 8u * k, v.x, v.y);
    if (k < kMaxChunkRows) sts_u2(ring_addr + 8u * (kRing16 + k), v.x, v.y);
}

// Slow path of the hand-off poll (every 65536th stale poll): true when the run was aborted or this warp has
// waited longer than timeout_ns, in which case the abort flag is raised for everybody else.
static __device__ __noinline__ bool psa_handoff_expired(int* abort_flag, unsigned long long timeout_ns, unsigned long long& t_wait, int lane)
{
    if (!abort_flag) return false;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t_wait == 0) t_wait = now;
    int stop = *reinterpret_cast<volatile int*>(abort_flag);
    if (now - t_wait > timeout_ns) { stop = 1; if (lane == 0) atomicExch(abort_flag, 1); }
    return __any_sync(kFullMask, stop != 0);
}

// ---- streamed left boundary (score-only single pair; see TSTA_STREAM_IN below) ---------------------------------
// rows [row0, row0 + 8) of a boundary column -> ring in shared memory (row0 a multiple of 8; the columns start on
// 16-byte boundaries and are padded by kBndPadRows, rows >= m are copied but never used); one copy group per call
__device__ __forceinline__ void psa_stream_issue(const unsigned long long* bnd_in, const int row0, const uint32_t sin_addr, const int lane)
{
    if (lane < 4) {
        const int r = row0 + 2 * lane;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sin_addr + 8u * (uint32_t)(r & (kStreamRing - 1))), "l"(bnd_in + r) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// Cold path: the group [row0, row0 + 8) was copied before its producer had written all of it -- this band has caught
// up with its left neighbour.  Wait until the producer is `hyst` rows past the group (or done), copy the window
// (the group and the `ahead` groups behind it) again and check again.  Out of line: the unrolled chunk stays small.
static __device__ __noinline__ void psa_stream_refill(const unsigned long long* bnd_in, const int row0, const int m, const unsigned epoch,
                                                      const uint32_t sin_addr, const int lane, const int ahead, const int hyst,
                                                      const bool watchdog, int* abort_flag, const unsigned long long timeout_ns)
{
    const unsigned long long* const pw = bnd_in + min(row0 + 7 + hyst, m - 1);
    const int row = row0 + (lane & 7);
    unsigned spins = 0;
    unsigned long long t_wait = 0;
    for (;;) {
        while ((unsigned)(ld_relaxed_u64(pw) >> 32) != epoch) {
            if (watchdog) {                                               // the producer is ANOTHER GPU: never hang on it
                __nanosleep(20);
                if ((++spins & 0xfffu) == 0 && psa_handoff_expired(abort_flag, timeout_ns, t_wait, lane)) return;
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");                  // (older copies must not land on top of the new ones)
        __syncwarp();
        for (int g = 0; g <= ahead; g++) psa_stream_issue(bnd_in, row0 + 8 * g, sin_addr, lane);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const uint32_t tag = lds_u32(sin_addr + 8u * (uint32_t)(row & (kStreamRing - 1)) + 4u);
        if (__all_sync(kFullMask, row >= m || tag == epoch)) return;
    }
}

// Per-warp shared-memory workspace of the packed sweep.
template <int C, bool RING>
struct __align__(16) PsaWarpSmem16 {
    // RING (TM == 1): [step & 127][lane] direction words of one step: C <= 4: one word (low 16 bits =
    // the low virtual lane's C codes, high 16 bits = the high one's); C == 8: two words (low, high)
    uint32_t dirw[RING ? kDirSteps * 32 * (C == 8 ? 2 : 1) : 4];
    // [r & 127] = (table of row r, table of row r-32), 4 x int8 each; the first 32 entries are
    // mirrored behind the ring so that "entry of row r0 + s" is ring[(r0 & 127) + s] with a
    // compile-time s (no wrap inside a chunk)
    uint2 rowring[kRing16 + kMaxChunkRows];
    uint32_t in[kMaxChunkRows];  // left boundary of the current chunk: lo16 = H, hi16 = Ft, relative to base
    uint2 in2[kMaxChunkRows];    // dual-pair mode: (H of pair A | pair B, F state of pair A | pair B)
    uint2 sin[kStreamRing];      // streamed left boundary: tagged elements of rows r (slot r & 31), copied in by cp.async
};

// Dual-pair mode (short reads): the skew is 31 steps instead of 63, so with 16-row chunks a ring of
// 64 steps is enough -- half the shared memory, twice the warps per SM for the traced kernels.
constexpr int kDirStepsDual = 64;
template <int C, bool TRACE>
struct __align__(16) PsaWarpSmem16D {
    uint32_t dirw[TRACE ? kDirStepsDual * 32 * (C == 8 ? 2 : 1) : 4];
    uint2 rowring[kRing16 + kMaxChunkRows];
    uint2 in2[kMaxChunkRows];
};

// Per-lane register state.
template <int C>
struct PsaLane16 {
    uint32_t Hup[C];     // H(i-1, .)            both halves relative to `base`
    uint32_t Et[C];      // Eg(i, .) - (O+E)
    uint32_t sel[C];     // PRMT selector: lo nibbles pick from the low row's table, hi from the high row's
    uint32_t diag;       // H(i-1, j0-1)
    uint32_t Hout, Fout; // right edge of the previous step (H, Ft)
    uint32_t ms;         // running maximum since the last re-centring
    uint2 Tnext;
    uint32_t cEb[C];     // TRACE: bit 0 / bit 16 = cE of the NEXT row of this column
};

// Does the packed form fit?  D1 = largest difference between neighbouring cells (M + |O+E|),
// span = rows + columns a warp keeps live, S = score scale (2 with traceback).
__host__ __device__ inline int psa16_pad_low(int M, int X, int E, int O, int S)
{
    // how far below its true value the diagonal input of a pad column starts on row 0 (see
    // psa_sweep_band16): enough to stay below "real neighbour + X" whatever the pad's own score is
    const int aM = M < 0 ? -M : M, aX = X < 0 ? -X : X;
    return S * (aM + aX - 2 * (O + E) + 2);
}
__host__ __device__ inline bool psa16_fits(int M, int X, int E, int O, int S, int C, int nsym)
{
    if (nsym > 4) return false;
    const int aM = M < 0 ? -M : M, aX = X < 0 ? -X : X;
    if (S * aM + 3 > 127 || S * aX + 3 > 127) return false;
    const long long D1 = (long long)(M > 0 ? M : 0) - (O + E);          // |H(i,j) - H(i,j-1)| <= M + |O+E|
    const long long span = 64LL * C + 64 + 2 * kMaxChunkRows + 16;      // columns + skew + chunk drift
    return (long long)S * (D1 * span + aM + aX - 2LL * (O + E)) + psa16_pad_low(M, X, E, O, S) + 64 <= 30000;
}

// Scoring constants of a sweep, packed in both halves.
struct PsaConst16 {
    uint32_t E2;      // score-only: E;    traced: 4E-1
    uint32_t OE2;     // score-only: O+E;  traced: 4(O+E)+2 ("opened")
    uint32_t Kf2;     // TM 1: max(4E, 4(O+E)+2) - 2;  TM 2: 4E-2
    uint32_t E1;      // traced: 4E+1
};
constexpr uint32_t kM1 = 0x00010001u, kM2 = 0x00020002u, kM3 = 0x00030003u;
// Traced BATCHES: 3 = the codes of TM 2 (14.5 ALU-pipe instructions per packed column; wf / we are read at the cell a
// walk enters) through the step ring of TM 1; 1 = the round-1 form (17.5; cF / cE are read at the cell a walk leaves).
#ifndef TSTA_BATCH_TM
#define TSTA_BATCH_TM 3
#endif
constexpr int kBatchTM = TSTA_BATCH_TM;
static_assert(kBatchTM == 1 || kBatchTM == 3, "trace mode of the batched kernels");
constexpr uint32_t kFakeTab = 0x80808080u;   // substitution table of the rows above row 0: -128 for every symbol

// TM (trace mode): 0 score-only, 1 traced through the shared-memory step ring (batches), 2 traced with
// the skewed direct-to-HBM layout (one long pair).
// CK: 0 steady state (no predication), 1 ramp chunks with full masking (lanes outside [0, m) keep
// their state), 2 ramp-IN chunks of the fixed-point start-up (only the running maximum and the
// boundary export are masked).
__device__ __forceinline__ void stg_cs_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_cs_u2(uint32_t* p, uint32_t x, uint32_t y)
{
    asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}

// One wavefront step: row t-k of the low virtual lane and row t-k-32 of the high one.
//   in_addr   sm.in of this chunk (entry s = the band's left boundary of row t)
//   ring_next ring entry of this lane's NEXT row at s = 0
//   bnd_row   boundary element of the row virtual lane 63 works on at s = 0
//   dir_addr  TM 1: this lane's direction word of step s = 0 in the step ring
//   dir_g     TM 2: this lane's direction word of step s = 0 in HBM ([band][step][lane])
//   DUAL      the two halves are two independent PAIRS (short reads): no low -> high hand-off in lane 0,
//             both halves work on row t-k, the boundary of both comes from sm.in2, m_hi = rows of pair B
template <int C, int TM, int CK, bool DUAL = false, bool INREG = false>
__device__ __forceinline__ void psa_step16(PsaLane16<C>& L, const int t, const int s, const int m, const int lane,
                                           const PsaConst16& k, const uint32_t selH, const uint32_t selF,
                                           const int src_lane, const uint32_t in_addr, const uint32_t ring_next,
                                           unsigned long long* bnd_row, const bool store_right,
                                           const uint32_t outHF, const unsigned epoch, const uint32_t dir_addr,
                                           uint32_t* dir_g = nullptr, const int m_hi = 0, const uint32_t in_reg = 0)
{
    constexpr bool CHECK = CK == 1;
    const uint32_t rH = __shfl_sync(kFullMask, L.Hout, src_lane);
    const uint32_t rF = __shfl_sync(kFullMask, L.Fout, src_lane);
    uint32_t in, inF;                                             // broadcast: the band's left boundary
    if (DUAL) { const uint2 v = lds_u2(in_addr + s * 8); in = v.x; inF = v.y; }
    else if (INREG) { in = in_reg; inF = in; }                   // streamed boundary: the caller holds the element
    else { in = lds_u32(in_addr + s * 4); inF = in; }
    const uint2 T = L.Tnext;
    L.Tnext = lds_u2(ring_next + s * 8);

    uint32_t mask = 0xffffffffu;
    if (CK != 0) {
        const bool lo = CK == 2 ? t >= lane : (unsigned)(t - lane) < (unsigned)m;
        const bool hi = DUAL ? (unsigned)(t - lane) < (unsigned)m_hi
                             : (CK == 2 ? t >= lane + 32 : (unsigned)(t - lane - 32) < (unsigned)m);
        mask = (lo ? 0xffffu : 0u) | (hi ? 0xffff0000u : 0u);
    }
    uint32_t Hn[C];
    uint32_t Hin, F;

    if constexpr (TM == 0) {
        // ---- F-independent pre-work ---------------------------------------------------------
        uint32_t hE[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t sub = prmtu(T.x, T.y, L.sel[c]);
            const uint32_t X = vadd2(c == 0 ? L.diag : L.Hup[c - 1], sub);
            hE[c] = vaddmax2(L.Et[c], k.OE2, X);                  // max(Eg, h)
        }
        Hin = prmtu(rH, in, selH);
        F = prmtu(rF, inF, selF);
        // ---- serial chain + epilogue --------------------------------------------------------
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t Fc = F;
            F = vaddmax2(Fc, k.E2, hE[c]);                        // Ft(i, j+1): the only serial op
            Hn[c] = vaddmax2(Fc, k.OE2, hE[c]);                   // H = max(Fg, max(Eg, h))
            const uint32_t En = vaddmax2(L.Et[c], k.E2, Hn[c]);
            L.Et[c] = CHECK ? bsel(mask, En, L.Et[c]) : En;
        }
    } else if constexpr (TM == 1) {
        // ---- F-independent pre-work ---------------------------------------------------------
        // tags in the two low bits: 4h+3 (the table entries carry the +3), 4Fg+2, 4Eg+1
        uint32_t hEt[C], Ec1[C], g[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t hs3 = vadd2(c == 0 ? L.diag : L.Hup[c - 1], prmtu(T.x, T.y, L.sel[c]));
            Ec1[c] = (L.Et[c] & ~kM3) | kM1;                      // L.Et holds the flagged Eg state here
            hEt[c] = __vmaxs2(hs3, Ec1[c]);                       // max(h, Eg), diagonal wins ties
            g[c] = vadd2(hEt[c] & ~kM3, k.OE2);                   // 4(max(h,Eg)+O+E)+2
        }
        const uint32_t Hraw = prmtu(rH, in, selH);                // bit 0 / 16: cF of this lane's first column
        Hin = Hraw & ~kM3;
        F = prmtu(rF, inF, selF);                                 // flagged Fg state
        uint32_t f1 = Hraw & kM1;
        // ---- serial chain + epilogue --------------------------------------------------------
        uint32_t acc[(C + 3) / 4];
#pragma unroll
        for (int q = 0; q < (C + 3) / 4; q++) acc[q] = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t Fin = F;
            const uint32_t Fc2 = Fin | kM2;                       // 4Fg+2
            F = vaddmax2(Fc2, k.Kf2, g[c]);                       // max(4a, 4t+2)   (Kf2 holds Kf - 2)
            const uint32_t Hp = __vmaxs2(hEt[c], Fc2);            // winner tagged 3 / 2 / 1: psa.c:294-298
            Hn[c] = Hp & ~kM3;
            const uint32_t T2 = vadd2(Hn[c], k.OE2);              // 4(H+O+E)+2
            const uint32_t En = vaddmax2(Ec1[c], k.E2, T2);       // (E2 holds 4E - 1)
            const uint32_t WE = vaddmax2(L.Et[c], k.E1, T2);      // bit 0: extending Eg wins or ties-after-open
            const uint32_t WF = vaddmax2(Fin, k.E1, T2);          // same for Fg: cF of column j+1
            // code = back | cF << 2 | cE << 3, at nibble c of each half
            const uint32_t x = f1 * 4u + (Hp & kM3) + L.cEb[c] * 8u;
            acc[c / 4] += x << (4 * (c & 3));
            f1 = WF & kM1;
            const uint32_t e1 = WE & kM1;
            L.Et[c] = CHECK ? bsel(mask, En, L.Et[c]) : En;
            L.cEb[c] = CHECK ? bsel(mask, e1, L.cEb[c]) : e1;
        }
        // direction words of this step -> step ring (garbage of inactive halves is never flushed)
        if (C == 8) sts_u2(dir_addr + s * 256, prmtu(acc[0], acc[1], 0x5410u), prmtu(acc[0], acc[1], 0x7632u));
        else sts_u32(dir_addr + s * 128, acc[0]);
        // the cF flag of the next lane's first column rides in bit 0 of the H hand-off
        L.Hout = f1;                                              // merged below
    } else {
        // ---- TM 2 / TM 3: everything a cell's code needs is known in the cell itself ------------
        uint32_t hs3[C], Ec1[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            hs3[c] = vadd2(c == 0 ? L.diag : L.Hup[c - 1], prmtu(T.x, T.y, L.sel[c]));   // 4h+3
            Ec1[c] = (L.Et[c] & ~kM3) | kM1;                      // 4Eg+1
        }
        Hin = prmtu(rH, in, selH);                                // multiples of four
        F = prmtu(rF, inF, selF);                                 // flagged Fg state
        uint32_t acc[(C + 3) / 4];
#pragma unroll
        for (int q = 0; q < (C + 3) / 4; q++) acc[q] = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const uint32_t Fin = F;
            const uint32_t Fc2 = Fin | kM2;                       // 4Fg+2
            const uint32_t Hp = vmax3_2(hs3[c], Fc2, Ec1[c]);     // winner tagged 3 / 2 / 1: psa.c:294-298
            Hn[c] = Hp & ~kM3;
            const uint32_t T2 = vadd2(Hn[c], k.OE2);              // 4(H+O+E)+2
            F = vaddmax2(Fc2, k.Kf2, T2);                         // max(4(Fg+E), 4(H+O+E)+2)
            const uint32_t En = vaddmax2(Ec1[c], k.E2, T2);       // max(4(Eg+E), 4(H+O+E)+2)
            const uint32_t WE = vaddmax2(L.Et[c], k.E1, T2);      // bit 0: we of this cell
            const uint32_t WF = vaddmax2(Fin, k.E1, T2);          // bit 0: wf of this cell
            const uint32_t x = (Hp & kM3) + (WF & kM1) * 4u + (WE & kM1) * 8u;
            acc[c / 4] += x << (4 * (c & 3));
            L.Et[c] = CHECK ? bsel(mask, En, L.Et[c]) : En;
        }
        // (halves on rows outside [0, m) store garbage nobody reads)
        if constexpr (TM == 3) {
            // batches: the same codes through the step ring (flushed row-major like TM 1)
            if (C == 8) sts_u2(dir_addr + s * 256, prmtu(acc[0], acc[(C + 3) / 4 - 1], 0x5410u), prmtu(acc[0], acc[(C + 3) / 4 - 1], 0x7632u));
            else sts_u32(dir_addr + s * 128, acc[0]);
        } else if (C > 4) stg_cs_u2(dir_g + s * 64, prmtu(acc[0], acc[(C + 3) / 4 - 1], 0x5410u), prmtu(acc[0], acc[(C + 3) / 4 - 1], 0x7632u));
        else stg_cs_u32(dir_g + s * 32, acc[0]);
    }
    uint32_t msn = L.ms;
#pragma unroll
    for (int c = 0; c + 1 < C; c += 2) msn = vmax3_2(msn, Hn[c], Hn[c + 1]);
    if (C & 1) msn = __vmaxs2(msn, Hn[C - 1]);
    const uint32_t Ho = TM == 1 ? (Hn[C - 1] | L.Hout) : Hn[C - 1];
    if (CHECK) {
#pragma unroll
        for (int c = 0; c < C; c++) L.Hup[c] = bsel(mask, Hn[c], L.Hup[c]);
        L.ms = bsel(mask, msn, L.ms);
        L.diag = bsel(mask, Hin, L.diag);
        L.Hout = bsel(mask, Ho, TM == 1 ? 0u : L.Hout);
        L.Fout = bsel(mask, F, L.Fout);
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) L.Hup[c] = Hn[c];
        L.ms = CK == 2 ? bsel(mask, msn, L.ms) : msn;
        L.diag = Hin;
        L.Hout = Ho;
        L.Fout = F;
    }
    // right edge of the band = high halves of lane 31, row t-63: (H | Fg << 16) absolute modulo 2^16
    if (!DUAL) {
        const uint32_t e = vadd2(prmtu(L.Hout, L.Fout, 0x7632u), outHF);
        const bool doit = store_right && (CK == 0 || (unsigned)(t - 63) < (unsigned)m);
        st_relaxed_pair_if(bnd_row + s, e, epoch, doit);
    }
}

// Writes the finished direction rows [lo, hi) of one band from the step ring to HBM: row r of
// virtual lane v was produced at step r + v.  16-byte units, lane u handles units u, u+32, ...
template <int C>
__device__ __forceinline__ void psa_flush_dirs16(const uint32_t* dirw, const PsaBand& b, int lo, int hi, int lane)
{
    constexpr int kRowBytes = 32 * C;                   // 64*C columns, 4 bits each
    constexpr int kUnits = kRowBytes / 16;
    constexpr int kW = C == 8 ? 2 : 1;                  // words per lane per step
    const int units = (hi - lo) * kUnits;
    for (int u = lane; u < units; u += 32) {
        const int r = lo + u / kUnits, part = u % kUnits;
        uint32_t w[4];
        if (C == 8) {           // 4 bytes per virtual lane: 4 virtual lanes per unit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = part * 4 + q;
                w[q] = dirw[(((r + v) & (kDirSteps - 1)) * 32 + (v & 31)) * kW + (v >> 5)];
            }
        } else if (C == 4) {    // 2 bytes per virtual lane: 8 per unit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v0 = part * 8 + 2 * q, v1 = v0 + 1;
                const uint32_t a = dirw[((r + v0) & (kDirSteps - 1)) * 32 + (v0 & 31)];
                const uint32_t c = dirw[((r + v1) & (kDirSteps - 1)) * 32 + (v1 & 31)];
                const uint32_t la = (v0 >> 5) ? (a >> 16) : (a & 0xffffu), lc = (v1 >> 5) ? (c >> 16) : (c & 0xffffu);
                w[q] = la | (lc << 16);
            }
        } else {                // C == 2: 1 byte per virtual lane: 16 per unit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t x = 0;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int v = part * 16 + 4 * q + e;
                    const uint32_t a = dirw[((r + v) & (kDirSteps - 1)) * 32 + (v & 31)];
                    x |= (((v >> 5) ? (a >> 16) : a) & 0xffu) << (8 * e);
                }
                w[q] = x;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(b.dirs + (long long)r * b.dirs_stride + (long long)b.band * kRowBytes + part * 16);
        __stcs(dst, make_uint4(w[0], w[1], w[2], w[3]));
    }
}

// Dual-pair mode: the finished direction rows [lo, hi) of ONE of the two pairs (half = 0: pair A in
// the low halves, 1: pair B in the high halves); row r of virtual lane k was produced at step r + k.
// A row is 32*C columns = 16*C bytes.
template <int C, int STEPS>
__device__ __forceinline__ void psa_flush_dirs16_dual(const uint32_t* dirw, uint8_t* dirs, int half, int lo, int hi, int lane)
{
    constexpr int kRowBytes = 16 * C;
    constexpr int kUnits = kRowBytes / 16;              // C == 4: 4 units, C == 8: 8 units
    constexpr int kW = C == 8 ? 2 : 1;
    const int units = (hi - lo) * kUnits;
    for (int u = lane; u < units; u += 32) {
        const int r = lo + u / kUnits, part = u % kUnits;
        uint32_t w[4];
        if (C == 8) {           // 4 bytes per lane: 4 lanes per unit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = part * 4 + q;
                w[q] = dirw[(((r + v) & (STEPS - 1)) * 32 + v) * kW + half];
            }
        } else {                // C == 4: 2 bytes per lane: 8 lanes per unit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v0 = part * 8 + 2 * q, v1 = v0 + 1;
                const uint32_t a = dirw[((r + v0) & (STEPS - 1)) * 32 + v0];
                const uint32_t c = dirw[((r + v1) & (STEPS - 1)) * 32 + v1];
                w[q] = (half ? (a >> 16) : (a & 0xffffu)) | ((half ? (c >> 16) : (c & 0xffffu)) << 16);
            }
        }
        __stcs(reinterpret_cast<uint4*>(dirs + (long long)r * kRowBytes + part * 16), make_uint4(w[0], w[1], w[2], w[3]));
    }
}

// One band (64*C columns) of one pair.  Returns the band's maximum H in absolute (scaled) units.
//   TM    0 score-only, 1 traced through the step ring (b.dirs row-major), 2 traced, skewed layout
//         (b.dirs = this band's [step][lane] words)
//   PIPE  one long pair: bands are separate warps, the left boundary arrives in tagged elements and is
//         imported in groups of kImportGroup rows (short lag behind the producing band)
#ifndef TSTA_IMPORT_GROUP
#define TSTA_IMPORT_GROUP 32
#endif
#ifndef TSTA_SCORE_UNROLL
#define TSTA_SCORE_UNROLL 32       // steps of the score-only steady-state loop that are unrolled (32 = the whole chunk)
#endif
#ifndef TSTA_FAST_RAMP
#define TSTA_FAST_RAMP 1
#endif
#ifndef TSTA_DEBUG_ASSERTS
#define TSTA_DEBUG_ASSERTS 0
#endif
#if TSTA_DEBUG_ASSERTS
#include <assert.h>
#endif
#ifndef TSTA_WATCHDOG
#define TSTA_WATCHDOG 1
#endif
// Streamed left boundary (score-only single pair) -- an EXPERIMENT, off by default (profiles/r02_handoff.md).  Instead
// of importing a whole chunk of boundary rows at the chunk start -- which makes a band wait until its left neighbour
// has finished ALL rows of that chunk -- the rows are copied into a small shared-memory ring by cp.async in groups of
// 8, TSTA_STREAM_AHEAD groups ahead of their use (cp.async.cg: through L2, and tracked by the copy-group counter; a
// register ring of ordinary loads serialises on the ONE scoreboard ptxas gives them: 32 ms).  Every 8th step checks
// the eight tags of the group that starts there (lanes 0..7, one vote); a stale group means this band has caught up
// with its producer: it waits until the producer is TSTA_STREAM_HYST rows ahead and copies the window again (out of
// line).  Measured on the 100 kbp pair: the start lag per band drops from 124 to 105 steps, but the check at every
// group boundary (LDS -> ISETP -> VOTE -> BRA, ~100 cycles that nothing overlaps) makes a step 91.6 instead of 77.7
// cycles: 7.26 ms against 6.19 ms.  The lag that is left is the skew (63 steps) plus the slow start-up chunks.
#ifndef TSTA_STREAM_IN
#define TSTA_STREAM_IN 0
#endif
#ifndef TSTA_STREAM_AHEAD
#define TSTA_STREAM_AHEAD 1
#endif
#ifndef TSTA_STREAM_HYST
#define TSTA_STREAM_HYST 12
#endif
constexpr int kImportGroup = TSTA_IMPORT_GROUP;
constexpr int kStreamAhead = TSTA_STREAM_AHEAD;   // groups of 8 rows requested ahead of the group in use
// boundary columns start on 16-byte boundaries (the streamed import copies two elements at a time)
__host__ __device__ inline long long psa_bnd_stride(long long m) { return (m + 1) & ~1LL; }
constexpr int kBndPadRows = 128;      // rows the streamed import may request past the last step (m + 63 + depth) of a boundary column
static_assert(kStreamAhead >= 1 && kStreamAhead <= 2, "ring of 32 rows: the group in use, the groups ahead, and one being overwritten");
// Per-band time stamps (tools/timeline.py): compiled in only for the experiment build (-DTSTA_TIMELINE=1).
#ifndef TSTA_TIMELINE
#define TSTA_TIMELINE 0
#endif
constexpr bool kTimeline = TSTA_TIMELINE != 0;
template <int C, int TM, bool PIPE, int CH>
__device__ __forceinline__ int psa_sweep_band16(const PsaBand& b, const PsaScoring& sc, const uint32_t* tab4,
                                                PsaWarpSmem16<C, TM == 1 || TM == 3>& sm, const int lane)
{
    constexpr bool TRACE = TM != 0;
    constexpr int S = TRACE ? 4 : 1;
    constexpr int G = kImportGroup < CH ? kImportGroup : CH;
    static_assert(CH == 16 || CH == 32, "chunk rows");
    static_assert((TM != 1 && TM != 3) || C == 2 || C == 4 || C == 8, "ring-traced strips pack to 8 / 16 / 32 bits per virtual lane");
    static_assert(TM != 3 || !PIPE, "TM 3 is the batch form (codes of TM 2, step ring of TM 1)");
    static_assert(TM != 2 || C <= 8, "skewed direction words hold at most 8 columns per virtual lane");
    static_assert(TM != 2 || PIPE, "the skewed layout belongs to the single-pair kernel");
    const int m = b.m;
    const int Es = S * sc.E, OEs = S * (sc.O + sc.E) + (TRACE ? 2 : 0);
    PsaConst16 k;
    // traced: the candidates' tags are folded into the constants: Eg+1 extends by 4E-1, Fg+2 by Kf-2
    k.E2 = dup16(TRACE ? Es - 1 : Es); k.OE2 = dup16(OEs); k.E1 = dup16(Es + 1);
    k.Kf2 = dup16(TM >= 2 ? Es - 2 : max(Es, OEs) - 2);
    const int shiftOE = TRACE ? 0 : OEs;                         // score-only: gap states are carried minus (O+E)
    const int padlow = psa16_pad_low(sc.M, sc.X, sc.E, sc.O, S);
    const int jlo = (b.band * 64 + lane) * C, jhi = jlo + 32 * C;

    // ---- initial state: row -1 (psa.c:219-227), relative to base = H(-1, first column of the band)
    const int goff = b.col_off;                                  // global column of the slice (0 unless multi-GPU)
    int base = S * (sc.O + (goff + b.band * 64 * C + 1) * sc.E);
    PsaLane16<C> L;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int a = jlo + c, h = jhi + c;
        const uint32_t cl = a < b.n ? sc.lut[b.s0[a]] : 8u;            // 8 = pad: sign byte of entry 0 (0 or -1)
        const uint32_t chh = h < b.n ? 4u + sc.lut[b.s0[h]] : 12u;
        const uint32_t nl = cl < 8u ? (cl | ((8u | cl) << 4)) : 0x88u;
        const uint32_t nh = chh < 12u ? (chh | ((8u | chh) << 4)) : 0xccu;
        L.sel[c] = nl | (nh << 8);
        const int ha = S * (sc.O + (goff + a + 1) * sc.E), hh = S * (sc.O + (goff + h + 1) * sc.E);   // H(-1, j)
        // Eg(0,j) = H(-1,j) + O + E, opened.  score-only: stored minus (O+E); TRACE: flagged
        const int eo = OEs - shiftOE;
        L.Et[c] = pack16(ha + eo - base, hh + eo - base);
        L.cEb[c] = 0;
        // Row -1 is only ever read as the DIAGONAL input of column j+1.  Pad columns (j >= n)
        // score 0 or -1 instead of X (the PRMT table has no spare entry), so their diagonal
        // input on row 0 starts `padlow` lower: every pad cell then stays below the real cell
        // it derives from, and the running maximum may include pad cells (the reference
        // includes its own pad columns too, psa.c:333).
        L.Hup[c] = pack16(ha - base - (a + 1 >= b.n ? padlow : 0), hh - base - (h + 1 >= b.n ? padlow : 0));
    }
    {
        const int da = S * (goff + jlo == 0 ? 0 : sc.O + (goff + jlo) * sc.E), dh = S * (sc.O + (goff + jhi) * sc.E);   // H(-1, j0-1)
        L.diag = pack16(da - base - (jlo >= b.n ? padlow : 0), dh - base - (jhi >= b.n ? padlow : 0));
    }
    L.ms = kMin16x2;
    {
        // right edge of row -1 as this lane would hand it on: H(-1, last column) and the F state of the
        // column after it.  The mask-free start-up (CK == 2) relies on it; masked start-ups ignore it.
        const int la = goff + jlo + C - 1, lh = goff + jhi + C - 1;              // last columns (global)
        const int Ha = S * (sc.O + (la + 1) * sc.E), Hh = S * (sc.O + (lh + 1) * sc.E);
        // F state: max(Fg + E, H + O + E) of row -1, where Fg(-1, j) = H(-1, j)
        const int Fa = max(Ha + Es, Ha + OEs) - shiftOE, Fh = max(Hh + Es, Hh + OEs) - shiftOE;
        L.Hout = pack16(Ha - base, Hh - base);
        L.Fout = pack16(Fa - base, Fh - base);
    }
    int ms32 = INT_MIN;
    const int t_begin = b.ch_begin * CH;
    if constexpr (TM == 2) {
        if (b.ckpt_in) {
            // resume from the score-only pass's lane registers at the start of chunk ch_begin: scores times four;
            // a gap state (carried minus O+E there) is "opened" iff it equals the H it was derived from
            const uint32_t* ck = b.ckpt_in + lane;
            const uint32_t oe1 = dup16(sc.O + sc.E);
            auto up4 = [](uint32_t h) { return (h & 0x3fff3fffu) << 2; };
            auto gap4 = [&](uint32_t g, uint32_t h) { return vadd2(up4(vadd2(g, oe1)), __vcmpeq2(g, h) & kM2); };
#pragma unroll
            for (int c = 0; c < C; c++) {
                const uint32_t h = __ldcs(ck + c * 32), e = __ldcs(ck + (C + c) * 32);
                L.Hup[c] = up4(h);
                L.Et[c] = gap4(e, h);
            }
            const uint32_t ho = __ldcs(ck + (2 * C + 1) * 32), fo = __ldcs(ck + (2 * C + 2) * 32);
            L.diag = up4(__ldcs(ck + 2 * C * 32));
            L.Hout = up4(ho);
            L.Fout = gap4(fo, ho);
            base = 4 * (int)__ldcs(ck + (2 * C + 3) * 32);
        }
    }

    const int nsteps = m + 63;                                   // virtual lane 63 finishes row m-1 at step m+62
    const int nchunks = (nsteps + CH - 1) / CH;
    const bool export_right = (b.band + 1 < b.nbands || b.has_right) && b.bnd_out != nullptr;
    const bool import_left = b.band > 0 || b.has_left;
    // Debug build (-DTSTA_DEBUG_ASSERTS=1; compute-sanitizer is not available on the pool): the batched sweep parks a
    // band's right boundary IN PLACE in the column the next band reads (row r is prefetched one chunk before it is
    // needed and rewritten at step r + 63); there the elements carry band + 1 as their tag, so that a row that was
    // overwritten before it was read -- or read before it was written -- trips the assert below.
    const unsigned epoch = PIPE ? b.epoch : (TSTA_DEBUG_ASSERTS ? (unsigned)b.band + 1u : 1u);
    const bool store_right = export_right && lane == 31;
    const uint32_t in_addr = (uint32_t)__cvta_generic_to_shared(sm.in);
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(sm.rowring);
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab4);
    auto tab_of = [&](uint32_t byte) { return lds_u32(tab_addr + 4u * (byte & 0xffu)); };
    const uint32_t dirw_addr = (uint32_t)__cvta_generic_to_shared(sm.dirw) + lane * (C == 8 ? 8 : 4);
    uint32_t* const dir_lane = TM == 2 ? reinterpret_cast<uint32_t*>(b.dirs) + lane * (C > 4 ? 2 : 1) : nullptr;
    const uint32_t selH = lane == 0 ? 0x1054u : 0x3210u;         // lane 0: lo <- boundary H, hi <- lane 31's low half
    const uint32_t selF = lane == 0 ? 0x1076u : 0x3210u;
    const int src_lane = (lane + 31) & 31;
    const int mrows = m + 32;                                    // rows whose table entry is read (high halves lag by 32)
    // mask-free start-up: single-pair bands without pad columns (the pad trick above lowers row -1,
    // which the fixed-point rows would undo); the band that holds the pads is the last one, nobody waits for it
    const bool fast_ramp = TSTA_FAST_RAMP && PIPE && (b.band + 1) * 64 * C <= b.n && m >= 64 + CH;

    // ---- prologue: row tables of chunk 0; prefetch for chunk 1 / boundary of chunk 0 -----------
    __syncwarp();
    const bool stager = lane < CH;
    if (t_begin == 0) {
        // rows above row 0 (read by lanes whose first row is still to come): the fixed-point table
        sts_u2(ring_addr + 8u * (kRing16 - 32 + lane), kFakeTab, kFakeTab);
        if (stager && lane < mrows) ring16_put_at(ring_addr, lane, make_uint2(lane < m ? tab_of(b.s1[lane]) : 0u, kFakeTab));   // rows < 32: no high row yet
    } else {
        // resumed sweep (t_begin >= 128): the steps of the first chunk read the entries of rows t_begin-31 .. t_begin+31
        const int ra = t_begin - 32 + lane, rb = t_begin + lane;
        ring16_put_at(ring_addr, ra, make_uint2(ra < m ? tab_of(b.s1[ra]) : 0u, ra - 32 < m ? tab_of(b.s1[ra - 32]) : 0u));
        ring16_put_at(ring_addr, rb, make_uint2(rb < m ? tab_of(b.s1[rb]) : 0u, rb - 32 < m ? tab_of(b.s1[rb - 32]) : 0u));
    }
    uint32_t raw_next = (stager && t_begin + CH + lane < m) ? b.s1[t_begin + CH + lane] : 0xffffffffu;
    // left boundary: PIPE imports it in groups of G rows (lanes 0..G-1 hold the group that is consumed
    // next); sequential mode (batches) reads whole chunks of its own parked column
    constexpr bool kStream = TSTA_STREAM_IN != 0 && PIPE && TM == 0;
    const bool stream = kStream && import_left;
    const uint32_t sin_addr = (uint32_t)__cvta_generic_to_shared(sm.sin);
    if (stream) {
#pragma unroll
        for (int g = 0; g < kStreamAhead; g++) psa_stream_issue(b.bnd_in, t_begin + 8 * g, sin_addr, lane);
    }
    unsigned long long bnd_next = 0;
    if (import_left && !stream && (PIPE ? lane < G : stager) && t_begin + lane < m) bnd_next = ld_relaxed_u64(b.bnd_in + t_begin + lane);
    __syncwarp();
    L.Tnext = lds_u2(ring_addr + 8u * ((t_begin - lane) & (kRing16 - 1)));      // table of this lane's row at the first step (row t_begin - lane)

    // PIPE: make rows [row0, row0 + G) of the left boundary available in sm.in[slot0 ..] and prefetch the next group
    // (epoch, m and the column pointer pinned in registers: as kernel parameters they are re-read from the constant
    // bank -- LDCU, waited for -- at every import, in the middle of the chunk boundary's serial chain)
#ifndef TSTA_IMPORT_PIN
#define TSTA_IMPORT_PIN 1
#endif
    unsigned epoch_pin = epoch;
    int m_pin = m;
    const unsigned long long* bin_pin = b.bnd_in;
    uint32_t in_lane = in_addr + 4u * (uint32_t)lane;
    if (TSTA_IMPORT_PIN && PIPE) {
        // (through shared memory: an empty asm leaves no PTX, and ptxas folds a mov -- or a shuffle of a uniform value --
        // back into the constant-bank operand; sm.in is not in use yet)
        const unsigned long long pb = (unsigned long long)bin_pin;
        __syncwarp();
        if (lane == 0) {
            sts_u32(in_addr, epoch_pin); sts_u32(in_addr + 4, (uint32_t)m_pin);
            sts_u32(in_addr + 8, (uint32_t)pb); sts_u32(in_addr + 12, (uint32_t)(pb >> 32));
        }
        __syncwarp();
        epoch_pin = lds_u32(in_addr);
        m_pin = (int)lds_u32(in_addr + 4);
        bin_pin = reinterpret_cast<const unsigned long long*>(((unsigned long long)lds_u32(in_addr + 12) << 32) | lds_u32(in_addr + 8));
        __syncwarp();
        sts_u32(in_lane, in_lane);                     // (each lane its own slot: no S2UR + ULEA rebuild of the shared window per import)
        __syncwarp();
        in_lane = lds_u32(in_lane);
        __syncwarp();
    }
    // the wait for rows that are not there yet (start-up, or a producer that was held up): off the straight path
    auto import_poll = [&](const int row, const bool mine) {
        if (TSTA_WATCHDOG && b.has_left && b.band == 0) {
            // the producer is ANOTHER GPU (column slices): never hang on a neighbour that does not come
            unsigned spins = 0;
            unsigned long long t_wait = 0;
            while (!__all_sync(kFullMask, !mine || (unsigned)(bnd_next >> 32) == epoch_pin)) {
                __nanosleep(20);
                if (mine && (unsigned)(bnd_next >> 32) != epoch_pin) bnd_next = ld_relaxed_u64(bin_pin + row);
                if ((++spins & 0xfffu) == 0 && psa_handoff_expired(b.abort_flag, b.timeout_ns, t_wait, lane)) break;
            }
        } else {
            // the producer is a band of this launch that took its ticket earlier: it runs, or has run, whatever this
            // warp does -- the wait always ends.  No watchdog here: any code inside this loop delays every band
            // behind this one (measured on the 100 kbp pair: 6.6 ms with a check every 4096 polls, 6.2 ms without)
            do {
                __nanosleep(20);
                if (mine && (unsigned)(bnd_next >> 32) != epoch_pin) bnd_next = ld_relaxed_u64(bin_pin + row);
            } while (!__all_sync(kFullMask, !mine || (unsigned)(bnd_next >> 32) == epoch_pin));
        }
    };
    auto import_group = [&](const int row0, const int slot0, const int cur_base) {
        const int row = row0 + lane;
        const bool mine = lane < G && row < m_pin;
        const long long dbg_c0 = (kTimeline && b.timeline) ? clock64() : 0;
        if (!__all_sync(kFullMask, !mine || (unsigned)(bnd_next >> 32) == epoch_pin)) import_poll(row, mine);
#if TSTA_DEBUG_ASSERTS
        if (mine && !(b.abort_flag && *reinterpret_cast<volatile int*>(b.abort_flag))) assert((unsigned)(bnd_next >> 32) == epoch);
#endif
        // absolute modulo 2^16 -> relative to this warp's base (score-only: Ft = Fg - (O+E))
        if (TM == 2 && b.bnd_unscaled) {
            // the column was written by the score-only pass (unscaled H | Fg): times four, and Fg is "opened"
            // iff it equals H + O + E
            const uint32_t raw = (uint32_t)bnd_next;
            const uint32_t rel = vadd2(raw, dup16(-(cur_base >> 2)));
            const uint32_t opened = (((raw >> 16) - raw - (uint32_t)(sc.O + sc.E)) & 0xffffu) == 0u ? 0x00020000u : 0u;
            if (mine) sts_u32(in_addr + 4u * (slot0 + lane), ((rel & 0x3fff3fffu) << 2) + opened);
        } else if (mine) sts_u32(in_lane + 4u * slot0, vadd2((uint32_t)bnd_next, pack16(-cur_base, -cur_base - shiftOE)));
        __syncwarp();
        if (lane < G && row + G < m_pin) bnd_next = ld_relaxed_u64(bin_pin + row + G);
        if (kTimeline && b.timeline && lane == 0) {   // diagnostics: polls that found stale rows, cycles spent importing, imports
            b.timeline[b.band * 16 + 8] += 0;   // (stale polls are no longer counted)
            b.timeline[b.band * 16 + 9] += (unsigned long long)(clock64() - dbg_c0);
            b.timeline[b.band * 16 + 10] += 1;
        }
    };

    auto stamp = [&](int slot) {
        if (kTimeline && PIPE && b.timeline && lane == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            b.timeline[b.band * 16 + slot] = now;
        }
    };
    stamp(0);
    if (kTimeline && PIPE && b.timeline && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        b.timeline[b.band * 16 + 11] = smid;
    }
    const int ch_last = min(nchunks, b.ch_end);
    for (int ch = b.ch_begin; ch < ch_last; ch++) {
        const int t0 = ch * CH;
        if (ch >= 1 && ch <= 5) stamp(1 + ch);                      // 2..6: start of chunks 1..5
        // ---- re-centre on lane 0's current row -------------------------------------------------
        {
            int delta = lo16(__shfl_sync(kFullMask, L.Et[0], 0));   // column j0 of the band is never a pad column
            if (TRACE) delta &= ~3;
            ms32 = max(ms32, max(lo16(L.ms), hi16(L.ms)) + base);   // (stale maxima of masked halves are harmless: they were real)
            const uint32_t nd = dup16(-delta);
#pragma unroll
            for (int c = 0; c < C; c++) { L.Hup[c] = vadd2(L.Hup[c], nd); L.Et[c] = vadd2(L.Et[c], nd); }
            L.diag = vadd2(L.diag, nd);
            L.Hout = vadd2(L.Hout, nd);
            L.Fout = vadd2(L.Fout, nd);
            L.ms = kMin16x2;
            base += delta;
        }
        // ---- checkpoint (score-only pass of a traced pair): lane registers at the start of every 128th step ----
        if constexpr (TM == 0 && PIPE && CH == 32) {
            if (b.ckpt_out && ch > 0 && (ch % kCkptChunks) == 0) {
                uint32_t* ck = b.ckpt_out + (size_t)(ch / kCkptChunks) * (kCkptWords<C> * 32) + lane;
#pragma unroll
                for (int c = 0; c < C; c++) { __stcs(ck + c * 32, L.Hup[c]); __stcs(ck + (C + c) * 32, L.Et[c]); }
                __stcs(ck + 2 * C * 32, L.diag);
                __stcs(ck + (2 * C + 1) * 32, L.Hout);
                __stcs(ck + (2 * C + 2) * 32, L.Fout);
                __stcs(ck + (2 * C + 3) * 32, (uint32_t)base);
            }
        }
        // ---- stage: next chunk's row tables, this chunk's left boundary -------------------------
        {
            const int rown = t0 + CH + lane;                    // row of the next chunk
            if (stager && rown < mrows) {
                const uint32_t hi_tab = rown >= 32 ? lds_u32(ring_addr + 8u * ((rown - 32) & (kRing16 - 1))) : kFakeTab;
                ring16_put_at(ring_addr, rown, make_uint2(rown < m ? tab_of(raw_next) : 0u, hi_tab));
            }
            const int rownn = rown + CH;
            raw_next = (stager && rownn < m) ? b.s1[rownn] : 0u;
            const int row = stager ? t0 + lane : m;
            if (stream) {
                // nothing to stage: the steps read the boundary from the streamed ring
            } else if (import_left && PIPE) {
                if (t0 < m) import_group(t0, 0, base);          // (also orders the ring stores above)
                if (ch == 0) stamp(1);                          // the first rows of the left boundary have arrived
            } else if (import_left) {
#if TSTA_DEBUG_ASSERTS
                if (row < m) assert((unsigned)(bnd_next >> 32) == (unsigned)b.band);      // written by band - 1 of THIS pair
#endif
                if (row < m) sts_u32(in_addr + 4u * lane, vadd2((uint32_t)bnd_next, pack16(-base, -base - shiftOE)));
                const int rowp = row + CH;
                if (stager && rowp < m) bnd_next = ld_relaxed_u64(b.bnd_in + rowp);
            } else if (row < m) {
                const int hb = S * (sc.O + (row + 1) * sc.E) - base;   // H(i,-1); Fg(i,0) = H(i,-1)+O+E, opened
                sts_u32(in_addr + 4u * lane, pack16(hb, hb + OEs - shiftOE));
            }
        }
        __syncwarp();
        const uint32_t outHF = pack16(base, base + shiftOE);      // relative (H, F state) -> absolute
        const uint32_t ring_next = ring_addr + (((t0 + 1 - lane) & (kRing16 - 1)) << 3);
        unsigned long long* bnd_row = b.bnd_out + (t0 - 63);
        const uint32_t dir_addr = dirw_addr + (t0 & (kDirSteps - 1)) * (C == 8 ? 256 : 128);
        uint32_t* const dir_g = TM == 2 ? dir_lane + ((long long)t0 - b.dir_step0) * (C > 4 ? 64 : 32) : nullptr;
        const bool groups = import_left && PIPE && !stream;
        // streamed boundary: every 8th step requests the group kStreamAhead groups ahead, makes sure the group that
        // starts here has landed and checks its eight tags (lanes 0..7, one vote); the steps read their element's
        // value from the ring and re-centre it on this warp's base
        const uint32_t rebase = pack16(-base, -base - shiftOE);
        uint32_t sin_t0 = sin_addr + 8u * (uint32_t)(t0 & (kStreamRing - 1));   // (a chunk never wraps: CH divides the ring)
        asm volatile("" : "+r"(sin_t0));                                          // keep it in a register (no S2UR rebuild per step)
        auto stream_in = [&](const int s, const bool live) -> uint32_t {
            if ((s & 7) == 0) {
                psa_stream_issue(b.bnd_in, t0 + s + 8 * kStreamAhead, sin_addr, lane);
                asm volatile("cp.async.wait_group %0;" ::"n"(kStreamAhead) : "memory");
                __syncwarp();
                const uint32_t tag = lds_u32(sin_t0 + 8u * (uint32_t)s + 8u * (uint32_t)(lane & 7) + 4u);
                const bool ok = (!live && t0 + s + (lane & 7) >= m) || tag == epoch;
                if (!__all_sync(kFullMask, ok))
                    psa_stream_refill(b.bnd_in, t0 + s, m, epoch, sin_addr, lane, kStreamAhead, TSTA_STREAM_HYST,
                                      TSTA_WATCHDOG && b.has_left && b.band == 0, b.abort_flag, b.timeout_ns);
                if (kTimeline && s == 0 && ch == 0) stamp(1);               // the first rows of the left boundary have arrived
            }
            return vadd2(lds_u32(sin_t0 + 8u * (uint32_t)s), rebase);
        };
        // the steps of one chunk; ST: the left boundary comes from the streamed ring (stream_in) instead of sm.in
        auto run_chunk = [&](auto st_tag) {
            constexpr bool ST = decltype(st_tag)::value;
            if (t0 + CH - 1 < m && (t0 >= 63 || fast_ramp)) {
                // score-only: the whole chunk is unrolled (every address becomes an immediate);
                // the traced step is several times as long, so it is unrolled just enough to overlap
                // neighbouring steps without outgrowing the instruction cache
                if (t0 >= 63) {
                    if constexpr (!TRACE && TSTA_SCORE_UNROLL >= CH) {
#pragma unroll
                        for (int s = 0; s < CH; s++) {
                            if (s > 0 && s % G == 0 && groups) import_group(t0 + s, s, base);
                            uint32_t inr = 0;
                            if constexpr (ST) inr = stream_in(s, true);
                            psa_step16<C, TM, 0, false, ST>(L, t0 + s, s, m, lane, k, selH, selF, src_lane, in_addr, ring_next,
                                                            bnd_row, store_right, outHF, epoch, dir_addr, dir_g, 0, inr);
                        }
                    } else {
                        constexpr int U = G < 8 ? G : 8;
#pragma unroll 1
                        for (int s0 = 0; s0 < CH; s0 += U) {
                            if (s0 > 0 && s0 % G == 0 && groups) import_group(t0 + s0, s0, base);
#pragma unroll
                            for (int q = 0; q < U; q++) {
                                uint32_t inr = 0;
                                if constexpr (ST) inr = stream_in(s0 + q, true);
                                psa_step16<C, TM, 0, false, ST>(L, t0 + s0 + q, s0 + q, m, lane, k, selH, selF, src_lane, in_addr, ring_next,
                                                                bnd_row, store_right, outHF, epoch, dir_addr, dir_g, 0, inr);
                            }
                        }
                    }
                } else {
                    // mask-free start-up chunks (t0 < 63): lanes whose first row is still to come sit on fixed-point rows
                    constexpr int U = G < 8 ? G : 8;
#pragma unroll 1
                    for (int s0 = 0; s0 < CH; s0 += U) {
                        if (s0 > 0 && s0 % G == 0 && groups) import_group(t0 + s0, s0, base);
#pragma unroll
                        for (int q = 0; q < U; q++) {
                            uint32_t inr = 0;
                            if constexpr (ST) inr = stream_in(s0 + q, true);
                            psa_step16<C, TM, 2, false, ST>(L, t0 + s0 + q, s0 + q, m, lane, k, selH, selF, src_lane, in_addr, ring_next,
                                                            bnd_row, store_right, outHF, epoch, dir_addr, dir_g, 0, inr);
                        }
                    }
                }
            } else if constexpr (ST) {
                // masked chunks, streamed: groups of 8 steps so that the ring slots stay compile-time
#pragma unroll 1
                for (int s0 = 0; s0 < CH; s0 += 8) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint32_t inr = stream_in(s0 + q, false);
                        psa_step16<C, TM, 1, false, true>(L, t0 + s0 + q, s0 + q, m, lane, k, selH, selF, src_lane, in_addr, ring_next,
                                                          bnd_row, store_right, outHF, epoch, dir_addr, dir_g, 0, inr);
                    }
                }
            } else {
#pragma unroll 1
                for (int s = 0; s < CH; s++) {
                    if (s > 0 && s % G == 0 && groups && t0 + s < m) import_group(t0 + s, s, base);
                    psa_step16<C, TM, 1>(L, t0 + s, s, m, lane, k, selH, selF, src_lane, in_addr, ring_next,
                                         bnd_row, store_right, outHF, epoch, dir_addr, dir_g);
                }
            }
        };
        if constexpr (kStream) {
            if (stream) run_chunk(std::true_type{});
            else run_chunk(std::false_type{});
        } else run_chunk(std::false_type{});
        __syncwarp();
        if constexpr (TM == 1 || TM == 3) {
            // rows every virtual lane has finished: [t0 - 63, t0 + CH - 63)
            psa_flush_dirs16<C>(sm.dirw, b, min(m, max(0, t0 - 63)), min(m, max(0, t0 + CH - 63)), lane);
            __syncwarp();
        }
    }
    ms32 = max(ms32, max(lo16(L.ms), hi16(L.ms)) + base);
    stamp(7);

    if (b.lastrow) {
#pragma unroll
        for (int c = 0; c < C; c++) {
            b.lastrow[jlo - b.lastrow_col0 + c] = lo16(L.Hup[c]) + base;
            b.lastrow[jhi - b.lastrow_col0 + c] = hi16(L.Hup[c]) + base;
        }
    }
    return ms32;
}

// ---------------------------------------------------------------------------------------
// DUAL-PAIR sweep (score-only): two independent short pairs per warp, pair A in the low halves
// and pair B in the high halves, 32 virtual lanes x C columns each (reads of up to 32*C bases
// fit one band).  Compared with the 64-virtual-lane sweep a 250 bp pair needs 250+31 steps
// shared by two pairs instead of 250+63 for one, and far fewer of them are ramp steps.
// Each half has its own base.  B may be absent (nB == 0 or mB == 0).
// ---------------------------------------------------------------------------------------
template <int C, bool TRACE>
__device__ __forceinline__ void psa_sweep_dual16(const uint8_t* s0A, const int nA, const uint8_t* s1A, const int mA,
                                                 const uint8_t* s0B, const int nB, const uint8_t* s1B, const int mB,
                                                 uint8_t* dirsA, uint8_t* dirsB,
                                                 const PsaScoring& sc, const uint32_t* tab4, PsaWarpSmem16D<C, TRACE>& sm,
                                                 const int lane, int& msA, int& msB, int& cornerA, int& cornerB)
{
    constexpr int CH = TRACE ? 16 : 32;                          // traced: 16-row chunks keep the step ring at 64 entries
    constexpr int kSteps = kDirStepsDual;
    constexpr int S = TRACE ? 4 : 1;
    static_assert(!TRACE || C == 4 || C == 8, "traced dual strips");
    const int Es = S * sc.E, OEs = S * (sc.O + sc.E) + (TRACE ? 2 : 0);
    PsaConst16 k;
    k.E2 = dup16(TRACE ? Es - 1 : Es); k.OE2 = dup16(OEs); k.Kf2 = dup16(kBatchTM >= 2 ? Es - 2 : max(Es, OEs) - 2); k.E1 = dup16(Es + 1);
    const int shiftOE = TRACE ? 0 : OEs;                         // score-only: gap states are carried minus (O+E)
    const int padlow = psa16_pad_low(sc.M, sc.X, sc.E, sc.O, S);
    const int j0 = lane * C;
    int baseA = S * (sc.O + sc.E), baseB = baseA;                // H(-1, 0)
    PsaLane16<C> L;
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int j = j0 + c;
        const uint32_t cl = j < nA ? sc.lut[s0A[j]] : 8u;
        const uint32_t chh = j < nB ? 4u + sc.lut[s0B[j]] : 12u;
        const uint32_t nl = cl < 8u ? (cl | ((8u | cl) << 4)) : 0x88u;
        const uint32_t nh = chh < 12u ? (chh | ((8u | chh) << 4)) : 0xccu;
        L.sel[c] = nl | (nh << 8);
        const int h = S * (sc.O + (j + 1) * sc.E);                // H(-1, j); Eg(0,j) = H(-1,j)+O+E, opened
        const int eo = OEs - shiftOE;
        L.Et[c] = pack16(h + eo - baseA, h + eo - baseB);
        L.Hup[c] = pack16(h - baseA - (j + 1 >= nA ? padlow : 0), h - baseB - (j + 1 >= nB ? padlow : 0));
        L.cEb[c] = 0;
    }
    {
        const int d = S * (j0 == 0 ? 0 : sc.O + j0 * sc.E);       // H(-1, j0-1)
        L.diag = pack16(d - baseA - (j0 >= nA ? padlow : 0), d - baseB - (j0 >= nB ? padlow : 0));
    }
    L.ms = kMin16x2;
    L.Hout = 0;
    L.Fout = 0;
    int ms32A = INT_MIN, ms32B = INT_MIN;

    const int m = max(mA, mB), mmin = min(mA, mB);
    const int nsteps = m + 31;
    const int nchunks = (nsteps + CH - 1) / CH;
    const uint32_t in_addr = (uint32_t)__cvta_generic_to_shared(sm.in2);
    const uint32_t ring_addr = (uint32_t)__cvta_generic_to_shared(sm.rowring);
    const uint32_t dirw_addr = (uint32_t)__cvta_generic_to_shared(sm.dirw) + lane * (C == 8 ? 8 : 4);
    const uint32_t selHF = lane == 0 ? 0x7654u : 0x3210u;        // lane 0: both halves from the boundary, others from lane k-1
    const int src_lane = (lane + 31) & 31;

    __syncwarp();
    const bool stager = lane < CH;
    if (stager && lane < m) ring16_put(sm, lane, make_uint2(lane < mA ? tab4[s1A[lane]] : 0u, lane < mB ? tab4[s1B[lane]] : 0u));
    uint32_t rawA = (stager && CH + lane < mA) ? s1A[CH + lane] : 0u, rawB = (stager && CH + lane < mB) ? s1B[CH + lane] : 0u;
    __syncwarp();
    L.Tnext = sm.rowring[0];

    for (int ch = 0; ch < nchunks; ch++) {
        const int t0 = ch * CH;
        {   // re-centre each half on its own lane-0 column
            const uint32_t e0 = __shfl_sync(kFullMask, L.Et[0], 0);
            int dA = lo16(e0), dB = mB > 0 ? hi16(e0) : 0;
            if (TRACE) { dA &= ~3; dB &= ~3; }
            ms32A = max(ms32A, lo16(L.ms) + baseA);
            ms32B = max(ms32B, hi16(L.ms) + baseB);
            const uint32_t nd = pack16(-dA, -dB);
#pragma unroll
            for (int c = 0; c < C; c++) { L.Hup[c] = vadd2(L.Hup[c], nd); L.Et[c] = vadd2(L.Et[c], nd); }
            L.diag = vadd2(L.diag, nd);
            L.Hout = vadd2(L.Hout, nd);
            L.Fout = vadd2(L.Fout, nd);
            L.ms = kMin16x2;
            baseA += dA;
            baseB += dB;
        }
        {   // next chunk's row tables, this chunk's left boundary (column -1 of both pairs: closed form)
            const int rown = t0 + CH + lane;
            if (stager && rown < m) ring16_put(sm, rown, make_uint2(rown < mA ? tab4[rawA & 0xffu] : 0u, rown < mB ? tab4[rawB & 0xffu] : 0u));
            const int rownn = rown + CH;
            rawA = (stager && rownn < mA) ? s1A[rownn] : 0u;
            rawB = (stager && rownn < mB) ? s1B[rownn] : 0u;
            const int row = t0 + lane;
            if (stager && row < m) {
                const int hb = S * (sc.O + (row + 1) * sc.E);     // H(i,-1); Fg(i,0) = H(i,-1)+O+E, opened
                const int fb = hb + OEs - shiftOE;
                sm.in2[lane] = make_uint2(pack16(hb - baseA, hb - baseB), pack16(fb - baseA, fb - baseB));
            }
        }
        __syncwarp();
        const uint32_t ring_next = ring_addr + (((t0 + 1 - lane) & (kRing16 - 1)) << 3);
        const uint32_t dir_addr = dirw_addr + (t0 & (kSteps - 1)) * (C == 8 ? 256 : 128);
        if (t0 >= 31 && t0 + CH - 1 < mmin) {
            if constexpr (!TRACE) {
#pragma unroll
                for (int s = 0; s < CH; s++)
                    psa_step16<C, TRACE ? kBatchTM : 0, 0, true>(L, t0 + s, s, mA, lane, k, selHF, selHF, src_lane, in_addr, ring_next,
                                                      nullptr, false, 0u, 0u, dir_addr, nullptr, mB);
            } else {
#pragma unroll 2
                for (int s = 0; s < CH; s++)
                    psa_step16<C, TRACE ? kBatchTM : 0, 0, true>(L, t0 + s, s, mA, lane, k, selHF, selHF, src_lane, in_addr, ring_next,
                                                      nullptr, false, 0u, 0u, dir_addr, nullptr, mB);
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < CH; s++)
                psa_step16<C, TRACE ? kBatchTM : 0, 1, true>(L, t0 + s, s, mA, lane, k, selHF, selHF, src_lane, in_addr, ring_next,
                                                 nullptr, false, 0u, 0u, dir_addr, nullptr, mB);
        }
        __syncwarp();
        if (TRACE) {
            // rows every lane has finished: [t0 - 31, t0 + CH - 31)
            psa_flush_dirs16_dual<C, kSteps>(sm.dirw, dirsA, 0, min(mA, max(0, t0 - 31)), min(mA, max(0, t0 + CH - 31)), lane);
            if (mB > 0) psa_flush_dirs16_dual<C, kSteps>(sm.dirw, dirsB, 1, min(mB, max(0, t0 - 31)), min(mB, max(0, t0 + CH - 31)), lane);
            __syncwarp();
        }
    }
    ms32A = max(ms32A, lo16(L.ms) + baseA);
    ms32B = max(ms32B, hi16(L.ms) + baseB);
    msA = warp_max(ms32A);
    msB = warp_max(ms32B);
    // corners: H(m-1, n-1) sits in lane (n-1)/C, column (n-1)%C
    uint32_t hA = L.Hup[0], hB = L.Hup[0];
#pragma unroll
    for (int c = 1; c < C; c++) {
        hA = ((nA - 1) % C) == c ? L.Hup[c] : hA;
        hB = ((nB - 1) % C) == c ? L.Hup[c] : hB;
    }
    cornerA = nA > 0 ? lo16(__shfl_sync(kFullMask, hA, max(nA - 1, 0) / C)) + baseA : 0;
    cornerB = nB > 0 ? hi16(__shfl_sync(kFullMask, hB, max(nB - 1, 0) / C)) + baseB : 0;
}

// 256-entry byte -> 4 x int8 substitution table (codes 0..3) for the packed PRMT.
__device__ __forceinline__ void psa_build_tab4(uint32_t* tab4, const PsaScoring& sc, int S, int tid, int nthreads)
{
    // traced (S == 4): the entries carry the diagonal candidate's tie-break tag (+3)
    const int Ms = S * sc.M + (S == 4 ? 3 : 0), Xs = S * sc.X + (S == 4 ? 3 : 0);
    for (int bb = tid; bb < 256; bb += nthreads) {
        const uint32_t c = sc.lut[bb];
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) v |= (uint32_t)(((uint32_t)k == c ? Ms : Xs) & 0xff) << (8 * k);
        tab4[bb] = v;
    }
}

// ---------------------------------------------------------------------------------------
// one long pair: one warp (= one CTA) per band of 64*C columns
// ---------------------------------------------------------------------------------------
template <int C, int TM, int CH>
__global__ void __launch_bounds__(32) psa_pair16_kernel(const __grid_constant__ PsaPairArgs a)
{
    __shared__ PsaWarpSmem16<C, TM == 1> sm;
    __shared__ uint32_t tab4[256];
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);
    psa_build_tab4(tab4, a.sc, TM != 0 ? 4 : 1, lane, 32);
    __syncwarp();

    PsaBand b;
    b.s0 = a.s0; b.s1 = a.s1; b.n = a.n; b.m = a.m;
    b.band = band; b.nbands = a.nbands;
    b.col_off = a.col_off;
    b.has_left = a.inbox != nullptr;
    b.has_right = a.outbox != nullptr;
    b.bnd_in = band == 0 ? a.inbox : a.bnd + (long long)(band - 1) * a.bnd_stride;
    b.bnd_out = (band == a.nbands - 1 && a.outbox) ? a.outbox : a.bnd + (long long)band * a.bnd_stride;
    b.epoch = a.epoch;
    b.abort_flag = a.abort_flag; b.timeout_ns = a.timeout_ns; b.timeline = a.timeline;
    // TM 2: dirs_stride = bytes of one band's [step][lane] region (psa_skew_band_bytes)
    b.dirs = TM == 2 ? a.dirs + (long long)band * a.dirs_stride : a.dirs;
    b.dirs_stride = a.dirs_stride;
    b.lastrow = a.lastrow; b.lastrow_col0 = 0;
    b.ckpt_out = (TM == 0 && a.ckpt) ? a.ckpt + (long long)band * a.ckpt_band_words : nullptr;
    int ms = psa_sweep_band16<C, TM, true, CH>(b, a.sc, tab4, sm, lane);
    ms = warp_max(ms);
    if (lane == 0) atomicMax(a.out_max, ms);
}

// Skewed direction layout of one long pair (TM 2): band b, step t, lane k -> one word (C <= 4: low
// 16 bits = virtual lane k, high 16 bits = virtual lane k+32, nibble c = column c) or two words
// (C > 4: word 0 = virtual lane k, word 1 = virtual lane k+32).  Row r of virtual lane v is in step r + v.
__host__ __device__ inline long long psa_skew_steps(long long m) { return (m + 63 + 31) & ~31LL; }   // whole chunks: every step of every chunk stores
__host__ __device__ inline long long psa_skew_band_bytes(long long m, int C) { return psa_skew_steps(m) * 32 * (C > 4 ? 8 : 4); }

// Checkpointed traceback: geometry of the checkpoints and of the tiles' direction words.
__host__ __device__ inline long long psa_tile_rows(long long m) { return (psa_skew_steps(m) + kTileSteps - 1) / kTileSteps; }
__host__ __device__ inline long long psa_ckpt_band_words(long long m, int C) { return psa_tile_rows(m) * (2 * C + 4) * 32; }
__host__ __device__ inline long long psa_tile_bytes(int C) { return (long long)kTileSteps * 32 * (C > 4 ? 8 : 4); }

struct PsaRetraceArgs {
    PsaScoring sc;
    const uint8_t* s0;
    const uint8_t* s1;
    int n, m, nbands;
    const unsigned long long* bnd;   // boundary columns of the score-only pass (all rows present, tagged `epoch`)
    long long bnd_stride;
    unsigned epoch;
    const uint32_t* ckpt;
    long long ckpt_band_words;
    const PsaTileRow* rows;          // [nrows] tile table (one row per 128-step range)
    int nrows;
    uint8_t* dirs;                   // slot s holds psa_tile_bytes(C) bytes: [step & 127][lane] words
    int* abort_flag;
    unsigned long long timeout_ns;
};

// One warp (= one CTA) per tile: blockIdx.y picks the step range (last range first: the walk starts there),
// blockIdx.x the band inside the range's corridor.
template <int C>
__global__ void __launch_bounds__(32) psa_retrace16_kernel(const __grid_constant__ PsaRetraceArgs a)
{
    __shared__ PsaWarpSmem16<C, false> sm;
    __shared__ uint32_t tab4[256];
    const int lane = threadIdx.x;
    const int k = a.nrows - 1 - (int)blockIdx.y;
    const PsaTileRow row = a.rows[k];
    const int band = row.bmin + (int)blockIdx.x;
    if (band > row.bmax) return;
    psa_build_tab4(tab4, a.sc, 4, lane, 32);
    __syncwarp();

    PsaBand b;
    b.s0 = a.s0; b.s1 = a.s1; b.n = a.n; b.m = a.m;
    b.band = band; b.nbands = a.nbands;
    b.col_off = 0;
    b.has_left = false; b.has_right = false;
    b.bnd_in = band == 0 ? nullptr : a.bnd + (long long)(band - 1) * a.bnd_stride;
    b.bnd_out = nullptr;
    b.epoch = a.epoch;
    b.abort_flag = a.abort_flag; b.timeout_ns = a.timeout_ns; b.timeline = nullptr;
    b.dirs = a.dirs + (row.slot0 + blockIdx.x) * psa_tile_bytes(C);
    b.dirs_stride = 0;
    b.dir_step0 = (long long)k * kTileSteps;
    b.lastrow = nullptr; b.lastrow_col0 = 0;
    b.ch_begin = k * kCkptChunks; b.ch_end = b.ch_begin + kCkptChunks;
    b.ckpt_in = k > 0 ? a.ckpt + (long long)band * a.ckpt_band_words + (long long)k * (kCkptWords<C> * 32) : nullptr;
    b.bnd_unscaled = true;
    psa_sweep_band16<C, 2, true, 32>(b, a.sc, tab4, sm, lane);
}

// ---------------------------------------------------------------------------------------
// a batch of independent pairs: persistent warps, one pair at a time per warp
// ---------------------------------------------------------------------------------------
template <int C, bool TRACE, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) psa_batch16_kernel(const __grid_constant__ PsaBatchArgs a)
{
    constexpr int S = TRACE ? 4 : 1;
    __shared__ PsaWarpSmem16<C, TRACE> sm[WARPS];
    __shared__ uint32_t tab4[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    psa_build_tab4(tab4, a.sc, S, threadIdx.x, 32 * WARPS);
    __syncthreads();
    const long long gw = (long long)blockIdx.x * WARPS + w;
    unsigned long long* scratch = a.scratch + gw * a.scratch_stride;
    int* lastrow = a.lastrow + gw * (64 * C);

    while (true) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(a.ticket, 1ull);
        k = __shfl_sync(kFullMask, k, 0);
        if (k >= (unsigned long long)a.npairs) break;
        const int n = a.len0[k], m = a.len1[k];
        const int nbands = (n + 64 * C - 1) / (64 * C);
        PsaBand b;
        b.s0 = a.seqs0 + a.off0[k]; b.s1 = a.seqs1 + a.off1[k];
        b.n = n; b.m = m; b.nbands = nbands;
        // in place: row r of the parked column is prefetched at chunk r/CH - 1 and rewritten at
        // step r+63 (see psa_fill.cuh: psa_batch_kernel)
        b.bnd_in = scratch; b.bnd_out = scratch;
        b.epoch = 1u;
        b.dirs = TRACE ? a.dirs + a.dirs_off[k] : nullptr;
        b.dirs_stride = (long long)nbands * (64 * C / 2);
        b.lastrow_col0 = (nbands - 1) * 64 * C;
        int ms = INT_MIN;
        if (m > 0) {
            for (int band = 0; band < nbands; band++) {
                b.band = band;
                b.lastrow = (band == nbands - 1) ? lastrow : nullptr;
                ms = max(ms, psa_sweep_band16<C, TRACE ? kBatchTM : 0, false, 32>(b, a.sc, tab4, sm[w], lane));
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

// Short reads (every s0 <= 32*C): persistent warps take TWO pairs per ticket and sweep them together.
template <int C, bool TRACE, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) psa_batch16_dual_kernel(const __grid_constant__ PsaBatchArgs a)
{
    constexpr int S = TRACE ? 4 : 1;
    __shared__ PsaWarpSmem16D<C, TRACE> sm[WARPS];
    __shared__ uint32_t tab4[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    psa_build_tab4(tab4, a.sc, S, threadIdx.x, 32 * WARPS);
    __syncthreads();
    while (true) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(a.ticket, 2ull);
        k = __shfl_sync(kFullMask, k, 0);
        if (k >= (unsigned long long)a.npairs) break;
        const bool hasB = k + 1 < (unsigned long long)a.npairs;
        const int nA = a.len0[k], mA = a.len1[k];
        const int nB = hasB ? a.len0[k + 1] : 0, mB = hasB ? a.len1[k + 1] : 0;
        int msA = INT_MIN, msB = INT_MIN, cA = 0, cB = 0;
        if (max(mA, mB) > 0)
            psa_sweep_dual16<C, TRACE>(a.seqs0 + a.off0[k], nA, a.seqs1 + a.off1[k], mA,
                                       a.seqs0 + (hasB ? a.off0[k + 1] : 0), nB, a.seqs1 + (hasB ? a.off1[k + 1] : 0), mB,
                                       TRACE ? a.dirs + a.dirs_off[k] : nullptr, (TRACE && hasB) ? a.dirs + a.dirs_off[k + 1] : nullptr,
                                       a.sc, tab4, sm[w], lane, msA, msB, cA, cB);
        if (lane == 0) {
            const bool realA = nA > 0 && mA > 0, realB = nB > 0 && mB > 0;
            a.maxsorce[k] = realA ? max(-100, msA / S) : -100;      // seed: psa.c:110,590
            a.corner[k] = realA ? cA / S : 0;
            if (hasB) {
                a.maxsorce[k + 1] = realB ? max(-100, msB / S) : -100;
                a.corner[k + 1] = realB ? cB / S : 0;
            }
        }
        __syncwarp();
    }
}

}  // namespace tsta
