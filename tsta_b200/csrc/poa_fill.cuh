// poa_fill.cuh -- sequence-to-graph DP (POA fill) and its traceback walker on sm_100a.
//
// Replaces the tile loop of control() (reference msa/c-t-simd.c:1019-1043) that feeds
// block_alignment()/block_line_alignment() (165-657) to the pthread pool, and the part of
// node_fuse() that walks the per-node trace rows (751-785).
//
// Rows are graph nodes in topological (p->sort[]) order, columns are read positions.  Every
// predecessor of a row has a smaller row index, so the plain skew "row + column strip" is a
// legal wavefront -- the same band/lane decomposition as the pairwise fill:
//   * the read is cut into bands of 32*C columns; one warp sweeps one band over all rows;
//   * lane k owns C columns and runs one row behind lane k-1; the strip's right edge (H, Fg)
//     moves to the next lane by shuffle and to the next band through tagged 64-bit elements
//     in L2 (no fences);
//   * unlike the pairwise case a row's inputs are ARBITRARY earlier rows (the node's pre[]),
//     so every lane keeps the (H, Eout) of its last D rows in a shared-memory history ring;
//     predecessors further back than the ring (rare: ~2 % of edges span more than 4 rows) are
//     read back from the global (H, Eout) matrix that every row also writes.
// Recurrences, tie-breaks and the two quirks (late heads Q1, 'N' columns) are the ones
// restated in oracle/poa_oracle.c; predecessor order decides ties (lowest index wins).
//
// Keys.  Every score the kernel carries is  score*256 + low byte, and the low byte is
//   [type:2][63 - predecessor index:6]   type 3 = diagonal, 2 = vertical (Eout), 1 = horizontal (Fg)
// so ONE 3-way maximum picks the cell's H, the move that produced it with the tie-break priority of
// c-t-simd.c:484-506 (diagonal > vertical > horizontal) and, inside a type, the LOWEST predecessor
// index that attains the maximum (the reference's tie-break).  The history holds H*256 and
// Eout*256 + 64; a successor adds 192 + (63-j) resp. 64 + (63-j) for its predecessor j.
//
// Trace: 2 bytes per cell, row-major, written with 128-bit stores from a shared ring:
//   byte 0  low byte of the winning key: type and 63 - j of the move (match / mismatch of a
//           diagonal is re-derived by the walker from the read and the node's base)
//   byte 1  bits 0-5: 63 - j*, j* = predecessor whose Eout is largest (the reference's esource index)
//           bit 6 = 0 iff  be + E >= H + O + E  at THIS cell: a vertical move arriving here keeps going
//                   up (the reference's rule 776-782 reads "not opened there, or opened and tied":
//                   both are this one comparison at the cell being entered)
//           bit 7 = 1 iff  Fg + E >= H + O + E  at THIS cell: a horizontal move arriving here from the
//                   right keeps going left (767-769, same argument)
#pragma once
#include "common.cuh"
#include "psa_fill.cuh"   // ld_relaxed_u64 / st_relaxed_u64_if, lds/sts helpers

namespace tsta {

constexpr int kPoaC = 4;            // columns per lane (widest variant; the kernel is a template: 2 or 4)
constexpr int kPoaHist = 32;        // history ring depth in rows (power of two)
constexpr int kPoaNear = kPoaHist - 2;   // predecessors at most this many rows back are read from the ring
constexpr int kPoaSlotNone = 32;    // extra history slots: "no predecessor" (keys of -infinity),
constexpr int kPoaSlotHead = 33;    // the virtual row above a head node,
constexpr int kPoaSlotLate = 34;    // and its late-head variant (quirk Q1)
constexpr int kPoaSlots = 35;
constexpr int kPoaFast = 8;         // predecessors the unrolled fast path handles
constexpr int kPoaChunk = 16;       // rows per staging chunk
constexpr int kPoaBndRing = 64;     // boundary ring (rows)
constexpr int kPoaTrRing = 64;      // trace ring (rows)
constexpr int kPoaDescRing = 64;
constexpr int kPoaKeyMin = INT_MIN / 2;
template <int C> constexpr int kPoaSlotBytesC = 33 * C * 8 + 16;   // 1072 / 544: lanes sit one slot apart, and (entry - slot) / 16 is odd,
                                                                     // so the 16-byte loads of a quarter warp fall on eight different bank groups
constexpr int kPoaSlotBytes = kPoaSlotBytesC<kPoaC>;

// Row descriptor, built on the host from the CSR graph and staged through shared memory a chunk
// ahead of the wavefront, so that the inner loop never waits for a dependent global load:
//   x: base | in-degree << 8 | flags << 16
//   y: history slot of predecessor 0..3 (one byte each; kPoaSlotNone when absent, the
//      virtual-row slots for heads)
//   z: history slot of predecessor 4..7 -- the fast path needs nothing else; rows flagged kPoaSlow
//      carry the first CSR edge of the row here instead (generic loop)
//   w: for rows flagged kPoaKeep: their row in the HBM spill
constexpr unsigned kPoaHead = 1u;        // in == 0
constexpr unsigned kPoaKeep = 2u;        // some successor is further away than the history ring: also store the keys in HBM
constexpr unsigned kPoaSlow = 4u;        // more than eight predecessors, or one further back than the ring: generic loop
constexpr unsigned kPoaLate = 8u;        // head with row index >= L (quirk Q1)

struct PoaFillArgs {
    const uint4* desc;        // [nrows] row descriptors
    const int* row_ptr;       // [nrows+1]
    const int* pred_row;      // [nedges], -1 = virtual predecessor of a head
    const uint8_t* base;      // [nrows]
    const int* hneg1;         // [nrows] H(r,-1); NULL: poa_rows_kernel computes it in band 0 (device-resident graphs)
    int* edge_spill;          // [nkeep] H(r,-1) keys of the spilled rows (written and read by band 0 when hneg1 is NULL)
    const uint8_t* read;      // [len]
    int len, nrows, nbands, L;
    int M, X, E, O;
    int2* HE;                 // [nkeep][stride] keys (H*256, Eout*256 + 64) of the rows with far successors (desc.w = their slot)
    long long stride;         // padded read length (multiple of 32*C)
    uint16_t* trace;          // [nrows][stride]
    unsigned long long* bnd;  // [nbands][nrows] tagged right-boundary elements
    unsigned epoch;           // 1..32767
    unsigned* ticket;
    int* lastsorce;           // [nrows] H(r, len-1)
};

template <int C>
struct PoaSmemC {
    // keys of the last kPoaHist rows (+ the three constant slots): entry [slot][lane + 1] holds the
    // lane's C columns as (H*256, Eout*256 + 64); entry [slot][0] is the band's left edge, so that
    // "the previous lane's last column" is one uniform load for every lane
    // (slot stride 1072 B = 67 x 16: lanes sit one row apart, i.e. lane k reads slot t-k, and an odd
    // multiple of 16 bytes spreads the eight 16-byte loads of a quarter warp over all banks)
    unsigned char hist[kPoaSlots * kPoaSlotBytesC<C>];
    uint16_t tr[kPoaTrRing][32 * C];         // trace ring
    int2 bring[kPoaBndRing];                 // left boundary keys (H(r, j0-1)*256, Fg(r, j0)*256 + 64) by row
    uint4 desc[kPoaDescRing];                // row descriptors
    int kconst[8];                           // scoring constants, read back once (see the kernel)
};
using PoaSmem = PoaSmemC<kPoaC>;

// boundary element: low word = key of H (H*256); high word: bits 0-15 Fg-H, bits 17-31 epoch.
// Fk - Hk = (Fg-H)*256 + 64, so its bytes 1-2 are the 16-bit difference; ep16 = epoch << 1.
__device__ __forceinline__ unsigned long long poa_bnd_pack(int Hk, int Fk, unsigned ep16)
{
    const unsigned hi = __byte_perm((unsigned)(Fk - Hk), ep16, 0x5421);
    return ((unsigned long long)hi << 32) | (unsigned)Hk;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_s4(uint32_t addr)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_s4(uint32_t addr, int x, int y, int z, int w)
{
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w));
}

// The (H, Eout) keys of one lane's C columns in a history slot: C == 4: two 16-byte loads, C == 2: one.
template <int C>
__device__ __forceinline__ void poa_lds_keys(uint32_t A, int (&h)[C], int (&e)[C])
{
    if constexpr (C == 4) {
        const int4 v01 = lds_s4(A), v23 = lds_s4(A + 16);
        h[0] = v01.x; e[0] = v01.y; h[1] = v01.z; e[1] = v01.w; h[2] = v23.x; e[2] = v23.y; h[3] = v23.z; e[3] = v23.w;
    } else {
        const int4 v01 = lds_s4(A);
        h[0] = v01.x; e[0] = v01.y; h[1] = v01.z; e[1] = v01.w;
    }
}
template <int C>
__device__ __forceinline__ void poa_sts_keys(uint32_t A, const int (&ko)[2 * C])
{
    sts_s4(A, ko[0], ko[1], ko[2], ko[3]);
    if constexpr (C == 4) sts_s4(A + 16, ko[4], ko[5], ko[6], ko[7]);
}


// One row of one lane's strip (C columns, serial in F).  In: the reduced predecessor keys kd/ke, the
// Fg key of the strip's first column and (quirks only) the H key of the left neighbour; out: the
// strip's history keys (ko), trace words (tw) and the right edge (Fk, Hleft).
template <int C, bool QUIRKS>
__device__ __forceinline__ void poa_row(const int (&kd)[C], const int (&ke)[C], const uint8_t (&rc)[C], const bool (&isL)[C],
                                        const bool (&isLm1)[C], const int b, const int MS, const int XS, const int ES,
                                        const int ES64, const int ES128, const int OES, const int Fadj, int& Fk, int& Hleft,
                                        int (&ko)[2 * C], uint32_t (&tw)[(C + 1) / 2])
{
    uint32_t lo8[C], hi8[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (QUIRKS && isL[c]) Fk += Fadj;                             // quirk Q1: carried F re-based at column L
        const int hk = kd[c] + ((int)rc[c] == b ? MS : XS);
        int Hk = __vimax3_s32(hk, ke[c], Fk);
        if (QUIRKS && rc[c] == 'N') {                                 // quirk Q3: 'N' columns copy H from the left
            const int H = Hleft >> 8;
            int lb = 64;
            if ((ke[c] >> 8) == H) lb = ke[c] & 255;
            if ((hk >> 8) == H) lb = hk & 255;
            Hk = H * 256 + lb;
        }
        const int H256 = Hk & ~255;
        const int t = H256 + OES;                                     // (H+O+E)*256 + 64: Fg / Eout opened from here
        const int kem = ke[c] & ~255;
        const int WE = __viaddmax_s32(kem, ES128, t);              // low byte 128 iff be+E >= H+O+E, else 64
        int WF = __viaddmax_s32(Fk, ES64, t);                      // low byte 128 iff Fg+E >= H+O+E, else 64
        ko[2 * c] = H256;
        ko[2 * c + 1] = __viaddmax_s32(kem, ES64, t);              // Eout*256 + 64
        const int Fn = __viaddmax_s32(Fk, ES, t);                     // Fg(c+1)*256 + 64
        if (QUIRKS && isLm1[c] && Fadj != 0) {
            // the next column re-bases F: its |fsource| == 2 test sees the re-based value (c-t-simd.c:228-233)
            const bool fabs2 = (Fn + Fadj == t), fneg = (Fk + ES == t);
            WF = (!fabs2 || fneg) ? 128 : 64;
        }
        lo8[c] = (uint32_t)Hk;
        hi8[c] = ((uint32_t)ke[c] & ~0xc0u) | ((uint32_t)WF & 0x80u) | ((uint32_t)WE & 0x40u);
        Fk = Fn;
        Hleft = H256;
    }
#pragma unroll
    for (int c = 0; c + 1 < C; c += 2)
        tw[c / 2] = __byte_perm(__byte_perm(lo8[c], lo8[c + 1], 0x0040), __byte_perm(hi8[c], hi8[c + 1], 0x0040), 0x5140);
    if constexpr (C & 1) tw[C / 2] = __byte_perm(lo8[C - 1], hi8[C - 1], 0x0040);
}

template <int C>
__global__ void __launch_bounds__(32) poa_fill_kernel(const __grid_constant__ PoaFillArgs a)
{
    static_assert(C == 2 || C == 4, "strip widths of the POA fill");
    constexpr int kSlotBytes = kPoaSlotBytesC<C>;
    constexpr int kEnt = C * 8;                                 // bytes of one lane's entry in a history slot
    extern __shared__ __align__(16) unsigned char poa_smem_raw[];
    PoaSmemC<C>& sm = *reinterpret_cast<PoaSmemC<C>*>(poa_smem_raw);
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);

    const int E = a.E, O = a.O, OE = a.O + a.E, M = a.M, X = a.X;
    const int R = a.nrows, len = a.len, L = a.L;
    // Scoring constants go through shared memory once: values the compiler can trace back to the
    // kernel parameters are re-derived inside the step (an LDC from the parameter bank plus a shift,
    // with their latency on the row's dependency chain) instead of being kept in registers.
    if (lane == 0) {
        sm.kconst[0] = 256 * M; sm.kconst[1] = 256 * X; sm.kconst[2] = 256 * E; sm.kconst[3] = 256 * OE + 64;
        sm.kconst[4] = 256 * E + 64; sm.kconst[5] = 256 * E + 128;
    }
    __syncwarp();
    const volatile int* kc = sm.kconst;
    const int MS = kc[0], XS = kc[1], ES = kc[2], OES = kc[3], ES64 = kc[4], ES128 = kc[5];
    const int j0 = (band * 32 + lane) * C;
    uint8_t rc[C];                       // read bytes of this lane's columns (0 = pad: matches nothing)
    bool isL[C], isLm1[C];               // column == L / L-1 (the late-head quirk re-bases F at column L)
    bool hasN = false;                   // quirk Q3 applies to this lane's strip at all?
#pragma unroll
    for (int c = 0; c < C; c++) {
        rc[c] = (j0 + c < len) ? a.read[j0 + c] : 0;
        isL[c] = (j0 + c == L);
        isLm1[c] = (j0 + c + 1 == L);
        hasN |= rc[c] == 'N';
    }
    const bool owns_last = (len - 1 >= j0) && (len - 1 < j0 + C);
    const int clast = owns_last ? len - 1 - j0 : 0;
    bool hasL = false;
#pragma unroll
    for (int c = 0; c < C; c++) hasL |= isL[c] || isLm1[c];
    const unsigned long long* bnd_in = a.bnd + (long long)(band - 1) * R;
    unsigned long long* bnd_out = a.bnd + (long long)band * R;
    const bool store_right = (band + 1 < a.nbands) && lane == 31;
    const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(&sm.hist[0]);
    const uint32_t my_hist = hist_addr + lane * kEnt;           // this lane's view: +kEnt-8 previous lane's last H key, +kEnt.. own columns

    // ---- constant history slots -------------------------------------------------------------------
    {
        // "no predecessor": keys of -infinity everywhere (also the left edge)
        {
            int kmin[2 * C];
#pragma unroll
            for (int q = 0; q < 2 * C; q++) kmin[q] = kPoaKeyMin;
            poa_sts_keys<C>(hist_addr + kPoaSlotNone * kSlotBytes + (lane + 1) * kEnt, kmin);
            if (lane == 0) poa_sts_keys<C>(hist_addr + kPoaSlotNone * kSlotBytes, kmin);
        }
        // virtual rows above a head (c-t-simd.c:176-188): the H key of column c is only ever read
        // as the DIAGONAL input of column c+1, the E key is Eout_v(c) = Hv(c) + O + E (opened)
#pragma unroll
        for (int v = 0; v < 2; v++) {
            const bool late = v == 1;
            int k[2 * C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int col = j0 + c;
                const int Hd_next = (late && col + 1 <= L - 1) ? (col + 1) * E : O + (col + 1) * E;   // diagonal input of column col+1
                const int Hp = (late && col <= L - 1) ? (col + 1) * E : O + (col + 1) * E;            // Hv(col)
                k[2 * c] = Hd_next * 256;
                k[2 * c + 1] = (Hp + OE) * 256 + 64;
            }
            const uint32_t slot = hist_addr + (kPoaSlotHead + v) * kSlotBytes;
            poa_sts_keys<C>(slot + (lane + 1) * kEnt, k);
            if (lane == 0) {   // left edge: diagonal input of column j0
                const int Hd0 = (late && j0 <= L - 1) ? j0 * E : (j0 == 0 ? 0 : O + j0 * E);
                sts_s4(slot + kEnt - 16, 0, 0, Hd0 * 256, 0);
            }
        }
    }

    int Hout = 0, Fout = 0;              // right edge of the previous step: keys of H and of Fg(next column)
    // explicit shared addresses: the compiler otherwise re-derives the shared window inside the step loop
    const uint32_t desc_addr = (uint32_t)__cvta_generic_to_shared(&sm.desc[0]);
    const uint32_t bring_addr = (uint32_t)__cvta_generic_to_shared(&sm.bring[0]);
    const uint32_t tr_addr = (uint32_t)__cvta_generic_to_shared(&sm.tr[0][0]) + lane * (C * 2);
    const unsigned ep16 = a.epoch << 1;
    // "no predecessor" is the same -infinity for every lane: all lanes read ONE address (a broadcast,
    // one wavefront) instead of their own copy 32 bytes apart (two-way bank conflicts on every load)
    const uint32_t none_addr = hist_addr + kPoaSlotNone * kSlotBytes + kEnt;
    // flags (descriptor bits 16..) that take a lane out of the PLAIN step: slow / spilled rows always,
    // late heads only in the lanes that own column L-1 or L; every row when the strip has an 'N'
    const unsigned unplain = hasN ? 0xffffffffu : (((kPoaSlow | kPoaKeep) | (hasL ? kPoaLate : 0u)) << 16);
    const bool last_band = band + 1 == a.nbands;
    const int nchunks = (R + 31 + kPoaChunk - 1) / kPoaChunk;
    unsigned long long pre_b = 0;
    if (band > 0 && lane < kPoaChunk && lane < R) pre_b = ld_relaxed_u64(bnd_in + lane);
    int pre_h = 0;                       // band 0: H(r,-1) of the next chunk's rows, fetched a chunk ahead
    if (band == 0 && lane < kPoaChunk && lane < R) pre_h = __ldg(a.hneg1 + lane);
    // descriptors: lanes 16..31 stage rows [t0+16, t0+32) at the top of chunk t0 (one chunk ahead)
    if (lane < kPoaChunk && lane < R) sm.desc[lane] = __ldg(a.desc + lane);
    uint4 pre_d = make_uint4(0, 0, 0, 0);
    if (lane >= 16 && lane - 16 + kPoaChunk < R) pre_d = __ldg(a.desc + (lane - 16 + kPoaChunk));

    for (int ch = 0; ch < nchunks; ch++) {
        const int t0 = ch * kPoaChunk;
        // ---- stage next chunk's descriptors, this chunk's left boundary --------------------------
        if (lane >= 16) {
            const int rown = t0 + kPoaChunk + (lane - 16);
            if (rown < R) sm.desc[rown & (kPoaDescRing - 1)] = pre_d;
            const int rownn = rown + kPoaChunk;
            if (rownn < R) pre_d = __ldg(a.desc + rownn);
        }
        {
            const int row = lane < kPoaChunk ? t0 + lane : R;
            if (band > 0) {
                while (!__all_sync(kFullMask, row >= R || (unsigned)(pre_b >> 49) == a.epoch)) {
                    __nanosleep(20);
                    if (row < R && (unsigned)(pre_b >> 49) != a.epoch) pre_b = ld_relaxed_u64(bnd_in + row);
                }
                if (row < R) {
                    const int Hk = (int)(unsigned)pre_b;
                    const unsigned hi = (unsigned)(pre_b >> 32);
                    sm.bring[row & (kPoaBndRing - 1)] = make_int2(Hk, Hk + (int)(short)(hi & 0xffffu) * 256 + 64);
                }
                const int rowp = row + kPoaChunk;
                if (rowp < R) pre_b = ld_relaxed_u64(bnd_in + rowp);
            } else {
                if (row < R) sm.bring[row & (kPoaBndRing - 1)] = make_int2(pre_h * 256, (pre_h + OE) * 256 + 64);   // Fg(r,0) = H(r,-1)+O+E
                const int rowp = row + kPoaChunk;
                if (rowp < R) pre_h = __ldg(a.hneg1 + rowp);
            }
        }
        __syncwarp();

        unsigned long long* bout = bnd_out + (t0 - lane);          // boundary element of this lane's row at s = 0
        // the descriptor of a lane's row is fetched one step ahead, off the step's dependency chain
        uint4 dsc_next = lds_u4(desc_addr + ((uint32_t)(t0 - lane) & 63u) * 16);
#pragma unroll 1
        for (int s = 0; s < kPoaChunk; s++, bout++) {
            const int r = t0 + s - lane;
            int Hleft = __shfl_up_sync(kFullMask, Hout, 1);
            int F = __shfl_up_sync(kFullMask, Fout, 1);
            __syncwarp();   // orders last step's history stores before this step's loads
            const bool active = (unsigned)r < (unsigned)R;
            const uint32_t rm = (uint32_t)r & 63u;
            const uint4 dsc = dsc_next;
            dsc_next = lds_u4(desc_addr + ((rm + 1u) & 63u) * 16);
            // how many predecessors does any lane of the warp have this step?  (78 % of the rows of an
            // MSA graph have one, 19 % two, 3 % three: the fourth round of the reduction is skipped
            // warp-wide whenever no lane needs it)
            const int npre_l = active ? (int)((dsc.x >> 8) & 0xffu) : 0;
            const int maxnp = __reduce_max_sync(kFullMask, npre_l);   // warp-uniform: rounds beyond it are skipped
            // PLAIN step: every lane active, at most eight near predecessors, no spill, no quirk -- the
            // whole step is one straight-line block (selects and predicated stores instead of branches)
            const bool plain = active && !(dsc.x & unplain) && !hasN;
            if (__all_sync(kFullMask, plain)) {
                const uint2 bv = lds_u2(bring_addr + rm * 8);
                const int Hleft_in = lane == 0 ? (int)bv.x : Hleft;
                int Fk = lane == 0 ? (int)bv.y : F;
                const int b = (int)(dsc.x & 0xffu);
                int kd[C], ke[C];
                {
                    const uint32_t A = my_hist + (dsc.y & 0xffu) * kSlotBytes;
                    const int Hd0 = lds_s32(A + kEnt - 8);
                    int ph[C], pe[C];
                    poa_lds_keys<C>(A + kEnt, ph, pe);
                    kd[0] = Hd0 + 255;
#pragma unroll
                    for (int c = 1; c < C; c++) kd[c] = ph[c - 1] + 255;
#pragma unroll
                    for (int c = 0; c < C; c++) ke[c] = pe[c] + 127;
                }
#pragma unroll
                for (int j = 1; j < kPoaFast; j++) {
                    if (j >= 3 && j >= maxnp) break;
                    const uint32_t slot = ((j < 4 ? dsc.y : dsc.z) >> (8 * (j & 3))) & 0xffu;
                    const uint32_t A = slot == kPoaSlotNone ? none_addr : my_hist + slot * kSlotBytes;
                    const int Hd0 = lds_s32(A + kEnt - 8);
                    int ph[C], pe[C];
                    poa_lds_keys<C>(A + kEnt, ph, pe);
                    const int wd = 192 + 63 - j, we = 64 + 63 - j;
                    kd[0] = __viaddmax_s32(Hd0, wd, kd[0]);
#pragma unroll
                    for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(ph[c - 1], wd, kd[c]);
#pragma unroll
                    for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(pe[c], we, ke[c]);
                }
                int ko[2 * C];
                uint32_t tw[C / 2];
                int Hl = Hleft_in;
                poa_row<C, false>(kd, ke, rc, isL, isLm1, b, MS, XS, ES, ES64, ES128, OES, 0, Fk, Hl, ko, tw);
                Hout = Hl; Fout = Fk;
                const uint32_t own = my_hist + ((uint32_t)r & (kPoaHist - 1)) * kSlotBytes;
                poa_sts_keys<C>(own + kEnt, ko);
                if (lane == 0) sts_u32(own + kEnt - 8, (uint32_t)Hleft_in);
                if constexpr (C == 4) sts_u2(tr_addr + rm * (32 * C * 2), tw[0], tw[1]);
                else sts_u32(tr_addr + rm * (32 * C * 2), tw[0]);
                if (owns_last) {
                    int hk = ko[0];
#pragma unroll
                    for (int c = 1; c < C; c++) hk = clast == c ? ko[2 * c] : hk;
                    a.lastsorce[r] = hk >> 8;
                }
                st_relaxed_u64_if(bout, poa_bnd_pack(Hout, Fout, ep16), store_right);
            } else if (active) {
                if (lane == 0) {
                    const uint2 v = lds_u2(bring_addr + rm * 8);
                    Hleft = (int)v.x;
                    F = (int)v.y;
                }
                const int b = (int)(dsc.x & 0xffu);
                const unsigned flags = dsc.x >> 16;

                int kd[C], ke[C];
                if (!(flags & kPoaSlow)) {
                    // ---- fast path: up to eight predecessors, all inside the history ring ----------
                    {   // predecessor 0: every row has one (heads: the virtual row's slot)
                        const uint32_t A = my_hist + (dsc.y & 0xffu) * kSlotBytes;
                        const int Hd0 = lds_s32(A + kEnt - 8);
                        int ph[C], pe[C];
                        poa_lds_keys<C>(A + kEnt, ph, pe);
                        kd[0] = Hd0 + 255;
#pragma unroll
                        for (int c = 1; c < C; c++) kd[c] = ph[c - 1] + 255;
#pragma unroll
                        for (int c = 0; c < C; c++) ke[c] = pe[c] + 127;
                    }
#pragma unroll
                    for (int j = 1; j < kPoaFast; j++) {
                        if (j >= 2 && j >= maxnp) break;
                        const uint32_t slot = ((j < 4 ? dsc.y : dsc.z) >> (8 * (j & 3))) & 0xffu;
                        const uint32_t A = slot == kPoaSlotNone ? none_addr : my_hist + slot * kSlotBytes;
                        const int Hd0 = lds_s32(A + kEnt - 8);
                        int ph[C], pe[C];
                        poa_lds_keys<C>(A + kEnt, ph, pe);
                        const int wd = 192 + 63 - j, we = 64 + 63 - j;
                        kd[0] = __viaddmax_s32(Hd0, wd, kd[0]);
#pragma unroll
                        for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(ph[c - 1], wd, kd[c]);
#pragma unroll
                        for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(pe[c], we, ke[c]);
                    }
                } else {
                    // ---- slow path: any in-degree, predecessors further back than the ring ----------
#pragma unroll
                    for (int c = 0; c < C; c++) { kd[c] = kPoaKeyMin; ke[c] = kPoaKeyMin; }
                    const int npre = (int)((dsc.x >> 8) & 0xffu);
                    const int e0 = (int)dsc.z;
                    for (int j = 0; j < npre; j++) {
                        const int p = __ldg(a.pred_row + e0 + j);
                        int ph[C], pe[C];
                        int Hd0;
                        if (p < 0 || r - p <= kPoaNear) {
                            const int slot = p < 0 ? ((flags & kPoaLate) ? kPoaSlotLate : kPoaSlotHead) : (p & (kPoaHist - 1));
                            const uint32_t A = my_hist + slot * kSlotBytes;
                            Hd0 = lds_s32(A + kEnt - 8);
                            poa_lds_keys<C>(A + kEnt, ph, pe);
                        } else {
                            const int2* src = a.HE + (long long)__ldg(&a.desc[p].w) * a.stride + j0;
                            const int4 v01 = __ldcg(reinterpret_cast<const int4*>(src));
                            ph[0] = v01.x; pe[0] = v01.y; ph[1] = v01.z; pe[1] = v01.w;
                            if constexpr (C == 4) {
                                const int4 v23 = __ldcg(reinterpret_cast<const int4*>(src + 2));
                                ph[2] = v23.x; pe[2] = v23.y; ph[3] = v23.z; pe[3] = v23.w;
                            }
                            if (lane > 0) Hd0 = __ldcg(src - 1).x;
                            else if (band > 0) Hd0 = (int)(unsigned)ld_relaxed_u64(bnd_in + p);
                            else Hd0 = __ldg(a.hneg1 + p) * 256;
                        }
                        const int wd = 192 + 63 - j, we = 64 + 63 - j;
                        kd[0] = __viaddmax_s32(Hd0, wd, kd[0]);
#pragma unroll
                        for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(ph[c - 1], wd, kd[c]);
#pragma unroll
                        for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(pe[c], we, ke[c]);
                    }
                }

                // ---- the row itself: serial along the strip ---------------------------------------------
                int Fk = F;
                const int Hleft_in = Hleft;
                int ko[2 * C];
                uint32_t tw[C / 2];
                // the two quirks (late heads re-base F at column L, 'N' columns copy H) touch a handful of
                // cells of a whole job: they get their own copy of the row code, the common one is free of them
                if (hasN || (hasL && (flags & kPoaLate)))
                    poa_row<C, true>(kd, ke, rc, isL, isLm1, b, MS, XS, ES, ES64, ES128, OES, (flags & kPoaLate) ? 256 * O : 0, Fk, Hleft, ko, tw);
                else
                    poa_row<C, false>(kd, ke, rc, isL, isLm1, b, MS, XS, ES, ES64, ES128, OES, 0, Fk, Hleft, ko, tw);
                Hout = Hleft; Fout = Fk;
                // history ring (own columns + the band's left edge), HBM spill for far successors, trace ring
                {
                    const uint32_t own = my_hist + ((uint32_t)r & (kPoaHist - 1)) * kSlotBytes;
                    poa_sts_keys<C>(own + kEnt, ko);
                    if (lane == 0) sts_u32(own + kEnt - 8, (uint32_t)Hleft_in);
                    if (flags & kPoaKeep) {    // a successor further away than the history ring will read it back
                        int2* g = a.HE + (long long)dsc.w * a.stride + j0;
                        *reinterpret_cast<int4*>(g) = make_int4(ko[0], ko[1], ko[2], ko[3]);
                        if constexpr (C == 4) *reinterpret_cast<int4*>(g + 2) = make_int4(ko[4], ko[5], ko[6], ko[7]);
                    }
                    if constexpr (C == 4) sts_u2(tr_addr + rm * (32 * C * 2), tw[0], tw[1]);
                    else sts_u32(tr_addr + rm * (32 * C * 2), tw[0]);
                }
                if (last_band) {
                    if (owns_last) {
                        int hk = ko[0];
#pragma unroll
                        for (int c = 1; c < C; c++) hk = clast == c ? ko[2 * c] : hk;
                        a.lastsorce[r] = hk >> 8;
                    }
                }
                st_relaxed_u64_if(bout, poa_bnd_pack(Hout, Fout, ep16), store_right);
            }
        }

        // ---- flush finished trace rows [lo, hi): 256 B per band row, 16 B per lane-unit ---------
        __syncwarp();
        {
            const int hi = min(R, max(0, t0 + kPoaChunk - 31));
            const int lo = min(R, max(0, t0 - 31));
            constexpr int kUnits = 32 * C * 2 / 16;    // 16-byte units per row: 16 (C == 4) or 8 (C == 2)
            constexpr int kRowsPerPass = 32 / kUnits;  // rows the warp covers with one load per lane
            const int units = (hi - lo) * kUnits;
            if (hi - lo == kPoaChunk) {
                // a whole chunk: lane u handles part u % kUnits of rows lo + u / kUnits + kRowsPerPass * k -- independent load/store pairs
                const int part = lane % kUnits;
                uint16_t* dst = a.trace + (long long)(lo + lane / kUnits) * a.stride + (long long)band * 32 * C + part * 8;
                uint4 v[kPoaChunk / kRowsPerPass];
#pragma unroll
                for (int k = 0; k < kPoaChunk / kRowsPerPass; k++)
                    v[k] = *reinterpret_cast<const uint4*>(&sm.tr[(lo + lane / kUnits + kRowsPerPass * k) & (kPoaTrRing - 1)][part * 8]);
#pragma unroll
                for (int k = 0; k < kPoaChunk / kRowsPerPass; k++)
                    __stcs(reinterpret_cast<uint4*>(dst + (long long)(kRowsPerPass * k) * a.stride), v[k]);
            } else
            for (int u = lane; u < units; u += 32) {
                const int row = lo + u / kUnits, part = u % kUnits;
                const uint4 v = *reinterpret_cast<const uint4*>(&sm.tr[row & (kPoaTrRing - 1)][part * 8]);
                __stcs(reinterpret_cast<uint4*>(a.trace + (long long)row * a.stride + (long long)band * 32 * C + part * 8), v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// traceback walker (node_fuse control flow, c-t-simd.c:751-956, with the forced
// continuations folded into the automaton).  One CTA; a window of the trace matrix and the
// window's slice of the CSR graph are staged in shared memory, one thread walks.
// ---------------------------------------------------------------------------------------
constexpr int kPoaWinRows = 128;
constexpr int kPoaWinCols = 128;
constexpr int kPoaWinEdges = 2048;
constexpr int kPoaWalkThreads = 256;

struct PoaStep {
    int row, col;
    uint8_t kind, pred;
    uint16_t pad;
};

struct PoaWalkArgs {
    const uint16_t* trace;
    long long stride;
    const int* row_ptr;
    const int* pred_row;
    const uint4* desc;        // [nrows]  row descriptors: a diagonal move is a match iff read[col] == base of the row (x & 255)
    const uint8_t* read;      // [len]
    int len, nrows, start_row;
    int fixed8;               // the predecessor table is at stride 8 (device-built graphs): row_ptr[r] = 8 r
    PoaStep* path;
    int* nsteps;
};

__global__ void __launch_bounds__(kPoaWalkThreads) poa_walk_kernel(const PoaWalkArgs a)
{
    __shared__ __align__(16) uint16_t win[kPoaWinRows][kPoaWinCols];
    __shared__ int s_ptr[kPoaWinRows + 1];
    __shared__ int s_pred[kPoaWinEdges];
    __shared__ uint8_t s_base[kPoaWinRows], s_read[kPoaWinCols];
    __shared__ uint8_t s_chain[kPoaWinRows][32];   // s_chain[w][k]: window row reached from window row w by k first-predecessor hops (255: outside)
    __shared__ int s_row, s_col, s_pend, s_n;
    if (threadIdx.x == 0) { s_row = a.start_row; s_col = a.len - 1; s_pend = 0; s_n = 0; }
    __syncthreads();
    while (true) {
        const int r0 = s_row, c0 = s_col;
        if (r0 < 0 || c0 < 0) break;
        // window: rows (r0-128, r0], columns [w0, w0+128) with c0 near the right end
        int w0 = ((c0 + 8) & ~7) - kPoaWinCols;
        if (w0 < 0) w0 = 0;
        const int rlo = max(0, r0 - kPoaWinRows + 1);
        // the window's cells by cp.async (no round trip through registers: all 8 copies of a thread are in flight at once,
        // next to the graph slice below)
        const uint32_t win_addr = (uint32_t)__cvta_generic_to_shared(&win[0][0]);
        for (int u = threadIdx.x; u < kPoaWinRows * (kPoaWinCols / 8); u += kPoaWalkThreads) {
            const int rr = u / (kPoaWinCols / 8), q = u % (kPoaWinCols / 8);
            const int row = r0 - rr;
            const long long col = (long long)w0 + q * 8;
            if (row >= 0 && col < a.stride)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(win_addr + (uint32_t)u * 16u),
                             "l"(a.trace + (long long)row * a.stride + col) : "memory");
        }
        const int ebase = a.fixed8 ? 8 * rlo : __ldg(a.row_ptr + rlo);
        const int ecount = min((a.fixed8 ? 8 * (r0 + 1) : __ldg(a.row_ptr + r0 + 1)) - ebase, kPoaWinEdges);
        for (int u = threadIdx.x; u <= r0 - rlo + 1; u += kPoaWalkThreads) s_ptr[u] = a.fixed8 ? 8 * (rlo + u) : a.row_ptr[rlo + u];
        for (int u = threadIdx.x; u < kPoaWinRows; u += kPoaWalkThreads) s_base[u] = r0 - u >= 0 ? (uint8_t)(__ldg(&a.desc[r0 - u].x) & 0xffu) : 0;
        for (int u = threadIdx.x; u < kPoaWinCols; u += kPoaWalkThreads) s_read[u] = w0 + u < a.len ? a.read[w0 + u] : 0;
        for (int u = threadIdx.x; u < ecount; u += kPoaWalkThreads) s_pred[u] = a.pred_row[ebase + u];
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        // First-predecessor chains of the window's rows, by doubling (5 rounds): a read that mostly matches the graph
        // walks diagonally along them even where consecutive path nodes are not consecutive rows.
        for (int w = threadIdx.x; w < kPoaWinRows; w += kPoaWalkThreads) {
            const int row = r0 - w;
            int hop = 255;
            if (row >= rlo) {
                const int e = s_ptr[row - rlo] - ebase;
                const int p0 = e < kPoaWinEdges ? s_pred[e] : -1;
                if (p0 >= rlo) hop = r0 - p0;
            }
            s_chain[w][0] = (uint8_t)w;
            s_chain[w][1] = (uint8_t)hop;
        }
        __syncthreads();
#pragma unroll 1
        for (int half = 1; half < 32; half *= 2) {                  // entries (half, 2 half] from entries (0, half]
            for (int u = threadIdx.x; u < kPoaWinRows * half; u += kPoaWalkThreads) {
                const int w = u / half, k = half + 1 + u % half;
                if (k < 32) {
                    const int mid = s_chain[w][half];
                    s_chain[w][k] = mid == 255 ? (uint8_t)255 : s_chain[mid][k - half];
                }
            }
            __syncthreads();
        }
        if (threadIdx.x < 32) {
            // pend: how this cell was entered (1 = horizontal move, 2 = vertical move, 0 = otherwise);
            // the cell's own bits 14 / 15 then say whether the move is forced to continue.
            // One warp walks: lane k ASSUMES that the k cells before it were diagonal moves, each into its row's FIRST
            // predecessor, and decodes the cell it would then stand on; the lanes up to the first cell that says
            // otherwise are on the path, and the walk resumes from that cell's real successor.  10-30 steps per pass
            // instead of one, for the same chain of dependent shared-memory loads.
            const int lane = threadIdx.x;
            int r = r0, c = c0, pend = s_pend, n = s_n;
            while (r >= rlo && c >= w0) {
                const int wk = s_chain[r0 - r][lane];               // window row of this lane's cell (255: left the window)
                const int rr = r0 - wk, cc = c - lane;
                const bool in = wk != 255 && cc >= w0;
                uint32_t type = 0;
                int pj = 0, p = 0, kind = 0, hop = 255;
                if (in) {
                    const uint32_t w = win[wk][cc - w0];
                    const int e0 = s_ptr[rr - rlo] - ebase;
                    const bool match = s_read[cc - w0] == s_base[wk];
                    hop = s_chain[wk][1];
                    const int pd = lane == 0 ? pend : 0;                  // lanes > 0 are entered by a diagonal move
                    const bool fF = pd == 1 && cc > 0 && (w & 0x8000u);
                    const bool fE = pd == 2 && rr > 0 && !(w & 0x4000u);
                    type = fF ? 1u : (fE ? 2u : ((w >> 6) & 3u));        // 3 diagonal, 2 vertical, 1 horizontal
                    pj = 63 - (int)((fE ? (w >> 8) : w) & 63u);
                    kind = type == 3u ? (match ? 1 : 2) : (type == 2u ? 0 : 3);
                    const int e = e0 + pj;
                    p = s_pred[min(e, kPoaWinEdges - 1)];
                    if (type != 1u && e >= kPoaWinEdges) p = a.pred_row[ebase + e];   // (a window with > 2048 edges: never seen)
                }
                // the next cell IS the one lane + 1 looks at: a diagonal move into the row the chain continues with
                const bool cont = in && type == 3u && hop != 255 && p == r0 - hop;
                const unsigned okmask = __ballot_sync(kFullMask, cont);
                const int d = okmask == 0xffffffffu ? 32 : __ffs((int)~okmask) - 1;   // first lane whose successor is elsewhere
                if (lane <= d && in) {
                    PoaStep st;
                    st.row = rr; st.col = cc; st.kind = (uint8_t)kind; st.pred = (uint8_t)(type == 1u ? 0 : pj); st.pad = 0;
                    a.path[n + lane] = st;
                }
                // where the walk goes on: lane d's own successor if lane d stands on a cell of the window, else (it left
                // the window) the successor of lane d - 1, which is where lane d would stand
                const int src = d == 32 ? 31 : d;
                const int d_in = __shfl_sync(kFullMask, (int)in, src);
                const int from = (d == 32 || d_in) ? src : d - 1;     // (d == 0 is always inside: the loop condition)
                const int nr = __shfl_sync(kFullMask, type == 1u ? rr : p, from);
                const int nc = __shfl_sync(kFullMask, type == 2u ? cc : cc - 1, from);
                const int np = __shfl_sync(kFullMask, type == 1u ? 1 : (type == 2u ? 2 : 0), from);
                r = nr; c = nc; pend = np;
                n += (d == 32) ? 32 : (d_in ? d + 1 : d);
            }
            if (lane == 0) { s_row = r; s_col = c; s_pend = pend; s_n = n; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.nsteps = s_n;
}

}  // namespace tsta
