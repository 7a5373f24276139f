// poa_rows.cuh -- row-synchronous sequence-to-graph DP (POA fill) on sm_100a: the default fill kernel.
//
// Same job, inputs, outputs and trace format as poa_fill_kernel (poa_fill.cuh; reference
// msa/c-t-simd.c:165-657, block_alignment / block_line_alignment), different decomposition:
//
//   * a band is 32*C columns and one warp; ALL 32 lanes work on the SAME row (no skew).  The row's
//     descriptor is therefore warp-uniform: the in-degree loop, the slow / spill / late-head cases are
//     uniform branches, and the 78 % of rows whose predecessor is the row just finished take their
//     inputs from registers (forwarding) -- no shared-memory round trip on the row-to-row chain;
//   * the horizontal (F) chain of a row, which is what forces the skew in poa_fill_kernel, is solved by a
//     max-plus scan: with O <= 0 (tsta::check_params) opening a gap from a cell whose H came from F never
//     beats extending that F, so
//         Fg(j) = max_{k<j} ( H^(k) + O + E*(j-k) ),   H^ = max(diagonal, vertical) only,
//     a prefix maximum with a linear decay.  Each lane forms the local prefixes of its C columns, a
//     5-round Kogge-Stone scan over shuffles carries them across the warp, the band's left edge joins as
//     Fin + E*C*lane, and every column's F is one more VIADDMNMX.  All F values are exact, so H, the
//     winning move and the continuation flags -- functions of the values -- are bit-identical;
//   * bands hand their right edge to the next band per ROW (tagged 64-bit elements as before), imported in
//     groups of 8 rows: band b runs ~10 rows behind band b-1 instead of 45 steps;
//   * a lane's trace cells of a row are 2*C contiguous bytes and the warp's 64*C: written straight to HBM
//     (one 8-byte store per lane, full sectors), no staging ring.
//
// Rows that need the serial F semantics -- late heads (quirk Q1: the carried F is re-based at column L) in
// the band that owns column L-1 / L -- run the row code of poa_fill.cuh 32 times with the carry moved one
// lane per pass (a handful of rows per job).  Reads with 'N' columns (quirk Q3: H copied from the left
// neighbour, a second horizontal chain) and strip widths that do not divide L stay on poa_fill_kernel.
//
// Per-row dependency chain: forward shuffle of H(r-1, j0-1) -> H^ -> local prefix -> 5 x (SHFL + VIADDMNMX)
// -> F -> H: ~230 cycles for ~140 instructions, against 549 cycles per row-step of the skewed kernel.
#pragma once
#include "poa_fill.cuh"

namespace tsta {

constexpr int kPoaGroup = 8;                 // rows per boundary import group
constexpr int kPoaRowsDescRing = 128;
constexpr int kPoaPlane = 33 * 16;           // one 16-byte unit per lane (+ the band's left edge in unit 0)
// slot of one row: plane 0 = 16-byte units with columns 0, 1; plane 1 = columns 2, 3 (16-byte units) or column 2 alone
// (8-byte units, strip width 3); slots stay 16-byte aligned
template <int C> constexpr int kPoaRowsSlotBytes = C == 2 ? kPoaPlane : (C == 3 ? kPoaPlane + 33 * 8 + 8 : 2 * kPoaPlane);
// byte offset, inside a slot, of the LAST column's H key of unit u (u = lane + 1: the lane's own, u = lane: the previous
// lane's, u = 0: the band's left edge)
template <int C> __device__ __forceinline__ uint32_t poa_rows_lasth(uint32_t u)
{
    return C == 2 ? u * 16 + 8 : (C == 3 ? kPoaPlane + u * 8 : kPoaPlane + u * 16 + 8);
}

// History: slot [row & 31] (+ the three constant slots) holds the keys of a row as C/2 planes of 16-byte units,
// unit [lane + 1] of plane q = (H, Eout) keys of the lane's columns 2q and 2q+1.  All lanes read the same slot at
// the same time: consecutive 16-byte units, conflict-free.  Unit 0 of the last plane carries the band's left edge
// H(row, j0 - 1) in its third word, so that "the previous lane's last H" is one uniform load for every lane.
template <int C>
struct PoaRowsSmem {
    unsigned char hist[kPoaSlots * kPoaRowsSlotBytes<C>];
    int2 bring[kPoaBndRing];                 // left boundary keys (H(r, j0-1)*256, Fg(r, j0)*256 + 64) by row
    uint4 desc[kPoaRowsDescRing];            // row descriptors, copied in by cp.async two 32-row chunks ahead
    int kconst[20];
};

constexpr unsigned kPoaNoFwd = 15u;          // descriptor bits 28..31: index of the predecessor that is row r-1 (15: none)
constexpr unsigned kPoaPlainRow = 16u;       // flag: the row's ONLY predecessor is row r-1, no spill, not a head: inputs come from registers

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void sts_u32_if(uint32_t addr, uint32_t v, bool doit)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((unsigned)doit) : "memory");
}

template <int C>
__device__ __forceinline__ void poa_rows_lds(uint32_t slot_addr, uint32_t lane16, int& Hd0, int (&h)[C], int (&e)[C])
{
    const uint32_t own = slot_addr + lane16;                    // lane16 = lane * 16: unit [lane] (the previous lane's)
    const int4 v01 = lds_s4(own + 16);
    h[0] = v01.x; e[0] = v01.y; h[1] = v01.z; e[1] = v01.w;
    if constexpr (C == 4) {
        const int4 v23 = lds_s4(own + 16 + kPoaPlane);
        h[2] = v23.x; e[2] = v23.y; h[3] = v23.z; e[3] = v23.w;
    }
    if constexpr (C == 3) {
        const uint2 v2 = lds_u2(slot_addr + kPoaPlane + (lane16 >> 1) + 8);
        h[2] = (int)v2.x; e[2] = (int)v2.y;
    }
    Hd0 = lds_s32(slot_addr + poa_rows_lasth<C>(lane16 >> 4));
}

// the keys of one row into its slot: k[2c] = H key, k[2c+1] = E key of column c
template <int C>
__device__ __forceinline__ void poa_rows_sts(uint32_t slot_addr, uint32_t lane16, const int (&k)[2 * C])
{
    sts_s4(slot_addr + lane16 + 16, k[0], k[1], k[2], k[3]);
    if constexpr (C == 4) sts_s4(slot_addr + kPoaPlane + lane16 + 16, k[4], k[5], k[6], k[7]);
    if constexpr (C == 3) sts_u2(slot_addr + kPoaPlane + (lane16 >> 1) + 8, (uint32_t)k[4], (uint32_t)k[5]);
}

// One band (32*C columns, one warp) over all rows.  LAST: the band that holds column len-1 also writes lastsorce.
// EDGE0: band 0 when the caller did not supply hneg1 (device-resident graphs, tsta_poa_fill_resident): the left edge
// H(r,-1) = max_j H(p_j,-1) + E (c-t-simd.c:196-200; O+E for a head, 187) is the DP's "column -1" and is carried like
// any other column -- every lane computes it from the edge words the history already holds.
template <int C, bool LAST, bool EDGE0>
__device__ __forceinline__ void poa_rows_sweep(const PoaFillArgs& a, PoaRowsSmem<C>& sm, const int band, const int lane)
{
    constexpr int kSlotB = kPoaRowsSlotBytes<C>;
    const int E = a.E, O = a.O, OE = a.O + a.E;
    const int R = a.nrows, len = a.len, L = a.L;
    const int j0 = (band * 32 + lane) * C;
    uint8_t rc[C];
    bool isL[C], isLm1[C];
    bool hasL = false;
#pragma unroll
    for (int c = 0; c < C; c++) {
        rc[c] = (j0 + c < len) ? a.read[j0 + c] : 0;
        isL[c] = (j0 + c == L);
        isLm1[c] = (j0 + c + 1 == L);
        hasL |= isL[c] || isLm1[c];
    }
    const bool band_has_L = __any_sync(kFullMask, hasL);        // late heads take the serial row only here
    const bool owns_last = LAST && (len - 1 >= j0) && (len - 1 < j0 + C);
    const int clast = owns_last ? len - 1 - j0 : 0;
    const unsigned long long* bnd_in = a.bnd + (long long)(band - 1) * R;
    unsigned long long* bout = a.bnd + (long long)band * R;     // right-edge element of the current row
    const bool store_right = (band + 1 < a.nbands) && lane == 31;
    const unsigned ep16 = a.epoch << 1;

    // Scoring constants, their multiples and the shared-memory base go through shared memory once: values that ptxas
    // can trace back to kernel parameters or to the shared window are re-derived inside the row loop (LDC + IMAD,
    // S2R SR_CgaCtaId + LEA with ~20 cycles of scoreboard wait each) instead of being kept in registers.
    const int ESC0 = 256 * E * C;
    if (lane == 0) {
        int* k = sm.kconst;
        k[0] = 256 * a.M; k[1] = 256 * a.X; k[2] = 256 * E; k[3] = 256 * OE + 64; k[4] = 256 * E + 64; k[5] = 256 * E + 128;
        k[6] = 256 * a.M + 255; k[7] = 256 * a.X + 255; k[8] = 512 * E; k[9] = 768 * E;
        k[10] = ESC0; k[11] = 2 * ESC0; k[12] = 4 * ESC0; k[13] = 8 * ESC0; k[14] = 16 * ESC0;
        k[18] = 3 * ESC0; k[19] = 12 * ESC0;
        k[15] = (int)(uint32_t)__cvta_generic_to_shared(&sm.hist[0]);
        const long long sb = a.stride * 2;                        // bytes between two trace rows
        k[16] = (int)(unsigned)(sb & 0xffffffffll); k[17] = (int)(sb >> 32);
    }
    __syncwarp();
    const volatile int* kc = sm.kconst;
    const int MS = kc[0], XS = kc[1], ES = kc[2], OES = kc[3], ES64 = kc[4], ES128 = kc[5];
    const int MS255 = kc[6], XS255 = kc[7], ES2 = kc[8], ES3 = kc[9];
    const int D1 = kc[10], D2 = kc[11], D4 = kc[12], D8 = kc[13], D16 = kc[14], D3 = kc[18], D12 = kc[19];
    const uint32_t base = (uint32_t)kc[15];
    const long long stride_bytes = (long long)(((unsigned long long)(unsigned)kc[17] << 32) | (unsigned)kc[16]);
    constexpr uint32_t kDescOff = (uint32_t)offsetof(PoaRowsSmem<C>, desc);
    constexpr uint32_t kBringOff = (uint32_t)offsetof(PoaRowsSmem<C>, bring);
    const uint32_t lane16 = (uint32_t)lane * 16u;
    const int cin_dec = D1 * lane;                              // the band's left-edge F reaches this lane's first column

    // ---- constant history slots (as in poa_fill_kernel; c-t-simd.c:176-188 for the virtual rows) ----
    {
        const uint32_t none = base + kPoaSlotNone * kSlotB;
        int kmin[2 * C];
#pragma unroll
        for (int q = 0; q < 2 * C; q++) kmin[q] = kPoaKeyMin;
        poa_rows_sts<C>(none, lane16, kmin);
        if (lane == 0) sts_u32(none + poa_rows_lasth<C>(0), (uint32_t)kPoaKeyMin);
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
            const uint32_t slot = base + (kPoaSlotHead + v) * kSlotB;
            poa_rows_sts<C>(slot, lane16, k);
            if (lane == 0) {   // left edge: diagonal input of column j0
                const int Hd0 = (late && j0 <= L - 1) ? j0 * E : (j0 == 0 ? 0 : O + j0 * E);
                sts_u32(slot + poa_rows_lasth<C>(0), (uint32_t)(Hd0 * 256));
            }
        }
    }

    // ---- prologue: descriptors of rows [0, 64) copied in; first boundary group in flight ----
    if (lane < R) cp_async16(base + kDescOff + lane * 16, a.desc + lane);
    if (lane + 32 < R) cp_async16(base + kDescOff + (lane + 32) * 16, a.desc + lane + 32);
    unsigned long long pre_b = 0;
    int pre_h = 0;
    if (lane < kPoaGroup && lane < R) {
        if (band > 0) pre_b = ld_relaxed_u64(bnd_in + lane);
        else if constexpr (!EDGE0) pre_h = __ldg(a.hneg1 + lane);
    }
    cp_async_wait_all();
    __syncwarp();

    int ko[2 * C];                       // keys of the row just finished (forwarded to the next row)
#pragma unroll
    for (int q = 0; q < 2 * C; q++) ko[q] = kPoaKeyMin;
    int fwdH = kPoaKeyMin;               // H(r-1, j0-1): the previous lane's last H of the row just finished
    int edge_prev = kPoaKeyMin;          // EDGE0: H(r-1, -1) as a key
    const uint32_t kEdgeOff = poa_rows_lasth<C>(0);              // the left-edge word of a history slot
    const int OEk = OES - 64;            // (O+E)*256
    uint4 dsc_next = lds_u4(base + kDescOff);
    char* trow = reinterpret_cast<char*>(a.trace + j0);   // this lane's cells of row r
    uint32_t desc_off = 16, bring_off = 0;   // ring offsets of row r+1's descriptor / row r's boundary element
    int* lastp = a.lastsorce;

    // The row proper, from the reduced predecessor keys: hk = best diagonal key + substitution score, ke = best
    // vertical key.  F by scan (poa_rows.cuh header), then every column; history, trace, right edge.
    auto finish_row = [&](const int r, const int (&hk)[C], const int (&ke)[C], const int Hedge, const int Fin,
                          const unsigned flags, const unsigned spill_row) {
        uint32_t tw[(C + 1) / 2];
        int Fout;                                               // Fg key entering the next lane's first column
        // this row's output addresses; the running pointers move on NOW, a row away from the stores that use them
        // (an increment right behind its store waits ~16 cycles for the store to release the address registers)
        char* const tcur = trow;
        unsigned long long* const bcur = bout;
        trow += stride_bytes;
        bout++;
        if ((flags & kPoaLate) && band_has_L) {
            // ---- serial row (late head in the band that re-bases F): the carry moves one lane per pass.
            //      poa_row adds the substitution score itself: hand it hk minus that score ----
            int kd[C], kee[C];
#pragma unroll
            for (int c = 0; c < C; c++) { kd[c] = hk[c]; kee[c] = ke[c]; }
            int Fk_in = lane == 0 ? Fin : kPoaKeyMin;
            Fout = kPoaKeyMin;
#pragma unroll 1
            for (int pass = 0; pass < 32; pass++) {
                int Fk = Fk_in, Hl = 0;
                poa_row<C, true>(kd, kee, rc, isL, isLm1, -1, 0, 0, ES, ES64, ES128, OES, 256 * O, Fk, Hl, ko, tw);
                Fout = Fk;
                const int up = __shfl_up_sync(kFullMask, Fout, 1);
                Fk_in = lane == 0 ? Fin : up;
            }
        } else {
            // ---- pass 1: H^ (no F), "opened from H^" keys T, local prefixes P, lane total G ----
            int T[C], P[C];
#pragma unroll
            for (int c = 0; c < C; c++) T[c] = (max(hk[c], ke[c]) & ~255) + OES;
            P[0] = kPoaKeyMin;
#pragma unroll
            for (int c = 1; c < C; c++) P[c] = c == 1 ? T[0] : __viaddmax_s32(P[c - 1], ES, T[c - 1]);
            const int G = __viaddmax_s32(P[C - 1], ES, T[C - 1]);
            // ---- exclusive max-plus scan over the lanes, radix 4: three dependent shuffle stages (1..4, then 4/8/12,
            //      then 16).  Stage 1 forms the EXCLUSIVE prefix over four lanes, so lanes 0..3 must not take the value
            //      shfl_up hands back to them; in the later stages a lane that gets its own value back keeps it (E <= 0) ----
            const int g1 = __shfl_up_sync(kFullMask, G, 1), g2 = __shfl_up_sync(kFullMask, G, 2);
            const int g3 = __shfl_up_sync(kFullMask, G, 3), g4 = __shfl_up_sync(kFullMask, G, 4);
            int S = lane >= 1 ? g1 : kPoaKeyMin;                  // (selects, not branches: the lanes diverge here)
            S = __viaddmax_s32(lane >= 2 ? g2 : kPoaKeyMin, D1, S);
            S = __viaddmax_s32(lane >= 3 ? g3 : kPoaKeyMin, D2, S);
            S = __viaddmax_s32(lane >= 4 ? g4 : kPoaKeyMin, D3, S);
            {
                const int s4 = __shfl_up_sync(kFullMask, S, 4), s8 = __shfl_up_sync(kFullMask, S, 8), s12 = __shfl_up_sync(kFullMask, S, 12);
                S = __viaddmax_s32(s4, D4, S);
                S = __viaddmax_s32(s8, D8, S);
                S = __viaddmax_s32(s12, D12, S);
            }
            S = __viaddmax_s32(__shfl_up_sync(kFullMask, S, 16), D16, S);
            const int Cin = __viaddmax_s32(Fin, cin_dec, S);      // lane 0: S = -inf, cin_dec = 0
            // ---- pass 2: every column from its exact F ----
            uint32_t lo8[C], hi8[C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int Fk = c == 0 ? Cin : __viaddmax_s32(Cin, c == 1 ? ES : (c == 2 ? ES2 : ES3), P[c]);
                const int Hk = __vimax3_s32(hk[c], ke[c], Fk);
                const int H256 = Hk & ~255;
                const int t = H256 + OES;                                     // (H+O+E)*256 + 64
                const int kem = ke[c] & ~255;
                const int WE = __viaddmax_s32(kem, ES128, t);              // low byte 128 iff be+E >= H+O+E
                const int WF = __viaddmax_s32(Fk, ES64, t);                // low byte 128 iff Fg+E >= H+O+E
                ko[2 * c] = H256;
                ko[2 * c + 1] = __viaddmax_s32(kem, ES64, t);              // Eout*256 + 64
                if (c == C - 1) Fout = __viaddmax_s32(Fk, ES, t);
                lo8[c] = (uint32_t)Hk;
                const uint32_t t1 = ((uint32_t)WE & 0x40u) | ((uint32_t)ke[c] & ~0x40u);
                hi8[c] = (t1 & ~0x80u) | ((uint32_t)WF & 0x80u);          // byte 0: WF bit, WE bit, 63 - j* of the best E
            }
#pragma unroll
            for (int c = 0; c + 1 < C; c += 2)
                tw[c / 2] = __byte_perm(__byte_perm(lo8[c], lo8[c + 1], 0x0040), __byte_perm(hi8[c], hi8[c + 1], 0x0040), 0x5140);
            if constexpr (C & 1) tw[C / 2] = __byte_perm(lo8[C - 1], hi8[C - 1], 0x0040);
        }
        // ---- outputs: history slot, spill, trace, last column, right edge ----
        const uint32_t slot = base + ((uint32_t)r & (kPoaHist - 1)) * kSlotB;
        poa_rows_sts<C>(slot, lane16, ko);
        sts_u32_if(slot + poa_rows_lasth<C>(0), (uint32_t)Hedge, lane == 0);     // lane 0: the band's left edge (unit 0)
        if (flags & kPoaKeep) {    // a successor further away than the history ring will read it back
            if constexpr (EDGE0) a.edge_spill[spill_row] = Hedge;         // (every lane, the same value)
            int2* g = a.HE + (long long)spill_row * a.stride + j0;
#pragma unroll
            for (int c = 0; c < C; c++) g[c] = make_int2(ko[2 * c], ko[2 * c + 1]);
        }
        if constexpr (C == 4) __stcs(reinterpret_cast<uint2*>(tcur), make_uint2(tw[0], tw[1]));
        else if constexpr (C == 2) __stcs(reinterpret_cast<uint32_t*>(tcur), tw[0]);
        else {   // strip width 3: six bytes per lane, 2-byte aligned
            __stcs(reinterpret_cast<unsigned short*>(tcur), (unsigned short)tw[0]);
            __stcs(reinterpret_cast<unsigned short*>(tcur) + 1, (unsigned short)(tw[0] >> 16));
            __stcs(reinterpret_cast<unsigned short*>(tcur) + 2, (unsigned short)tw[1]);
        }
        if constexpr (LAST) {
            if (owns_last) {
                int hl = ko[0];
#pragma unroll
                for (int c = 1; c < C; c++) hl = clast == c ? ko[2 * c] : hl;
                lastp[r] = hl >> 8;
            }
        }
        st_relaxed_u64_if(bcur, poa_bnd_pack(ko[2 * (C - 1)], Fout, ep16), store_right);
        const int upH = __shfl_up_sync(kFullMask, ko[2 * (C - 1)], 1);
        fwdH = lane == 0 ? Hedge : upH;
        if constexpr (EDGE0) edge_prev = Hedge;
        __syncwarp();                                           // this row's history stores before the next row's loads
    };

    for (int r0 = 0; r0 < R; r0 += 32) {
        {   // the descriptors of rows [r0+32, r0+64) have arrived (copy issued one chunk ago); start [r0+64, r0+96)
            cp_async_wait_all();
            const int rown = r0 + 64 + lane;
            if (rown < R) cp_async16(base + kDescOff + ((uint32_t)rown & (kPoaRowsDescRing - 1)) * 16, a.desc + rown);
        }
#pragma unroll 1
        for (int t0 = r0; t0 < min(R, r0 + 32); t0 += kPoaGroup) {
            {   // import the left boundary of rows [t0, t0+8) (lanes 0..7), prefetch the next group
                const int row = lane < kPoaGroup ? t0 + lane : R;
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
                    if (row + kPoaGroup < R) pre_b = ld_relaxed_u64(bnd_in + row + kPoaGroup);
                } else if constexpr (!EDGE0) {
                    if (row < R) sm.bring[row & (kPoaBndRing - 1)] = make_int2(pre_h * 256, (pre_h + OE) * 256 + 64);   // Fg(r,0) = H(r,-1)+O+E
                    if (row + kPoaGroup < R) pre_h = __ldg(a.hneg1 + row + kPoaGroup);
                }
                __syncwarp();
            }
            const int rend = min(R, t0 + kPoaGroup);
#pragma unroll 1
            for (int r = t0; r < rend; r++) {
                const uint4 dsc = dsc_next;
                dsc_next = lds_u4(base + kDescOff + desc_off);
                const uint2 bv = lds_u2(base + kBringOff + bring_off);   // one address: a broadcast
                desc_off = (desc_off + 16) & (kPoaRowsDescRing * 16 - 1);
                bring_off = (bring_off + 8) & (kPoaBndRing * 8 - 1);
                int Hedge = (int)bv.x, Fin = (int)bv.y;
                const int b = (int)(dsc.x & 0xffu);
                const unsigned flags = (dsc.x >> 16) & 0xffu;
                int hk[C], ke[C];
                if (flags & kPoaPlainRow) {
                    if constexpr (EDGE0) { Hedge = edge_prev + ES; Fin = Hedge + OES; }
                    // ---- the only predecessor is the row just finished: registers, weights folded into the score ----
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        hk[c] = (c == 0 ? fwdH : ko[2 * (c - 1)]) + ((int)rc[c] == b ? MS255 : XS255);
                        ke[c] = ko[2 * c + 1] + 127;
                    }
                    finish_row(r, hk, ke, Hedge, Fin, 0u, 0u);
                    continue;
                }
                // ---- predecessors: kd = best diagonal input, ke = best vertical input (keys, poa_fill.cuh) ----
                const int np = (int)((dsc.x >> 8) & 0xffu);
                int kd[C];
                int emax = kPoaKeyMin;                                  // EDGE0: max_j H(p_j, -1)
                if (!(flags & kPoaSlow)) {
                    // Predecessors 0 and 1 in one straight line: both slots are read unconditionally with all loads in
                    // flight together (an absent second predecessor is the "none" slot: keys of -infinity) -- one
                    // shared-memory latency, no branches.  Row r-1 (index jf, if it is among them) still comes from the
                    // registers: its slot is swapped for "none" and its keys join afterwards, so the row-to-row chain
                    // never goes through shared memory.  44 % of the rows of an MSA graph come through here.
                    const int jf = (int)(dsc.x >> 28);
                    {
                        const uint32_t none = base + kPoaSlotNone * kSlotB;
                        const uint32_t sa = jf == 0 ? none : base + (dsc.y & 0xffu) * kSlotB;
                        const uint32_t sb = jf == 1 ? none : base + ((dsc.y >> 8) & 0xffu) * kSlotB;
                        int Ha, ha[C], ea[C], Hb, hb[C], eb[C];
                        poa_rows_lds<C>(sa, lane16, Ha, ha, ea);
                        poa_rows_lds<C>(sb, lane16, Hb, hb, eb);
                        if constexpr (EDGE0) emax = max(lds_s32(sa + kEdgeOff), lds_s32(sb + kEdgeOff));
                        kd[0] = __viaddmax_s32(Hb, 254, Ha + 255);
#pragma unroll
                        for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(hb[c - 1], 254, ha[c - 1] + 255);
#pragma unroll
                        for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(eb[c], 126, ea[c] + 127);
                        if (jf < 2) {
                            const int wd = 255 - jf, we = 127 - jf;
                            if constexpr (EDGE0) emax = max(emax, edge_prev);
                            kd[0] = __viaddmax_s32(fwdH, wd, kd[0]);
#pragma unroll
                            for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(ko[2 * (c - 1)], wd, kd[c]);
#pragma unroll
                            for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(ko[2 * c + 1], we, ke[c]);
                        }
                    }
#pragma unroll 1
                    for (int j = 2; j < np; j++) {                     // (3 % of the rows)
                        int Hd0, ph[C], pe[C];
                        if (j == jf) {
                            Hd0 = fwdH;
                            if constexpr (EDGE0) emax = max(emax, edge_prev);
#pragma unroll
                            for (int c = 0; c < C; c++) { ph[c] = ko[2 * c]; pe[c] = ko[2 * c + 1]; }
                        } else {
                            const uint32_t slot = ((j < 4 ? dsc.y : dsc.z) >> (8 * (j & 3))) & 0xffu;
                            poa_rows_lds<C>(base + slot * kSlotB, lane16, Hd0, ph, pe);
                            if constexpr (EDGE0) emax = max(emax, lds_s32(base + slot * kSlotB + kEdgeOff));
                        }
                        const int wd = 192 + 63 - j, we = 64 + 63 - j;
                        kd[0] = __viaddmax_s32(Hd0, wd, kd[0]);
#pragma unroll
                        for (int c = 1; c < C; c++) kd[c] = __viaddmax_s32(ph[c - 1], wd, kd[c]);
#pragma unroll
                        for (int c = 0; c < C; c++) ke[c] = __viaddmax_s32(pe[c], we, ke[c]);
                    }
                } else {
                    // any in-degree, predecessors further back than the ring (HBM spill); row r-1 is read from the ring
#pragma unroll
                    for (int c = 0; c < C; c++) { kd[c] = kPoaKeyMin; ke[c] = kPoaKeyMin; }
                    const int e0 = (int)dsc.z;
                    for (int j = 0; j < np; j++) {
                        const int p = __ldg(a.pred_row + e0 + j);
                        int ph[C], pe[C], Hd0;
                        if (p < 0 || r - p <= kPoaNear) {
                            const int slot = p < 0 ? ((flags & kPoaLate) ? kPoaSlotLate : kPoaSlotHead) : (p & (kPoaHist - 1));
                            poa_rows_lds<C>(base + slot * kSlotB, lane16, Hd0, ph, pe);
                            if constexpr (EDGE0) emax = max(emax, lds_s32(base + slot * kSlotB + kEdgeOff));
                        } else {
                            if constexpr (EDGE0) emax = max(emax, __ldcg(a.edge_spill + __ldg(&a.desc[p].w)));
                            const int2* src = a.HE + (long long)__ldg(&a.desc[p].w) * a.stride + j0;
#pragma unroll
                            for (int c = 0; c < C; c++) {
                                const int2 v = __ldcg(src + c);
                                ph[c] = v.x; pe[c] = v.y;
                            }
                            if (lane > 0) Hd0 = __ldcg(src - 1).x;
                            else if (band > 0) Hd0 = (int)(unsigned)ld_relaxed_u64(bnd_in + p);
                            else if constexpr (EDGE0) Hd0 = __ldcg(a.edge_spill + __ldg(&a.desc[p].w));
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
                if constexpr (EDGE0) { Hedge = (flags & kPoaHead) ? OEk : emax + ES; Fin = Hedge + OES; }
#pragma unroll
                for (int c = 0; c < C; c++) hk[c] = kd[c] + ((int)rc[c] == b ? MS : XS);
                finish_row(r, hk, ke, Hedge, Fin, flags, dsc.w);
            }
        }
    }
}

template <int C>
__global__ void __launch_bounds__(32) poa_rows_kernel(const __grid_constant__ PoaFillArgs a)
{
    static_assert(C >= 2 && C <= 4, "strip widths of the row-synchronous POA fill");
    extern __shared__ __align__(16) unsigned char poa_smem_raw[];
    PoaRowsSmem<C>& sm = *reinterpret_cast<PoaRowsSmem<C>*>(poa_smem_raw);
    const int lane = threadIdx.x;
    int band = 0;
    if (lane == 0) band = (int)atomicAdd(a.ticket, 1u);
    band = __shfl_sync(kFullMask, band, 0);
    const bool edge0 = band == 0 && a.hneg1 == nullptr;
    if (band + 1 == a.nbands) {
        if (edge0) poa_rows_sweep<C, true, true>(a, sm, band, lane);
        else poa_rows_sweep<C, true, false>(a, sm, band, lane);
    } else if (edge0) poa_rows_sweep<C, false, true>(a, sm, band, lane);
    else poa_rows_sweep<C, false, false>(a, sm, band, lane);
}

}  // namespace tsta
