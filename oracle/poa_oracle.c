/*
 * oracle/poa_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * Plain-C int32 restatement of the TSTA sequence-to-graph DP (POA fill):
 *   block_line_alignment()   reference msa/c-t-simd.c:165-599  (one graph node x one column block)
 * and of the part of node_fuse() that WALKS the trace (msa/c-t-simd.c:751-785 control flow,
 * including the forced-continuation rewrites at 767-769 and 776-782).
 *
 * The reference keeps int8 differences with three saturating limbs per predecessor; inside
 * its domain that is the exact recurrence below on absolute scores.  Rows are
 * graph nodes in p->sort[] order; predecessors are listed in pre[] order (ties pick the
 * lowest index); columns are read positions.
 *
 *   H(r,-1)   = O+E for heads, else max_j H(p_j,-1)+E                         (187, 196-200)
 *   heads use a virtual predecessor: Hv(-1)=0, Hv(c)=O+(c+1)E, Eout_v(c)=Hv(c)+O+E (174-188)
 *   Fg(r,0)   = H(r,-1)+O+E ; Fg(r,c+1) = max(Fg(r,c)+E, H(r,c)+O+E)
 *   H(r,c)    = max( max_j H(p_j,c-1)+sub , max_j Eout(p_j,c) , Fg(r,c) ), sub = M/X
 *               read[c]=='N' -> H(r,c)=H(r,c-1)                               (461, 481)
 *   Eout(r,c) = max( H(r,c)+O+E , max_j Eout(p_j,c)+E )
 *   source    : 42+j match-diag (lowest j) > 84+j mismatch-diag > j (Eout(p_j,c)==H) > 126 (484-505)
 *   esource   : j* = lowest j with Eout(p_j,c)==max; +-((opened?43:1)+j*)       (509-538)
 *   fsource   : +-(opened?2:1)                                                 (521-522,564-566)
 *
 * Quirk Q1 (late heads): a head row whose index r >= L sees, in columns
 * c <= L-1, a virtual row WITHOUT the open penalty, and its carried Fg drops by |O| when it
 * enters column L (c-t-simd.c:176-185, 228-233, 570-571).  L is the reference's tile edge
 * (SIMD lanes x -W): results of the reference itself depend on it.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-1000000000)

typedef struct {
    int32_t row;
    int32_t col;
    uint8_t kind;   /* 0 vertical, 1 match, 2 mismatch, 3 horizontal */
    uint8_t pred;
    uint16_t pad_;
} oracle_poa_step;

static inline int imax(int a, int b) { return a > b ? a : b; }

/*
 * Fill.  source/esource/fsource are nrows x len signed bytes (row-major, unstriped).
 * lastsorce[r] = H(r, len-1) for EVERY row (the reference stores it for sinks only, 577-578).
 * Memory: two int32 rows per graph row are kept for the whole fill (predecessors may be
 * arbitrarily far back), i.e. 8 * nrows * len bytes -- sized for test inputs.
 */
int tsta_oracle_poa_fill(int nrows, const int32_t* row_ptr, const int32_t* pred_row, const uint8_t* base,
                         const int32_t* hneg1, const char* read, int len,
                         int M, int X, int E, int O, int L,
                         int8_t* source, int8_t* esource, int8_t* fsource, int32_t* lastsorce)
{
    const int OE = O + E;
    const size_t cells = (size_t)nrows * (size_t)len;
    int32_t* Hm = (int32_t*)malloc((cells ? cells : 1) * sizeof(int32_t));
    int32_t* Em = (int32_t*)malloc((cells ? cells : 1) * sizeof(int32_t));
    if (!Hm || !Em) { free(Hm); free(Em); return -1; }
    for (int r = 0; r < nrows; r++) {
        const int e0 = row_ptr[r], e1 = row_ptr[r + 1];
        const int npre = e1 - e0;
        const int head = (npre == 1 && pred_row[e0] < 0);
        const int late = head && L > 0 && (r / L > 0);
        int32_t* Hr = Hm + (size_t)r * len;
        int32_t* Er = Em + (size_t)r * len;
        int Hleft = hneg1[r];
        int Fg = Hleft + OE;
        for (int c = 0; c < len; c++) {
            const int sub = ((unsigned char)read[c] == base[r]) ? M : X;
            if (late && c == L) Fg += O;                       /* Q1: carried F re-based (228-233, 570-571) */
            int best_diag = NEG_INF, best_diag_j = 0;
            int best_e = NEG_INF, best_e_j = 0, best_e_H = 0;
            for (int j = 0; j < npre; j++) {
                int Hd, Ee, Hp;
                if (head) {
                    if (late && c <= L - 1) {                  /* Q1 virtual row without O */
                        Hd = c * E;
                        Hp = (c + 1) * E;
                    } else {
                        Hd = c == 0 ? 0 : O + c * E;
                        Hp = O + (c + 1) * E;
                    }
                    Ee = Hp + OE;
                } else {
                    const int p = pred_row[e0 + j];
                    Hd = c == 0 ? hneg1[p] : Hm[(size_t)p * len + c - 1];
                    Hp = Hm[(size_t)p * len + c];
                    Ee = Em[(size_t)p * len + c];
                }
                if (Hd > best_diag) { best_diag = Hd; best_diag_j = j; }
                if (Ee > best_e) { best_e = Ee; best_e_j = j; best_e_H = Hp; }
            }
            const int h = best_diag + sub;
            int H = imax(h, imax(best_e, Fg));
            if (read[c] == 'N') H = Hleft;
            /* source: F by default, then E, then mismatch-diag, then match-diag (484-505);
             * inside a class the lowest predecessor index wins */
            int src = 126;
            if (best_e == H) src = 0 + best_e_j;
            if (h == H) src = (sub == M ? 42 : 84) + best_diag_j;
            /* esource (509-538) */
            const int eopen = (best_e - best_e_H == OE);
            int es = (eopen ? 43 : 1) + best_e_j;
            if (best_e + E == H + OE) es = -es;
            /* fsource (521-522, 564-566) */
            int fs = (Fg - Hleft == OE) ? 2 : 1;
            if (Fg + E == H + OE) fs = -fs;
            source[(size_t)r * len + c] = (int8_t)src;
            esource[(size_t)r * len + c] = (int8_t)es;
            fsource[(size_t)r * len + c] = (int8_t)fs;
            Hr[c] = H;
            Er[c] = imax(H + OE, best_e + E);
            Fg = imax(Fg + E, H + OE);
            Hleft = H;
        }
        lastsorce[r] = len > 0 ? Hr[len - 1] : hneg1[r];
    }
    free(Hm);
    free(Em);
    return 0;
}

/*
 * Walk (node_fuse's control flow, msa/c-t-simd.c:751-956): from (start_row, len-1) follow
 * `source`, applying the two rewrites the reference performs ahead of itself:
 *   horizontal at (r,c), c>=2, and (|fsource(r,c)|==1 or (|fsource(r,c)|==2 and fsource(r,c-1)<0))
 *        -> source(r,c-1) := 126                                            (767-769)
 *   vertical at (r,c) to predecessor row p>0, and (|esource(r,c)|<=42 or
 *        (|esource(r,c)|>42 and esource(p,c)<0)) -> source(p,c) := |esource(p,c)%42|-1   (776-782)
 * The matrices are modified in place exactly like the reference does.
 * Returns the number of steps written to path (capacity nrows+len).
 */
int tsta_oracle_poa_walk(int nrows, const int32_t* row_ptr, const int32_t* pred_row, int len,
                         int8_t* source, const int8_t* esource, const int8_t* fsource,
                         int start_row, oracle_poa_step* path)
{
    (void)nrows;
    int n = 0, num1 = start_row, num2 = len - 1;
    while (num1 != -1 && num2 != -1) {
        const size_t at = (size_t)num1 * len + num2;
        const int S = source[at];
        const int kind = S / 42, pj = S % 42;
        path[n].row = num1;
        path[n].col = num2;
        path[n].kind = (uint8_t)kind;
        path[n].pred = (uint8_t)(kind == 3 ? 0 : pj);
        path[n].pad_ = 0;
        n++;
        if (kind == 3) {
            const int f = fsource[at];
            if (num2 - 1 > 0 && ((f == 1 || f == -1) || ((f == 2 || f == -2) && fsource[at - 1] < 0)))
                source[at - 1] = 126;
            num2--;
        } else if (kind == 0) {
            const int p = pred_row[row_ptr[num1] + pj];
            const int e = esource[at];
            if (p > 0 && ((e <= 42 && e >= -42) || ((e > 42 || e < -42) && esource[(size_t)p * len + num2] < 0))) {
                int s5 = esource[(size_t)p * len + num2] % 42;
                s5 = (s5 >= 0 ? s5 : -s5) - 1;
                source[(size_t)p * len + num2] = (int8_t)s5;
            }
            num1 = p;
        } else {
            num1 = pred_row[row_ptr[num1] + pj];
            num2--;
        }
    }
    return n;
}
