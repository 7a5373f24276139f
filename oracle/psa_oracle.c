/*
 * oracle/psa_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * Plain-C int32 restatement of the TSTA pairwise global-alignment hot path:
 *   - score-matrix fill (affine / linear gaps)          reference psa/psa.c:219-363
 *   - the three direction matrices back/fback/eback       reference psa/psa.c:294-327
 *   - the traceback walk and the >1/>2 strings            reference psa/psa.c:435-519
 *
 * The reference works on int8 *differences* in a striped SIMD layout; inside its
 * parameter domain (O<=0, E<=0, small |M|,|X|) that is arithmetically identical to
 * the textbook Gotoh recurrences on absolute int32 scores, which is what is written
 * here.  The oracle is pinned against the real reference binary (oracle/_ref, built
 * by oracle/build_ref.sh) and against the reference's published goldens by
 * tests/test_oracle_psa.py.
 *
 * Conventions (reference psa/psa.c:194-199): s0 is the LONGER read (columns j, length
 * n), s1 the shorter (rows i, length m).  The caller does the swap.
 *
 * Two traceback implementations are provided on purpose:
 *   tsta_oracle_psa_ref3()  keeps three byte matrices exactly like the reference and
 *                           walks them exactly like trace() does (including the
 *                           in-place rewrite of `back`);
 *   tsta_oracle_psa()       keeps ONE 4-bit code per cell (back | cF<<2 | cE<<3) and
 *                           walks it with a 3-state walker -- the form the CUDA path
 *                           stores in HBM.
 * Tests require both to agree byte for byte.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-1000000000)
#define MS_SEED (-100) /* reference: #define MIN -100, ms = MIN (psa/psa.c:110,590) */

static inline int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
    int maxsorce;        /* max over all real cells of H, seeded with -100 (psa/psa.c:357,416-422,590) */
    int corner;          /* H(m-1, n-1): the true global score (not printed by the reference)           */
    uint64_t lastrow_fnv;/* FNV-1a over the int32 values H(m-1, 0..n-1)                                   */
    int64_t alen;        /* alignment length (0 when no traceback was requested)                         */
} tsta_oracle_result;

static uint64_t fnv1a_i32(uint64_t h, int32_t v)
{
    for (int k = 0; k < 4; k++) {
        h ^= (uint64_t)((uint32_t)v >> (8 * k)) & 0xffu;
        h *= 1099511628211ULL;
    }
    return h;
}

/*
 * Core fill.  When nib != NULL it receives one byte per cell (row-major, m x n):
 *   bits 0-1  back   : 1 diagonal, 0 left (F), 2 up (E)             psa/psa.c:294-298
 *   bit  2    cF     : j>=1 && (fback(i,j)==1 || fback(i,j-1)==-2)   psa/psa.c:450-451
 *   bit  3    cE     : i>=1 && (eback(i,j)==1 || eback(i-1,j)==-2)   psa/psa.c:457-458
 * When b3 != NULL it receives the three reference matrices (back, fback, eback) as
 * signed bytes, each m x n, concatenated.
 * lastrow (may be NULL) receives H(m-1, j).
 */
static void psa_fill(const char* s0, int64_t n, const char* s1, int64_t m,
                     int M, int X, int E, int O,
                     uint8_t* nib, int8_t* b3, int32_t* lastrow, tsta_oracle_result* res)
{
    const int OE = O + E;
    int32_t* Hrow = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t)); /* H(i-1, j), index j+1 */
    int32_t* Erow = (int32_t*)malloc((size_t)n * sizeof(int32_t));       /* Eg(i-1, j)           */
    int8_t* fbrow = (int8_t*)malloc((size_t)n);                           /* fback(i, .)          */
    int8_t* ebprev = (int8_t*)malloc((size_t)n);                          /* eback(i-1, .)        */
    int ms = MS_SEED;

    /* boundary row -1 (psa/psa.c:221-226): H(-1,-1)=0, H(-1,j)=O+(j+1)E, no vertical gap yet */
    Hrow[0] = 0;
    for (int64_t j = 0; j < n; j++) {
        Hrow[j + 1] = O + (int)(j + 1) * E;
        Erow[j] = NEG_INF;
        ebprev[j] = 0;
    }
    for (int64_t i = 0; i < m; i++) {
        /* boundary column -1 (psa/psa.c:228-231): H(i,-1)=O+(i+1)E, no horizontal gap yet */
        int Hleft = O + (int)(i + 1) * E;
        int Hdiag = Hrow[0];
        int Fg = NEG_INF;
        const char c1 = s1[i];
        for (int64_t j = 0; j < n; j++) {
            const int Hup = Hrow[j + 1];
            const int Eg = imax(Erow[j] + E, Hup + OE);
            Fg = imax(Fg + E, Hleft + OE);
            const int h = Hdiag + (s0[j] == c1 ? M : X);
            int H = imax(h, imax(Eg, Fg));
            const int back = (H == h) ? 1 : (H == Fg ? 0 : 2);
            /* fback/eback: 1 = extended, 2 = opened, -2 = opened and the NEXT cell's
             * extend/open candidates tie (psa/psa.c:301-327) */
            const int fopen = (Fg - Hleft == OE);
            const int ftie = (Fg + E == H + OE);
            const int eopen = (Eg - Hup == OE);
            const int etie = (Eg + E == H + OE);
            const int8_t fb = (int8_t)(fopen ? (ftie ? -2 : 2) : 1);
            const int8_t eb = (int8_t)(eopen ? (etie ? -2 : 2) : 1);
            if (nib) {
                const int cF = (j >= 1) && (fb == 1 || fbrow[j - 1] == -2);
                const int cE = (i >= 1) && (eb == 1 || ebprev[j] == -2);
                nib[(size_t)i * n + j] = (uint8_t)(back | (cF << 2) | (cE << 3));
            }
            if (b3) {
                b3[(size_t)i * n + j] = (int8_t)back;
                b3[(size_t)m * n + (size_t)i * n + j] = fb;
                b3[2 * (size_t)m * n + (size_t)i * n + j] = eb;
            }
            fbrow[j] = fb;
            ebprev[j] = eb;
            if (H > ms) ms = H;
            Hdiag = Hup;
            Hrow[j + 1] = H;
            Erow[j] = Eg;
            Hleft = H;
        }
        Hrow[0] = O + (int)(i + 1) * E;
    }
    res->maxsorce = ms;
    res->corner = (n > 0 && m > 0) ? Hrow[n] : 0;
    uint64_t h = 1469598103934665603ULL;
    for (int64_t j = 0; j < n; j++) {
        h = fnv1a_i32(h, Hrow[j + 1]);
        if (lastrow) lastrow[j] = Hrow[j + 1];
    }
    res->lastrow_fnv = h;
    res->alen = 0;
    free(Hrow); free(Erow); free(fbrow); free(ebprev);
}

/*
 * Score-only form of psa_fill(): the same recurrences (psa/psa.c:219-363) without the direction
 * bookkeeping, so that the full BASELINE sizes (1e10 cells) finish in well under a minute.
 * tests/test_oracle_psa.py requires it to agree with psa_fill() on maxsorce, corner and the
 * whole last row.
 */
static void psa_fill_scores(const char* s0, int64_t n, const char* s1, int64_t m,
                            int M, int X, int E, int O, int32_t* lastrow, tsta_oracle_result* res)
{
    const int OE = O + E;
    int32_t* Hrow = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t)); /* H(i-1, j), index j+1 */
    int32_t* Erow = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t)); /* Eg(i-1, j)           */
    int ms = MS_SEED;
    Hrow[0] = 0;
    for (int64_t j = 0; j < n; j++) {
        Hrow[j + 1] = O + (int)(j + 1) * E;
        Erow[j] = NEG_INF;
    }
    int32_t* T = (int32_t*)malloc((size_t)(n + 1) * sizeof(int32_t));    /* max(h, Eg) of the current row */
    for (int64_t i = 0; i < m; i++) {
        const char c1 = s1[i];
        /* pass 1 (no dependency along the row): Eg and the diagonal candidate from row i-1 */
        for (int64_t j = 0; j < n; j++) {
            const int Eg = imax(Erow[j] + E, Hrow[j + 1] + OE);
            const int h = Hrow[j] + (s0[j] == c1 ? M : X);
            Erow[j] = Eg;
            T[j] = imax(h, Eg);
        }
        /* pass 2: the horizontal gap chain */
        /* Fg(i,j+1) = max(Fg(i,j)+E, H(i,j)+OE) with H = max(T, Fg) and O <= 0 is max(Fg(i,j)+E, T(j)+OE):
         * two operations on the serial chain instead of five */
        int Fg = O + (int)(i + 1) * E + OE;                      /* Fg(i,0) = H(i,-1)+O+E */
        for (int64_t j = 0; j < n; j++) {
            const int H = imax(T[j], Fg);
            Fg = imax(Fg + E, T[j] + OE);
            ms = imax(ms, H);
            Hrow[j + 1] = H;
        }
        Hrow[0] = O + (int)(i + 1) * E;
    }
    free(T);
    res->maxsorce = ms;
    res->corner = (n > 0 && m > 0) ? Hrow[n] : 0;
    uint64_t h = 1469598103934665603ULL;
    for (int64_t j = 0; j < n; j++) {
        h = fnv1a_i32(h, Hrow[j + 1]);
        if (lastrow) lastrow[j] = Hrow[j + 1];
    }
    res->lastrow_fnv = h;
    res->alen = 0;
    free(Hrow); free(Erow);
}

/* Score-only entry point (maxsorce, corner, last row); see psa_fill_scores(). */
int tsta_oracle_psa_scores(const char* s0, int64_t n, const char* s1, int64_t m,
                           int M, int X, int E, int O, tsta_oracle_result* res, int32_t* lastrow)
{
    psa_fill_scores(s0, n, s1, m, M, X, E, O, lastrow, res);
    return 0;
}

/* Turn a reversed op list (0 left, 1 diag, 2 up) into the two aligned strings. */
static void emit_strings(const char* s0, int64_t n, const char* s1, int64_t m,
                         const uint8_t* ops, int64_t nops, char** a_out, char** b_out)
{
    char* a = (char*)malloc((size_t)nops + 1);
    char* b = (char*)malloc((size_t)nops + 1);
    int64_t i = m - 1, j = n - 1, p = nops;
    for (int64_t k = 0; k < nops; k++) {
        p--;
        if (ops[k] == 1) { a[p] = s0[j--]; b[p] = s1[i--]; }
        else if (ops[k] == 0) { a[p] = s0[j--]; b[p] = '-'; }
        else { a[p] = '-'; b[p] = s1[i--]; }
    }
    a[nops] = b[nops] = '\0';
    *a_out = a; *b_out = b;
}

/*
 * Packed (4-bit code) oracle.  a/b may be NULL for score-only.  Caller frees *a,*b
 * with tsta_oracle_free.  dirs_out (may be NULL) receives the m*n code bytes.
 */
int tsta_oracle_psa(const char* s0, int64_t n, const char* s1, int64_t m,
                    int M, int X, int E, int O, int want_trace,
                    tsta_oracle_result* res, char** a, char** b,
                    int32_t* lastrow, uint8_t* dirs_out)
{
    uint8_t* nib = NULL;
    if (want_trace || dirs_out) {
        nib = dirs_out ? dirs_out : (uint8_t*)malloc((n > 0 && m > 0) ? (size_t)n * (size_t)m : 1);
        if (!nib) return -1;
    }
    psa_fill(s0, n, s1, m, M, X, E, O, nib, NULL, lastrow, res);
    if (want_trace) {
        /* 3-state walker equivalent to psa/psa.c:441-463: `forced` models the cells the
         * reference overwrites ahead of itself (back[i][j-1]=0 / back[i-1][j]=2). */
        uint8_t* ops = (uint8_t*)malloc((size_t)(n + m) + 1);
        int64_t nops = 0, i = m - 1, j = n - 1;
        int forced = -1;
        while (i >= 0 && j >= 0) {
            const uint8_t c = nib[(size_t)i * n + j];
            const int d = forced >= 0 ? forced : (c & 3);
            if (d == 1) { ops[nops++] = 1; i--; j--; forced = -1; }
            else if (d == 0) { ops[nops++] = 0; forced = (c & 4) ? 0 : -1; j--; }
            else { ops[nops++] = 2; forced = (c & 8) ? 2 : -1; i--; }
        }
        while (j >= 0) { ops[nops++] = 0; j--; }   /* psa/psa.c:497-503 */
        while (i >= 0) { ops[nops++] = 2; i--; }   /* psa/psa.c:504-510 */
        emit_strings(s0, n, s1, m, ops, nops, a, b);
        res->alen = nops;
        free(ops);
    }
    if (nib && !dirs_out) free(nib);
    return 0;
}

/*
 * Literal three-matrix oracle (3 bytes per cell), walking exactly like trace():
 * first walk rewrites `back` ahead of itself, second walk emits (psa/psa.c:441-510).
 */
int tsta_oracle_psa_ref3(const char* s0, int64_t n, const char* s1, int64_t m,
                         int M, int X, int E, int O,
                         tsta_oracle_result* res, char** a_out, char** b_out)
{
    size_t cells = (size_t)n * m;
    int8_t* b3 = (int8_t*)malloc(cells ? 3 * cells : 1);
    if (!b3) return -1;
    psa_fill(s0, n, s1, m, M, X, E, O, NULL, b3, NULL, res);
    int8_t* back = b3; int8_t* fback = b3 + cells; int8_t* eback = b3 + 2 * cells;
#define AT(mat, i, j) mat[(size_t)(i) * n + (j)]
    int64_t i = m - 1, j = n - 1, cnt = 0;
    while (i >= 0 && j >= 0) {
        if (AT(back, i, j) == 1) { i--; j--; }
        else if (AT(back, i, j) == 0) {
            int8_t f = AT(fback, i, j);
            if (j - 1 >= 0 && ((f == 1 || f == -1) || ((f == 2 || f == -2) && AT(fback, i, j - 1) < 0)))
                AT(back, i, j - 1) = 0;
            j--;
        } else {
            int8_t e = AT(eback, i, j);
            if (i - 1 >= 0 && ((e == 1 || e == -1) || ((e == 2 || e == -2) && AT(eback, i - 1, j) < 0)))
                AT(back, i - 1, j) = 2;
            i--; cnt++;
        }
    }
    if (i >= 0) cnt += i + 1;
    cnt += n;
    char* a = (char*)malloc((size_t)cnt + 1);
    char* b = (char*)malloc((size_t)cnt + 1);
    a[cnt] = b[cnt] = '\0';
    int64_t p = cnt;
    i = m - 1; j = n - 1;
    while (i >= 0 && j >= 0) {
        if (AT(back, i, j) == 1) { a[p - 1] = s0[j]; b[p - 1] = s1[i]; i--; j--; }
        else if (AT(back, i, j) == 0) { a[p - 1] = s0[j]; b[p - 1] = '-'; j--; }
        else { a[p - 1] = '-'; b[p - 1] = s1[i]; i--; }
        p--;
    }
    while (j >= 0) { a[p - 1] = s0[j]; b[p - 1] = '-'; j--; p--; }
    while (i >= 0) { b[p - 1] = s1[i]; a[p - 1] = '-'; i--; p--; }
#undef AT
    res->alen = cnt;
    *a_out = a; *b_out = b;
    free(b3);
    return 0;
}

void tsta_oracle_free(void* p) { free(p); }
