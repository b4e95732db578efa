// sw_core.cuh -- one ksw_align2 job per lane: the reference's striped Smith-Waterman kernels (ksw_u8 src/ksw.c:440-572, ksw_i16
// src/ksw.c:574-696, SSE2 build) replayed in scalar integer code over query positions.  Shared between the CUDA kernel (sw.cu) and a
// host build (tests/host_emul/sw_host.cpp) that is checked against the oracle on the CPU box.
//
// What has to be replayed rather than recomputed (the results are not those of the textbook recurrence):
//   main pass   positions of a lane (l * slen + j, j = 0 .. slen-1; slen = ceil(qlen / p), p = 16 or 8 lanes) in order, F restarted at
//               0 at the head of every lane; E(i+1, .) from this pass's H; the row maximum is taken here only;
//   lazy F      the F left at the tail of each lane moves one lane up and decays through it, raising H only; vector-wide loop of up
//               to 16 rounds x slen steps, left after the first step at which no lane's F exceeds H - oe_ins;
//   saturation  unsigned bytes over a matrix biased by `shift` (u8), signed 16-bit adds / unsigned saturating subtractions (i16);
//   bookkeeping rows at or above the XSUBO threshold collected in runs (here: the row maxima are kept and the runs rebuilt at the end),
//               Hmax at the best row, XSTOP / byte-overflow early stop, qe = smallest position among equal maxima, score2 / te2.
// State per lane, element stride NS (NS = number of lanes of the grid: coalesced rows): H0, H1, E, Hmax of n = slen * p int16 each,
// and one int16 per target row for the row maxima.
#pragma once
#include <stdint.h>
#include "bwamem_b200.h"

#ifdef __CUDACC__
#define SW_DEV __device__ __forceinline__
#define SW_MEM __device__ __forceinline__
#else
#define SW_DEV static inline
#define SW_MEM inline
#endif

struct SwParams { int8_t mat[32]; int32_t m, o_del, e_del, o_ins, e_ins; };

struct SwSeq {                  // a sequence as the second pass of ksw_align2 sees it: the first `rev` bases reversed, the rest as given
    const uint8_t *p;
    int rev;
    SW_MEM int operator()(int i) const { const int c = p[i < rev ? rev - 1 - i : i]; return c > 4 ? 4 : c; }
};

SW_DEV int sw_subs(int a, int b) { return a > b ? a - b : 0; }
SW_DEV int sw_max(int a, int b) { return a > b ? a : b; }

template <int P>                // P = 16: ksw_u8, P = 8: ksw_i16
SW_DEV void sw_striped(int qlen, const SwSeq &Q, int tlen, const SwSeq &T, const SwParams &S, int xtra,
                       int16_t *ws, size_t NS, size_t n_cap, int16_t *rowmax, bwa_b200_sw_result_t &r)
{
    r.score = 0; r.te = -1; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    const int slen = (qlen + P - 1) / P, n = slen * P;
    if (slen == 0) return;
    const int oe_del = S.o_del + S.e_del, oe_ins = S.o_ins + S.e_ins;
    int shift = 127, qmax = 0;
    for (int a = 0; a < S.m * S.m; ++a) { shift = S.mat[a] < shift ? S.mat[a] : shift; qmax = S.mat[a] > qmax ? S.mat[a] : qmax; }
    shift = (256 - shift) & 255;
    const int minsc = (xtra & 0x40000) ? xtra & 0xffff : 0x10000, endsc = (xtra & 0x20000) ? xtra & 0xffff : 0x10000;
    int16_t *H0 = ws, *H1 = ws + n_cap * NS, *E = ws + 2 * n_cap * NS, *Hmax = ws + 3 * n_cap * NS;
    for (int pos = 0; pos < n; ++pos) { H0[pos * NS] = 0; E[pos * NS] = 0; Hmax[pos * NS] = 0; }
    int gmax = 0, te = -1, n_rows = 0;
    int fv[P];
    for (int i = 0; i < tlen; ++i) {
        const int8_t *ma = S.mat + T(i) * S.m;
        int rowm = 0, hd = 0;                                                // hd = H(i-1, pos-1)
        for (int l = 0; l < P; ++l) {
            int f = 0;
            for (int j = 0; j < slen; ++j) {
                const int pos = l * slen + j;
                const int sc = pos >= qlen ? 0 : ma[Q(pos)];
                int h;
                if (P == 16) { h = hd + sc + shift; h = h > 255 ? 255 : h; h = sw_subs(h, shift); }
                else { h = hd + sc; h = h > 32767 ? 32767 : (h < -32768 ? -32768 : h); }
                hd = H0[pos * NS];
                int e = E[pos * NS];
                h = sw_max(h, e); h = sw_max(h, f);
                rowm = sw_max(rowm, h);
                H1[pos * NS] = (int16_t)h;
                E[pos * NS] = (int16_t)sw_max(sw_subs(e, S.e_del), sw_subs(h, oe_del));
                f = sw_max(sw_subs(f, S.e_ins), sw_subs(h, oe_ins));
            }
            fv[l] = f;
        }
        bool done = false;
        for (int k = 0; k < 16 && !done; ++k) {
#pragma unroll
            for (int l = P - 1; l > 0; --l) fv[l] = fv[l - 1];
            fv[0] = 0;
            for (int j = 0; j < slen; ++j) {
                bool any = false;
#pragma unroll
                for (int l = 0; l < P; ++l) {
                    const size_t at = (size_t)(l * slen + j) * NS;
                    int h = H1[at];
                    if (fv[l] > h) { h = fv[l]; H1[at] = (int16_t)h; }
                    h = sw_subs(h, oe_ins);
                    fv[l] = sw_subs(fv[l], S.e_ins);
                    any = any || fv[l] > h;
                }
                if (!any) { done = true; break; }
            }
        }
        rowmax[(size_t)i * NS] = (int16_t)rowm;
        n_rows = i + 1;
        if (rowm > gmax) {
            gmax = rowm; te = i;
            for (int pos = 0; pos < n; ++pos) Hmax[pos * NS] = H1[pos * NS];
            if ((P == 16 && gmax + shift >= 255) || gmax >= endsc) break;
        }
        int16_t *t = H1; H1 = H0; H0 = t;
    }
    r.score = P == 16 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (P == 16 && r.score == 255) return;
    int mx = -1;
    for (int pos = 0; pos < n; ++pos) {                                      // smallest position among the maxima
        const int v = Hmax[pos * NS];
        if (v > mx) { mx = v; r.qe = pos; }
    }
    // the reference's list b (src/ksw.c:526-537): a row at or above minsc extends the last entry only when that entry's RECORDED row
    // is the row before (the recorded row moves only when the score improves), otherwise it opens a new entry; score2 / te2 = the best
    // entry outside the window around te (src/ksw.c:559-568).  Rebuilt here from the row maxima.
    if (minsc <= 0xffff) {
        const int w = (r.score + qmax - 1) / qmax, low = te - w, high = te + w;
        int last_sc = -1, last_e = -1;
        for (int i = 0; i < n_rows; ++i) {
            const int v = (int)rowmax[(size_t)i * NS];
            if (v < minsc) continue;
            if (last_e < 0 || last_e + 1 != i) {
                if (last_e >= 0 && (last_e < low || last_e > high) && last_sc > r.score2) { r.score2 = last_sc; r.te2 = last_e; }
                last_sc = v; last_e = i;
            } else if (last_sc < v) { last_sc = v; last_e = i; }
        }
        if (last_e >= 0 && (last_e < low || last_e > high) && last_sc > r.score2) { r.score2 = last_sc; r.te2 = last_e; }
    }
}

// The score ksw_i16 returns (src/ksw.c:574-696) when nothing else of the result is read and no threshold is set (xtra = KSW_XSTART:
// mem_seed_sw, src/bwamem.c:774-808): the same replay as sw_striped<8> without Hmax, the row maxima and the second pass.  One H array is
// enough: in position order a cell's previous-row value is read right before it is overwritten and the lazy-F rounds only touch the new
// row.  H and E hold ceil(qlen / 8) * 8 int16 each at element stride NS; Q(pos) / T(row) return codes 0..4.
template <class QS, class TS>
SW_DEV int sw_i16_score(int qlen, const QS &Q, int tlen, const TS &T, const SwParams &S, int16_t *H, int16_t *E, size_t NS)
{
    constexpr int P = 8;
    const int slen = (qlen + P - 1) / P, n = slen * P;
    if (slen == 0) return 0;
    const int oe_del = S.o_del + S.e_del, oe_ins = S.o_ins + S.e_ins;
    for (int pos = 0; pos < n; ++pos) { H[pos * NS] = 0; E[pos * NS] = 0; }
    int gmax = 0;
    int fv[P];
    for (int i = 0; i < tlen; ++i) {
        const int8_t *ma = S.mat + T(i) * S.m;
        int rowm = 0, hd = 0;
        for (int l = 0; l < P; ++l) {
            int f = 0;
            for (int j = 0; j < slen; ++j) {
                const int pos = l * slen + j;
                int h = hd + (pos >= qlen ? 0 : ma[Q(pos)]);
                h = h > 32767 ? 32767 : (h < -32768 ? -32768 : h);
                hd = H[pos * NS];
                const int e = E[pos * NS];
                h = sw_max(h, e); h = sw_max(h, f);
                rowm = sw_max(rowm, h);
                H[pos * NS] = (int16_t)h;
                E[pos * NS] = (int16_t)sw_max(sw_subs(e, S.e_del), sw_subs(h, oe_del));
                f = sw_max(sw_subs(f, S.e_ins), sw_subs(h, oe_ins));
            }
            fv[l] = f;
        }
        bool done = false;
        for (int k = 0; k < 16 && !done; ++k) {
#pragma unroll
            for (int l = P - 1; l > 0; --l) fv[l] = fv[l - 1];
            fv[0] = 0;
            for (int j = 0; j < slen; ++j) {
                bool any = false;
#pragma unroll
                for (int l = 0; l < P; ++l) {
                    const size_t at = (size_t)(l * slen + j) * NS;
                    int h = H[at];
                    if (fv[l] > h) { h = fv[l]; H[at] = (int16_t)h; }
                    h = sw_subs(h, oe_ins);
                    fv[l] = sw_subs(fv[l], S.e_ins);
                    any = any || fv[l] > h;
                }
                if (!any) { done = true; break; }
            }
        }
        gmax = sw_max(gmax, rowm);
    }
    return gmax;
}

// ksw_align2 (src/ksw.c:698-736), qry = NULL, avx2 = 0
SW_DEV void sw_align2(int qlen, const uint8_t *q, int tlen, const uint8_t *t, const SwParams &S, int xtra,
                      int16_t *ws, size_t NS, size_t n_cap, int16_t *rowmax, bwa_b200_sw_result_t &r)
{
    const bool byte_mode = (xtra & 0x10000) != 0;
    SwSeq Q{q, 0}, T{t, 0};
    if (byte_mode) sw_striped<16>(qlen, Q, tlen, T, S, xtra, ws, NS, n_cap, rowmax, r);
    else sw_striped<8>(qlen, Q, tlen, T, S, xtra, ws, NS, n_cap, rowmax, r);
    if ((xtra & 0x80000) == 0 || ((xtra & 0x40000) && r.score < (xtra & 0xffff))) return;
    if (r.qe < 0 || r.te < 0) return;          // byte overflow: the reference's second pass is undefined there
    bwa_b200_sw_result_t rr;
    SwSeq Q2{q, r.qe + 1}, T2{t, r.te + 1};
    if (byte_mode) sw_striped<16>(r.qe + 1, Q2, tlen, T2, S, 0x20000 | r.score, ws, NS, n_cap, rowmax, rr);
    else sw_striped<8>(r.qe + 1, Q2, tlen, T2, S, 0x20000 | r.score, ws, NS, n_cap, rowmax, rr);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
}
