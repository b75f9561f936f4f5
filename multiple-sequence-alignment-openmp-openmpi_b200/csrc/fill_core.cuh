// fill_core.cuh — the per-lane R x 4 block update shared by the fill kernels (fill.cu) and the
// tile-recompute traceback (traceback_tiles.cu).  See fill.cu for the decomposition.
#pragma once
#include "seqalk_kernels.h"

namespace seqalk {

// a + b, optionally forced onto the FMA pipe as a*one + b (one == 1 at run time, unknown
// to ptxas) so that the ALU pipe is left to VIMNMX3 / SHF / ISETP / PRMT.
template <int ADDMODE>
__device__ __forceinline__ int add_i(int a, int b, int one) {
  if (ADDMODE != 1) return a + b;
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
  return r;
}
template <int ADDMODE>
__device__ __forceinline__ int sub_i(int a, int b, int minus_one) {  // a - b
  if (ADDMODE != 1) return a - b;
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(minus_one), "r"(a));
  return r;
}

// One lane's R x 4 block.  h[r] = H' of row r at the previous column; dg = H' of the row
// above the block at the previous column; up_in[c] = H' of the row above at column c.
//
// Per cell, fast modes (pxy >= 0, pgap >= 0, scores far from int32 overflow):
//     t  = diag + cost(x, y)                           MODE_CMP: ISETP + add + predicated add
//                                                      MODE_TAB: PRMT (4-letter table) + add
//     v  = min3(up, left, t)                           VIMNMX3
//     d1 = v - t      (< 0  <=>  diagonal not optimal) add
//     d2 = left - up  (< 0  <=>  left strictly better) add
//     bits = (bits << 1 | sign(d1)) << 1 | sign(d2)    2 funnel shifts (SHF.L.W)
// On a match the reference takes the diagonal unconditionally (:478-479); with
// non-negative penalties H[i-1][j-1] <= min(H[i-1][j], H[i][j-1]) + pgap, so the
// plain minimum selects the same value and sign(d1) = 0 (DIAG) — identical output.
// MODE_GENERAL (any integers): explicit select on match, compares instead of sign tricks.
//
// xc[r]: MODE_TAB: 4 signed bytes, byte q = cost of row r's character against letter q;
//        otherwise the raw character code (0x100+r for rows past m: never matches).
// EMIT = false: scores only (no direction bits): the checkpointed fill of the tile mode.
template <int R, int MODE, int ADDMODE, bool EMIT>
__device__ __forceinline__ void macro_step(int (&h)[R], int &dg, const int (&xc)[R], const int (&up_in)[4],
                                           unsigned yword, int A, int B, int one, int minus_one,
                                           int (&outv)[4], unsigned (&bits)[R / 4]) {
#pragma unroll
  for (int w = 0; w < R / 4; ++w) bits[w] = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int yc = 0;
    unsigned sel = 0;
    if (MODE == MODE_TAB && ADDMODE == 2) {
      // y letters arrive as shift amounts 8*code: one funnel shift (amount mod 32) builds the
      // one-hot byte mask; dp4a(table, one-hot, diag) = diag + table[code]
      sel = __funnelshift_l(0u, 1u, yword >> (8 * c));
    } else {
      yc = (int)((yword >> (8 * c)) & 0xffu);
      if (MODE == MODE_TAB) sel = (unsigned)(yc >> 3) * 0x1111u + 0x8880u;  // PRMT: byte code, sign-replicated
    }
    int diag = dg;
    int up = up_in[c];
    dg = up;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int w = (c * R + r) / CELLS_PER_WORD;  // cells are shifted in, first cell ends up highest
      const int left = h[r];
      int v;
      if (MODE == MODE_GENERAL) {
        const bool match = (xc[r] == yc);
        const int t = diag + B;
        const int m1 = min(up, left);
        v = match ? diag + A : min(t, m1);
        if (EMIT) {
          const unsigned nd = ((!match) && (t > m1)) ? 2u : 0u;
          const unsigned lf = (up > left) ? 1u : 0u;
          bits[w] = (bits[w] << 2) | nd | lf;
        }
      } else {
        int t;
        if (MODE == MODE_TAB && ADDMODE == 2) {
          // cost lookup and add in ONE instruction (IDP.4A): sum of table bytes x one-hot bytes
          t = __dp4a(xc[r], (int)sel, diag);
        } else if (MODE == MODE_TAB) {
          int cost;
          asm("prmt.b32 %0, %1, %2, %3;" : "=r"(cost) : "r"(xc[r]), "r"(0), "r"(sel));
          t = add_i<ADDMODE>(diag, cost, one);
        } else {
          t = add_i<ADDMODE>(diag, B, one);
          if (ADDMODE == 0)
            asm("{.reg .pred p; setp.eq.s32 p, %1, %2; @p add.s32 %0, %3, %4;}"
                : "+r"(t)
                : "r"(xc[r]), "r"(yc), "r"(diag), "r"(A));
          else
            asm("{.reg .pred p; setp.eq.s32 p, %1, %2; @p mad.lo.s32 %0, %3, %5, %4;}"
                : "+r"(t)
                : "r"(xc[r]), "r"(yc), "r"(diag), "r"(A), "r"(one));
        }
        v = min(min(up, left), t);
        if (EMIT) {
          const int d1 = sub_i<ADDMODE>(v, t, minus_one);
          const int d2 = sub_i<ADDMODE>(left, up, minus_one);
          bits[w] = __funnelshift_l((unsigned)d1, bits[w], 1);
          bits[w] = __funnelshift_l((unsigned)d2, bits[w], 1);
        }
      }
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
  }
}


// Row operand of macro_step for one DP row: the 4-letter cost table (MODE_TAB, `ch` is 8 x the
// letter code, i.e. 0/8/16/24) or the raw character.
// Rows past the end of x: MODE_TAB gives them an all-zero table (cost 0 against every letter).
// H' never increases down a column or along a row, so with cost 0 such a row takes
// min3(diag, up, left) = up: every padded row is a copy of DP row m, and the strip's bottom row —
// which the fill stores anyway — carries H'[m][1..n].  The corner H'[m][n] read from there is the
// fill's own score (score-only entry points, and the on-device check of the traceback's path
// cost).  Other modes: a character that never matches (their padded rows hold no usable value).
template <int MODE>
__device__ __forceinline__ int row_operand(bool inside, int ch, int r, int A, int B) {
  if (MODE == MODE_TAB) {
    if (!inside) return 0;
    const unsigned Ab = (unsigned)A & 0xffu, Bb = (unsigned)B & 0xffu;
    return (int)((Bb * 0x01010101u) ^ ((Ab ^ Bb) << (ch & 24)));  // ch = 8 * letter code
  }
  return inside ? ch : (0x100 + r);
}

}  // namespace seqalk
