// fill16.cu — the score-only fill of the tile mode with TWO DP cells per register (16x2 SIMD),
// hand-written for sm_100a.  Same role and same results as fill_kernel<.., EMIT = false> in
// fill.cu (reference: the DP recurrence of getMinimumPenalty2, submit/xuliny-seqalkway.cpp:462-491,
// sequential twin :594-609); used when the input has <= 4 letters and the gap penalty is small
// enough for 16-bit relative scores (fill16_usable).
//
// Why: the 32-bit fill needs IDP.4A + VIMNMX3 per cell, each reading three different registers; the
// SM sub-partition's register file delivers ~2 operands per clock, which caps that instruction mix
// at IPC ~0.63 (profiles/experiments/issue_rf.cu: 11.2 TCUPS for the bare 32-bit loop, 14.3-15.2
// for the packed one).  Packing halves the register reads per cell.
//
// Decomposition: a "strip pair" = two consecutive strips A, B of 32*RP rows of one pair, swept by
// ONE warp as a 64-lane virtual warp: virtual lane t (rows of A) lives in the low halves of lane
// t's registers, virtual lane 32+t (rows of B) in the high halves.  At macro step q physical lane t
// works on macro column q-t in the low halves and q-t-32 in the high halves.  Lane 31's low half
// (bottom row of A) feeds lane 0's high half (top row of B) one step later, exactly like every
// other lane-to-lane hand-off, so a strip pair is one skewed wavefront of 64 x RP rows.
//
// Scores are kept in the max domain G = -H' (H' = H - (i+j)*pgap as in fill.cu), relative to a
// warp-uniform base:  G[i][j] = max3(G[i-1][j-1] + cost, G[i-1][j], G[i][j-1]),
// cost = 2*pgap on a match, max(0, 2*pgap - pxy) otherwise (a negative mismatch gain can never win:
// G[i-1][j-1] <= G[i-1][j]).  Per register (two cells):
//     PRMT        cost pair  = bytes [tabA[r][yA], 0, tabB[r][yB], 0]     (ALU pipe)
//     IMAD        t = diag + cost pair   (plain 32-bit add: halves never carry) (FMA pipe)
//     VIMNMX3.U16x2   v = max3(up, left, t)                               (ALU pipe)
// A half that has not started or is past its last column runs with a null selector (cost 0): it
// keeps the boundary value 0 resp. copies its last column (G is monotone), so every lane executes
// every step unconditionally — no per-lane activity branch.  Rows past m have an all-zero table and
// copy row m (see row_operand in fill_core.cuh): the pair's last kept row carries H'[m][.].
//
// 16-bit range: at any step all values of a strip pair lie within 2*pgap*(256 + 64*RP) of each
// other; every K = 2^rebase_shift steps the warp subtracts the smallest value it holds from all of
// them (base += delta) and logs the base, so relative scores stay below 65536.
//
// What goes to HBM (same places as the 32-bit fill, so traceback_tiles.cu differs only in how it
// reads a tile's left edge):
//   * kept rows (every TILE_ROWS DP rows): absolute int32 H', converted with IDP.2A — also the
//     hand-off row to the next strip pair and the carrier of the corner H'[m][n];
//   * column checkpoints (every CKQ macro columns): the packed registers as they are, images of
//     mpad/2 words.  CKQ = 64: per checkpoint two images [low halves valid | high halves valid],
//     since the two halves reach a checkpoint column 32 steps apart; CKQ = 32: both halves are at a
//     checkpoint column at the same step, image e = [rows of A at checkpoint e | rows of B at
//     checkpoint e - 1].  The traceback unpacks with the logged base.
//
// Builds of the one-launch kernel (FILL16_* flags, chosen per job shape by engine.cu):
// column checkpoints by predicated stores (PREDCK) instead of a one-lane branch; for jobs of a few long
// pairs the lane hand-off shuffled as soon as a column is done (PIPE) and sentinel words in the
// hand-off rows instead of progress counters (SENT).
#include <algorithm>

#include "fill_core.cuh"

namespace seqalk {

namespace {

constexpr int F16_THREADS = 128;
static_assert(CKQ == 64 || CKQ == 32, "the two halves of a register reach a macro column 32 steps apart");
constexpr unsigned SEL_NULL = 0xCC88CC88u;  // both columns of a selector word: cost 0 in both halves

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// c - low / high 16 bits of a (IDP.2A: a = two unsigned halves, b = signed bytes -1 / 0)
__device__ __forceinline__ int sub_lo16(int c, unsigned a) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0x000000ffu), "r"(c));
  return d;
}
__device__ __forceinline__ int sub_hi16(int c, unsigned a) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0x0000ff00u), "r"(c));
  return d;
}

// One lane's RP x 4 block of packed cells (column-major).  sel01 / sel23: the selector words of
// the macro column (16 bits per DP column).
template <int RP>
__device__ __forceinline__ void macro_step16(unsigned (&h)[RP], unsigned &dg, const unsigned (&tl)[RP],
                                             const unsigned (&th)[RP], const unsigned (&up_in)[4], unsigned sel01,
                                             unsigned sel23, unsigned one, unsigned (&outv)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned w = c < 2 ? sel01 : sel23;
    const unsigned sel = (c & 1) ? (w >> 16) : w;  // PRMT reads the low 16 bits of its selector
    unsigned diag = dg, up = up_in[c];
    dg = up;
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const unsigned left = h[r];
      const unsigned cost = prmt(tl[r], th[r], sel);
      unsigned t;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(diag), "r"(one), "r"(cost));
      const unsigned v = __vimax3_u16x2(left, up, t);  // left first: same operand slot as the next row's diag
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
  }
}

// macro_step16 that also hands the block's bottom row to the next lane as soon as a column is done
// (PIPE, jobs of a few long pairs: a warp alone on its scheduler would otherwise wait out the shuffle
// latency at the top of every step; only the last column's shuffle is still on the critical path).
template <int RP>
__device__ __forceinline__ void macro_step16_sh(unsigned (&h)[RP], unsigned &dg, const unsigned (&tl)[RP],
                                                const unsigned (&th)[RP], const unsigned (&up_in)[4], unsigned sel01,
                                                unsigned sel23, unsigned one, unsigned (&outv)[4], unsigned (&sh)[4],
                                                int src_lane) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned w = c < 2 ? sel01 : sel23;
    const unsigned sel = (c & 1) ? (w >> 16) : w;
    unsigned diag = dg, up = up_in[c];
    dg = up;
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      const unsigned left = h[r];
      const unsigned cost = prmt(tl[r], th[r], sel);
      unsigned t;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(diag), "r"(one), "r"(cost));
      const unsigned v = __vimax3_u16x2(left, up, t);
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
    sh[c] = __shfl_sync(0xffffffffu, up, src_lane);
  }
}

// Cost table of one DP row in the max domain: byte q = gain of the row's letter against letter q.
__device__ __forceinline__ unsigned row_table16(bool inside, unsigned ch, unsigned cm, unsigned cx) {
  return inside ? ((cx * 0x01010101u) ^ ((cm ^ cx) << (ch & 24u))) : 0u;  // ch = 8 * letter code
}

// PREDCK: the column checkpoints are written by predicated stores through a per-lane running pointer
// instead of a branch that one lane takes at every step (see below).
// PIPE: the lane hand-off is shuffled as soon as a column is done (macro_step16_sh) and the staged
// entries of the next step are read from shared memory one step ahead.
// SENT: no progress counters — the predecessor's bottom row itself says whether it is there: the engine
// fills the hand-off rows with FILL16_SENTINEL words before the launch, a block of 32 macro columns is
// staged once none of its words is the sentinel any more (one L2 round trip per attempt instead of
// poll + acquire + load, and no release fence on the producer's side).  H' can never take that value.
template <int RP, bool PREDCK = false, bool PIPE = false, bool SENT = false>
__device__ __forceinline__ void process_strip_pair(const FillArgs &a, const PairDesc &P, const int s,
                                                   const int row_base, const int lane, uint2 *ring_sel,
                                                   int4 *ring_top) {
  const unsigned FULL = 0xffffffffu;
  const int m = P.m, n = P.n;
  const int nq = (n + 3) >> 2;
  const int nsteps = nq + 63;
  const bool first = (s == 0), last = (s == P.nstrips - 1);
  const int KS = a.rebase_shift;
  const unsigned one = (unsigned)a.one;
  const unsigned cm = (unsigned)(2 * a.pgap), cx = (unsigned)max(0, 2 * a.pgap - a.pxy);

  // ---- this lane's rows: RP of strip A (low halves), RP of strip B (high halves)
  const int rowA0 = row_base + lane * RP, rowB0 = rowA0 + 32 * RP;
  unsigned tl[RP], th[RP];
  {
    const uint8_t *__restrict__ xs = a.codes + P.x_off;  // zero padded far beyond m (engine.cu)
    unsigned wa[RP / 4], wb[RP / 4];
    static_assert(RP == 32 || RP == 16 || RP == 8 || RP == 4, "rows per half lane");
    if constexpr (RP == 32) {
#pragma unroll
      for (int q4 = 0; q4 < 2; ++q4) {
        const uint4 va = rowA0 + 16 * q4 < m ? __ldg(reinterpret_cast<const uint4 *>(xs + rowA0) + q4) : make_uint4(0, 0, 0, 0);
        const uint4 vb = rowB0 + 16 * q4 < m ? __ldg(reinterpret_cast<const uint4 *>(xs + rowB0) + q4) : make_uint4(0, 0, 0, 0);
        wa[4 * q4] = va.x; wa[4 * q4 + 1] = va.y; wa[4 * q4 + 2] = va.z; wa[4 * q4 + 3] = va.w;
        wb[4 * q4] = vb.x; wb[4 * q4 + 1] = vb.y; wb[4 * q4 + 2] = vb.z; wb[4 * q4 + 3] = vb.w;
      }
    } else if constexpr (RP == 16) {
      const uint4 va = rowA0 < m ? __ldg(reinterpret_cast<const uint4 *>(xs + rowA0)) : make_uint4(0, 0, 0, 0);
      const uint4 vb = rowB0 < m ? __ldg(reinterpret_cast<const uint4 *>(xs + rowB0)) : make_uint4(0, 0, 0, 0);
      wa[0] = va.x; wa[1] = va.y; wa[2] = va.z; wa[3] = va.w;
      wb[0] = vb.x; wb[1] = vb.y; wb[2] = vb.z; wb[3] = vb.w;
    } else if constexpr (RP == 8) {
      const uint2 va = rowA0 < m ? __ldg(reinterpret_cast<const uint2 *>(xs + rowA0)) : make_uint2(0, 0);
      const uint2 vb = rowB0 < m ? __ldg(reinterpret_cast<const uint2 *>(xs + rowB0)) : make_uint2(0, 0);
      wa[0] = va.x; wa[1] = va.y;
      wb[0] = vb.x; wb[1] = vb.y;
    } else {
      wa[0] = rowA0 < m ? __ldg(reinterpret_cast<const unsigned *>(xs + rowA0)) : 0u;
      wb[0] = rowB0 < m ? __ldg(reinterpret_cast<const unsigned *>(xs + rowB0)) : 0u;
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) {
      tl[r] = row_table16(rowA0 + r < m, (wa[r >> 2] >> (8 * (r & 3))) & 0xffu, cm, cx);
      th[r] = row_table16(rowB0 + r < m, (wb[r >> 2] >> (8 * (r & 3))) & 0xffu, cm, cx);
    }
  }

  // ---- where things go
  const uint64_t bstride = (uint64_t)(4 * nq + 4);
  int32_t *bbase = a.brow + P.brow_off;
  const int4 *__restrict__ brow_rd =
      reinterpret_cast<const int4 *>(bbase + (uint64_t)(first ? 0 : row_base / TILE_ROWS - 1) * bstride);
  const bool keep = ((lane * RP + RP) & (TILE_ROWS - 1)) == 0;  // this lane's bottom rows are kept rows
  int4 *__restrict__ keepA = reinterpret_cast<int4 *>(bbase + (uint64_t)(rowA0 / TILE_ROWS) * bstride);
  int4 *__restrict__ keepB = reinterpret_cast<int4 *>(bbase + (uint64_t)(rowB0 / TILE_ROWS) * bstride);
  // column checkpoints: image (k, half) = mpad/2 packed words; this lane's RP words start at
  // row_base/2 + lane*RP (strip pairs are laid out one after the other)
  const int halfwords = (int)(P.mpad / 2);
  uint32_t *__restrict__ ck = reinterpret_cast<uint32_t *>(a.dirs + P.dir_off) + (uint64_t)(row_base / 2 + lane * RP);
  // base log of this strip pair, behind the pair's kept rows
  const int nbl = base_log_entries(nq, KS);
  int32_t *blog = bbase + (uint64_t)(P.mpad / TILE_ROWS) * bstride + (uint64_t)s * nbl;
  const uint2 *__restrict__ selstream = reinterpret_cast<const uint2 *>(a.sel16 + 2 * (uint64_t)P.y_off);
  const unsigned *prev_progress = a.progress + P.strip0 + (first ? 0 : s - 1);
  unsigned *my_progress = a.progress + P.strip0 + s;

  unsigned h[RP];
#pragma unroll
  for (int r = 0; r < RP; ++r) h[r] = 0u;
  unsigned dg = 0u, outv[4] = {0u, 0u, 0u, 0u};
  unsigned sh[4] = {0u, 0u, 0u, 0u};  // PIPE: outv of the lane above (lane 31's for lane 0), already shuffled
  int negbase = 0;  // -(base): absolute G = relative - negbase, H' = negbase - relative
  const unsigned merge = lane == 0 ? 0x1054u : 0x3210u;  // lane 0: low half <- top row, high half <- lane 31's low half
  const int src_lane = (lane + 31) & 31;
  int avail = first ? nq : 0;
  if (lane == 0) blog[0] = 0;
  const unsigned nq_keep = keep ? (unsigned)nq : 0u;  // store predicate: (unsigned)column < nq_keep
  // The selector ring holds every entry twice, 64 slots apart: entry q - lane is read at slot
  // (q & 63) + 64 - lane without a wrap-around (a lane-constant address plus a uniform offset).
  const uint8_t *sel_rd = reinterpret_cast<const uint8_t *>(ring_sel + 64 - lane);
  const uint8_t *top_rd = reinterpret_cast<const uint8_t *>(ring_top);
  int qcA = -lane;  // macro column of the low halves at step q
  // PREDCK: this lane stores image k - 1 when its low halves have finished macro column 32 * k - 1,
  // k = 1 .. ck_last (the same set of stores as the branch below makes): next such column + 1, and
  // where the image goes
  const int ck_last = ((nq - 1) >> 5) >= 1 ? ((nq - 1) >> 5) + 1 : 0;
  int ck_next = ck_last >= 1 ? 32 : 0x7fffffff;
  uint32_t *ckp = ck;

  __syncwarp();
  ring_sel[32 + lane] = ring_sel[96 + lane] = make_uint2(SEL_NULL, SEL_NULL);  // entries -31..-1: halves not started yet

  const int nblocks = (nsteps + 31) >> 5;
  for (int blk = 0; blk < nblocks; ++blk) {
    const int q0 = blk * 32;
    // ---- re-base (block starts that are multiples of the interval): subtract the smallest value
    // any lane holds from everything the warp holds
    if (q0 > 0 && (q0 & ((1 << KS) - 1)) == 0) {
      unsigned mn = __vminu2(dg, outv[0]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) mn = __vminu2(mn, __shfl_xor_sync(FULL, mn, d));
      const unsigned delta = min(mn & 0xffffu, mn >> 16);
      const unsigned dd = delta * 0x10001u;
#pragma unroll
      for (int r = 0; r < RP; ++r) h[r] -= dd;  // both halves >= delta: no borrow across
#pragma unroll
      for (int c = 0; c < 4; ++c) outv[c] -= dd;
#pragma unroll
      for (int c = 0; c < 4; ++c) sh[c] -= dd;
      dg -= dd;
      negbase -= (int)delta;
      if (lane == 0) blog[q0 >> KS] = -negbase;
    }
    // ---- stage the block's 32 stream entries: selector words of macro column e (low halves) /
    // e - 32 (high halves), and the predecessor's bottom row at macro column e (lane 0's top row)
    const int e = q0 + lane;
    if (!first && !SENT) {
      const int need = min(nq, q0 + 32);
      if (avail < need) {
        // relaxed polls (an acquire load invalidates the SM's L1), one acquire once the value is there
        unsigned ns = 32;
        for (;;) {
          unsigned v = 0;
          if (lane == 0) v = ld_relaxed_u32(prev_progress);
          v = __shfl_sync(FULL, v, 0);
          if ((int)v >= need) break;
          __nanosleep(ns);
          if (ns < 2048) ns *= 2;
        }
        unsigned v = 0;
        if (lane == 0) v = ld_acquire_u32(prev_progress);
        avail = (int)__shfl_sync(FULL, v, 0);
      }
    }
    const uint2 sv = e < nq + 64 ? __ldg(selstream + e) : make_uint2(SEL_NULL, SEL_NULL);
    int4 tv = make_int4(0, 0, 0, 0);
    if (!first) {
      unsigned ns = 32;
      for (;;) {
        if (e < nq) {
          tv = __ldcg(brow_rd + e);
        } else {  // past the last column: keep copying it (the halves keep running with cost 0)
          const int lastv = __ldcg(reinterpret_cast<const int *>(brow_rd) + 4 * nq - 1);
          tv = make_int4(lastv, lastv, lastv, lastv);
        }
        if (!SENT) break;
        const int S = (int)FILL16_SENTINEL;
        if (__all_sync(FULL, tv.x != S && tv.y != S && tv.z != S && tv.w != S)) break;
        __nanosleep(ns);
        if (ns < 1024) ns *= 2;
      }
    }
    __syncwarp();
    ring_sel[(blk & 1) * 32 + lane] = ring_sel[(blk & 1) * 32 + 64 + lane] = sv;
    // absolute H' -> relative G (the base only changes at block starts)
    ring_top[lane] = make_int4(negbase - tv.x, negbase - tv.y, negbase - tv.z, negbase - tv.w);
    __syncwarp();

    const int send = min(32, nsteps - q0);
    int4 tv4_next = make_int4(0, 0, 0, 0);
    uint2 sw_next = make_uint2(0u, 0u);
    if constexpr (PIPE) {  // the first step's staged entries
      tv4_next = *reinterpret_cast<const int4 *>(top_rd);
      sw_next = *reinterpret_cast<const uint2 *>(sel_rd + 8 * (q0 & 63));
    }
#pragma unroll 1  // keeps the rows in fixed registers (vector stores of the checkpoints without moves)
    for (int st = 0; st < send; ++st) {
      // ---- neighbours: every lane takes the bottom rows of the lane above (both halves); lane 0
      // takes lane 31's low half (bottom row of strip A) as its high half and the staged top row as
      // its low half
      if constexpr (PIPE) {
        const int4 tv4 = tv4_next;
        const uint2 sw = sw_next;
        tv4_next = *reinterpret_cast<const int4 *>(top_rd + 16 * ((st + 1) & 31));  // (slot 0 after the block's last step: unused)
        sw_next = *reinterpret_cast<const uint2 *>(sel_rd + 8 * ((q0 + st + 1) & 63));
        unsigned up_in[4];
        up_in[0] = prmt(sh[0], (unsigned)tv4.x, merge);
        up_in[1] = prmt(sh[1], (unsigned)tv4.y, merge);
        up_in[2] = prmt(sh[2], (unsigned)tv4.z, merge);
        up_in[3] = prmt(sh[3], (unsigned)tv4.w, merge);
        macro_step16_sh<RP>(h, dg, tl, th, up_in, sw.x, sw.y, one, outv, sh, src_lane);
      } else {
        const int4 tv4 = *reinterpret_cast<const int4 *>(top_rd + 16 * st);
        unsigned up_in[4];
        up_in[0] = prmt(__shfl_sync(FULL, outv[0], src_lane), (unsigned)tv4.x, merge);
        up_in[1] = prmt(__shfl_sync(FULL, outv[1], src_lane), (unsigned)tv4.y, merge);
        up_in[2] = prmt(__shfl_sync(FULL, outv[2], src_lane), (unsigned)tv4.z, merge);
        up_in[3] = prmt(__shfl_sync(FULL, outv[3], src_lane), (unsigned)tv4.w, merge);
        const uint2 sw = *reinterpret_cast<const uint2 *>(sel_rd + 8 * ((q0 + st) & 63));
        macro_step16<RP>(h, dg, tl, th, up_in, sw.x, sw.y, one, outv);
      }

      // ---- kept rows (absolute H'): tile traceback, hand-off to the next strip pair, corner.
      // Predicated stores, no branch: the conversions run on the FMA pipe, which has room.
      {
        const int qcB = qcA - 32;
        const int a0 = sub_lo16(negbase, outv[0]), a1 = sub_lo16(negbase, outv[1]), a2 = sub_lo16(negbase, outv[2]),
                  a3 = sub_lo16(negbase, outv[3]);
        const int b0 = sub_hi16(negbase, outv[0]), b1 = sub_hi16(negbase, outv[1]), b2 = sub_hi16(negbase, outv[2]),
                  b3 = sub_hi16(negbase, outv[3]);
        asm volatile("{.reg .pred p; setp.lt.u32 p, %5, %6; @p st.global.v4.b32 [%0], {%1,%2,%3,%4};}" ::"l"(keepA + qcA),
                     "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(qcA), "r"(nq_keep)
                     : "memory");
        asm volatile("{.reg .pred p; setp.lt.u32 p, %5, %6; @p st.global.v4.b32 [%0], {%1,%2,%3,%4};}" ::"l"(keepB + qcB),
                     "r"(b0), "r"(b1), "r"(b2), "r"(b3), "r"(qcB), "r"(nq_keep)
                     : "memory");
      }
      // ---- column checkpoints: the packed registers as they are.  One lane at a time takes this
      // branch (every lane once per 32 steps), so it is kept short: consecutive stores of a lane go
      // to consecutive images (k, low halves), (k, high halves), (k+1, low halves), ...
      if constexpr (PREDCK) {
        static_assert(!PREDCK || (CKQ == 32 && RP >= 8), "predicated checkpoint stores: a checkpoint every 32 macro columns");
        const int hit = (qcA + 1 == ck_next) ? 1 : 0;
#pragma unroll
        for (int r8 = 0; r8 < RP / 8; ++r8)
          asm volatile("{.reg .pred p; setp.ne.s32 p, %9, 0; @p st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};}" ::"l"(ckp + 8 * r8),
                       "r"(h[8 * r8]), "r"(h[8 * r8 + 1]), "r"(h[8 * r8 + 2]), "r"(h[8 * r8 + 3]), "r"(h[8 * r8 + 4]),
                       "r"(h[8 * r8 + 5]), "r"(h[8 * r8 + 6]), "r"(h[8 * r8 + 7]), "r"(hit)
                       : "memory");
        if (hit) {
          ckp += halfwords;
          ck_next = ck_next + 32 <= 32 * ck_last ? ck_next + 32 : 0x7fffffff;
        }
      } else if (((qcA + 1) & 31) == 0) {
        const int ckcol = (qcA + 1) >> 5;
        bool store;
        uint32_t *dst;
        if constexpr (CKQ == 64) {
          // even: the low halves are at a checkpoint column, odd: the high halves
          const int qc = (ckcol & 1) ? qcA - 32 : qcA;
          store = ckcol >= 2 && qc < nq;
          dst = ck + (int64_t)(ckcol - 2) * (int64_t)halfwords;
        } else {
          // a checkpoint every 32 macro columns: both halves are at one — the low halves at
          // checkpoint ckcol - 1, the high halves at checkpoint ckcol - 2 — so image ckcol - 1 holds
          // [rows of A at checkpoint ckcol - 1 | rows of B at checkpoint ckcol - 2]; stored while a
          // tile to the right of either column exists
          store = ckcol >= 1 && (qcA + 1 < nq || (ckcol >= 2 && qcA - 31 < nq));
          dst = ck + (int64_t)(ckcol - 1) * (int64_t)halfwords;
        }
        if (store && RP == 4) {
          asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        } else if (store) {
#pragma unroll
          for (int r8 = 0; r8 < RP / 8; ++r8)
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8 * r8), "r"(h[8 * r8]),
                         "r"(h[8 * r8 + 1]), "r"(h[8 * r8 + 2]), "r"(h[8 * r8 + 3]), "r"(h[8 * r8 + 4]),
                         "r"(h[8 * r8 + 5]), "r"(h[8 * r8 + 6]), "r"(h[8 * r8 + 7])
                         : "memory");
        }
      }
      ++qcA;
    }
    // ---- publish the progress of the bottom row (virtual lane 63 = lane 31's high halves), once
    // per block: the next strip pair consumes it in blocks of 32 columns as well
    if (!last && !SENT) {
      const int done = min(nq, q0 + send - 63);
      if (done > 0 && lane == 31) st_release_u32(my_progress, (unsigned)done);
    }
  }
}

template <int RP, int MINB>
__global__ void __launch_bounds__(F16_THREADS, MINB) fill16_kernel(const FillArgs a) {
  __shared__ uint2 s_sel[F16_THREADS / 32][128];
  __shared__ int4 s_top[F16_THREADS / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.ntasks) break;
    const StripTask task = a.tasks[t];
    const PairDesc P = a.pairs[task.pair];
    const int s = (int)task.strip;
    // main launch: strip pairs of 64 * main_rp rows (1024 or 2048); a pair's shorter last strip
    // pairs (1024 rows in the three-level geometry, then 512 rows; see strip_geometry) run in later
    // launches with 16 / 8 rows per half lane
    int row_base;
    if (RP == a.main_rp)
      row_base = s * (64 * RP);
    else if (RP == 16)
      row_base = main_strips(P.nstrips, P.half_last) * 2048;
    else
      row_base = (int)P.mpad - 512;
    process_strip_pair<RP>(a, P, s, row_base, lane, s_sel[warp], s_top[warp]);
  }
}

// One persistent kernel for ALL strip pairs of a batch: the main ones (64 * MAINRP rows) and a
// pair's shorter last ones (1024 rows in the three-level geometry, 512 rows), which wait for their
// predecessor through the same progress counters as every other strip pair.  One queue, one
// ramp-down per batch instead of three (the tail launches of a small batch — one GPU's share of a
// chunk on 8 GPUs — are shorter than one wave of warps).
template <int MAINRP, int MINB, bool PREDCK = false, bool PIPE = false, bool SENT = false>
__global__ void __launch_bounds__(F16_THREADS, MINB) fill16_all_kernel(const FillArgs a) {
  __shared__ uint2 s_sel[F16_THREADS / 32][128];
  __shared__ int4 s_top[F16_THREADS / 32][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.ntasks) break;
    const StripTask task = a.tasks[t];
    const PairDesc P = a.pairs[task.pair];
    const int s = (int)task.strip;
    const int nmain = main_strips(P.nstrips, P.half_last);
    if (s < nmain)
      process_strip_pair<MAINRP, PREDCK && MAINRP >= 8, PIPE, SENT>(a, P, s, s * (64 * MAINRP), lane, s_sel[warp], s_top[warp]);
    else if (MAINRP == 32 && (P.half_last & 2) && s == nmain)
      process_strip_pair<16, PREDCK>(a, P, s, nmain * 2048, lane, s_sel[warp], s_top[warp]);
    else
      process_strip_pair<8, PREDCK, false, SENT>(a, P, s, (int)P.mpad - 512, lane, s_sel[warp], s_top[warp]);
  }
}

// ---- selector streams, one per sequence (built once per load): entry e = two words, 16 bits per
// DP column c of macro column e: nibbles [code of y[4e+c] | 8, 8, 4 + code of y[4(e-32)+c] | 0xC, 0xC]
__global__ void build_sel16_kernel(const uint8_t *__restrict__ codes, uint8_t *__restrict__ sel16,
                                   const uint32_t *__restrict__ off, const int32_t *__restrict__ len, int k) {
  const int seq = blockIdx.y;
  if (seq >= k) return;
  const int n = len[seq], nq = (n + 3) >> 2;
  const uint8_t *y = codes + off[seq];
  uint2 *out = reinterpret_cast<uint2 *>(sel16 + 2 * (uint64_t)off[seq]);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nq + 64; e += gridDim.x * blockDim.x) {
    unsigned w[2] = {0u, 0u};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ja = 4 * e + c, jb = 4 * (e - 32) + c;
      const unsigned lo = (ja < n) ? (unsigned)(y[ja] >> 3) : 8u;
      const unsigned hi = (jb >= 0 && jb < n) ? 4u + (unsigned)(y[jb] >> 3) : 0xCu;
      const unsigned s16 = 0xC080u | lo | (hi << 8);
      w[c >> 1] |= s16 << (16 * (c & 1));
    }
    out[e] = make_uint2(w[0], w[1]);
  }
}

// FILL16_SENT: fills the hand-off rows of every pair (the kept row above each strip pair but the first)
// with sentinel words.  blockIdx.y = pair.
__global__ void fill16_sentinel_kernel(const PairDesc *__restrict__ pairs, int32_t *__restrict__ brow, int rows_per_pair) {
  const PairDesc P = pairs[blockIdx.y];
  const int nq = (P.n + 3) >> 2;
  const uint64_t bstride = (uint64_t)(4 * nq + 4);
  const int4 sent = make_int4((int)FILL16_SENTINEL, (int)FILL16_SENTINEL, (int)FILL16_SENTINEL, (int)FILL16_SENTINEL);
  const long long total = (long long)(P.nstrips - 1) * nq;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t / nq) + 1, q = (int)(t - (long long)(s - 1) * nq);
    reinterpret_cast<int4 *>(brow + P.brow_off + ((uint64_t)s * rows_per_pair - 1) * bstride)[q] = sent;
  }
}

typedef void (*Fill16Ptr)(const FillArgs);
// the one-launch kernel; flags: FILL16_PREDCK / FILL16_PIPE
Fill16Ptr pick16_all(int rp, int minb, int flags) {
  const bool predck = flags & FILL16_PREDCK, pipe = flags & FILL16_PIPE;
  if (rp == 32) {
    // 128 registers without spills = 4 CTAs per SM; with the predicated checkpoint stores the 3-CTA build
    // (134 registers) is the faster one (cfg4 job 4.69 s -> 4.61 s, same box)
    if (predck) return minb == 4 ? fill16_all_kernel<32, 4, true> : fill16_all_kernel<32, 3, true>;
    return minb == 3 ? fill16_all_kernel<32, 3> : fill16_all_kernel<32, 4>;
  }
  if (rp == 16) {
    if (predck && (flags & FILL16_SENT)) return fill16_all_kernel<16, 3, true, false, true>;
    if (predck) return minb == 6 ? fill16_all_kernel<16, 6, true> : fill16_all_kernel<16, 3, true>;
    return minb == 5 ? fill16_all_kernel<16, 5> : fill16_all_kernel<16, 6>;
  }
  // thin strip pairs (512 / 256 rows) for jobs of a few long pairs: more warps along the wavefront
  if (rp == 8) {
    if (predck && pipe && (flags & FILL16_SENT)) return fill16_all_kernel<8, 4, true, true, true>;  // uniform strip pairs only
    if (predck && pipe) return minb == 8 ? fill16_all_kernel<8, 8, true, true> : fill16_all_kernel<8, 4, true, true>;
    return predck ? fill16_all_kernel<8, 8, true> : fill16_all_kernel<8, 8>;
  }
  if (rp == 4) return fill16_all_kernel<4, 8>;
  return nullptr;
}
Fill16Ptr pick16(int rp, int minb, int flags) {
  if (rp >= 100) return pick16_all(rp - 100, minb, flags);  // 100 + main rp: the one-launch kernel
  // 32 rows per half lane: 128 registers without spills = 4 CTAs per SM (cfg4 fill 4.14 s; 3 CTAs: 4.20 s)
  if (rp == 32) return minb == 3 ? fill16_kernel<32, 3> : fill16_kernel<32, 4>;
  if (rp == 16) {
    // 6 CTAs per SM (<= 80 registers): 161 ms per cfg3 job against 168 (5 CTAs), 173 (4), 180 (7)
    if (minb == 4) return fill16_kernel<16, 4>;
    if (minb == 5) return fill16_kernel<16, 5>;
    if (minb == 7) return fill16_kernel<16, 7>;
    return fill16_kernel<16, 6>;
  }
  if (rp == 8) return fill16_kernel<8, 8>;
  return nullptr;
}

}  // namespace

// 16-bit relative scores need: byte-sized gains, and the spread of a strip pair plus the growth
// over one re-base interval below 2^16.  Returns the re-base interval's log2 (>= 6), or 0.
int fill16_rebase_shift(int pxy, int pgap, int rp) {
  if (pxy < 0 || pgap < 0 || 2 * (int64_t)pgap > 127) return 0;
  if (pgap == 0) return 10;
  // values within 2*pgap*(256 + 64*rp) of each other + 2*pgap*4 per step + 1024 steps of slack for halves
  // that have finished
  for (int ks = 10; ks >= 6; --ks)
    if (2ll * pgap * (4ll * (1 << ks) + 256 + 64 * rp + 4 * 256) <= 65535 - 2ll * pgap) return ks;
  return 0;
}

cudaError_t launch_build_sel16(const uint8_t *codes, uint8_t *sel16, const uint32_t *off_dev, const int32_t *len_dev,
                               int k, int maxlen, cudaStream_t stream) {
  if (k == 0) return cudaSuccess;
  const int entries = ((maxlen + 3) >> 2) + 64;
  dim3 grid((unsigned)std::min(64, (entries + 255) / 256), (unsigned)k);
  build_sel16_kernel<<<grid, 256, 0, stream>>>(codes, sel16, off_dev, len_dev, k);
  return cudaGetLastError();
}

cudaError_t launch_fill16_sentinels(const PairDesc *pairs, uint32_t npairs, int32_t *brow, int main_rp, uint64_t max_words,
                                    cudaStream_t stream) {
  if (npairs == 0) return cudaSuccess;
  const unsigned bx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(64, (max_words / 4 + 255) / 256));
  fill16_sentinel_kernel<<<dim3(bx, npairs), 256, 0, stream>>>(pairs, brow, 64 * main_rp / TILE_ROWS);
  return cudaGetLastError();
}

int fill16_blocks_per_sm(int rp, int minb, int flags) {
  Fill16Ptr k = pick16(rp, minb, flags);
  if (!k) return 0;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, F16_THREADS, 0);
  return nb;
}

cudaError_t launch_fill16(const FillArgs &args, int rp, int minb, int sm_count, cudaStream_t stream, int flags,
                          int max_ctas_per_sm) {
  if (args.ntasks == 0) return cudaSuccess;
  Fill16Ptr k = pick16(rp, minb, flags);
  if (!k) return cudaErrorInvalidValue;
  int nb = fill16_blocks_per_sm(rp, minb, flags);
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  if (max_ctas_per_sm > 0) nb = std::min(nb, max_ctas_per_sm);  // fewer resident strip pairs, each of them faster
  long long want = ((long long)args.ntasks + (F16_THREADS / 32) - 1) / (F16_THREADS / 32);
  long long grid = (long long)sm_count * nb;
  if (grid > want) grid = want;
  k<<<(unsigned)grid, F16_THREADS, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace seqalk
