// fill.cu — integer DP matrix fill, hand-written for sm_100a.
//
// Replaces the reference's OpenMP tile-wavefront fill
//   submit/xuliny-seqalkway.cpp:435-491 (getMinimumPenalty2: init, tiling, recurrence :478-487)
// and its sequential twin :584-609 (getMinimumPenalty).
//
// Work decomposition ("strip wavefront"):
//   * a pair's DP matrix is cut into strips of 32*R rows; one warp sweeps one strip
//     left to right; lane t owns R consecutive rows and trails lane t-1 by one macro
//     column (4 DP columns), i.e. the warp is an anti-diagonal wavefront of R x 4 blocks;
//   * the value crossing from lane t-1's bottom row to lane t's top row travels by
//     __shfl_up_sync (4 per macro step), and so does the 4-byte word of y;
//   * strip s hands its bottom row to strip s+1 through a per-pair boundary-row
//     buffer (L2 resident) plus a per-strip progress counter (release/acquire), so
//     consecutive strips of one pair run concurrently on different warps / CTAs /
//     SMs as a second-level wavefront ("CTA wavefront" for long sequences), while
//     a batch with many pairs degenerates to independent warps ("warp per pair");
//   * strips are claimed from a global queue ordered (strip index, pair): every
//     strip's predecessor is claimed earlier, by a warp that is already running,
//     so the waits cannot deadlock and no cooperative launch is needed;
//   * scores stay in registers in a gap-shifted domain H'[i][j] = H[i][j] - (i+j)*pgap,
//     which makes all boundaries 0 and both gap moves free of adds:
//         H'[i][j] = match ? H'[i-1][j-1] - 2*pgap
//                          : min(H'[i-1][j-1] + pxy - 2*pgap, H'[i-1][j], H'[i][j-1])
//     (comparisons are shift-invariant, so ties break exactly as in the reference);
//   * per cell a 2-bit code is emitted: bit1 = "not diagonal", bit0 = "left strictly
//     better than up".  The traceback reads DIAG if !bit1, else LEFT if bit0, else UP —
//     the reference's priority order (match, mismatch-diag, up, left; :502-531).
//     128 bits per lane are accumulated in registers and written with one 16 B store;
//     a warp's store is one contiguous 512 B segment (strip-major, step-skewed layout).
#include <cstdlib>

#include "seqalk_kernels.h"

namespace seqalk {

constexpr int FILL_THREADS = 128;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// a + b, optionally forced onto the FMA pipe as a*one + b (one == 1 at run time, unknown
// to ptxas) so that the ALU pipe is left to VIMNMX3 / SHF / ISETP / PRMT.
template <int ADDMODE>
__device__ __forceinline__ int add_i(int a, int b, int one) {
  if (ADDMODE == 0) return a + b;
  int r;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
  return r;
}
template <int ADDMODE>
__device__ __forceinline__ int sub_i(int a, int b, int minus_one) {  // a - b
  if (ADDMODE == 0) return a - b;
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
template <int R, int MODE, int ADDMODE>
__device__ __forceinline__ void macro_step(int (&h)[R], int &dg, const int (&xc)[R], const int (&up_in)[4],
                                           unsigned yword, int A, int B, int one, int minus_one,
                                           int (&outv)[4], unsigned (&bits)[R / 4]) {
#pragma unroll
  for (int w = 0; w < R / 4; ++w) bits[w] = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int yc = (int)((yword >> (8 * c)) & 0xffu);
    unsigned sel = 0;
    if (MODE == MODE_TAB) sel = (unsigned)yc * 0x1111u + 0x8880u;  // byte yc, sign-replicated
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
        const unsigned nd = ((!match) && (t > m1)) ? 2u : 0u;
        const unsigned lf = (up > left) ? 1u : 0u;
        bits[w] = (bits[w] << 2) | nd | lf;
      } else {
        int t;
        if (MODE == MODE_TAB) {
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
        const int d1 = sub_i<ADDMODE>(v, t, minus_one);
        const int d2 = sub_i<ADDMODE>(left, up, minus_one);
        bits[w] = __funnelshift_l((unsigned)d1, bits[w], 1);
        bits[w] = __funnelshift_l((unsigned)d2, bits[w], 1);
      }
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
  }
}

template <int R, int MODE, int ADDMODE>
__device__ __forceinline__ void process_strip(const FillArgs &a, const PairDesc &P, const int s, const int lane) {
  constexpr int G = 16 / R;  // macro steps per 16 B direction store
  const unsigned FULL = 0xffffffffu;
  const int m = P.m, n = P.n;
  const int nq = (n + 3) >> 2;
  const int ngroups = (nq + 31 + G - 1) / G;
  const int A = -2 * a.pgap, B = a.pxy - 2 * a.pgap;
  const int one = a.one, minus_one = -a.one;
  const bool first = (s == 0), last = (s == P.nstrips - 1);

  const uint8_t *__restrict__ seqbase = (MODE == MODE_TAB) ? a.codes : a.seq;
  const uint8_t *__restrict__ xs = seqbase + P.x_off;
  const unsigned *__restrict__ yw32 = reinterpret_cast<const unsigned *>(seqbase + P.y_off);
  int4 *__restrict__ brow = reinterpret_cast<int4 *>(a.brow + P.brow_off);
  uint4 *__restrict__ dout =
      reinterpret_cast<uint4 *>(a.dirs + P.dir_off + (uint64_t)s * ((uint64_t)ngroups * 512ull)) + lane;
  const unsigned *prev_progress = a.progress + P.strip0 + (first ? 0 : s - 1);
  unsigned *my_progress = a.progress + P.strip0 + s;

  // this lane's rows
  const int row0 = s * (32 * R) + lane * R;
  int xc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool inside = row0 + r < m;
    const int ch = inside ? (int)xs[row0 + r] : 0;
    if (MODE == MODE_TAB) {
      const unsigned Ab = (unsigned)A & 0xffu, Bb = (unsigned)B & 0xffu;
      unsigned tab = Bb * 0x01010101u;
      if (inside) tab ^= (Ab ^ Bb) << (8 * (ch & 3));
      xc[r] = (int)tab;
    } else {
      xc[r] = inside ? ch : (0x100 + r);
    }
  }

  int h[R];
#pragma unroll
  for (int r = 0; r < R; ++r) h[r] = 0;
  int dg = 0;
  int outv[4] = {0, 0, 0, 0};
  unsigned yw = 0;
  int avail = first ? nq : 0;  // macro cols of the predecessor's bottom row known to be published

  // lane 0's look-ahead: boundary values + y words of the next group.  They are loaded
  // right after the current group's last use of these registers, one macro step (hundreds
  // of issue slots) before they are needed: no second buffer, no register moves.
  int4 top[G];
  unsigned ynext[G];
#pragma unroll
  for (int u = 0; u < G; ++u) {
    top[u] = make_int4(0, 0, 0, 0);
    ynext[u] = 0u;
  }
  auto fetch_group = [&](int g) {
    const int need = min(nq, g * G + G);
    if (avail < need) {  // predecessor strip still in flight: wait for its next publication
      // Poll with relaxed loads (an acquire load invalidates the SM's whole L1 — CCTL.IVALL —
      // which would evict the co-resident traceback / digest kernels' lines on every poll);
      // one acquire load once the value is there orders the boundary reads after it.
      unsigned ns = 32;
      for (;;) {
        unsigned v = 0;
        if (lane == 0) v = ld_relaxed_u32(prev_progress);
        v = __shfl_sync(FULL, v, 0);
        if ((int)v >= need) break;
        __nanosleep(ns);
        if (ns < 4096) ns *= 2;
      }
      unsigned v = 0;
      if (lane == 0) v = ld_acquire_u32(prev_progress);
      avail = (int)__shfl_sync(FULL, v, 0);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int qc = g * G + u;
        if (qc < nq) {
          ynext[u] = __ldg(yw32 + qc);
          if (!first) top[u] = __ldcg(brow + qc);
        }
      }
    }
  };
  fetch_group(0);

  for (int g = 0; g < ngroups; ++g) {
    unsigned pk[4];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int q = g * G + u;
      int up_in[4];
      up_in[0] = __shfl_up_sync(FULL, outv[0], 1);
      up_in[1] = __shfl_up_sync(FULL, outv[1], 1);
      up_in[2] = __shfl_up_sync(FULL, outv[2], 1);
      up_in[3] = __shfl_up_sync(FULL, outv[3], 1);
      unsigned ny = __shfl_up_sync(FULL, yw, 1);
      if (lane == 0) {
        up_in[0] = top[u].x;
        up_in[1] = top[u].y;
        up_in[2] = top[u].z;
        up_in[3] = top[u].w;
        ny = ynext[u];
      }
      if (u == G - 1 && (g + 1) * G < nq) fetch_group(g + 1);
      const int qc = q - lane;  // this lane's macro column
      unsigned bits[R / 4];
#pragma unroll
      for (int w = 0; w < R / 4; ++w) bits[w] = 0u;
      if (qc >= 0 && qc < nq) {
        yw = ny;
        macro_step<R, MODE, ADDMODE>(h, dg, xc, up_in, yw, A, B, one, minus_one, outv, bits);
        if (lane == 31 && !last) brow[qc] = make_int4(outv[0], outv[1], outv[2], outv[3]);
      }
#pragma unroll
      for (int w = 0; w < R / 4; ++w) pk[u * (R / 4) + w] = bits[w];
    }
    dout[(size_t)g * 32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);

    // publish progress of the bottom row (lane 31 trails by 31 macro steps)
    if (!last) {
      const int done = min(nq, (g + 1) * G - 31);
      if (done > 0) {
        const int prev_done = max(0, min(nq, g * G - 31));
        if (done != prev_done && (done == nq || (done / PUBLISH_EVERY) != (prev_done / PUBLISH_EVERY))) {
          if (lane == 31) st_release_u32(my_progress, (unsigned)done);
        }
      }
    }
  }
}

template <int R, int MODE, int ADDMODE>
__global__ void __launch_bounds__(FILL_THREADS) fill_kernel(const FillArgs a) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.ntasks) break;
    const StripTask task = a.tasks[t];
    const PairDesc P = a.pairs[task.pair];
    process_strip<R, MODE, ADDMODE>(a, P, (int)task.strip, lane);
  }
}

typedef void (*FillKernelPtr)(const FillArgs);

static FillKernelPtr pick_kernel(const FillVariant &v) {
#define SEQALK_PICK(RR)                                                                     \
  if (v.R == RR) {                                                                          \
    if (v.mode == MODE_GENERAL) return fill_kernel<RR, MODE_GENERAL, 0>;                    \
    if (v.mode == MODE_CMP) return v.addmode ? fill_kernel<RR, MODE_CMP, 1> : fill_kernel<RR, MODE_CMP, 0>; \
    if (v.mode == MODE_TAB) return v.addmode ? fill_kernel<RR, MODE_TAB, 1> : fill_kernel<RR, MODE_TAB, 0>; \
  }
  SEQALK_PICK(8)
  SEQALK_PICK(16)
#undef SEQALK_PICK
  return nullptr;
}

// Resident fill CTAs per SM.  The fill is a persistent kernel; it must leave room for the
// small traceback / digest CTAs of the previous batch (<= 64 threads x 64 registers), which
// run concurrently on the high-priority stream: keep >= 4096 registers of every SM free.
int fill_blocks_per_sm(const FillVariant &v) {
  FillKernelPtr k = pick_kernel(v);
  if (!k) return 0;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, FILL_THREADS, 0);
  cudaFuncAttributes attr;
  if (cudaFuncGetAttributes(&attr, k) == cudaSuccess) {
    const int regs_alloc = (attr.numRegs + 7) / 8 * 8;
    int headroom = 4096;
    if (const char *e = getenv("SEQALK_FILL_HEADROOM_REGS")) headroom = atoi(e);
    while (nb > 1 && 65536 - nb * FILL_THREADS * regs_alloc < headroom) --nb;
  }
  return nb;
}

cudaError_t launch_fill(const FillArgs &args, const FillVariant &v, int sm_count, cudaStream_t stream) {
  if (args.ntasks == 0) return cudaSuccess;
  FillKernelPtr k = pick_kernel(v);
  if (!k) return cudaErrorInvalidValue;
  int nb = fill_blocks_per_sm(v);
  if (nb < 1) return cudaErrorLaunchOutOfResources;
  // persistent grid: every SM full, but no more warps than strips
  long long want = ((long long)args.ntasks + (FILL_THREADS / 32) - 1) / (FILL_THREADS / 32);
  long long grid = (long long)sm_count * nb;
  if (grid > want) grid = want;
  k<<<(unsigned)grid, FILL_THREADS, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace seqalk
