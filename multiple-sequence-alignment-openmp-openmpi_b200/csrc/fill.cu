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
//   * tile mode (default, EMIT = false): scores only — per cell one IDP.4A (cost-table lookup
//     + add, FMA pipe) and one VIMNMX3 (ALU pipe); the row of H' is kept every TILE_ROWS
//     (128) DP rows and each lane checkpoints its rows every 256 DP columns, from which
//     traceback_tiles.cu recomputes directions for the tiles on the optimal path only;
//   * directions mode (EMIT = true): per cell a 2-bit code is emitted: bit1 = "not diagonal",
//     bit0 = "left strictly better than up".  The traceback reads DIAG if !bit1, else LEFT if
//     bit0, else UP — the reference's priority order (match, mismatch-diag, up, left;
//     :502-531).  128 bits per lane are accumulated in registers and written with one 16 B
//     store; a warp's store is one contiguous 512 B segment (strip-major, step-skewed layout).
#include <cstdlib>

#include "fill_core.cuh"

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

// EMIT = true : directions mode.  2-bit codes for every cell go to HBM (P.dir_off); the boundary
//               row buffer holds one row per pair, rewritten in place strip after strip.
// EMIT = false: tile mode.  Scores only; every TILE_ROWS = 128 DP rows (lanes 3, 7, ..., 31 for R = 32) the
//               row of H' is kept (P.brow_off: R/8 rows per strip), and every CKQ macro columns each
//               lane stores its R row values (P.dir_off reused as the checkpoint region:
//               [checkpoint][padded row] int32).  The traceback recomputes directions only for
//               the TILE_ROWS x 4*CKQ tiles its path visits (traceback_tiles.cu).
template <int R, int MODE, int ADDMODE, bool EMIT>
__device__ __forceinline__ void process_strip(const FillArgs &a, const PairDesc &P, const int s, const int row_base,
                                              const int lane) {
  constexpr int G = EMIT ? 16 / R : 2;  // macro steps per loop iteration (per 16 B direction store)
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
  // boundary rows: in place (EMIT) or one row per TILE_ROWS DP rows (tile mode, stride 4*nq+4 ints)
  const uint64_t bstride = EMIT ? 0 : (uint64_t)(4 * nq + 4);
  int32_t *bbase = a.brow + P.brow_off;
  const int4 *__restrict__ brow_rd =
      reinterpret_cast<const int4 *>(EMIT ? bbase : bbase + (uint64_t)(first ? 0 : row_base / TILE_ROWS - 1) * bstride);
  int4 *__restrict__ brow_wr = reinterpret_cast<int4 *>(
      EMIT ? bbase : bbase + (uint64_t)((row_base + lane * R) / TILE_ROWS) * bstride);  // kept row (DP row / TILE_ROWS) - 1
  uint4 *__restrict__ dout =
      reinterpret_cast<uint4 *>(a.dirs + P.dir_off + (uint64_t)s * ((uint64_t)ngroups * 512ull)) + lane;
  const unsigned *prev_progress = a.progress + P.strip0 + (first ? 0 : s - 1);
  unsigned *my_progress = a.progress + P.strip0 + s;

  // this lane's rows
  const int row0 = row_base + lane * R;
  int xc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool inside = row0 + r < m;
    const int ch = inside ? (int)xs[row0 + r] : 0;
    xc[r] = row_operand<MODE>(inside, ch, r, A, B);
  }
  // tile mode: column checkpoints of this lane's rows, [checkpoint][padded row]
  const uint64_t mpad = (uint64_t)P.mpad;
  int4 *__restrict__ ck_wr = reinterpret_cast<int4 *>(a.dirs + P.dir_off) + (uint64_t)row0 / 4;

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
          if (!first) top[u] = __ldcg(brow_rd + qc);
        }
      }
    }
  };
  fetch_group(0);

  for (int g = 0; g < ngroups; ++g) {
    unsigned pk[4] = {0u, 0u, 0u, 0u};
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
        macro_step<R, MODE, ADDMODE, EMIT>(h, dg, xc, up_in, yw, A, B, one, minus_one, outv, bits);
        if (EMIT) {
          if (lane == 31) brow_wr[qc] = make_int4(outv[0], outv[1], outv[2], outv[3]);  // last strip: H'[m][.] (row_operand)
        } else {
          // every TILE_ROWS rows (R = 32: lanes 3, 7, ..., 31) keep the row for the tile traceback
          if ((lane & (TILE_ROWS / R - 1)) == (TILE_ROWS / R - 1)) brow_wr[qc] = make_int4(outv[0], outv[1], outv[2], outv[3]);
          // column checkpoint: H' of this lane's rows at the last column of macro col qc
          if (((qc + 1) & (CKQ - 1)) == 0) {
            // 256-bit stores (STG.E.256, sm_100): the branch is taken by one lane at a time, so
            // every instruction in it costs the whole warp an issue slot
            int4 *dst = ck_wr + ((uint64_t)((qc + 1) / CKQ - 1) * mpad) / 4;
#pragma unroll
            for (int r8 = 0; r8 < R / 8; ++r8)
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 2 * r8), "r"(h[8 * r8]),
                           "r"(h[8 * r8 + 1]), "r"(h[8 * r8 + 2]), "r"(h[8 * r8 + 3]), "r"(h[8 * r8 + 4]),
                           "r"(h[8 * r8 + 5]), "r"(h[8 * r8 + 6]), "r"(h[8 * r8 + 7])
                           : "memory");
          }
        }
      }
      if (EMIT) {
#pragma unroll
        for (int w = 0; w < R / 4; ++w) pk[u * (R / 4) + w] = bits[w];
      }
    }
    if (EMIT) dout[(size_t)g * 32] = make_uint4(pk[0], pk[1], pk[2], pk[3]);

    // publish progress of the bottom row (lane 31 trails by 31 macro steps)
    if (!last) {
      const int done = min(nq, (g + 1) * G - 31);
      if (done > 0) {
        const int prev_done = max(0, min(nq, g * G - 31));
        if (done != prev_done && (done == nq || (done >> a.publish_shift) != (prev_done >> a.publish_shift))) {
          if (lane == 31) st_release_u32(my_progress, (unsigned)done);
        }
      }
    }
  }
}

// MINB: minimum resident CTAs per SM the register allocation must allow (1 = no constraint)
template <int R, int MODE, int ADDMODE, bool EMIT, int MINB = 1>
__global__ void __launch_bounds__(FILL_THREADS, MINB) fill_kernel(const FillArgs a) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(a.counter, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= a.ntasks) break;
    const StripTask task = a.tasks[t];
    const PairDesc P = a.pairs[task.pair];
    const int s = (int)task.strip;
    // a pair's 512-row last strip (tile mode, see strip_geometry) is run by the R = 16 kernel
    // in a second, small launch; it starts where the 1024-row strips end
    const int row_base = (!EMIT && P.half_last && s == P.nstrips - 1) ? (int)P.mpad - 512 : s * (32 * R);
    process_strip<R, MODE, ADDMODE, EMIT>(a, P, s, row_base, lane);
  }
}

typedef void (*FillKernelPtr)(const FillArgs);

static FillKernelPtr pick_kernel(const FillVariant &v) {
  if (!v.emit && v.R == 32) {  // tile mode, 32 rows per lane
    if (v.mode == MODE_GENERAL) return fill_kernel<32, MODE_GENERAL, 0, false>;
    if (v.mode == MODE_CMP) return fill_kernel<32, MODE_CMP, 0, false>;
    if (v.mode == MODE_TAB && v.addmode == 2) {
      // left alone ptxas takes 168 registers (3 CTAs per SM); capped at 128 (4 CTAs, 8 bytes of
      // spill outside the loop) the fill is 0.8 % faster (same-box A/B, cfg3 and cfg4)
      if (v.minb == 1) return fill_kernel<32, MODE_TAB, 2, false>;
      if (v.minb == 5) return fill_kernel<32, MODE_TAB, 2, false, 5>;
      if (v.minb == 6) return fill_kernel<32, MODE_TAB, 2, false, 6>;
      return fill_kernel<32, MODE_TAB, 2, false, 4>;
    }
    if (v.mode == MODE_TAB) return fill_kernel<32, MODE_TAB, 0, false>;
    return nullptr;
  }
  if (!v.emit) {  // tile mode, 16 rows per lane
    if (v.R != 16) return nullptr;
    if (v.mode == MODE_GENERAL) return fill_kernel<16, MODE_GENERAL, 0, false>;
    if (v.mode == MODE_CMP) return v.addmode ? fill_kernel<16, MODE_CMP, 1, false> : fill_kernel<16, MODE_CMP, 0, false>;
    if (v.mode == MODE_TAB) {
      if (v.addmode == 2) return fill_kernel<16, MODE_TAB, 2, false>;
      return v.addmode ? fill_kernel<16, MODE_TAB, 1, false> : fill_kernel<16, MODE_TAB, 0, false>;
    }
    return nullptr;
  }
#define SEQALK_PICK(RR)                                                                     \
  if (v.R == RR) {                                                                          \
    if (v.mode == MODE_GENERAL) return fill_kernel<RR, MODE_GENERAL, 0, true>;              \
    if (v.mode == MODE_CMP)                                                                 \
      return v.addmode ? fill_kernel<RR, MODE_CMP, 1, true> : fill_kernel<RR, MODE_CMP, 0, true>; \
    if (v.mode == MODE_TAB)                                                                 \
      return v.addmode ? fill_kernel<RR, MODE_TAB, 1, true> : fill_kernel<RR, MODE_TAB, 0, true>; \
  }
  if (v.R == 16 && v.mode == MODE_TAB && v.addmode == 2) return fill_kernel<16, MODE_TAB, 2, true>;
  SEQALK_PICK(8)
  SEQALK_PICK(16)
#undef SEQALK_PICK
  return nullptr;
}

// Resident fill CTAs per SM.  In directions mode the fill is a persistent kernel that must
// leave room for the small traceback / digest CTAs of the previous batch (<= 64 threads x 64
// registers), which run concurrently on the high-priority stream: keep >= 4096 registers of
// every SM free.  In tile mode the stages run one after the other: no headroom.
int fill_blocks_per_sm(const FillVariant &v) {
  FillKernelPtr k = pick_kernel(v);
  if (!k) return 0;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, FILL_THREADS, 0);
  cudaFuncAttributes attr;
  if (v.emit && cudaFuncGetAttributes(&attr, k) == cudaSuccess) {
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
