// traceback_tiles.cu — traceback for the tile mode: directions are NOT read from HBM but
// recomputed, one 256-row x (4*CKQ)-column tile at a time, only for the tiles the optimal path
// visits (about (m + n) / 256 of them), from the checkpoints the score-only fill left behind:
// one row of H' every 256 DP rows and one column of H' every 4*CKQ DP columns.
//
// Same reference semantics as traceback.cu (submit/xuliny-seqalkway.cpp:502-558, :191-214):
// the recompute runs the very same per-cell update as the fill (fill_core.cuh, EMIT = true), so
// the 2-bit codes — and with them the tie-breaking — are identical to the directions mode.
//
// One warp per pair.  Per tile: lane t owns 8 rows, the warp sweeps the needed macro columns
// as a skewed wavefront (exactly like the fill) writing the codes to shared memory, then walks
// the path inside the tile until it leaves through the top or the left edge.
#include "fill_core.cuh"

namespace seqalk {

namespace {

constexpr int RT = 8;  // rows per lane in the recompute: 32 lanes x 8 rows = TILE_ROWS
static_assert(32 * RT == TILE_ROWS, "tile height");

// Right-to-left writer (32-bit words); every lane tracks the position, lane 0 stores.
struct BackWriter0 {
  uint8_t *base;
  uint32_t pos;
  uint32_t acc;
  bool store;
  __device__ __forceinline__ void init(uint8_t *b, uint32_t end, bool st) {
    base = b;
    pos = end;
    acc = 0;
    store = st;
  }
  __device__ __forceinline__ void put(uint32_t ch) {
    --pos;
    acc |= ch << (8 * (pos & 3));
    if ((pos & 3) == 0) {
      if (store) *reinterpret_cast<uint32_t *>(base + pos) = acc;
      acc = 0;
    }
  }
  __device__ __forceinline__ void finish() {
    if ((pos & 3) && store) *reinterpret_cast<uint32_t *>(base + (pos & ~3u)) = acc;
  }
};

}  // namespace

template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(32) traceback_tiles_kernel(const TracebackArgs a) {
  extern __shared__ uint2 sdirs[];  // [CKQ][32]: codes of the tile's macro column qc, lane t
  // the tile's characters, staged once per tile for the walk: x rows, then y columns
  uint64_t *schars = reinterpret_cast<uint64_t *>(sdirs + CKQ * 32);
  const uint8_t *sxb = reinterpret_cast<const uint8_t *>(schars);
  const uint8_t *syb = sxb + TILE_ROWS;
  const uint32_t *sdirs32 = reinterpret_cast<const uint32_t *>(sdirs);
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const uint32_t p = blockIdx.x;
  const PairDesc P = a.pairs[p];
  const int m = P.m, n = P.n;
  const int nq = (n + 3) >> 2;
  const uint64_t bstride = (uint64_t)(4 * nq + 4);
  const uint64_t mpad = (uint64_t)P.nstrips * 32 * (uint64_t)a.R;  // rows padded to the fill's strips
  const int A = -2 * a.pgap, B = a.pxy - 2 * a.pgap;

  const uint8_t *__restrict__ xs = a.seq + P.x_off;  // raw characters for the output
  const uint8_t *__restrict__ ys = a.seq + P.y_off;
  const uint8_t *__restrict__ opbase = (MODE == MODE_TAB) ? a.codes : a.seq;
  const uint8_t *__restrict__ xop = opbase + P.x_off;
  const unsigned *__restrict__ yw32 = reinterpret_cast<const unsigned *>(opbase + P.y_off);
  const int32_t *__restrict__ brow = a.brow + P.brow_off;
  const int32_t *__restrict__ ck = reinterpret_cast<const int32_t *>(a.dirs + P.dir_off);

  BackWriter0 w1, w2;
  w1.init(a.out + P.out_off, (uint32_t)(m + n), lane == 0);
  w2.init(a.out + P.out_off + P.cap, (uint32_t)(m + n), lane == 0);

  int i = m, j = n;
  int cost = 0;
  while (i > 0 && j > 0) {
    // ---- the tile containing (i, j), cut to the rows <= i and columns <= j
    const int hs = (i - 1) / TILE_ROWS, cb = (j - 1) / (4 * CKQ);
    const int row0 = hs * TILE_ROWS, col0 = cb * 4 * CKQ, q0 = cb * CKQ;
    const int tq = ((j - 1 - col0) >> 2) + 1;  // macro columns needed
    const int tmax = (i - 1 - row0) / RT;      // last lane needed
    int xc[RT], h[RT];
    {
      const int r0 = row0 + lane * RT;
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const bool inside = r0 + r < m;
        const int ch = inside ? (int)xop[r0 + r] : 0;
        xc[r] = row_operand<MODE>(inside, ch, r, A, B);
      }
      if (cb > 0) {  // left edge = column checkpoint cb-1
        const int4 *src = reinterpret_cast<const int4 *>(ck + (uint64_t)(cb - 1) * mpad + (uint64_t)r0);
        const int4 v0 = __ldcg(src), v1 = __ldcg(src + 1);
        h[0] = v0.x; h[1] = v0.y; h[2] = v0.z; h[3] = v0.w;
        h[4] = v1.x; h[5] = v1.y; h[6] = v1.z; h[7] = v1.w;
      } else {
#pragma unroll
        for (int r = 0; r < RT; ++r) h[r] = 0;
      }
    }
    {  // stage the tile's characters (8 per lane; rows / columns past the ends are never walked)
      const int r0 = row0 + lane * 8, c0 = col0 + lane * 8;
      schars[lane] = r0 < m ? *reinterpret_cast<const uint64_t *>(xs + r0) : 0ull;
      schars[32 + lane] = c0 < n ? *reinterpret_cast<const uint64_t *>(ys + c0) : 0ull;
    }
    const int32_t *toprow = hs > 0 ? brow + (uint64_t)(hs - 1) * bstride : nullptr;
    int dg = __shfl_up_sync(FULL, h[RT - 1], 1);
    if (lane == 0) dg = (toprow && cb > 0) ? __ldcg(toprow + col0 - 1) : 0;
    int outv[4] = {0, 0, 0, 0};
    unsigned yw = 0;
    int4 top = make_int4(0, 0, 0, 0);
    unsigned ynext = 0;
    const int4 *top4 = toprow ? reinterpret_cast<const int4 *>(toprow) + q0 : nullptr;
    if (lane == 0) {
      ynext = __ldg(yw32 + q0);
      if (top4) top = __ldcg(top4);
    }
    const int nsteps = tq + tmax;
    for (int q = 0; q < nsteps; ++q) {
      int up_in[4];
      up_in[0] = __shfl_up_sync(FULL, outv[0], 1);
      up_in[1] = __shfl_up_sync(FULL, outv[1], 1);
      up_in[2] = __shfl_up_sync(FULL, outv[2], 1);
      up_in[3] = __shfl_up_sync(FULL, outv[3], 1);
      unsigned ny = __shfl_up_sync(FULL, yw, 1);
      if (lane == 0) {
        up_in[0] = top.x; up_in[1] = top.y; up_in[2] = top.z; up_in[3] = top.w;
        ny = ynext;
        if (q + 1 < tq) {  // look-ahead for the next macro column
          ynext = __ldg(yw32 + q0 + q + 1);
          if (top4) top = __ldcg(top4 + q + 1);
        }
      }
      const int qc = q - lane;
      if (qc >= 0 && qc < tq && lane <= tmax) {
        yw = ny;
        unsigned bits[RT / 4];
        macro_step<RT, MODE, ADDMODE, true>(h, dg, xc, up_in, yw, A, B, 1, -1, outv, bits);
        sdirs[qc * 32 + lane] = make_uint2(bits[0], bits[1]);
      }
    }
    __syncwarp();
    // ---- walk inside the tile (every lane walks; lane 0 writes)
    int li = i - 1 - row0, lj = j - 1 - col0;
    while (li >= 0 && lj >= 0) {
      const unsigned t = (unsigned)li >> 3, r = (unsigned)li & 7u, qc = (unsigned)lj >> 2, c = (unsigned)lj & 3u;
      const unsigned idx = c * RT + r;
      const uint32_t word = sdirs32[(qc * 32 + t) * 2 + (idx >> 4)];
      const uint32_t code = (word >> (30 - 2 * (idx & 15u))) & 3u;
      const uint32_t xch = sxb[li], ych = syb[lj];
      if (!(code & DIR_NOT_DIAG)) {
        w1.put(xch);
        w2.put(ych);
        cost += (xch == ych) ? 0 : a.pxy;
        --li;
        --lj;
      } else if (!(code & DIR_LEFT_BETTER)) {
        w1.put(xch);
        w2.put('_');
        cost += a.pgap;
        --li;
      } else {
        w1.put('_');
        w2.put(ych);
        cost += a.pgap;
        --lj;
      }
    }
    i = row0 + li + 1;
    j = col0 + lj + 1;
    __syncwarp();
  }
  cost += (i + j) * a.pgap;
  for (; i > 0; --i) {
    w1.put(xs[i - 1]);
    w2.put('_');
  }
  for (; j > 0; --j) {
    w1.put('_');
    w2.put(ys[j - 1]);
  }
  const uint32_t start = w1.pos;
  w1.finish();
  w2.finish();
  if (lane == 0) {
    PairAlign res;
    res.start = start;
    res.len = (uint32_t)(m + n) - start;
    a.align[p] = res;
    a.path_cost[p] = cost;
  }
}

cudaError_t launch_traceback_tiles(const TracebackArgs &args, cudaStream_t stream) {
  if (args.npairs == 0) return cudaSuccess;
  const size_t smem = (size_t)CKQ * 32 * sizeof(uint2) + 2 * TILE_ROWS;
  switch (args.mode) {
    case MODE_GENERAL: traceback_tiles_kernel<MODE_GENERAL, 0><<<args.npairs, 32, smem, stream>>>(args); break;
    case MODE_CMP: traceback_tiles_kernel<MODE_CMP, 0><<<args.npairs, 32, smem, stream>>>(args); break;
    case MODE_TAB:
      if (args.addmode == 2)
        traceback_tiles_kernel<MODE_TAB, 2><<<args.npairs, 32, smem, stream>>>(args);
      else
        traceback_tiles_kernel<MODE_TAB, 0><<<args.npairs, 32, smem, stream>>>(args);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace seqalk
