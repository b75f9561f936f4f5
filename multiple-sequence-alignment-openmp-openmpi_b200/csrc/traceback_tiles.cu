// traceback_tiles.cu — traceback for the tile mode: directions are NOT read from HBM but
// recomputed, one TILE_ROWS-row x (4*CKQ)-column tile (128 x 128) at a time, only for the tiles the optimal path
// visits (about m / 128 + n / 128 of them), from the checkpoints the score-only fill left behind:
// one row of H' every TILE_ROWS DP rows and one column of H' every 4*CKQ DP columns.
//
// Same reference semantics as traceback.cu (submit/xuliny-seqalkway.cpp:502-558, :191-214):
// the recompute runs the very same per-cell update as the fill (fill_core.cuh, EMIT = true), so
// the 2-bit codes — and with them the tie-breaking — are identical to the directions mode.
//
// Half a warp per pair (jobs with many pairs; a whole warp + three look-ahead warps for jobs of a
// few long pairs).  Per tile: lane t owns RT rows, the lanes sweep the needed macro columns as a
// skewed wavefront (exactly like the fill) writing the codes to shared memory, then walk the path
// inside the tile until it leaves through the top or the left edge.
#include "fill_core.cuh"

namespace seqalk {

namespace {

// A tile is recomputed by a group of LN = TILE_ROWS / RT lanes owning RT rows each: the whole warp
// (RT = TILE_ROWS / 32: the four-warp kernel of long pairs) or half a warp (RT = TILE_ROWS / 16:
// the many-pair kernel, two pairs per warp — half the wavefront ramp per tile).
constexpr int RT_WARP = TILE_ROWS / 32;
constexpr int RT_HALF = TILE_ROWS / 16;
static_assert(RT_WARP == 4 && RT_HALF == 8, "tile height 128: 4 rows per lane (warp) / 8 rows per lane (half warp)");

// Right-to-left output of the two alignment strings.  Every lane of the walking warp carries
// the same position; the characters of 32 consecutive steps are stored by 32 lanes at once.
struct PathOut {
  uint8_t *o1, *o2;  // first character of each string's buffer
  int pos;           // characters [pos, m + n) are written
  int cost;          // this lane's share of the path cost
};

}  // namespace

// Per-pair constants of the tile traceback.
struct TileCtx {
  int m, n, nq;
  uint64_t bstride, mpad;
  int A, B;
  const uint8_t *xs, *ys;      // raw characters (output)
  const uint8_t *xop;          // row operands: letter codes (MODE_TAB) or raw characters
  const unsigned *yw32;        // y as 4-byte words (codes or raw)
  const int32_t *brow, *ck;    // kept rows, column checkpoints
  // 16x2-packed column checkpoints (fill16.cu), ck16 = log2 of the fill's re-base interval (0: int32
  // checkpoints): image (k, half) of mpad/2 packed words, relative to the base the fill logged
  int ck16, main_rows;
  int nstrips, half_last;
};

template <int MODE>
__device__ __forceinline__ TileCtx make_ctx(const TracebackArgs &a, const PairDesc &P) {
  TileCtx T;
  T.m = P.m;
  T.n = P.n;
  T.nq = (P.n + 3) >> 2;
  T.bstride = (uint64_t)(4 * T.nq + 4);
  T.mpad = (uint64_t)P.mpad;  // rows padded to the fill's strips
  T.A = -2 * a.pgap;
  T.B = a.pxy - 2 * a.pgap;
  T.xs = a.seq + P.x_off;
  T.ys = a.seq + P.y_off;
  const uint8_t *opbase = (MODE == MODE_TAB) ? a.codes : a.seq;
  T.xop = opbase + P.x_off;
  T.yw32 = reinterpret_cast<const unsigned *>(opbase + P.y_off);
  T.brow = a.brow + P.brow_off;
  T.ck = reinterpret_cast<const int32_t *>(a.dirs + P.dir_off);
  T.ck16 = a.ck16;
  T.main_rows = a.main_rows;
  T.nstrips = P.nstrips;
  T.half_last = (int)P.half_last;
  return T;
}

// Shared-memory image of one tile: direction codes [CKQ][32] + the tile's characters, and
// (STAGE) the tile's top edge: per macro column the 4 values of the kept row + the y word.
constexpr int TILE_DIR_BYTES = CKQ * TILE_ROWS;  // 2 bits per cell of a TILE_ROWS x 4*CKQ tile
constexpr int TILE_CHAR_BYTES = TILE_ROWS + 4 * CKQ;
constexpr int TILE_EDGE_BYTES = CKQ * 20;
constexpr int tile_smem_bytes(bool stage) { return TILE_DIR_BYTES + TILE_CHAR_BYTES + (stage ? TILE_EDGE_BYTES : 0); }
#ifndef SEQALK_TB_STAGE
#define SEQALK_TB_STAGE 1
#endif

// Recomputes the directions of tile (hs, cb), macro columns [0, tq) and lanes [0, tmax], into
// `tile`, by a group of LN lanes owning RT rows each (`lane` = index inside the group; every lane of
// the WARP must call this together, with the same `nsteps` >= tq + tmax: the groups of a warp run
// the step loop in lockstep, each on its own tile).  Same sweep as the fill: lane t owns RT rows,
// trails lane t-1 by one macro column; left edge = column checkpoint cb-1, top edge = kept row hs-1.
// tq = 0: nothing to do for this group (it only keeps the others company).
// STAGE: the group first copies the top edge to shared memory with a few wide loads; otherwise
// lane 0 fetches it from global memory one macro column ahead, which exposes the load latency on
// every step unless other warps cover it.
template <int MODE, int ADDMODE, bool STAGE, int RT, int LN = TILE_ROWS / RT>
__device__ __forceinline__ void recompute_tile(const TileCtx &T, int hs, int cb, int tq, int tmax, int nsteps,
                                               uint8_t *tile, int lane) {
  static_assert(LN * RT <= TILE_ROWS, "LN lanes x RT rows: the tile, or (LN * RT < TILE_ROWS) its upper part");
  constexpr int WPE = RT / 4;         // 32-bit words of direction codes per lane and macro column
  static_assert(RT == 8 || RT == 4, "rows per lane");
  const unsigned FULL = 0xffffffffu;
  uint32_t *sdirs = reinterpret_cast<uint32_t *>(tile);
  const int row0 = hs * TILE_ROWS, col0 = cb * 4 * CKQ, q0 = cb * CKQ;
  int xc[RT], h[RT];
  {
    const int r0 = row0 + lane * RT;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const bool inside = r0 + r < T.m;
      const int ch = inside ? (int)T.xop[r0 + r] : 0;
      xc[r] = row_operand<MODE>(inside, ch, r, T.A, T.B);
    }
    if (cb > 0 && T.ck16) {
      // left edge from the packed fill's checkpoint cb-1.  Row r0 lies in the strip pair starting
      // at row sp0, in its first half A (low
      // 16 bits) or second half B (high 16 bits), on virtual lane vl of 64; that lane reached the
      // checkpoint column (macro column 64*cb - 1) at macro step 64*cb - 1 + vl: the base logged
      // for that step makes the value absolute:  H' = -(base + relative).
      // (main strip pairs of main_rows rows; then, by the tail code, one of 1024 rows (three-level
      // geometry) and / or one of 512 rows)
      const int nmain = main_strips(T.nstrips, (uint32_t)T.half_last);
      const int main_end = nmain * T.main_rows;
      int sp_rows, s, sp0;
      if (r0 < main_end) {
        sp_rows = T.main_rows;
        s = r0 / T.main_rows;
        sp0 = s * T.main_rows;
      } else if ((T.half_last & 2) && r0 < main_end + 1024) {
        sp_rows = 1024;
        s = nmain;
        sp0 = main_end;
      } else {
        sp_rows = 512;
        s = T.nstrips - 1;
        sp0 = (int)T.mpad - 512;
      }
      const int rem = r0 - sp0;
      const int half = rem >= (sp_rows >> 1) ? 1 : 0;
      const int vl = rem / (sp_rows >> 6);
      const int step = CKQ * cb - 1 + vl;
      const int32_t *blog = T.brow + (uint64_t)(T.mpad / TILE_ROWS) * T.bstride +
                            (uint64_t)s * (uint64_t)base_log_entries(T.nq, T.ck16);
      const int base = __ldcg(blog + (step >> T.ck16));
      // CKQ = 64: images (checkpoint, half); CKQ = 32: image e = [A rows at checkpoint e | B rows at e - 1]
      const uint64_t image = CKQ == 64 ? (uint64_t)(cb - 1) * 2 + half : (uint64_t)(cb - 1) + half;
      const uint32_t *img = reinterpret_cast<const uint32_t *>(T.ck) + image * (T.mpad / 2);
      const uint4 *src = reinterpret_cast<const uint4 *>(img + (sp0 >> 1) + (rem - half * (sp_rows >> 1)));
#pragma unroll
      for (int r4 = 0; r4 < RT / 4; ++r4) {
        const uint4 v = __ldcg(src + r4);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) h[4 * r4 + r] = -(base + (int)(half ? (w[r] >> 16) : (w[r] & 0xffffu)));
      }
    } else if (cb > 0) {  // left edge = column checkpoint cb-1
      const int4 *src = reinterpret_cast<const int4 *>(T.ck + (uint64_t)(cb - 1) * T.mpad + (uint64_t)r0);
#pragma unroll
      for (int r4 = 0; r4 < RT / 4; ++r4) {
        const int4 v = __ldcg(src + r4);
        h[4 * r4] = v.x; h[4 * r4 + 1] = v.y; h[4 * r4 + 2] = v.z; h[4 * r4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < RT; ++r) h[r] = 0;
    }
  }
  if (tq > 0) {  // stage the tile's characters (rows / columns past the ends are never walked)
    const int r0 = row0 + lane * RT;
    if (RT == 8)
      reinterpret_cast<uint64_t *>(tile + TILE_DIR_BYTES)[lane] =
          r0 < T.m ? *reinterpret_cast<const uint64_t *>(T.xs + r0) : 0ull;
    else
      reinterpret_cast<uint32_t *>(tile + TILE_DIR_BYTES)[lane] =
          r0 < T.m ? *reinterpret_cast<const uint32_t *>(T.xs + r0) : 0u;
#pragma unroll
    for (int c = lane * 8; c < 4 * CKQ; c += 8 * LN)  // 8 y characters per lane and round
      *reinterpret_cast<uint64_t *>(tile + TILE_DIR_BYTES + TILE_ROWS + c) =
          col0 + c < T.n ? *reinterpret_cast<const uint64_t *>(T.ys + col0 + c) : 0ull;
  }
  const int32_t *toprow = hs > 0 ? T.brow + (uint64_t)(hs - 1) * T.bstride : nullptr;
  int dg = __shfl_up_sync(FULL, h[RT - 1], 1, LN);
  if (lane == 0) dg = (toprow && cb > 0) ? __ldcg(toprow + col0 - 1) : 0;
  int outv[4] = {0, 0, 0, 0};
  unsigned yw = 0;
  int4 top = make_int4(0, 0, 0, 0);
  unsigned ynext = 0;
  const int4 *top4 = toprow ? reinterpret_cast<const int4 *>(toprow) + q0 : nullptr;
  int4 *stop = reinterpret_cast<int4 *>(tile + TILE_DIR_BYTES + TILE_CHAR_BYTES);
  unsigned *sy = reinterpret_cast<unsigned *>(stop + CKQ);
  if (STAGE) {
    for (int q = lane; q < tq; q += LN) {
      stop[q] = top4 ? __ldcg(top4 + q) : make_int4(0, 0, 0, 0);
      sy[q] = __ldg(T.yw32 + q0 + q);
    }
    __syncwarp();
  } else if (lane == 0 && tq > 0) {
    ynext = __ldg(T.yw32 + q0);
    if (top4) top = __ldcg(top4);
  }
  for (int q = 0; q < nsteps; ++q) {
    int up_in[4];
    up_in[0] = __shfl_up_sync(FULL, outv[0], 1, LN);
    up_in[1] = __shfl_up_sync(FULL, outv[1], 1, LN);
    up_in[2] = __shfl_up_sync(FULL, outv[2], 1, LN);
    up_in[3] = __shfl_up_sync(FULL, outv[3], 1, LN);
    unsigned ny = __shfl_up_sync(FULL, yw, 1, LN);
    if (STAGE) {
      const int qs = max(min(q, tq - 1), 0);  // broadcast reads; only lane 0 uses them, and only while q < tq
      top = stop[qs];
      ynext = sy[qs];
    }
    if (lane == 0) {
      up_in[0] = top.x; up_in[1] = top.y; up_in[2] = top.z; up_in[3] = top.w;
      ny = ynext;
      if (!STAGE && q + 1 < tq) {  // look-ahead for the next macro column
        ynext = __ldg(T.yw32 + q0 + q + 1);
        if (top4) top = __ldcg(top4 + q + 1);
      }
    }
    const int qc = q - lane;
    if (qc >= 0 && qc < tq && lane <= tmax) {
      yw = ny;
      unsigned bits[RT / 4];
      macro_step<RT, MODE, ADDMODE, true>(h, dg, xc, up_in, yw, T.A, T.B, 1, -1, outv, bits);
      if (RT == 8)
        reinterpret_cast<uint2 *>(sdirs)[qc * LN + lane] = make_uint2(bits[0], bits[WPE - 1]);
      else
        sdirs[qc * LN + lane] = bits[0];
    }
  }
}

// Emits the steps recorded in `hist` (a sentinel 1 followed by one 2-bit code per step, oldest
// first; at most 15 steps), the walk having been at tile cell (hli, hlj) before the first of
// them: lane s rebuilds the cell of step s from the codes before it (x is consumed by every
// step that is not LEFT, y by every step that is not UP) and stores its two characters.
__device__ __forceinline__ void emit_steps(const uint8_t *sxb, const uint8_t *syb, uint32_t hist, int hli, int hlj,
                                           PathOut &W, int pxy, int pgap, int lane) {
  const int n = (31 - __clz(hist)) >> 1;
  if (lane < n) {
    const int sh = 2 * (n - 1 - lane);
    const uint32_t code = (hist >> sh) & 3u;  // 0, 1: diagonal; 2: up; 3: left
    const uint32_t hi = (hist >> 1) & 0x55555555u, lo = hist & 0x55555555u;
    const uint32_t prior = 0xFFFFFFFCu << sh;  // the steps before this one
    const int li = hli - (lane - __popc(hi & lo & prior)), lj = hlj - (lane - __popc(hi & ~lo & prior));
    const bool dx = code != (DIR_NOT_DIAG | DIR_LEFT_BETTER), dy = code != DIR_NOT_DIAG;
    const uint32_t xch = sxb[li], ych = syb[lj];
    const int at = W.pos - 1 - lane;
    W.o1[at] = (uint8_t)(dx ? xch : (uint32_t)'_');
    W.o2[at] = (uint8_t)(dy ? ych : (uint32_t)'_');
    W.cost += (dx && dy) ? (xch == ych ? 0 : pxy) : pgap;
  }
  W.pos -= n;
}

// Walks the path inside a recomputed tile from (i, j) until it leaves through the top or the
// left edge.  Every lane of the group (`lane` = index inside it, at least 16 lanes) follows the walk.  One shared-memory word holds the 16 codes of an
// RT x (16 / RT) block of cells; while the path stays inside the block the word is kept rotated
// so that the current cell's code sits in the top two bits (a step is: read the code, rotate by
// a table look-up, move, append the code to the history).  The history is emitted by up to 15
// lanes at once (emit_steps).
template <int RT, int LN = TILE_ROWS / RT>  // LN lanes x RT rows recomputed the tile (layout of the code words)
__device__ __forceinline__ void walk_tile(const uint8_t *tile, int row0, int col0, int &i, int &j, PathOut &W,
                                          int pxy, int pgap, int lane) {
  constexpr int WPE = RT / 4;
  constexpr int CW = 16 / RT;         // columns of a word's block
  constexpr int MAXB = RT + CW - 1;   // most steps inside one block
  // rotate amounts (2 bits per cell, cell index = column * RT + row) by code
  constexpr uint32_t ROT = (2u * (RT + 1)) | (2u * (RT + 1)) << 8 | 2u << 16 | (2u * RT) << 24;
  const uint32_t *sdirs32 = reinterpret_cast<const uint32_t *>(tile);
  const uint8_t *sxb = tile + TILE_DIR_BYTES;
  const uint8_t *syb = sxb + TILE_ROWS;
  int li = i - 1 - row0, lj = j - 1 - col0;
  uint32_t hist = 1u;
  int hli = li, hlj = lj;
  while (li >= 0 && lj >= 0) {
    if (hist >= (1u << (2 * (16 - MAXB)))) {  // the next block might not fit
      emit_steps(sxb, syb, hist, hli, hlj, W, pxy, pgap, lane);
      hist = 1u;
      hli = li;
      hlj = lj;
    }
    const int bl = li & ~(RT - 1), bj = lj & ~(CW - 1);
    const unsigned idx = (unsigned)(lj & (CW - 1)) * RT + (unsigned)(li & (RT - 1));
    const uint32_t word = sdirs32[(((unsigned)lj >> 2) * LN + (unsigned)li / RT) * WPE + ((unsigned)(lj & 3) / CW)];
    uint32_t w = __funnelshift_l(word, word, 2 * idx);
#pragma unroll 2
    do {
      const uint32_t code = w >> 30;
      w = __funnelshift_r(w, w, __byte_perm(ROT, 0u, code));  // shift amount is taken mod 32
      li -= (code != (DIR_NOT_DIAG | DIR_LEFT_BETTER));
      lj -= (code != DIR_NOT_DIAG);
      hist = hist * 4u + code;
    } while ((((li ^ bl) & ~(RT - 1)) | ((lj ^ bj) & ~(CW - 1))) == 0);
  }
  if (hist > 1u) emit_steps(sxb, syb, hist, hli, hlj, W, pxy, pgap, lane);
  i = row0 + li + 1;
  j = col0 + lj + 1;
}

// Tails (only one of i, j is non-zero) and the per-pair result record, by a group of LN lanes
// (`lane` = index inside the group, `gmask` = the group's lanes in the warp, `g0` = its first lane).
template <int LN>
__device__ __forceinline__ void finish_pair(const TracebackArgs &a, const TileCtx &T, uint32_t p, int i, int j,
                                            PathOut &W, int lane, unsigned gmask, int g0) {
  for (int t = lane; t < i; t += LN) {
    W.o1[W.pos - 1 - t] = T.xs[i - 1 - t];
    W.o2[W.pos - 1 - t] = '_';
  }
  W.pos -= i;
  for (int t = lane; t < j; t += LN) {
    W.o1[W.pos - 1 - t] = '_';
    W.o2[W.pos - 1 - t] = T.ys[j - 1 - t];
  }
  W.pos -= j;
  if (a.strip_scan) {
    // '_' inside the inputs: the reference strips everything up to the LAST column that holds
    // '_' in both strings (submit/xuliny-seqalkway.cpp:191-202), not only the leading pad
    __syncwarp(gmask);
    const int L = T.m + T.n;
    int found = -1;
    for (int base = L - LN; base + LN - 1 >= W.pos && found < 0; base -= LN) {
      const int idx = base + lane;
      const bool hit = idx >= W.pos && idx < L && W.o1[idx] == '_' && W.o2[idx] == '_';
      const unsigned b = __ballot_sync(gmask, hit) >> g0;
      if (b) found = base + 31 - __clz(b);
    }
    if (found >= 0) W.pos = found + 1;
  }
  int cost = W.cost;
#pragma unroll
  for (int d = LN / 2; d > 0; d >>= 1) cost += __shfl_xor_sync(gmask, cost, d);
  if (lane == 0) {
    PairAlign res;
    res.start = (uint32_t)W.pos;
    res.len = (uint32_t)(T.m + T.n - W.pos);
    a.align[p] = res;
    a.path_cost[p] = cost + (i + j) * a.pgap;
  }
}

// One warp per pair: recompute the part of the tile up-left of the entry cell, walk it, repeat.
// (SEQALK_TB_HALF=0; the default for jobs with many pairs is the half-warp kernel below.)
template <int MODE, int ADDMODE, bool STAGE>
__global__ void __launch_bounds__(32) traceback_tiles_kernel(const TracebackArgs a) {
  constexpr int RT = RT_WARP;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x;
  const uint32_t p = blockIdx.x;
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, T.m + T.n, 0};
  int i = T.m, j = T.n;
  while (i > 0 && j > 0) {
    const int hs = (i - 1) / TILE_ROWS, cb = (j - 1) / (4 * CKQ);
    const int row0 = hs * TILE_ROWS, col0 = cb * 4 * CKQ;
    const int tq = ((j - 1 - col0) >> 2) + 1;  // macro columns needed
    const int tmax = (i - 1 - row0) / RT;      // last lane needed
    recompute_tile<MODE, ADDMODE, STAGE, RT>(T, hs, cb, tq, tmax, tq + tmax, smem, lane);
    __syncwarp();
    walk_tile<RT>(smem, row0, col0, i, j, W, a.pxy, a.pgap, lane);
    __syncwarp();
  }
  finish_pair<32>(a, T, p, i, j, W, lane, 0xffffffffu, 0);
}

// Two pairs per warp, one per half warp (jobs with many pairs): 16 lanes x 8 rows per tile.  The
// wavefront ramp of a tile costs 15 steps instead of 31, and both halves share every issued
// instruction of the recompute (the step loop runs in lockstep over max(steps) of the two tiles;
// neighbouring pairs of a batch have similar lengths: they are laid out longest first).  The walks
// of the two halves run together as long as their control flow agrees.
// Measured (one B200, same box): cfg4 traceback 811 ms (one warp per pair, 128 x 256 tiles) ->
// 743 ms (128 x 128 tiles) -> 608 ms (this kernel, 22 CTAs per SM); with 128 x 256 tiles this kernel
// was SLOWER (870 ms): 11 CTAs per SM by shared memory, issue slots 64 % busy instead of 82 %.
constexpr int HALF_TILE_BYTES = (tile_smem_bytes(true) + 15) & ~15;
template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(32) traceback_tiles_half_kernel(const TracebackArgs a) {
  constexpr int RT = RT_HALF, LN = 16;
  extern __shared__ __align__(16) uint8_t smem[];
  const int half = threadIdx.x >> 4, lane = threadIdx.x & 15;
  const uint32_t p0 = 2u * blockIdx.x;
  const bool valid = p0 + half < a.npairs;
  const uint32_t p = valid ? p0 + half : p0;
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  uint8_t *tile = smem + half * HALF_TILE_BYTES;
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, T.m + T.n, 0};
  int i = valid ? T.m : 0, j = valid ? T.n : 0;
  for (;;) {
    const bool active = i > 0 && j > 0;
    if (!__any_sync(0xffffffffu, active)) break;
    int hs = 0, cb = 0, tq = 0, li = -1;
    if (active) {
      hs = (i - 1) / TILE_ROWS;
      cb = (j - 1) / (4 * CKQ);
      tq = ((j - 1 - cb * 4 * CKQ) >> 2) + 1;
      li = i - 1 - hs * TILE_ROWS;  // row of the entry cell inside the tile
    }
    // both tiles entered in their upper half: 4 rows per lane keep twice as many lanes busy
    if (__all_sync(0xffffffffu, li < TILE_ROWS / 2)) {
      constexpr int RS = RT / 2;
      const int tmax = active ? li / RS : -1;
      const int mine = active ? tq + tmax : 0;
      const int nsteps = max(mine, __shfl_xor_sync(0xffffffffu, mine, 16));
      recompute_tile<MODE, ADDMODE, true, RS, LN>(T, hs, cb, tq, tmax, nsteps, tile, lane);
      __syncwarp();
      if (active) walk_tile<RS, LN>(tile, hs * TILE_ROWS, cb * 4 * CKQ, i, j, W, a.pxy, a.pgap, lane);
    } else {
      const int tmax = active ? li / RT : -1;
      const int mine = active ? tq + tmax : 0;
      const int nsteps = max(mine, __shfl_xor_sync(0xffffffffu, mine, 16));
      recompute_tile<MODE, ADDMODE, true, RT, LN>(T, hs, cb, tq, tmax, nsteps, tile, lane);
      __syncwarp();
      if (active) walk_tile<RT, LN>(tile, hs * TILE_ROWS, cb * 4 * CKQ, i, j, W, a.pxy, a.pgap, lane);
    }
    __syncwarp();
  }
  if (valid) finish_pair<LN>(a, T, p, i, j, W, lane, 0xffffu << (16 * half), 16 * half);
}

// Four warps per pair (jobs of a few long pairs, where one warp per pair leaves the GPU idle and
// the tile chain is pure latency): while warp 0 walks the current tile, warps 1..3 recompute —
// in full — the up to three tiles the path can enter next (above, left, above-left), so the next
// tile is already in shared memory when the walk leaves the current one.
constexpr int LONG_SLOTS = 4;
constexpr int LONG_TILE_BYTES = tile_smem_bytes(true);
template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(128) traceback_tiles_long_kernel(const TracebackArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_i, s_j;
  __shared__ int s_tag[LONG_SLOTS][2];  // (hs, cb) held by each slot, hs = -1: empty
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t p = blockIdx.x;
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, T.m + T.n, 0};
  if (threadIdx.x == 0) {
    s_i = T.m;
    s_j = T.n;
    for (int s = 0; s < LONG_SLOTS; ++s) s_tag[s][0] = -1;
  }
  __syncthreads();
  auto full_tq = [&](int cb) { return min(CKQ, T.nq - cb * CKQ); };
  constexpr int RT = RT_WARP;
  auto full_tmax = [&](int hs) { return min(31, (T.m - 1 - hs * TILE_ROWS) / RT); };
  // first tile: warp 1 computes it in full
  {
    const int hs = (T.m - 1) / TILE_ROWS, cb = (T.n - 1) / (4 * CKQ);
    if (warp == 1) {
      recompute_tile<MODE, ADDMODE, true, RT>(T, hs, cb, full_tq(cb), full_tmax(hs), full_tq(cb) + full_tmax(hs), smem, lane);
      if (lane == 0) {
        s_tag[0][0] = hs;
        s_tag[0][1] = cb;
      }
    }
  }
  __syncthreads();
  for (;;) {
    const int i = s_i, j = s_j;
    if (i <= 0 || j <= 0) break;
    const int hs = (i - 1) / TILE_ROWS, cb = (j - 1) / (4 * CKQ);
    // slot of the current tile (always present: it was a neighbour of the previous one)
    int cur = 0;
    for (int s = 0; s < LONG_SLOTS; ++s)
      if (s_tag[s][0] == hs && s_tag[s][1] == cb) cur = s;
    // the three tiles the path can enter next; producer warp w handles candidate w-1
    const int cand_hs[3] = {hs - 1, hs, hs - 1}, cand_cb[3] = {cb, cb - 1, cb - 1};
    // slots free for new tiles = those holding neither the current tile nor a candidate
    int target[3] = {-1, -1, -1};
    {
      bool used[LONG_SLOTS];
      for (int s = 0; s < LONG_SLOTS; ++s) {
        used[s] = (s == cur);
        for (int c = 0; c < 3; ++c)
          if (cand_hs[c] >= 0 && cand_cb[c] >= 0 && s_tag[s][0] == cand_hs[c] && s_tag[s][1] == cand_cb[c])
            used[s] = true;
      }
      for (int c = 0; c < 3; ++c) {
        if (cand_hs[c] < 0 || cand_cb[c] < 0) continue;
        bool have = false;
        for (int s = 0; s < LONG_SLOTS; ++s)
          if (s_tag[s][0] == cand_hs[c] && s_tag[s][1] == cand_cb[c]) have = true;
        if (have) continue;
        for (int s = 0; s < LONG_SLOTS; ++s)
          if (!used[s]) {
            used[s] = true;
            target[c] = s;
            break;
          }
      }
    }
    __syncthreads();  // everybody has read the tags before they change
    if (warp == 0) {
      int wi = i, wj = j;
      walk_tile<RT>(smem + (size_t)cur * LONG_TILE_BYTES, hs * TILE_ROWS, cb * 4 * CKQ, wi, wj, W, a.pxy, a.pgap,
                    lane);
      if (lane == 0) {
        s_i = wi;
        s_j = wj;
      }
    } else {
      const int c = warp - 1;
      if (target[c] >= 0) {
        recompute_tile<MODE, ADDMODE, true, RT>(T, cand_hs[c], cand_cb[c], full_tq(cand_cb[c]), full_tmax(cand_hs[c]),
                                                full_tq(cand_cb[c]) + full_tmax(cand_hs[c]),
                                                smem + (size_t)target[c] * LONG_TILE_BYTES, lane);
        if (lane == 0) {
          s_tag[target[c]][0] = cand_hs[c];
          s_tag[target[c]][1] = cand_cb[c];
        }
      }
    }
    __syncthreads();
  }
  if (warp == 0) finish_pair<32>(a, T, p, s_i, s_j, W, lane, 0xffffffffu, 0);
}

template <int MODE, int ADDMODE>
static cudaError_t launch_tiles(const TracebackArgs &args, cudaStream_t stream) {
  if (args.long_pairs) {
    const size_t smem = (size_t)LONG_SLOTS * LONG_TILE_BYTES;
    cudaError_t e = cudaFuncSetAttribute(traceback_tiles_long_kernel<MODE, ADDMODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    traceback_tiles_long_kernel<MODE, ADDMODE><<<args.npairs, 128, smem, stream>>>(args);
  } else if (args.half_warp) {
    traceback_tiles_half_kernel<MODE, ADDMODE><<<(args.npairs + 1) / 2, 32, 2 * HALF_TILE_BYTES, stream>>>(args);
  } else {
    constexpr bool STAGE = SEQALK_TB_STAGE != 0;
    traceback_tiles_kernel<MODE, ADDMODE, STAGE><<<args.npairs, 32, tile_smem_bytes(STAGE), stream>>>(args);
  }
  return cudaGetLastError();
}

cudaError_t launch_traceback_tiles(const TracebackArgs &args, cudaStream_t stream) {
  if (args.npairs == 0) return cudaSuccess;
  switch (args.mode) {
    case MODE_GENERAL: return launch_tiles<MODE_GENERAL, 0>(args, stream);
    case MODE_CMP: return launch_tiles<MODE_CMP, 0>(args, stream);
    case MODE_TAB: return args.addmode == 2 ? launch_tiles<MODE_TAB, 2>(args, stream) : launch_tiles<MODE_TAB, 0>(args, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace seqalk
