// traceback_tiles.cu — traceback for the tile mode: directions are NOT read from HBM but
// recomputed, one TILE_ROWS-row x (4*CKQ)-column tile (128 x 128) at a time, only for the tiles the optimal path
// visits (about m / 128 + n / 128 of them), from the checkpoints the score-only fill left behind:
// one row of H' every TILE_ROWS DP rows and one column of H' every 4*CKQ DP columns.
//
// Same reference semantics as traceback.cu (submit/xuliny-seqalkway.cpp:502-558, :191-214):
// the recompute runs the very same per-cell update as the fill (fill_core.cuh, EMIT = true), so
// the 2-bit codes — and with them the tie-breaking — are identical to the directions mode.
//
// Half a warp per pair (jobs with many pairs).  Per tile: lane t owns RT rows, the lanes sweep the
// needed macro columns as a skewed wavefront (exactly like the fill) writing the codes to shared
// memory, then walk the path inside the tile until it leaves through the top or the left edge.
// Jobs of a few LONG pairs run the banded traceback in chapters instead (further down): all tiles
// near the path recomputed at once together with their transfer tables, one table look-up per tile
// to follow the path, all visited tiles walked at once; shorter few-pair jobs keep the four-warp kernel
// (walk + three look-ahead recompute warps).
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

// Stages the tile's characters behind its direction codes (rows / columns past the ends are never
// walked): LN lanes, lane t brings the RT characters of its rows and 8 characters of y per round.
template <int RT, int LN>
__device__ __forceinline__ void stage_tile_chars(const TileCtx &T, int row0, int col0, uint8_t *tile, int lane) {
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

// macro_step (fill_core.cuh, EMIT = true) plus the tile's TRANSFER FUNCTION: next to H' every cell
// carries E = (exit << 16 | steps): where the traceback path through this cell leaves the tile and
// how many cells of the tile it visits on the way, E(cell) = E(predecessor of the cell) + 1 — the
// predecessor chosen exactly as walk_tile reads the 2-bit code (0, 1: diagonal; 2: up; 3: left).
// Cells outside the tile carry their own exit code and 0 steps: exit = c + 1 for the row above the
// tile at tile column c (c = -1: the corner), exit = 129 + r for the column left of it at tile row r.
// Same arithmetic as macro_step, statement by statement (the direction bits must be identical).
template <int R, int MODE, int ADDMODE>
__device__ __forceinline__ void macro_step_track(int (&h)[R], int &dg, const int (&xc)[R], const int (&up_in)[4],
                                                 unsigned yword, int A, int B, int (&outv)[4], unsigned (&bits)[R / 4],
                                                 int (&eh)[R], int &edg, const int (&eup_in)[4], int (&eout)[4]) {
  static_assert(ADDMODE != 1, "the traceback runs the plain or the IDP.4A adds");
#pragma unroll
  for (int w = 0; w < R / 4; ++w) bits[w] = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int yc = 0;
    unsigned sel = 0;
    if (MODE == MODE_TAB && ADDMODE == 2) {
      sel = __funnelshift_l(0u, 1u, yword >> (8 * c));
    } else {
      yc = (int)((yword >> (8 * c)) & 0xffu);
      if (MODE == MODE_TAB) sel = (unsigned)(yc >> 3) * 0x1111u + 0x8880u;
    }
    int diag = dg, up = up_in[c];
    dg = up;
    int ediag = edg, eup = eup_in[c];
    edg = eup;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int w = (c * R + r) / CELLS_PER_WORD;
      const int left = h[r], eleft = eh[r];
      int v;
      bool not_diag, left_better;
      if (MODE == MODE_GENERAL) {
        const bool match = (xc[r] == yc);
        const int t = diag + B;
        const int m1 = min(up, left);
        v = match ? diag + A : min(t, m1);
        not_diag = (!match) && (t > m1);
        left_better = up > left;
        bits[w] = (bits[w] << 2) | (not_diag ? 2u : 0u) | (left_better ? 1u : 0u);
      } else {
        int t;
        if (MODE == MODE_TAB && ADDMODE == 2) {
          t = __dp4a(xc[r], (int)sel, diag);
        } else if (MODE == MODE_TAB) {
          int cost;
          asm("prmt.b32 %0, %1, %2, %3;" : "=r"(cost) : "r"(xc[r]), "r"(0), "r"(sel));
          t = diag + cost;
        } else {
          t = diag + B;
          asm("{.reg .pred p; setp.eq.s32 p, %1, %2; @p add.s32 %0, %3, %4;}" : "+r"(t) : "r"(xc[r]), "r"(yc), "r"(diag), "r"(A));
        }
        v = min(min(up, left), t);
        const int d1 = v - t, d2 = left - up;
        bits[w] = __funnelshift_l((unsigned)d1, bits[w], 1);
        bits[w] = __funnelshift_l((unsigned)d2, bits[w], 1);
        not_diag = d1 < 0;
        left_better = d2 < 0;
      }
      const int e = (not_diag ? (left_better ? eleft : eup) : ediag) + 1;
      diag = left;
      up = v;
      h[r] = v;
      ediag = eleft;
      eup = e;
      eh[r] = e;
    }
    outv[c] = up;
    eout[c] = eup;
  }
}

// Entries of a tile's transfer table (256 words): the cells through which a path can enter the tile —
// bottom row (index = tile column) and right column (index = 128 + tile row).
constexpr int TILE_TABLE_WORDS = TILE_ROWS + 4 * CKQ;

// Recomputes the directions of tile (hs, cb), macro columns [0, tq) and lanes [0, tmax], into
// `tile`, by a group of LN lanes owning RT rows each (`lane` = index inside the group; every lane of
// the WARP must call this together, with the same `nsteps` >= tq + tmax: the groups of a warp run
// the step loop in lockstep, each on its own tile).  Same sweep as the fill: lane t owns RT rows,
// trails lane t-1 by one macro column; left edge = column checkpoint cb-1, top edge = kept row hs-1.
// tq = 0: nothing to do for this group (it only keeps the others company).
// STAGE: the group first copies the top edge to shared memory with a few wide loads; otherwise
// lane 0 fetches it from global memory one macro column ahead, which exposes the load latency on
// every step unless other warps cover it.
// TRACK: also writes the tile's transfer table (macro_step_track) to `table`: E of the bottom row's cells
// (full-height tiles) and of the right column's cells (full-width tiles).
template <int MODE, int ADDMODE, bool STAGE, int RT, int LN = TILE_ROWS / RT, bool TRACK = false>
__device__ __forceinline__ void recompute_tile(const TileCtx &T, int hs, int cb, int tq, int tmax, int nsteps,
                                               uint8_t *tile, int lane, uint32_t *table = nullptr) {
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
  if (tq > 0) stage_tile_chars<RT, LN>(T, row0, col0, tile, lane);
  const int32_t *toprow = hs > 0 ? T.brow + (uint64_t)(hs - 1) * T.bstride : nullptr;
  int dg = __shfl_up_sync(FULL, h[RT - 1], 1, LN);
  if (lane == 0) dg = (toprow && cb > 0) ? __ldcg(toprow + col0 - 1) : 0;
  int eh[RT], edg = 0, eout[4] = {0, 0, 0, 0};
  if (TRACK) {
    static_assert(!TRACK || LN * RT == TILE_ROWS, "transfer tables: whole tiles");
#pragma unroll
    for (int r = 0; r < RT; ++r) eh[r] = (4 * CKQ + 1 + lane * RT + r) << 16;  // left of the tile: exit through the left edge at this row
    edg = __shfl_up_sync(FULL, eh[RT - 1], 1, LN);
    if (lane == 0) edg = 0;  // the corner
  }
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
    int eup_in[4] = {0, 0, 0, 0};
    if (TRACK) {
#pragma unroll
      for (int c = 0; c < 4; ++c) eup_in[c] = __shfl_up_sync(FULL, eout[c], 1, LN);
      if (lane == 0) {  // above the tile: exit through the top edge at this column
#pragma unroll
        for (int c = 0; c < 4; ++c) eup_in[c] = (4 * q + c + 1) << 16;
      }
    }
    if (qc >= 0 && qc < tq && lane <= tmax) {
      yw = ny;
      unsigned bits[RT / 4];
      if (TRACK) {
        macro_step_track<RT, MODE, ADDMODE>(h, dg, xc, up_in, yw, T.A, T.B, outv, bits, eh, edg, eup_in, eout);
        if (lane == LN - 1) reinterpret_cast<int4 *>(table)[qc] = make_int4(eout[0], eout[1], eout[2], eout[3]);
      } else {
        macro_step<RT, MODE, ADDMODE, true>(h, dg, xc, up_in, yw, T.A, T.B, 1, -1, outv, bits);
      }
      if (RT == 8)
        reinterpret_cast<uint2 *>(sdirs)[qc * LN + lane] = make_uint2(bits[0], bits[WPE - 1]);
      else
        sdirs[qc * LN + lane] = bits[0];
    }
  }
  if (TRACK && tq == CKQ && lane <= tmax) {
    static_assert(!TRACK || RT == 4, "one int4 of right-column entries per lane");
    reinterpret_cast<int4 *>(table + 4 * CKQ)[lane] = make_int4(eh[0], eh[1], eh[2], eh[3]);
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

// ---- Banded traceback in chapters (jobs of a few LONG pairs).
//
// The four-warp kernel above is a serial chain of tile recomputes: ~12 us per tile on a lone warp
// (63 dependent macro steps), ~3 100 tiles for a 200k x 200k pair, and even with the tiles at hand the
// walk itself costs ~45 ns per step on a lone warp.  Neither the recompute of a tile nor the walk
// inside it depends on the rest of the path — only WHERE the path enters the tile does.  So the
// traceback runs in CHAPTERS of up to band_levels + 1 tile anti-diagonals (levels hs + cb), three
// launches each:
//   PRODUCE  one warp per tile, all tiles at once: every tile within band_halfwidth of the line the
//            path is expected to follow is recomputed; its 4 KB direction image and its transfer table
//            (entry cell on the bottom row / right column -> exit cell + steps, macro_step_track) go to HBM;
//   HOP      one warp per pair, serial but short: from the chapter's first cell the path is followed
//            tile by tile with ONE table look-up per tile, which yields every visited tile's entry cell
//            and the position of its characters in the alignment strings;
//   EMIT     one warp per visited tile, all at once: copies the image to shared memory, walks the tile
//            from its entry cell and writes the characters (walk_tile, as in every other kernel).
// The band follows the path: a chapter is centred on the line through its first tile with the slope of
// the previous chapter's displacement (the first: towards DP cell (0, 0)); when the path leaves the band
// the chapter ends there and the next one is re-centred (paths of unequal-length pairs are long
// straight stretches — diagonal runs and gap runs — with few turns).  A tile that still is not covered
// (a pair's first tile, entered in its interior; everything in the last chapter the host enqueued) is
// recomputed and walked by the hopping warp itself (BandState::misses), so the result never depends
// on the band.  A FINISH launch writes the tails and the result records.
struct BandGeom {
  int hs0, cb0, L0, W1;
  float slope;  // of the centre line: tile rows per level (the produce and the hop launches evaluate the same
                // expression on the same inputs, so they agree on every tile's slot)
};
__device__ __forceinline__ BandGeom band_geom(const TracebackArgs &a, int i0, int j0, int dhs, int dcb) {
  BandGeom g;
  g.hs0 = (i0 - 1) / TILE_ROWS;
  g.cb0 = (j0 - 1) / (4 * CKQ);
  g.L0 = g.hs0 + g.cb0;
  g.W1 = 2 * a.band_halfwidth + 1;
  if (dhs + dcb > 0)  // the previous chapter's displacement
    g.slope = __fdividef((float)dhs, (float)(dhs + dcb));
  else  // towards DP cell (0, 0)
    g.slope = __fdividef((float)i0, (float)i0 + (float)j0);
  return g;
}
// tile row of the band's centre, d levels below the chapter's first tile
__device__ __forceinline__ int band_center(const BandGeom &g, int d) {
  return g.hs0 - (int)__fmaf_rn((float)d, g.slope, 0.5f);
}
// index of tile (hs, cb) among the chapter's images, -1: not part of the band
__device__ __forceinline__ int band_slot(const TracebackArgs &a, const BandGeom &g, int hs, int cb) {
  const int d = g.L0 - (hs + cb);
  if (hs < 0 || cb < 0 || hs > g.hs0 || cb > g.cb0 || d < 0 || d > a.band_levels) return -1;
  const int k = hs - band_center(g, d) + a.band_halfwidth;
  if (k < 0 || k >= g.W1) return -1;
  return d * g.W1 + k;
}
// a pair's state at the start of chapter `chapter`
struct BandStart {
  int i, j, dhs, dcb;
};
__device__ __forceinline__ BandStart band_start(const TracebackArgs &a, const PairDesc &P, uint32_t p, int chapter) {
  BandStart b{P.m, P.n, 0, 0};
  if (chapter > 0) {
    const BandState *S = a.band_state + p;
    b.i = S->done ? 0 : S->i;
    b.j = S->done ? 0 : S->j;
    b.dhs = S->dhs;
    b.dcb = S->dcb;
  }
  return b;
}
__device__ __forceinline__ int band_images(const TracebackArgs &a) { return (a.band_levels + 1) * (2 * a.band_halfwidth + 1); }
__device__ __forceinline__ const uint8_t *band_image(const TracebackArgs &a, uint32_t p, int slot) {
  return a.band_tiles + ((uint64_t)p * (uint64_t)band_images(a) + (uint64_t)slot) * TILE_DIR_BYTES;
}
__device__ __forceinline__ uint32_t *band_table(const TracebackArgs &a, uint32_t p, int slot) {
  return a.band_tables + ((uint64_t)p * (uint64_t)band_images(a) + (uint64_t)slot) * TILE_TABLE_WORDS;
}

constexpr int BAND_WARPS = 4;  // warps per CTA of the produce and emit kernels
template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(32 * BAND_WARPS) traceback_band_produce_kernel(const TracebackArgs a, const int chapter) {
  constexpr int RT = RT_WARP;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t p = blockIdx.y;
  const PairDesc P = a.pairs[p];
  const BandStart b = band_start(a, P, p, chapter);
  if (b.i <= 0 || b.j <= 0) return;  // finished (or nothing to walk)
  const BandGeom g = band_geom(a, b.i, b.j, b.dhs, b.dcb);
  const int slot = (int)blockIdx.x * BAND_WARPS + warp;
  if (slot >= band_images(a)) return;
  const int d = slot / g.W1, k = slot - d * g.W1;
  const int hs = band_center(g, d) + k - a.band_halfwidth, cb = g.L0 - d - hs;
  if (hs < 0 || cb < 0 || hs > g.hs0 || cb > g.cb0) return;
  const TileCtx T = make_ctx<MODE>(a, P);
  uint8_t *tile = smem + (size_t)warp * LONG_TILE_BYTES;
  const int tq = min(CKQ, T.nq - cb * CKQ), tmax = min(31, (T.m - 1 - hs * TILE_ROWS) / RT);
  recompute_tile<MODE, ADDMODE, true, RT, TILE_ROWS / RT, true>(T, hs, cb, tq, tmax, tq + tmax, tile, lane,
                                                                band_table(a, p, slot));
  __syncwarp();
  uint4 *dst = reinterpret_cast<uint4 *>(const_cast<uint8_t *>(band_image(a, p, slot)));
  const uint4 *src = reinterpret_cast<const uint4 *>(tile);
#pragma unroll
  for (int r = 0; r < TILE_DIR_BYTES / 16 / 32; ++r) dst[lane + 32 * r] = src[lane + 32 * r];
}

template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(32) traceback_band_hop_kernel(const TracebackArgs a, const int chapter, const int last) {
  constexpr int RT = RT_WARP;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x;
  const uint32_t p = blockIdx.x;
  BandState *S = a.band_state + p;
  if (chapter > 0 && S->done) {
    if (lane == 0) S->nvisits = 0;  // nothing left for the emit launches
    return;
  }
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  const BandStart b = band_start(a, P, p, chapter);
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, T.m + T.n, 0};
  int i = b.i, j = b.j, misses = 0;
  if (chapter > 0) {
    W.pos = S->pos;
    W.cost = S->cost[lane];
    misses = S->misses;
  }
  const BandGeom g = band_geom(a, max(i, 1), max(j, 1), b.dhs, b.dcb);
  const int lowest = g.L0 - a.band_levels;  // last level of this chapter's band
  BandVisit *list = a.band_visits + (uint64_t)p * (uint64_t)a.band_visit_cap;
  int nlist = 0, ntiles = 0;
  while (i > 0 && j > 0) {
    const int hs = (i - 1) / TILE_ROWS, cb = (j - 1) / (4 * CKQ);
    if (!last && hs + cb < lowest) break;
    const int row0 = hs * TILE_ROWS, col0 = cb * 4 * CKQ;
    const int li = i - 1 - row0, lj = j - 1 - col0;
    const int entry = li == TILE_ROWS - 1 ? lj : (lj == 4 * CKQ - 1 ? 4 * CKQ + li : -1);
    const int slot = entry >= 0 ? band_slot(a, g, hs, cb) : -1;
    if (slot >= 0 && nlist < a.band_visit_cap) {
      const uint32_t E = __ldcg(band_table(a, p, slot) + entry);
      if (lane == 0) list[nlist] = BandVisit{hs, cb, li | (lj << 16), W.pos, slot};
      ++nlist;
      const int steps = (int)(E & 0xffffu), code = (int)(E >> 16);
      W.pos -= steps;
      if (code <= 4 * CKQ) {  // left through the top edge at tile column code - 1 (0: through the corner)
        i = row0;
        j = col0 + code;
      } else {  // left through the left edge at tile row code - 129
        i = row0 + code - 4 * CKQ;
        j = col0;
      }
    } else {
      // outside the band (or entered in the interior: a pair's first tile).  Unless this is the chapter's
      // first tile or the last chapter, end the chapter here: the next one is centred on this tile.
      if (!last && ntiles > 0) break;
      const int tq = (lj >> 2) + 1, tmax = li / RT;
      recompute_tile<MODE, ADDMODE, true, RT>(T, hs, cb, tq, tmax, tq + tmax, smem, lane);
      __syncwarp();
      walk_tile<RT>(smem, row0, col0, i, j, W, a.pxy, a.pgap, lane);
      __syncwarp();
      ++misses;
    }
    ++ntiles;
  }
  S->cost[lane] = W.cost;
  if (lane == 0) {
    S->i = i;
    S->j = j;
    S->pos = W.pos;
    S->done = (i > 0 && j > 0) ? 0 : 1;
    S->misses = misses;
    S->nvisits = nlist;
    if (chapter == 0) S->cost_emit = 0;
    // slope for the next chapter: this chapter's displacement, up to the tile the next one starts in
    // (after a turn that is the tile which ended this chapter: the new direction shows at once)
    const int dhs = g.hs0 - (max(i, 1) - 1) / TILE_ROWS, dcb = g.cb0 - (max(j, 1) - 1) / (4 * CKQ);
    if (dhs + dcb > 0 || chapter == 0) {
      S->dhs = dhs;
      S->dcb = dcb;
    }
  }
}

template <int MODE, int ADDMODE>
__global__ void __launch_bounds__(32 * BAND_WARPS) traceback_band_emit_kernel(const TracebackArgs a) {
  constexpr int RT = RT_WARP;
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t p = blockIdx.y;
  BandState *S = a.band_state + p;
  const int e = (int)blockIdx.x * BAND_WARPS + warp;
  if (e >= S->nvisits) return;
  const BandVisit V = a.band_visits[(uint64_t)p * (uint64_t)a.band_visit_cap + (uint64_t)e];
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  uint8_t *tile = smem + (size_t)warp * LONG_TILE_BYTES;
  const uint4 *src = reinterpret_cast<const uint4 *>(band_image(a, p, V.slot));
  uint4 *dst = reinterpret_cast<uint4 *>(tile);
#pragma unroll
  for (int r = 0; r < TILE_DIR_BYTES / 16 / 32; ++r) dst[lane + 32 * r] = __ldcg(src + lane + 32 * r);
  const int row0 = V.hs * TILE_ROWS, col0 = V.cb * 4 * CKQ;
  stage_tile_chars<RT, TILE_ROWS / RT>(T, row0, col0, tile, lane);
  __syncwarp();
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, V.pos, 0};
  int i = row0 + (V.cell & 0xffff) + 1, j = col0 + (V.cell >> 16) + 1;
  walk_tile<RT>(tile, row0, col0, i, j, W, a.pxy, a.pgap, lane);
  int cost = W.cost;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, d);
  if (lane == 0 && cost) atomicAdd(&S->cost_emit, cost);
}

template <int MODE>
__global__ void __launch_bounds__(32) traceback_band_finish_kernel(const TracebackArgs a) {
  const int lane = threadIdx.x;
  const uint32_t p = blockIdx.x;
  const BandState *S = a.band_state + p;
  const PairDesc P = a.pairs[p];
  const TileCtx T = make_ctx<MODE>(a, P);
  PathOut W{a.out + P.out_off, a.out + P.out_off + P.cap, S->pos, S->cost[lane] + (lane == 0 ? S->cost_emit : 0)};
  finish_pair<32>(a, T, p, S->i, S->j, W, lane, 0xffffffffu, 0);
}

template <int MODE, int ADDMODE>
static cudaError_t launch_band(const TracebackArgs &args, cudaStream_t stream) {
  const int images = (args.band_levels + 1) * (2 * args.band_halfwidth + 1);
  const dim3 pgrid((unsigned)((images + BAND_WARPS - 1) / BAND_WARPS), args.npairs);
  const dim3 egrid((unsigned)((args.band_visit_cap + BAND_WARPS - 1) / BAND_WARPS), args.npairs);
  for (int c = 0; c < args.band_chapters; ++c) {
    traceback_band_produce_kernel<MODE, ADDMODE>
        <<<pgrid, 32 * BAND_WARPS, BAND_WARPS * LONG_TILE_BYTES, stream>>>(args, c);
    traceback_band_hop_kernel<MODE, ADDMODE><<<args.npairs, 32, LONG_TILE_BYTES, stream>>>(args, c, c + 1 == args.band_chapters);
    traceback_band_emit_kernel<MODE, ADDMODE><<<egrid, 32 * BAND_WARPS, BAND_WARPS * LONG_TILE_BYTES, stream>>>(args);
  }
  traceback_band_finish_kernel<MODE><<<args.npairs, 32, 0, stream>>>(args);
  return cudaGetLastError();
}

template <int MODE, int ADDMODE>
static cudaError_t launch_tiles(const TracebackArgs &args, cudaStream_t stream) {
  if (args.band) return launch_band<MODE, ADDMODE>(args, stream);
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
