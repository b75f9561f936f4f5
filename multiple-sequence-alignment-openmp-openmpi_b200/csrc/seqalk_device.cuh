// seqalk_device.cuh — device-side data layout shared by the fill, traceback and digest kernels.
//
// Replaces the reference's per-pair `int dp[m+1][n+1]` slab (new2d,
// submit/xuliny-seqalkway.cpp:167-181) with a 2-bit direction matrix in a
// strip-major, step-skewed layout (DESIGN.md "Data layout in HBM").
#pragma once
#include <cstdint>

namespace seqalk {

// One alignment task = one (i, j) pair of the reference's task queue
// (Triple{i,j,task_id}, submit/xuliny-seqalkway.cpp:105-107, 280-287).
// x = genes[i] indexes DP rows (length m), y = genes[j] indexes DP columns (length n).
struct PairDesc {
  uint32_t x_off;      // byte offset of x in the packed sequence buffer (16 B aligned)
  uint32_t y_off;      // byte offset of y
  int32_t m;           // rows
  int32_t n;           // columns
  uint64_t dir_off;    // byte offset of this pair's direction strips (16 B aligned)
  uint64_t brow_off;   // int32 index of this pair's boundary-row buffer (4*nq ints)
  uint64_t out_off;    // byte offset of this pair's alignment scratch (2 buffers of cap bytes)
  uint32_t strip0;     // index of this pair's first strip in the progress array
  int32_t nstrips;     // ceil(m / (32*R))
  uint32_t cap;        // bytes per alignment buffer = round16(m + n)
  int32_t task_id;     // position in the reference's task order (for result placement)
  uint32_t mpad;       // rows padded to the strips (stride of the tile-mode column checkpoints)
  uint32_t half_last;  // tile mode, R = 32: the last strip has 512 rows (16 per lane) instead of 1024
};

struct StripTask {
  uint32_t pair;   // index into PairDesc[]
  uint32_t strip;  // 0 .. nstrips-1
};

// Per-pair result written by the traceback kernel, consumed by the digest kernel.
struct PairAlign {
  uint32_t start;  // alignment strings occupy [start, m+n) of each buffer
  uint32_t len;    // alignment length L = m + n - start
};

// Geometry helpers (host + device).  R = DP rows owned by one lane.
//   strip      = 32*R consecutive rows handled by one warp sweep
//   macro col  = 4 consecutive DP columns
//   macro step = one lane's R x 4 block; lane t works on macro col (q - t) at macro step q
//   group      = G = 16/R macro steps -> 128 direction bits per lane -> one 16 B store
__host__ __device__ constexpr int strip_rows(int R) { return 32 * R; }
__host__ __device__ constexpr int group_steps(int R) { return 16 / R; }
__host__ __device__ inline int macro_cols(int n) { return (n + 3) >> 2; }
__host__ __device__ inline int strip_groups(int n, int R) {
  return (macro_cols(n) + 31 + group_steps(R) - 1) / group_steps(R);
}
__host__ __device__ inline uint64_t strip_dir_bytes(int n, int R) {
  return (uint64_t)strip_groups(n, R) * 512ull;
}
__host__ __device__ inline int num_strips(int m, int R) { return (m + strip_rows(R) - 1) / strip_rows(R); }

// Strip geometry of a pair: all strips have 32*R rows, except that with `mixed` (tile mode,
// R = 32) a last strip covering <= 512 remaining rows is a half strip of 16 rows per lane —
// halves the average row padding (5.0 % -> 2.5 % of the cells on BASELINE config 4).
// `half_last` is a tail code: bit 0 = the pair ends with a 512-row strip, bit 1 = (three-level
// geometry only) a 1024-row strip before that / at the end.
struct StripGeom {
  int nstrips;
  uint32_t mpad;
  uint32_t half_last;
};
__host__ __device__ inline StripGeom strip_geometry(int m, int R, bool mixed) {
  StripGeom g;
  const int h = strip_rows(R);
  g.nstrips = (m + h - 1) / h;
  const int rem = m - (g.nstrips - 1) * h;  // rows of the last strip, 1..h
  g.half_last = (mixed && R == 32 && rem <= h / 2) ? 1u : 0u;
  g.mpad = (uint32_t)((g.nstrips - 1) * h + (g.half_last ? h / 2 : h));
  return g;
}
// Three-level geometry of the packed fill with 32 rows per half lane: strip pairs of 2048 rows,
// then at most one of 1024 and one of 512 rows, so that the padding stays <= 512 rows.
__host__ __device__ inline StripGeom strip_geometry3(int m) {
  StripGeom g;
  int n0 = m / 2048;
  const int rem = m - n0 * 2048;
  uint32_t tail = 0;
  if (rem == 0) tail = 0;
  else if (rem <= 512) tail = 1;
  else if (rem <= 1024) tail = 2;
  else if (rem <= 1536) tail = 3;
  else ++n0;
  g.nstrips = n0 + (int)(tail & 1) + (int)(tail >> 1);
  g.mpad = (uint32_t)(n0 * 2048 + ((tail & 2) ? 1024 : 0) + ((tail & 1) ? 512 : 0));
  g.half_last = tail;
  return g;
}
// strips of the main height in a pair with tail code `tail`
__host__ __device__ inline int main_strips(int nstrips, uint32_t tail) { return nstrips - (int)(tail & 1) - (int)((tail >> 1) & 1); }

// Direction codes.  16 cells per 32-bit word, shifted in MSB-first in the order
// idx = c*R + r (column-major inside the lane's R x 4 block).
constexpr int CELLS_PER_WORD = 16;
enum : unsigned { DIR_NOT_DIAG = 2u, DIR_LEFT_BETTER = 1u };

// Byte offset (inside a pair's direction region) of the 32-bit word holding DP cell
// (i, j), 1-based, and the shift that brings its 2-bit code to bits [1:0].
struct DirAddr {
  uint64_t byte_off;
  uint32_t shift;
};
__host__ __device__ inline DirAddr dir_addr(int i, int j, int n, int R) {
  const int G = group_steps(R);
  const int ri = i - 1, cj = j - 1;
  const int s = ri / strip_rows(R);
  const int t = (ri / R) & 31;
  const int r = ri % R;
  const int q = (cj >> 2) + t;  // macro step at which lane t visits this macro column
  const int g = q / G, u = q % G;
  const int idx = (cj & 3) * R + r;
  const int wl = idx / CELLS_PER_WORD, k = idx % CELLS_PER_WORD;
  DirAddr a;
  a.byte_off = (uint64_t)s * strip_dir_bytes(n, R) + ((uint64_t)g * 32 + t) * 16 + 4 * (u * (R / 4) + wl);
  a.shift = 30 - 2 * k;
  return a;
}

}  // namespace seqalk
