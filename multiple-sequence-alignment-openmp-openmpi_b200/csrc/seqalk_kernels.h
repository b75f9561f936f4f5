// seqalk_kernels.h — kernel argument blocks and launch entry points (internal to libseqalk).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "seqalk_device.cuh"

namespace seqalk {

// fill kernel variants
enum : int { MODE_GENERAL = 0, MODE_CMP = 1, MODE_TAB = 2 };
struct FillVariant {
  int R = 8;        // DP rows per lane (8 or 16); strip = 32*R rows
  int mode = MODE_CMP;
  int addmode = 0;  // 1: adds forced onto the FMA pipe (IMAD); 2: MODE_TAB lookup+add as one IDP.4A
  int minb = 0;     // > 0: cap registers so that this many fill CTAs fit on an SM (tile mode, R = 32)
  bool emit = true; // true: directions mode (2-bit codes to HBM); false: tile mode (scores +
                    // checkpoints, directions recomputed by the traceback for the tiles it visits)
};
// tile mode: a column checkpoint every CKQ macro columns (4*CKQ DP columns), a row every TILE_ROWS rows.
// 32 = 128 x 128 tiles: free for the packed fill (both halves of a register reach a checkpoint column
// at the same step) and, with two pairs per warp in the traceback, 25 % less traceback time on
// BASELINE config 4 than 128 x 256 tiles (811 -> 608 ms); the 32-bit fill pays ~5 % for it.
#ifndef SEQALK_CKQ
#define SEQALK_CKQ 32
#endif
constexpr int CKQ = SEQALK_CKQ;  // 64 or 32
// bytes of a pair's column checkpoints: int32 fill [nq / CKQ][mpad]; packed fill: images of mpad / 2
// words (fill16.cu) — two per checkpoint at CKQ = 64, ceil(nq / 32) at CKQ = 32
__host__ __device__ inline uint64_t ck_region_bytes(int nq, uint64_t mpad, bool packed) {
  if (packed && CKQ == 32) return (uint64_t)((nq + 31) / 32) * mpad * 2;
  return (uint64_t)(nq / CKQ) * mpad * 4;
}
#ifndef SEQALK_TILE_ROWS
#define SEQALK_TILE_ROWS 128
#endif
constexpr int TILE_ROWS = SEQALK_TILE_ROWS;  // 128 (default: -21 % traceback time on cfg4) or 256; 32 lanes x 4 or 8 rows in the recompute
// bottom-row progress is published every 2^publish_shift macro columns (and at strip end):
// 64 when a batch has plenty of independent pairs, 16 when consecutive strips of a few long
// pairs must run as a tight wavefront (shorter start-up ramp)
constexpr int PUBLISH_SHIFT_MANY = 6, PUBLISH_SHIFT_FEW = 4;
constexpr int TRACEBACK_THREADS = 32;
constexpr int DIGEST_THREADS = 64;
// bytes that may be read past the end of the packed sequence / alignment buffers
constexpr int BUFFER_SLACK = 64;
// zero bytes behind every sequence: the fill reads y in 4-byte words and x in 16-byte words, and the
// selector stream of a sequence (fill16.cu: 8 bytes per macro column, 64 extra entries) lives at twice
// the sequence's offset in a buffer twice as large
constexpr int SEQ_PAD = 272;

struct FillArgs {
  const uint8_t *seq;       // packed sequences (each 16 B aligned, zero padded)
  const uint8_t *codes;     // same layout, letters remapped to 8*(0..3) (MODE_TAB only, else null)
  const PairDesc *pairs;    // pairs of this batch
  const StripTask *tasks;   // strip queue, ordered (strip, pair)
  uint32_t ntasks;
  uint32_t *counter;        // queue head (zeroed before launch)
  uint32_t *progress;       // per strip: macro cols of its bottom row published (zeroed)
  int32_t *brow;            // boundary rows, 4*nq ints per pair
  uint8_t *dirs;            // packed 2-bit directions
  int32_t pxy, pgap;
  int32_t one;              // == 1, opaque to the compiler (see add_i in fill.cu)
  int32_t publish_shift;    // log2 of the publication interval in macro columns
  // 16x2-packed fill (fill16.cu) only:
  const uint8_t *sel16;     // selector streams, one per sequence, at byte offset 2 * (offset of the sequence)
  int32_t rebase_shift;     // log2 of the re-base interval in macro steps
  int32_t main_rp;          // rows per half lane of the job's main strip pairs (16: 1024-row, 32: 2048-row strip pairs)
};

// Banded traceback of long pairs (traceback_tiles.cu): what a pair carries from one chapter
// (produce + hop + emit launches) to the next.
struct BandState {
  int32_t i, j;       // the path continues from DP cell (i, j)
  int32_t pos;        // characters [pos, m + n) of the alignment buffers are accounted for
  int32_t done;       // 1: the path has reached row 0 or column 0
  int32_t misses;     // tiles the hopping warp had to recompute and walk itself, diagnostics
  int32_t dhs, dcb;   // tile rows / tile columns the previous chapter moved: slope of the next band
  int32_t nvisits;    // tiles recorded for this chapter's emit launch
  int32_t cost_emit;  // path cost collected by the emit launches
  int32_t pad[7];
  int32_t cost[32];   // per-lane shares of the path cost collected by the hopping warp
};
// A tile the path visits inside the band: what the emit launch needs to walk it.
struct BandVisit {
  int32_t hs, cb;  // tile
  int32_t cell;    // entry cell inside the tile: row | column << 16
  int32_t pos;     // characters of this tile end at `pos` of the alignment buffers
  int32_t slot;    // image of the tile
};

struct TracebackArgs {
  const uint8_t *seq;
  const uint8_t *codes;     // letters remapped to 0..3 (tile mode with MODE_TAB), else null
  const int32_t *brow;      // tile mode: kept boundary rows (one per TILE_ROWS DP rows)
  int32_t mode;             // tile mode: MODE_* of the recompute (same arithmetic as the fill)
  int32_t addmode;          // tile mode: 2 = IDP.4A cost lookup in the recompute
  int32_t long_pairs;       // tile mode: 1 = four warps per pair (walk + look-ahead recompute)
  int32_t half_warp;        // tile mode, many pairs: 1 = two pairs per warp (16 lanes x 8 rows per tile)
  const PairDesc *pairs;
  uint32_t npairs;
  const uint8_t *dirs;
  uint8_t *out;             // alignment scratch: per pair 2 right-aligned buffers of P.cap bytes
  PairAlign *align;         // per pair: start offset + length of the alignment strings
  int32_t *path_cost;       // per pair: cost of the walked path == dp[m][n] (every step of the walk
                            // satisfies dp[cell] = dp[predecessor] + step cost, and dp[i][0] = i*pgap)
  int32_t pxy, pgap;
  int32_t R;                // rows per lane the fill used (direction layout)
  int32_t ck16;             // tile mode: 0 = int32 column checkpoints (fill.cu); else the 16x2-packed fill's
                            // checkpoints (fill16.cu), value = log2 of its re-base interval
  int32_t main_rows;        // packed checkpoints: rows of the job's main strip pairs (1024 or 2048)
  int32_t strip_scan;       // 1: the input contains '_' (the gap character): apply the reference's
                            // strip rule literally (submit/xuliny-seqalkway.cpp:191-202): the alignment
                            // starts after the LAST column holding '_' in both strings
  // tile mode, jobs of a few LONG pairs: banded traceback in chapters (traceback_tiles.cu)
  int32_t band;             // 1: on
  int32_t band_levels;      // tile anti-diagonals (levels hs + cb) per chapter after the chapter's first
  int32_t band_halfwidth;   // tiles recomputed per level = 2 * band_halfwidth + 1
  int32_t band_chapters;    // chapters enqueued (the last one finishes whatever is left)
  int32_t band_visit_cap;   // visits a pair can record per chapter
  uint8_t *band_tiles;      // per pair (band_levels + 1) * (2 * band_halfwidth + 1) direction images of 4 KB
  uint32_t *band_tables;    // ... and as many transfer tables of 256 words
  BandState *band_state;    // per pair
  BandVisit *band_visits;   // per pair band_visit_cap records
};

// Where the fill left H'[m][n] of a pair (MODE_TAB only: padded rows copy DP row m, so the bottom
// row of the pair's last strip is H'[m][1..n]; see row_operand in fill_core.cuh).
enum : int { CORNER_NONE = 0, CORNER_TILES = 1, CORNER_DIRS = 2 };
__host__ __device__ inline uint64_t corner_index(const PairDesc &P, int corner_mode) {
  const uint64_t nq = (uint64_t)((P.n + 3) >> 2);
  return corner_mode == CORNER_TILES ? (uint64_t)(P.mpad / TILE_ROWS - 1) * (4 * nq + 4) + (uint64_t)(P.n - 1)
                                     : (uint64_t)(P.n - 1);
}

struct DigestArgs {
  const PairDesc *pairs;
  uint32_t npairs;
  const uint8_t *out;
  const PairAlign *align;
  const int32_t *penalties;     // per pair of the batch: dp[m][n] (path cost from the traceback)
  const uint32_t *result_slot;  // per pair of the batch: slot in the job-wide result arrays
  uint8_t *digests;             // job-wide: 64 raw bytes per slot (may be a peer GPU's memory)
  int32_t *penalties_out;       // job-wide: penalty per slot (may be a peer GPU's memory)
  // self-check: the fill's own dp[m][n] (corner_mode != CORNER_NONE) must equal the traceback's
  // path cost; mismatches are counted in check[0], the smallest failing task id goes to check[1]
  const int32_t *brow;
  int32_t corner_mode;
  int32_t pgap;
  uint32_t *check;
  int32_t small_regs;           // 1: the 64-register build (dirs mode: runs beside the persistent fill)
  int32_t host_hash;            // 1: the host hashed this batch's strings and uploaded the digests (a few very
                                // long strings: one SHA-512 lane is ~10x slower than a host core); the kernel
                                // only publishes the penalties and runs the self-check
};

// Score-only result: penalties without alignment strings (reference result dp[m][n],
// submit/xuliny-seqalkway.cpp:560).  corner_mode != CORNER_NONE: read from the fill's bottom row;
// else copy the traceback's path cost.
struct ScoreArgs {
  const PairDesc *pairs;
  uint32_t npairs;
  const int32_t *brow;
  int32_t corner_mode;
  int32_t pgap;
  const int32_t *path_cost;
  const uint32_t *result_slot;
  int32_t *penalties_out;
};

cudaError_t launch_fill(const FillArgs &args, const FillVariant &v, int sm_count, cudaStream_t stream);
// 16x2-packed score-only fill: rp = rows per half lane (16: strip pairs of 1024 rows, 8: 512 rows)
// flags (one-launch kernel, rp >= 100): FILL16_PREDCK = column checkpoints by predicated stores through a running
// pointer instead of a one-lane branch; FILL16_PIPE = lane hand-off shuffled as soon as a column is done, staged
// entries read one step ahead (8 rows per half lane).  max_ctas_per_sm > 0 caps the resident CTAs.
// FILL16_SENT (with FILL16_PREDCK; 8 or 16 rows per half lane, where a pair's strip pairs start every 64 * rp
// rows): hand-off rows pre-filled with FILL16_SENTINEL by the engine replace the progress counters.
enum : int { FILL16_PREDCK = 1, FILL16_PIPE = 2, FILL16_SENT = 4 };
constexpr uint32_t FILL16_SENTINEL = 0x80808080u;  // a byte pattern (cudaMemset2DAsync) no H' can take: |H'| < 2^29
cudaError_t launch_fill16(const FillArgs &args, int rp, int minb, int sm_count, cudaStream_t stream, int flags = 0,
                          int max_ctas_per_sm = 0);
int fill16_blocks_per_sm(int rp, int minb, int flags = 0);
// FILL16_SENT: sentinel words into every pair's hand-off rows (max_words: the most any pair needs, sizes the grid)
cudaError_t launch_fill16_sentinels(const PairDesc *pairs, uint32_t npairs, int32_t *brow, int main_rp, uint64_t max_words,
                                    cudaStream_t stream);
int fill16_rebase_shift(int pxy, int pgap, int rp);  // 0: penalties do not fit the 16-bit scheme with rp rows per half lane
cudaError_t launch_build_sel16(const uint8_t *codes, uint8_t *sel16, const uint32_t *off_dev, const int32_t *len_dev,
                               int k, int maxlen, cudaStream_t stream);
// entries of a strip pair's base log for a pair with nq macro columns
__host__ __device__ inline int base_log_entries(int nq, int rebase_shift) { return ((nq + 64) >> rebase_shift) + 1; }
int fill_blocks_per_sm(const FillVariant &v);
cudaError_t launch_traceback(const TracebackArgs &args, cudaStream_t stream);
cudaError_t launch_traceback_tiles(const TracebackArgs &args, cudaStream_t stream);
cudaError_t launch_digest(const DigestArgs &args, cudaStream_t stream);
cudaError_t launch_scores(const ScoreArgs &args, cudaStream_t stream);

}  // namespace seqalk
