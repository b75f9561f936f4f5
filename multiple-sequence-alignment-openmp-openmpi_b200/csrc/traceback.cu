// traceback.cu — walks the packed 2-bit direction matrix on the GPU and writes the two
// gapped alignment strings.
//
// Replaces the reference's serial traceback + tail fill + strip scan:
//   submit/xuliny-seqalkway.cpp:502-531 (traceback, priority match / mismatch-diag / up / left),
//   :535-558 (tails: leftover prefix against '_'), :191-214 (do_task: strip common '_' pad,
//   build align1/align2).
//
// Closed form used here: with the walk stopping at (i, j), one of
// them 0,   align1 = x[0:i] + '_'*j + path_x,   align2 = '_'*i + y[0:j] + path_y,
// written right-aligned into two per-pair byte buffers so that both strings end at
// offset m+n; `start` marks their first byte.
//
// One thread per pair: the walk is a dependent chain of <= m+n steps; pairs of a batch
// run in parallel.  The thread caches the last 16 B of directions (64 cells = one lane's
// R rows x 64/R columns) and emits output through 8-byte registers so that the stores are
// 64-bit, not per character.
#include "seqalk_kernels.h"

namespace seqalk {

// Accumulates bytes written right-to-left and flushes aligned 8-byte words.
struct BackWriter {
  uint8_t *base;   // buffer base (16 B aligned)
  uint32_t pos;    // next byte goes to base[pos-1]
  uint64_t acc;    // bytes for the 8-byte word containing pos-1 (little endian placement)
  __device__ __forceinline__ void init(uint8_t *b, uint32_t end) {
    base = b;
    pos = end;
    acc = 0;
  }
  __device__ __forceinline__ void put(uint8_t ch) {
    --pos;
    acc |= (uint64_t)ch << (8 * (pos & 7));
    if ((pos & 7) == 0) {
      *reinterpret_cast<uint64_t *>(base + pos) = acc;
      acc = 0;
    }
  }
  // flush the partial word (bytes below pos inside the word are written as 0; they are
  // before `start`, nobody reads them)
  __device__ __forceinline__ void finish() {
    if (pos & 7) *reinterpret_cast<uint64_t *>(base + (pos & ~7u)) = acc;
  }
};

template <int R>
__global__ void __launch_bounds__(TRACEBACK_THREADS) traceback_kernel(const TracebackArgs a) {
  // 16 B of directions = one lane's R rows x (64/R) columns; the walk usually stays inside
  // it for several steps, so it is cached in registers.
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npairs) return;
  const PairDesc P = a.pairs[p];
  const int m = P.m, n = P.n;
  const uint8_t *__restrict__ xs = a.seq + P.x_off;
  const uint8_t *__restrict__ ys = a.seq + P.y_off;
  const uint8_t *__restrict__ dirs = a.dirs + P.dir_off;
  BackWriter w1, w2;
  w1.init(a.out + P.out_off, (uint32_t)(m + n));
  w2.init(a.out + P.out_off + P.cap, (uint32_t)(m + n));

  int i = m, j = n;
  long long cost = 0;  // path cost == dp[m][n]: the penalty of this pair
  uint64_t cached_off = ~0ull;
  uint4 cached = make_uint4(0, 0, 0, 0);
  while (i > 0 && j > 0) {
    const DirAddr da = dir_addr(i, j, n, R);
    const uint64_t line = da.byte_off & ~15ull;
    if (line != cached_off) {
      cached_off = line;
      cached = __ldcg(reinterpret_cast<const uint4 *>(dirs + line));
    }
    const uint32_t wsel = (uint32_t)(da.byte_off >> 2) & 3u;
    const uint32_t word = wsel == 0 ? cached.x : wsel == 1 ? cached.y : wsel == 2 ? cached.z : cached.w;
    const uint32_t code = (word >> da.shift) & 3u;
    const uint8_t xch = xs[i - 1], ych = ys[j - 1];
    if (!(code & DIR_NOT_DIAG)) {
      w1.put(xch);
      w2.put(ych);
      cost += (xch == ych) ? 0 : a.pxy;
      --i;
      --j;
    } else if (!(code & DIR_LEFT_BETTER)) {  // up: consume x, gap in y
      w1.put(xch);
      w2.put('_');
      cost += a.pgap;
      --i;
    } else {  // left: gap in x, consume y
      w1.put('_');
      w2.put(ych);
      cost += a.pgap;
      --j;
    }
  }
  // tails: only one of i, j is non-zero
  cost += (long long)(i + j) * a.pgap;
  for (; i > 0; --i) {
    w1.put(xs[i - 1]);
    w2.put('_');
  }
  for (; j > 0; --j) {
    w1.put('_');
    w2.put(ys[j - 1]);
  }
  uint32_t start = w1.pos;
  w1.finish();
  w2.finish();
  if (a.strip_scan) {  // '_' inside the inputs: the reference's scan for the last common '_' column
    const uint8_t *o1 = a.out + P.out_off, *o2 = o1 + P.cap;
    for (uint32_t c = (uint32_t)(m + n); c > start; --c)
      if (o1[c - 1] == '_' && o2[c - 1] == '_') {
        start = c;
        break;
      }
  }
  PairAlign r;
  r.start = start;
  r.len = (uint32_t)(m + n) - start;
  a.align[p] = r;
  a.path_cost[p] = (int32_t)cost;
}

cudaError_t launch_traceback(const TracebackArgs &args, cudaStream_t stream) {
  if (args.npairs == 0) return cudaSuccess;
  const unsigned grid = (args.npairs + TRACEBACK_THREADS - 1) / TRACEBACK_THREADS;
  if (args.R == 8)
    traceback_kernel<8><<<grid, TRACEBACK_THREADS, 0, stream>>>(args);
  else if (args.R == 16)
    traceback_kernel<16><<<grid, TRACEBACK_THREADS, 0, stream>>>(args);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace seqalk
