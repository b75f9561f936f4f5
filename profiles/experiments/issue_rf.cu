// issue_rf.cu — round-2 experiments behind the fill kernel's ceiling (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rf issue_rf.cu && ./issue_rf
//
// Part 1: does the ALU + FMA-heavy dual issue survive THREE DISTINCT register operands per
//         instruction?  (microbench.cu's streams share two of the three sources, which the
//         operand-reuse cache serves; the fill's IDP.4A / VIMNMX3 read three different registers.)
// Part 2: bare R x 4 block loops: 32-bit column-major / wavefront (as round 1) and 16x2-packed
//         (two strips of one pair in the halves of a register: PRMT cost lookup + add +
//         VIMNMX3.U16x2), with the measured cell rate of each.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define MIN3(r, a, b) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(r) : "r"(a), "r"(b))
#define DP4A(r, a, b) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define MIN2(r, a) asm volatile("min.s32 %0, %0, %1;" : "+r"(r) : "r"(a))
#define MAX3U(r, a, b) asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(r) : "r"(a), "r"(b))
#define PRMTD(r, a, b) asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define MAD1(r, a, one) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(one))

template <int KIND>
__global__ void __launch_bounds__(128) opk(int *out, const int *in, int iters, int one) {
  int r[8], q[8], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r[i] = in[threadIdx.x + 32 * i];
    q[i] = in[threadIdx.x + 32 * i + 256];
    a[i] = in[threadIdx.x + 32 * i + 512];
    b[i] = in[threadIdx.x + 32 * i + 768];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      if (KIND == 0) {  // VIMNMX3, three distinct registers
#pragma unroll
        for (int i = 0; i < 8; ++i) MIN3(r[i], a[i], b[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) MIN3(q[i], b[i], a[(i + 1) & 7]);
      }
      if (KIND == 1) {  // IDP.4A, three distinct registers
#pragma unroll
        for (int i = 0; i < 8; ++i) DP4A(r[i], a[i], b[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) DP4A(q[i], b[i], a[(i + 1) & 7]);
      }
      if (KIND == 2) {  // IDP.4A + VIMNMX3 alternating, all operands distinct
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          DP4A(r[i], a[i], b[i]);
          MIN3(q[i], b[(i + 3) & 7], a[(i + 5) & 7]);
        }
      }
      if (KIND == 3) {  // IDP.4A shares its second source (the fill's column mask), VIMNMX3 distinct
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          DP4A(r[i], a[i], b[0]);
          MIN3(q[i], b[(i + 3) & 7], a[(i + 5) & 7]);
        }
      }
      if (KIND == 4) {  // both share two sources (= microbench.cu)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          DP4A(r[i], a[0], b[0]);
          MIN3(q[i], a[0], b[0]);
        }
      }
      if (KIND == 5) {  // IDP.4A distinct + two-source VIMNMX
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          DP4A(r[i], a[i], b[i]);
          MIN2(q[i], b[i]);
        }
      }
      if (KIND == 6) {  // IMAD (two register sources) + VIMNMX3 distinct
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          MAD1(r[i], a[i], one);
          MIN3(q[i], b[i], a[(i + 1) & 7]);
        }
      }
      if (KIND == 8) {  // VIMNMX3.U16x2, three distinct registers
#pragma unroll
        for (int i = 0; i < 8; ++i) MAX3U(r[i], a[i], b[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) MAX3U(q[i], b[(i + 3) & 7], a[(i + 5) & 7]);
      }
      if (KIND == 9) {  // the packed fill's mix: PRMT + IMAD + VIMNMX3.U16x2 on independent registers
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          PRMTD(a[i], r[i], b[i]);
          MAD1(b[i], q[i], one);
          MAX3U(q[(i + 1) & 7], r[(i + 2) & 7], a[(i + 3) & 7]);
        }
      }
      if (KIND == 11) {  // PRMT with its two table operands shared (per-column tables) + IMAD + VIMNMX3.U16x2
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(r[0]), "r"(b[0]));
          MAD1(b[(i & 3) + 4], q[i], one);
          MAX3U(q[(i + 1) & 7], r[(i + 2) & 7], a[(i + 3) & 7]);
        }
      }
      if (KIND == 10) {  // PRMT + VIMNMX3.U16x2 (ALU pipe only)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          PRMTD(a[i], r[i], b[i]);
          MAX3U(q[i], r[(i + 2) & 7], b[(i + 3) & 7]);
        }
      }
      if (KIND == 7) {  // VIMNMX3 with ONE shared source
#pragma unroll
        for (int i = 0; i < 8; ++i) MIN3(r[i], a[0], b[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) MIN3(q[i], a[0], b[i]);
      }
    }
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= r[i] ^ q[i];
  out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int KIND>
void run_op(const char *name, int ctas_per_sm, int *out, int *in) {
  const int grid = 148 * ctas_per_sm, iters = 8192;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  opk<KIND><<<grid, 128>>>(out, in, 64, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    opk<KIND><<<grid, 128>>>(out, in, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double per_iter = (KIND == 9 || KIND == 11) ? 48.0 : 32.0;
  const double instr = (double)grid * 4 * iters * per_iter;  // warp instructions
  // per SMSP: warp instructions per clock at 1.965 GHz
  const double ipc = instr / (148.0 * 4) / (best * 1e-3 * 1.965e9);
  printf("op %-44s %d CTAs/SM: %8.1f Glane-instr/s  IPC/SMSP@1965 = %.3f\n", name, ctas_per_sm,
         instr * 32 / best / 1e6, ipc);
}

// ---------------------------------------------------------------- 32-bit blocks (round 1)
template <int VAR, int R>
__device__ __forceinline__ void block32(int (&h)[R], int &dg, const int (&xc)[R], const int (&up_in)[4],
                                        unsigned yword, int (&outv)[4]) {
  if (VAR == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      unsigned sel = __funnelshift_l(0u, 1u, yword >> (8 * c));
      int diag = dg, up = up_in[c];
      dg = up;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int left = h[r];
        int t = __dp4a(xc[r], (int)sel, diag);
        int v = min(min(up, left), t);
        diag = left;
        up = v;
        h[r] = v;
      }
      outv[c] = up;
    }
  } else {
    int diag[4], up[4];
    unsigned sel[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sel[c] = __funnelshift_l(0u, 1u, yword >> (8 * c));
      up[c] = up_in[c];
      diag[c] = (c == 0) ? dg : up_in[c - 1];
    }
    dg = up_in[3];
#pragma unroll
    for (int k = 0; k < R + 3; ++k) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int r = k - c;
        if (r >= 0 && r < R) {
          const int left = h[r];
          int t = __dp4a(xc[r], (int)sel[c], diag[c]);
          int v = min(min(up[c], left), t);
          diag[c] = left;
          up[c] = v;
          h[r] = v;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) outv[c] = up[c];
  }
}

template <int VAR, int R, int MINB>
__global__ void __launch_bounds__(128, MINB) loop32(int *out, const unsigned *yws, int n) {
  int h[R], xc[R], dg = 0, outv[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < R; ++r) {
    h[r] = 0;
    xc[r] = out[(r + threadIdx.x) & 1023];
  }
  for (int q = 0; q < n; ++q) {
    int up_in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) up_in[c] = __shfl_up_sync(0xffffffffu, outv[c], 1);
    block32<VAR, R>(h, dg, xc, up_in, yws[q & 1023], outv);
  }
  int s = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) s += h[r];
  out[(blockIdx.x * 128 + threadIdx.x) & 1023] = s;
}

// ---------------------------------------------------------------- 16x2-packed block
// A register holds two DP cells: low half = row r of strip s at column c, high half = row r of
// strip s+1 at column c - D (the second strip trails the first).  Column-major as above.
//   cost pair = PRMT(tabLo[r], tabHi[r], sel_c)   bytes: [lo cost, 0, hi cost, 0]  (ALU pipe)
//   t         = diag + cost pair                   plain 32-bit add, halves cannot carry (FMA pipe as IMAD, ADD = 1)
//   v         = max3.u16x2(up, left, t)            (ALU pipe)        [max domain: G = -H']
// ADD = 2: VIMNMX.U16x2(up, left) + VIADDMNMX-style fused add+max instead (all ALU).
template <int R, int ADD>
__device__ __forceinline__ void block16(unsigned (&h)[R], unsigned &dg, const unsigned (&tl)[R], const unsigned (&th)[R],
                                        const unsigned (&up_in)[4], unsigned yword, unsigned (&outv)[4], unsigned one) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned sel = 0xC480u | ((yword >> (2 * c)) & 3u) | (((yword >> (8 + 2 * c)) & 3u) << 8);
    unsigned diag = dg, up = up_in[c];
    dg = up;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned left = h[r];
      const unsigned cost = __byte_perm(tl[r], th[r], sel);
      unsigned v;
      if (ADD == 0) {
        v = __vimax3_u16x2(up, left, diag + cost);
      } else if (ADD == 1) {
        unsigned t;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(diag), "r"(one), "r"(cost));
        v = __vimax3_u16x2(up, left, t);
      } else {
        v = __viaddmax_u16x2(diag, cost, __vmaxu2(up, left));
      }
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
  }
}

template <int R, int ADD, int MINB>
__global__ void __launch_bounds__(128, MINB) loop16(unsigned *out, const unsigned *yws, int n, unsigned one) {
  unsigned h[R], tl[R], th[R], dg = 0, outv[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < R; ++r) {
    h[r] = 0;
    tl[r] = out[(r + threadIdx.x) & 1023] & 0x03030303u;
    th[r] = out[(r + 77 + threadIdx.x) & 1023] & 0x03030303u;
  }
  for (int q = 0; q < n; ++q) {
    unsigned up_in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) up_in[c] = __shfl_up_sync(0xffffffffu, outv[c], 1);
    block16<R, ADD>(h, dg, tl, th, up_in, yws[q & 1023], outv, one);
    if ((q & 255) == 255) {  // stand-in for the periodic 16-bit re-basing (one packed subtract per register)
      const unsigned base = __shfl_sync(0xffffffffu, h[0], 0) & 0x7fff7fffu;
#pragma unroll
      for (int r = 0; r < R; ++r) h[r] = __vsub2(h[r], base);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) s += h[r];
  out[(blockIdx.x * 128 + threadIdx.x) & 1023] = s;
}

// Variant: the cost table belongs to the COLUMN (two registers per DP column: gains of the column's
// letter pair against the four row letters) and the selector to the ROW: PRMT(colA, colB, rsel[r]).
// The two table operands are the same for all rows of a column (operand reuse cache), and a row
// needs one register instead of two.
template <int R, int ORDER>
__device__ __forceinline__ void block16c(unsigned (&h)[R], unsigned &dg, const unsigned (&rsel)[R],
                                         const unsigned (&up_in)[4], unsigned ywa, unsigned ywb, unsigned (&outv)[4],
                                         unsigned one, unsigned cbase, unsigned cdelta) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned ta = cbase ^ (cdelta << ((ywa >> (8 * c)) & 24u));
    const unsigned tb = cbase ^ (cdelta << ((ywb >> (8 * c)) & 24u));
    unsigned diag = dg, up = up_in[c];
    dg = up;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const unsigned left = h[r];
      unsigned cost;
      asm("prmt.b32 %0, %1, %2, %3;" : "=r"(cost) : "r"(ta), "r"(tb), "r"(rsel[r]));
      unsigned t;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(diag), "r"(one), "r"(cost));
      const unsigned v = ORDER ? __vimax3_u16x2(left, up, t) : __vimax3_u16x2(up, left, t);
      diag = left;
      up = v;
      h[r] = v;
    }
    outv[c] = up;
  }
}

template <int R, int ORDER, int MINB>
__global__ void __launch_bounds__(128, MINB) loop16c(unsigned *out, const unsigned *yws, int n, unsigned one) {
  unsigned h[R], rsel[R], dg = 0, outv[4] = {0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < R; ++r) {
    h[r] = 0;
    rsel[r] = 0xC480u | (out[(r + threadIdx.x) & 1023] & 0x0303u);
  }
  const unsigned cbase = out[5] & 0x01010101u, cdelta = (out[6] & 3u) + 1u;
  for (int q = 0; q < n; ++q) {
    unsigned up_in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) up_in[c] = __shfl_up_sync(0xffffffffu, outv[c], 1);
    block16c<R, ORDER>(h, dg, rsel, up_in, yws[q & 1023], yws[(q + 32) & 1023], outv, one, cbase, cdelta);
    if ((q & 255) == 255) {
      const unsigned base = __shfl_sync(0xffffffffu, h[0], 0) & 0x7fff7fffu;
#pragma unroll
      for (int r = 0; r < R; ++r) h[r] = __vsub2(h[r], base);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) s += h[r];
  out[(blockIdx.x * 128 + threadIdx.x) & 1023] = s;
}

template <typename K>
void time_loop(const char *name, K kernel, int R, int cells_per_reg, int blocks_per_sm, int regs, int maxb,
               void (*launch)(int, int, void *, void *, int), void *out, void *y) {
  (void)kernel;
  const int n = 20000, grid = 148 * blocks_per_sm;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch(grid, 0, out, y, 100);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    launch(grid, 0, out, y, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  const double cells = (double)grid * 128 * n * R * 4 * cells_per_reg;
  printf("loop %-34s R=%2d CTAs/SM=%d (max %d, regs %3d): %8.1f GCUPS\n", name, R, blocks_per_sm, maxb, regs,
         cells / best / 1e6);
}

#define RUN32(VAR, R, MINB, NB, NAME)                                                                   \
  {                                                                                                     \
    auto k = loop32<VAR, R, MINB>;                                                                      \
    int nb = 0;                                                                                         \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 128, 0);                                      \
    cudaFuncAttributes at;                                                                              \
    cudaFuncGetAttributes(&at, k);                                                                      \
    time_loop(NAME, k, R, 1, NB, at.numRegs, nb,                                                        \
              [](int g, int, void *o, void *y, int n) { loop32<VAR, R, MINB><<<g, 128>>>((int *)o, (const unsigned *)y, n); }, \
              out, y);                                                                                  \
  }
#define RUN16(R, ADD, MINB, NB, NAME)                                                                   \
  {                                                                                                     \
    auto k = loop16<R, ADD, MINB>;                                                                      \
    int nb = 0;                                                                                         \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 128, 0);                                      \
    cudaFuncAttributes at;                                                                              \
    cudaFuncGetAttributes(&at, k);                                                                      \
    time_loop(NAME, k, R, 2, NB, at.numRegs, nb,                                                        \
              [](int g, int, void *o, void *y, int n) { loop16<R, ADD, MINB><<<g, 128>>>((unsigned *)o, (const unsigned *)y, n, 1u); }, \
              out, y);                                                                                  \
  }

int main() {
  int *out, *in;
  unsigned *y;
  cudaMalloc(&out, 148 * 16 * 128 * 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&y, 4096);
  cudaMemset(out, 1, 4096);
  cudaMemset(in, 3, 4096 * 4);
  cudaMemset(y, 0x1b, 4096);
  for (int nb : {4, 8}) {
    run_op<0>("VIMNMX3 distinct", nb, out, in);
    run_op<1>("IDP.4A distinct", nb, out, in);
    run_op<2>("IDP.4A + VIMNMX3 distinct", nb, out, in);
    run_op<3>("IDP.4A(shared b) + VIMNMX3 distinct", nb, out, in);
    run_op<4>("IDP.4A + VIMNMX3, two shared sources", nb, out, in);
    run_op<5>("IDP.4A distinct + VIMNMX (2 sources)", nb, out, in);
    run_op<6>("IMAD (2 reg sources) + VIMNMX3 distinct", nb, out, in);
    run_op<7>("VIMNMX3, one shared source", nb, out, in);
    run_op<8>("VIMNMX3.U16x2 distinct", nb, out, in);
    run_op<9>("PRMT + IMAD + VIMNMX3.U16x2", nb, out, in);
    run_op<10>("PRMT + VIMNMX3.U16x2", nb, out, in);
    run_op<11>("PRMT(2 shared) + IMAD + VIMNMX3.U16x2", nb, out, in);
  }
  cudaMemset(out, 1, 4096);
  RUN32(0, 32, 4, 4, "32-bit column-major");
  RUN32(1, 32, 4, 4, "32-bit wavefront");
  RUN32(0, 16, 6, 6, "32-bit column-major");
  RUN32(1, 16, 6, 6, "32-bit wavefront");
  RUN32(1, 16, 8, 8, "32-bit wavefront");
  RUN32(0, 16, 8, 8, "32-bit column-major");
  RUN16(16, 0, 4, 4, "16x2 add (ptxas' choice)");
  RUN16(16, 1, 4, 4, "16x2 add on FMA pipe");
  RUN16(16, 2, 4, 4, "16x2 VIMNMX + VIADDMNMX");
  RUN16(16, 0, 6, 6, "16x2 add (ptxas' choice)");
  RUN16(16, 1, 6, 6, "16x2 add on FMA pipe");
  RUN16(16, 1, 8, 8, "16x2 add on FMA pipe");
  RUN16(8, 1, 8, 8, "16x2 add on FMA pipe");
  RUN16(8, 1, 12, 12, "16x2 add on FMA pipe");
  RUN16(32, 1, 3, 3, "16x2 add on FMA pipe");
  RUN16(32, 1, 4, 4, "16x2 add on FMA pipe");
#define RUN16C(R, ORDER, MINB, NB, NAME)                                                                \
  {                                                                                                     \
    auto k = loop16c<R, ORDER, MINB>;                                                                   \
    int nb = 0;                                                                                         \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, 128, 0);                                      \
    cudaFuncAttributes at;                                                                              \
    cudaFuncGetAttributes(&at, k);                                                                      \
    time_loop(NAME, k, R, 2, NB, at.numRegs, nb,                                                        \
              [](int g, int, void *o, void *y, int n) { loop16c<R, ORDER, MINB><<<g, 128>>>((unsigned *)o, (const unsigned *)y, n, 1u); }, \
              out, y);                                                                                  \
  }
  RUN16C(16, 0, 4, 4, "16x2 column tables");
  RUN16C(16, 1, 4, 4, "16x2 column tables, left first");
  RUN16C(16, 1, 6, 6, "16x2 column tables, left first");
  RUN16C(16, 1, 8, 8, "16x2 column tables, left first");
  RUN16C(32, 0, 4, 4, "16x2 column tables");
  RUN16C(32, 1, 4, 4, "16x2 column tables, left first");
  RUN16C(32, 1, 5, 5, "16x2 column tables, left first");
  RUN16C(8, 1, 8, 8, "16x2 column tables, left first");
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
