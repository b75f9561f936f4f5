// microbench.cu — measures the integer issue rates the fill kernel is bounded by
// ("measure ... with a dependent-free microbenchmark on the box").
// Each thread runs N iterations of 32 independent single-instruction asm statements on 8
// registers; `volatile` keeps ptxas from merging or dropping them.  Reported in lane-ops/s.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/seqalk.h"

namespace {

enum Kind { K_LOP3 = 0, K_MIN3 = 1, K_IMAD = 2, K_MIXED = 3, K_ADDMIN = 4, K_SHF = 5, K_ALUMIN = 6, K_FILLMIX = 7,
            K_DP4A = 8, K_DP4A_MIN3 = 9, K_DP4A_IMAD = 10, K_PRMT = 11, K_MAX3U16 = 12, K_FILL16MIX = 13, K_DP4A_MIN3_DISTINCT = 14 };

template <int KIND>
__global__ void __launch_bounds__(256) issue_kernel(int *out, int iters, int a, int b) {
  int r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
#define ALU(r) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(a), "r"(b))
#define MIN3(r) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(r) : "r"(a), "r"(b))
#define ADDMIN(r) asm volatile("{.reg .s32 t; add.s32 t, %0, %2; min.s32 %0, t, %1;}" : "+r"(r) : "r"(a), "r"(b))
#define SHF(r) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(r) : "r"(a))
#define DP4A(r) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define PRMT(r) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b))
#define MAD(r) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r) : "r"(a), "r"(b))
#define MAX3U(r, x, y) asm volatile("{.reg .b32 t; max.u16x2 t, %0, %1; max.u16x2 %0, t, %2;}" : "+r"(r) : "r"(x), "r"(y))
#define PRMT3(d, x, y) asm volatile("prmt.b32 %0, %1, %2, %0;" : "+r"(d) : "r"(x), "r"(y))
#define MADU(d, x) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d) : "r"(x), "r"(b))
#define DP4A3(d, x, y) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(d) : "r"(x), "r"(y))
#define MIN33(d, x, y) asm volatile("{.reg .s32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(d) : "r"(x), "r"(y))
#define ROW(OP) OP(r0); OP(r1); OP(r2); OP(r3); OP(r4); OP(r5); OP(r6); OP(r7)
  for (int i = 0; i < iters; ++i) {
    if (KIND == K_LOP3) { ROW(ALU); ROW(ALU); ROW(ALU); ROW(ALU); }
    if (KIND == K_MIN3) { ROW(MIN3); ROW(MIN3); ROW(MIN3); ROW(MIN3); }
    if (KIND == K_IMAD) { ROW(MAD); ROW(MAD); ROW(MAD); ROW(MAD); }
    if (KIND == K_MIXED) { ROW(ALU); ROW(MAD); ROW(ALU); ROW(MAD); }
    if (KIND == K_ADDMIN) { ROW(ADDMIN); ROW(ADDMIN); ROW(ADDMIN); ROW(ADDMIN); }
    if (KIND == K_SHF) { ROW(SHF); ROW(SHF); ROW(SHF); ROW(SHF); }
    if (KIND == K_ALUMIN) { ROW(ALU); ROW(MIN3); ROW(ALU); ROW(MIN3); }
    if (KIND == K_FILLMIX) { ROW(MIN3); ROW(SHF); ROW(MAD); ROW(MAD); }
    if (KIND == K_DP4A) { ROW(DP4A); ROW(DP4A); ROW(DP4A); ROW(DP4A); }
    if (KIND == K_DP4A_MIN3) { ROW(DP4A); ROW(MIN3); ROW(DP4A); ROW(MIN3); }
    if (KIND == K_DP4A_IMAD) { ROW(DP4A); ROW(MAD); ROW(DP4A); ROW(MAD); }
    if (KIND == K_PRMT) { ROW(PRMT); ROW(PRMT); ROW(PRMT); ROW(PRMT); }
    if (KIND == K_MAX3U16) {  // VIMNMX3.U16x2, three different registers each
      MAX3U(r0, r1, r2); MAX3U(r3, r4, r5); MAX3U(r6, r7, r0); MAX3U(r1, r2, r3);
      MAX3U(r4, r5, r6); MAX3U(r7, r0, r1); MAX3U(r2, r3, r4); MAX3U(r5, r6, r7);
      MAX3U(r0, r2, r4); MAX3U(r1, r3, r5); MAX3U(r6, r0, r2); MAX3U(r7, r1, r3);
      MAX3U(r4, r6, r0); MAX3U(r5, r7, r1); MAX3U(r2, r4, r6); MAX3U(r3, r5, r7);
      MAX3U(r0, r3, r6); MAX3U(r1, r4, r7); MAX3U(r2, r5, r0); MAX3U(r3, r6, r1);
      MAX3U(r4, r7, r2); MAX3U(r5, r0, r3); MAX3U(r6, r1, r4); MAX3U(r7, r2, r5);
      MAX3U(r0, r5, r2); MAX3U(r1, r6, r3); MAX3U(r2, r7, r4); MAX3U(r3, r0, r5);
      MAX3U(r4, r1, r6); MAX3U(r5, r2, r7); MAX3U(r6, r3, r0); MAX3U(r7, r4, r1);
    }
    if (KIND == K_FILL16MIX) {  // the packed fill's mix, different registers: PRMT + IMAD + VIMNMX3.U16x2 (x 10, + 2)
      PRMT3(r0, r1, r2); MADU(r3, r4); MAX3U(r5, r6, r7); PRMT3(r1, r2, r3); MADU(r4, r5); MAX3U(r6, r7, r0);
      PRMT3(r2, r3, r4); MADU(r5, r6); MAX3U(r7, r0, r1); PRMT3(r3, r4, r5); MADU(r6, r7); MAX3U(r0, r1, r2);
      PRMT3(r4, r5, r6); MADU(r7, r0); MAX3U(r1, r2, r3); PRMT3(r5, r6, r7); MADU(r0, r1); MAX3U(r2, r3, r4);
      PRMT3(r6, r7, r0); MADU(r1, r2); MAX3U(r3, r4, r5); PRMT3(r7, r0, r1); MADU(r2, r3); MAX3U(r4, r5, r6);
      PRMT3(r0, r2, r4); MADU(r5, r7); MAX3U(r1, r3, r6); PRMT3(r1, r3, r5); MADU(r6, r0); MAX3U(r2, r4, r7);
      PRMT3(r3, r6, r1); MAX3U(r4, r7, r2);
    }
    if (KIND == K_DP4A_MIN3_DISTINCT) {  // the 32-bit fill's mix with three different registers per instruction
      DP4A3(r0, r1, r2); MIN33(r3, r4, r5); DP4A3(r6, r7, r0); MIN33(r1, r2, r3);
      DP4A3(r4, r5, r6); MIN33(r7, r0, r1); DP4A3(r2, r3, r4); MIN33(r5, r6, r7);
      DP4A3(r0, r2, r4); MIN33(r1, r3, r5); DP4A3(r6, r0, r2); MIN33(r7, r1, r3);
      DP4A3(r4, r6, r0); MIN33(r5, r7, r1); DP4A3(r2, r4, r6); MIN33(r3, r5, r7);
      DP4A3(r0, r3, r6); MIN33(r1, r4, r7); DP4A3(r2, r5, r0); MIN33(r3, r6, r1);
      DP4A3(r4, r7, r2); MIN33(r5, r0, r3); DP4A3(r6, r1, r4); MIN33(r7, r2, r5);
      DP4A3(r0, r5, r2); MIN33(r1, r6, r3); DP4A3(r2, r7, r4); MIN33(r3, r0, r5);
      DP4A3(r4, r1, r6); MIN33(r5, r2, r7); DP4A3(r6, r3, r0); MIN33(r7, r4, r1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
}

template <int KIND>
double run_kind(int sm_count, int *d_out) {
  const int grid = sm_count * 8, block = 256, iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  issue_kernel<KIND><<<grid, block>>>(d_out, 64, 3, 5);  // warm-up
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    issue_kernel<KIND><<<grid, block>>>(d_out, iters, 3, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * block * (double)iters * 32.0;
    const double rate = ops / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace

extern "C" int seqalk_measure_int_peaks(int32_t device, double rates_out[16]) {
  if (!rates_out) return 1;
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 1;
  int *d_out = nullptr;
  if (cudaMalloc(&d_out, (size_t)prop.multiProcessorCount * 8 * 256 * sizeof(int)) != cudaSuccess) return 1;
  const int sms = prop.multiProcessorCount;
  rates_out[0] = run_kind<K_LOP3>(sms, d_out);
  rates_out[1] = run_kind<K_MIN3>(sms, d_out);
  rates_out[2] = run_kind<K_IMAD>(sms, d_out);
  rates_out[3] = run_kind<K_MIXED>(sms, d_out);
  rates_out[4] = run_kind<K_ADDMIN>(sms, d_out);
  rates_out[5] = run_kind<K_SHF>(sms, d_out);
  rates_out[6] = run_kind<K_ALUMIN>(sms, d_out);
  rates_out[7] = run_kind<K_FILLMIX>(sms, d_out);
  rates_out[8] = run_kind<K_DP4A>(sms, d_out);
  rates_out[9] = run_kind<K_DP4A_MIN3>(sms, d_out);
  rates_out[10] = run_kind<K_DP4A_IMAD>(sms, d_out);
  rates_out[11] = run_kind<K_PRMT>(sms, d_out);
  rates_out[12] = run_kind<K_MAX3U16>(sms, d_out);
  rates_out[13] = run_kind<K_FILL16MIX>(sms, d_out);
  rates_out[14] = run_kind<K_DP4A_MIN3_DISTINCT>(sms, d_out);
  rates_out[15] = 0;
  cudaFree(d_out);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
