// seqalk_kernels.h — kernel argument blocks and launch entry points (internal to libseqalk).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "seqalk_device.cuh"

namespace seqalk {

// rows per lane used by the production kernels (strip = 32*FILL_R rows)
constexpr int FILL_R = 8;
// bottom-row progress is published every PUBLISH_EVERY macro columns (and at strip end)
constexpr int PUBLISH_EVERY = 64;
constexpr int TRACEBACK_THREADS = 32;
constexpr int DIGEST_THREADS = 64;
// bytes that may be read past the end of the packed sequence / alignment buffers
constexpr int BUFFER_SLACK = 64;

struct FillArgs {
  const uint8_t *seq;       // packed sequences (each 16 B aligned, zero padded)
  const PairDesc *pairs;    // pairs of this batch
  const StripTask *tasks;   // strip queue, ordered (strip, pair)
  uint32_t ntasks;
  uint32_t *counter;        // queue head (zeroed before launch)
  uint32_t *progress;       // per strip: macro cols of its bottom row published (zeroed)
  int32_t *brow;            // boundary rows, 4*nq ints per pair
  uint8_t *dirs;            // packed 2-bit directions
  int32_t *penalties;       // per pair of the batch: dp[m][n]
  int32_t pxy, pgap;
};

struct TracebackArgs {
  const uint8_t *seq;
  const PairDesc *pairs;
  uint32_t npairs;
  const uint8_t *dirs;
  uint8_t *out;             // alignment scratch: per pair 2 right-aligned buffers of P.cap bytes
  PairAlign *align;         // per pair: start offset + length of the alignment strings
  int32_t *path_cost;       // per pair: cost of the walked path (== dp[m][n], host-checked)
  int32_t pxy, pgap;
};

struct DigestArgs {
  const PairDesc *pairs;
  uint32_t npairs;
  const uint8_t *out;
  const PairAlign *align;
  const int32_t *penalties;     // per pair of the batch (from the fill)
  const int32_t *path_cost;     // per pair of the batch (from the traceback); must equal penalties
  uint32_t *error_flag;         // set to 1 if any pair's path cost != dp[m][n]
  const uint32_t *result_slot;  // per pair of the batch: slot in the job-wide result arrays
  uint8_t *digests;             // job-wide: 64 raw bytes per slot (may be a peer GPU's memory)
  int32_t *penalties_out;       // job-wide: penalty per slot (may be a peer GPU's memory)
};

cudaError_t launch_fill(const FillArgs &args, bool general, int sm_count, cudaStream_t stream);
int fill_blocks_per_sm(bool general);
cudaError_t launch_traceback(const TracebackArgs &args, cudaStream_t stream);
cudaError_t launch_digest(const DigestArgs &args, cudaStream_t stream);

}  // namespace seqalk
