// engine.h — host-side orchestration of the all-pairs alignment on one GPU (internal).
//
// Plays the role of the reference's worker loop (recv Triple -> do_task -> send Packet,
// submit/xuliny-seqalkway.cpp:342-357, 399-411) for one device: takes a list of (i, j)
// tasks, cuts it into batches that fit the HBM budget for packed directions, and for each
// batch runs  fill  on the fill stream and  traceback -> digest (-> D2H)  on a second,
// high-priority stream, double-buffered over NSLOT buffer sets, so that the latency-bound
// traceback of batch b runs under the fill of batch b+1 (the fill leaves a sliver of every
// SM's registers free for it).
#pragma once
#include <cuda_runtime.h>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include "seqalk_kernels.h"

namespace seqalk {

// Host copy of the input in the layout the kernels read: every sequence starts on a
// 16 B boundary and is followed by >= 16 zero bytes (the fill reads y in 4-byte words).
// The bytes live in page-locked host memory taken from a process-wide pool (pinned_acquire /
// pinned_release), so the H2D copy of the sequences is a real asynchronous DMA and repeated calls
// do not pay cudaHostAlloc again (the reference's input path: std::cin >> genes[i],
// submit/xuliny-seqalkway.cpp:44-48, then MPI_Bcast :249-266).
struct SeqPack {
  uint8_t *bytes = nullptr;    // pinned, `size` bytes used of `cap`
  size_t size = 0, cap = 0;
  std::vector<uint32_t> off;
  std::vector<int32_t> len;
  int nletters = 256;          // distinct byte values in the input
  uint8_t code_of[256];        // <= 4 letters: letter -> 8*(0..3) (bit offset of its cost-table byte)
  bool has_underscore = false; // '_' (the gap character) occurs inside a sequence
  int k = 0;
  SeqPack() = default;
  SeqPack(const SeqPack &) = delete;
  SeqPack &operator=(const SeqPack &) = delete;
  ~SeqPack();
  void build(const char *const *seqs, const int32_t *lens, int k);
};
// Page-locked staging buffers, cached between calls; pinned_clear() frees them (seqalk_shutdown).
uint8_t *pinned_acquire(size_t bytes, size_t *cap_out);
void pinned_release(uint8_t *p, size_t cap);
void pinned_clear();

struct TaskIJ {
  int32_t i, j;  // x = seq[i] (rows), y = seq[j] (columns)
};

struct JobStats {
  uint64_t cells = 0;          // sum of m*n over the job's tasks
  uint64_t dir_bytes = 0;      // packed direction bytes written
  uint32_t batches = 0;
  uint32_t fill_launches = 0, traceback_launches = 0, digest_launches = 0;
  double fill_ms = 0;          // sum of CUDA-event durations of the fill launches
  double traceback_ms = 0, digest_ms = 0;
  double wall_ms = 0;          // host wall time of run()
  uint32_t tile_mode = 0;      // 1: score-only fill + tile-recompute traceback; 0: directions in HBM
  uint32_t fill_rows_per_lane = 0;
  uint32_t fill_packed16 = 0;  // 1: the 16x2-packed fill (fill16.cu) ran the tile-mode fill
  uint32_t self_checked = 0;   // 1: every penalty was compared on the device with the fill's own dp[m][n]
};

class DeviceWorker {
 public:
  explicit DeviceWorker(int device);
  ~DeviceWorker();
  DeviceWorker(const DeviceWorker &) = delete;
  DeviceWorker &operator=(const DeviceWorker &) = delete;

  int device() const { return dev_; }
  int nslots() const { return nslots_; }
  static int slots_from_env();
  // H2D of the packed sequences (kept resident until the next load()).
  void load(const SeqPack &pack);
  // Plans batches for `tasks`; result slot of tasks[t] is t unless result_slots is given.
  // `chunk_ends` (optional, increasing, last == tasks.size()) cuts the job into chunks that
  // can be run one at a time with run_chunk(); a batch never straddles a chunk end.
  void plan(const std::vector<TaskIJ> &tasks, int pxy, int pgap,
            const std::vector<uint32_t> *result_slots = nullptr,
            const std::vector<size_t> *chunk_ends = nullptr);
  // Runs every batch.  Results are written to (penalties_dev[slot], digests_dev[64*slot]);
  // both may be null (-> worker-owned device arrays, read back with fetch()) or point
  // into another GPU's memory when peer access is enabled (fused gather over NVLink).
  void run(int32_t *penalties_dev = nullptr, uint8_t *digests_dev = nullptr);
  // Score-only pass (reference result dp[m][n], submit/xuliny-seqalkway.cpp:560): penalties
  // without alignment strings or digests.  With <= 4 letters the fill alone produces them (no
  // traceback); other alphabets run fill + traceback and return the path cost.
  void run_scores(int32_t *penalties_dev = nullptr);
  // Runs only chunk c (see plan); the pointers address the WHOLE job's result arrays.
  void run_chunk(size_t c, int32_t *penalties_dev = nullptr, uint8_t *digests_dev = nullptr);
  size_t num_chunks() const { return chunk_batch_.empty() ? 0 : chunk_batch_.size() - 1; }
  // Asynchronous form: a library thread enqueues every chunk back to back (no gap between
  // chunks on the GPU).  wait_chunk(c, stream) returns once chunk c is fully enqueued and makes
  // `stream` (a cudaStream_t of the caller, e.g. the one its NCCL gather runs on; null = block
  // the host instead) wait for the chunk's results.  finish_async() joins and rethrows errors.
  void start_async(int32_t *penalties_dev, uint8_t *digests_dev);
  void wait_chunk(size_t c, cudaStream_t consumer);
  void finish_async();
  // Single-GPU end-to-end path: runs every batch with worker-owned result arrays, copies each
  // batch's results to the host as soon as its digest kernel is done and folds the hash chain
  // (submit/xuliny-seqalkway.cpp:334-337) in task order while later batches are still on the
  // GPU.  Requires identity result slots.  penalties_host[T], digest_hex_out[129].
  void run_and_fold(int32_t *penalties_host, char digest_hex_out[129]);
  // D2H of worker-owned results.
  void fetch(int32_t *penalties_host, uint8_t *digests_host);
  // Alignment strings of the LAST batch's pair `index_in_job` (only valid when the job is
  // a single batch; used by seqalk_pair).  Returns false if not available.
  bool fetch_alignment(size_t index_in_job, std::vector<uint8_t> &a1, std::vector<uint8_t> &a2);

  const JobStats &stats() const { return stats_; }
  size_t num_tasks() const { return tasks_.size(); }

 private:
  struct Batch {
    size_t first, count;
    uint64_t dir_bytes, brow_ints, out_bytes;
    uint32_t nstrips;
    uint64_t budget_bytes;
  };
  struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
  };
  struct Slot {
    cudaEvent_t fill_done = nullptr;   // fill of the batch using this slot finished
    cudaEvent_t post_done = nullptr;   // traceback + digest (+ D2H) of that batch finished
    bool post_pending = false;
    DevBuf pairs, tasks, slots, progress, brow, dirs, out, align, cost, counter;
    DevBuf band_tiles, band_tables, band_state, band_visits;  // banded traceback of long pairs (traceback_tiles.cu)
    void *h_stage = nullptr;  // pinned staging: pairs | tasks | slots
    size_t h_cap = 0;
    cudaEvent_t staged = nullptr;  // H2D of the staging buffer finished
    bool staged_pending = false;
    std::vector<PairDesc> last_pairs;  // descriptors of the last batch run on this slot
  };
  void ensure(DevBuf &b, size_t bytes);
  void ensure_stage(Slot &s, size_t bytes);
  void release(DevBuf &b);
  bool dropped_cached_ = false;  // ensure() released cached big arrays of other slots to make room

  int dev_;
  int sm_count_ = 0;
  size_t dir_budget_ = 0;
  size_t mem_avail_ = 0;  // free device memory + our own cached big arrays, asked once (plan)
  int nslots_ = 2;
  std::vector<Slot> slots_;
  cudaStream_t fill_stream_ = nullptr, post_stream_ = nullptr;
  cudaStream_t fill2_stream_ = nullptr;  // tile mode with overlap: batches alternate between fill_stream_ and this one
  std::vector<cudaStream_t> more_streams_;  // ... and these, with more than two pipelines (SEQALK_SLOTS)
  std::vector<cudaStream_t> pipe_post_streams_;  // tile mode with overlap: traceback + digest of pipeline p (high priority)
  bool overlap_ = false;  // tile mode: two full pipelines (buffer sets + streams), so that a batch's kernel tails are
                          // filled by the next batch's kernels
  DevBuf seq_, codes_, sel16_, seqmeta_;
  const SeqPack *pack_ = nullptr;
  SeqPack pack_copy_;
  std::vector<TaskIJ> tasks_;
  std::vector<uint32_t> result_slots_;
  std::vector<uint32_t> order_;  // scratch: longest-first order of the current batch
  std::vector<std::pair<uint64_t, uint64_t>> queue_keys_;  // scratch: (cells to the end of the pair, pair << 32 | strip)
  std::vector<Batch> batches_;
  int pxy_ = 0, pgap_ = 0;
  FillVariant variant_;
  bool few_pairs_ = false;  // fewer strips than ~2x the resident warps: wavefront-bound job
  bool mixed_ = false;  // tile mode, R = 32: half-height last strips (second fill launch)
  int fill16_flags_ = 0, fill16_minb_ = 0, fill16_cap_ = 0;  // one-launch packed fill: build (FILL16_*), register budget, resident CTAs per SM
  int main_rp_ = 16;      // packed fill: rows per half lane of the main strip pairs (32: 2048-row strip pairs + 1024 / 512-row tails)
  StripGeom geometry(int m) const {
    if (fill16_shift_ && main_rp_ <= 8) return strip_geometry(m, 2 * main_rp_, false);  // uniform strip pairs of 64 * main_rp_ rows
    return (fill16_shift_ && main_rp_ == 32) ? strip_geometry3(m) : strip_geometry(m, variant_.R, mixed_);
  }
  int fill16_shift_ = 0;  // > 0: the 16x2-packed fill (fill16.cu) runs the tile-mode fill; value = log2 of its re-base interval
  bool tiles_ = false;  // tile mode: score-only fill + checkpoints, directions recomputed per tile
  Batch need_{0, 0, 0, 0, 0, 0, 0};  // per-buffer maxima over the job's batches
  DevBuf res_pen_, res_dig_;
  void run_impl(int32_t *penalties_dev, uint8_t *digests_dev, int32_t *fold_penalties, char *fold_hex,
                size_t batch_begin, size_t batch_end, bool scores_only = false);
  // where the last run wrote its results (fetch() can only serve worker-owned arrays)
  bool results_in_caller_pen_ = false, results_in_caller_dig_ = false, have_digests_ = false;
  bool ran_ = false;           // a pass over the planned job has completed since plan()
  bool strip_scan_ = false;    // the input contains '_'
  DevBuf check_;               // self-check counters: [0] mismatches, [1] smallest failing task id
  void verify_check();
  uint64_t pair_base_log_ints(uint32_t nstrips, int n) const {
    // rounded to 4 ints: the next pair's kept rows are read and written as int4
    return ((uint64_t)nstrips * (uint64_t)base_log_entries(macro_cols(n), fill16_shift_) + 3) / 4 * 4;
  }         // throws if the device-side fill/traceback cross-check failed
  std::vector<size_t> chunk_batch_;  // chunk c = batches [chunk_batch_[c], chunk_batch_[c+1])
  std::vector<cudaEvent_t> chunk_ev_;  // results of chunk c are in the result arrays
  std::thread async_thread_;
  std::mutex async_mu_;
  std::condition_variable async_cv_;
  size_t async_enqueued_ = 0;  // chunks fully enqueued by the async thread
  bool async_active_ = false;
  std::string async_error_;
  void *h_res_ = nullptr;  // pinned host mirror of the result arrays: T*4 | T*64
  size_t h_res_cap_ = 0;
  void *h_hash_ = nullptr;  // pinned landing buffer of a batch's alignment strings + records (host hashing)
  size_t h_hash_cap_ = 0;
  void host_hash_batch(const Slot &S, const Slot &G, size_t count, size_t out_bytes, uint8_t *digests_dev,
                       const uint32_t *slots_host, cudaStream_t st);
  std::vector<cudaEvent_t> ev_;
  JobStats stats_;
};

// cells = sum over tasks of len[i]*len[j] (64-bit; testing9/test.cpp:219 overflowed here)
uint64_t total_cells(const SeqPack &pack, const std::vector<TaskIJ> &tasks);

// The reference's task order (submit/xuliny-seqalkway.cpp:282-287).
void all_tasks(int k, std::vector<TaskIJ> &out);
inline int64_t task_id_of(int i, int j) { return (int64_t)i * (i - 1) / 2 + j; }
void task_from_id(int64_t id, int32_t *i, int32_t *j);

// Cell-balanced shard of tasks [id_begin, id_end) over nranks (greedy list scheduling in id
// order onto the least loaded rank, ties by rank index), each rank's list in id order.
void partition_lpt(const int32_t *lens, int k, int64_t id_begin, int64_t id_end, int nranks,
                   std::vector<std::vector<int64_t>> &per_rank);

void throw_cuda(cudaError_t e, const char *what, const char *file, int line);
#define SEQALK_CUDA(x) ::seqalk::throw_cuda((x), #x, __FILE__, __LINE__)

}  // namespace seqalk
