/*
 * seqalk.h — C ABI of libseqalk: all-pairs minimum-penalty global alignment on B200.
 *
 * Drop-in boundary for the hot path of yangxvlin/multiple-sequence-alignment-openMP-openMPI
 * ("xuliny-seqalkway").  Plain pointers and sizes only.  Every entry point returns 0 on
 * success and non-zero on failure; seqalk_last_error() then describes the failure.  There
 * is no CPU fallback: without a usable sm_100a GPU every compute call fails.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   seqalk_all_pairs   <- std::string getMinimumPenalties(std::string *genes, int k, int pxy,
 *                         int pgap, int *penalties)      submit/xuliny-seqalkway.cpp:13, :232-364
 *   seqalk_pair        <- int getMinimumPenalty(std::string x, std::string y, int pxy, int pgap,
 *                         int *xans, int *yans)          submit/xuliny-seqalkway.cpp:14, :570-663
 *                         (and getMinimumPenalty2 :419-566)
 *   seqalk_partition   <- the master's task queue / hand-out  :280-331 (MPI_Send of Triple)
 *   seqalk_job_*       <- the worker loop recv Triple -> do_task -> send Packet  :342-357, :399-411
 *   seqalk_fold_chain  <- the ordered fold of task hashes  :334-337
 *   seqalk_sha512_hex  <- sw::sha512::calculate  sha512.hh:159-164
 * C++ wrappers with the reference's exact signatures are in seqalk.hpp.
 */
#ifndef SEQALK_H
#define SEQALK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEQALK_DIGEST_HEX_LEN 128 /* lowercase hex chars, callers pass char[129] */
#define SEQALK_DIGEST_RAW_LEN 64

/* ---- runtime ---------------------------------------------------------------------- */

/* Selects the GPUs used by seqalk_all_pairs / seqalk_pair of this process: devices
 * 0..ngpus-1 (ngpus <= 0: every visible device).  Replaces MPI_Init_thread / MPI_Comm_size
 * (submit/xuliny-seqalkway.cpp:38, :245).  Optional: the first compute call initialises
 * with SEQALK_GPUS from the environment, else 1 GPU. */
int seqalk_init(int ngpus);
/* Releases every device resource (MPI_Finalize, :75). */
void seqalk_shutdown(void);
/* Number of GPUs in use (0 before initialisation). */
int seqalk_device_count(void);
/* Message of the last failure on the calling thread ("" if none). */
const char *seqalk_last_error(void);
const char *seqalk_version(void);

/* ---- the reference's operator surface ------------------------------------------------ */

/* All k(k-1)/2 pairs in the reference's task order (for i=1..k-1, for j=0..i-1; x = seqs[i]
 * is the row sequence, y = seqs[j] the column sequence).  Writes penalties_out[T] and the
 * chained SHA-512 digest as 128 hex chars + NUL.  seqs[i] need not be NUL-terminated. */
int seqalk_all_pairs(const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy, int32_t pgap,
                     int32_t *penalties_out, char digest_hex_out[129]);

/* One pair.  xans / yans (each int32[m+n+1], may be NULL) receive the reference's
 * right-aligned alignment arrays: index 0 untouched, 1..m+n filled with character codes,
 * leading positions padded with '_'.  penalty_out may be NULL. */
int seqalk_pair(const char *x, int32_t m, const char *y, int32_t n, int32_t pxy, int32_t pgap,
                int32_t *xans, int32_t *yans, int32_t *penalty_out);

/* Score-only forms: the penalties dp[m][n] (submit/xuliny-seqalkway.cpp:560)
 * without alignment strings or digest.  With <= 4 distinct letters in the input the DP fill alone
 * produces them (no traceback, no hashing); other alphabets run fill + traceback and return the
 * path cost.  The full entry points above cross-check the two on the device for every pair. */
int seqalk_all_pairs_scores(const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy, int32_t pgap,
                            int32_t *penalties_out);
int seqalk_pair_score(const char *x, int32_t m, const char *y, int32_t n, int32_t pxy, int32_t pgap,
                      int32_t *penalty_out);

/* Phase times (milliseconds) of the last seqalk_all_pairs call of this process, for tracing the
 * scaling losses: [0] whole call, [1] set-up (pack, partition, H2D, plan), [2] from the first
 * launch until every GPU finished its last chunk, [3] hash-chain fold (and last D2H) left exposed
 * after that, [4] D2H copies of the gathered results, [5] kernel time (fill + traceback + digest,
 * CUDA events) of the busiest GPU, [6] mean kernel time over the GPUs, [7] GPUs, [8] chunks (N > 1)
 * or batches (N = 1), [9..11] fill / traceback / digest kernel time summed over the GPUs. */
#define SEQALK_PROFILE_LEN 12
int seqalk_profile(double *out, int32_t n);

/* ---- sharded execution: one process per GPU ------------------------------------------- */

/* T = k(k-1)/2 */
int64_t seqalk_num_tasks(int32_t k);
/* (i, j) of a task id in the reference's order. */
int seqalk_task_ij(int64_t task_id, int32_t *i_out, int32_t *j_out);
/* Cell-balanced (LPT on len_i*len_j) shard of task ids [id_begin, id_end) for `rank` of
 * `nranks`; ids_out (capacity id_end-id_begin) receives the rank's ids in increasing order. */
int seqalk_partition(const int32_t *lens, int32_t k, int64_t id_begin, int64_t id_end, int32_t nranks,
                     int32_t rank, int64_t *ids_out, int64_t *count_out);

/* Same shard for every rank at once: ids_out (capacity id_end-id_begin) receives rank 0's ids,
 * then rank 1's, ...; counts_out[nranks] their counts. */
int seqalk_partition_all(const int32_t *lens, int32_t k, int64_t id_begin, int64_t id_end, int32_t nranks,
                         int64_t *ids_out, int64_t *counts_out);

/* Chunk boundaries for `ntasks` tasks: bounds_out[nchunks+1], chunk c = ids
 * [bounds[c], bounds[c+1]).  Equal chunks, except that with >= 5 chunks (and >= 4096 tasks) the
 * last three hold 8 %, 5 % and 3 % of the tasks (with 4 chunks: a 4 % last chunk), so that the part
 * of the serial hash-chain fold that cannot overlap GPU work stays short. */
int seqalk_chunk_bounds(int64_t ntasks, int32_t nchunks, int64_t *bounds_out);
/* The shards of the `nchunks` id ranges of seqalk_chunk_bounds for every rank, computed in
 * parallel: ids_out (capacity T) = chunk 0 rank 0, chunk 0 rank 1,
 * ..., chunk 1 rank 0, ...; counts_out[nchunks*nranks] in the same order. */
int seqalk_partition_chunks(const int32_t *lens, int32_t k, int32_t nranks, int32_t nchunks, int64_t *ids_out,
                            int64_t *counts_out);

typedef struct seqalk_job seqalk_job_t;

typedef struct seqalk_stats {
  uint64_t cells;             /* sum of m*n over the job's tasks */
  uint64_t dir_bytes;         /* packed direction bytes written by the fill */
  uint32_t batches;
  uint32_t fill_launches, traceback_launches, digest_launches;
  double fill_ms;             /* sum of CUDA-event durations of the fill launches */
  double traceback_ms, digest_ms;
  double wall_ms;             /* host wall time of the last seqalk_job_run */
  uint32_t tile_mode;         /* 1: score-only fill + tile-recompute traceback, 0: directions in HBM */
  uint32_t fill_rows_per_lane;
  uint32_t fill_packed16;     /* 1: the tile-mode fill ran as the 16x2-packed kernel (two DP cells per register) */
  uint32_t self_checked;      /* 1: every penalty was compared on the device with the fill's own dp[m][n] */
} seqalk_stats_t;

/* Creates a job on CUDA device `device`: copies the k sequences to the GPU and plans the
 * batches for the given task ids (NULL: all T tasks).  Result slot s belongs to
 * task_ids[s]. */
int seqalk_job_create(int32_t device, const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy,
                      int32_t pgap, const int64_t *task_ids, int64_t ntasks, seqalk_job_t **job_out);
/* Runs fill + traceback + digest for every task with inputs resident in HBM.  Results go to
 * DEVICE memory: penalties_dev (int32 per slot) and digests_dev (64 raw bytes per slot); pass
 * NULL for job-owned buffers (read them back with seqalk_job_fetch). */
int seqalk_job_run(seqalk_job_t *job, void *penalties_dev, void *digests_dev);
/* Cuts the job into chunks of consecutive result slots: chunk c covers slots
 * [chunk_ends[c-1], chunk_ends[c]) (chunk_ends[nchunks-1] must equal the task count), so that
 * a caller can gather / fold chunk c while the GPU works on chunk c+1. */
int seqalk_job_set_chunks(seqalk_job_t *job, const int64_t *chunk_ends, int32_t nchunks);
/* Like seqalk_job_run for one chunk; the pointers address the WHOLE job's result arrays. */
int seqalk_job_run_chunk(seqalk_job_t *job, int32_t chunk, void *penalties_dev, void *digests_dev);
/* Asynchronous form of run over all chunks: a library thread enqueues chunk after chunk with
 * no gap on the GPU.  seqalk_job_wait_chunk returns once chunk c is fully enqueued and makes
 * consumer_stream (a cudaStream_t of the caller, e.g. the stream its NCCL gather runs on;
 * NULL = block the host until the results are there) wait for that chunk's results.
 * seqalk_job_finish_async joins the thread and reports any error. */
int seqalk_job_start_async(seqalk_job_t *job, void *penalties_dev, void *digests_dev);
int seqalk_job_wait_chunk(seqalk_job_t *job, int32_t chunk, void *consumer_stream);
int seqalk_job_finish_async(seqalk_job_t *job);
/* Score-only pass over the job's tasks (see seqalk_all_pairs_scores); penalties_dev NULL = job-owned. */
int seqalk_job_run_scores(seqalk_job_t *job, void *penalties_dev);
/* D2H of the job-owned results (fails if the last run wrote to caller-supplied device memory). */
int seqalk_job_fetch(seqalk_job_t *job, int32_t *penalties_out, uint8_t *digests_out);
int seqalk_job_stats(const seqalk_job_t *job, seqalk_stats_t *stats_out);
void seqalk_job_destroy(seqalk_job_t *job);

/* Folds `count` per-task digests (64 raw bytes each, task order) into the running chain
 * state: state_hex is "" before the first task and the 128-char digest afterwards. */
int seqalk_fold_chain(const uint8_t *pair_digests, int64_t count, char state_hex[129]);

/* SHA-512 as 128 lowercase hex chars + NUL (host). */
int seqalk_sha512_hex(const void *data, size_t len, char out[129]);

/* ---- measurement ---------------------------------------------------------------------- */

/* Dependent-free issue-rate microbenchmarks on `device`, in lane-operations per second:
 * [0] LOP3 (ALU pipe), [1] VIMNMX3 (3-input min), [2] IMAD (FMA pipe), [3] LOP3+IMAD
 * alternating (dual pipe), [4] VIADDMNMX (fused add+min), [5] SHF (funnel shift),
 * [6] LOP3+VIMNMX3 alternating, [7] the directions-mode fill mix (1 VIMNMX3 : 1 SHF : 2 IMAD),
 * [8] IDP.4A (byte dot product + add), [9] IDP.4A+VIMNMX3 alternating, [10] IDP.4A+IMAD
 * alternating, [11] PRMT, [12] VIMNMX3.U16x2 (two 16-bit max3 per lane; counted as one lane-op),
 * [13] the packed fill's mix PRMT + IMAD + VIMNMX3.U16x2 on different registers, [14] IDP.4A +
 * VIMNMX3 with three different register operands per instruction (register-file bandwidth bound,
 * unlike [9] whose operands sit in the reuse cache), [15] reserved.
 * rates_out[0] is the integer-ALU roofline denominator of SURVEY.md section 8d. */
int seqalk_measure_int_peaks(int32_t device, double rates_out[16]);

#ifdef __cplusplus
}
#endif
#endif /* SEQALK_H */
