// engine.cu — see engine.h
#include "engine.h"
#include "sha512_host.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <queue>
#include <stdexcept>

namespace seqalk {

void throw_cuda(cudaError_t e, const char *what, const char *file, int line) {
  if (e == cudaSuccess) return;
  char buf[512];
  std::snprintf(buf, sizeof buf, "CUDA error %s (%d) at %s:%d: %s", cudaGetErrorString(e), (int)e, file, line,
                what);
  throw std::runtime_error(buf);
}

static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t EV = 6;  // timing events per batch
// bounds the descriptor staging of a batch of very many tiny pairs (56 B + strips per pair)
constexpr size_t MAX_PAIRS_PER_BATCH = 1u << 18;

namespace {
struct PinnedPool {
  std::mutex mu;
  std::vector<std::pair<uint8_t *, size_t>> idle;
};
PinnedPool &pinned_pool() {
  static PinnedPool p;
  return p;
}
}  // namespace

uint8_t *pinned_acquire(size_t bytes, size_t *cap_out) {
  PinnedPool &pp = pinned_pool();
  {
    std::lock_guard<std::mutex> g(pp.mu);
    size_t best = pp.idle.size();
    for (size_t i = 0; i < pp.idle.size(); ++i)
      if (pp.idle[i].second >= bytes && (best == pp.idle.size() || pp.idle[i].second < pp.idle[best].second)) best = i;
    if (best != pp.idle.size()) {
      uint8_t *p = pp.idle[best].first;
      *cap_out = pp.idle[best].second;
      pp.idle.erase(pp.idle.begin() + (long)best);
      return p;
    }
  }
  void *p = nullptr;
  const size_t want = round_up(bytes + bytes / 8 + 4096, 4096);
  cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw std::runtime_error(std::string("cannot allocate page-locked staging memory (") + cudaGetErrorString(e) +
                             "): libseqalk needs a CUDA device, there is no CPU fallback");
  }
  *cap_out = want;
  return static_cast<uint8_t *>(p);
}

void pinned_release(uint8_t *p, size_t cap) {
  if (!p) return;
  PinnedPool &pp = pinned_pool();
  std::lock_guard<std::mutex> g(pp.mu);
  if (pp.idle.size() >= 4) {  // keep a handful; drop the smallest
    size_t small = 0;
    for (size_t i = 1; i < pp.idle.size(); ++i)
      if (pp.idle[i].second < pp.idle[small].second) small = i;
    if (pp.idle[small].second < cap) {
      cudaFreeHost(pp.idle[small].first);
      pp.idle[small] = {p, cap};
    } else {
      cudaFreeHost(p);
    }
    return;
  }
  pp.idle.emplace_back(p, cap);
}

void pinned_clear() {
  PinnedPool &pp = pinned_pool();
  std::lock_guard<std::mutex> g(pp.mu);
  for (auto &b : pp.idle) cudaFreeHost(b.first);
  pp.idle.clear();
}

SeqPack::~SeqPack() { pinned_release(bytes, cap); }

void SeqPack::build(const char *const *seqs, const int32_t *lens, int k_) {
  k = k_;
  off.resize(k);
  len.assign(lens, lens + k);
  size_t total = 0;
  for (int i = 0; i < k; ++i) {
    if (lens[i] < 0) throw std::runtime_error("negative sequence length");
    if (lens[i] > 0 && !seqs[i]) throw std::runtime_error("null sequence pointer");
    off[i] = (uint32_t)total;
    total += round_up((size_t)lens[i], 16) + SEQ_PAD;
    if (total > 0xfff00000ull) throw std::runtime_error("input larger than 4 GiB is not supported");
  }
  size = total + BUFFER_SLACK;
  if (bytes) pinned_release(bytes, cap);
  bytes = pinned_acquire(size, &cap);
  std::memset(bytes, 0, size);
  for (int i = 0; i < k; ++i)
    if (lens[i]) std::memcpy(bytes + off[i], seqs[i], (size_t)lens[i]);
  // alphabet: with <= 4 distinct letters (DNA) the fill can use a per-row cost table; the
  // letters are remapped on the GPU (remap_letters) from this 256-entry table
  // (eight tables: consecutive stores to ONE table hit the same few bytes and serialise on
  // store forwarding — 9-14 ms for the 10 MB of BASELINE config 4 against 1.5-2 ms this way)
  uint8_t seen8[8][256];
  std::memset(seen8, 0, sizeof seen8);
  for (int i = 0; i < k; ++i) {
    const uint8_t *p = reinterpret_cast<const uint8_t *>(seqs[i]);
    int32_t c = 0;
    for (; c + 8 <= lens[i]; c += 8)
      for (int b = 0; b < 8; ++b) seen8[b][p[c + b]] = 1;  // branch-free pass over the input
    for (; c < lens[i]; ++c) seen8[0][p[c]] = 1;
  }
  uint8_t seen[256];
  for (int ch = 0; ch < 256; ++ch) {
    uint8_t v = 0;
    for (int b = 0; b < 8; ++b) v |= seen8[b][ch];
    seen[ch] = v;
  }
  nletters = 0;
  std::memset(code_of, 0, sizeof code_of);
  for (int ch = 0; ch < 256; ++ch)
    if (seen[ch]) {
      if (nletters < 4) code_of[ch] = (uint8_t)(8 * nletters);  // stored as the byte's bit offset
      ++nletters;
    }
  has_underscore = seen[(unsigned char)'_'] != 0;
}

namespace {
struct Lut256 {
  uint8_t v[256];
};
__global__ void remap_letters_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n16,
                                     const Lut256 lut) {
  // 16 bytes per thread per iteration; padding bytes (0) map to code_of[0] like any letter
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = s4[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) o |= (uint32_t)lut.v[(w[k] >> (8 * b)) & 0xffu] << (8 * b);
      w[k] = o;
    }
    d4[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
}  // namespace

uint64_t total_cells(const SeqPack &pack, const std::vector<TaskIJ> &tasks) {
  uint64_t c = 0;
  for (const TaskIJ &t : tasks) c += (uint64_t)pack.len[t.i] * (uint64_t)pack.len[t.j];
  return c;
}

void all_tasks(int k, std::vector<TaskIJ> &out) {
  out.clear();
  out.reserve((size_t)k * (k - 1) / 2);
  for (int i = 1; i < k; ++i)
    for (int j = 0; j < i; ++j) out.push_back(TaskIJ{i, j});
}

void task_from_id(int64_t id, int32_t *i, int32_t *j) {
  // largest i with i(i-1)/2 <= id
  int64_t ii = (int64_t)((1.0 + std::sqrt(1.0 + 8.0 * (double)id)) / 2.0);
  while (ii * (ii - 1) / 2 > id) --ii;
  while ((ii + 1) * ii / 2 <= id) ++ii;
  *i = (int32_t)ii;
  *j = (int32_t)(id - ii * (ii - 1) / 2);
}

void partition_lpt(const int32_t *lens, int k, int64_t id_begin, int64_t id_end, int nranks,
                   std::vector<std::vector<int64_t>> &per_rank) {
  // Greedy list scheduling in id order: every task goes to the rank with the fewest cells so
  // far (ties: lowest rank).  The imbalance is bounded by one task (<= 4e8 cells of ~1e12 per
  // rank and chunk at k = 1000, i.e. < 0.05 %), each rank's ids come out increasing without a
  // sort, and the whole k = 1000 plan costs a few milliseconds.
  (void)k;
  per_rank.assign(nranks, {});
  if (id_end <= id_begin) return;
  for (int r = 0; r < nranks; ++r) per_rank[r].reserve((size_t)((id_end - id_begin) / nranks + 16));
  std::vector<uint64_t> load((size_t)nranks, 0);
  int32_t i, j;
  task_from_id(id_begin, &i, &j);
  for (int64_t id = id_begin; id < id_end; ++id) {
    int best = 0;
    for (int r = 1; r < nranks; ++r)
      if (load[r] < load[best]) best = r;
    load[best] += (uint64_t)lens[i] * (uint64_t)lens[j];
    per_rank[best].push_back(id);
    if (++j == i) {
      ++i;
      j = 0;
    }
  }
}

// ------------------------------------------------------------------ DeviceWorker

DeviceWorker::DeviceWorker(int device) : dev_(device) {
  SEQALK_CUDA(cudaSetDevice(dev_));
  cudaDeviceProp prop;
  SEQALK_CUDA(cudaGetDeviceProperties(&prop, dev_));
  if (prop.major < 10)
    throw std::runtime_error(std::string("seqalk needs an sm_100a GPU (B200); found ") + prop.name);
  sm_count_ = prop.multiProcessorCount;
  nslots_ = slots_from_env();
  slots_.resize(nslots_);
  int prio_lo = 0, prio_hi = 0;
  SEQALK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  SEQALK_CUDA(cudaStreamCreateWithPriority(&fill_stream_, cudaStreamNonBlocking, prio_lo));
  SEQALK_CUDA(cudaStreamCreateWithPriority(&post_stream_, cudaStreamNonBlocking, prio_hi));
  SEQALK_CUDA(cudaStreamCreateWithPriority(&fill2_stream_, cudaStreamNonBlocking, prio_lo));
  for (int i = 2; i < nslots_; ++i) {
    cudaStream_t st = nullptr;
    SEQALK_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_lo));
    more_streams_.push_back(st);
  }
  for (int i = 0; i < nslots_; ++i) {
    cudaStream_t st = nullptr;
    SEQALK_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio_hi));
    pipe_post_streams_.push_back(st);
  }
  for (Slot &s : slots_) {
    SEQALK_CUDA(cudaEventCreateWithFlags(&s.staged, cudaEventDisableTiming));
    SEQALK_CUDA(cudaEventCreateWithFlags(&s.fill_done, cudaEventDisableTiming));
    SEQALK_CUDA(cudaEventCreateWithFlags(&s.post_done, cudaEventDisableTiming));
  }
}

int DeviceWorker::slots_from_env() {
  int n = 3;  // tile mode: three overlapped batch pipelines (dirs mode: buffer sets)
  if (const char *e = std::getenv("SEQALK_SLOTS")) n = std::max(1, std::min(4, std::atoi(e)));
  return n;
}

void DeviceWorker::release(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

DeviceWorker::~DeviceWorker() {
  if (async_thread_.joinable()) async_thread_.join();
  cudaSetDevice(dev_);
  for (cudaEvent_t e : chunk_ev_) cudaEventDestroy(e);
  cudaDeviceSynchronize();
  for (Slot &s : slots_) {
    for (DevBuf *b : {&s.pairs, &s.tasks, &s.slots, &s.progress, &s.brow, &s.dirs, &s.out, &s.align, &s.cost,
                      &s.counter, &s.band_tiles, &s.band_tables, &s.band_state, &s.band_visits})
      release(*b);
    if (s.h_stage) cudaFreeHost(s.h_stage);
    if (s.staged) cudaEventDestroy(s.staged);
    if (s.fill_done) cudaEventDestroy(s.fill_done);
    if (s.post_done) cudaEventDestroy(s.post_done);
  }
  if (fill_stream_) cudaStreamDestroy(fill_stream_);
  if (post_stream_) cudaStreamDestroy(post_stream_);
  if (fill2_stream_) cudaStreamDestroy(fill2_stream_);
  for (cudaStream_t st : more_streams_) cudaStreamDestroy(st);
  for (cudaStream_t st : pipe_post_streams_) cudaStreamDestroy(st);
  release(seq_);
  release(codes_);
  release(sel16_);
  release(seqmeta_);
  release(res_pen_);
  release(res_dig_);
  release(check_);
  if (h_res_) cudaFreeHost(h_res_);
  if (h_hash_) cudaFreeHost(h_hash_);
  for (cudaEvent_t e : ev_) cudaEventDestroy(e);
}

void DeviceWorker::ensure(DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return;
  // growing a buffer: nothing may still be using the old one
  SEQALK_CUDA(cudaDeviceSynchronize());
  release(b);
  size_t want = round_up(bytes + bytes / 16, 256);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = round_up(bytes, 256);
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    // the big arrays cached from an earlier job with another buffer configuration (more
    // pipelines, smaller sets) may be in the way: nothing is in flight here, drop them all — the
    // slots that are still needed grow theirs again
    cudaGetLastError();
    for (Slot &s : slots_)
      for (DevBuf *o : {&s.brow, &s.dirs, &s.out})
        if (o != &b && o->p) {
          release(*o);
          dropped_cached_ = true;
        }
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    mem_avail_ = 0;  // the next plan() looks at the free memory again
    char msg[256];
    std::snprintf(msg, sizeof msg, "getMinimumPenalty: device allocation of %zu bytes failed (%s)", want,
                  cudaGetErrorString(e));
    throw std::runtime_error(msg);
  }
  b.cap = want;
}

void DeviceWorker::ensure_stage(Slot &s, size_t bytes) {
  if (bytes <= s.h_cap) return;
  if (s.staged_pending) {
    SEQALK_CUDA(cudaEventSynchronize(s.staged));
    s.staged_pending = false;
  }
  if (s.h_stage) SEQALK_CUDA(cudaFreeHost(s.h_stage));
  s.h_stage = nullptr;
  const size_t want = round_up(bytes + bytes / 8, 4096);
  SEQALK_CUDA(cudaMallocHost(&s.h_stage, want));
  s.h_cap = want;
}

void DeviceWorker::load(const SeqPack &pack) {
  SEQALK_CUDA(cudaSetDevice(dev_));
  // lengths/offsets are needed for planning; the bytes stay with the caller
  pack_copy_.k = pack.k;
  pack_copy_.off = pack.off;
  pack_copy_.len = pack.len;
  pack_copy_.nletters = pack.nletters;
  pack_copy_.has_underscore = pack.has_underscore;
  pack_ = &pack_copy_;
  strip_scan_ = pack.has_underscore;
  ensure(seq_, (pack.size + 15) / 16 * 16);  // the letter remap reads whole 16 B words
  // source is page-locked (SeqPack): a true asynchronous DMA
  SEQALK_CUDA(cudaMemcpyAsync(seq_.p, pack.bytes, pack.size, cudaMemcpyHostToDevice, fill_stream_));
  if (pack.nletters <= 4) {
    const size_t n16 = (pack.size + 15) / 16;
    ensure(codes_, n16 * 16);
    Lut256 lut;
    std::memcpy(lut.v, pack.code_of, 256);
    remap_letters_kernel<<<(unsigned)std::min<size_t>((n16 + 255) / 256, 4096), 256, 0, fill_stream_>>>(
        static_cast<const uint8_t *>(seq_.p), static_cast<uint8_t *>(codes_.p), n16, lut);
    SEQALK_CUDA(cudaGetLastError());
    // selector streams of the 16x2-packed fill (fill16.cu): 8 bytes per macro column of every
    // sequence, at twice the sequence's offset
    ensure(sel16_, 2 * n16 * 16);
    ensure(seqmeta_, (size_t)pack.k * 8 + 16);
    uint32_t *off_dev = static_cast<uint32_t *>(seqmeta_.p);
    int32_t *len_dev = reinterpret_cast<int32_t *>(off_dev + pack.k);
    SEQALK_CUDA(cudaMemcpyAsync(off_dev, pack.off.data(), (size_t)pack.k * 4, cudaMemcpyHostToDevice, fill_stream_));
    SEQALK_CUDA(cudaMemcpyAsync(len_dev, pack.len.data(), (size_t)pack.k * 4, cudaMemcpyHostToDevice, fill_stream_));
    int maxlen = 0;
    for (int32_t L : pack.len) maxlen = std::max(maxlen, L);
    SEQALK_CUDA(launch_build_sel16(static_cast<const uint8_t *>(codes_.p), static_cast<uint8_t *>(sel16_.p), off_dev,
                                   len_dev, pack.k, maxlen, fill_stream_));
  }
  SEQALK_CUDA(cudaStreamSynchronize(fill_stream_));
}

void DeviceWorker::plan(const std::vector<TaskIJ> &tasks, int pxy, int pgap,
                        const std::vector<uint32_t> *result_slots, const std::vector<size_t> *chunk_ends) {
  if (!pack_) throw std::runtime_error("plan() before load()");
  finish_async();  // never re-plan under a running asynchronous pass
  ran_ = false;
  tasks_ = tasks;
  if (result_slots) {
    if (result_slots->size() != tasks.size()) throw std::runtime_error("result_slots size mismatch");
    result_slots_ = *result_slots;
  } else {
    result_slots_.clear();
  }
  pxy_ = pxy;
  pgap_ = pgap;
  batches_.clear();
  stats_ = JobStats();
  // kernel variant
  variant_ = FillVariant();
  variant_.R = 16;  // measured best on B200 (profiles/): 16 rows per lane, cost table when possible
  variant_.addmode = 2;  // cost table lookup + add as one IDP.4A (FMA pipe) when the table applies
  if (const char *e = std::getenv("SEQALK_FILL_R")) variant_.R = std::atoi(e) == 8 ? 8 : 16;
  if (const char *e = std::getenv("SEQALK_FILL_MINB")) variant_.minb = std::atoi(e);
  if (const char *e = std::getenv("SEQALK_FILL_ADD")) variant_.addmode = std::max(0, std::min(2, std::atoi(e)));
  // traceback mode: "tiles" = score-only fill + per-tile recompute (less work per cell, 8x
  // less memory; measured faster on every BASELINE config, profiles/), "dirs" = 2-bit
  // directions for every cell in HBM (SEQALK_TRACEBACK=dirs).
  tiles_ = true;
  if (const char *e = std::getenv("SEQALK_TRACEBACK")) {
    if (!std::strcmp(e, "tiles")) tiles_ = true;
    if (!std::strcmp(e, "dirs")) tiles_ = false;
  }
  if (tiles_) {
    variant_.R = 32;  // measured best for the score-only fill (profiles/): 9.0 vs 7.9 TCUPS at R = 16
    if (const char *e = std::getenv("SEQALK_FILL_R"))
      if (std::atoi(e) == 16) variant_.R = 16;
  }
  variant_.emit = !tiles_;
  {  // checkpoint / direction budget per buffer set (larger batches = fewer kernel tails), capped
     // by the device's memory.  dirs mode: 56 GB per set (the fill of batch b+1 runs beside the
     // traceback of batch b).  Tile mode: one set per pipeline — 3 x 40 GB by default
     // (SEQALK_SLOTS / SEQALK_DIR_BUDGET_GB), 104 GB without overlap.
    SEQALK_CUDA(cudaSetDevice(dev_));
    // our own (cached) buffers count as available, so "free + own" does not change from job to job:
    // asked once per worker — cudaMemGetInfo costs 0.5 - 15 ms with tens of GB allocated, which showed
    // in the end-to-end time of the 40 ms big13 job (SEQALK_TRACE=2)
    auto ask = [&]() {
      size_t free_b = 0, total_b = 0;
      SEQALK_CUDA(cudaMemGetInfo(&free_b, &total_b));
      size_t own = 0;
      for (const Slot &s : slots_) own += s.dirs.cap + s.brow.cap;
      mem_avail_ = free_b + own;
    };
    if (!mem_avail_) ask();
    // tile mode: consecutive batches go to different pipelines (own set of the big arrays, own
    // streams), so the tail of one batch's fill / traceback / digest is filled by the next
    // batches' kernels (SEQALK_OVERLAP=0: one set, stages back to back).  One GPU's share of an
    // 8-GPU cfg4 job, emulated on one GPU (scripts/shard_span.py): 704 ms without overlap, 640 ms
    // with two pipelines, 610 ms with three and the traceback / digest on high-priority streams.
    overlap_ = tiles_ && nslots_ >= 2;
    if (const char *e = std::getenv("SEQALK_OVERLAP")) overlap_ = overlap_ && std::atoi(e) != 0;
    const int nbig = tiles_ ? (overlap_ ? nslots_ : 1) : nslots_;
    // three pipelines of 40 GB: a tenth of an 8-GPU job's tasks per GPU still fits one batch
    double budget_gb = tiles_ ? (nbig >= 3 ? 40.0 : 104.0 / nbig) : 56.0;
    if (const char *e = std::getenv("SEQALK_DIR_BUDGET_GB")) budget_gb = std::atof(e);
    dir_budget_ = (size_t)(budget_gb * (double)(1ull << 30));
    size_t cap = (size_t)((double)mem_avail_ * 0.80 / nbig);
    if (dir_budget_ > cap) {  // memory is tight (a shared device, other workers' caches): look again
      ask();
      cap = (size_t)((double)mem_avail_ * 0.80 / nbig);
    }
    if (dir_budget_ > cap) dir_budget_ = cap;
  }
  mixed_ = tiles_;  // 512-row last strips where they halve the row padding
  if (const char *e = std::getenv("SEQALK_MIXED")) mixed_ = tiles_ && std::atoi(e) != 0;
  for (const TaskIJ &tk : tasks_)
    if (tk.i < 0 || tk.i >= pack_->k || tk.j < 0 || tk.j >= pack_->k) throw std::runtime_error("task index out of range");
  {
    // few long pairs: the GPU is filled by the strips of the same pair running as a wavefront,
    // so use thinner strips (more of them) and publish progress more often
    uint64_t strips = 0;
    for (const TaskIJ &tk : tasks_) strips += (uint64_t)num_strips(std::max(pack_->len[tk.i], 1), variant_.R);
    few_pairs_ = strips < 2ull * (uint64_t)sm_count_ * 16;
  }
  int64_t maxlen = 1;
  for (const TaskIJ &tk : tasks_) maxlen = std::max<int64_t>(maxlen, std::max(pack_->len[tk.i], pack_->len[tk.j]));
  // fast path needs non-negative penalties and scores far from int32 overflow
  // (|H'| <= 2*max(pxy,pgap)*(m+n)); anything else takes the exact-select kernel.
  const int64_t pmax = std::max<int64_t>(std::max(pxy, pgap), 1);
  bool general = pxy < 0 || pgap < 0 || pmax * 4 * maxlen >= (1ll << 29);
  if (const char *e = std::getenv("SEQALK_FORCE_GENERAL"))
    if (std::atoi(e)) general = true;
  // cost table (PRMT) needs <= 4 letters and both shifted costs in a signed byte
  bool table = pack_->nletters <= 4 && 2 * (int64_t)pgap <= 127 && (int64_t)pxy - 2 * (int64_t)pgap <= 127 &&
               (int64_t)pxy - 2 * (int64_t)pgap >= -128;
  if (const char *e = std::getenv("SEQALK_FILL_TAB"))
    if (!std::atoi(e)) table = false;
  variant_.mode = general ? MODE_GENERAL : (table ? MODE_TAB : MODE_CMP);
  // 16x2-packed fill (fill16.cu): tile mode, cost table, 1024 / 512-row strip pairs (the geometry
  // of the 32-rows-per-lane fill), penalties small enough for 16-bit relative scores
  fill16_shift_ = 0;
  // 2048-row strip pairs (+ 1024 / 512-row tails): less bookkeeping per cell than 1024-row ones;
  // few-pair jobs take 1024-row ones: more strip pairs in flight (big13 fill 46 -> 41 ms against the
  // 32-bit fill, the lone 200k x 200k pair 26 -> 24 ms; with 2048-row pairs 31 ms)
  main_rp_ = few_pairs_ ? 16 : 32;
  // Kernel build and residency (same-box A/B, profiles/r02b_*):
  //  * many pairs: predicated checkpoint stores, 3 CTAs per SM (134 registers): cfg4 job 4.69 -> 4.61 s;
  //  * few long pairs, more strip pairs than the GPU has warp schedulers (big13): 1024-row strip pairs,
  //    predicated checkpoint stores, only 2 CTAs per SM — the queue is ordered critical path first,
  //    and fewer resident warps means the earliest (most critical) strip pairs run at full speed:
  //    fill 40.5 ms (6 CTAs per SM, branch) -> 35.5 (4, predicated) -> 30.5 (2);
  //  * a handful of long pairs whose 512-row strip pairs all find a scheduler of their own (the lone
  //    200k x 200k pair): 512-row strip pairs, predicated stores, pipelined hand-off: fill 24.5 -> 16.8 ms.
  fill16_flags_ = few_pairs_ ? (FILL16_PREDCK | FILL16_SENT) : FILL16_PREDCK;
  fill16_minb_ = 3;
  fill16_cap_ = few_pairs_ ? 2 : 0;
  if (few_pairs_) {
    uint64_t thin = 0;  // 512-row strip pairs of the job
    for (const TaskIJ &tk : tasks_) thin += (uint64_t)(std::max(pack_->len[tk.i], 1) + 511) / 512;
    if (thin <= 6ull * (uint64_t)sm_count_) {
      main_rp_ = 8;
      fill16_flags_ = FILL16_PREDCK | FILL16_PIPE | FILL16_SENT;
      fill16_minb_ = 4;
      fill16_cap_ = 0;
    }
  }
  if (const char *e = std::getenv("SEQALK_FILL16_RP")) {
    const int v = std::atoi(e);
    main_rp_ = (v == 16 || v == 8 || v == 4) ? v : 32;  // 8 / 4: strip pairs of 512 / 256 rows (one-launch kernel only)
    if (main_rp_ != 8) fill16_flags_ &= ~FILL16_PIPE;
  }
  if (const char *e = std::getenv("SEQALK_FILL16_PREDCK")) fill16_flags_ = (fill16_flags_ & ~FILL16_PREDCK) | (std::atoi(e) ? FILL16_PREDCK : 0);
  if (const char *e = std::getenv("SEQALK_FILL16_PIPE")) fill16_flags_ = (fill16_flags_ & ~FILL16_PIPE) | (std::atoi(e) ? FILL16_PIPE : 0);
  if (const char *e = std::getenv("SEQALK_FILL16_SENT")) fill16_flags_ = (fill16_flags_ & ~FILL16_SENT) | (std::atoi(e) ? FILL16_SENT : 0);
  if (main_rp_ == 4) fill16_flags_ = 0;
  if ((main_rp_ != 8 && main_rp_ != 16) || !(fill16_flags_ & FILL16_PREDCK)) fill16_flags_ &= ~FILL16_SENT;
  if (main_rp_ == 8 && !(fill16_flags_ & FILL16_PIPE)) fill16_flags_ &= ~FILL16_SENT;  // (no such build)
  if (!(fill16_flags_ & FILL16_PREDCK)) fill16_minb_ = 0;  // the branch builds: their own defaults (4 / 6 / 8 CTAs per SM)
  if (variant_.minb > 0) fill16_minb_ = variant_.minb;     // SEQALK_FILL_MINB
  if (const char *e = std::getenv("SEQALK_FILL_CTAS_PER_SM")) fill16_cap_ = std::max(0, std::atoi(e));
  if (tiles_ && variant_.mode == MODE_TAB && variant_.R == 32 && mixed_ && codes_.p) {
    fill16_shift_ = fill16_rebase_shift(pxy, pgap, main_rp_);
    if (!fill16_shift_ && main_rp_ == 32) {
      main_rp_ = 16;
      fill16_shift_ = fill16_rebase_shift(pxy, pgap, main_rp_);
    }
  }
  if (const char *e = std::getenv("SEQALK_FILL16"))
    if (!std::atoi(e)) fill16_shift_ = 0;
  // few long pairs without the packed fill: the GPU is filled by the strips of the same pair running
  // as a wavefront, so use thinner strips (more of them) and publish progress more often
  if (!fill16_shift_ && few_pairs_ && tiles_ && !std::getenv("SEQALK_FILL_R")) variant_.R = 16;
  const int R = variant_.R;
  Batch cur{0, 0, 0, 0, 0, 0, 0};
  std::vector<size_t> ends;
  if (chunk_ends) ends = *chunk_ends;
  if (ends.empty() || ends.back() != tasks_.size()) ends.push_back(tasks_.size());
  chunk_batch_.assign(1, 0);
  size_t next_end = 0;
  for (size_t t = 0; t < tasks_.size(); ++t) {
    while (next_end < ends.size() && ends[next_end] <= t) {  // chunk boundary: close the batch
      if (cur.count) {
        batches_.push_back(cur);
        cur = Batch{t, 0, 0, 0, 0, 0, 0};
      }
      chunk_batch_.push_back(batches_.size());
      ++next_end;
    }
    const TaskIJ &tk = tasks_[t];
    if (tk.i < 0 || tk.i >= pack_->k || tk.j < 0 || tk.j >= pack_->k) throw std::runtime_error("task index out of range");
    const int m = pack_->len[tk.i], n = pack_->len[tk.j];
    // an empty sequence (undefined in the parallel reference: ceil(0/0) tile size at
    // submit/xuliny-seqalkway.cpp:457; the sequential twin :570-663 returns (m+n)*pgap with a
    // pure-tail alignment) has no DP cells: no strips, the traceback only writes the tails
    StripGeom geo = geometry(m);
    if (m < 1 || n < 1) geo.nstrips = 0, geo.half_last = 0;
    const uint32_t ns = (uint32_t)geo.nstrips;
    uint64_t brow_ints = tiles_ ? (uint64_t)(geo.mpad / TILE_ROWS) * (4ull * (uint64_t)macro_cols(n) + 4)
                                : 4ull * (uint64_t)macro_cols(n) + 4;
    if (fill16_shift_) brow_ints += pair_base_log_ints(ns, n);  // base logs of the strip pairs, behind the kept rows
    // directions, or (tile mode) the column checkpoints [nq / CKQ][rows padded to the strips]
    const uint64_t db = tiles_ ? ck_region_bytes(macro_cols(n), geo.mpad, fill16_shift_ != 0) + 32
                               : (uint64_t)ns * strip_dir_bytes(n, R);
    // what counts against the budget: directions, or (tile mode) checkpoints + kept rows
    const uint64_t bb = tiles_ ? db + brow_ints * 4 : db;
    if (cur.count && (cur.budget_bytes + bb > dir_budget_ || (uint64_t)cur.nstrips + ns > 0x7fffffffull ||
                      cur.count >= MAX_PAIRS_PER_BATCH)) {
      batches_.push_back(cur);
      cur = Batch{t, 0, 0, 0, 0, 0, 0};
    }
    cur.count++;
    cur.dir_bytes += db;
    cur.budget_bytes += bb;
    cur.brow_ints += brow_ints;
    cur.out_bytes += 2ull * round_up((size_t)m + n, 16);
    cur.nstrips += ns;
    stats_.cells += (uint64_t)m * (uint64_t)n;
    stats_.dir_bytes += bb;
  }
  if (cur.count) batches_.push_back(cur);
  while (next_end < ends.size()) {
    chunk_batch_.push_back(batches_.size());
    ++next_end;
  }
  stats_.batches = (uint32_t)batches_.size();
  // one event per chunk, created here: wait_chunk() reads this vector from the caller's thread
  // while the enqueue thread of start_async() is running
  while (chunk_ev_.size() + 1 < chunk_batch_.size()) {
    cudaEvent_t e;
    SEQALK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    chunk_ev_.push_back(e);
  }
  stats_.tile_mode = tiles_ ? 1 : 0;
  stats_.fill_packed16 = fill16_shift_ ? 1 : 0;
  stats_.fill_rows_per_lane = fill16_shift_ ? (uint32_t)main_rp_ : (uint32_t)variant_.R;
  // buffer sizes: the largest batch decides, so that run() never reallocates mid-job
  need_ = Batch{0, 0, 0, 0, 0, 0, 0};
  for (const Batch &b : batches_) {
    need_.count = std::max(need_.count, b.count);
    need_.dir_bytes = std::max(need_.dir_bytes, b.dir_bytes);
    need_.brow_ints = std::max(need_.brow_ints, b.brow_ints);
    need_.out_bytes = std::max(need_.out_bytes, b.out_bytes);
    need_.nstrips = std::max(need_.nstrips, b.nstrips);
  }
}

void DeviceWorker::run(int32_t *penalties_dev, uint8_t *digests_dev) {
  run_impl(penalties_dev, digests_dev, nullptr, nullptr, 0, batches_.size());
}

void DeviceWorker::run_chunk(size_t c, int32_t *penalties_dev, uint8_t *digests_dev) {
  if (c + 1 >= chunk_batch_.size()) throw std::runtime_error("chunk index out of range");
  run_impl(penalties_dev, digests_dev, nullptr, nullptr, chunk_batch_[c], chunk_batch_[c + 1]);
}

void DeviceWorker::run_scores(int32_t *penalties_dev) {
  run_impl(penalties_dev, nullptr, nullptr, nullptr, 0, batches_.size(), true);
}

void DeviceWorker::verify_check() {
  if (!check_.p) return;
  uint32_t c[2] = {0, 0};
  SEQALK_CUDA(cudaMemcpy(c, check_.p, sizeof c, cudaMemcpyDeviceToHost));
  if (c[0]) {
    char msg[256];
    std::snprintf(msg, sizeof msg,
                  "internal check failed: for %u task(s) the traceback's path cost differs from the fill's "
                  "dp[m][n] (first: task %u of this job)", c[0], c[1]);
    throw std::runtime_error(msg);
  }
}

void DeviceWorker::run_and_fold(int32_t *penalties_host, char digest_hex_out[129]) {
  if (!result_slots_.empty()) throw std::runtime_error("run_and_fold needs identity result slots");
  run_impl(nullptr, nullptr, penalties_host, digest_hex_out, 0, batches_.size());
}

void DeviceWorker::run_impl(int32_t *penalties_dev, uint8_t *digests_dev, int32_t *fold_penalties,
                            char *fold_hex, size_t batch_begin, size_t batch_end, bool scores_only) {
  SEQALK_CUDA(cudaSetDevice(dev_));
  const bool streaming = fold_hex != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  // SEQALK_TRACE=2: host-side time stamps of this run on stderr (ms since its start)
  const bool trace = [] {
    const char *e = std::getenv("SEQALK_TRACE");
    return e && std::atoi(e) >= 2;
  }();
  std::vector<std::pair<const char *, double>> marks;
  auto mark = [&](const char *what) {
    if (trace)
      marks.push_back({what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count()});
  };
  const int R = variant_.R;
  const size_t T = tasks_.size();
  results_in_caller_pen_ = penalties_dev != nullptr;
  results_in_caller_dig_ = digests_dev != nullptr;
  have_digests_ = !scores_only;
  if (!penalties_dev) {
    ensure(res_pen_, std::max<size_t>(T, 1) * sizeof(int32_t));
    penalties_dev = static_cast<int32_t *>(res_pen_.p);
  }
  if (!digests_dev && !scores_only) {
    ensure(res_dig_, std::max<size_t>(T, 1) * 64);
    digests_dev = static_cast<uint8_t *>(res_dig_.p);
  }
  // The fill's own dp[m][n] is available in table mode (row_operand, fill_core.cuh): the digest
  // kernel compares it with the traceback's path cost for every pair (SEQALK_SELFCHECK=0: off).
  int corner_mode = variant_.mode == MODE_TAB ? (tiles_ ? CORNER_TILES : CORNER_DIRS) : CORNER_NONE;
  const bool fill_only = scores_only && corner_mode != CORNER_NONE;
  if (!scores_only)
    if (const char *e = std::getenv("SEQALK_SELFCHECK"))
      if (!std::atoi(e)) corner_mode = CORNER_NONE;
  ensure(check_, 2 * sizeof(uint32_t));
  if (batch_begin == 0) {
    const uint32_t init[2] = {0u, 0xffffffffu};
    SEQALK_CUDA(cudaMemcpy(check_.p, init, sizeof init, cudaMemcpyHostToDevice));  // before any stream's digest kernel
    stats_.self_checked = (!scores_only && corner_mode != CORNER_NONE) ? 1 : 0;
  }
  int32_t *h_pen = nullptr;
  uint8_t *h_dig = nullptr;
  size_t next_fold = batch_begin, hex_len = 0;
  if (streaming) {
    const size_t want = std::max<size_t>(T, 1) * 68;
    if (want > h_res_cap_) {
      if (h_res_) SEQALK_CUDA(cudaFreeHost(h_res_));
      h_res_ = nullptr;
      SEQALK_CUDA(cudaMallocHost(&h_res_, want));
      h_res_cap_ = want;
    }
    h_pen = static_cast<int32_t *>(h_res_);
    h_dig = static_cast<uint8_t *>(h_res_) + std::max<size_t>(T, 1) * 4;
    fold_hex[0] = 0;
  }
  // Folds the batches below `limit` whose results have landed.  `limit` must not exceed the
  // batches enqueued in THIS run: the events of later batches still carry their (completed)
  // record of the previous run.
  auto fold_ready = [&](bool wait, size_t limit) {
    while (next_fold < limit) {
      cudaEvent_t done = ev_[EV * next_fold + 5];
      if (wait) {
        SEQALK_CUDA(cudaEventSynchronize(done));
      } else {
        cudaError_t q = cudaEventQuery(done);
        if (q == cudaErrorNotReady) break;
        SEQALK_CUDA(q);
      }
      const Batch &FB = batches_[next_fold];
      fold_chain(h_dig + 64 * FB.first, FB.count, fold_hex, &hex_len);
      std::memcpy(fold_penalties + FB.first, h_pen + FB.first, FB.count * sizeof(int32_t));
      ++next_fold;
    }
  };
  // events per batch: fill start/end (fill stream); traceback start/end, digest end,
  // results on host (post stream)
  while (ev_.size() < EV * batches_.size()) {
    cudaEvent_t e;
    SEQALK_CUDA(cudaEventCreate(&e));
    ev_.push_back(e);
  }

  if (batch_begin == 0) {  // timing counters cover one pass over the job's batches
    stats_.fill_ms = stats_.traceback_ms = stats_.digest_ms = stats_.wall_ms = 0;
    stats_.fill_launches = stats_.traceback_launches = stats_.digest_launches = 0;
  }
  for (size_t b = batch_begin; b < batch_end; ++b) {
    const Batch &B = batches_[b];
    Slot &S = slots_[b % nslots_];
    // tile mode: one stream, so one set of the big arrays serves every batch (the descriptor
    // buffers stay per slot: the host fills the next batch's staging while this one runs)
    Slot &G = (tiles_ && !overlap_) ? slots_[0] : S;
    // the batch's stream (tile mode): its whole pipeline runs there
    const int pipe = overlap_ ? (int)(b % nslots_) : 0;
    cudaStream_t tile_stream = pipe == 0 ? fill_stream_ : pipe == 1 ? fill2_stream_ : more_streams_[(size_t)pipe - 2];
    // ---- size device buffers (for the largest batch of the job: at most one allocation)
    do {  // (again if an allocation had to drop cached arrays of this very set)
      dropped_cached_ = false;
      ensure(S.pairs, need_.count * sizeof(PairDesc));
      ensure(S.tasks, (size_t)need_.nstrips * sizeof(StripTask));
      ensure(S.slots, need_.count * sizeof(uint32_t));
      ensure(S.progress, (size_t)need_.nstrips * sizeof(uint32_t));
      ensure(G.brow, need_.brow_ints * sizeof(int32_t));
      ensure(G.dirs, need_.dir_bytes);
      ensure(G.out, need_.out_bytes + BUFFER_SLACK);
      ensure(G.align, need_.count * sizeof(PairAlign));
      ensure(G.cost, need_.count * sizeof(int32_t));
      ensure(S.counter, 4 * sizeof(uint32_t));
    } while (dropped_cached_);
    mark("buffers sized");
    // ---- build descriptors in pinned staging
    const size_t pairs_bytes = round_up(need_.count * sizeof(PairDesc), 256);
    const size_t tasks_bytes = round_up((size_t)need_.nstrips * sizeof(StripTask), 256);
    const size_t slots_bytes = round_up(need_.count * sizeof(uint32_t), 256);
    ensure_stage(S, pairs_bytes + tasks_bytes + slots_bytes);
    if (S.staged_pending) {
      SEQALK_CUDA(cudaEventSynchronize(S.staged));
      S.staged_pending = false;
    }
    PairDesc *hp = reinterpret_cast<PairDesc *>(S.h_stage);
    StripTask *ht = reinterpret_cast<StripTask *>(static_cast<uint8_t *>(S.h_stage) + pairs_bytes);
    uint32_t *hs = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(S.h_stage) + pairs_bytes + tasks_bytes);
    uint64_t dir_off = 0, brow_off = 0, out_off = 0;
    uint32_t strip0 = 0;
    int max_strips = 0;
    // inside a batch the pairs are laid out longest first (m + n): the traceback kernel's
    // blocks and every level of the strip queue then start their longest work first, which
    // shortens the tails; results still land at the task's slot
    order_.resize(B.count);
    for (size_t p = 0; p < B.count; ++p) order_[p] = (uint32_t)p;
    std::stable_sort(order_.begin(), order_.end(), [&](uint32_t a, uint32_t c) {
      const TaskIJ &ta = tasks_[B.first + a], &tc = tasks_[B.first + c];
      return (int64_t)pack_->len[ta.i] + pack_->len[ta.j] > (int64_t)pack_->len[tc.i] + pack_->len[tc.j];
    });
    for (size_t p = 0; p < B.count; ++p) {
      const size_t t_index = B.first + order_[p];
      const TaskIJ &tk = tasks_[t_index];
      PairDesc d;
      d.x_off = pack_->off[tk.i];
      d.y_off = pack_->off[tk.j];
      d.m = pack_->len[tk.i];
      d.n = pack_->len[tk.j];
      d.dir_off = dir_off;
      d.brow_off = brow_off;
      d.out_off = out_off;
      d.strip0 = strip0;
      StripGeom geo = geometry(d.m);
      if (d.m < 1 || d.n < 1) geo.nstrips = 0, geo.half_last = 0;
      d.nstrips = geo.nstrips;
      d.mpad = geo.mpad;
      d.half_last = geo.half_last;
      d.cap = (uint32_t)round_up((size_t)d.m + d.n, 16);
      d.task_id = (int32_t)t_index;
      hp[p] = d;
      hs[p] = result_slots_.empty() ? (uint32_t)t_index : result_slots_[t_index];
      if (tiles_) {
        dir_off += round_up(ck_region_bytes(macro_cols(d.n), d.mpad, fill16_shift_ != 0), 32);  // 256-bit checkpoint stores
        brow_off += (uint64_t)(d.mpad / TILE_ROWS) * (4ull * (uint64_t)macro_cols(d.n) + 4);
        if (fill16_shift_) brow_off += pair_base_log_ints((uint32_t)d.nstrips, d.n);
      } else {
        dir_off += (uint64_t)d.nstrips * strip_dir_bytes(d.n, R);
        brow_off += 4ull * (uint64_t)macro_cols(d.n) + 4;
      }
      out_off += 2ull * d.cap;
      strip0 += (uint32_t)d.nstrips;
      max_strips = std::max(max_strips, d.nstrips);
    }
    // strip queue: a strip's predecessor is always claimed earlier (no deadlock without a
    // cooperative launch).  Default order = critical path first (big13 fill 46.0 -> 44.5 ms, no
    // difference on many-pair jobs).
    // Shorter last strips (tile mode) go to the end: they run in later launches (1024-row strip
    // pairs of the three-level geometry with 16 rows per half lane, then the 512-row ones), after
    // all their predecessors have finished.
    size_t n_main = 0, n_mid = 0, n_half = 0;
    // packed fill: one persistent launch takes the main strip pairs and the shorter last ones
    // (SEQALK_FILL16_SPLIT=1: one launch per strip height, as the 32-bit fill does)
    const bool split16 = [] {
      const char *e = std::getenv("SEQALK_FILL16_SPLIT");
      return e && std::atoi(e) != 0;
    }();
    const bool one_launch = fill16_shift_ && (!split16 || main_rp_ <= 8);  // thin strip pairs: the one-launch kernel only
    if (one_launch) {
      // one queue for every strip pair of the batch, critical path first: ordered by the DP cells
      // from the strip pair's first row to the end of its pair (a predecessor has more of them, so
      // it is still claimed earlier; the last tasks of the queue are the smallest last strip pairs)
      queue_keys_.clear();
      for (size_t p = 0; p < B.count; ++p) {
        const int ns = hp[p].nstrips, nm = main_strips(ns, hp[p].half_last);
        for (int s = 0; s < ns; ++s) {
          int row_base;
          if (s < nm) row_base = s * 64 * main_rp_;
          else if (main_rp_ == 32 && (hp[p].half_last & 2) && s == nm) row_base = nm * 2048;
          else row_base = (int)hp[p].mpad - 512;
          const uint64_t rem = (uint64_t)(hp[p].mpad - (uint32_t)row_base) * (uint64_t)hp[p].n;
          queue_keys_.push_back({rem, ((uint64_t)p << 32) | (uint32_t)s});
        }
      }
      std::sort(queue_keys_.begin(), queue_keys_.end(), [](const std::pair<uint64_t, uint64_t> &a, const std::pair<uint64_t, uint64_t> &c) {
        return a.first != c.first ? a.first > c.first : a.second < c.second;
      });
      for (const auto &kv : queue_keys_) ht[n_main++] = StripTask{(uint32_t)(kv.second >> 32), (uint32_t)kv.second};
    } else {
      const bool by_remaining = [] {
        const char *e = std::getenv("SEQALK_QUEUE_REMAINING");  // 0: plain (strip, pair) order
        return !e || std::atoi(e) != 0;
      }();
      if (by_remaining) {
        // critical path first: strips ordered by the number of main strips still to come in their
        // pair (a predecessor has more of them, so it is still claimed earlier)
        for (int r = max_strips; r >= 1; --r)
          for (size_t p = 0; p < B.count; ++p) {
            const int ms = main_strips(hp[p].nstrips, hp[p].half_last);
            if (ms >= r) ht[n_main++] = StripTask{(uint32_t)p, (uint32_t)(ms - r)};
          }
      } else {
        for (int s = 0; s < max_strips; ++s)
          for (size_t p = 0; p < B.count; ++p)
            if (s < main_strips(hp[p].nstrips, hp[p].half_last)) ht[n_main++] = StripTask{(uint32_t)p, (uint32_t)s};
      }
      for (size_t p = 0; p < B.count; ++p)
        if (hp[p].half_last & 2)
          ht[n_main + n_mid++] = StripTask{(uint32_t)p, (uint32_t)main_strips(hp[p].nstrips, hp[p].half_last)};
      for (size_t p = 0; p < B.count; ++p)
        if (hp[p].half_last & 1) ht[n_main + n_mid + n_half++] = StripTask{(uint32_t)p, (uint32_t)(hp[p].nstrips - 1)};
    }
    S.last_pairs.assign(hp, hp + B.count);
    mark("descriptors built");
    // ---- enqueue: fill stream
    cudaStream_t st = tiles_ ? tile_stream : fill_stream_;
    if (S.post_pending) {  // the slot's buffers are free once its previous batch is post-processed
      SEQALK_CUDA(cudaStreamWaitEvent(st, S.post_done, 0));
      S.post_pending = false;
    }
    SEQALK_CUDA(cudaMemcpyAsync(S.pairs.p, hp, B.count * sizeof(PairDesc), cudaMemcpyHostToDevice, st));
    SEQALK_CUDA(cudaMemcpyAsync(S.tasks.p, ht, (size_t)B.nstrips * sizeof(StripTask), cudaMemcpyHostToDevice, st));
    SEQALK_CUDA(cudaMemcpyAsync(S.slots.p, hs, B.count * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    SEQALK_CUDA(cudaEventRecord(S.staged, st));
    S.staged_pending = true;
    SEQALK_CUDA(cudaMemsetAsync(S.counter.p, 0, 4 * sizeof(uint32_t), st));
    SEQALK_CUDA(cudaMemsetAsync(S.progress.p, 0, (size_t)B.nstrips * sizeof(uint32_t), st));

    FillArgs fa;
    fa.seq = static_cast<const uint8_t *>(seq_.p);
    fa.codes = static_cast<const uint8_t *>(codes_.p);
    fa.one = 1;
    fa.publish_shift = few_pairs_ ? PUBLISH_SHIFT_FEW : PUBLISH_SHIFT_MANY;
    fa.pairs = static_cast<const PairDesc *>(S.pairs.p);
    fa.tasks = static_cast<const StripTask *>(S.tasks.p);
    fa.ntasks = (uint32_t)n_main;
    fa.counter = static_cast<uint32_t *>(S.counter.p);
    fa.progress = static_cast<uint32_t *>(S.progress.p);
    fa.brow = static_cast<int32_t *>(G.brow.p);
    fa.dirs = static_cast<uint8_t *>(G.dirs.p);
    fa.pxy = pxy_;
    fa.pgap = pgap_;
    fa.sel16 = static_cast<const uint8_t *>(sel16_.p);
    fa.rebase_shift = fill16_shift_;
    fa.main_rp = main_rp_;
    SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 0], st));
    if (one_launch && (fill16_flags_ & FILL16_SENT)) {
      // sentinel hand-off (fill16.cu): the kept rows that pass from one strip pair to the next start out as
      // FILL16_SENTINEL words
      uint64_t max_words = 0;
      for (size_t p = 0; p < B.count; ++p)
        if (hp[p].nstrips > 1) max_words = std::max<uint64_t>(max_words, (uint64_t)(hp[p].nstrips - 1) * 4 * (uint64_t)macro_cols(hp[p].n));
      if (max_words) {
        SEQALK_CUDA(launch_fill16_sentinels(fa.pairs, (uint32_t)B.count, fa.brow, main_rp_, max_words, st));
        stats_.fill_launches++;
      }
    }
    if (one_launch) {
      SEQALK_CUDA(launch_fill16(fa, 100 + main_rp_, fill16_minb_, sm_count_, st, fill16_flags_, fill16_cap_));
    } else if (fill16_shift_) {
      SEQALK_CUDA(launch_fill16(fa, main_rp_, variant_.minb, sm_count_, st));
    } else {
      SEQALK_CUDA(launch_fill(fa, variant_, sm_count_, st));
    }
    if (n_mid) {  // second launch (three-level geometry): the 1024-row strip pairs
      FillArgs fm = fa;
      fm.tasks = fa.tasks + n_main;
      fm.ntasks = (uint32_t)n_mid;
      fm.counter = fa.counter + 2;
      SEQALK_CUDA(launch_fill16(fm, 16, 0, sm_count_, st));
      stats_.fill_launches++;
    }
    if (n_half) {  // last launch: the 512-row last strips
      FillArgs fh = fa;
      fh.tasks = fa.tasks + n_main + n_mid;
      fh.ntasks = (uint32_t)n_half;
      fh.counter = fa.counter + 1;
      if (fill16_shift_) {
        SEQALK_CUDA(launch_fill16(fh, 8, 0, sm_count_, st));
      } else {
        FillVariant vh = variant_;
        vh.R = 16;
        vh.minb = 0;
        SEQALK_CUDA(launch_fill(fh, vh, sm_count_, st));
      }
      stats_.fill_launches++;
    }
    SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 1], st));
    SEQALK_CUDA(cudaEventRecord(S.fill_done, st));
    mark("fill enqueued");
    // ---- enqueue: post stream (high priority; runs under the next batch's fill).  In tile
    // mode the traceback is a heavy kernel of its own: the stages run back to back instead.
    // Tile mode, overlapped pipelines: traceback + digest of a batch go to the pipeline's own
    // HIGH-priority stream, so that when SMs free up they are placed before the CTAs of a later
    // batch's persistent fill (otherwise a batch's results wait for the fill after next)
    const bool post_prio = [] {
      const char *e = std::getenv("SEQALK_POST_PRIORITY");
      return !e || std::atoi(e) != 0;
    }();
    st = tiles_ ? ((overlap_ && post_prio) ? pipe_post_streams_[(size_t)pipe] : tile_stream) : post_stream_;
    SEQALK_CUDA(cudaStreamWaitEvent(st, S.fill_done, 0));
    SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 2], st));

    TracebackArgs ta;
    ta.seq = fa.seq;
    ta.pairs = fa.pairs;
    ta.npairs = (uint32_t)B.count;
    ta.dirs = fa.dirs;
    ta.out = static_cast<uint8_t *>(G.out.p);
    ta.align = static_cast<PairAlign *>(G.align.p);
    ta.path_cost = static_cast<int32_t *>(G.cost.p);
    ta.pxy = pxy_;
    ta.pgap = pgap_;
    ta.R = R;
    ta.codes = fa.codes;
    ta.brow = fa.brow;
    ta.mode = variant_.mode;
    ta.addmode = variant_.addmode;
    ta.ck16 = fill16_shift_;
    ta.main_rows = 64 * main_rp_;
    ta.long_pairs = few_pairs_ ? 1 : 0;
    ta.strip_scan = strip_scan_ ? 1 : 0;
    if (const char *e = std::getenv("SEQALK_TB_LONG")) ta.long_pairs = std::atoi(e) ? 1 : 0;
    ta.half_warp = 1;  // two pairs per warp (SEQALK_TB_HALF=0: one warp per pair)
    if (const char *e = std::getenv("SEQALK_TB_HALF")) ta.half_warp = std::atoi(e) ? 1 : 0;
    // 256-row strip pairs: the 8 rows of a half-warp lane would span two virtual lanes of the fill
    // (two re-base histories); the whole-warp tracebacks read 4 rows per lane
    if (fill16_shift_ && main_rp_ == 4) ta.half_warp = 0;
    // a few LONG pairs: banded traceback in chapters (traceback_tiles.cu) — all tiles near the path
    // recomputed at once, one warp each, instead of one after the other (the 200k x 200k pair:
    // 38 ms -> see profiles/).  SEQALK_TB_BAND=0 / 1 forces it off / on (tests run it on small inputs),
    // SEQALK_TB_BAND_LEVELS / SEQALK_TB_BAND_HALFWIDTH set the chapter length and the band.
    ta.band = 0;
    ta.band_levels = ta.band_halfwidth = ta.band_chapters = ta.band_visit_cap = 0;
    ta.band_tiles = nullptr;
    ta.band_tables = nullptr;
    ta.band_state = nullptr;
    ta.band_visits = nullptr;
    if (tiles_ && !fill_only) {
      int64_t longest = 0;
      int max_level = 0;
      for (size_t p = 0; p < B.count; ++p) {
        longest = std::max<int64_t>(longest, (int64_t)hp[p].m + hp[p].n);
        if (hp[p].m > 0 && hp[p].n > 0)
          max_level = std::max(max_level, (hp[p].m - 1) / TILE_ROWS + (hp[p].n - 1) / (4 * CKQ));
      }
      bool band = ta.long_pairs && B.count <= 512 && longest >= 32768;
      if (const char *e = std::getenv("SEQALK_TB_BAND")) band = std::atoi(e) != 0 && B.count <= 4096;
      if (band) {
        ta.band = 1;
        // a handful of pairs: long chapters, wide band (the produce launch is far from filling the GPU);
        // more pairs: the produce launches are the cost — short chapters, three tiles per level
        // (same box, traceback of the 200k x 200k pair: 3.6 ms at 63 / 4, 2.8 ms at 127 / 4; of the 78 pairs
        // of big13: 8.5 ms at 63 / 2, 6.8 ms at 31 / 1)
        ta.band_levels = B.count <= 8 ? 127 : 31;
        ta.band_halfwidth = B.count <= 8 ? 4 : 1;
        if (const char *e = std::getenv("SEQALK_TB_BAND_LEVELS")) ta.band_levels = std::max(0, std::min(1023, std::atoi(e)));
        if (const char *e = std::getenv("SEQALK_TB_BAND_HALFWIDTH")) ta.band_halfwidth = std::max(0, std::min(16, std::atoi(e)));
        // chapters: what a path that stays inside its bands needs, plus a quarter for chapters cut
        // short by a turn of the path; the last one finishes whatever is left
        const int nominal = max_level / (ta.band_levels + 1) + 1;
        ta.band_chapters = nominal + nominal / 4 + 4;
        if (const char *e = std::getenv("SEQALK_TB_BAND_CHAPTERS")) ta.band_chapters = std::max(1, std::atoi(e));
        ta.band_visit_cap = ta.band_levels + 2;
        const size_t images = (size_t)(ta.band_levels + 1) * (size_t)(2 * ta.band_halfwidth + 1);
        ensure(G.band_tiles, B.count * images * (size_t)(CKQ * TILE_ROWS));
        ensure(G.band_tables, B.count * images * (size_t)(TILE_ROWS + 4 * CKQ) * sizeof(uint32_t));
        ensure(G.band_state, B.count * sizeof(BandState));
        ensure(G.band_visits, B.count * (size_t)ta.band_visit_cap * sizeof(BandVisit));
        ta.band_tiles = static_cast<uint8_t *>(G.band_tiles.p);
        ta.band_tables = static_cast<uint32_t *>(G.band_tables.p);
        ta.band_state = static_cast<BandState *>(G.band_state.p);
        ta.band_visits = static_cast<BandVisit *>(G.band_visits.p);
        stats_.traceback_launches += 3 * (uint32_t)ta.band_chapters;  // + the finish launch, counted below
      }
    }
    if (!fill_only) SEQALK_CUDA(tiles_ ? launch_traceback_tiles(ta, st) : launch_traceback(ta, st));
    SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 3], st));
    mark("traceback enqueued");

    if (scores_only) {
      ScoreArgs sa;
      sa.pairs = fa.pairs;
      sa.npairs = (uint32_t)B.count;
      sa.brow = fa.brow;
      sa.corner_mode = corner_mode;
      sa.pgap = pgap_;
      sa.path_cost = ta.path_cost;
      sa.result_slot = static_cast<const uint32_t *>(S.slots.p);
      sa.penalties_out = penalties_dev;
      SEQALK_CUDA(launch_scores(sa, st));
    } else {
      DigestArgs da;
      da.pairs = fa.pairs;
      da.npairs = (uint32_t)B.count;
      da.out = ta.out;
      da.align = ta.align;
      da.penalties = ta.path_cost;
      da.result_slot = static_cast<const uint32_t *>(S.slots.p);
      da.digests = digests_dev;
      da.penalties_out = penalties_dev;
      da.brow = fa.brow;
      da.corner_mode = corner_mode;
      da.pgap = pgap_;
      da.check = static_cast<uint32_t *>(check_.p);
      // the 64-register build: alone it is 12 % slower than the 91-register one (cfg4: 72 vs 63 ms),
      // but four of its CTAs fit exactly into the registers one exiting fill CTA frees, and the
      // overlapped job is 1.6 % faster with it (4.70 s vs 4.78 s, same box); SEQALK_DIGEST_CAP=0
      da.small_regs = 1;
      if (const char *e = std::getenv("SEQALK_DIGEST_CAP")) da.small_regs = std::atoi(e) ? 1 : 0;
      // A few very long strings: SHA-512 is serial per string, and one GPU lane hashes ~27 MB/s
      // (less when the lanes of a warp hash strings of different lengths) against ~260 MB/s of a host core.  Such a batch's
      // strings are brought to the host, hashed there by a few threads and the 64-byte digests
      // uploaded; the kernel then only publishes the penalties and runs the self-check.
      // SEQALK_HOST_HASH=0 / 1 forces it off / on.
      da.host_hash = 0;
      {
        uint64_t longest = 0, total = 0;
        for (size_t p = 0; p < B.count; ++p) {
          const uint64_t len = (uint64_t)hp[p].m + (uint64_t)hp[p].n;
          longest = std::max(longest, len);
          total += 2 * len;
        }
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const double threads = (double)std::min<uint64_t>(std::min<uint64_t>(hw, 16), 2 * B.count);
        // measured: the 200k x 200k pair 15.3 ms -> 0.6 ms, the 78 pairs of big13 (18.7 MB of strings) 10.2 -> 3.2 ms
        // (only where it pays clearly: small jobs stay on the GPU)
        const double device_ns = 55.0 * (double)longest, host_ns = 2.5 * (double)total / threads + 200e3;
        bool host = few_pairs_ && B.count > 0 && device_ns > 1.5e6 && 1.5 * host_ns < device_ns;
        if (const char *e = std::getenv("SEQALK_HOST_HASH")) host = std::atoi(e) != 0 && B.count > 0;
        if (host) {
          host_hash_batch(S, G, B.count, B.out_bytes, digests_dev, hs, st);
          da.host_hash = 1;
        }
      }
      SEQALK_CUDA(launch_digest(da, st));
    }
    SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 4], st));
    mark("digest enqueued (after the host hashes, if any)");
    if (streaming) {
      SEQALK_CUDA(cudaMemcpyAsync(h_pen + B.first, penalties_dev + B.first, B.count * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, st));
      SEQALK_CUDA(cudaMemcpyAsync(h_dig + 64 * B.first, digests_dev + 64 * B.first, B.count * 64,
                                  cudaMemcpyDeviceToHost, st));
      SEQALK_CUDA(cudaEventRecord(ev_[EV * b + 5], st));
    }
    SEQALK_CUDA(cudaEventRecord(S.post_done, st));
    S.post_pending = true;
    // chunk boundary: mark the chunk's results with an event and tell waiters
    for (size_t c = 0; c + 1 < chunk_batch_.size(); ++c)
      if (chunk_batch_[c + 1] == b + 1 && chunk_batch_[c + 1] > chunk_batch_[c]) {
        // with two pipelines the chunk's earlier batches may still run on the other stream
        if (overlap_)
          for (Slot &O : slots_)
            if (&O != &S && O.post_pending) SEQALK_CUDA(cudaStreamWaitEvent(st, O.post_done, 0));
        SEQALK_CUDA(cudaEventRecord(chunk_ev_[c], st));
        {
          std::lock_guard<std::mutex> g(async_mu_);
          async_enqueued_ = c + 1;
        }
        async_cv_.notify_all();
      }
    if (streaming) fold_ready(false, b + 1);  // fold what has landed while the GPU works on later batches
    stats_.fill_launches++;
    if (!fill_only) stats_.traceback_launches++;
    stats_.digest_launches++;
  }
  mark("all batches enqueued");
  if (streaming) fold_ready(true, batch_end);
  mark("results folded");
  SEQALK_CUDA(cudaStreamSynchronize(fill_stream_));
  SEQALK_CUDA(cudaStreamSynchronize(post_stream_));
  SEQALK_CUDA(cudaStreamSynchronize(fill2_stream_));
  for (cudaStream_t st : more_streams_) SEQALK_CUDA(cudaStreamSynchronize(st));
  for (cudaStream_t st : pipe_post_streams_) SEQALK_CUDA(cudaStreamSynchronize(st));
  for (Slot &S : slots_) S.staged_pending = S.post_pending = false;
  mark("streams idle");
  if (batch_end == batches_.size() && !scores_only) verify_check();
  ran_ = true;
  mark("self-check read");
  for (const auto &m : marks) std::fprintf(stderr, "[seqalk run] %8.3f ms  %s\n", m.second, m.first);
  for (size_t b = batch_begin; b < batch_end; ++b) {
    float ms = 0;
    SEQALK_CUDA(cudaEventElapsedTime(&ms, ev_[EV * b + 0], ev_[EV * b + 1]));
    stats_.fill_ms += ms;
    SEQALK_CUDA(cudaEventElapsedTime(&ms, ev_[EV * b + 2], ev_[EV * b + 3]));
    stats_.traceback_ms += ms;
    SEQALK_CUDA(cudaEventElapsedTime(&ms, ev_[EV * b + 3], ev_[EV * b + 4]));
    stats_.digest_ms += ms;
  }
  stats_.wall_ms +=
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
}

// Host side of the per-pair hash for a batch of a few very long pairs (same result as
// digest_kernel: problemhash = H(hex(H(align1)) ++ hex(H(align2))), submit/xuliny-seqalkway.cpp:217-219).
// Runs on the batch's post stream after the traceback: strings + records to the host, hash, digests
// back to the slots of the job-wide array (which may live on a peer GPU).
void DeviceWorker::host_hash_batch(const Slot &S, const Slot &G, size_t count, size_t out_bytes, uint8_t *digests_dev,
                                   const uint32_t *slots_host, cudaStream_t st) {
  const size_t rec_off = round_up(out_bytes, 64), dig_off = rec_off + round_up(count * sizeof(PairAlign), 64);
  const size_t want = dig_off + count * 64;
  if (want > h_hash_cap_) {
    if (h_hash_) SEQALK_CUDA(cudaFreeHost(h_hash_));
    h_hash_ = nullptr;
    h_hash_cap_ = 0;
    SEQALK_CUDA(cudaMallocHost(&h_hash_, want + want / 8));
    h_hash_cap_ = want + want / 8;
  }
  uint8_t *h = static_cast<uint8_t *>(h_hash_);
  SEQALK_CUDA(cudaMemcpyAsync(h, G.out.p, out_bytes, cudaMemcpyDeviceToHost, st));
  SEQALK_CUDA(cudaMemcpyAsync(h + rec_off, G.align.p, count * sizeof(PairAlign), cudaMemcpyDeviceToHost, st));
  SEQALK_CUDA(cudaStreamSynchronize(st));
  const PairAlign *rec = reinterpret_cast<const PairAlign *>(h + rec_off);
  const std::vector<PairDesc> &pairs = S.last_pairs;
  // one job per string, handed out longest first (the pairs of a batch are laid out longest first)
  std::vector<char> hex(count * 256);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    for (;;) {
      const size_t job = next.fetch_add(1);
      if (job >= 2 * count) break;
      const size_t p = job >> 1, which = job & 1;
      const PairDesc &P = pairs[p];
      char out[129];
      Sha512::hex(h + P.out_off + which * (size_t)P.cap + rec[p].start, rec[p].len, out);
      std::memcpy(&hex[p * 256 + which * 128], out, 128);
    }
  };
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nthreads = std::min<size_t>(std::min<size_t>(hw, 16), 2 * count);
  std::vector<std::thread> pool;
  for (size_t t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (std::thread &t : pool) t.join();
  for (size_t p = 0; p < count; ++p) Sha512::raw(&hex[p * 256], 256, h + dig_off + 64 * p);
  for (size_t p = 0; p < count; ++p)
    SEQALK_CUDA(cudaMemcpyAsync(digests_dev + 64ull * slots_host[p], h + dig_off + 64 * p, 64, cudaMemcpyDefault, st));
  // the landing buffer is reused by the next batch
  SEQALK_CUDA(cudaStreamSynchronize(st));
}

void DeviceWorker::start_async(int32_t *penalties_dev, uint8_t *digests_dev) {
  finish_async();
  {
    std::lock_guard<std::mutex> g(async_mu_);
    async_enqueued_ = 0;
    async_error_.clear();
    async_active_ = true;
  }
  async_thread_ = std::thread([this, penalties_dev, digests_dev]() {
    try {
      run_impl(penalties_dev, digests_dev, nullptr, nullptr, 0, batches_.size());
    } catch (const std::exception &e) {
      std::lock_guard<std::mutex> g(async_mu_);
      async_error_ = e.what();
      if (async_error_.empty()) async_error_ = "unknown error";
    }
    {
      std::lock_guard<std::mutex> g(async_mu_);
      async_enqueued_ = chunk_batch_.size();  // releases every waiter (also on error)
    }
    async_cv_.notify_all();
  });
}

void DeviceWorker::wait_chunk(size_t c, cudaStream_t consumer) {
  if (c + 1 >= chunk_batch_.size()) throw std::runtime_error("chunk index out of range");
  {
    std::unique_lock<std::mutex> lk(async_mu_);
    async_cv_.wait(lk, [&] { return async_enqueued_ > c; });
    if (!async_error_.empty()) throw std::runtime_error(async_error_);
  }
  if (chunk_batch_[c + 1] == chunk_batch_[c]) return;  // empty chunk
  SEQALK_CUDA(cudaSetDevice(dev_));
  if (consumer)
    SEQALK_CUDA(cudaStreamWaitEvent(consumer, chunk_ev_[c], 0));
  else
    SEQALK_CUDA(cudaEventSynchronize(chunk_ev_[c]));
}

void DeviceWorker::finish_async() {
  if (async_thread_.joinable()) async_thread_.join();
  std::string err;
  {
    std::lock_guard<std::mutex> g(async_mu_);
    async_active_ = false;
    err.swap(async_error_);
  }
  if (!err.empty()) throw std::runtime_error(err);
}

void DeviceWorker::fetch(int32_t *penalties_host, uint8_t *digests_host) {
  SEQALK_CUDA(cudaSetDevice(dev_));
  const size_t T = tasks_.size();
  if (!T) return;
  if (penalties_host) {
    if (results_in_caller_pen_ || !res_pen_.p || !ran_)
      throw std::runtime_error("fetch: the last run wrote its penalties to caller-supplied device memory (or did not run)");
    SEQALK_CUDA(cudaMemcpy(penalties_host, res_pen_.p, T * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  if (digests_host) {
    if (results_in_caller_dig_ || !have_digests_ || !res_dig_.p || !ran_)
      throw std::runtime_error("fetch: the last run wrote its digests to caller-supplied device memory (or produced none)");
    SEQALK_CUDA(cudaMemcpy(digests_host, res_dig_.p, T * 64, cudaMemcpyDeviceToHost));
  }
}

bool DeviceWorker::fetch_alignment(size_t index_in_job, std::vector<uint8_t> &a1, std::vector<uint8_t> &a2) {
  SEQALK_CUDA(cudaSetDevice(dev_));
  if (batches_.empty()) return false;
  const size_t b = batches_.size() - 1;
  const Batch &B = batches_[b];
  if (index_in_job < B.first || index_in_job >= B.first + B.count) return false;
  Slot &S = slots_[b % nslots_];
  Slot &G = (tiles_ && !overlap_) ? slots_[0] : S;
  size_t p = 0;
  while (p < S.last_pairs.size() && (size_t)S.last_pairs[p].task_id != index_in_job) ++p;
  if (p == S.last_pairs.size()) return false;
  const PairDesc &d = S.last_pairs[p];
  PairAlign al;
  SEQALK_CUDA(cudaMemcpy(&al, static_cast<PairAlign *>(G.align.p) + p, sizeof al, cudaMemcpyDeviceToHost));
  a1.resize(al.len);
  a2.resize(al.len);
  const uint8_t *out = static_cast<const uint8_t *>(G.out.p);
  if (al.len) {
    SEQALK_CUDA(cudaMemcpy(a1.data(), out + d.out_off + al.start, al.len, cudaMemcpyDeviceToHost));
    SEQALK_CUDA(cudaMemcpy(a2.data(), out + d.out_off + d.cap + al.start, al.len, cudaMemcpyDeviceToHost));
  }
  return true;
}

}  // namespace seqalk
