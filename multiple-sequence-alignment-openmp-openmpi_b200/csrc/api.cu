// api.cu — the C ABI of include/seqalk.h over the engine (host code only).
#include "../../include/seqalk.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <thread>
#include <vector>

#include "engine.h"
#include "sha512_host.h"

using namespace seqalk;

namespace {

thread_local std::string g_last_error;

// Idle workers of destroyed jobs, kept with their device buffers so that the next job on the
// same GPU does not pay cudaMalloc / cudaMallocHost again (a service calls the library many
// times; the reference's timed region likewise excludes MPI start-up).  seqalk_shutdown()
// releases them.
struct WorkerPool {
  std::mutex mu;
  std::vector<std::unique_ptr<DeviceWorker>> idle;
  std::unique_ptr<DeviceWorker> acquire(int device) {
    {
      std::lock_guard<std::mutex> g(mu);
      for (size_t i = 0; i < idle.size(); ++i)
        if (idle[i]->device() == device) {
          std::unique_ptr<DeviceWorker> w = std::move(idle[i]);
          idle.erase(idle.begin() + (long)i);
          if (w->nslots() == DeviceWorker::slots_from_env()) return w;
          w.reset();  // configuration changed: free its buffers before building a new one
          break;
        }
    }
    return std::unique_ptr<DeviceWorker>(new DeviceWorker(device));
  }
  void release(std::unique_ptr<DeviceWorker> w) {
    std::lock_guard<std::mutex> g(mu);
    idle.push_back(std::move(w));
  }
  void clear() {
    std::lock_guard<std::mutex> g(mu);
    idle.clear();
  }
};
static WorkerPool &pool() {
  static WorkerPool p;
  return p;
}

struct Runtime {
  std::mutex mu;
  std::vector<std::unique_ptr<DeviceWorker>> workers;  // devices 0..n-1
  bool peer_ok = false;                                 // every worker can store into device 0
  // fused gather: task-ordered result arrays on device 0, kept between calls (allocating and
  // freeing peer-mapped memory inside the call cost up to 0.6 s on some boxes)
  void *pen0 = nullptr, *dig0 = nullptr;
  uint8_t *dig_host = nullptr;  // page-locked landing buffer of the digests (folded from here)
  size_t res_cap = 0;  // tasks
  // phase times of the last seqalk_all_pairs call (seqalk_profile)
  double prof[SEQALK_PROFILE_LEN] = {0};
  void reserve_results(size_t tasks) {
    if (tasks <= res_cap) return;
    release_results();
    SEQALK_CUDA(cudaSetDevice(0));
    SEQALK_CUDA(cudaMalloc(&pen0, tasks * sizeof(int32_t)));
    SEQALK_CUDA(cudaMalloc(&dig0, tasks * 64));
    SEQALK_CUDA(cudaMallocHost(reinterpret_cast<void **>(&dig_host), tasks * 64));
    res_cap = tasks;
  }
  void release_results() {
    if (pen0 || dig0) {
      cudaSetDevice(0);
      cudaFree(pen0);
      cudaFree(dig0);
      cudaFreeHost(dig_host);
    }
    pen0 = dig0 = nullptr;
    dig_host = nullptr;
    res_cap = 0;
  }
};
Runtime &rt() {
  static Runtime r;
  return r;
}

int fail(const std::exception &e) {
  g_last_error = e.what();
  return 1;
}
int fail(const char *msg) {
  g_last_error = msg;
  return 1;
}

void init_locked(Runtime &r, int ngpus) {
  int visible = 0;
  cudaError_t e = cudaGetDeviceCount(&visible);
  if (e != cudaSuccess || visible < 1) {
    cudaGetLastError();
    throw std::runtime_error(
        "no CUDA device visible: libseqalk has no CPU fallback (needs an sm_100a GPU)");
  }
  if (ngpus <= 0 || ngpus > visible) ngpus = visible;
  r.release_results();
  for (auto &w : r.workers) pool().release(std::move(w));
  r.workers.clear();
  for (int d = 0; d < ngpus; ++d) r.workers.push_back(pool().acquire(d));
  // fused gather: let every device store its results straight into device 0's arrays
  r.peer_ok = ngpus > 1;
  for (int d = 1; d < ngpus && r.peer_ok; ++d) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, d, 0) != cudaSuccess || !can) {
      r.peer_ok = false;
      break;
    }
    SEQALK_CUDA(cudaSetDevice(d));
    cudaError_t pe = cudaDeviceEnablePeerAccess(0, 0);
    if (pe == cudaErrorPeerAccessAlreadyEnabled) {
      cudaGetLastError();
    } else if (pe != cudaSuccess) {
      cudaGetLastError();
      r.peer_ok = false;
    }
  }
  if (const char *env = std::getenv("SEQALK_P2P"))
    if (!std::atoi(env)) r.peer_ok = false;
}

void ensure_init(Runtime &r) {
  if (!r.workers.empty()) return;
  int n = 1;
  if (const char *e = std::getenv("SEQALK_GPUS")) n = std::atoi(e);
  init_locked(r, n);
}

struct DeviceArray {
  int dev;
  void *p = nullptr;
  DeviceArray(int d, size_t bytes) : dev(d) {
    SEQALK_CUDA(cudaSetDevice(dev));
    SEQALK_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
  }
  ~DeviceArray() {
    cudaSetDevice(dev);
    cudaFree(p);
  }
};

// Chunk c covers task ids [bounds[c], bounds[c+1]).  The host folds chunk c (about 0.6 us per task,
// serial by construction) while the GPUs work on the chunks after it; what stays exposed at the end
// is  max over c of  fold(chunks >= c) - gpu(chunks > c).
// A GPU works on up to three chunks at once (overlapped batch pipelines, engine.cu), so a chunk's
// results arrive two to three chunks after its fill started, and neighbouring chunks arrive together
// (8-GPU traces in profiles/r02_n8_chunk_trace.txt: with equal body chunks 42 % of the tasks arrived
// in the last 95 ms and left 23 - 30 ms of fold exposed).  Chunk sizes therefore shrink
// geometrically from the first one (x 0.8 per chunk): the work in flight — and with it the fold
// that is left when the GPUs finish — shrinks with them (the last three of 14 chunks hold 4.4 % of
// the tasks), while the first chunks stay large enough for full-size batches.
void chunk_bounds(int64_t T, int nchunks, std::vector<int64_t> &b) {
  b.assign((size_t)nchunks + 1, 0);
  if (nchunks >= 4 && T >= 4096) {
    double sum = 0, w = 1.0;
    for (int c = 0; c < nchunks; ++c, w *= 0.8) sum += w;
    double acc = 0;
    w = 1.0;
    for (int c = 0; c < nchunks; ++c, w *= 0.8) {
      acc += w;
      b[(size_t)c + 1] = std::max(b[(size_t)c], (int64_t)((double)T * (acc / sum)));
    }
  } else {
    for (int c = 0; c <= nchunks; ++c) b[c] = T * c / nchunks;
  }
  b[nchunks] = T;
}

// Shards of every chunk for every rank, one thread per chunk: per[c][r] = increasing task ids.
void partition_all_chunks(const int32_t *lens, int k, int nranks, const std::vector<int64_t> &bounds,
                          std::vector<std::vector<std::vector<int64_t>>> &per) {
  const int nchunks = (int)bounds.size() - 1;
  per.assign((size_t)nchunks, {});
  std::vector<std::thread> th;
  std::vector<std::string> errs((size_t)nchunks);
  for (int c = 0; c < nchunks; ++c)
    th.emplace_back([&, c]() {
      try {
        partition_lpt(lens, k, bounds[c], bounds[c + 1], nranks, per[c]);
      } catch (const std::exception &e) {
        errs[c] = e.what();
      }
    });
  for (std::thread &t : th) t.join();
  for (const std::string &e : errs)
    if (!e.empty()) throw std::runtime_error(e);
}

// All pairs on N >= 2 GPUs of this process with the fused gather: every GPU runs its
// cell-balanced share of each chunk of task ids; its digest kernel stores the results straight
// into GPU 0's task-ordered arrays (peer stores over NVLink); the host folds chunk c while the
// GPUs work on chunk c+1.
void all_pairs_peer(Runtime &r, const SeqPack &pack, int k, int pxy, int pgap, int32_t *penalties,
                    char digest_hex[129]) {
  const int N = (int)r.workers.size();
  const int64_t T = (int64_t)k * (k - 1) / 2;
  int nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(14, T / 16384));
  if (const char *e = std::getenv("SEQALK_PEER_CHUNKS"))  // tests: force several chunks on small jobs
    nchunks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, T), std::atoi(e)));
  // SEQALK_TRACE=1: phase times of this call on stderr
  const bool trace = std::getenv("SEQALK_TRACE") && std::atoi(std::getenv("SEQALK_TRACE"));
  const auto t0 = std::chrono::steady_clock::now();
  auto mark = [&](const char *what, int a = -1) {
    if (!trace) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (a >= 0)
      std::fprintf(stderr, "[seqalk] %9.2f ms  %s %d\n", ms, what, a);
    else
      std::fprintf(stderr, "[seqalk] %9.2f ms  %s\n", ms, what);
  };
  std::vector<int64_t> bounds;
  chunk_bounds(T, nchunks, bounds);
  std::vector<std::vector<std::vector<int64_t>>> per;
  partition_all_chunks(pack.len.data(), k, N, bounds, per);
  mark("partitioned");
  auto ms_since = [&](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
  };
  for (double &v : r.prof) v = 0;
  r.reserve_results((size_t)T);
  int32_t *const pen0 = static_cast<int32_t *>(r.pen0);
  uint8_t *const dig0 = static_cast<uint8_t *>(r.dig0);
  {  // upload + plan on every GPU at once; nothing is started until all of them succeeded
    std::vector<std::thread> th;
    std::vector<std::string> errs((size_t)N);
    for (int d = 0; d < N; ++d)
      th.emplace_back([&, d]() {
        try {
          std::vector<TaskIJ> tasks;
          std::vector<uint32_t> slots;
          std::vector<size_t> ends;
          for (int c = 0; c < nchunks; ++c) {
            for (int64_t id : per[c][d]) {
              TaskIJ t;
              task_from_id(id, &t.i, &t.j);
              tasks.push_back(t);
              slots.push_back((uint32_t)id);
            }
            ends.push_back(tasks.size());
          }
          DeviceWorker &w = *r.workers[d];
          w.load(pack);
          w.plan(tasks, pxy, pgap, &slots, &ends);
        } catch (const std::exception &e) {
          errs[d] = e.what();
          if (errs[d].empty()) errs[d] = "unknown error";
        }
      });
    for (std::thread &t : th) t.join();
    for (int d = 0; d < N; ++d)
      if (!errs[d].empty()) throw std::runtime_error("GPU " + std::to_string(d) + ": " + errs[d]);
  }
  mark("loaded and planned");
  r.prof[1] = ms_since(t0);
  const auto t_start = std::chrono::steady_clock::now();
  for (int d = 0; d < N; ++d) {
    try {
      r.workers[d]->start_async(pen0, dig0);
    } catch (...) {  // the GPUs already started write into pen0 / dig0: let them finish first
      for (int e = 0; e < d; ++e) {
        try {
          r.workers[e]->finish_async();
        } catch (...) {
        }
      }
      throw;
    }
  }
  uint8_t *const digests = r.dig_host;
  size_t hl = 0;
  digest_hex[0] = 0;
  // one folder thread takes the chunks in order as their results reach the host, so that this
  // thread never waits for a fold and sees every chunk the moment the GPUs finish it
  std::mutex fmu;
  std::condition_variable fcv;
  int ready = 0;      // chunks copied to the host
  bool stop = false;  // no more chunks will come
  std::thread folder([&]() {
    for (int c = 0;; ++c) {
      {
        std::unique_lock<std::mutex> g(fmu);
        fcv.wait(g, [&] { return ready > c || stop; });
        if (ready <= c) return;
      }
      const size_t lo = (size_t)bounds[c], n = (size_t)(bounds[c + 1] - bounds[c]);
      if (n) fold_chain(digests + 64 * lo, n, digest_hex, &hl);
      if (trace) mark("folded chunk", c);
    }
  });
  std::string error;
  try {
    for (int c = 0; c < nchunks; ++c) {
      for (int d = 0; d < N; ++d) r.workers[d]->wait_chunk((size_t)c, nullptr);
      mark("chunk done on every GPU", c);
      if (c == nchunks - 1) r.prof[2] = ms_since(t_start);  // every GPU has finished its last chunk
      const size_t lo = (size_t)bounds[c], n = (size_t)(bounds[c + 1] - bounds[c]);
      if (n) {
        const auto t_copy = std::chrono::steady_clock::now();
        SEQALK_CUDA(cudaSetDevice(0));
        SEQALK_CUDA(cudaMemcpy(penalties + lo, pen0 + lo, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        SEQALK_CUDA(cudaMemcpy(digests + 64 * lo, dig0 + 64 * lo, n * 64, cudaMemcpyDeviceToHost));
        r.prof[4] += ms_since(t_copy);
      }
      {
        std::lock_guard<std::mutex> g(fmu);
        ready = c + 1;
      }
      fcv.notify_one();
    }
  } catch (const std::exception &e) {
    error = e.what();
  }
  {
    std::lock_guard<std::mutex> g(fmu);
    stop = true;
  }
  fcv.notify_one();
  folder.join();
  mark("folded");
  r.prof[3] = ms_since(t_start) - r.prof[2];  // fold (and last copy) left exposed after the GPUs finished
  for (int d = 0; d < N; ++d) {
    try {
      r.workers[d]->finish_async();
    } catch (const std::exception &e) {
      if (error.empty()) error = "GPU " + std::to_string(d) + ": " + e.what();
    }
  }
  mark("GPUs idle");
  if (!error.empty()) throw std::runtime_error(error);
  double kmax = 0, ksum = 0;
  for (int d = 0; d < N; ++d) {
    const JobStats &st = r.workers[d]->stats();
    const double k = st.fill_ms + st.traceback_ms + st.digest_ms;
    kmax = std::max(kmax, k);
    ksum += k;
    r.prof[9] += st.fill_ms;
    r.prof[10] += st.traceback_ms;
    r.prof[11] += st.digest_ms;
  }
  r.prof[0] = ms_since(t0);
  r.prof[5] = kmax;
  r.prof[6] = ksum / N;
  r.prof[7] = N;
  r.prof[8] = nchunks;
}

// all tasks of `tasks` on the runtime's GPUs; results in task order on the host
void run_tasks(Runtime &r, const SeqPack &pack, const std::vector<TaskIJ> &tasks, int pxy, int pgap,
               int32_t *penalties, uint8_t *digests) {
  const int N = (int)r.workers.size();
  const size_t T = tasks.size();
  if (N == 1) {
    DeviceWorker &w = *r.workers[0];
    w.load(pack);
    w.plan(tasks, pxy, pgap);
    w.run();
    w.fetch(penalties, digests);
    return;
  }
  // shard by cells (LPT), one host thread per GPU
  std::vector<std::vector<TaskIJ>> shard(N);
  std::vector<std::vector<uint32_t>> slots(N);
  {
    struct Item {
      uint64_t cells;
      uint32_t idx;
    };
    std::vector<Item> items(T);
    for (size_t t = 0; t < T; ++t)
      items[t] = Item{(uint64_t)pack.len[tasks[t].i] * (uint64_t)pack.len[tasks[t].j], (uint32_t)t};
    std::sort(items.begin(), items.end(),
              [](const Item &a, const Item &b) { return a.cells != b.cells ? a.cells > b.cells : a.idx < b.idx; });
    std::vector<uint64_t> load(N, 0);
    std::vector<std::vector<uint32_t>> idx(N);
    for (const Item &it : items) {
      int best = 0;
      for (int d = 1; d < N; ++d)
        if (load[d] < load[best]) best = d;
      load[best] += it.cells;
      idx[best].push_back(it.idx);
    }
    for (int d = 0; d < N; ++d) {
      std::sort(idx[d].begin(), idx[d].end());
      for (uint32_t t : idx[d]) {
        shard[d].push_back(tasks[t]);
        slots[d].push_back(t);
      }
    }
  }
  std::unique_ptr<DeviceArray> pen0, dig0;
  if (r.peer_ok) {
    pen0.reset(new DeviceArray(0, T * sizeof(int32_t)));
    dig0.reset(new DeviceArray(0, T * 64));
  }
  std::vector<std::string> errors(N);
  std::vector<std::vector<int32_t>> lpen(N);
  std::vector<std::vector<uint8_t>> ldig(N);
  std::vector<std::thread> threads;
  for (int d = 0; d < N; ++d) {
    threads.emplace_back([&, d]() {
      try {
        DeviceWorker &w = *r.workers[d];
        w.load(pack);
        if (r.peer_ok) {
          // results land in device 0's task-ordered arrays through peer stores issued by
          // the digest kernel itself (gather fused into the producing kernel, over NVLink)
          w.plan(shard[d], pxy, pgap, &slots[d]);
          w.run(static_cast<int32_t *>(pen0->p), static_cast<uint8_t *>(dig0->p));
        } else {
          w.plan(shard[d], pxy, pgap);
          w.run();
          lpen[d].resize(shard[d].size());
          ldig[d].resize(shard[d].size() * 64);
          w.fetch(lpen[d].data(), ldig[d].data());
        }
      } catch (const std::exception &e) {
        errors[d] = e.what();
        if (errors[d].empty()) errors[d] = "unknown error";
      }
    });
  }
  for (std::thread &t : threads) t.join();
  for (int d = 0; d < N; ++d)
    if (!errors[d].empty()) throw std::runtime_error("GPU " + std::to_string(d) + ": " + errors[d]);
  if (r.peer_ok) {
    SEQALK_CUDA(cudaSetDevice(0));
    SEQALK_CUDA(cudaMemcpy(penalties, pen0->p, T * sizeof(int32_t), cudaMemcpyDeviceToHost));
    SEQALK_CUDA(cudaMemcpy(digests, dig0->p, T * 64, cudaMemcpyDeviceToHost));
  } else {
    for (int d = 0; d < N; ++d)
      for (size_t s = 0; s < slots[d].size(); ++s) {
        penalties[slots[d][s]] = lpen[d][s];
        std::memcpy(digests + 64ull * slots[d][s], ldig[d].data() + 64 * s, 64);
      }
  }
}

}  // namespace

struct seqalk_job {
  std::unique_ptr<DeviceWorker> worker;
  std::vector<TaskIJ> tasks;
  int pxy = 0, pgap = 0;
  ~seqalk_job() {
    if (worker) {
      try {
        worker->finish_async();
      } catch (...) {
      }
      pool().release(std::move(worker));
    }
  }
};

extern "C" {

int seqalk_init(int ngpus) {
  try {
    Runtime &r = rt();
    std::lock_guard<std::mutex> g(r.mu);
    init_locked(r, ngpus);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

void seqalk_shutdown(void) {
  Runtime &r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  r.release_results();
  r.workers.clear();
  pool().clear();
  pinned_clear();
}

int seqalk_device_count(void) {
  Runtime &r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  return (int)r.workers.size();
}

const char *seqalk_last_error(void) { return g_last_error.c_str(); }
const char *seqalk_version(void) { return "seqalk-b200 0.2 (sm_100a)"; }

int seqalk_all_pairs(const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy, int32_t pgap,
                     int32_t *penalties_out, char digest_hex_out[129]) {
  try {
    if (!digest_hex_out) return fail("null argument");
    if (k < 0) return fail("negative sequence count");
    if (k < 2) {  // no pairs: the reference returns the empty chain state and writes no penalty (:237, :282-287)
      digest_hex_out[0] = 0;
      return 0;
    }
    if (!seqs || !lens || !penalties_out) return fail("null argument");
    Runtime &r = rt();
    std::lock_guard<std::mutex> g(r.mu);
    ensure_init(r);
    SeqPack pack;
    pack.build(seqs, lens, k);
    std::vector<TaskIJ> tasks;
    all_tasks(k, tasks);
    if (r.workers.size() == 1) {
      // one GPU: results stream back batch by batch, the chain is folded under the GPU's shadow
      const auto t0 = std::chrono::steady_clock::now();
      auto ms_since = [&]() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      };
      for (double &v : r.prof) v = 0;
      DeviceWorker &w = *r.workers[0];
      w.load(pack);
      const double t_load = ms_since();
      w.plan(tasks, pxy, pgap);
      r.prof[1] = ms_since();
      if (const char *e = std::getenv("SEQALK_TRACE"))
        if (std::atoi(e) >= 2) std::fprintf(stderr, "[seqalk call] load %.3f ms, plan %.3f ms (after packing)\n", t_load, r.prof[1] - t_load);
      w.run_and_fold(penalties_out, digest_hex_out);
      const JobStats &st = w.stats();
      r.prof[0] = ms_since();
      r.prof[2] = r.prof[0] - r.prof[1];
      r.prof[5] = r.prof[6] = st.fill_ms + st.traceback_ms + st.digest_ms;
      r.prof[7] = 1;
      r.prof[8] = st.batches;
      r.prof[9] = st.fill_ms;
      r.prof[10] = st.traceback_ms;
      r.prof[11] = st.digest_ms;
      return 0;
    }
    if (r.peer_ok) {
      all_pairs_peer(r, pack, k, pxy, pgap, penalties_out, digest_hex_out);
      return 0;
    }
    std::vector<uint8_t> digests(tasks.size() * 64);
    run_tasks(r, pack, tasks, pxy, pgap, penalties_out, digests.data());
    size_t hl = 0;
    digest_hex_out[0] = 0;
    fold_chain(digests.data(), tasks.size(), digest_hex_out, &hl);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_pair(const char *x, int32_t m, const char *y, int32_t n, int32_t pxy, int32_t pgap, int32_t *xans,
                int32_t *yans, int32_t *penalty_out) {
  try {
    if (m < 0 || n < 0) return fail("negative sequence length");
    if ((m && !x) || (n && !y)) return fail("null argument");
    Runtime &r = rt();
    std::lock_guard<std::mutex> g(r.mu);
    ensure_init(r);
    const char *seqs[2] = {y, x};  // task (i=1, j=0): x = seqs[1] rows, y = seqs[0] columns
    const int32_t lens[2] = {n, m};
    SeqPack pack;
    pack.build(seqs, lens, 2);
    std::vector<TaskIJ> tasks(1, TaskIJ{1, 0});
    DeviceWorker &w = *r.workers[0];
    w.load(pack);
    w.plan(tasks, pxy, pgap);
    w.run();
    int32_t pen = 0;
    w.fetch(&pen, nullptr);
    if (penalty_out) *penalty_out = pen;
    if (xans || yans) {
      std::vector<uint8_t> a1, a2;
      if (!w.fetch_alignment(0, a1, a2)) return fail("alignment not available");
      const size_t l = (size_t)m + (size_t)n, L = a1.size();
      // right-aligned int arrays, '_' pad in front (submit/xuliny-seqalkway.cpp:535-558)
      for (size_t a = 1; a <= l; ++a) {
        const bool pad = a <= l - L;
        if (xans) xans[a] = pad ? (int32_t)'_' : (int32_t)(char)a1[a - 1 - (l - L)];
        if (yans) yans[a] = pad ? (int32_t)'_' : (int32_t)(char)a2[a - 1 - (l - L)];
      }
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_all_pairs_scores(const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy, int32_t pgap,
                            int32_t *penalties_out) {
  try {
    if (k < 0) return fail("negative sequence count");
    if (k < 2) return 0;
    if (!seqs || !lens || !penalties_out) return fail("null argument");
    Runtime &r = rt();
    std::lock_guard<std::mutex> g(r.mu);
    ensure_init(r);
    SeqPack pack;
    pack.build(seqs, lens, k);
    const int N = (int)r.workers.size();
    const int64_t T = (int64_t)k * (k - 1) / 2;
    if (N == 1) {
      std::vector<TaskIJ> tasks;
      all_tasks(k, tasks);
      DeviceWorker &w = *r.workers[0];
      w.load(pack);
      w.plan(tasks, pxy, pgap);
      w.run_scores();
      w.fetch(penalties_out, nullptr);
      return 0;
    }
    // cell-balanced shards, one host thread per GPU, results scattered on the host (4 B per task)
    std::vector<std::vector<int64_t>> per;
    partition_lpt(pack.len.data(), k, 0, T, N, per);
    std::vector<std::string> errors((size_t)N);
    std::vector<std::vector<int32_t>> lpen((size_t)N);
    std::vector<std::thread> threads;
    for (int d = 0; d < N; ++d)
      threads.emplace_back([&, d]() {
        try {
          std::vector<TaskIJ> tasks(per[d].size());
          for (size_t t = 0; t < tasks.size(); ++t) task_from_id(per[d][t], &tasks[t].i, &tasks[t].j);
          DeviceWorker &w = *r.workers[d];
          w.load(pack);
          w.plan(tasks, pxy, pgap);
          w.run_scores();
          lpen[d].resize(tasks.size());
          w.fetch(lpen[d].data(), nullptr);
        } catch (const std::exception &e) {
          errors[d] = e.what();
          if (errors[d].empty()) errors[d] = "unknown error";
        }
      });
    for (std::thread &t : threads) t.join();
    for (int d = 0; d < N; ++d)
      if (!errors[d].empty()) throw std::runtime_error("GPU " + std::to_string(d) + ": " + errors[d]);
    for (int d = 0; d < N; ++d)
      for (size_t t = 0; t < per[d].size(); ++t) penalties_out[per[d][t]] = lpen[d][t];
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_pair_score(const char *x, int32_t m, const char *y, int32_t n, int32_t pxy, int32_t pgap,
                      int32_t *penalty_out) {
  try {
    if (m < 0 || n < 0) return fail("negative sequence length");
    if ((m && !x) || (n && !y) || !penalty_out) return fail("null argument");
    Runtime &r = rt();
    std::lock_guard<std::mutex> g(r.mu);
    ensure_init(r);
    const char *seqs[2] = {y, x};
    const int32_t lens[2] = {n, m};
    SeqPack pack;
    pack.build(seqs, lens, 2);
    std::vector<TaskIJ> tasks(1, TaskIJ{1, 0});
    DeviceWorker &w = *r.workers[0];
    w.load(pack);
    w.plan(tasks, pxy, pgap);
    w.run_scores();
    w.fetch(penalty_out, nullptr);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_profile(double *out, int32_t n) {
  if (!out || n < 1) return fail("bad argument");
  Runtime &r = rt();
  std::lock_guard<std::mutex> g(r.mu);
  for (int i = 0; i < n; ++i) out[i] = i < SEQALK_PROFILE_LEN ? r.prof[i] : 0.0;
  return 0;
}

int64_t seqalk_num_tasks(int32_t k) { return k < 2 ? 0 : (int64_t)k * (k - 1) / 2; }

int seqalk_task_ij(int64_t task_id, int32_t *i_out, int32_t *j_out) {
  if (task_id < 0 || !i_out || !j_out) return fail("bad argument");
  task_from_id(task_id, i_out, j_out);
  return 0;
}

int seqalk_partition(const int32_t *lens, int32_t k, int64_t id_begin, int64_t id_end, int32_t nranks,
                     int32_t rank, int64_t *ids_out, int64_t *count_out) {
  try {
    if (!lens || !ids_out || !count_out || nranks < 1 || rank < 0 || rank >= nranks) return fail("bad argument");
    if (id_begin < 0 || id_end > seqalk_num_tasks(k) || id_begin > id_end) return fail("task id range out of bounds");
    std::vector<std::vector<int64_t>> per;
    partition_lpt(lens, k, id_begin, id_end, nranks, per);
    *count_out = (int64_t)per[rank].size();
    std::copy(per[rank].begin(), per[rank].end(), ids_out);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_partition_all(const int32_t *lens, int32_t k, int64_t id_begin, int64_t id_end, int32_t nranks,
                         int64_t *ids_out, int64_t *counts_out) {
  try {
    if (!lens || !ids_out || !counts_out || nranks < 1) return fail("bad argument");
    if (id_begin < 0 || id_end > seqalk_num_tasks(k) || id_begin > id_end) return fail("task id range out of bounds");
    std::vector<std::vector<int64_t>> per;
    partition_lpt(lens, k, id_begin, id_end, nranks, per);
    int64_t *w = ids_out;
    for (int r = 0; r < nranks; ++r) {
      counts_out[r] = (int64_t)per[r].size();
      w = std::copy(per[r].begin(), per[r].end(), w);
    }
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_chunk_bounds(int64_t ntasks, int32_t nchunks, int64_t *bounds_out) {
  if (ntasks < 0 || nchunks < 1 || !bounds_out) return fail("bad argument");
  std::vector<int64_t> b;
  chunk_bounds(ntasks, nchunks, b);
  std::copy(b.begin(), b.end(), bounds_out);
  return 0;
}

int seqalk_partition_chunks(const int32_t *lens, int32_t k, int32_t nranks, int32_t nchunks, int64_t *ids_out,
                            int64_t *counts_out) {
  try {
    if (!lens || !ids_out || !counts_out || nranks < 1 || nchunks < 1) return fail("bad argument");
    const int64_t T = seqalk_num_tasks(k);
    std::vector<int64_t> bounds;
    chunk_bounds(T, nchunks, bounds);
    std::vector<std::vector<std::vector<int64_t>>> per;
    partition_all_chunks(lens, k, nranks, bounds, per);
    int64_t *w = ids_out;
    for (int c = 0; c < nchunks; ++c)
      for (int r = 0; r < nranks; ++r) {
        counts_out[(size_t)c * nranks + r] = (int64_t)per[c][r].size();
        w = std::copy(per[c][r].begin(), per[c][r].end(), w);
      }
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_create(int32_t device, const char *const *seqs, const int32_t *lens, int32_t k, int32_t pxy,
                      int32_t pgap, const int64_t *task_ids, int64_t ntasks, seqalk_job_t **job_out) {
  try {
    if (!seqs || !lens || !job_out || k < 2) return fail("bad argument");
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1) {
      cudaGetLastError();
      return fail("no CUDA device visible: libseqalk has no CPU fallback (needs an sm_100a GPU)");
    }
    if (device < 0 || device >= visible) return fail("device index out of range");
    SeqPack pack;
    pack.build(seqs, lens, k);
    std::vector<TaskIJ> tasks;
    if (task_ids) {
      const int64_t T = seqalk_num_tasks(k);
      tasks.resize((size_t)ntasks);
      for (int64_t t = 0; t < ntasks; ++t) {
        if (task_ids[t] < 0 || task_ids[t] >= T) return fail("task id out of range");
        task_from_id(task_ids[t], &tasks[t].i, &tasks[t].j);
      }
    } else {
      all_tasks(k, tasks);
    }
    std::unique_ptr<seqalk_job> job(new seqalk_job);
    job->worker = pool().acquire(device);
    job->worker->load(pack);
    job->worker->plan(tasks, pxy, pgap);
    job->tasks = tasks;
    job->pxy = pxy;
    job->pgap = pgap;
    *job_out = job.release();
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_set_chunks(seqalk_job_t *job, const int64_t *chunk_ends, int32_t nchunks) {
  try {
    if (!job || !chunk_ends || nchunks < 1) return fail("bad argument");
    std::vector<size_t> ends(nchunks);
    for (int c = 0; c < nchunks; ++c) {
      if (chunk_ends[c] < 0 || (c && chunk_ends[c] < chunk_ends[c - 1])) return fail("chunk ends must increase");
      ends[c] = (size_t)chunk_ends[c];
    }
    if (ends.back() != job->tasks.size()) return fail("last chunk end must equal the number of tasks");
    job->worker->plan(job->tasks, job->pxy, job->pgap, nullptr, &ends);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_run_chunk(seqalk_job_t *job, int32_t chunk, void *penalties_dev, void *digests_dev) {
  try {
    if (!job || chunk < 0) return fail("bad argument");
    job->worker->run_chunk((size_t)chunk, static_cast<int32_t *>(penalties_dev),
                           static_cast<uint8_t *>(digests_dev));
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_start_async(seqalk_job_t *job, void *penalties_dev, void *digests_dev) {
  try {
    if (!job) return fail("null job");
    job->worker->start_async(static_cast<int32_t *>(penalties_dev), static_cast<uint8_t *>(digests_dev));
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_wait_chunk(seqalk_job_t *job, int32_t chunk, void *consumer_stream) {
  try {
    if (!job || chunk < 0) return fail("bad argument");
    job->worker->wait_chunk((size_t)chunk, static_cast<cudaStream_t>(consumer_stream));
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_finish_async(seqalk_job_t *job) {
  try {
    if (!job) return fail("null job");
    job->worker->finish_async();
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_run(seqalk_job_t *job, void *penalties_dev, void *digests_dev) {
  try {
    if (!job) return fail("null job");
    job->worker->run(static_cast<int32_t *>(penalties_dev), static_cast<uint8_t *>(digests_dev));
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_run_scores(seqalk_job_t *job, void *penalties_dev) {
  try {
    if (!job) return fail("null job");
    job->worker->run_scores(static_cast<int32_t *>(penalties_dev));
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_fetch(seqalk_job_t *job, int32_t *penalties_out, uint8_t *digests_out) {
  try {
    if (!job) return fail("null job");
    job->worker->fetch(penalties_out, digests_out);
    return 0;
  } catch (const std::exception &e) {
    return fail(e);
  }
}

int seqalk_job_stats(const seqalk_job_t *job, seqalk_stats_t *s) {
  if (!job || !s) return fail("null argument");
  const JobStats &j = job->worker->stats();
  s->cells = j.cells;
  s->dir_bytes = j.dir_bytes;
  s->batches = j.batches;
  s->fill_launches = j.fill_launches;
  s->traceback_launches = j.traceback_launches;
  s->digest_launches = j.digest_launches;
  s->fill_ms = j.fill_ms;
  s->traceback_ms = j.traceback_ms;
  s->digest_ms = j.digest_ms;
  s->wall_ms = j.wall_ms;
  s->tile_mode = j.tile_mode;
  s->fill_rows_per_lane = j.fill_rows_per_lane;
  s->self_checked = j.self_checked;
  s->fill_packed16 = j.fill_packed16;
  return 0;
}

void seqalk_job_destroy(seqalk_job_t *job) { delete job; }

int seqalk_fold_chain(const uint8_t *pair_digests, int64_t count, char state_hex[129]) {
  if (!state_hex || count < 0 || (count && !pair_digests)) return fail("bad argument");
  size_t hl = std::strlen(state_hex);
  if (hl != 0 && hl != 128) return fail("chain state must be \"\" or 128 hex chars");
  fold_chain(pair_digests, (size_t)count, state_hex, &hl);
  return 0;
}

int seqalk_sha512_hex(const void *data, size_t len, char out[129]) {
  if (!out || (len && !data)) return fail("bad argument");
  Sha512::hex(data, len, out);
  return 0;
}

}  // extern "C"
