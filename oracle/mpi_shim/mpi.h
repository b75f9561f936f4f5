/*
 * Single-rank MPI shim — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * No MPI is installed in this image.  The reference sources
 * (/root/reference/submit/xuliny-seqalkway.cpp and
 *  /root/reference/testing12/seqalign-mpi-skeleton.cpp) only need a handful of
 * MPI entry points to run as "mpirun -np 1".  This header provides exactly
 * those with rank 0 / size 1 semantics so the UNMODIFIED reference sources can
 * be compiled where they lie (oracle/Makefile) into oracle/_ref/.
 *
 *   - Bcast on a 1-rank communicator is a no-op.
 *   - Send/Recv to self is a thread-safe, tag-matched FIFO: the reference's
 *     root rank talks to itself between two OpenMP threads
 *     (submit/xuliny-seqalkway.cpp:290-357).
 *   - Derived datatypes only need a byte extent.
 *
 * Nothing in the product (include/, the package, bench.py's GPU arm) includes
 * this file.
 */
#ifndef SEQALK_ORACLE_MPI_SHIM_H
#define SEQALK_ORACLE_MPI_SHIM_H

#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef std::ptrdiff_t MPI_Aint;
struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
};

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_STATUS_IGNORE ((MPI_Status *)0)

namespace mpi_shim {

struct Message {
  int tag;
  std::vector<unsigned char> bytes;
};

struct State {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Message> box;
  std::vector<std::size_t> extents;  // handle -> byte extent
  State() {
    extents.resize(3);
    extents[0] = 0;
    extents[MPI_CHAR] = sizeof(char);
    extents[MPI_INT] = sizeof(int);
  }
};

inline State &state() {
  static State s;
  return s;
}

inline std::size_t extent_of(MPI_Datatype t) {
  State &s = state();
  std::lock_guard<std::mutex> g(s.mu);
  return s.extents.at((std::size_t)t);
}

inline MPI_Datatype new_type(std::size_t extent) {
  State &s = state();
  std::lock_guard<std::mutex> g(s.mu);
  s.extents.push_back(extent);
  return (MPI_Datatype)(s.extents.size() - 1);
}

}  // namespace mpi_shim

inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
inline int MPI_Init_thread(int *, char ***, int required, int *provided) {
  if (provided) *provided = required;
  return MPI_SUCCESS;
}
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int *rank) {
  *rank = 0;
  return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm, int *size) {
  *size = 1;
  return MPI_SUCCESS;
}
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }

inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *out) {
  *out = mpi_shim::new_type((std::size_t)count * mpi_shim::extent_of(old));
  return MPI_SUCCESS;
}
inline int MPI_Type_create_struct(int count, const int *blocklens, const MPI_Aint *disps,
                                  const MPI_Datatype *types, MPI_Datatype *out) {
  std::size_t extent = 0;
  for (int b = 0; b < count; ++b) {
    std::size_t end = (std::size_t)disps[b] + (std::size_t)blocklens[b] * mpi_shim::extent_of(types[b]);
    if (end > extent) extent = end;
  }
  *out = mpi_shim::new_type(extent);
  return MPI_SUCCESS;
}
inline int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }

inline int MPI_Send(const void *buf, int count, MPI_Datatype type, int /*dest*/, int tag, MPI_Comm) {
  mpi_shim::State &s = mpi_shim::state();
  std::size_t nbytes = (std::size_t)count * mpi_shim::extent_of(type);
  mpi_shim::Message m;
  m.tag = tag;
  m.bytes.assign((const unsigned char *)buf, (const unsigned char *)buf + nbytes);
  {
    std::lock_guard<std::mutex> g(s.mu);
    s.box.push_back(std::move(m));
  }
  s.cv.notify_all();
  return MPI_SUCCESS;
}

inline int MPI_Recv(void *buf, int count, MPI_Datatype type, int /*source*/, int tag, MPI_Comm,
                    MPI_Status *status) {
  mpi_shim::State &s = mpi_shim::state();
  std::size_t cap = (std::size_t)count * mpi_shim::extent_of(type);
  std::unique_lock<std::mutex> lk(s.mu);
  for (;;) {
    for (std::deque<mpi_shim::Message>::iterator it = s.box.begin(); it != s.box.end(); ++it) {
      if (tag == MPI_ANY_TAG || it->tag == tag) {
        std::size_t nb = it->bytes.size() < cap ? it->bytes.size() : cap;
        std::memcpy(buf, it->bytes.data(), nb);
        if (status) {
          status->MPI_SOURCE = 0;
          status->MPI_TAG = it->tag;
          status->MPI_ERROR = MPI_SUCCESS;
        }
        s.box.erase(it);
        return MPI_SUCCESS;
      }
    }
    s.cv.wait(lk);
  }
}

#endif  // SEQALK_ORACLE_MPI_SHIM_H
