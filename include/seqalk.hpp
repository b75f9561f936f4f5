// seqalk.hpp — the reference's three C++ entry points as thin inline wrappers over the
// C ABI (seqalk.h).  Same names, argument meaning and error behaviour as the prototypes at
// submit/xuliny-seqalkway.cpp:13-15, so the frozen driver (`main`, :35-77) links unchanged:
//   std::string getMinimumPenalties(std::string *genes, int k, int pxy, int pgap, int *penalties);
//   int getMinimumPenalty(std::string x, std::string y, int pxy, int pgap, int *xans, int *yans);
//   void do_MPI_task(int rank);
// Errors: the reference prints to std::cerr and exit(1)s on allocation failure (:172-175);
// the wrappers do the same for any library failure.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "seqalk.h"

inline std::string getMinimumPenalties(std::string *genes, int k, int pxy, int pgap, int *penalties) {
  std::vector<const char *> ptrs(k);
  std::vector<int32_t> lens(k);
  for (int i = 0; i < k; ++i) {
    ptrs[i] = genes[i].data();
    lens[i] = (int32_t)genes[i].length();
  }
  char digest[129];
  static_assert(sizeof(int) == sizeof(int32_t), "int must be 32 bits");
  if (seqalk_all_pairs(ptrs.data(), lens.data(), k, pxy, pgap, reinterpret_cast<int32_t *>(penalties), digest)) {
    std::cerr << "getMinimumPenalties: " << seqalk_last_error() << std::endl;
    std::exit(1);
  }
  return std::string(digest);
}

inline int getMinimumPenalty(std::string x, std::string y, int pxy, int pgap, int *xans, int *yans) {
  int32_t penalty = 0;
  if (seqalk_pair(x.data(), (int32_t)x.length(), y.data(), (int32_t)y.length(), pxy, pgap,
                  reinterpret_cast<int32_t *>(xans), reinterpret_cast<int32_t *>(yans), &penalty)) {
    std::cerr << "getMinimumPenalty: " << seqalk_last_error() << std::endl;
    std::exit(1);
  }
  return penalty;
}

// Non-root MPI ranks have nothing to do: the GPUs of the box are driven from one process.
inline void do_MPI_task(int /*rank*/) {}
