// seqalkway_main.cpp — stdin/stdout driver, byte-compatible with the reference's frozen
// main (submit/xuliny-seqalkway.cpp:35-77) minus MPI:
//   stdin : pxy pgap k, then k whitespace-delimited sequences
//   stdout: "Time: %ld us\n", the 128-hex digest, then every penalty followed by one space
// The timed region covers exactly the getMinimumPenalties call, as in the reference
// (:54-62); GPU context creation happens before it (seqalk_init), like MPI_Init.
#include <sys/time.h>

#include <cstdint>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "seqalk.hpp"

static uint64_t GetTimeStamp() {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec * (uint64_t)1000000 + tv.tv_usec;
}

int main(int, char **) {
  std::ios::sync_with_stdio(true);
  int ngpus = 0;  // SEQALK_GPUS or all visible
  if (const char *e = std::getenv("SEQALK_GPUS")) ngpus = std::atoi(e);
  if (seqalk_init(ngpus)) {
    std::cerr << "seqalkway: " << seqalk_last_error() << std::endl;
    return 1;
  }
  int misMatchPenalty, gapPenalty, k;
  std::cin >> misMatchPenalty;
  std::cin >> gapPenalty;
  std::cin >> k;
  if (!std::cin || k < 0) {
    std::cerr << "seqalkway: malformed input" << std::endl;
    return 1;
  }
  std::vector<std::string> genes(k);
  for (int i = 0; i < k; i++) std::cin >> genes[i];
  const long numPairs = (long)k * (k - 1) / 2;
  std::vector<int> penalties(numPairs > 0 ? numPairs : 1);

  uint64_t start = GetTimeStamp();
  std::string alignmentHash = getMinimumPenalties(genes.data(), k, misMatchPenalty, gapPenalty, penalties.data());
  printf("Time: %ld us\n", (long)(GetTimeStamp() - start));

  std::cout << alignmentHash << std::endl;
  for (long i = 0; i < numPairs; i++) std::cout << penalties[i] << " ";
  std::cout << std::endl;
  seqalk_shutdown();
  return 0;
}
