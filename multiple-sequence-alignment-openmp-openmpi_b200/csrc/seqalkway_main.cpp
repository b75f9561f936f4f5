// seqalkway_main.cpp — stdin/stdout driver, byte-compatible with the reference's frozen
// main (submit/xuliny-seqalkway.cpp:35-77) minus MPI:
//   stdin : pxy pgap k, then k whitespace-delimited sequences
//   stdout: "Time: %ld us\n", the 128-hex digest, then every penalty followed by one space
// The timed region covers exactly the getMinimumPenalties call, as in the reference
// (:54-62); GPU context creation happens before it (seqalk_init), like MPI_Init.
#include <sys/time.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "seqalk.hpp"

static uint64_t GetTimeStamp() {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec * (uint64_t)1000000 + tv.tv_usec;
}

// Whole-stdin tokenizer: same token semantics as `std::cin >> std::string` (maximal runs of
// non-whitespace bytes), but one read + one pass instead of per-character stream extraction
// (the k=1000 input is 10 MB).
static bool read_all_stdin(std::string &buf) {
  char chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof chunk, stdin)) > 0) buf.append(chunk, n);
  return !ferror(stdin);
}
static inline bool is_space(unsigned char c) {
  return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f';
}
static bool next_token(const std::string &buf, size_t &pos, size_t &begin, size_t &len) {
  while (pos < buf.size() && is_space((unsigned char)buf[pos])) ++pos;
  if (pos >= buf.size()) return false;
  begin = pos;
  while (pos < buf.size() && !is_space((unsigned char)buf[pos])) ++pos;
  len = pos - begin;
  return true;
}

int main(int, char **) {
  int ngpus = 0;  // SEQALK_GPUS or all visible
  if (const char *e = std::getenv("SEQALK_GPUS")) ngpus = std::atoi(e);
  if (seqalk_init(ngpus)) {
    std::cerr << "seqalkway: " << seqalk_last_error() << std::endl;
    return 1;
  }
  std::string input;
  size_t pos = 0, b = 0, l = 0;
  long header[3] = {0, 0, -1};
  bool ok = read_all_stdin(input);
  for (int h = 0; ok && h < 3; ++h) {
    ok = next_token(input, pos, b, l);
    if (ok) {
      char *end = nullptr;
      const std::string tok = input.substr(b, l);
      header[h] = std::strtol(tok.c_str(), &end, 10);
      ok = end && *end == 0;
    }
  }
  const int misMatchPenalty = (int)header[0], gapPenalty = (int)header[1], k = (int)header[2];
  if (!ok || k < 0) {
    std::cerr << "seqalkway: malformed input" << std::endl;
    return 1;
  }
  std::vector<std::string> genes(k);
  for (int i = 0; i < k; i++)
    if (next_token(input, pos, b, l)) genes[i].assign(input, b, l);  // missing tokens stay empty, like cin
  const long numPairs = (long)k * (k - 1) / 2;
  std::vector<int> penalties(numPairs > 0 ? numPairs : 1);

  uint64_t start = GetTimeStamp();
  std::string alignmentHash = getMinimumPenalties(genes.data(), k, misMatchPenalty, gapPenalty, penalties.data());
  printf("Time: %ld us\n", (long)(GetTimeStamp() - start));

  // digest, then every penalty followed by one space, then a newline (reference main :65-70)
  std::string out = alignmentHash + "\n";
  char num[16];
  for (long i = 0; i < numPairs; i++) {
    const int w = snprintf(num, sizeof num, "%d ", penalties[i]);
    out.append(num, (size_t)w);
  }
  out += "\n";
  fwrite(out.data(), 1, out.size(), stdout);
  fflush(stdout);
  seqalk_shutdown();
  return 0;
}
