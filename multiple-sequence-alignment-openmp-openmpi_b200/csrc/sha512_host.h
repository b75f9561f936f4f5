// sha512_host.h — host SHA-512 for the ordered hash chain.
//
// Replaces sw::sha512::calculate (sha512.hh:159-164) on the one place where the
// reference hashes on the root rank: the serial fold
//   alignmentHash = H(alignmentHash ++ problemhash_id), id = 0..T-1
// (submit/xuliny-seqalkway.cpp:334-337).  The fold is inherently serial (each message
// starts with the previous digest), so it gets a tight implementation: fixed-shape
// 128/256-byte messages, no allocation, no stringstream, hex rendered by table.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace seqalk {

class Sha512 {
 public:
  Sha512() { reset(); }
  void reset();
  void update(const void *data, size_t len);
  void final_raw(uint8_t out[64]);
  // one-shot helpers
  static void raw(const void *data, size_t len, uint8_t out[64]);
  static void hex(const void *data, size_t len, char out[129]);  // NUL-terminated

  static void compress(uint64_t st[8], const uint8_t *block);

 private:
  uint64_t st_[8];
  uint8_t buf_[128];
  size_t fill_;
  uint64_t total_;
};

// 64 raw bytes -> 128 lowercase hex chars (no terminator written)
void hex128(const uint8_t raw[64], char out[128]);

// Folds T per-pair digests (64 raw bytes each, task order) into the final digest:
// H_0 = "", H_{t+1} = hex(SHA512(H_t ++ hex(pair_t))).  `state_hex` carries H across
// calls so that a caller can fold batches as they arrive: pass hex_len = 0 initially.
void fold_chain(const uint8_t *pair_digests, size_t count, char state_hex[129], size_t *hex_len);

}  // namespace seqalk
