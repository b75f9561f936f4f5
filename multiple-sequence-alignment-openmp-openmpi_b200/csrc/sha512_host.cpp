// sha512_host.cpp — see sha512_host.h
#include "sha512_host.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace seqalk {

namespace {
const uint64_t K[80] = {
    0x428a2f98d728ae22ULL, 0x7137449123ef65cdULL, 0xb5c0fbcfec4d3b2fULL, 0xe9b5dba58189dbbcULL,
    0x3956c25bf348b538ULL, 0x59f111f1b605d019ULL, 0x923f82a4af194f9bULL, 0xab1c5ed5da6d8118ULL,
    0xd807aa98a3030242ULL, 0x12835b0145706fbeULL, 0x243185be4ee4b28cULL, 0x550c7dc3d5ffb4e2ULL,
    0x72be5d74f27b896fULL, 0x80deb1fe3b1696b1ULL, 0x9bdc06a725c71235ULL, 0xc19bf174cf692694ULL,
    0xe49b69c19ef14ad2ULL, 0xefbe4786384f25e3ULL, 0x0fc19dc68b8cd5b5ULL, 0x240ca1cc77ac9c65ULL,
    0x2de92c6f592b0275ULL, 0x4a7484aa6ea6e483ULL, 0x5cb0a9dcbd41fbd4ULL, 0x76f988da831153b5ULL,
    0x983e5152ee66dfabULL, 0xa831c66d2db43210ULL, 0xb00327c898fb213fULL, 0xbf597fc7beef0ee4ULL,
    0xc6e00bf33da88fc2ULL, 0xd5a79147930aa725ULL, 0x06ca6351e003826fULL, 0x142929670a0e6e70ULL,
    0x27b70a8546d22ffcULL, 0x2e1b21385c26c926ULL, 0x4d2c6dfc5ac42aedULL, 0x53380d139d95b3dfULL,
    0x650a73548baf63deULL, 0x766a0abb3c77b2a8ULL, 0x81c2c92e47edaee6ULL, 0x92722c851482353bULL,
    0xa2bfe8a14cf10364ULL, 0xa81a664bbc423001ULL, 0xc24b8b70d0f89791ULL, 0xc76c51a30654be30ULL,
    0xd192e819d6ef5218ULL, 0xd69906245565a910ULL, 0xf40e35855771202aULL, 0x106aa07032bbd1b8ULL,
    0x19a4c116b8d2d0c8ULL, 0x1e376c085141ab53ULL, 0x2748774cdf8eeb99ULL, 0x34b0bcb5e19b48a8ULL,
    0x391c0cb3c5c95a63ULL, 0x4ed8aa4ae3418acbULL, 0x5b9cca4f7763e373ULL, 0x682e6ff3d6b2b8a3ULL,
    0x748f82ee5defb2fcULL, 0x78a5636f43172f60ULL, 0x84c87814a1f0ab72ULL, 0x8cc702081a6439ecULL,
    0x90befffa23631e28ULL, 0xa4506cebde82bde9ULL, 0xbef9a3f7b2c67915ULL, 0xc67178f2e372532bULL,
    0xca273eceea26619cULL, 0xd186b8c721c0c207ULL, 0xeada7dd6cde0eb1eULL, 0xf57d4f7fee6ed178ULL,
    0x06f067aa72176fbaULL, 0x0a637dc5a2c898a6ULL, 0x113f9804bef90daeULL, 0x1b710b35131c471bULL,
    0x28db77f523047d84ULL, 0x32caab7b40c72493ULL, 0x3c9ebe0a15c9bebcULL, 0x431d67c49c100d4cULL,
    0x4cc5d4becb3e42b6ULL, 0x597f299cfc657e2aULL, 0x5fcb6fab3ad6faecULL, 0x6c44198c4a475817ULL};

inline uint64_t rotr(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }
inline uint64_t load_be64(const uint8_t *p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return __builtin_bswap64(v);
}
inline void store_be64(uint8_t *p, uint64_t v) {
  v = __builtin_bswap64(v);
  std::memcpy(p, &v, 8);
}
}  // namespace

// The compression function in two halves, so that the fold below can take the message schedule
// of a block (which does not depend on the chain state) off the serial path:
//   schedule_kw: kw[t] = K[t] + W[t], t = 0..79, from the 128 message bytes;
//   rounds_kw:   the 80 rounds on precomputed kw.
// target_clones: a BMI2 (RORX) build of the same code, picked at load time on CPUs that have it.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define SEQALK_CLONES __attribute__((target_clones("default", "bmi2")))
#else
#define SEQALK_CLONES
#endif

SEQALK_CLONES void schedule_kw(const uint8_t *block, uint64_t kw[80]) {
  uint64_t w[16];
  for (int t = 0; t < 16; ++t) {
    w[t] = load_be64(block + 8 * t);
    kw[t] = w[t] + K[t];
  }
  for (int t0 = 16; t0 < 80; t0 += 16)
    for (int k = 0; k < 16; ++k) {
      const uint64_t w15 = w[(k + 1) & 15], w2 = w[(k + 14) & 15];
      w[k] += (rotr(w15, 1) ^ rotr(w15, 8) ^ (w15 >> 7)) + w[(k + 9) & 15] + (rotr(w2, 19) ^ rotr(w2, 61) ^ (w2 >> 6));
      kw[t0 + k] = w[k] + K[t0 + k];
    }
}

SEQALK_CLONES void rounds_kw(uint64_t st[8], const uint64_t kw[80]) {
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#define SEQALK_ROUND(a, b, c, d, e, f, g, h, kw)                                     \
  {                                                                                  \
    uint64_t t1 = h + (rotr(e, 14) ^ rotr(e, 18) ^ rotr(e, 41)) + ((e & f) ^ (~e & g)) + (kw); \
    uint64_t t2 = (rotr(a, 28) ^ rotr(a, 34) ^ rotr(a, 39)) + ((a & b) ^ (a & c) ^ (b & c)); \
    d += t1;                                                                         \
    h = t1 + t2;                                                                     \
  }
  for (int k = 0; k < 80; k += 8) {
    SEQALK_ROUND(a, b, c, d, e, f, g, h, kw[k + 0]);
    SEQALK_ROUND(h, a, b, c, d, e, f, g, kw[k + 1]);
    SEQALK_ROUND(g, h, a, b, c, d, e, f, kw[k + 2]);
    SEQALK_ROUND(f, g, h, a, b, c, d, e, kw[k + 3]);
    SEQALK_ROUND(e, f, g, h, a, b, c, d, kw[k + 4]);
    SEQALK_ROUND(d, e, f, g, h, a, b, c, kw[k + 5]);
    SEQALK_ROUND(c, d, e, f, g, h, a, b, kw[k + 6]);
    SEQALK_ROUND(b, c, d, e, f, g, h, a, kw[k + 7]);
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

SEQALK_CLONES void compress_block(uint64_t st[8], const uint8_t *block) {
  uint64_t w[16];
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int t = 0; t < 16; ++t) w[t] = load_be64(block + 8 * t);
  for (int t0 = 0; t0 < 80; t0 += 16) {
    if (t0) {
      for (int k = 0; k < 16; ++k) {
        const uint64_t w15 = w[(k + 1) & 15], w2 = w[(k + 14) & 15];
        w[k] += (rotr(w15, 1) ^ rotr(w15, 8) ^ (w15 >> 7)) + w[(k + 9) & 15] +
                (rotr(w2, 19) ^ rotr(w2, 61) ^ (w2 >> 6));
      }
    }
    for (int k = 0; k < 16; k += 8) {
      SEQALK_ROUND(a, b, c, d, e, f, g, h, K[t0 + k + 0] + w[k + 0]);
      SEQALK_ROUND(h, a, b, c, d, e, f, g, K[t0 + k + 1] + w[k + 1]);
      SEQALK_ROUND(g, h, a, b, c, d, e, f, K[t0 + k + 2] + w[k + 2]);
      SEQALK_ROUND(f, g, h, a, b, c, d, e, K[t0 + k + 3] + w[k + 3]);
      SEQALK_ROUND(e, f, g, h, a, b, c, d, K[t0 + k + 4] + w[k + 4]);
      SEQALK_ROUND(d, e, f, g, h, a, b, c, K[t0 + k + 5] + w[k + 5]);
      SEQALK_ROUND(c, d, e, f, g, h, a, b, K[t0 + k + 6] + w[k + 6]);
      SEQALK_ROUND(b, c, d, e, f, g, h, a, K[t0 + k + 7] + w[k + 7]);
    }
  }
#undef SEQALK_ROUND
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

void Sha512::compress(uint64_t st[8], const uint8_t *block) { compress_block(st, block); }

void Sha512::reset() {
  st_[0] = 0x6a09e667f3bcc908ULL; st_[1] = 0xbb67ae8584caa73bULL; st_[2] = 0x3c6ef372fe94f82bULL;
  st_[3] = 0xa54ff53a5f1d36f1ULL; st_[4] = 0x510e527fade682d1ULL; st_[5] = 0x9b05688c2b3e6c1fULL;
  st_[6] = 0x1f83d9abfb41bd6bULL; st_[7] = 0x5be0cd19137e2179ULL;
  fill_ = 0;
  total_ = 0;
}

void Sha512::update(const void *data, size_t len) {
  const uint8_t *p = static_cast<const uint8_t *>(data);
  total_ += len;
  if (fill_) {
    size_t take = 128 - fill_ < len ? 128 - fill_ : len;
    std::memcpy(buf_ + fill_, p, take);
    fill_ += take; p += take; len -= take;
    if (fill_ == 128) { compress(st_, buf_); fill_ = 0; }
  }
  while (len >= 128) { compress(st_, p); p += 128; len -= 128; }
  if (len) { std::memcpy(buf_, p, len); fill_ = len; }
}

void Sha512::final_raw(uint8_t out[64]) {
  uint8_t tail[256];
  std::memset(tail, 0, sizeof tail);
  std::memcpy(tail, buf_, fill_);
  tail[fill_] = 0x80;
  const size_t nb = fill_ < 112 ? 1 : 2;
  store_be64(tail + 128 * nb - 8, total_ << 3);
  for (size_t b = 0; b < nb; ++b) compress(st_, tail + 128 * b);
  for (int i = 0; i < 8; ++i) store_be64(out + 8 * i, st_[i]);
}

void Sha512::raw(const void *data, size_t len, uint8_t out[64]) {
  Sha512 s;
  s.update(data, len);
  s.final_raw(out);
}

void hex128(const uint8_t raw[64], char out[128]) {
  static const char d[] = "0123456789abcdef";
  for (int i = 0; i < 64; ++i) {
    out[2 * i] = d[raw[i] >> 4];
    out[2 * i + 1] = d[raw[i] & 15];
  }
}

void Sha512::hex(const void *data, size_t len, char out[129]) {
  uint8_t r[64];
  raw(data, len, r);
  hex128(r, out);
  out[128] = 0;
}

namespace {
const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                        0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
constexpr size_t FOLD_BLOCK = 512;  // tasks per hand-over between the schedule thread and the chain

// kw of the padding block of a 256-byte message (0x80, zeros, bit length 2048): constant
struct PadKw {
  uint64_t kw[80];
  PadKw() {
    uint8_t pad[128];
    std::memset(pad, 0, sizeof pad);
    pad[0] = 0x80;
    store_be64(pad + 120, 256ull * 8);
    schedule_kw(pad, kw);
  }
};

// the chain step for tasks [t0, t1) given the schedules of their second blocks (hex of the pair digest)
void fold_steps(const uint64_t *kw2, size_t n, uint8_t msg_state[128]) {
  static const PadKw pad;
  for (size_t t = 0; t < n; ++t) {
    uint64_t st[8];
    std::memcpy(st, IV, sizeof st);
    compress_block(st, msg_state);        // block 1 = hex(H_t): the only schedule on the serial path
    rounds_kw(st, kw2 + 80 * t);          // block 2 = hex(problemhash_t)
    rounds_kw(st, pad.kw);                // block 3 = padding
    uint8_t r[64];
    for (int i = 0; i < 8; ++i) store_be64(r + 8 * i, st[i]);
    hex128(r, reinterpret_cast<char *>(msg_state));
  }
}
}  // namespace

void fold_chain(const uint8_t *pair_digests, size_t count, char state_hex[129], size_t *hex_len) {
  // message = H_t (0 or 128 chars) ++ hex(pair digest) (128 chars).  For H_t != "" the message is
  // three blocks: hex(H_t) | hex(pair digest) | constant padding.  Only the first block's message
  // schedule depends on the chain; the second block's schedules are computed FOLD_BLOCK tasks
  // ahead by a helper thread (they cost about a third of a compression), the third's is constant.
  size_t hl = *hex_len;
  uint8_t msg[128];
  size_t t = 0;
  if (hl == 0 && count) {  // first step of the chain: a 128-byte message
    char hexd[128];
    hex128(pair_digests, hexd);
    uint8_t r[64];
    Sha512::raw(hexd, 128, r);
    hex128(r, reinterpret_cast<char *>(msg));
    hl = 128;
    t = 1;
  } else if (hl) {
    std::memcpy(msg, state_hex, 128);
  }
  const size_t rest = count - t;
  if (rest) {
    auto schedules = [&](size_t first, size_t n, uint64_t *kw2) {
      char hexd[128];
      for (size_t i = 0; i < n; ++i) {
        hex128(pair_digests + 64 * (first + i), hexd);
        schedule_kw(reinterpret_cast<const uint8_t *>(hexd), kw2 + 80 * i);
      }
    };
    if (rest < 4 * FOLD_BLOCK) {  // short: no thread
      std::vector<uint64_t> kw2(80 * std::min(rest, FOLD_BLOCK));
      for (size_t done = 0; done < rest; done += FOLD_BLOCK) {
        const size_t n = std::min(FOLD_BLOCK, rest - done);
        schedules(t + done, n, kw2.data());
        fold_steps(kw2.data(), n, msg);
      }
    } else {
      // ring of 4 blocks: the helper fills block b while the chain consumes block b - 1 ... b - 3
      constexpr size_t RING = 4;
      std::vector<uint64_t> kw2(80 * FOLD_BLOCK * RING);
      const size_t nblocks = (rest + FOLD_BLOCK - 1) / FOLD_BLOCK;
      std::atomic<size_t> produced{0}, consumed{0};
      std::thread helper([&]() {
        for (size_t b = 0; b < nblocks; ++b) {
          while (b >= consumed.load(std::memory_order_acquire) + RING) std::this_thread::yield();
          const size_t n = std::min(FOLD_BLOCK, rest - b * FOLD_BLOCK);
          schedules(t + b * FOLD_BLOCK, n, kw2.data() + 80 * FOLD_BLOCK * (b % RING));
          produced.store(b + 1, std::memory_order_release);
        }
      });
      for (size_t b = 0; b < nblocks; ++b) {
        while (produced.load(std::memory_order_acquire) <= b) std::this_thread::yield();
        const size_t n = std::min(FOLD_BLOCK, rest - b * FOLD_BLOCK);
        fold_steps(kw2.data() + 80 * FOLD_BLOCK * (b % RING), n, msg);
        consumed.store(b + 1, std::memory_order_release);
      }
      helper.join();
    }
  }
  if (hl) std::memcpy(state_hex, msg, 128);
  state_hex[hl] = 0;
  *hex_len = hl;
}

}  // namespace seqalk
