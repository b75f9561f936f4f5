// digest.cu — per-pair problemhash on the GPU.
//
// Replaces, per pair, the three sw::sha512::calculate calls of do_task
//   submit/xuliny-seqalkway.cpp:217-219:
//     problemhash = H( hex(H(align1)) ++ hex(H(align2)) )
// where H is SHA-512 rendered as 128 lowercase hex characters (sha512.hh:125-164,
// transform :212-257 — FIPS 180-4).  Only the 64 raw bytes of problemhash leave the GPU;
// the host renders hex when it folds the chain (:334-337).
//
// Two adjacent lanes per pair: lane 2p hashes align1, lane 2p+1 hashes align2 (both have
// the same length), they exchange state by shuffle, the even lane hashes the 256-char
// concatenation.  Message bytes are fetched as aligned 64-bit words and funnel-shifted,
// because the strings start at an arbitrary byte offset of their right-aligned buffers.
#include "seqalk_kernels.h"

namespace seqalk {

__constant__ uint64_t K512[80] = {
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

__device__ __forceinline__ uint64_t rotr64(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }

__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
  const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}

struct Sha512State {
  uint64_t s[8];
  __device__ __forceinline__ void init() {
    s[0] = 0x6a09e667f3bcc908ULL; s[1] = 0xbb67ae8584caa73bULL; s[2] = 0x3c6ef372fe94f82bULL;
    s[3] = 0xa54ff53a5f1d36f1ULL; s[4] = 0x510e527fade682d1ULL; s[5] = 0x9b05688c2b3e6c1fULL;
    s[6] = 0x1f83d9abfb41bd6bULL; s[7] = 0x5be0cd19137e2179ULL;
  }
  // one compression over 16 big-endian message words
  __device__ void block(uint64_t (&w)[16]) {
    uint64_t a = s[0], b = s[1], c = s[2], d = s[3], e = s[4], f = s[5], g = s[6], h = s[7];
#pragma unroll 1
    for (int t0 = 0; t0 < 80; t0 += 16) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (t0) {
          const uint64_t w15 = w[(k + 1) & 15], w2 = w[(k + 14) & 15];
          const uint64_t s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
          const uint64_t s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
          w[k] = w[k] + s0 + w[(k + 9) & 15] + s1;
        }
        const uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
        const uint64_t ch = (e & f) ^ (~e & g);
        const uint64_t t1 = h + S1 + ch + K512[t0 + k] + w[k];
        const uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
        const uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
        const uint64_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
      }
    }
    s[0] += a; s[1] += b; s[2] += c; s[3] += d; s[4] += e; s[5] += f; s[6] += g; s[7] += h;
  }
};

// SHA-512 of `len` bytes starting at `p` (any alignment; up to 15 bytes past the end of
// the message may be read, which the buffer slack guarantees to be mapped).
__device__ void sha512_bytes(Sha512State &st, const uint8_t *p, uint32_t len) {
  st.init();
  const uint64_t *base = reinterpret_cast<const uint64_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
  const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(p) & 7) * 8;
  const uint32_t nblocks = (len + 16) / 128 + 1;
  const uint32_t last_word = 16 * nblocks - 1;
  for (uint32_t b = 0; b < nblocks; ++b) {
    uint64_t w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t wi = 16 * b + k;       // message word index
      const uint32_t off = 8 * wi;          // byte offset of this word
      uint64_t v = 0;
      if (off <= len) {
        // word holds message bytes and/or the 0x80 terminator
        if (off < len) {
          const uint64_t lo = base[wi];
          v = lo;
          if (sh) v = (lo >> sh) | (base[wi + 1] << (64 - sh));
          v = bswap64(v);
        }
        if (len - off < 8) {
          const unsigned rem = len - off;  // valid bytes in this word (0..7)
          const uint64_t keep = rem ? (~0ull << (8 * (8 - rem))) : 0ull;
          v = (v & keep) | (0x80ull << (8 * (7 - rem)));
        }
      }
      if (wi == last_word) v = (uint64_t)len << 3;
      w[k] = v;
    }
    st.block(w);
  }
}

__device__ __forceinline__ uint64_t hex8(uint32_t v) {
  // 4 bytes (big-endian order v[31:24] first) -> 8 ASCII hex chars packed big-endian
  uint64_t out = 0;
#pragma unroll
  for (int nib = 0; nib < 8; ++nib) {
    const uint32_t d = (v >> (28 - 4 * nib)) & 15u;
    const uint32_t ch = d < 10 ? ('0' + d) : ('a' + d - 10);
    out = (out << 8) | ch;
  }
  return out;
}

// The pair's penalty to its slot of the job-wide array, and the self-check: the fill's own dp[m][n]
// must equal the cost of the path the traceback walked.
__device__ __forceinline__ void publish_penalty(const DigestArgs &a, uint32_t p) {
  const int32_t pen = a.penalties[p];
  a.penalties_out[a.result_slot[p]] = pen;
  if (a.corner_mode != CORNER_NONE) {
    const PairDesc P = a.pairs[p];
    if (P.m > 0 && P.n > 0) {
      const int32_t score = a.brow[P.brow_off + corner_index(P, a.corner_mode)] + (P.m + P.n) * a.pgap;
      if (score != pen) {
        atomicAdd(a.check, 1u);
        atomicMin(a.check + 1, (uint32_t)P.task_id);
      }
    }
  }
}

// CAPPED (the default): <= 64 registers (220 bytes of spills) so that a CTA fits into the sliver of
// the register file the persistent fill kernel leaves free (dirs mode, see fill_blocks_per_sm) and
// four CTAs fit into what one exiting CTA of the packed fill frees (tile mode, overlapped
// pipelines).  Uncapped: the 91 registers it wants — faster alone, slower inside the overlapped job.
template <bool CAPPED>
__global__ void __launch_bounds__(DIGEST_THREADS, CAPPED ? 16 : 1) digest_kernel(const DigestArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = tid >> 1, which = tid & 1;
  const bool valid = p < a.npairs;
  if (a.host_hash) {  // the digests of this batch were computed by the host (engine.cu): penalties + self-check only
    if (valid && !which) publish_penalty(a, p);
    return;
  }
  Sha512State st;
  st.init();
  if (valid) {
    const PairDesc P = a.pairs[p];
    const PairAlign al = a.align[p];
    sha512_bytes(st, a.out + P.out_off + (uint64_t)which * P.cap + al.start, al.len);
  }
  // partner state (lane^1); all lanes of the warp take part in the shuffles
  uint64_t other[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) other[k] = __shfl_xor_sync(0xffffffffu, st.s[k], 1);
  if (!valid || which) return;
  // message = hex(H(align1)) ++ hex(H(align2)): 256 chars = 2 blocks, then the padding block
  Sha512State fin;
  fin.init();
  uint64_t w[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w[2 * k] = hex8((uint32_t)(st.s[k] >> 32));
    w[2 * k + 1] = hex8((uint32_t)st.s[k]);
  }
  fin.block(w);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w[2 * k] = hex8((uint32_t)(other[k] >> 32));
    w[2 * k + 1] = hex8((uint32_t)other[k]);
  }
  fin.block(w);
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = 0;
  w[0] = 0x8000000000000000ULL;
  w[15] = 256ull * 8;
  fin.block(w);
  // raw digest, big-endian bytes, at the pair's slot in task order
  uint64_t *dst = reinterpret_cast<uint64_t *>(a.digests + (uint64_t)a.result_slot[p] * 64);
#pragma unroll
  for (int k = 0; k < 8; ++k) dst[k] = bswap64(fin.s[k]);
  publish_penalty(a, p);
}

__global__ void __launch_bounds__(128) score_kernel(const ScoreArgs a) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npairs) return;
  const PairDesc P = a.pairs[p];
  int32_t pen;
  if (P.m == 0 || P.n == 0)
    pen = (P.m + P.n) * a.pgap;
  else if (a.corner_mode != CORNER_NONE)
    pen = a.brow[P.brow_off + corner_index(P, a.corner_mode)] + (P.m + P.n) * a.pgap;
  else
    pen = a.path_cost[p];
  a.penalties_out[a.result_slot[p]] = pen;
}

cudaError_t launch_scores(const ScoreArgs &args, cudaStream_t stream) {
  if (args.npairs == 0) return cudaSuccess;
  score_kernel<<<(args.npairs + 127) / 128, 128, 0, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launch_digest(const DigestArgs &args, cudaStream_t stream) {
  if (args.npairs == 0) return cudaSuccess;
  const unsigned threads = 2 * args.npairs;
  const unsigned grid = (threads + DIGEST_THREADS - 1) / DIGEST_THREADS;
  if (args.small_regs)
    digest_kernel<true><<<grid, DIGEST_THREADS, 0, stream>>>(args);
  else
    digest_kernel<false><<<grid, DIGEST_THREADS, 0, stream>>>(args);
  return cudaGetLastError();
}

}  // namespace seqalk
