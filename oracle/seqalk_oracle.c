/*
 * seqalk_oracle.c — CPU restatement of the reference's all-pairs minimum-penalty
 * global alignment.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product (libseqalk.so, the package, the
 * seqalkway driver) never links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against the
 * reference's own golden outputs (mseq.dat, mseq1.dat, big13 subsets; see
 * tests/golden/README.md) and against the UNMODIFIED reference sources
 * compiled into oracle/_ref/ (oracle/Makefile) on fuzzed inputs.
 *
 * What it follows (paths relative to /root/reference):
 *   pair order / orientation / strip / hash chain :
 *       testing12/seqalign-mpi-skeleton.cpp:117-175
 *       submit/xuliny-seqalkway.cpp:183-227, 280-287, 334-337
 *   DP fill recurrence (match takes the diagonal unconditionally):
 *       testing12/seqalign-mpi-skeleton.cpp:206-223, submit/...:476-487
 *   traceback priority (match, mismatch-diag, up, left) and tails:
 *       testing12/seqalign-mpi-skeleton.cpp:233-273, submit/...:502-558
 *   SHA-512 -> 128 lowercase hex chars: sha512.hh:102-164,212-257 (= FIPS 180-4)
 *
 * Restatement choices (memory-lean, not a copy): two rolling score rows plus a
 * 2-bit direction matrix instead of the full int matrix, so 200 000^2 fits in
 * 10 GB; the alignment strings are assembled as prefix + path instead of
 * right-aligned int arrays followed by a strip scan (SURVEY.md section 8a
 * "equivalent closed form").
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ SHA-512 */

static const uint64_t K512[80] = {
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

static inline uint64_t rotr64(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }

static void sha512_block(uint64_t st[8], const unsigned char *p) {
  uint64_t w[80];
  for (int t = 0; t < 16; ++t) {
    uint64_t v = 0;
    for (int b = 0; b < 8; ++b) v = (v << 8) | p[8 * t + b];
    w[t] = v;
  }
  for (int t = 16; t < 80; ++t) {
    uint64_t s0 = rotr64(w[t - 15], 1) ^ rotr64(w[t - 15], 8) ^ (w[t - 15] >> 7);
    uint64_t s1 = rotr64(w[t - 2], 19) ^ rotr64(w[t - 2], 61) ^ (w[t - 2] >> 6);
    w[t] = w[t - 16] + s0 + w[t - 7] + s1;
  }
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int t = 0; t < 80; ++t) {
    uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
    uint64_t ch = (e & f) ^ (~e & g);
    uint64_t t1 = h + S1 + ch + K512[t] + w[t];
    uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
    uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint64_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

/* out: 128 lowercase hex chars + NUL (what sw::sha512::calculate returns). */
void oracle_sha512_hex(const void *data, size_t len, char out[129]) {
  uint64_t st[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                    0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                    0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
  const unsigned char *p = (const unsigned char *)data;
  size_t full = len / 128;
  for (size_t b = 0; b < full; ++b) sha512_block(st, p + 128 * b);
  unsigned char tail[256];
  size_t rem = len - 128 * full;
  memset(tail, 0, sizeof tail);
  if (rem) memcpy(tail, p + 128 * full, rem);
  tail[rem] = 0x80;
  size_t nb = (rem < 112) ? 1 : 2;
  uint64_t bits = (uint64_t)len << 3;
  for (int b = 0; b < 8; ++b) tail[128 * nb - 1 - b] = (unsigned char)(bits >> (8 * b));
  for (size_t b = 0; b < nb; ++b) sha512_block(st, tail + 128 * b);
  static const char hexd[] = "0123456789abcdef";
  for (int i = 0; i < 8; ++i)
    for (int b = 0; b < 8; ++b) {
      unsigned v = (unsigned)(st[i] >> (56 - 8 * b)) & 0xff;
      out[16 * i + 2 * b] = hexd[v >> 4];
      out[16 * i + 2 * b + 1] = hexd[v & 15];
    }
  out[128] = 0;
}

/* ------------------------------------------------------------- one pair */

enum { DIR_DIAG = 0, DIR_UP = 1, DIR_LEFT = 2 };

/*
 * Aligns x (rows, length m) against y (columns, length n).
 * Outputs (any may be NULL): *penalty = dp[m][n]; align1/align2 = the two
 * gapped strings with the common leading pad removed (capacity m+n+1 each,
 * NUL-terminated), *alen their length; hash_hex = problemhash
 * = H(H(align1) ++ H(align2)).
 * The alignment starts after the LAST column holding '_' in both strings
 * (skeleton:135-145): for inputs without a literal '_' that is the leading
 * pad; with '_' inside the inputs the reference cuts later, and so does this.
 * m = 0 or n = 0 follow the sequential skeleton (:186-280: penalty (m+n)*pgap,
 * pure tails; the submission's tiled twin divides by zero there).
 * Returns 0, or -1 on allocation failure / negative length.
 */
int oracle_pair(const char *x, int m, const char *y, int n, int pxy, int pgap, int *penalty,
                char *align1, char *align2, int *alen, char hash_hex[129]) {
  if (m < 0 || n < 0) return -1;
  size_t stride = ((size_t)n + 3) / 4; /* bytes per direction row (cells j=1..n) */
  unsigned char *dirs = (unsigned char *)calloc((size_t)m * stride + 1, 1);
  int *prev = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  int *cur = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  if (!dirs || !prev || !cur) {
    free(dirs); free(prev); free(cur);
    return -1;
  }
  for (int j = 0; j <= n; ++j) prev[j] = j * pgap; /* dp[0][j], skeleton:198-205 */
  for (int i = 1; i <= m; ++i) {
    unsigned char *drow = dirs + (size_t)(i - 1) * stride;
    const char xc = x[i - 1];
    cur[0] = i * pgap; /* dp[i][0] */
    for (int j = 1; j <= n; ++j) {
      int v, d;
      if (xc == y[j - 1]) { /* match: diagonal, no min (skeleton:210-213) */
        v = prev[j - 1];
        d = DIR_DIAG;
      } else {
        int dg = prev[j - 1] + pxy, up = prev[j] + pgap, lf = cur[j - 1] + pgap;
        /* min3 value; direction = first branch the traceback would take
           (skeleton:239-265): diag, then up, then left */
        v = dg < up ? dg : up;
        if (lf < v) v = lf;
        d = (dg == v) ? DIR_DIAG : (up == v) ? DIR_UP : DIR_LEFT;
      }
      cur[j] = v;
      drow[(j - 1) >> 2] |= (unsigned char)(d << (((j - 1) & 3) * 2));
    }
    int *t = prev; prev = cur; cur = t;
  }
  int pen = prev[n];
  if (penalty) *penalty = pen;

  if (align1 || align2 || alen || hash_hex) {
    size_t cap = (size_t)m + n + 1;
    char *a1 = align1 ? align1 : (char *)malloc(cap);
    char *a2 = align2 ? align2 : (char *)malloc(cap);
    char *r1 = (char *)malloc(cap), *r2 = (char *)malloc(cap); /* reversed path */
    if (!a1 || !a2 || !r1 || !r2) return -1;
    int i = m, j = n;
    size_t len = 0;
    while (i > 0 && j > 0) {
      int d = (dirs[(size_t)(i - 1) * stride + ((j - 1) >> 2)] >> (((j - 1) & 3) * 2)) & 3;
      if (d == DIR_DIAG) { r1[len] = x[i - 1]; r2[len] = y[j - 1]; --i; --j; }
      else if (d == DIR_UP) { r1[len] = x[i - 1]; r2[len] = '_'; --i; }
      else { r1[len] = '_'; r2[len] = y[j - 1]; --j; }
      ++len;
    }
    /* tails (skeleton:266-273): leftover prefix of x (or y) against '_';
       the all-'_' columns further left are what the strip scan removes. */
    size_t o = 0;
    for (int a = 0; a < i; ++a) { a1[o] = x[a]; a2[o] = '_'; ++o; }
    for (int a = 0; a < j; ++a) { a1[o] = '_'; a2[o] = y[a]; ++o; }
    for (size_t a = 0; a < len; ++a) { a1[o + a] = r1[len - 1 - a]; a2[o + a] = r2[len - 1 - a]; }
    o += len;
    a1[o] = 0; a2[o] = 0;
    /* strip scan (skeleton:135-145): columns up to the last '_'/'_' column are dropped */
    size_t first = 0;
    for (size_t a = o; a > 0; --a)
      if (a1[a - 1] == '_' && a2[a - 1] == '_') { first = a; break; }
    if (first) {
      memmove(a1, a1 + first, o - first + 1);
      memmove(a2, a2 + first, o - first + 1);
      o -= first;
    }
    if (alen) *alen = (int)o;
    if (hash_hex) {
      char cat[257];
      oracle_sha512_hex(a1, o, cat);
      oracle_sha512_hex(a2, o, cat + 128);
      oracle_sha512_hex(cat, 256, hash_hex);
    }
    free(r1); free(r2);
    if (!align1) free(a1);
    if (!align2) free(a2);
  }
  free(dirs); free(prev); free(cur);
  return 0;
}

/* Score only (linear memory): dp[m][n]. */
int oracle_pair_score(const char *x, int m, const char *y, int n, int pxy, int pgap) {
  int *prev = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  int *cur = (int *)malloc(sizeof(int) * ((size_t)n + 1));
  for (int j = 0; j <= n; ++j) prev[j] = j * pgap;
  for (int i = 1; i <= m; ++i) {
    cur[0] = i * pgap;
    const char xc = x[i - 1];
    for (int j = 1; j <= n; ++j) {
      if (xc == y[j - 1]) cur[j] = prev[j - 1];
      else {
        int dg = prev[j - 1] + pxy, up = prev[j] + pgap, lf = cur[j - 1] + pgap;
        int v = dg < up ? dg : up;
        cur[j] = lf < v ? lf : v;
      }
    }
    int *t = prev; prev = cur; cur = t;
  }
  int r = prev[n];
  free(prev); free(cur);
  return r;
}

/* ------------------------------------------------------------- all pairs */

/*
 * penalties[T], T=k(k-1)/2, in the reference's task order: for i=1..k-1, for
 * j=0..i-1 with x=seqs[i], y=seqs[j].  digest = H_T where H_0 = "" and
 * H_{t+1} = hex(SHA512(H_t ++ problemhash_t)).  pair_hashes (optional) gets T
 * records of 129 bytes.  nthreads<=0 -> all available; pairs are independent,
 * the chain is folded serially afterwards in task order.
 */
int oracle_all_pairs(const char *const *seqs, const int *lens, int k, int pxy, int pgap,
                     int *penalties, char digest[129], char *pair_hashes, int nthreads) {
  if (k < 2) return -1;
  long T = (long)k * (k - 1) / 2;
  char *hashes = pair_hashes ? pair_hashes : (char *)malloc((size_t)T * 129);
  if (!hashes) return -1;
  int *ii = (int *)malloc(sizeof(int) * T), *jj = (int *)malloc(sizeof(int) * T);
  long t = 0;
  for (int i = 1; i < k; ++i)
    for (int j = 0; j < i; ++j) { ii[t] = i; jj[t] = j; ++t; }
  int err = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (long id = 0; id < T; ++id) {
    int rc = oracle_pair(seqs[ii[id]], lens[ii[id]], seqs[jj[id]], lens[jj[id]], pxy, pgap,
                         &penalties[id], NULL, NULL, NULL, hashes + (size_t)id * 129);
    if (rc) {
#pragma omp atomic write
      err = 1;
    }
  }
  if (!err) {
    char buf[257], h[129];
    size_t hl = 0; /* H_0 = "" */
    for (long id = 0; id < T; ++id) {
      memcpy(buf + hl, hashes + (size_t)id * 129, 128);
      oracle_sha512_hex(buf, hl + 128, h);
      memcpy(buf, h, 128);
      hl = 128;
    }
    memcpy(digest, h, 129);
  }
  free(ii); free(jj);
  if (!pair_hashes) free(hashes);
  return err ? -1 : 0;
}

/* ------------------------------------------------------------- CLI (same stdin/stdout as the reference main) */
#ifdef ORACLE_MAIN
#include <sys/time.h>
static char *read_token(FILE *f, int *len) {
  int c;
  do { c = fgetc(f); } while (c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v');
  if (c == EOF) return NULL;
  size_t cap = 1024, n = 0;
  char *s = (char *)malloc(cap);
  while (c != EOF && !(c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v')) {
    if (n + 1 >= cap) { cap *= 2; s = (char *)realloc(s, cap); }
    s[n++] = (char)c;
    c = fgetc(f);
  }
  s[n] = 0;
  *len = (int)n;
  return s;
}
int main(void) {
  int pxy, pgap, k;
  if (scanf("%d %d %d", &pxy, &pgap, &k) != 3) return 2;
  char **seqs = (char **)malloc(sizeof(char *) * k);
  int *lens = (int *)malloc(sizeof(int) * k);
  for (int i = 0; i < k; ++i) seqs[i] = read_token(stdin, &lens[i]);
  long T = (long)k * (k - 1) / 2;
  int *pen = (int *)malloc(sizeof(int) * T);
  char digest[129];
  struct timeval a, b;
  gettimeofday(&a, NULL);
  if (oracle_all_pairs((const char *const *)seqs, lens, k, pxy, pgap, pen, digest, NULL, 0)) return 1;
  gettimeofday(&b, NULL);
  printf("Time: %ld us\n", (long)((b.tv_sec - a.tv_sec) * 1000000L + (b.tv_usec - a.tv_usec)));
  printf("%s\n", digest);
  for (long t = 0; t < T; ++t) printf("%d ", pen[t]);
  printf("\n");
  return 0;
}
#endif
