/*
 * frieda_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see frieda_oracle.h).
 *
 * Scalar restatement of the reference's commit / proof / verify path.  The reference
 * glue is /root/reference/src/{commit,proof,utils}.rs; the arithmetic it calls lives in
 * stwo-prover @ 19d12d7 (not vendored), restated here from its published algorithm
 * (SURVEY.md Appendix A).  The structure deliberately mirrors stwo's CpuBackend --
 * per-call twiddle precompute by repeated point addition, a full-size radix-2 circle
 * FFT over the zero-extended coefficients, per-pair domain-point recomputation plus a
 * Fermat inversion inside the FRI folds, one scalar BLAKE2s compression per Merkle
 * node -- so that timing it is a fair "port" CPU baseline.
 */
#include "frieda_oracle.h"

#include <stdlib.h>
#include <string.h>

#define P 0x7fffffffu
typedef uint32_t m31;

/* ------------------------------------------------------------------ M31 */
static inline m31 m_add(m31 a, m31 b) {
  uint32_t s = a + b;
  return s >= P ? s - P : s;
}
static inline m31 m_sub(m31 a, m31 b) { return a >= b ? a - b : a + P - b; }
static inline m31 m_neg(m31 a) { return a ? P - a : 0; }
static inline m31 m_mul(m31 a, m31 b) {
  uint64_t p = (uint64_t)a * b;
  uint32_t r = (uint32_t)(p & P) + (uint32_t)(p >> 31);
  r = (r & P) + (r >> 31);
  return r == P ? 0 : r;
}
static m31 m_pow(m31 a, uint32_t e) {
  m31 r = 1;
  while (e) {
    if (e & 1) r = m_mul(r, a);
    a = m_mul(a, a);
    e >>= 1;
  }
  return r;
}
static inline m31 m_inv(m31 a) { return m_pow(a, P - 2); }
uint32_t fo_m31_mul(uint32_t a, uint32_t b) { return m_mul(a, b); }
uint32_t fo_m31_inv(uint32_t a) { return m_inv(a); }

/* ------------------------------------------------------------------ CM31 / QM31 (A.7) */
typedef struct { m31 a, b; } cm31;
static inline cm31 c_add(cm31 x, cm31 y) { return (cm31){m_add(x.a, y.a), m_add(x.b, y.b)}; }
static inline cm31 c_sub(cm31 x, cm31 y) { return (cm31){m_sub(x.a, y.a), m_sub(x.b, y.b)}; }
static inline cm31 c_mul(cm31 x, cm31 y) {
  return (cm31){m_sub(m_mul(x.a, y.a), m_mul(x.b, y.b)), m_add(m_mul(x.a, y.b), m_mul(x.b, y.a))};
}
typedef fo_qm31 qm31;
static inline qm31 q_zero(void) { return (qm31){{0, 0, 0, 0}}; }
static inline qm31 q_add(qm31 x, qm31 y) {
  return (qm31){{m_add(x.v[0], y.v[0]), m_add(x.v[1], y.v[1]), m_add(x.v[2], y.v[2]),
                 m_add(x.v[3], y.v[3])}};
}
static inline qm31 q_sub(qm31 x, qm31 y) {
  return (qm31){{m_sub(x.v[0], y.v[0]), m_sub(x.v[1], y.v[1]), m_sub(x.v[2], y.v[2]),
                 m_sub(x.v[3], y.v[3])}};
}
static inline qm31 q_mul_m(qm31 x, m31 s) {
  return (qm31){{m_mul(x.v[0], s), m_mul(x.v[1], s), m_mul(x.v[2], s), m_mul(x.v[3], s)}};
}
static inline qm31 q_mul(qm31 x, qm31 y) {
  /* (A + B u)(C + D u) = (AC + R BD) + (AD + BC) u,  R = 2 + i */
  cm31 A = {x.v[0], x.v[1]}, B = {x.v[2], x.v[3]}, C = {y.v[0], y.v[1]}, D = {y.v[2], y.v[3]};
  cm31 ac = c_mul(A, C), bd = c_mul(B, D);
  cm31 rbd = {m_sub(m_add(bd.a, bd.a), bd.b), m_add(bd.a, m_add(bd.b, bd.b))};
  cm31 lo = c_add(ac, rbd);
  cm31 hi = c_add(c_mul(A, D), c_mul(B, C));
  return (qm31){{lo.a, lo.b, hi.a, hi.b}};
}
static inline int q_eq(qm31 x, qm31 y) { return memcmp(&x, &y, sizeof x) == 0; }
static inline int q_is_zero(qm31 x) { return (x.v[0] | x.v[1] | x.v[2] | x.v[3]) == 0; }
fo_qm31 fo_qm31_mul(fo_qm31 a, fo_qm31 b) { return q_mul(a, b); }

/* ------------------------------------------------------------------ circle group (A.3) */
typedef struct { m31 x, y; } cpoint;
#define GEN_X 2u
#define GEN_Y 1268011823u
#define IDX_MASK 0x7fffffffu /* indices live mod 2^31 */

static inline cpoint p_add(cpoint a, cpoint b) {
  return (cpoint){m_sub(m_mul(a.x, b.x), m_mul(a.y, b.y)), m_add(m_mul(a.x, b.y), m_mul(a.y, b.x))};
}
static inline cpoint p_double(cpoint a) { return p_add(a, a); }
/* CirclePointIndex::to_point = generator.mul(index): LSB-first double-and-add. */
static cpoint p_from_index(uint32_t index) {
  cpoint res = {1, 0}, cur = {GEN_X, GEN_Y};
  index &= IDX_MASK;
  while (index) {
    if (index & 1) res = p_add(res, cur);
    cur = p_double(cur);
    index >>= 1;
  }
  return res;
}
void fo_circle_point(uint32_t index, uint32_t *x, uint32_t *y) {
  cpoint p = p_from_index(index);
  *x = p.x;
  *y = p.y;
}
static inline m31 double_x(m31 x) { return m_sub(m_add(m_mul(x, x), m_mul(x, x)), 1); }

typedef struct { uint32_t initial, step, log_size; } coset;
static inline coset half_odds(uint32_t k) {
  /* initial = subgroup_gen(k+2) = 2^(31-(k+2)); step = subgroup_gen(k) = 2^(31-k) */
  coset c;
  c.initial = (uint32_t)(((uint64_t)1 << (29 - k)) & IDX_MASK);
  c.step = (uint32_t)(((uint64_t)1 << (31 - k)) & IDX_MASK);
  c.log_size = k;
  return c;
}
static inline uint32_t coset_index_at(coset c, uint32_t i) {
  return (uint32_t)(((uint64_t)c.initial + (uint64_t)i * c.step) & IDX_MASK);
}
/* CircleDomain::index_at: second half is the conjugate (negated index) of the first. */
static inline uint32_t domain_index_at(coset half, uint32_t i) {
  uint32_t n = 1u << half.log_size;
  if (i < n) return coset_index_at(half, i);
  return (uint32_t)((0x80000000u - coset_index_at(half, i - n)) & IDX_MASK);
}
static inline uint32_t brev(uint32_t i, uint32_t bits) {
  uint32_t r = 0;
  for (uint32_t b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
  return r;
}
static void bit_reverse_u32(uint32_t *a, uint32_t log_n) {
  uint32_t n = 1u << log_n;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t j = brev(i, log_n);
    if (i < j) {
      uint32_t t = a[i];
      a[i] = a[j];
      a[j] = t;
    }
  }
}

/* ------------------------------------------------------------------ BLAKE2s (A.6, A.9) */
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4},
    {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13},
    {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11},
    {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5},
    {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
#define B2S_G(a, b, c, d, x, y) \
  do {                          \
    a = a + b + (x);            \
    d = rotr(d ^ a, 16);        \
    c = c + d;                  \
    b = rotr(b ^ c, 12);        \
    a = a + b + (y);            \
    d = rotr(d ^ a, 8);         \
    c = c + d;                  \
    b = rotr(b ^ c, 7);         \
  } while (0)

void fo_blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1,
                         uint32_t f0, uint32_t f1) {
  uint32_t v[16];
  for (int i = 0; i < 8; i++) {
    v[i] = h[i];
    v[i + 8] = B2S_IV[i];
  }
  v[12] ^= t0;
  v[13] ^= t1;
  v[14] ^= f0;
  v[15] ^= f1;
  for (int r = 0; r < 10; r++) {
    const uint8_t *s = B2S_SIGMA[r];
    B2S_G(v[0], v[4], v[8], v[12], m[s[0]], m[s[1]]);
    B2S_G(v[1], v[5], v[9], v[13], m[s[2]], m[s[3]]);
    B2S_G(v[2], v[6], v[10], v[14], m[s[4]], m[s[5]]);
    B2S_G(v[3], v[7], v[11], v[15], m[s[6]], m[s[7]]);
    B2S_G(v[0], v[5], v[10], v[15], m[s[8]], m[s[9]]);
    B2S_G(v[1], v[6], v[11], v[12], m[s[10]], m[s[11]]);
    B2S_G(v[2], v[7], v[8], v[13], m[s[12]], m[s[13]]);
    B2S_G(v[3], v[4], v[9], v[14], m[s[14]], m[s[15]]);
  }
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

static inline uint32_t ld32le(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline void st32le(uint8_t *p, uint32_t v) {
  p[0] = (uint8_t)v;
  p[1] = (uint8_t)(v >> 8);
  p[2] = (uint8_t)(v >> 16);
  p[3] = (uint8_t)(v >> 24);
}

/* RFC 7693 unkeyed BLAKE2s-256 (what the `blake2` crate's Blake2s256 computes). */
void fo_blake2s_256(const uint8_t *data, size_t len, uint8_t out[32]) {
  uint32_t h[8];
  memcpy(h, B2S_IV, sizeof h);
  h[0] ^= 0x01010020u;
  uint64_t t = 0;
  uint8_t block[64];
  uint32_t m[16];
  size_t off = 0;
  while (len - off > 64) {
    for (int i = 0; i < 16; i++) m[i] = ld32le(data + off + 4 * i);
    t += 64;
    fo_blake2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), 0, 0);
    off += 64;
  }
  size_t rem = len - off; /* 0..64; 0 only for the empty message */
  memset(block, 0, sizeof block);
  if (rem) memcpy(block, data + off, rem);
  for (int i = 0; i < 16; i++) m[i] = ld32le(block + 4 * i);
  t += rem;
  fo_blake2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), 0xFFFFFFFFu, 0);
  for (int i = 0; i < 8; i++) st32le(out + 4 * i, h[i]);
}

/* Blake2sMerkleHasher::hash_node: zero-state compression chain (SURVEY finding 3). */
static void hash_leaf4(const uint32_t vals[4], uint8_t out[32]) {
  uint32_t h[8] = {0}, m[16] = {0};
  m[0] = vals[0];
  m[1] = vals[1];
  m[2] = vals[2];
  m[3] = vals[3];
  fo_blake2s_compress(h, m, 0, 0, 0, 0);
  for (int i = 0; i < 8; i++) st32le(out + 4 * i, h[i]);
}
static void hash_children(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]) {
  uint32_t h[8] = {0}, m[16];
  for (int i = 0; i < 8; i++) {
    m[i] = ld32le(left + 4 * i);
    m[8 + i] = ld32le(right + 4 * i);
  }
  fo_blake2s_compress(h, m, 0, 0, 0, 0);
  for (int i = 0; i < 8; i++) st32le(out + 4 * i, h[i]);
}

/* ------------------------------------------------------------------ packing (A.1, A.2) */
size_t fo_bytes_to_felts(const uint8_t *data, size_t len, uint32_t *out, size_t cap) {
  /* src/utils.rs:10-19: BitVec<u8,Lsb0>.chunks(30).load::<u32>() */
  size_t n_bits = len * 8;
  size_t n = (n_bits + 29) / 30;
  for (size_t k = 0; k < n && k < cap; k++) {
    uint32_t v = 0;
    size_t base = k * 30;
    for (unsigned t = 0; t < 30; t++) {
      size_t j = base + t;
      if (j >= n_bits) break;
      v |= (uint32_t)((data[j >> 3] >> (j & 7)) & 1u) << t;
    }
    out[k] = v;
  }
  return n;
}
static uint32_t ceil_log2_sz(size_t n) {
  uint32_t l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
uint32_t fo_poly_log(size_t len) {
  /* src/utils.rs:23: 1 << max(ceil(log2(count)), 2); count = 0 -> log2 = -inf -> cast 0 */
  size_t n = (len * 8 + 29) / 30;
  uint32_t lg = n ? ceil_log2_sz(n) : 0;
  if (lg < 2) lg = 2;
  return lg - 2;
}

/* ------------------------------------------------------------------ twiddles (A.4) */
static void batch_inverse(const m31 *src, m31 *dst, size_t n) {
  if (!n) return;
  dst[0] = src[0];
  for (size_t i = 1; i < n; i++) dst[i] = m_mul(dst[i - 1], src[i]);
  m31 inv = m_inv(dst[n - 1]);
  for (size_t i = n - 1; i > 0; i--) {
    m31 t = m_mul(inv, dst[i - 1]);
    inv = m_mul(inv, src[i]);
    dst[i] = t;
  }
  dst[0] = inv;
}
int fo_precompute_twiddles(uint32_t k, uint32_t *tw, uint32_t *itw) {
  if (k > 29) return FO_ERR_PANIC;
  coset c = half_odds(k);
  cpoint init = p_from_index(c.initial), step = p_from_index(c.step);
  size_t pos = 0;
  for (uint32_t lvl = 0; lvl < k; lvl++) {
    uint32_t log_half = k - lvl - 1;
    size_t half = (size_t)1 << log_half;
    cpoint p = init;
    for (size_t i = 0; i < half; i++) { /* coset.iter().take(size/2).map(|p| p.x) */
      tw[pos + i] = p.x;
      p = p_add(p, step);
    }
    bit_reverse_u32(tw + pos, log_half);
    pos += half;
    init = p_double(init);
    step = p_double(step);
  }
  tw[pos++] = 1;
  if (itw) {
    const size_t chunk = 4096;
    size_t n = (size_t)1 << k;
    for (size_t o = 0; o < n; o += chunk) batch_inverse(tw + o, itw + o, n - o < chunk ? n - o : chunk);
  }
  return FO_OK;
}

/* ------------------------------------------------------------------ circle FFT (A.5) */
static inline void butterfly(m31 *v0, m31 *v1, m31 t) {
  m31 tmp = m_mul(*v1, t);
  *v1 = m_sub(*v0, tmp);
  *v0 = m_add(*v0, tmp);
}
int fo_circle_fft(uint32_t *v, uint32_t log_size, const uint32_t *tw) {
  if (log_size == 0 || log_size > 30) return FO_ERR_PANIC;
  coset half = half_odds(log_size - 1);
  if (log_size == 1) {
    cpoint p = p_from_index(half.initial);
    butterfly(&v[0], &v[1], p.y);
    return FO_OK;
  }
  if (log_size == 2) {
    cpoint p = p_from_index(half.initial);
    butterfly(&v[0], &v[2], p.x);
    butterfly(&v[1], &v[3], p.x);
    butterfly(&v[0], &v[1], p.y);
    butterfly(&v[2], &v[3], m_neg(p.y));
    return FO_OK;
  }
  uint32_t K = log_size - 1;
  size_t len = (size_t)1 << K;
  /* line layers, largest stride first: layer index i = j + 1 uses block j of the tree */
  for (int j = (int)K - 1; j >= 0; j--) {
    size_t s = (size_t)1 << (K - 1 - j);
    const uint32_t *lt = tw + (len - 2 * s);
    uint32_t i = (uint32_t)j + 1;
    for (size_t h = 0; h < s; h++) {
      m31 t = lt[h];
      for (size_t l = 0; l < ((size_t)1 << i); l++) {
        size_t a = (h << (i + 1)) + l;
        butterfly(&v[a], &v[a + ((size_t)1 << i)], t);
      }
    }
  }
  /* circle layer: twiddles [y, -y, -x, x] per consecutive (x, y) of line block 0 */
  {
    size_t s = (size_t)1 << (K - 1);
    const uint32_t *lt = tw + (len - 2 * s);
    for (size_t q = 0; q < s / 2; q++) {
      m31 x = lt[2 * q], y = lt[2 * q + 1];
      m31 ct[4] = {y, m_neg(y), m_neg(x), x};
      for (int e = 0; e < 4; e++) {
        size_t h = 4 * q + e;
        butterfly(&v[2 * h], &v[2 * h + 1], ct[e]);
      }
    }
  }
  return FO_OK;
}

int fo_circle_eval_naive(const uint32_t *coeffs, uint32_t n_coeffs_log, uint32_t log_size,
                         uint32_t *out) {
  if (log_size == 0 || log_size > 16 || n_coeffs_log > log_size) return FO_ERR_PANIC;
  coset half = half_odds(log_size - 1);
  uint32_t N = 1u << log_size, nc = 1u << n_coeffs_log;
  for (uint32_t i = 0; i < N; i++) {
    cpoint p = p_from_index(domain_index_at(half, i));
    m31 phi[32];
    phi[0] = p.y;
    phi[1] = p.x;
    for (uint32_t b = 2; b < log_size; b++) phi[b] = double_x(phi[b - 1]);
    m31 acc = 0;
    for (uint32_t k = 0; k < nc; k++) {
      if (!coeffs[k]) continue;
      m31 term = coeffs[k];
      for (uint32_t b = 0; b < log_size; b++)
        if ((k >> b) & 1u) term = m_mul(term, phi[b]);
      acc = m_add(acc, term);
    }
    out[brev(i, log_size)] = acc;
  }
  return FO_OK;
}

/* ------------------------------------------------------------------ Merkle (A.6) */
typedef struct {
  uint32_t log; /* leaf layer log size */
  uint8_t **levels; /* levels[k] : 2^k nodes * 32 B, k = 0..log */
} mtree;

static void mtree_free(mtree *t) {
  if (!t || !t->levels) return;
  for (uint32_t k = 0; k <= t->log; k++) free(t->levels[k]);
  free(t->levels);
  t->levels = NULL;
}
/* MerkleProver::commit over 4 equal-length columns (src/commit.rs:17-22). */
static int mtree_commit(mtree *t, const uint32_t *const cols[4], uint32_t log) {
  t->log = log;
  t->levels = (uint8_t **)calloc(log + 1, sizeof(uint8_t *));
  if (!t->levels) return FO_ERR_ALLOC;
  for (uint32_t k = 0; k <= log; k++) {
    t->levels[k] = (uint8_t *)malloc(((size_t)32) << k);
    if (!t->levels[k]) return FO_ERR_ALLOC;
  }
  size_t n = (size_t)1 << log;
  for (size_t i = 0; i < n; i++) {
    uint32_t vals[4] = {cols[0][i], cols[1][i], cols[2][i], cols[3][i]};
    hash_leaf4(vals, t->levels[log] + 32 * i);
  }
  for (int k = (int)log - 1; k >= 0; k--) {
    size_t m = (size_t)1 << k;
    const uint8_t *prev = t->levels[k + 1];
    for (size_t i = 0; i < m; i++) hash_children(prev + 64 * i, prev + 64 * i + 32, t->levels[k] + 32 * i);
  }
  return FO_OK;
}

/* ------------------------------------------------------------------ channel (A.9) */
typedef struct {
  uint8_t digest[32];
  uint64_t n_sent;
} channel;
static void ch_init(channel *c) { memset(c, 0, sizeof *c); }
static void ch_update(channel *c, const uint8_t d[32]) {
  memcpy(c->digest, d, 32);
  c->n_sent = 0;
}
static void ch_mix_root(channel *c, const uint8_t root[32]) {
  uint8_t buf[64], out[32];
  memcpy(buf, c->digest, 32);
  memcpy(buf + 32, root, 32);
  fo_blake2s_256(buf, 64, out);
  ch_update(c, out);
}
static int ch_mix_felts(channel *c, const qm31 *f, size_t n) {
  size_t len = 32 + 16 * n;
  uint8_t *buf = (uint8_t *)malloc(len);
  if (!buf) return FO_ERR_ALLOC;
  memcpy(buf, c->digest, 32);
  for (size_t i = 0; i < n; i++)
    for (int j = 0; j < 4; j++) st32le(buf + 32 + 16 * i + 4 * j, f[i].v[j]);
  uint8_t out[32];
  fo_blake2s_256(buf, len, out);
  free(buf);
  ch_update(c, out);
  return FO_OK;
}
static void ch_mix_u64(channel *c, uint64_t n) {
  uint32_t h[8], m[16] = {0};
  for (int i = 0; i < 8; i++) h[i] = ld32le(c->digest + 4 * i);
  m[0] = (uint32_t)n;
  m[1] = (uint32_t)(n >> 32);
  fo_blake2s_compress(h, m, 0, 0, 0, 0);
  uint8_t out[32];
  for (int i = 0; i < 8; i++) st32le(out + 4 * i, h[i]);
  ch_update(c, out);
}
static void ch_draw_random_bytes(channel *c, uint8_t out[32]) {
  uint8_t buf[64];
  memcpy(buf, c->digest, 32);
  memset(buf + 32, 0, 32);
  for (int i = 0; i < 8; i++) buf[32 + i] = (uint8_t)(c->n_sent >> (8 * i));
  c->n_sent++;
  fo_blake2s_256(buf, 64, out);
}
static qm31 ch_draw_felt(channel *c) {
  for (;;) {
    uint8_t b[32];
    uint32_t u[8];
    ch_draw_random_bytes(c, b);
    int ok = 1;
    for (int i = 0; i < 8; i++) {
      u[i] = ld32le(b + 4 * i);
      if (u[i] >= 2 * P) ok = 0;
    }
    if (!ok) continue;
    qm31 r;
    for (int i = 0; i < 4; i++) r.v[i] = u[i] % P;
    return r;
  }
}
static uint32_t ch_trailing_zeros(const channel *c) {
  uint32_t tz = 0;
  for (int i = 0; i < 16; i++) {
    uint8_t b = c->digest[i];
    if (b == 0) {
      tz += 8;
      continue;
    }
    while (!(b & 1)) {
      tz++;
      b >>= 1;
    }
    return tz;
  }
  return 128;
}

/* ------------------------------------------------------------------ FRI folds (A.8) */
/* ibutterfly(v0, v1, itw): (v0, v1) <- (v0 + v1, (v0 - v1) * itw) */
static inline qm31 fold_pair(qm31 a, qm31 b, m31 itw, qm31 alpha) {
  qm31 f0 = q_add(a, b);
  qm31 f1 = q_mul_m(q_sub(a, b), itw);
  return q_add(q_mul(alpha, f1), f0);
}
static inline qm31 col_at(uint32_t *const cols[4], size_t i) {
  return (qm31){{cols[0][i], cols[1][i], cols[2][i], cols[3][i]}};
}
static inline void col_set(uint32_t *const cols[4], size_t i, qm31 v) {
  for (int c = 0; c < 4; c++) cols[c][i] = v.v[c];
}
/* CpuBackend::fold_circle_into_line with dst = 0: recomputes each domain point by a scalar
 * multiplication and inverts y by Fermat, exactly as the reference's CPU path does. */
static void fold_circle_into_line(uint32_t *const src[4], uint32_t log, qm31 alpha, uint32_t *const dst[4]) {
  coset half = half_odds(log - 1);
  size_t n = (size_t)1 << log;
  qm31 alpha_sq = q_mul(alpha, alpha);
  for (size_t i = 0; i < n / 2; i++) {
    cpoint p = p_from_index(domain_index_at(half, brev((uint32_t)(i << 1), log)));
    qm31 f = fold_pair(col_at(src, 2 * i), col_at(src, 2 * i + 1), m_inv(p.y), alpha);
    col_set(dst, i, q_add(q_mul(q_zero(), alpha_sq), f));
  }
}
static void fold_line(uint32_t *const src[4], uint32_t log, qm31 alpha, uint32_t *const dst[4]) {
  coset c = half_odds(log);
  size_t n = (size_t)1 << log;
  for (size_t i = 0; i < n / 2; i++) {
    cpoint p = p_from_index(coset_index_at(c, brev((uint32_t)(i << 1), log)));
    col_set(dst, i, fold_pair(col_at(src, 2 * i), col_at(src, 2 * i + 1), m_inv(p.x), alpha));
  }
}
/* LineEvaluation::interpolate -> LinePoly (coefficients in storage = bit-reversed order). */
static void line_interpolate(qm31 *vals, uint32_t log) {
  size_t n = (size_t)1 << log;
  for (size_t i = 0; i < n; i++) {
    size_t j = brev((uint32_t)i, log);
    if (i < j) {
      qm31 t = vals[i];
      vals[i] = vals[j];
      vals[j] = t;
    }
  }
  coset dom = half_odds(log);
  uint32_t dlog = log;
  while (dlog > 0) {
    size_t size = (size_t)1 << dlog;
    for (size_t base = 0; base < n; base += size) {
      qm31 *l = vals + base, *r = vals + base + size / 2;
      for (size_t i = 0; i < size / 2; i++) {
        cpoint p = p_from_index(coset_index_at(dom, (uint32_t)i));
        m31 itw = m_inv(p.x);
        qm31 t = l[i];
        l[i] = q_add(t, r[i]);
        r[i] = q_mul_m(q_sub(t, r[i]), itw);
      }
    }
    dlog--;
    dom = half_odds(dlog); /* half_odds(k).double() == half_odds(k-1) */
  }
  m31 len_inv = m_inv((m31)(n % P));
  for (size_t i = 0; i < n; i++) vals[i] = q_mul_m(vals[i], len_inv);
}

/* ------------------------------------------------------------------ queries (A.11) */
static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : x > y;
}
/* Queries::generate: exactly n_queries draws, masked, into an ordered set. */
static int queries_generate(channel *c, uint32_t log_domain, uint64_t n_queries, uint32_t **out, uint32_t *n_out) {
  if (n_queries == 0 || n_queries > (1u << 24)) return FO_ERR_PANIC;
  uint32_t *q = (uint32_t *)malloc(sizeof(uint32_t) * n_queries);
  if (!q) return FO_ERR_ALLOC;
  uint32_t mask = log_domain >= 32 ? 0xffffffffu : ((1u << log_domain) - 1);
  uint64_t cnt = 0;
  while (cnt < n_queries) {
    uint8_t b[32];
    ch_draw_random_bytes(c, b);
    for (int i = 0; i < 8 && cnt < n_queries; i++) q[cnt++] = ld32le(b + 4 * i) & mask;
  }
  qsort(q, n_queries, sizeof(uint32_t), cmp_u32);
  uint32_t m = 0;
  for (uint64_t i = 0; i < n_queries; i++)
    if (m == 0 || q[m - 1] != q[i]) q[m++] = q[i];
  *out = q;
  *n_out = m;
  return FO_OK;
}
/* Queries::fold(n): shift and dedup consecutive. */
static uint32_t queries_fold(const uint32_t *q, uint32_t n, uint32_t shift, uint32_t *out) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t v = q[i] >> shift;
    if (m == 0 || out[m - 1] != v) out[m++] = v;
  }
  return m;
}

/* ------------------------------------------------------------------ decommit (A.12) */
typedef struct {
  uint32_t *pos;
  uint32_t n_pos;
} poslist;

/* compute_decommitment_positions_and_witness_evals with fold_step = 1 */
static int decommit_positions(uint32_t *const cols[4], const uint32_t *q, uint32_t nq, poslist *pl,
                              qm31 **wit, uint32_t *n_wit) {
  pl->pos = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (nq ? nq : 1));
  *wit = (qm31 *)malloc(sizeof(qm31) * 2 * (nq ? nq : 1));
  if (!pl->pos || !*wit) return FO_ERR_ALLOC;
  pl->n_pos = 0;
  *n_wit = 0;
  uint32_t i = 0;
  while (i < nq) {
    uint32_t g = q[i] >> 1, j = i;
    while (j < nq && (q[j] >> 1) == g) j++;
    uint32_t k = i;
    for (uint32_t position = 2 * g; position < 2 * g + 2; position++) {
      pl->pos[pl->n_pos++] = position;
      if (k < j && q[k] == position) {
        k++;
        continue;
      }
      (*wit)[(*n_wit)++] = col_at(cols, position);
    }
    i = j;
  }
  return FO_OK;
}
/* MerkleProver::decommit for one column size (all 4 columns at the leaf layer). */
static int merkle_decommit(const mtree *t, const uint32_t *pos, uint32_t n_pos, uint8_t **hw, uint32_t *n_hw) {
  size_t cap = (size_t)(n_pos ? n_pos : 1) * (t->log + 1);
  uint8_t *w = (uint8_t *)malloc(cap * 32);
  uint32_t *prev = (uint32_t *)malloc(sizeof(uint32_t) * (n_pos ? n_pos : 1));
  uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * (n_pos ? n_pos : 1));
  if (!w || !prev || !cur) return FO_ERR_ALLOC;
  uint32_t n_prev = 0, nw = 0;
  for (int k = (int)t->log; k >= 0; k--) {
    uint32_t n_cur = 0, pi = 0, ci = 0;
    uint32_t n_colq = (k == (int)t->log) ? n_pos : 0;
    for (;;) {
      /* next_decommitment_node: min(prev.peek()/2, layer_queries.peek()) */
      int have = 0;
      uint32_t node = 0;
      if (pi < n_prev) {
        node = prev[pi] / 2;
        have = 1;
      }
      if (ci < n_colq && (!have || pos[ci] < node)) {
        node = pos[ci];
        have = 1;
      }
      if (!have) break;
      if (k < (int)t->log) {
        const uint8_t *ph = t->levels[k + 1];
        if (pi < n_prev && prev[pi] == 2 * node) pi++;
        else memcpy(w + 32 * (size_t)nw++, ph + 32 * (size_t)(2 * node), 32);
        if (pi < n_prev && prev[pi] == 2 * node + 1) pi++;
        else memcpy(w + 32 * (size_t)nw++, ph + 32 * (size_t)(2 * node + 1), 32);
      }
      if (ci < n_colq && pos[ci] == node) ci++; /* queried values are not stored in the proof */
      cur[n_cur++] = node;
    }
    uint32_t *tmp = prev;
    prev = cur;
    cur = tmp;
    n_prev = n_cur;
  }
  free(prev);
  free(cur);
  *hw = w;
  *n_hw = nw;
  return FO_OK;
}

/* ------------------------------------------------------------------ traced prover */
typedef struct {
  uint32_t log;
  uint32_t *cols[4];
  mtree tree;
  qm31 alpha;
} fri_layer;

struct fo_trace {
  uint32_t poly_log, n_felts, D;
  fo_pcs_config cfg;
  uint32_t *coeffs;
  uint32_t *tw, *itw;
  uint32_t n_layers;
  fri_layer *layers;
  uint32_t last_log;
  uint32_t *last_cols[4];
  uint8_t digest_after_fri[32];
  uint64_t nonce;
  uint32_t n_queries;
  uint32_t *queries;
  fo_proof *proof;
  uint8_t root[32];
};

static void layer_proof_free(fo_layer_proof *l) {
  free(l->fri_witness);
  free(l->hash_witness);
  free(l->column_witness);
}
void fo_proof_free(fo_proof *p) {
  if (!p) return;
  layer_proof_free(&p->first_layer);
  for (uint32_t i = 0; i < p->n_inner_layers; i++) layer_proof_free(&p->inner_layers[i]);
  free(p->inner_layers);
  free(p->last_layer_poly);
  free(p->evaluations);
  free(p);
}
static void *dup_mem(const void *src, size_t n) {
  void *d = malloc(n ? n : 1);
  if (d && n) memcpy(d, src, n);
  return d;
}
static void layer_proof_clone(fo_layer_proof *d, const fo_layer_proof *s) {
  *d = *s;
  d->fri_witness = (fo_qm31 *)dup_mem(s->fri_witness, sizeof(fo_qm31) * s->n_fri_witness);
  d->hash_witness = (uint8_t *)dup_mem(s->hash_witness, 32 * (size_t)s->n_hash_witness);
  d->column_witness = (uint32_t *)dup_mem(s->column_witness, 4 * (size_t)s->n_column_witness);
}
fo_proof *fo_proof_clone(const fo_proof *p) {
  fo_proof *d = (fo_proof *)malloc(sizeof *d);
  *d = *p;
  layer_proof_clone(&d->first_layer, &p->first_layer);
  d->inner_layers = (fo_layer_proof *)malloc(sizeof(fo_layer_proof) * (p->n_inner_layers ? p->n_inner_layers : 1));
  for (uint32_t i = 0; i < p->n_inner_layers; i++) layer_proof_clone(&d->inner_layers[i], &p->inner_layers[i]);
  d->last_layer_poly = (fo_qm31 *)dup_mem(p->last_layer_poly, sizeof(fo_qm31) * p->n_last_layer_poly);
  d->evaluations = (fo_qm31 *)dup_mem(p->evaluations, sizeof(fo_qm31) * p->n_evaluations);
  return d;
}

void fo_trace_free(fo_trace *t) {
  if (!t) return;
  free(t->coeffs);
  free(t->tw);
  free(t->itw);
  if (t->layers) {
    for (uint32_t i = 0; i < t->n_layers; i++) {
      for (int c = 0; c < 4; c++) free(t->layers[i].cols[c]);
      mtree_free(&t->layers[i].tree);
    }
    free(t->layers);
  }
  for (int c = 0; c < 4; c++) free(t->last_cols[c]);
  free(t->queries);
  fo_proof_free(t->proof);
  free(t);
}

static int alloc_cols(uint32_t *cols[4], uint32_t log) {
  for (int c = 0; c < 4; c++) {
    cols[c] = (uint32_t *)calloc((size_t)1 << log, sizeof(uint32_t));
    if (!cols[c]) return FO_ERR_ALLOC;
  }
  return FO_OK;
}

/* polynomial_from_bytes + LDE (src/commit.rs:12-16 == src/proof.rs:38,44-50). */
static int build_lde(fo_trace *t, const uint8_t *data, size_t len, uint32_t log_blowup) {
  size_t n_felts = (len * 8 + 29) / 30;
  uint32_t plog = fo_poly_log(len);
  uint64_t Dll = (uint64_t)plog + log_blowup;
  if (Dll == 0 || Dll > 28) return FO_ERR_PANIC; /* half_odds(D-1) underflows at D == 0 */
  uint32_t D = (uint32_t)Dll;
  t->poly_log = plog;
  t->n_felts = (uint32_t)n_felts;
  t->D = D;
  size_t n_coef = (size_t)4 << plog;
  t->coeffs = (uint32_t *)calloc(n_coef, sizeof(uint32_t));
  if (!t->coeffs) return FO_ERR_ALLOC;
  fo_bytes_to_felts(data, len, t->coeffs, n_coef);
  /* twiddles: recomputed on every call, as the reference does (src/commit.rs:15) */
  size_t tn = (size_t)1 << (D - 1);
  t->tw = (uint32_t *)malloc(tn * sizeof(uint32_t));
  t->itw = (uint32_t *)malloc(tn * sizeof(uint32_t));
  if (!t->tw || !t->itw) return FO_ERR_ALLOC;
  int rc = fo_precompute_twiddles(D - 1, t->tw, t->itw);
  if (rc) return rc;
  t->layers = (fri_layer *)calloc(D + 1, sizeof(fri_layer));
  if (!t->layers) return FO_ERR_ALLOC;
  t->n_layers = 1;
  fri_layer *L0 = &t->layers[0];
  L0->log = D;
  rc = alloc_cols(L0->cols, D);
  if (rc) return rc;
  for (int c = 0; c < 4; c++) {
    memcpy(L0->cols[c], t->coeffs + ((size_t)c << plog), sizeof(uint32_t) << plog);
    rc = fo_circle_fft(L0->cols[c], D, t->tw);
    if (rc) return rc;
  }
  return FO_OK;
}

static int layer_commit(fri_layer *L) {
  const uint32_t *const cols[4] = {L->cols[0], L->cols[1], L->cols[2], L->cols[3]};
  return mtree_commit(&L->tree, cols, L->log);
}

int fo_commit(const uint8_t *data, size_t len, uint32_t log_blowup, uint8_t root_out[32]) {
  fo_trace *t = (fo_trace *)calloc(1, sizeof *t);
  if (!t) return FO_ERR_ALLOC;
  int rc = build_lde(t, data, len, log_blowup);
  if (!rc) rc = layer_commit(&t->layers[0]);
  if (!rc) memcpy(root_out, t->layers[0].tree.levels[0], 32);
  fo_trace_free(t);
  return rc;
}

static int make_layer_proof(const fri_layer *L, const uint32_t *q, uint32_t nq, fo_layer_proof *out) {
  poslist pl = {0};
  qm31 *wit = NULL;
  uint32_t n_wit = 0;
  int rc = decommit_positions(L->cols, q, nq, &pl, &wit, &n_wit);
  if (rc) return rc;
  uint8_t *hw = NULL;
  uint32_t n_hw = 0;
  rc = merkle_decommit(&L->tree, pl.pos, pl.n_pos, &hw, &n_hw);
  free(pl.pos);
  if (rc) return rc;
  memcpy(out->commitment, L->tree.levels[0], 32);
  out->fri_witness = wit;
  out->n_fri_witness = n_wit;
  out->hash_witness = hw;
  out->n_hash_witness = n_hw;
  out->column_witness = (uint32_t *)malloc(1);
  out->n_column_witness = 0;
  return FO_OK;
}

int fo_trace_run(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
                 int stop_after_fri, fo_trace **out) {
  fo_trace *t = (fo_trace *)calloc(1, sizeof *t);
  if (!t) return FO_ERR_ALLOC;
  t->cfg = *cfg;
  int rc = build_lde(t, data, len, cfg->log_blowup_factor);
  channel ch;
  ch_init(&ch);
  if (seed) ch_mix_u64(&ch, *seed); /* src/proof.rs:40-42 */
  if (rc) goto fail;
  {
    uint32_t D = t->D;
    uint64_t last_log = (uint64_t)cfg->log_last_layer_degree_bound + cfg->log_blowup_factor;
    /* commit_last_layer asserts evaluation.len() == last_layer_domain_size */
    if ((uint64_t)D - 1 < last_log) {
      rc = FO_ERR_PANIC;
      goto fail;
    }
    /* first layer */
    fri_layer *L = &t->layers[0];
    if ((rc = layer_commit(L))) goto fail;
    memcpy(t->root, L->tree.levels[0], 32);
    ch_mix_root(&ch, L->tree.levels[0]);
    L->alpha = ch_draw_felt(&ch);
    uint32_t *cur[4];
    uint32_t cur_log = D - 1;
    if ((rc = alloc_cols(cur, cur_log))) goto fail;
    fold_circle_into_line(L->cols, D, L->alpha, cur);
    while (cur_log > last_log) {
      fri_layer *Li = &t->layers[t->n_layers++];
      Li->log = cur_log;
      for (int c = 0; c < 4; c++) Li->cols[c] = cur[c];
      if ((rc = layer_commit(Li))) goto fail;
      ch_mix_root(&ch, Li->tree.levels[0]);
      Li->alpha = ch_draw_felt(&ch);
      if ((rc = alloc_cols(cur, cur_log - 1))) goto fail;
      fold_line(Li->cols, cur_log, Li->alpha, cur);
      cur_log--;
    }
    t->last_log = cur_log;
    for (int c = 0; c < 4; c++) t->last_cols[c] = cur[c];
    /* last layer: interpolate, degree check, mix_felts */
    size_t ln = (size_t)1 << cur_log;
    qm31 *vals = (qm31 *)malloc(sizeof(qm31) * ln);
    if (!vals) {
      rc = FO_ERR_ALLOC;
      goto fail;
    }
    for (size_t i = 0; i < ln; i++) vals[i] = col_at(cur, i);
    line_interpolate(vals, cur_log);
    /* into_ordered_coefficients = bit_reverse(storage) */
    qm31 *ordered = (qm31 *)malloc(sizeof(qm31) * ln);
    for (size_t i = 0; i < ln; i++) ordered[i] = vals[brev((uint32_t)i, cur_log)];
    size_t bound = (size_t)1 << cfg->log_last_layer_degree_bound;
    for (size_t i = bound; i < ln; i++)
      if (!q_is_zero(ordered[i])) {
        free(vals);
        free(ordered);
        rc = FO_ERR_PANIC; /* "invalid degree" */
        goto fail;
      }
    /* LinePoly::from_ordered_coefficients: bit-reverse the kept prefix back to storage order */
    fo_proof *pr = (fo_proof *)calloc(1, sizeof *pr);
    t->proof = pr;
    pr->n_last_layer_poly = (uint32_t)bound;
    pr->last_layer_poly = (qm31 *)malloc(sizeof(qm31) * bound);
    for (size_t i = 0; i < bound; i++)
      pr->last_layer_poly[i] = ordered[brev((uint32_t)i, cfg->log_last_layer_degree_bound)];
    free(vals);
    free(ordered);
    if ((rc = ch_mix_felts(&ch, pr->last_layer_poly, bound))) goto fail;
    memcpy(t->digest_after_fri, ch.digest, 32);
    pr->pcs_config = *cfg;
    pr->log_size_bound = t->poly_log;
    pr->n_inner_layers = t->n_layers - 1;
    pr->inner_layers = (fo_layer_proof *)calloc(t->n_layers, sizeof(fo_layer_proof));
    memcpy(pr->first_layer.commitment, t->layers[0].tree.levels[0], 32);
    for (uint32_t i = 1; i < t->n_layers; i++)
      memcpy(pr->inner_layers[i - 1].commitment, t->layers[i].tree.levels[0], 32);
    if (stop_after_fri) {
      *out = t;
      return FO_OK;
    }
    /* grind (src/proof.rs:58): smallest nonce whose mix gives >= pow_bits trailing zeros */
    if (cfg->pow_bits > 40) {
      rc = FO_ERR_PANIC;
      goto fail;
    }
    uint64_t nonce = 0;
    for (;; nonce++) {
      channel c2 = ch;
      ch_mix_u64(&c2, nonce);
      if (ch_trailing_zeros(&c2) >= cfg->pow_bits) break;
    }
    t->nonce = nonce;
    pr->proof_of_work = nonce;
    ch_mix_u64(&ch, nonce); /* src/proof.rs:59 */
    /* decommit (src/proof.rs:60) */
    if ((rc = queries_generate(&ch, D, cfg->n_queries, &t->queries, &t->n_queries))) goto fail;
    if ((rc = make_layer_proof(&t->layers[0], t->queries, t->n_queries, &pr->first_layer))) goto fail;
    uint32_t *fq = (uint32_t *)malloc(sizeof(uint32_t) * t->n_queries);
    for (uint32_t i = 1; i < t->n_layers; i++) {
      uint32_t nfq = queries_fold(t->queries, t->n_queries, i, fq);
      if ((rc = make_layer_proof(&t->layers[i], fq, nfq, &pr->inner_layers[i - 1]))) {
        free(fq);
        goto fail;
      }
    }
    free(fq);
    /* evaluations at the query positions, ascending (src/proof.rs:62-66) */
    pr->n_evaluations = t->n_queries;
    pr->evaluations = (qm31 *)malloc(sizeof(qm31) * (t->n_queries ? t->n_queries : 1));
    for (uint32_t i = 0; i < t->n_queries; i++) pr->evaluations[i] = col_at(t->layers[0].cols, t->queries[i]);
  }
  *out = t;
  return FO_OK;
fail:
  fo_trace_free(t);
  return rc;
}

int fo_prove(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
             uint8_t root_out[32], fo_proof **proof_out) {
  fo_trace *t = NULL;
  int rc = fo_trace_run(data, len, seed, cfg, 0, &t);
  if (rc) return rc;
  if (root_out) memcpy(root_out, t->root, 32);
  *proof_out = t->proof;
  t->proof = NULL;
  fo_trace_free(t);
  return FO_OK;
}

int fo_fri_commit(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
                  uint8_t *roots_out, uint32_t roots_cap, uint32_t *n_layers_out,
                  fo_qm31 *last_poly_out, uint32_t last_poly_cap) {
  fo_trace *t = NULL;
  int rc = fo_trace_run(data, len, seed, cfg, 1, &t);
  if (rc) return rc;
  if (n_layers_out) *n_layers_out = t->n_layers;
  for (uint32_t i = 0; i < t->n_layers && i < roots_cap; i++)
    memcpy(roots_out + 32 * i, t->layers[i].tree.levels[0], 32);
  for (uint32_t i = 0; i < t->proof->n_last_layer_poly && i < last_poly_cap; i++)
    last_poly_out[i] = t->proof->last_layer_poly[i];
  fo_trace_free(t);
  return FO_OK;
}

/* ------------------------------------------------------------------ verifier (A.13) */
typedef struct {
  uint32_t *positions; /* decommitment positions (pairs) */
  uint32_t n_positions;
  qm31 *subset_evals; /* 2 per subset */
  uint32_t n_subsets;
} sparse_eval;

/* compute_decommitment_positions_and_rebuild_evals, fold_step = 1.
 * returns 1 ok, 0 insufficient witness, FO_ERR_PANIC when query_evals runs out (unwrap). */
static int rebuild_evals(const uint32_t *q, uint32_t nq, const qm31 *query_evals, uint32_t n_query_evals,
                         const qm31 *wit, uint32_t n_wit, uint32_t *wit_used, sparse_eval *se) {
  se->positions = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (nq ? nq : 1));
  se->subset_evals = (qm31 *)malloc(sizeof(qm31) * 2 * (nq ? nq : 1));
  se->n_positions = 0;
  se->n_subsets = 0;
  uint32_t qe = 0, w = 0, i = 0;
  while (i < nq) {
    uint32_t g = q[i] >> 1, j = i;
    while (j < nq && (q[j] >> 1) == g) j++;
    uint32_t k = i;
    se->positions[se->n_positions++] = 2 * g;
    se->positions[se->n_positions++] = 2 * g + 1;
    for (uint32_t position = 2 * g; position < 2 * g + 2; position++) {
      qm31 v;
      if (k < j && q[k] == position) {
        k++;
        if (qe >= n_query_evals) return FO_ERR_PANIC;
        v = query_evals[qe++];
      } else {
        if (w >= n_wit) return 0;
        v = wit[w++];
      }
      se->subset_evals[2 * se->n_subsets + (position & 1)] = v;
    }
    se->n_subsets++;
    i = j;
  }
  *wit_used = w;
  return 1;
}
static void sparse_free(sparse_eval *se) {
  free(se->positions);
  free(se->subset_evals);
}
/* MerkleVerifier::verify with 4 columns at log size `log`.  1 ok / 0 error. */
static int merkle_verify(const uint8_t root[32], uint32_t log, const uint32_t *pos, uint32_t n_pos,
                         const uint32_t *queried_values, size_t n_queried_values, const uint8_t *hw,
                         uint32_t n_hw, uint32_t n_column_witness) {
  uint32_t cap = n_pos ? n_pos : 1;
  uint32_t *pidx = (uint32_t *)malloc(sizeof(uint32_t) * cap), *cidx = (uint32_t *)malloc(sizeof(uint32_t) * cap);
  uint8_t *ph = (uint8_t *)malloc(32 * (size_t)cap), *chh = (uint8_t *)malloc(32 * (size_t)cap);
  uint32_t n_prev = 0, hw_used = 0;
  size_t qv_used = 0;
  int ok = 1, have_prev = 0;
  for (int k = (int)log; k >= 0 && ok; k--) {
    uint32_t n_cur = 0, pi = 0, ci = 0, hi = 0;
    uint32_t n_colq = (k == (int)log) ? n_pos : 0;
    for (;;) {
      int have = 0;
      uint32_t node = 0;
      if (pi < n_prev) {
        node = pidx[pi] / 2;
        have = 1;
      }
      if (ci < n_colq && (!have || pos[ci] < node)) {
        node = pos[ci];
        have = 1;
      }
      if (!have) break;
      while (pi < n_prev && pidx[pi] / 2 == node) pi++;
      uint8_t out[32];
      uint32_t vals[4] = {0, 0, 0, 0};
      const uint8_t *lh = NULL, *rh = NULL;
      if (have_prev) {
        if (hi < n_prev && pidx[hi] == 2 * node) lh = ph + 32 * (size_t)hi++;
        else if (hw_used < n_hw) lh = hw + 32 * (size_t)hw_used++;
        else { ok = 0; break; }
        if (hi < n_prev && pidx[hi] == 2 * node + 1) rh = ph + 32 * (size_t)hi++;
        else if (hw_used < n_hw) rh = hw + 32 * (size_t)hw_used++;
        else { ok = 0; break; }
      }
      int is_colq = (ci < n_colq && pos[ci] == node);
      if (is_colq) ci++;
      uint32_t ncol = (k == (int)log) ? 4 : 0;
      if (ncol) {
        if (is_colq) {
          if (qv_used + 4 > n_queried_values) { ok = 0; break; }
          memcpy(vals, queried_values + qv_used, 16);
          qv_used += 4;
        } else { ok = 0; break; } /* column witness is always empty on this path */
      }
      if (lh) {
        /* hash_node(children, values): values are empty below the leaf layer */
        hash_children(lh, rh, out);
      } else {
        hash_leaf4(vals, out);
      }
      cidx[n_cur] = node;
      memcpy(chh + 32 * (size_t)n_cur, out, 32);
      n_cur++;
    }
    uint32_t *ti = pidx; pidx = cidx; cidx = ti;
    uint8_t *th = ph; ph = chh; chh = th;
    n_prev = n_cur;
    have_prev = 1;
  }
  if (ok && hw_used != n_hw) ok = 0;
  if (ok && qv_used != n_queried_values) ok = 0;
  if (ok && n_column_witness != 0) ok = 0;
  if (ok && (n_prev != 1 || memcmp(ph, root, 32) != 0)) ok = 0;
  free(pidx); free(cidx); free(ph); free(chh);
  return ok;
}

/* LinePoly::eval_at_point: recursive fold over doublings of x. */
static qm31 line_poly_eval(const qm31 *coeffs, uint32_t log, const qm31 *doublings) {
  if (log == 0) return coeffs[0];
  size_t half = (size_t)1 << (log - 1);
  qm31 l = line_poly_eval(coeffs, log - 1, doublings + 1);
  qm31 r = line_poly_eval(coeffs + half, log - 1, doublings + 1);
  return q_add(l, q_mul(r, doublings[0]));
}

int fo_verify(const fo_proof *pr, const uint64_t *seed) {
  const fo_pcs_config *cfg = &pr->pcs_config;
  channel ch;
  ch_init(&ch);
  if (seed) ch_mix_u64(&ch, *seed);
  /* FriVerifier::commit */
  ch_mix_root(&ch, pr->first_layer.commitment);
  if (pr->log_size_bound == 0) return FO_ERR_PANIC; /* fold_to_line underflow */
  uint64_t Dll = (uint64_t)pr->log_size_bound + cfg->log_blowup_factor;
  if (Dll > 28) return FO_ERR_PANIC;
  uint32_t D = (uint32_t)Dll;
  qm31 alpha0 = ch_draw_felt(&ch);
  uint32_t n_inner = pr->n_inner_layers;
  qm31 *alphas = (qm31 *)malloc(sizeof(qm31) * (n_inner ? n_inner : 1));
  uint32_t bound = pr->log_size_bound - 1;
  for (uint32_t i = 0; i < n_inner; i++) {
    ch_mix_root(&ch, pr->inner_layers[i].commitment);
    alphas[i] = ch_draw_felt(&ch);
    if (bound == 0) { free(alphas); return 0; } /* InvalidNumFriLayers */
    bound--;
  }
  if (bound != cfg->log_last_layer_degree_bound) { free(alphas); return 0; }
  if (pr->n_last_layer_poly > (1u << cfg->log_last_layer_degree_bound)) { free(alphas); return 0; }
  /* LinePoly must have power-of-two length; stwo's LinePoly::new asserts it at construction. */
  if (pr->n_last_layer_poly == 0 || (pr->n_last_layer_poly & (pr->n_last_layer_poly - 1))) { free(alphas); return FO_ERR_PANIC; }
  ch_mix_felts(&ch, pr->last_layer_poly, pr->n_last_layer_poly);
  /* src/proof.rs:91-95 */
  ch_mix_u64(&ch, pr->proof_of_work);
  if (ch_trailing_zeros(&ch) < cfg->pow_bits) { free(alphas); return 0; }
  uint32_t *q = NULL, nq = 0;
  int rc = queries_generate(&ch, D, cfg->n_queries, &q, &nq);
  if (rc) { free(alphas); return rc; }
  int result = 0;
  uint32_t *lq = (uint32_t *)malloc(sizeof(uint32_t) * nq);
  qm31 *evals = (qm31 *)malloc(sizeof(qm31) * nq);
  /* first layer */
  {
    sparse_eval se;
    uint32_t used = 0;
    int r = rebuild_evals(q, nq, pr->evaluations, pr->n_evaluations, pr->first_layer.fri_witness,
                          pr->first_layer.n_fri_witness, &used, &se);
    if (r != 1) { sparse_free(&se); result = r; goto done; }
    if (used != pr->first_layer.n_fri_witness) { sparse_free(&se); goto done; }
    int ok = merkle_verify(pr->first_layer.commitment, D, se.positions, se.n_positions,
                           (const uint32_t *)se.subset_evals, 4 * (size_t)se.n_positions,
                           pr->first_layer.hash_witness, pr->first_layer.n_hash_witness,
                           pr->first_layer.n_column_witness);
    if (!ok) { sparse_free(&se); goto done; }
    if (n_inner == 0) { sparse_free(&se); result = FO_ERR_PANIC; goto done; } /* assert!(first_layer_columns.is_empty()) */
    /* fold_circle of each subset with alpha0; accumulate into zeros */
    coset half = half_odds(D - 1);
    qm31 a_sq = q_mul(alpha0, alpha0);
    for (uint32_t s = 0; s < se.n_subsets; s++) {
      uint32_t start = se.positions[2 * s];
      cpoint p = p_from_index(domain_index_at(half, brev(start, D)));
      qm31 f = fold_pair(se.subset_evals[2 * s], se.subset_evals[2 * s + 1], m_inv(p.y), alpha0);
      evals[s] = q_add(q_mul(q_zero(), a_sq), f);
    }
    sparse_free(&se);
  }
  uint32_t nlq = queries_fold(q, nq, 1, lq);
  uint32_t llog = D - 1;
  for (uint32_t i = 0; i < n_inner; i++) {
    const fo_layer_proof *lp = &pr->inner_layers[i];
    sparse_eval se;
    uint32_t used = 0;
    int r = rebuild_evals(lq, nlq, evals, nlq, lp->fri_witness, lp->n_fri_witness, &used, &se);
    if (r != 1) { sparse_free(&se); result = r == 0 ? 0 : r; goto done; }
    if (used != lp->n_fri_witness) { sparse_free(&se); goto done; }
    int ok = merkle_verify(lp->commitment, llog, se.positions, se.n_positions,
                           (const uint32_t *)se.subset_evals, 4 * (size_t)se.n_positions, lp->hash_witness,
                           lp->n_hash_witness, lp->n_column_witness);
    if (!ok) { sparse_free(&se); goto done; }
    coset dom = half_odds(llog);
    for (uint32_t s = 0; s < se.n_subsets; s++) {
      uint32_t start = se.positions[2 * s];
      cpoint p = p_from_index(coset_index_at(dom, brev(start, llog)));
      evals[s] = fold_pair(se.subset_evals[2 * s], se.subset_evals[2 * s + 1], m_inv(p.x), alphas[i]);
    }
    uint32_t m = queries_fold(lq, nlq, 1, lq);
    nlq = m;
    llog--;
    sparse_free(&se);
  }
  /* last layer */
  {
    coset dom = half_odds(llog);
    uint32_t plog = 0;
    while ((1u << plog) < pr->n_last_layer_poly) plog++;
    result = 1;
    for (uint32_t s = 0; s < nlq; s++) {
      cpoint p = p_from_index(coset_index_at(dom, brev(lq[s], llog)));
      qm31 dbl[32];
      qm31 x = {{p.x, 0, 0, 0}};
      for (uint32_t b = 0; b < plog; b++) {
        dbl[b] = x;
        qm31 xx = q_mul(x, x);
        x = q_sub(q_add(xx, xx), (qm31){{1, 0, 0, 0}});
      }
      if (!q_eq(evals[s], line_poly_eval(pr->last_layer_poly, plog, dbl))) { result = 0; break; }
    }
  }
done:
  free(alphas); free(q); free(lq); free(evals);
  return result;
}

/* ------------------------------------------------------------------ serialization */
typedef struct { uint8_t *p; size_t cap, n; } wbuf;
static void w_bytes(wbuf *w, const void *src, size_t n) {
  if (w->p && w->n + n <= w->cap) memcpy(w->p + w->n, src, n);
  w->n += n;
}
static void w_u32(wbuf *w, uint32_t v) { uint8_t b[4]; st32le(b, v); w_bytes(w, b, 4); }
static void w_u64(wbuf *w, uint64_t v) { w_u32(w, (uint32_t)v); w_u32(w, (uint32_t)(v >> 32)); }
static void w_qm31s(wbuf *w, const fo_qm31 *q, uint32_t n) {
  w_u32(w, n);
  for (uint32_t i = 0; i < n; i++) for (int j = 0; j < 4; j++) w_u32(w, q[i].v[j]);
}
static void w_layer(wbuf *w, const fo_layer_proof *l) {
  w_bytes(w, l->commitment, 32);
  w_qm31s(w, l->fri_witness, l->n_fri_witness);
  w_u32(w, l->n_hash_witness);
  w_bytes(w, l->hash_witness, 32 * (size_t)l->n_hash_witness);
  w_u32(w, l->n_column_witness);
  for (uint32_t i = 0; i < l->n_column_witness; i++) w_u32(w, l->column_witness[i]);
}
size_t fo_proof_serialize(const fo_proof *p, uint8_t *out, size_t cap) {
  wbuf w = {out, cap, 0};
  w_bytes(&w, "FRDA", 4);
  w_u32(&w, p->log_size_bound);
  w_u32(&w, p->pcs_config.log_blowup_factor);
  w_u32(&w, p->pcs_config.log_last_layer_degree_bound);
  w_u64(&w, p->pcs_config.n_queries);
  w_u32(&w, p->pcs_config.pow_bits);
  w_u64(&w, p->proof_of_work);
  w_qm31s(&w, p->evaluations, p->n_evaluations);
  w_qm31s(&w, p->last_layer_poly, p->n_last_layer_poly);
  w_u32(&w, 1 + p->n_inner_layers);
  w_layer(&w, &p->first_layer);
  for (uint32_t i = 0; i < p->n_inner_layers; i++) w_layer(&w, &p->inner_layers[i]);
  return w.n;
}

/* ------------------------------------------------------------------ trace accessors */
uint32_t fo_trace_poly_log(const fo_trace *t) { return t->poly_log; }
uint32_t fo_trace_n_felts(const fo_trace *t) { return t->n_felts; }
const uint32_t *fo_trace_coeffs(const fo_trace *t) { return t->coeffs; }
const uint32_t *fo_trace_twiddles(const fo_trace *t, int inverse) { return inverse ? t->itw : t->tw; }
uint32_t fo_trace_n_layers(const fo_trace *t) { return t->n_layers; }
uint32_t fo_trace_layer_log(const fo_trace *t, uint32_t layer) { return t->layers[layer].log; }
const uint32_t *fo_trace_layer_column(const fo_trace *t, uint32_t layer, uint32_t coord) {
  return t->layers[layer].cols[coord];
}
const uint8_t *fo_trace_tree_level(const fo_trace *t, uint32_t layer, uint32_t level) {
  return t->layers[layer].tree.levels[level];
}
fo_qm31 fo_trace_alpha(const fo_trace *t, uint32_t layer) { return t->layers[layer].alpha; }
const uint32_t *fo_trace_last_eval(const fo_trace *t, uint32_t *log_out) {
  /* columns are separate allocations; callers read each via fo_trace_last_eval_col */
  if (log_out) *log_out = t->last_log;
  return t->last_cols[0];
}
const uint32_t *fo_trace_last_eval_col(const fo_trace *t, uint32_t coord) { return t->last_cols[coord]; }
const uint8_t *fo_trace_digest_after_fri(const fo_trace *t) { return t->digest_after_fri; }
uint64_t fo_trace_nonce(const fo_trace *t) { return t->nonce; }
uint32_t fo_trace_n_queries(const fo_trace *t) { return t->n_queries; }
const uint32_t *fo_trace_queries(const fo_trace *t) { return t->queries; }
const fo_proof *fo_trace_proof(const fo_trace *t) { return t->proof; }
const uint8_t *fo_trace_root(const fo_trace *t) { return t->root; }
