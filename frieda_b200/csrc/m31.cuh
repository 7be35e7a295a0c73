// m31.cuh -- arithmetic in M31 (P = 2^31 - 1), CM31 and QM31 for host and device.
//
// Restates stwo's fields::{m31,cm31,qm31} as used by the reference (src/lib.rs:14,
// src/proof.rs:6,63-66): QM31 = CM31[u]/(u^2 - 2 - i), CM31 = M31[i]/(i^2 + 1).
// All stored values are canonical in [0, P).  Reductions use the unsigned-min trick
// (IADD + IMNMX.U32) instead of compare/select.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#define FR_D __device__ __forceinline__
#else
#define FR_HD inline
#define FR_D inline
#endif

namespace frieda {

constexpr uint32_t P31 = 0x7fffffffu;

FR_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

// a, b in [0, P] -> [0, P] (P is accepted as an alias of 0; canonical in => canonical out)
FR_HD uint32_t m31_add(uint32_t a, uint32_t b) {
  uint32_t s = a + b;
  return umin32(s, s - P31);
}
FR_HD uint32_t m31_sub(uint32_t a, uint32_t b) {
  uint32_t d = a - b;
  return umin32(d, d + P31);
}
FR_HD uint32_t m31_neg(uint32_t a) { return a ? P31 - a : 0u; }
// a, b canonical -> canonical
FR_HD uint32_t m31_mul(uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  uint32_t r = (uint32_t)(p & P31) + (uint32_t)(p >> 31);
  return umin32(r, r - P31);
}
// Multiplication by a constant stored doubled (t2 = 2t < 2^32):
// a * t2 = hi * 2^32 + lo, so a * t = hi * 2^31 + lo / 2 == hi + (lo >> 1) (mod P).
FR_HD uint32_t m31_mul_t2(uint32_t a, uint32_t t2) {
  uint64_t p = (uint64_t)a * t2;
  uint32_t r = (uint32_t)(p >> 32) + ((uint32_t)p >> 1);
  return umin32(r, r - P31);
}
FR_HD uint32_t m31_pow(uint32_t a, uint32_t e) {
  uint32_t r = 1;
  while (e) {
    if (e & 1) r = m31_mul(r, a);
    a = m31_mul(a, a);
    e >>= 1;
  }
  return r;
}
FR_HD uint32_t m31_inv(uint32_t a) { return m31_pow(a, P31 - 2); }

struct CM31 {
  uint32_t a, b;
};
FR_HD CM31 cm31_add(CM31 x, CM31 y) { return {m31_add(x.a, y.a), m31_add(x.b, y.b)}; }
FR_HD CM31 cm31_sub(CM31 x, CM31 y) { return {m31_sub(x.a, y.a), m31_sub(x.b, y.b)}; }
// Karatsuba: 3 base multiplications
FR_HD CM31 cm31_mul(CM31 x, CM31 y) {
  uint32_t ac = m31_mul(x.a, y.a), bd = m31_mul(x.b, y.b);
  uint32_t s = m31_mul(m31_add(x.a, x.b), m31_add(y.a, y.b));
  return {m31_sub(ac, bd), m31_sub(m31_sub(s, ac), bd)};
}

struct QM31 {
  uint32_t v[4];  // (v0 + v1 i) + (v2 + v3 i) u
};
FR_HD QM31 qm31_zero() { return {{0u, 0u, 0u, 0u}}; }
FR_HD QM31 qm31_add(QM31 x, QM31 y) {
  return {{m31_add(x.v[0], y.v[0]), m31_add(x.v[1], y.v[1]), m31_add(x.v[2], y.v[2]), m31_add(x.v[3], y.v[3])}};
}
FR_HD QM31 qm31_sub(QM31 x, QM31 y) {
  return {{m31_sub(x.v[0], y.v[0]), m31_sub(x.v[1], y.v[1]), m31_sub(x.v[2], y.v[2]), m31_sub(x.v[3], y.v[3])}};
}
FR_HD QM31 qm31_mul_m31(QM31 x, uint32_t s) {
  return {{m31_mul(x.v[0], s), m31_mul(x.v[1], s), m31_mul(x.v[2], s), m31_mul(x.v[3], s)}};
}
// (A + B u)(C + D u) = (AC + R BD) + ((A + B)(C + D) - AC - BD) u,  R = 2 + i
FR_HD QM31 qm31_mul(QM31 x, QM31 y) {
  CM31 A = {x.v[0], x.v[1]}, B = {x.v[2], x.v[3]}, C = {y.v[0], y.v[1]}, D = {y.v[2], y.v[3]};
  CM31 ac = cm31_mul(A, C), bd = cm31_mul(B, D);
  CM31 cross = cm31_sub(cm31_sub(cm31_mul(cm31_add(A, B), cm31_add(C, D)), ac), bd);
  CM31 rbd = {m31_sub(m31_add(bd.a, bd.a), bd.b), m31_add(bd.a, m31_add(bd.b, bd.b))};
  CM31 lo = cm31_add(ac, rbd);
  return {{lo.a, lo.b, cross.a, cross.b}};
}
FR_HD bool qm31_eq(QM31 x, QM31 y) {
  return x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2] && x.v[3] == y.v[3];
}
FR_HD bool qm31_is_zero(QM31 x) { return (x.v[0] | x.v[1] | x.v[2] | x.v[3]) == 0; }

// Multiplication by a FIXED QM31 (the folding alpha of a blob) as a 4x4 M31 matrix-vector product:
// with alpha = (c0 + c1 i) + (d0 + d1 i) u and R = 2 + i,
//   out0 = a0 c0 - a1 c1 + b0 (2 d0 - d1) - b1 (d0 + 2 d1)
//   out1 = a0 c1 + a1 c0 + b0 (d0 + 2 d1) + b1 (2 d0 - d1)
//   out2 = a0 d0 - a1 d1 + b0 c0 - b1 c1
//   out3 = a0 d1 + a1 d0 + b0 c1 + b1 c0
// Each output is 4 products accumulated in 64 bits (4 (P-1)^2 < 2^64) and reduced once, instead of
// 9 reduced multiplications and ~25 modular additions (Karatsuba form above).
struct QM31Mat {
  uint32_t m[4][4];
};
FR_HD QM31Mat qm31_mat(QM31 al) {
  const uint32_t c0 = al.v[0], c1 = al.v[1], d0 = al.v[2], d1 = al.v[3];
  const uint32_t e = m31_sub(m31_add(d0, d0), d1);  // 2 d0 - d1
  const uint32_t f = m31_add(d0, m31_add(d1, d1));  // d0 + 2 d1
  QM31Mat r = {{{c0, m31_neg(c1), e, m31_neg(f)},
                {c1, c0, f, e},
                {d0, m31_neg(d1), c0, m31_neg(c1)},
                {d1, d0, c1, c0}}};
  return r;
}
// x <= 4 (P-1)^2 (a sum of four products) -> canonical M31
FR_HD uint32_t m31_reduce64(uint64_t x) {
  // 2^31 == 1: fold 31-bit limbs; for x <= 4 (P-1)^2 the three limbs sum below 2^32
  uint32_t s = (uint32_t)(x & P31) + (uint32_t)((x >> 31) & P31) + (uint32_t)(x >> 62);
  s = (s & P31) + (s >> 31);
  return umin32(s, s - P31);
}
FR_HD QM31 qm31_mul_mat(const QM31Mat &M, QM31 x) {
  QM31 r;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) {
    uint64_t acc = (uint64_t)M.m[k][0] * x.v[0];
    acc += (uint64_t)M.m[k][1] * x.v[1];
    acc += (uint64_t)M.m[k][2] * x.v[2];
    acc += (uint64_t)M.m[k][3] * x.v[3];
    r.v[k] = m31_reduce64(acc);
  }
  return r;
}
// fri_fold_pair with the alpha matrix precomputed
FR_HD QM31 fri_fold_pair_mat(QM31 a, QM31 b, uint32_t itw, const QM31Mat &M) {
  QM31 f0 = qm31_add(a, b);
  QM31 f1 = qm31_mul_m31(qm31_sub(a, b), itw);
  return qm31_add(qm31_mul_mat(M, f1), f0);
}

// FRI fold of one pair (stwo fri.rs fold_line / fold_circle_into_line with dst = 0):
//   f0 = a + b, f1 = (a - b) * itw, out = f0 + alpha * f1        (src/proof.rs:52 call path)
FR_HD QM31 fri_fold_pair(QM31 a, QM31 b, uint32_t itw, QM31 alpha) {
  QM31 f0 = qm31_add(a, b);
  QM31 f1 = qm31_mul_m31(qm31_sub(a, b), itw);
  return qm31_add(qm31_mul(alpha, f1), f0);
}

// ---- circle group over M31 (stwo circle.rs): generator G = (2, 1268011823) of order 2^31.
struct CPoint {
  uint32_t x, y;
};
FR_HD CPoint cpoint_add(CPoint p, CPoint q) {
  return {m31_sub(m31_mul(p.x, q.x), m31_mul(p.y, q.y)), m31_add(m31_mul(p.x, q.y), m31_mul(p.y, q.x))};
}
FR_HD uint32_t bit_reverse(uint32_t i, uint32_t bits) {
  if (bits == 0) return 0;
#if defined(__CUDA_ARCH__)
  return __brev(i) >> (32 - bits);
#else
  uint32_t r = 0;
  for (uint32_t b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
  return r;
#endif
}
// Index (mod 2^31) of Coset::half_odds(k).at(i): initial 2^(29-k), step 2^(31-k).
FR_HD uint32_t half_odds_index(uint32_t k, uint32_t i) {
  uint64_t initial = (uint64_t)1 << (29 - k), step = (uint64_t)1 << (31 - k);
  return (uint32_t)((initial + (uint64_t)i * step) & 0x7fffffffu);
}

}  // namespace frieda
