// blake2s.cuh -- BLAKE2s compression held in registers, and the Fiat-Shamir channel.
//
// Merkle hashing (stwo vcs::blake2_merkle::Blake2sMerkleHasher::hash_node, called from
// src/commit.rs:17-22 and inside FriProver via src/proof.rs:52) is the bare compression
// function with an all-zero initial state and t = f = 0 (SURVEY finding 3).
// The channel (stwo channel::Blake2sChannel; src/proof.rs:39-42,52,58-60,80-96) uses real
// BLAKE2s-256 for mix_root / mix_felts / draw_random_bytes and the raw compression for mix_u64.
// "parity unpinned": the channel byte layouts are restated from stwo's published algorithm
// (SURVEY A.9); they live only in this header so they can be switched in one place.
#pragma once
#include <cstdint>

#include "m31.cuh"

namespace frieda {

#if defined(__CUDA_ARCH__)
FR_D uint32_t rotr16(uint32_t x) { return __byte_perm(x, 0, 0x1032); }
FR_D uint32_t rotr8(uint32_t x) { return __byte_perm(x, 0, 0x0321); }
FR_D uint32_t rotr12(uint32_t x) { return __funnelshift_r(x, x, 12); }
FR_D uint32_t rotr7(uint32_t x) { return __funnelshift_r(x, x, 7); }
#else
inline uint32_t rotr_(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline uint32_t rotr16(uint32_t x) { return rotr_(x, 16); }
inline uint32_t rotr8(uint32_t x) { return rotr_(x, 8); }
inline uint32_t rotr12(uint32_t x) { return rotr_(x, 12); }
inline uint32_t rotr7(uint32_t x) { return rotr_(x, 7); }
#endif

#define FR_B2S_IV0 0x6A09E667u
#define FR_B2S_IV1 0xBB67AE85u
#define FR_B2S_IV2 0x3C6EF372u
#define FR_B2S_IV3 0xA54FF53Au
#define FR_B2S_IV4 0x510E527Fu
#define FR_B2S_IV5 0x9B05688Cu
#define FR_B2S_IV6 0x1F83D9ABu
#define FR_B2S_IV7 0x5BE0CD19u

// Pipe balance (measured on B200, bench_micro/pipes.cu): LOP3 / SHF / PRMT / IADD3 issue at 0.5
// warp-instructions per clock per SM sub-partition on the ALU pipe, IMAD at 0.5 on the FMA pipe, and
// the two overlap (a 1:1 mix reaches 0.98).  A G function needs 4 xors + 4 rotates that only the ALU
// pipe can do, so the 6 adds go to the FMA pipe as IMAD x * one + y with `one` an opaque runtime 1
// (otherwise ptxas folds half of them back into ALU-pipe IADD3s).  MASK marks the message words that
// may be non-zero; adds of the others are dropped at compile time.
#if defined(__CUDA_ARCH__)
FR_D uint32_t fr_add(uint32_t x, uint32_t y, uint32_t one) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y));
  return r;
}
#else
inline uint32_t fr_add(uint32_t x, uint32_t y, uint32_t) { return x + y; }
#endif

#define FR_MADD(a, s) ((((MASK) >> (s)) & 1u) ? fr_add(a, m[s], one) : (a))
#define FR_G(a, b, c, d, sx, sy)   \
  do {                             \
    a = fr_add(FR_MADD(a, sx), b, one); \
    d = rotr16(d ^ a);             \
    c = fr_add(c, d, one);         \
    b = rotr12(b ^ c);             \
    a = fr_add(FR_MADD(a, sy), b, one); \
    d = rotr8(d ^ a);              \
    c = fr_add(c, d, one);         \
    b = rotr7(b ^ c);              \
  } while (0)

#define FR_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
  do {                                                                                  \
    FR_G(v0, v4, v8, v12, s0, s1);                                                      \
    FR_G(v1, v5, v9, v13, s2, s3);                                                      \
    FR_G(v2, v6, v10, v14, s4, s5);                                                     \
    FR_G(v3, v7, v11, v15, s6, s7);                                                     \
    FR_G(v0, v5, v10, v15, s8, s9);                                                     \
    FR_G(v1, v6, v11, v12, s10, s11);                                                   \
    FR_G(v2, v7, v8, v13, s12, s13);                                                    \
    FR_G(v3, v4, v9, v14, s14, s15);                                                    \
  } while (0)

// G with zero message words on compile-time constants (for the constant part of a leaf hash).
struct GConst {
  uint32_t a, b, c, d;
};
constexpr uint32_t c_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
constexpr GConst g_zero_msg(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  a = a + b;
  d = c_rotr(d ^ a, 16);
  c = c + d;
  b = c_rotr(b ^ c, 12);
  a = a + b;
  d = c_rotr(d ^ a, 8);
  c = c + d;
  b = c_rotr(b ^ c, 7);
  return {a, b, c, d};
}

// h <- compress(h, m, t0, t1, f0, f1).  Fully unrolled, state in registers.
// ZERO_LEAF: the caller guarantees h == 0, t = f = 0 and m[4..15] == 0 (a Merkle leaf): columns 2 and 3 of
// round 0's first half then see only constants and are folded at compile time (the opaque IMAD adds would
// otherwise hide that from the compiler).
// Msg: anything indexable by a compile-time word number (a register array, or an accessor that fetches the word
// from shared memory where it is used, which keeps the 16 message words out of the register file).
template <uint32_t MASK, bool ZERO_LEAF = false, class Msg = const uint32_t *>
FR_HD void blake2s_compress_t(uint32_t h[8], const Msg &m, uint32_t t0, uint32_t t1, uint32_t f0,
                              uint32_t f1, uint32_t one) {
  uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  uint32_t v8 = FR_B2S_IV0, v9 = FR_B2S_IV1, v10 = FR_B2S_IV2, v11 = FR_B2S_IV3;
  uint32_t v12 = FR_B2S_IV4 ^ t0, v13 = FR_B2S_IV5 ^ t1, v14 = FR_B2S_IV6 ^ f0, v15 = FR_B2S_IV7 ^ f1;
  if (ZERO_LEAF) {
    static_assert(!ZERO_LEAF || MASK == 0x000Fu, "ZERO_LEAF is the 4-word leaf message");
    FR_G(v0, v4, v8, v12, 0, 1);
    FR_G(v1, v5, v9, v13, 2, 3);
    constexpr GConst c2 = g_zero_msg(0u, 0u, FR_B2S_IV2, FR_B2S_IV6);
    constexpr GConst c3 = g_zero_msg(0u, 0u, FR_B2S_IV3, FR_B2S_IV7);
    v2 = c2.a; v6 = c2.b; v10 = c2.c; v14 = c2.d;
    v3 = c3.a; v7 = c3.b; v11 = c3.c; v15 = c3.d;
    FR_G(v0, v5, v10, v15, 8, 9);
    FR_G(v1, v6, v11, v12, 10, 11);
    FR_G(v2, v7, v8, v13, 12, 13);
    FR_G(v3, v4, v9, v14, 14, 15);
  } else {
    FR_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  }
  FR_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
  FR_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
  FR_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
  FR_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
  FR_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
  FR_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
  FR_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
  FR_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
  FR_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
  h[0] ^= v0 ^ v8;
  h[1] ^= v1 ^ v9;
  h[2] ^= v2 ^ v10;
  h[3] ^= v3 ^ v11;
  h[4] ^= v4 ^ v12;
  h[5] ^= v5 ^ v13;
  h[6] ^= v6 ^ v14;
  h[7] ^= v7 ^ v15;
}
FR_HD void blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1, uint32_t f0,
                            uint32_t f1, uint32_t one = 1u) {
  blake2s_compress_t<0xFFFFu>(h, m, t0, t1, f0, f1, one);
}

// Merkle leaf over the 4 coordinate columns: hash_node(None, [c0, c1, c2, c3]).
FR_HD void merkle_hash_leaf(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[8],
                            uint32_t one = 1u) {
  uint32_t m[16] = {c0, c1, c2, c3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = 0;
  blake2s_compress_t<0x000Fu, true>(out, m, 0, 0, 0, 0, one);
}
// Merkle inner node with the message behind an accessor (see blake2s_compress_t).
template <class Msg>
FR_HD void merkle_hash_node_msg(const Msg &m, uint32_t out[8], uint32_t one = 1u) {
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = 0;
  blake2s_compress_t<0xFFFFu, false, Msg>(out, m, 0, 0, 0, 0, one);
}
// Merkle inner node: hash_node(Some((left, right)), []); m = left || right.
FR_HD void merkle_hash_node(const uint32_t m[16], uint32_t out[8], uint32_t one = 1u) {
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = 0;
  blake2s_compress_t<0xFFFFu>(out, m, 0, 0, 0, 0, one);
}

// ---- Blake2sChannel ------------------------------------------------------
struct Channel {
  uint32_t digest[8];  // 32 digest bytes as little-endian words
  uint32_t n_sent;     // reset whenever the digest changes (u64 in stwo; < 2^32 here)
};
FR_HD void channel_init(Channel &c) {
#pragma unroll
  for (int i = 0; i < 8; i++) c.digest[i] = 0;
  c.n_sent = 0;
}
FR_HD void b2s256_init(uint32_t h[8]) {
  h[0] = FR_B2S_IV0 ^ 0x01010020u;
  h[1] = FR_B2S_IV1;
  h[2] = FR_B2S_IV2;
  h[3] = FR_B2S_IV3;
  h[4] = FR_B2S_IV4;
  h[5] = FR_B2S_IV5;
  h[6] = FR_B2S_IV6;
  h[7] = FR_B2S_IV7;
}
// BLAKE2s-256 of exactly one 64-byte block given as 16 LE words.
FR_HD void b2s256_block64(const uint32_t m[16], uint32_t out[8]) {
  b2s256_init(out);
  blake2s_compress(out, m, 64, 0, 0xFFFFFFFFu, 0);
}
// mix_root: digest = H(digest || root)
FR_HD void channel_mix_root(Channel &c, const uint32_t root[8]) {
  uint32_t m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    m[i] = c.digest[i];
    m[8 + i] = root[i];
  }
  b2s256_block64(m, c.digest);
  c.n_sent = 0;
}
// mix_u64: digest = compress(digest, [lo, hi, 0 x 14], 0, 0, 0, 0)   (raw compression)
FR_HD void channel_mix_u64(Channel &c, uint64_t n) {
  uint32_t m[16] = {(uint32_t)n, (uint32_t)(n >> 32), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  blake2s_compress(c.digest, m, 0, 0, 0, 0);
  c.n_sent = 0;
}
// mix_felts: digest = H(digest || felts as 4 x u32 LE each); any length (multi-block).
FR_HD void channel_mix_felts(Channel &c, const QM31 *felts, uint32_t n) {
  uint32_t h[8];
  b2s256_init(h);
  uint32_t total_words = 8 + 4 * n;  // message length in words
  uint32_t m[16];
  uint32_t w = 0;  // words consumed
  uint64_t t = 0;
  for (;;) {
    uint32_t take = total_words - w;
    bool last = take <= 16;
    if (!last) take = 16;
    for (uint32_t i = 0; i < 16; i++) {
      uint32_t idx = w + i;
      uint32_t val = 0;
      if (i < take) val = idx < 8 ? c.digest[idx] : felts[(idx - 8) >> 2].v[(idx - 8) & 3];
      m[i] = val;
    }
    w += take;
    t += 4ull * take;
    blake2s_compress(h, m, (uint32_t)t, (uint32_t)(t >> 32), last ? 0xFFFFFFFFu : 0u, 0);
    if (last) break;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) c.digest[i] = h[i];
  c.n_sent = 0;
}
// draw_random_bytes: H(digest || n_sent as u64 LE || 24 zero bytes); n_sent += 1
FR_HD void channel_draw_random_words(Channel &c, uint32_t out[8]) {
  uint32_t m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = c.digest[i];
  m[8] = c.n_sent;
#pragma unroll
  for (int i = 9; i < 16; i++) m[i] = 0;
  c.n_sent += 1;
  b2s256_block64(m, out);
}
// draw_felt: first 4 of 8 base felts; retry until all 8 words < 2P (p ~ 2^-28 per draw).
FR_HD QM31 channel_draw_felt(Channel &c) {
  for (;;) {
    uint32_t u[8];
    channel_draw_random_words(c, u);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) ok = ok && (u[i] < 2u * P31);
    if (!ok) continue;
    QM31 r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = u[i] >= P31 ? u[i] - P31 : u[i];
    return r;
  }
}
// trailing zeros of the u128 built little-endian from digest bytes 0..16
FR_HD uint32_t digest_trailing_zeros(const uint32_t d[8]) {
  uint32_t tz = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t w = d[i];
    if (w == 0) {
      tz += 32;
      continue;
    }
#if defined(__CUDA_ARCH__)
    return tz + (uint32_t)(__ffs((int)w) - 1);
#else
    uint32_t k = 0;
    while (!(w & 1u)) {
      w >>= 1;
      k++;
    }
    return tz + k;
#endif
  }
  return 128;
}

}  // namespace frieda
