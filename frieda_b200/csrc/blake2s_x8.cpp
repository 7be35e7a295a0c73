// blake2s_x8.cpp -- eight Merkle hashes at a time on the host (AVX2), for the host verifier.
//
// verify_proof (src/proof.rs:79-101) checks, per FRI layer, a sparse Merkle multi-proof: at every tree
// level the queried nodes are independent compressions of stwo's Blake2sMerkleHasher::hash_node (all-zero
// initial state, t = f = 0).  This file hashes them eight abreast, one job per 32-bit lane.  It is
// compiled with -mavx2 by the host compiler (frieda_b200/build.py) and only entered after a runtime
// CPU check in verify.cpp; the scalar path of blake2s.cuh remains for other CPUs.  Host-only: the
// device hashing is csrc/merkle.cu.
#include <immintrin.h>

#include <cstddef>
#include <cstdint>

namespace frieda {

namespace {

inline __m256i rotr16(__m256i x) {
  const __m256i k = _mm256_setr_epi8(2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8, 9, 14, 15, 12, 13, 2, 3, 0, 1, 6, 7, 4, 5, 10, 11, 8,
                                     9, 14, 15, 12, 13);
  return _mm256_shuffle_epi8(x, k);
}
inline __m256i rotr8(__m256i x) {
  const __m256i k = _mm256_setr_epi8(1, 2, 3, 0, 5, 6, 7, 4, 9, 10, 11, 8, 13, 14, 15, 12, 1, 2, 3, 0, 5, 6, 7, 4, 9, 10, 11,
                                     8, 13, 14, 15, 12);
  return _mm256_shuffle_epi8(x, k);
}
template <int N>
inline __m256i rotr(__m256i x) {
  return _mm256_or_si256(_mm256_srli_epi32(x, N), _mm256_slli_epi32(x, 32 - N));
}

#define G8(a, b, c, d, x, y)                          \
  do {                                                \
    a = _mm256_add_epi32(_mm256_add_epi32(a, b), x);  \
    d = rotr16(_mm256_xor_si256(d, a));               \
    c = _mm256_add_epi32(c, d);                       \
    b = rotr<12>(_mm256_xor_si256(b, c));             \
    a = _mm256_add_epi32(_mm256_add_epi32(a, b), y);  \
    d = rotr8(_mm256_xor_si256(d, a));                \
    c = _mm256_add_epi32(c, d);                       \
    b = rotr<7>(_mm256_xor_si256(b, c));              \
  } while (0)

// rows r[0..7] (one job each, 8 words) -> columns (one word each, 8 jobs), in place
inline void transpose8(__m256i r[8]) {
  __m256i t0 = _mm256_unpacklo_epi32(r[0], r[1]), t1 = _mm256_unpackhi_epi32(r[0], r[1]);
  __m256i t2 = _mm256_unpacklo_epi32(r[2], r[3]), t3 = _mm256_unpackhi_epi32(r[2], r[3]);
  __m256i t4 = _mm256_unpacklo_epi32(r[4], r[5]), t5 = _mm256_unpackhi_epi32(r[4], r[5]);
  __m256i t6 = _mm256_unpacklo_epi32(r[6], r[7]), t7 = _mm256_unpackhi_epi32(r[6], r[7]);
  __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
  __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
  __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
  __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
  r[0] = _mm256_permute2x128_si256(u0, u4, 0x20);
  r[1] = _mm256_permute2x128_si256(u1, u5, 0x20);
  r[2] = _mm256_permute2x128_si256(u2, u6, 0x20);
  r[3] = _mm256_permute2x128_si256(u3, u7, 0x20);
  r[4] = _mm256_permute2x128_si256(u0, u4, 0x31);
  r[5] = _mm256_permute2x128_si256(u1, u5, 0x31);
  r[6] = _mm256_permute2x128_si256(u2, u6, 0x31);
  r[7] = _mm256_permute2x128_si256(u3, u7, 0x31);
}

const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

}  // namespace

// outs[j] = compress(h = 0, m = left[j] || right[j], t = 0, f = 0) for j = 0..7 (8 LE words per half,
// any alignment).
void merkle_hash_x8_avx2(const void *const left[8], const void *const right[8], uint32_t *const outs[8]) {
  __m256i m[16];
  {
    __m256i lo[8], hi[8];
    for (int j = 0; j < 8; j++) {
      lo[j] = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(left[j]));
      hi[j] = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(right[j]));
    }
    transpose8(lo);
    transpose8(hi);
    for (int i = 0; i < 8; i++) {
      m[i] = lo[i];
      m[8 + i] = hi[i];
    }
  }
  const __m256i z = _mm256_setzero_si256();
  __m256i v0 = z, v1 = z, v2 = z, v3 = z, v4 = z, v5 = z, v6 = z, v7 = z;
  __m256i v8 = _mm256_set1_epi32((int)0x6A09E667u), v9 = _mm256_set1_epi32((int)0xBB67AE85u);
  __m256i v10 = _mm256_set1_epi32((int)0x3C6EF372u), v11 = _mm256_set1_epi32((int)0xA54FF53Au);
  __m256i v12 = _mm256_set1_epi32((int)0x510E527Fu), v13 = _mm256_set1_epi32((int)0x9B05688Cu);
  __m256i v14 = _mm256_set1_epi32((int)0x1F83D9ABu), v15 = _mm256_set1_epi32((int)0x5BE0CD19u);
  for (int r = 0; r < 10; r++) {
    const uint8_t *s = SIGMA[r];
    G8(v0, v4, v8, v12, m[s[0]], m[s[1]]);
    G8(v1, v5, v9, v13, m[s[2]], m[s[3]]);
    G8(v2, v6, v10, v14, m[s[4]], m[s[5]]);
    G8(v3, v7, v11, v15, m[s[6]], m[s[7]]);
    G8(v0, v5, v10, v15, m[s[8]], m[s[9]]);
    G8(v1, v6, v11, v12, m[s[10]], m[s[11]]);
    G8(v2, v7, v8, v13, m[s[12]], m[s[13]]);
    G8(v3, v4, v9, v14, m[s[14]], m[s[15]]);
  }
  __m256i h[8] = {_mm256_xor_si256(v0, v8),  _mm256_xor_si256(v1, v9),  _mm256_xor_si256(v2, v10),
                  _mm256_xor_si256(v3, v11), _mm256_xor_si256(v4, v12), _mm256_xor_si256(v5, v13),
                  _mm256_xor_si256(v6, v14), _mm256_xor_si256(v7, v15)};
  transpose8(h);  // columns (word i of 8 jobs) -> rows (8 words of job j)
  for (int j = 0; j < 8; j++) _mm256_storeu_si256(reinterpret_cast<__m256i *>(outs[j]), h[j]);
}

}  // namespace frieda
