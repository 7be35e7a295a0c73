// blake2s_x16.cpp -- sixteen Merkle hashes at a time on the host (AVX-512F), for the host verifier.
// Same contract as blake2s_x8.cpp (hash_node: zero initial state, t = f = 0), one job per 32-bit lane of
// a zmm register, rotations as single VPRORD instructions.  Compiled with -mavx512f by the host compiler
// (frieda_b200/build.py) and only entered after a runtime CPU check in verify.cpp.
#include <immintrin.h>

#include <cstddef>
#include <cstdint>

namespace frieda {

namespace {

#define G16(a, b, c, d, x, y)                               \
  do {                                                      \
    a = _mm512_add_epi32(_mm512_add_epi32(a, b), x);        \
    d = _mm512_ror_epi32(_mm512_xor_si512(d, a), 16);       \
    c = _mm512_add_epi32(c, d);                             \
    b = _mm512_ror_epi32(_mm512_xor_si512(b, c), 12);       \
    a = _mm512_add_epi32(_mm512_add_epi32(a, b), y);        \
    d = _mm512_ror_epi32(_mm512_xor_si512(d, a), 8);        \
    c = _mm512_add_epi32(c, d);                             \
    b = _mm512_ror_epi32(_mm512_xor_si512(b, c), 7);        \
  } while (0)

const uint8_t SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

}  // namespace

// outs[j] = compress(h = 0, m = left[j] || right[j], t = 0, f = 0) for j = 0..15 (8 LE words per half,
// any alignment).
void merkle_hash_x16_avx512(const void *const left[16], const void *const right[16], uint32_t *const outs[16]) {
  // word i of the 16 jobs: the halves are staged row by row, then gathered column-wise (16 gathers;
  // the shuffle network of the 8-way version would need 64 instructions here)
  alignas(64) uint32_t tmp[16][16];
  for (int j = 0; j < 16; j++) {
    _mm256_store_si256(reinterpret_cast<__m256i *>(tmp[j]), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(left[j])));
    _mm256_store_si256(reinterpret_cast<__m256i *>(tmp[j] + 8),
                       _mm256_loadu_si256(reinterpret_cast<const __m256i *>(right[j])));
  }
  const __m512i idx = _mm512_setr_epi32(0, 16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 176, 192, 208, 224, 240);
  __m512i m[16];
  for (int i = 0; i < 16; i++) m[i] = _mm512_i32gather_epi32(idx, &tmp[0][i], 4);
  const __m512i z = _mm512_setzero_si512();
  __m512i v0 = z, v1 = z, v2 = z, v3 = z, v4 = z, v5 = z, v6 = z, v7 = z;
  __m512i v8 = _mm512_set1_epi32((int)0x6A09E667u), v9 = _mm512_set1_epi32((int)0xBB67AE85u);
  __m512i v10 = _mm512_set1_epi32((int)0x3C6EF372u), v11 = _mm512_set1_epi32((int)0xA54FF53Au);
  __m512i v12 = _mm512_set1_epi32((int)0x510E527Fu), v13 = _mm512_set1_epi32((int)0x9B05688Cu);
  __m512i v14 = _mm512_set1_epi32((int)0x1F83D9ABu), v15 = _mm512_set1_epi32((int)0x5BE0CD19u);
#pragma GCC unroll 10
  for (int r = 0; r < 10; r++) {
    const uint8_t *s = SIGMA[r];
    G16(v0, v4, v8, v12, m[s[0]], m[s[1]]);
    G16(v1, v5, v9, v13, m[s[2]], m[s[3]]);
    G16(v2, v6, v10, v14, m[s[4]], m[s[5]]);
    G16(v3, v7, v11, v15, m[s[6]], m[s[7]]);
    G16(v0, v5, v10, v15, m[s[8]], m[s[9]]);
    G16(v1, v6, v11, v12, m[s[10]], m[s[11]]);
    G16(v2, v7, v8, v13, m[s[12]], m[s[13]]);
    G16(v3, v4, v9, v14, m[s[14]], m[s[15]]);
  }
  alignas(64) uint32_t h[8][16];
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[0]), _mm512_xor_si512(v0, v8));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[1]), _mm512_xor_si512(v1, v9));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[2]), _mm512_xor_si512(v2, v10));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[3]), _mm512_xor_si512(v3, v11));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[4]), _mm512_xor_si512(v4, v12));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[5]), _mm512_xor_si512(v5, v13));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[6]), _mm512_xor_si512(v6, v14));
  _mm512_store_si512(reinterpret_cast<__m512i *>(h[7]), _mm512_xor_si512(v7, v15));
  for (int j = 0; j < 16; j++)
    for (int i = 0; i < 8; i++) outs[j][i] = h[i][j];
}

}  // namespace frieda
