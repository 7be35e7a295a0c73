// host_math.hpp -- host-side circle-group helpers (stwo circle.rs: CirclePointIndex::to_point,
// Coset::half_odds, CircleDomain::at) used by the verifier and to seed the device twiddle kernel.
#pragma once
#include <cstdint>

#include "m31.cuh"

namespace frieda {
namespace host {

constexpr uint32_t GEN_X = 2u, GEN_Y = 1268011823u;  // generator of the order-2^31 circle group

// G^(2^j), j = 0..30, computed once.
struct GenPowerTable {
  CPoint g[31];
  GenPowerTable() {
    CPoint cur = {GEN_X, GEN_Y};
    for (int j = 0; j < 31; j++) {
      g[j] = cur;
      cur = cpoint_add(cur, cur);
    }
  }
};
inline const GenPowerTable &gen_powers() {
  static const GenPowerTable t;
  return t;
}

// G^index, index taken mod 2^31: one group addition per set bit.
inline CPoint point_from_index(uint32_t index) {
  const GenPowerTable &t = gen_powers();
  CPoint res = {1u, 0u};
  index &= 0x7fffffffu;
  for (int j = 0; index; j++, index >>= 1)
    if (index & 1u) res = cpoint_add(res, t.g[j]);
  return res;
}

// CircleDomain::new(Coset::half_odds(D-1)).at(i): the second half is the conjugate coset.
inline CPoint circle_domain_at(uint32_t D, uint32_t i) {
  const uint32_t half = 1u << (D - 1);
  uint32_t idx = i < half ? half_odds_index(D - 1, i)
                          : (uint32_t)((0x80000000u - half_odds_index(D - 1, i - half)) & 0x7fffffffu);
  return point_from_index(idx);
}

inline uint32_t ceil_log2(uint64_t n) {
  uint32_t l = 0;
  while (((uint64_t)1 << l) < n) l++;
  return l;
}

}  // namespace host
}  // namespace frieda
