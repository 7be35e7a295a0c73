// host_math.hpp -- host-side circle-group helpers (stwo circle.rs: CirclePointIndex::to_point,
// Coset::half_odds, CircleDomain::at) used by the verifier and to seed the device twiddle kernel.
#pragma once
#include <cstdint>

#include "m31.cuh"

namespace frieda {
namespace host {

constexpr uint32_t GEN_X = 2u, GEN_Y = 1268011823u;  // generator of the order-2^31 circle group

// G^index, index taken mod 2^31 (LSB-first double-and-add).
inline CPoint point_from_index(uint32_t index) {
  CPoint res = {1u, 0u}, cur = {GEN_X, GEN_Y};
  index &= 0x7fffffffu;
  while (index) {
    if (index & 1u) res = cpoint_add(res, cur);
    cur = cpoint_add(cur, cur);
    index >>= 1;
  }
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
