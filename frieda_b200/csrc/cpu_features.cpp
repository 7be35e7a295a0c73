// cpu_features.cpp -- runtime ISA checks for the host verifier's 8- and 16-way hashing (compiled without
// ISA flags).  FRIEDA_HOST_ISA=scalar|avx2 caps the level (used by the tests to exercise every path).
#include <cstdlib>
#include <cstring>

namespace frieda {
namespace {
int isa_level() {  // 0 scalar, 1 AVX2, 2 AVX-512F
  static const int level = [] {
    int l = 0;
#if defined(__x86_64__) || defined(__i386__)
    if (__builtin_cpu_supports("avx2")) l = 1;
    if (l == 1 && __builtin_cpu_supports("avx512f")) l = 2;
#endif
    if (const char *e = std::getenv("FRIEDA_HOST_ISA")) {
      if (!std::strcmp(e, "scalar")) l = 0;
      if (!std::strcmp(e, "avx2") && l > 1) l = 1;
    }
    return l;
  }();
  return level;
}
}  // namespace
bool cpu_has_avx2() { return isa_level() >= 1; }
bool cpu_has_avx512() { return isa_level() >= 2; }
}  // namespace frieda
