// criterion_lite.hpp -- the few lines of criterion the reference's benches use
// (benchmark_group / bench_with_input / BenchmarkId::from_parameter), printing one JSON line per
// benchmark: {"group": ..., "param": ..., "iters": ..., "mean_us": ..., "min_us": ...}.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace criterion_lite {

template <class T>
inline void black_box(T const &v) {
  asm volatile("" : : "g"(&v) : "memory");
}

inline double budget_seconds() {
  const char *e = std::getenv("FRIEDA_BENCH_SECONDS");  // measurement time per benchmark
  return e ? std::atof(e) : 1.0;
}

struct Group {
  std::string name;
  template <class F>
  void bench_with_input(size_t param, F &&f) {
    using clk = std::chrono::steady_clock;
    for (int i = 0; i < 3; i++) f();  // warm-up
    const double budget = budget_seconds();
    double total = 0, best = 1e30;
    size_t iters = 0;
    while (total < budget || iters < 5) {
      auto t0 = clk::now();
      f();
      double dt = std::chrono::duration<double>(clk::now() - t0).count();
      total += dt;
      if (dt < best) best = dt;
      iters++;
    }
    std::printf("{\"group\": \"%s\", \"param\": %zu, \"iters\": %zu, \"mean_us\": %.2f, \"min_us\": %.2f}\n", name.c_str(),
                param, iters, total / iters * 1e6, best * 1e6);
    std::fflush(stdout);
  }
};

// the reference's inputs: (0..size).map(|i| (i % 256) as u8) for four sizes, then the bundled blob
inline std::vector<std::vector<uint8_t>> reference_datas(const char *blob_path) {
  std::vector<std::vector<uint8_t>> datas;
  for (size_t size : {1024, 4096, 16384, 65536}) {
    std::vector<uint8_t> d(size);
    for (size_t i = 0; i < size; i++) d[i] = (uint8_t)(i % 256);
    datas.push_back(std::move(d));
  }
  std::ifstream f(blob_path, std::ios::binary);
  if (!f) {
    std::fprintf(stderr, "cannot read %s\n", blob_path);
    std::exit(2);
  }
  datas.emplace_back((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  return datas;
}

}  // namespace criterion_lite
