// benches/commit.cpp -- the reference's criterion bench `commit` (benches/commit.rs:4-17) over the
// C++ host API: group "commit", inputs i % 256 of 1024 / 4096 / 16384 / 65536 bytes and the bundled
// blob, commit(data, 4).  Usage: commit <path to blob>
#include "criterion_lite.hpp"
#include "frieda.hpp"

using namespace criterion_lite;

static void bench_commit(const char *blob) {
  Group group{"commit"};
  for (auto &data : reference_datas(blob))
    group.bench_with_input(data.size(), [&] { black_box(frieda::api::commit(data, 4)); });
}

int main(int argc, char **argv) {
  if (argc < 2) return std::fprintf(stderr, "usage: %s <blob>\n", argv[0]), 2;
  try {
    bench_commit(argv[1]);
  } catch (const frieda::Error &e) {
    return std::fprintf(stderr, "%s\n", e.what()), 1;
  }
  return 0;
}
