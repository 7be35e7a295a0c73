// benches/proof.cpp -- the reference's criterion benches `generate_proof`,
// `commit_and_generate_proof` and `verify_proof` (benches/proof.rs:14-61) over the C++ host API,
// with its PCS_CONFIG (blowup 2^4, last-layer bound 2^0, 20 queries, 20 proof-of-work bits) and
// seed = Some(data.len()).  Usage: proof <path to blob>
#include "criterion_lite.hpp"
#include "frieda.hpp"

using namespace criterion_lite;
using frieda::proof::commit_and_generate_proof;
using frieda::proof::generate_proof;
using frieda::proof::verify_proof;

static const frieda::PcsConfig PCS_CONFIG{20, frieda::FriConfig{4, 0, 20}};

static void bench_generate_proof(const char *blob) {
  Group group{"generate_proof"};
  for (auto &data : reference_datas(blob))
    group.bench_with_input(data.size(), [&] { black_box(generate_proof(data, (uint64_t)data.size(), PCS_CONFIG)); });
}

static void bench_commit_and_generate_proof(const char *blob) {
  Group group{"commit_and_generate_proof"};
  for (auto &data : reference_datas(blob))
    group.bench_with_input(data.size(),
                           [&] { black_box(commit_and_generate_proof(data, (uint64_t)data.size(), PCS_CONFIG)); });
}

static void bench_verify_proof(const char *blob) {
  Group group{"verify_proof"};
  for (auto &data : reference_datas(blob)) {
    auto cp = commit_and_generate_proof(data, (uint64_t)data.size(), PCS_CONFIG);
    const frieda::proof::Proof &proof = cp.second;
    group.bench_with_input(data.size(), [&] {
      if (!verify_proof(proof.clone(), (uint64_t)data.size())) std::abort();
    });
  }
}

int main(int argc, char **argv) {
  if (argc < 2) return std::fprintf(stderr, "usage: %s <blob>\n", argv[0]), 2;
  try {
    bench_generate_proof(argv[1]);
    bench_commit_and_generate_proof(argv[1]);
    bench_verify_proof(argv[1]);
  } catch (const frieda::Error &e) {
    return std::fprintf(stderr, "%s\n", e.what()), 1;
  }
  return 0;
}
