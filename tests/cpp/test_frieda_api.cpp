// C++ host-API tests written after the reference's own test modules
// (src/commit.rs:24-39, src/proof.rs:103-194, src/lib.rs:46-86), through include/frieda.hpp.
//   usage: test_frieda_api <blob file> gpu
//          test_frieda_api <blob file> cpu <serialized proof file>     (host-only checks, no device)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>

#include "frieda.hpp"

using frieda::FriConfig;
using frieda::PcsConfig;
using frieda::QM31;
using namespace frieda::proof;

static int failures = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      failures++;                                                         \
    }                                                                     \
  } while (0)

static std::vector<uint8_t> read_file(const char *path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), {});
}

// src/proof.rs:109-116
static const PcsConfig PCS_CONFIG = PcsConfig{20, FriConfig{4, 1, 20}};

template <class F>
static bool panics(F f) {
  try {
    f();
  } catch (const frieda::Panic &) {
    return true;
  }
  return false;
}

static void tamper_suite(const Proof &base, std::optional<uint64_t> seed) {
  CHECK(verify_proof(base, seed));  // test_verify_proof
  {                                  // test_verify_proof_with_invalid_pow
    Proof proof = base;
    proof.proof_of_work() += 1;
    CHECK(!verify_proof(proof, seed));
  }
  {  // test_verify_proof_with_invalid_evaluations
    Proof proof = base;
    proof.evaluations()[0] += QM31::from_u32_unchecked(1, 1, 1, 1);
    CHECK(!verify_proof(proof, seed));
  }
  {  // test_verify_proof_with_invalid_evaluations_order
    Proof proof = base;
    proof.evaluations().reverse();
    CHECK(!verify_proof(proof, seed));
  }
  {  // test_verify_proof_with_invalid_evaluations_length  (#[should_panic])
    Proof proof = base;
    proof.evaluations().pop();
    CHECK(panics([&] { verify_proof(proof, seed); }));
  }
  {  // test_verify_proof_with_invalid_1_evaluation_unordered
    Proof proof = base;
    proof.evaluations().swap(0, 1);
    CHECK(!verify_proof(proof, seed));
  }
}

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::vector<uint8_t> data = read_file(argv[1]);  // include_bytes!("../blob")
  CHECK(data.size() == 262146);
  const std::string mode = argv[2];
  if (mode == "cpu") {
    // no device: compute calls must fail loudly (no CPU fallback); the host verifier still works
    bool threw = false;
    try {
      frieda::api::commit(data, 4);
    } catch (const frieda::Error &e) {
      threw = e.code == FRIEDA_ERR_CUDA;
    }
    CHECK(threw);
    Proof proof = Proof::deserialize(read_file(argv[3]));
    CHECK(proof.inner_layers_len() != 0);
    CHECK(Proof::deserialize(proof.serialize()).serialize() == proof.serialize());
    tamper_suite(proof, std::nullopt);
    CHECK(!verify_proof(proof, 1));
  } else {
    // test_commit (src/commit.rs:28-38)
    const frieda::commit::Commitment golden = {209, 162, 213, 6,   157, 197, 135, 229, 93,  194, 156,
                                               198, 37,  90,  249, 55,  255, 127, 237, 14,  228, 27,
                                               223, 90,  249, 135, 23,  249, 215, 79,  96,  232};
    CHECK(frieda::commit::commit(data, 4) == golden);
    // test_generate_proof
    Proof proof = generate_proof(data, std::nullopt, PCS_CONFIG);
    CHECK(proof.inner_layers_len() != 0);
    // test_commit_and_generate_proof
    auto [commitment, proof2] = commit_and_generate_proof(data, std::nullopt, PCS_CONFIG);
    CHECK(commitment == frieda::commit::commit(data, PCS_CONFIG.fri_config.log_blowup_factor));
    CHECK(proof2.first_layer_commitment() == commitment);
    CHECK(proof2.serialize() == proof.serialize());
    // verify + tamper tests
    tamper_suite(proof, std::nullopt);
    // test_verify_proof_with_seed
    Proof p1 = generate_proof(data, 1, PCS_CONFIG), p2 = generate_proof(data, 2, PCS_CONFIG);
    CHECK(p1.evaluations_vec() != p2.evaluations_vec());
    CHECK(verify_proof(p1, 1));
    CHECK(verify_proof(p2, 2));
    CHECK(!verify_proof(p1, 2));
    CHECK(!verify_proof(p2, 1));
    // test_end_to_end (src/lib.rs:52-85)
    const std::string s = "This is the original data that needs to be made available.";
    const std::vector<uint8_t> original_data(s.begin(), s.end());
    auto c = frieda::api::commit(original_data, 4);
    Proof e2e = frieda::api::generate_proof(original_data, std::nullopt, PcsConfig{20, FriConfig{4, 0, 20}});
    CHECK(e2e.first_layer_commitment() == c);
    CHECK(frieda::api::verify(e2e, std::nullopt));
    // reference panics are surfaced as frieda::Panic
    CHECK(panics([&] { frieda::api::commit(std::vector<uint8_t>{1}, 0); }));
  }
  std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
  return failures ? 1 : 0;
}
