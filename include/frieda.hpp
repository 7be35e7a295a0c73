// frieda.hpp -- C++ host API above the C ABI (include/frieda_b200.h), mirroring the reference crate's
// public interface name for name:
//
//   frieda::api::commit(data, log_blowup_factor) -> Commitment              src/lib.rs:31
//   frieda::api::generate_proof(data, seed, pcs_config) -> Proof            src/lib.rs:36
//   frieda::api::verify(proof, seed) -> bool                                src/lib.rs:41
//   frieda::commit::commit / frieda::commit::Commitment                     src/commit.rs:9,11
//   frieda::proof::{Proof, generate_proof, commit_and_generate_proof, verify_proof}
//                                                                            src/proof.rs:19-26,28,32,79
//   frieda::PcsConfig { pow_bits, fri_config: FriConfig { log_blowup_factor,
//                       log_last_layer_degree_bound, n_queries } }          (stwo, built by literal
//                                                                            in src/lib.rs:71-78)
//
// The reference host is Rust; no Rust toolchain exists in the build image, so this header is the
// compiled-language host side (the Rust shim of INTEGRATION.md / rust/ is uncompiled).  Where the
// reference panics this API throws frieda::Panic; other failures throw frieda::Error.  There is no
// CPU fallback: compute calls need a CUDA device.  Header-only; link with -lfrieda_b200.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "frieda_b200.h"

namespace frieda {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error("frieda_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
// The reference implementation panics on this input (assert / unwrap).
struct Panic : Error {
  explicit Panic(const std::string &m) : Error(FRIEDA_ERR_PANIC, m) {}
};

struct FriConfig {
  uint32_t log_blowup_factor;
  uint32_t log_last_layer_degree_bound;
  size_t n_queries;
};
struct PcsConfig {
  uint32_t pow_bits;
  FriConfig fri_config;
};
struct QM31 {
  uint32_t a, b, c, d;  // (a + b i) + (c + d i) u
  static QM31 from_u32_unchecked(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {a, b, c, d}; }
  bool operator==(const QM31 &o) const { return a == o.a && b == o.b && c == o.c && d == o.d; }
  bool operator!=(const QM31 &o) const { return !(*this == o); }
  QM31 &operator+=(const QM31 &o) {  // coordinate-wise addition in M31
    auto add = [](uint32_t x, uint32_t y) {
      uint32_t s = x + y;
      return s >= 0x7fffffffu ? s - 0x7fffffffu : s;
    };
    a = add(a, o.a);
    b = add(b, o.b);
    c = add(c, o.c);
    d = add(d, o.d);
    return *this;
  }
};

namespace detail {
inline frieda_pcs_config to_c(const PcsConfig &p) {
  return frieda_pcs_config{p.fri_config.log_blowup_factor, p.fri_config.log_last_layer_degree_bound,
                           (uint64_t)p.fri_config.n_queries, p.pow_bits};
}
struct CtxHolder {
  frieda_ctx *ctx = nullptr;
  ~CtxHolder() {
    if (ctx) frieda_ctx_destroy(ctx);
  }
};
// One context per thread on device 0 (contexts are not internally locked; the reference functions are pure).
inline frieda_ctx *ctx() {
  thread_local CtxHolder h;
  if (!h.ctx) {
    int rc = frieda_ctx_create(0, &h.ctx);
    if (rc) throw Error(rc, frieda_last_error(nullptr));
  }
  return h.ctx;
}
inline void check(int rc) {
  if (rc == 0) return;
  std::string msg = frieda_last_error(ctx());
  if (rc == FRIEDA_ERR_PANIC) throw Panic(msg);
  throw Error(rc, msg);
}
}  // namespace detail

namespace commit {
using Commitment = std::array<uint8_t, 32>;
inline Commitment commit(const uint8_t *data, size_t len, uint32_t log_blowup_factor) {
  Commitment root{};
  detail::check(frieda_commit(detail::ctx(), data, len, log_blowup_factor, root.data()));
  return root;
}
inline Commitment commit(const std::vector<uint8_t> &data, uint32_t log_blowup_factor) {
  return commit(data.data(), data.size(), log_blowup_factor);
}
}  // namespace commit

namespace proof {
using commit::Commitment;

// Owned proof; the fields of the reference's `Proof` are reachable through `raw()` (frieda_proof) and the
// accessors below.  Copyable like the reference's `#[derive(Clone)]`.
class Proof {
 public:
  explicit Proof(frieda_proof *p) : p_(p, frieda_proof_free) {}
  Proof(const Proof &o) : p_(frieda_proof_clone(o.p_.get()), frieda_proof_free) {
    if (!p_) throw Error(FRIEDA_ERR_ALLOC, "clone failed");
  }
  Proof &operator=(const Proof &o) {
    if (this != &o) *this = Proof(o);
    return *this;
  }
  Proof(Proof &&) = default;
  Proof &operator=(Proof &&) = default;
  Proof clone() const { return Proof(*this); }

  frieda_proof *raw() { return p_.get(); }
  const frieda_proof *raw() const { return p_.get(); }
  uint64_t &proof_of_work() { return p_->proof_of_work; }
  uint64_t proof_of_work() const { return p_->proof_of_work; }
  uint32_t log_size_bound() const { return p_->log_size_bound; }
  PcsConfig pcs_config() const {
    return PcsConfig{p_->pcs_config.pow_bits,
                     FriConfig{p_->pcs_config.log_blowup_factor, p_->pcs_config.log_last_layer_degree_bound,
                               (size_t)p_->pcs_config.n_queries}};
  }
  // `proof.proof.first_layer.commitment.0`
  Commitment first_layer_commitment() const {
    Commitment c;
    std::copy(p_->first_layer.commitment, p_->first_layer.commitment + 32, c.begin());
    return c;
  }
  size_t inner_layers_len() const { return p_->n_inner_layers; }
  // `proof.evaluations` as a mutable view (ascending query order)
  struct Evaluations {
    frieda_proof *p;
    size_t len() const { return p->n_evaluations; }
    QM31 &operator[](size_t i) { return reinterpret_cast<QM31 *>(p->evaluations)[i]; }
    void reverse() {
      for (size_t i = 0, j = len(); i + 1 < j; i++, j--) std::swap((*this)[i], (*this)[j - 1]);
    }
    void swap(size_t i, size_t j) { std::swap((*this)[i], (*this)[j]); }
    void pop() {
      if (p->n_evaluations) p->n_evaluations--;
    }
    std::vector<QM31> to_vec() const {
      const QM31 *q = reinterpret_cast<const QM31 *>(p->evaluations);
      return std::vector<QM31>(q, q + p->n_evaluations);
    }
  };
  Evaluations evaluations() { return Evaluations{p_.get()}; }
  std::vector<QM31> evaluations_vec() const { return Evaluations{p_.get()}.to_vec(); }

  std::vector<uint8_t> serialize() const {
    std::vector<uint8_t> out(frieda_proof_serialize(p_.get(), nullptr, 0));
    frieda_proof_serialize(p_.get(), out.data(), out.size());
    return out;
  }
  static Proof deserialize(const std::vector<uint8_t> &bytes) {
    frieda_proof *p = nullptr;
    int rc = frieda_proof_deserialize(bytes.data(), bytes.size(), &p);
    if (rc) throw Error(rc, "malformed proof bytes");
    return Proof(p);
  }

 private:
  std::unique_ptr<frieda_proof, void (*)(frieda_proof *)> p_;
};
static_assert(sizeof(QM31) == sizeof(frieda_qm31), "QM31 layout");

inline std::pair<Commitment, Proof> commit_and_generate_proof(const uint8_t *data, size_t len,
                                                              std::optional<uint64_t> seed, PcsConfig pcs_config) {
  frieda_pcs_config cfg = detail::to_c(pcs_config);
  Commitment root{};
  frieda_proof *p = nullptr;
  uint64_t s = seed.value_or(0);
  detail::check(frieda_prove(detail::ctx(), data, len, seed ? &s : nullptr, &cfg, root.data(), &p));
  return {root, Proof(p)};
}
inline std::pair<Commitment, Proof> commit_and_generate_proof(const std::vector<uint8_t> &data,
                                                              std::optional<uint64_t> seed, PcsConfig pcs_config) {
  return commit_and_generate_proof(data.data(), data.size(), seed, pcs_config);
}
inline Proof generate_proof(const std::vector<uint8_t> &data, std::optional<uint64_t> seed, PcsConfig pcs_config) {
  return commit_and_generate_proof(data, seed, pcs_config).second;
}
// The positions `proof.evaluations` belong to (ascending; the reference leaves them implicit): indices into the
// bit-reversed evaluation domain, from a replay of the verifier's transcript.  Empty when the transcript is rejected.
inline std::vector<uint32_t> query_positions(const Proof &proof, std::optional<uint64_t> seed) {
  uint64_t s = seed.value_or(0);
  long long n = frieda_proof_query_positions(proof.raw(), seed ? &s : nullptr, nullptr, 0);
  if (n == FRIEDA_ERR_PANIC) throw Panic("called `Option::unwrap()` on a `None` value");
  if (n < 0) throw Error((int)n, "query_positions failed");
  std::vector<uint32_t> out((size_t)n);
  frieda_proof_query_positions(proof.raw(), seed ? &s : nullptr, out.data(), out.size());
  return out;
}
// Takes the proof by value like the reference; throws Panic where the reference panics
// (too few evaluations, src/proof.rs:166-173).
inline bool verify_proof(Proof proof, std::optional<uint64_t> seed) {
  uint64_t s = seed.value_or(0);
  int rc = frieda_verify(proof.raw(), seed ? &s : nullptr);
  if (rc == FRIEDA_ERR_PANIC) throw Panic("called `Option::unwrap()` on a `None` value");
  if (rc < 0) throw Error(rc, "verify failed");
  return rc == 1;
}
}  // namespace proof

// Core public API for FRIEDA (src/lib.rs:22-44)
namespace api {
using commit::Commitment;
using proof::Proof;
inline Commitment commit(const std::vector<uint8_t> &data, uint32_t log_blowup_factor) {
  return frieda::commit::commit(data, log_blowup_factor);
}
inline Proof generate_proof(const std::vector<uint8_t> &data, std::optional<uint64_t> seed, PcsConfig pcs_config) {
  return proof::generate_proof(data, seed, pcs_config);
}
inline bool verify(Proof proof, std::optional<uint64_t> seed) { return proof::verify_proof(std::move(proof), seed); }
}  // namespace api

}  // namespace frieda
