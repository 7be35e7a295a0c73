// split_plan.hpp -- which values and tree nodes a Proof's decommitment consists of, in PROOF ORDER, and which rank
// of a split blob holds each of them (host code shared by ctx.cu, which serves a rank's share, and split_proof.cpp,
// which merges the shares).
//
// Same walks as decommit.cu (FriProver::decommit / MerkleProver::decommit reached from src/proof.rs:60):
// layer l (log size d = D - l) is queried at distinct(q >> l); the sibling of a queried position that is not itself
// queried contributes its value to fri_witness; on the way up, a child (tree level c, index i) of a path node that no
// query covers contributes its hash to hash_witness, child levels d-1 .. 1, ascending index inside a level.
// `evaluations` (src/proof.rs:62-66) are the layer-0 values at the queries.
//
// Ownership under the split of frieda_fri_split_*: rank r holds indices [r 2^(d-g), (r+1) 2^(d-g)) of every split layer
// (l < n_split) and the subtree below node (level g, index r); the top g levels of a split layer's tree and everything
// of the unsplit layers are replicated -- rank 0 serves them.
#pragma once
#include <cstdint>
#include <vector>

namespace frieda {

struct SplitShape {
  uint32_t D, n_layers, n_split, gl;
};
enum SplitKind : uint32_t { SK_EVAL = 0, SK_FRI = 1, SK_HASH = 2 };
struct SplitItem {
  uint32_t kind, layer, level;  // level: tree level of a hash (unused otherwise)
  uint32_t index;               // position inside the layer, or node index inside the tree level
  uint32_t owner;
};

// runs of equal (q >> sh): every run yields the children (bit sh-1) that no query covers (decommit.cu: walk_runs)
template <class Emit>
inline void split_walk(const uint32_t *q, uint32_t nq, uint32_t sh, Emit emit) {
  uint32_t i = 0;
  while (i < nq) {
    const uint32_t node = sh >= 32 ? 0u : q[i] >> sh;
    bool has0 = false, has1 = false;
    while (i < nq && (sh >= 32 ? 0u : q[i] >> sh) == node) {
      if ((q[i] >> (sh - 1)) & 1u) has1 = true; else has0 = true;
      i++;
    }
    if (!has0) emit(2 * node);
    if (!has1) emit(2 * node + 1);
  }
}

// q: sorted distinct query positions of layer 0.  Appends the items of the whole proof in proof order and, per layer,
// the lengths of fri_witness and hash_witness.
inline void split_plan(const SplitShape &s, const uint32_t *q, uint32_t nq, std::vector<SplitItem> &items,
                       std::vector<uint32_t> &n_fri, std::vector<uint32_t> &n_hash) {
  items.clear();
  n_fri.assign(s.n_layers, 0);
  n_hash.assign(s.n_layers, 0);
  for (uint32_t i = 0; i < nq; i++) items.push_back({SK_EVAL, 0u, 0u, q[i], q[i] >> (s.D - s.gl)});
  for (uint32_t l = 0; l < s.n_layers; l++) {
    const uint32_t d = s.D - l;
    const bool split = l < s.n_split;
    split_walk(q, nq, l + 1, [&](uint32_t pos) {
      items.push_back({SK_FRI, l, 0u, pos, split ? pos >> (d - s.gl) : 0u});
      n_fri[l]++;
    });
    for (uint32_t j = 1; j < d; j++) {
      const uint32_t c = d - j;  // tree level of the emitted children (parents on level d - 1 - j)
      split_walk(q, nq, l + 1 + j, [&](uint32_t idx) {
        items.push_back({SK_HASH, l, c, idx, (split && c > s.gl) ? idx >> (c - s.gl) : 0u});
        n_hash[l]++;
      });
    }
  }
}

// little-endian words of a share (frieda_fri_split_decommit) before the item payload
constexpr uint32_t SPLIT_SHARE_MAGIC = 0x48535246u;  // "FRSH"
struct SplitShareHeader {
  uint32_t magic, world, rank, gl, D, n_layers, n_split, log_size_bound;
  uint32_t log_blowup, log_last, nq_lo, nq_hi, pow_bits, nonce_lo, nonce_hi, n_unique;
};

}  // namespace frieda
