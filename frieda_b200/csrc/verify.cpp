// verify.cpp -- host-side verifier: frieda::proof::verify_proof (src/proof.rs:79-101).
//
// Rebuilds the channel, replays FriVerifier::commit (mix_root / draw_felt per layer, layer-count
// and last-layer-length checks, mix_felts), checks the proof of work, regenerates the query
// positions, then per layer rebuilds the sibling pairs from (`evaluations` | folded values) and
// fri_witness, checks the sparse Merkle multi-proof, folds, and finally compares with the
// last-layer polynomial.  Sub-millisecond integer work on a few thousand hashes: it stays on
// the host by design, it is not a fallback for any device path.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/frieda_b200.h"
#include "blake2s.cuh"
#include "host_math.hpp"
#include "m31.cuh"

namespace frieda {
namespace {

struct PanicError {};  // the reference would panic here

QM31 to_q(const frieda_qm31 &q) { return {{q.v[0], q.v[1], q.v[2], q.v[3]}}; }

void hash_words_to_bytes(const uint32_t h[8], uint8_t out[32]) { std::memcpy(out, h, 32); }
void bytes_to_hash_words(const uint8_t in[32], uint32_t h[8]) { std::memcpy(h, in, 32); }

// Queries::generate
std::vector<uint32_t> generate_queries(Channel &ch, uint32_t log_domain, uint64_t n_queries) {
  std::vector<uint32_t> q;
  q.reserve(n_queries);
  uint32_t mask = log_domain >= 32 ? 0xffffffffu : ((1u << log_domain) - 1);
  while (q.size() < n_queries) {
    uint32_t w[8];
    channel_draw_random_words(ch, w);
    for (int i = 0; i < 8 && q.size() < n_queries; i++) q.push_back(w[i] & mask);
  }
  std::sort(q.begin(), q.end());
  q.erase(std::unique(q.begin(), q.end()), q.end());
  return q;
}
std::vector<uint32_t> fold_queries(const std::vector<uint32_t> &q, uint32_t n) {
  std::vector<uint32_t> out;
  for (uint32_t v : q) {
    uint32_t f = v >> n;
    if (out.empty() || out.back() != f) out.push_back(f);
  }
  return out;
}

struct SparseEval {
  std::vector<uint32_t> positions;   // decommitment positions (sibling pairs)
  std::vector<QM31> subset_evals;    // 2 per subset
};

// compute_decommitment_positions_and_rebuild_evals with fold_step = 1.
// false => InsufficientWitnessError; throws PanicError when query_evals runs out (`unwrap`).
bool rebuild(const std::vector<uint32_t> &q, const QM31 *query_evals, size_t n_query_evals, const frieda_qm31 *wit,
             uint32_t n_wit, uint32_t &wit_used, SparseEval &se) {
  size_t qe = 0, i = 0;
  wit_used = 0;
  while (i < q.size()) {
    uint32_t g = q[i] >> 1;
    size_t j = i;
    while (j < q.size() && (q[j] >> 1) == g) j++;
    size_t k = i;
    se.positions.push_back(2 * g);
    se.positions.push_back(2 * g + 1);
    for (uint32_t pos = 2 * g; pos < 2 * g + 2; pos++) {
      if (k < j && q[k] == pos) {
        k++;
        if (qe >= n_query_evals) throw PanicError{};
        se.subset_evals.push_back(query_evals[qe++]);
      } else {
        if (wit_used >= n_wit) return false;
        se.subset_evals.push_back(to_q(wit[wit_used++]));
      }
    }
    i = j;
  }
  return true;
}

struct Node {
  uint32_t index;
  uint32_t h[8];
};

// MerkleVerifier::verify for 4 columns of log size `log`, queried at `pos` (sorted).
bool merkle_verify(const uint8_t root[32], uint32_t log, const std::vector<uint32_t> &pos,
                   const std::vector<QM31> &values, const frieda_layer_proof &lp) {
  std::vector<Node> prev, cur;
  uint32_t hw_used = 0;
  size_t val_used = 0;
  bool have_prev = false;
  for (int k = (int)log; k >= 0; k--) {
    cur.clear();
    size_t pi = 0, hi = 0, ci = 0;
    const size_t n_colq = (k == (int)log) ? pos.size() : 0;
    for (;;) {
      bool have = false;
      uint32_t node = 0;
      if (pi < prev.size()) {
        node = prev[pi].index / 2;
        have = true;
      }
      if (ci < n_colq && (!have || pos[ci] < node)) {
        node = pos[ci];
        have = true;
      }
      if (!have) break;
      while (pi < prev.size() && prev[pi].index / 2 == node) pi++;
      Node out;
      out.index = node;
      if (have_prev) {
        uint32_t m[16];
        for (int s = 0; s < 2; s++) {
          if (hi < prev.size() && prev[hi].index == 2 * node + s) {
            std::memcpy(m + 8 * s, prev[hi].h, 32);
            hi++;
          } else {
            if (hw_used >= lp.n_hash_witness) return false;  // WitnessTooShort
            bytes_to_hash_words(lp.hash_witness + 32 * (size_t)hw_used++, m + 8 * s);
          }
        }
        // column values below the leaf layer: none (n_columns_in_layer == 0)
        if (ci < n_colq && pos[ci] == node) ci++;
        merkle_hash_node(m, out.h);
      } else {
        if (!(ci < n_colq && pos[ci] == node)) return false;  // would read the (empty) column witness
        ci++;
        if (val_used >= values.size()) return false;  // TooFewQueriedValues
        const QM31 &v = values[val_used++];
        merkle_hash_leaf(v.v[0], v.v[1], v.v[2], v.v[3], out.h);
      }
      cur.push_back(out);
    }
    prev.swap(cur);
    have_prev = true;
  }
  if (hw_used != lp.n_hash_witness) return false;   // WitnessTooLong
  if (val_used != values.size()) return false;      // TooManyQueriedValues
  if (lp.n_column_witness != 0) return false;       // WitnessTooLong
  if (prev.size() != 1) return false;
  uint8_t got[32];
  hash_words_to_bytes(prev[0].h, got);
  return std::memcmp(got, root, 32) == 0;           // RootMismatch
}

// LinePoly::eval_at_point: coefficients in storage order, fold over the doublings of x.
QM31 line_poly_eval(const QM31 *coeffs, uint32_t log, const QM31 *doublings) {
  if (log == 0) return coeffs[0];
  size_t half = (size_t)1 << (log - 1);
  QM31 l = line_poly_eval(coeffs, log - 1, doublings + 1);
  QM31 r = line_poly_eval(coeffs + half, log - 1, doublings + 1);
  return qm31_add(l, qm31_mul(r, doublings[0]));
}

int verify_impl(const frieda_proof *pr, const uint64_t *seed) {
  const frieda_pcs_config &cfg = pr->pcs_config;
  Channel ch;
  channel_init(ch);
  if (seed) channel_mix_u64(ch, *seed);  // src/proof.rs:80-83
  // FriVerifier::commit (src/proof.rs:84-91)
  uint32_t root[8];
  bytes_to_hash_words(pr->first_layer.commitment, root);
  channel_mix_root(ch, root);
  if (pr->log_size_bound == 0) throw PanicError{};  // CirclePolyDegreeBound::fold_to_line underflow
  uint64_t D64 = (uint64_t)pr->log_size_bound + cfg.log_blowup_factor;
  if (D64 > 29) throw PanicError{};
  const uint32_t D = (uint32_t)D64;
  const QM31 alpha0 = channel_draw_felt(ch);
  std::vector<QM31> alphas;
  uint32_t bound = pr->log_size_bound - 1;
  for (uint32_t i = 0; i < pr->n_inner_layers; i++) {
    bytes_to_hash_words(pr->inner_layers[i].commitment, root);
    channel_mix_root(ch, root);
    alphas.push_back(channel_draw_felt(ch));
    if (bound == 0) return 0;  // InvalidNumFriLayers
    bound--;
  }
  if (bound != cfg.log_last_layer_degree_bound) return 0;                 // InvalidNumFriLayers
  if (pr->n_last_layer_poly > (1u << cfg.log_last_layer_degree_bound)) return 0;  // LastLayerDegreeInvalid
  if (pr->n_last_layer_poly == 0 || (pr->n_last_layer_poly & (pr->n_last_layer_poly - 1))) throw PanicError{};
  std::vector<QM31> last_poly(pr->n_last_layer_poly);
  for (uint32_t i = 0; i < pr->n_last_layer_poly; i++) last_poly[i] = to_q(pr->last_layer_poly[i]);
  channel_mix_felts(ch, last_poly.data(), pr->n_last_layer_poly);
  // proof of work (src/proof.rs:92-95)
  channel_mix_u64(ch, pr->proof_of_work);
  if (digest_trailing_zeros(ch.digest) < cfg.pow_bits) return 0;
  // queries (src/proof.rs:96)
  if (cfg.n_queries == 0 || cfg.n_queries > (1u << 24)) throw PanicError{};
  std::vector<uint32_t> q = generate_queries(ch, D, cfg.n_queries);
  // first layer (src/proof.rs:98-100 -> FriVerifier::decommit)
  std::vector<QM31> evals;
  {
    std::vector<QM31> qe(pr->n_evaluations);
    for (uint32_t i = 0; i < pr->n_evaluations; i++) qe[i] = to_q(pr->evaluations[i]);
    SparseEval se;
    uint32_t used = 0;
    if (!rebuild(q, qe.data(), qe.size(), pr->first_layer.fri_witness, pr->first_layer.n_fri_witness, used, se))
      return 0;
    if (used != pr->first_layer.n_fri_witness) return 0;
    if (!merkle_verify(pr->first_layer.commitment, D, se.positions, se.subset_evals, pr->first_layer)) return 0;
    if (pr->n_inner_layers == 0) throw PanicError{};  // assert!(first_layer_columns.is_empty())
    for (size_t s = 0; s * 2 < se.positions.size(); s++) {
      CPoint p = host::circle_domain_at(D, bit_reverse(se.positions[2 * s], D));
      evals.push_back(fri_fold_pair(se.subset_evals[2 * s], se.subset_evals[2 * s + 1], m31_inv(p.y), alpha0));
    }
  }
  std::vector<uint32_t> lq = fold_queries(q, 1);
  uint32_t llog = D - 1;
  for (uint32_t i = 0; i < pr->n_inner_layers; i++) {
    const frieda_layer_proof &lp = pr->inner_layers[i];
    SparseEval se;
    uint32_t used = 0;
    if (!rebuild(lq, evals.data(), evals.size(), lp.fri_witness, lp.n_fri_witness, used, se)) return 0;
    if (used != lp.n_fri_witness) return 0;
    if (!merkle_verify(lp.commitment, llog, se.positions, se.subset_evals, lp)) return 0;
    std::vector<QM31> next;
    for (size_t s = 0; s * 2 < se.positions.size(); s++) {
      CPoint p = host::point_from_index(half_odds_index(llog, bit_reverse(se.positions[2 * s], llog)));
      next.push_back(fri_fold_pair(se.subset_evals[2 * s], se.subset_evals[2 * s + 1], m31_inv(p.x), alphas[i]));
    }
    evals.swap(next);
    lq = fold_queries(lq, 1);
    llog--;
  }
  // last layer
  uint32_t plog = 0;
  while ((1u << plog) < pr->n_last_layer_poly) plog++;
  for (size_t s = 0; s < lq.size(); s++) {
    CPoint p = host::point_from_index(half_odds_index(llog, bit_reverse(lq[s], llog)));
    QM31 dbl[32];
    QM31 x = {{p.x, 0, 0, 0}};
    const QM31 one = {{1, 0, 0, 0}};
    for (uint32_t b = 0; b < plog; b++) {
      dbl[b] = x;
      QM31 xx = qm31_mul(x, x);
      x = qm31_sub(qm31_add(xx, xx), one);
    }
    if (!qm31_eq(evals[s], line_poly_eval(last_poly.data(), plog, dbl))) return 0;  // LastLayerEvaluationsInvalid
  }
  return 1;
}

}  // namespace
}  // namespace frieda

extern "C" int frieda_verify(const frieda_proof *proof, const uint64_t *seed_or_null) {
  if (!proof) return FRIEDA_ERR_ARG;
  try {
    return frieda::verify_impl(proof, seed_or_null);
  } catch (const frieda::PanicError &) {
    return FRIEDA_ERR_PANIC;
  } catch (const std::bad_alloc &) {
    return FRIEDA_ERR_ALLOC;
  }
}
