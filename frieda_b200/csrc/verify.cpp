// verify.cpp -- host-side verifier: frieda::proof::verify_proof (src/proof.rs:79-101).
//
// Rebuilds the channel, replays FriVerifier::commit (mix_root / draw_felt per layer, layer-count
// and last-layer-length checks, mix_felts), checks the proof of work, regenerates the query
// positions, then per layer rebuilds the sibling pairs from (`evaluations` | folded values) and
// fri_witness, checks the sparse Merkle multi-proof, folds, and finally compares with the
// last-layer polynomial.  Sub-millisecond integer work on a few thousand hashes: it stays on
// the host by design, it is not a fallback for any device path.
// The nodes of one tree level are independent, so they are hashed eight abreast (blake2s_x8.cpp, AVX2,
// runtime-checked); the fold twiddles of a layer come from the previous layer's points by the doubling
// map x -> 2x^2 - 1 and are inverted together (one Fermat inversion per layer).
#include <algorithm>
#include <array>
#include <cstring>
#include <vector>

#include "../../include/frieda_b200.h"
#include "blake2s.cuh"
#include "host_math.hpp"
#include "m31.cuh"

namespace frieda {
// message j = 32 bytes at left[j] || 32 bytes at right[j] (any alignment), digest to outs[j]
void merkle_hash_x8_avx2(const void *const left[8], const void *const right[8], uint32_t *const outs[8]);       // blake2s_x8.cpp
void merkle_hash_x16_avx512(const void *const left[16], const void *const right[16], uint32_t *const outs[16]);  // blake2s_x16.cpp
bool cpu_has_avx2();                                                                     // cpu_features.cpp
bool cpu_has_avx512();
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
// One compression: message = left (8 words) || right (8 words), digest to out (8 words).  The halves
// are read where they lie (a child digest, the proof's hash witness, a leaf's values).
struct HashJob {
  const void *left, *right;
  uint32_t *out;
};
const uint32_t ZERO8[8] = {0, 0, 0, 0, 0, 0, 0, 0};

// hash_node of every job (a leaf is the same compression with a zero-padded message): sixteen abreast
// with AVX-512, eight with AVX2, else one at a time.
void hash_many(const std::vector<HashJob> &jobs) {
  const size_t n = jobs.size();
  size_t i = 0;
  uint32_t dummy[8];
  if (n >= 12 && cpu_has_avx512()) {
    for (; i + 8 < n; i += 16) {  // a tail of <= 8 is left to the 8-way path
      const void *lp[16], *rp[16];
      uint32_t *op[16];
      for (size_t j = 0; j < 16; j++) {
        const bool live = i + j < n;
        const HashJob &jb = jobs[live ? i + j : n - 1];
        lp[j] = jb.left;
        rp[j] = jb.right;
        op[j] = live ? jb.out : dummy;
      }
      merkle_hash_x16_avx512(lp, rp, op);
    }
  }
  if (n - i >= 3 && i < n && cpu_has_avx2()) {
    for (; i < n; i += 8) {
      const void *lp[8], *rp[8];
      uint32_t *op[8];
      for (size_t j = 0; j < 8; j++) {
        const bool live = i + j < n;
        const HashJob &jb = jobs[live ? i + j : n - 1];
        lp[j] = jb.left;
        rp[j] = jb.right;
        op[j] = live ? jb.out : dummy;
      }
      merkle_hash_x8_avx2(lp, rp, op);
    }
    return;
  }
  for (; i < n; i++) {
    uint32_t m[16];
    std::memcpy(m, jobs[i].left, 32);
    std::memcpy(m + 8, jobs[i].right, 32);
    merkle_hash_node(m, jobs[i].out);
  }
}

// MerkleVerifier::verify for 4 columns of log size `log`, queried at `pos` (sorted), as a level-by-level
// walk: prepare() lists the nodes of the current level and their messages, the caller hashes them
// (together with the other layers' nodes), advance() moves one level up.
struct MerkleWalk {
  const uint8_t *root;
  const std::vector<uint32_t> *pos;
  const std::vector<QM31> *values;
  const frieda_layer_proof *lp;
  std::vector<Node> prev, cur;
  int k;
  uint32_t log, hw_used = 0;
  size_t val_used = 0;
  bool have_prev = false, failed = false, done = false;

  MerkleWalk(const uint8_t *root_, uint32_t log_, const std::vector<uint32_t> *pos_, const std::vector<QM31> *values_,
             const frieda_layer_proof *lp_)
      : root(root_), pos(pos_), values(values_), lp(lp_), k((int)log_), log(log_) {
    // a level never has more nodes than the leaf level: no reallocation while jobs point into cur / leaves
    prev.reserve(pos_->size() + 1);
    cur.reserve(pos_->size() + 1);
    leaves.reserve(pos_->size() + 1);
  }
  bool active() const { return !failed && !done; }

  // One job per node of the current level.  `left`/`right` of a job point into prev (stable until
  // advance()), the proof, or `leaves`; `out` into cur, whose storage was reserved up front.
  std::vector<std::array<uint32_t, 8>> leaves;  // leaf messages: 4 values + 4 zero words

  void prepare(std::vector<HashJob> &jobs) {
    cur.clear();
    size_t pi = 0, hi = 0, ci = 0;
    const size_t n_prev = prev.size();
    const size_t n_colq = (k == (int)log) ? pos->size() : 0;
    for (;;) {
      bool have = false;
      uint32_t node = 0;
      if (pi < n_prev) {
        node = prev[pi].index >> 1;
        have = true;
      }
      if (ci < n_colq && (!have || (*pos)[ci] < node)) {
        node = (*pos)[ci];
        have = true;
      }
      if (!have) break;
      while (pi < n_prev && (prev[pi].index >> 1) == node) pi++;
      const void *half[2];
      if (have_prev) {
        for (int s = 0; s < 2; s++) {
          if (hi < n_prev && prev[hi].index == 2 * node + s) {
            half[s] = prev[hi].h;
            hi++;
          } else {
            if (hw_used >= lp->n_hash_witness) {  // WitnessTooShort
              failed = true;
              return;
            }
            half[s] = lp->hash_witness + 32 * (size_t)hw_used++;
          }
        }
        // column values below the leaf layer: none (n_columns_in_layer == 0)
        if (ci < n_colq && (*pos)[ci] == node) ci++;
      } else {
        if (!(ci < n_colq && (*pos)[ci] == node) || val_used >= values->size()) {
          failed = true;  // would read the (empty) column witness / TooFewQueriedValues
          return;
        }
        ci++;
        const QM31 &v = (*values)[val_used++];
        leaves.push_back({v.v[0], v.v[1], v.v[2], v.v[3], 0, 0, 0, 0});
        half[0] = leaves.back().data();
        half[1] = ZERO8;
      }
      cur.emplace_back();
      cur.back().index = node;
      jobs.push_back({half[0], half[1], cur.back().h});
    }
  }
  void advance() {
    prev.swap(cur);
    have_prev = true;
    if (--k >= 0) return;
    done = true;
    if (hw_used != lp->n_hash_witness) failed = true;   // WitnessTooLong
    if (val_used != values->size()) failed = true;      // TooManyQueriedValues
    if (lp->n_column_witness != 0) failed = true;       // WitnessTooLong
    if (prev.size() != 1) {
      failed = true;
      return;
    }
    uint8_t got[32];
    hash_words_to_bytes(prev[0].h, got);
    if (std::memcmp(got, root, 32) != 0) failed = true;  // RootMismatch
  }
};

// Runs the walks in lockstep (the layers' trees are independent); true iff every one verifies.
bool merkle_verify_all(std::vector<MerkleWalk> &walks) {
  std::vector<HashJob> jobs;
  for (;;) {
    jobs.clear();
    bool any = false;
    for (MerkleWalk &w : walks)
      if (w.active()) {
        w.prepare(jobs);
        if (w.failed) return false;
        any = true;
      }
    if (!any) break;
    hash_many(jobs);
    for (MerkleWalk &w : walks)
      if (w.active()) {
        w.advance();
        if (w.failed) return false;
      }
  }
  return true;
}

// v[i] <- 1 / v[i] for non-zero v (Montgomery's trick: one inversion, 3 multiplications each)
void batch_inverse(std::vector<uint32_t> &v) {
  const size_t n = v.size();
  if (n == 0) return;
  std::vector<uint32_t> pre(n);
  uint32_t acc = 1;
  for (size_t i = 0; i < n; i++) {
    pre[i] = acc;
    acc = m31_mul(acc, v[i]);
  }
  uint32_t inv = m31_inv(acc);
  for (size_t i = n; i-- > 0;) {
    const uint32_t vi = v[i];
    v[i] = m31_mul(inv, pre[i]);
    inv = m31_mul(inv, vi);
  }
}

// For sorted unique positions `q` of one layer: per sibling pair {2g, 2g+1} (in order) the index of
// its first member in q.
std::vector<size_t> pair_representatives(const std::vector<uint32_t> &q) {
  std::vector<size_t> rep;
  for (size_t i = 0; i < q.size();) {
    rep.push_back(i);
    const uint32_t g = q[i] >> 1;
    while (i < q.size() && (q[i] >> 1) == g) i++;
  }
  return rep;
}

// LinePoly::eval_at_point: coefficients in storage order, fold over the doublings of x.
QM31 line_poly_eval(const QM31 *coeffs, uint32_t log, const QM31 *doublings) {
  if (log == 0) return coeffs[0];
  size_t half = (size_t)1 << (log - 1);
  QM31 l = line_poly_eval(coeffs, log - 1, doublings + 1);
  QM31 r = line_poly_eval(coeffs + half, log - 1, doublings + 1);
  return qm31_add(l, qm31_mul(r, doublings[0]));
}

// positions_out != nullptr: stop after the Fiat-Shamir replay and hand back the query positions (sorted, distinct).
int verify_impl(const frieda_proof *pr, const uint64_t *seed, std::vector<uint32_t> *positions_out = nullptr) {
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
  if (positions_out) {
    *positions_out = q;
    return 1;
  }
  // FriVerifier::decommit (src/proof.rs:98-100).  The field work of all layers runs first (it does not
  // depend on any hash), then the layers' Merkle multi-proofs are checked in lockstep so that every tree
  // level of every layer is one batch of independent compressions.  The reference goes layer by layer
  // and stops at the first error or panic: `pending` holds the first non-Merkle event, which only
  // counts once every layer before it has passed its Merkle check.
  enum { NONE, REJECT, PANIC } pending = NONE;
  std::vector<SparseEval> ses;
  ses.reserve((size_t)pr->n_inner_layers + 1);  // the walks keep pointers into these
  std::vector<MerkleWalk> walks;
  walks.reserve((size_t)pr->n_inner_layers + 1);
  // run one layer's rebuild; true when the layer is complete and its walk is queued
  auto rebuild_layer = [&](const std::vector<uint32_t> &pos, const QM31 *qe, size_t n_qe, const frieda_layer_proof &lp,
                           uint32_t log) -> bool {
    ses.emplace_back();
    SparseEval &se = ses.back();
    uint32_t used = 0;
    bool ok;
    try {
      ok = rebuild(pos, qe, n_qe, lp.fri_witness, lp.n_fri_witness, used, se);
    } catch (const PanicError &) {
      pending = PANIC;
      return false;
    }
    if (!ok || used != lp.n_fri_witness) {
      pending = REJECT;
      return false;
    }
    walks.emplace_back(lp.commitment, log, &se.positions, &se.subset_evals, &lp);
    return true;
  };
  // Domain points of the queried positions.  Bit-reversed neighbours are conjugates on the circle
  // ((x, y), (x, -y)) and antipodes on a line (x, -x), so the twiddle of the pair {2g, 2g+1} is the
  // coordinate at 2g: that of the queried member, negated when the member is the odd one.  The line
  // point below position pos is the double of the point at pos: x' = 2 x^2 - 1 (half_odds(k-1) ==
  // half_odds(k).double()), so only the first layer needs a scalar multiplication per query.
  std::vector<QM31> evals;
  std::vector<uint32_t> xs;  // x-coordinate of the current layer's queried positions
  std::vector<uint32_t> lq;
  uint32_t llog = D - 1;
  {
    std::vector<QM31> qe(pr->n_evaluations);
    for (uint32_t i = 0; i < pr->n_evaluations; i++) qe[i] = to_q(pr->evaluations[i]);
    if (rebuild_layer(q, qe.data(), qe.size(), pr->first_layer, D)) {
      if (pr->n_inner_layers == 0) {
        pending = PANIC;  // assert!(first_layer_columns.is_empty())
      } else {
        const SparseEval &first = ses.back();
        const std::vector<size_t> rep = pair_representatives(q);
        std::vector<uint32_t> tw(rep.size());
        xs.resize(rep.size());
        for (size_t s = 0; s < rep.size(); s++) {
          const uint32_t pos = q[rep[s]];
          CPoint p = host::circle_domain_at(D, bit_reverse(pos, D));
          tw[s] = (pos & 1u) ? m31_neg(p.y) : p.y;
          xs[s] = p.x;  // line layer 1, position pos >> 1
        }
        batch_inverse(tw);
        for (size_t s = 0; s * 2 < first.positions.size(); s++)
          evals.push_back(fri_fold_pair(first.subset_evals[2 * s], first.subset_evals[2 * s + 1], tw[s], alpha0));
        lq = fold_queries(q, 1);
      }
    }
  }
  for (uint32_t i = 0; i < pr->n_inner_layers && pending == NONE; i++) {
    const frieda_layer_proof &lp = pr->inner_layers[i];
    if (!rebuild_layer(lq, evals.data(), evals.size(), lp, llog)) break;
    const SparseEval &se = ses.back();
    const std::vector<size_t> rep = pair_representatives(lq);
    std::vector<uint32_t> tw(rep.size()), nxs(rep.size());
    for (size_t s = 0; s < rep.size(); s++) {
      const uint32_t x = xs[rep[s]];
      tw[s] = (lq[rep[s]] & 1u) ? m31_neg(x) : x;
      const uint32_t xx = m31_mul(x, x);
      nxs[s] = m31_sub(m31_add(xx, xx), 1u);
    }
    batch_inverse(tw);
    std::vector<QM31> next;
    for (size_t s = 0; s * 2 < se.positions.size(); s++)
      next.push_back(fri_fold_pair(se.subset_evals[2 * s], se.subset_evals[2 * s + 1], tw[s], alphas[i]));
    evals.swap(next);
    xs.swap(nxs);
    lq = fold_queries(lq, 1);
    llog--;
  }
  if (!merkle_verify_all(walks)) return 0;
  if (pending == PANIC) throw PanicError{};
  if (pending == REJECT) return 0;
  // last layer
  uint32_t plog = 0;
  while ((1u << plog) < pr->n_last_layer_poly) plog++;
  for (size_t s = 0; s < lq.size(); s++) {
    QM31 dbl[32];
    QM31 x = {{xs[s], 0, 0, 0}};
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

extern "C" long long frieda_proof_query_positions(const frieda_proof *proof, const uint64_t *seed_or_null,
                                                  uint32_t *positions_out, size_t cap) {
  if (!proof) return FRIEDA_ERR_ARG;
  try {
    std::vector<uint32_t> q;
    const int rc = frieda::verify_impl(proof, seed_or_null, &q);
    if (rc != 1) return rc;  // 0: the transcript does not lead to queries (layer count, last layer, proof of work)
    if (positions_out)
      for (size_t i = 0; i < q.size() && i < cap; i++) positions_out[i] = q[i];
    return (long long)q.size();
  } catch (const frieda::PanicError &) {
    return FRIEDA_ERR_PANIC;
  } catch (const std::bad_alloc &) {
    return FRIEDA_ERR_ALLOC;
  }
}

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
