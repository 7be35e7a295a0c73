// verify_core.cuh -- verifier core shared by the GPU batch verifier (verify_batch.cu) and its host
// self-check: frieda::proof::verify_proof (src/proof.rs:79-101) restated as two data-parallel phases
// over flat-encoded proofs (the "FRDA" encoding of csrc/proof.cpp, read as little-endian u32 words).
//
//   phase A, one thread per proof: FriVerifier::commit (transcript replay, layer-count and
//     last-layer-length checks, mix_felts), proof of work, Queries::generate, then per layer
//     compute_decommitment_positions_and_rebuild_evals + fold, and the last-layer comparison.  It
//     leaves, per layer, the sorted leaf list (position + 4 column values) for phase B.
//   phase B, one thread per (proof, layer): MerkleVerifier::verify over that leaf list and the
//     layer's hash_witness, in place.
//   resolve: combines both in the reference's order of checks (a Merkle failure of an earlier layer wins
//     over a later rebuild failure; the "first layer columns not consumed" panic comes after layer 0's
//     Merkle check).
// This is an independent second implementation next to csrc/verify.cpp (std::vector based, host only).
#pragma once
#include <cstdint>

#include "blake2s.cuh"
#include "m31.cuh"

namespace frieda {

struct VGen {
  CPoint g[31];  // G^(2^j)
};

FR_HD CPoint v_point_from_index(const VGen &gp, uint32_t idx) {
  CPoint r = {1u, 0u};
  for (int j = 0; j < 31; j++)
    if ((idx >> j) & 1u) r = cpoint_add(r, gp.g[j]);
  return r;
}

// entry of a layer's node list: leaf = (position, 4 column values) before hashing, (index, hash) after
struct VNode {
  uint32_t idx;
  uint32_t w[8];
};

constexpr uint32_t V_MAX_LAYERS = 32;

// per-proof record written by phase A
struct VProofState {
  int32_t a_kind;       // 1 ok, 0 reject, -1 reference panics
  uint32_t a_layer;     // layer at which phase A stopped (valid when a_kind != 1)
  uint32_t a_post;      // 0: the failure precedes that layer's Merkle check, 1: follows it
  uint32_t n_layers;    // 1 + inner layers present in the proof
  uint32_t D;           // log size of layer 0
  uint32_t n_leaf[V_MAX_LAYERS];    // leaf-list length per layer (0: not produced)
  uint32_t layer_off[V_MAX_LAYERS]; // word offset of each layer record inside the proof
  int32_t merkle_ok[V_MAX_LAYERS];  // phase B result per layer (1 ok, 0 fail)
};

struct VReader {
  const uint32_t *w;
  uint32_t n, pos;
  bool ok;
  FR_HD uint32_t u32() {
    if (pos >= n) {
      ok = false;
      return 0;
    }
    return w[pos++];
  }
  FR_HD const uint32_t *take(uint32_t k) {
    if (!ok || k > n - pos) {
      ok = false;
      return w;
    }
    const uint32_t *p = w + pos;
    pos += k;
    return p;
  }
  // `count` items of `words_per_item` words each.  The count is an untrusted word of the proof: the product is
  // formed in 64 bits, so a count like 0x20000000 cannot wrap to a small `take` and pass the bounds check.
  FR_HD const uint32_t *take_items(uint32_t count, uint32_t words_per_item) {
    const uint64_t k = (uint64_t)count * words_per_item;
    if (!ok || k > (uint64_t)(n - pos)) {
      ok = false;
      return w;
    }
    return take((uint32_t)k);
  }
};

FR_HD QM31 v_q(const uint32_t *p) { return {{p[0], p[1], p[2], p[3]}}; }

// Phase A.  `leaves` holds max_pos entries per layer (layer-major); max_q = capacity of `queries`.
FR_HD void verify_phase_a(const uint32_t *words, uint32_t n_words, const uint64_t *seed, const VGen &gp,
                          VProofState &st, VNode *leaves, uint32_t max_pos, uint32_t *queries, uint32_t max_q,
                          QM31 *evals /* max_q */, QM31 *alphas /* V_MAX_LAYERS */,
                          uint32_t max_layers = V_MAX_LAYERS) {
  st.a_kind = 1;
  st.a_layer = 0;
  st.a_post = 0;
  st.n_layers = 0;
  st.D = 0;
  for (uint32_t l = 0; l < V_MAX_LAYERS; l++) {
    st.n_leaf[l] = 0;
    st.layer_off[l] = 0;
    st.merkle_ok[l] = 1;
  }
  auto stop = [&](int kind, uint32_t layer, uint32_t post) {
    st.a_kind = kind;
    st.a_layer = layer;
    st.a_post = post;
  };
  VReader r{words, n_words, 0, true};
  if (r.u32() != 0x41445246u) return stop(0, 0, 0);  // "FRDA"
  const uint32_t log_size_bound = r.u32(), log_blowup = r.u32(), log_last = r.u32();
  const uint32_t nq_lo = r.u32(), nq_hi = r.u32();
  const uint32_t pow_bits = r.u32();
  const uint32_t pow_lo = r.u32(), pow_hi = r.u32();
  const uint32_t n_evals = r.u32();
  const uint32_t *ev = r.take_items(n_evals, 4);
  const uint32_t n_last = r.u32();
  const uint32_t *last = r.take_items(n_last, 4);
  const uint32_t n_layers = r.u32();
  if (!r.ok || n_layers == 0 || n_layers > V_MAX_LAYERS || n_layers > max_layers) return stop(0, 0, 0);
  st.n_layers = n_layers;
  // locate the layers
  const uint32_t *commit[V_MAX_LAYERS];
  for (uint32_t l = 0; l < n_layers; l++) {
    st.layer_off[l] = r.pos;
    commit[l] = r.take(8);
    uint32_t nf = r.u32();
    r.take_items(nf, 4);
    uint32_t nh = r.u32();
    r.take_items(nh, 8);
    uint32_t nc = r.u32();
    r.take(nc);
    if (!r.ok) return stop(0, 0, 0);
  }
  // ---- FriVerifier::commit
  Channel ch;
  channel_init(ch);
  if (seed) channel_mix_u64(ch, *seed);
  channel_mix_root(ch, commit[0]);
  if (log_size_bound == 0) return stop(-1, 0, 0);  // CirclePolyDegreeBound::fold_to_line underflow
  if ((uint64_t)log_size_bound + log_blowup > 29) return stop(-1, 0, 0);
  const uint32_t D = log_size_bound + log_blowup;
  st.D = D;
  alphas[0] = channel_draw_felt(ch);
  uint32_t bound = log_size_bound - 1;
  for (uint32_t l = 1; l < n_layers; l++) {
    channel_mix_root(ch, commit[l]);
    alphas[l] = channel_draw_felt(ch);
    if (bound == 0) return stop(0, 0, 0);  // InvalidNumFriLayers
    bound--;
  }
  if (bound != log_last) return stop(0, 0, 0);
  if (log_last > 20 || n_last > (1u << log_last)) return stop(0, 0, 0);  // LastLayerDegreeInvalid
  if (n_last == 0 || (n_last & (n_last - 1))) return stop(-1, 0, 0);
  channel_mix_felts(ch, reinterpret_cast<const QM31 *>(last), n_last);
  channel_mix_u64(ch, (uint64_t)pow_lo | ((uint64_t)pow_hi << 32));
  if (digest_trailing_zeros(ch.digest) < pow_bits) return stop(0, 0, 0);
  // ---- Queries::generate
  if (nq_hi != 0 || nq_lo == 0) return stop(-1, 0, 0);
  if (nq_lo > max_q) return stop(0, 0, 0);  // capacity of this batch (callers size it from the proofs)
  uint32_t nq = 0;
  {
    const uint32_t mask = (1u << D) - 1;
    uint32_t cnt = 0;
    while (cnt < nq_lo) {
      uint32_t w8[8];
      channel_draw_random_words(ch, w8);
      for (int i = 0; i < 8 && cnt < nq_lo; i++, cnt++) {
        uint32_t v = w8[i] & mask;
        uint32_t lo = 0, hi = nq;
        while (lo < hi) {
          uint32_t mid = (lo + hi) >> 1;
          if (queries[mid] < v) lo = mid + 1; else hi = mid;
        }
        if (lo < nq && queries[lo] == v) continue;
        for (uint32_t j = nq; j > lo; j--) queries[j] = queries[j - 1];
        queries[lo] = v;
        nq++;
      }
    }
  }
  // ---- per layer: rebuild the sibling pairs, record the leaves, fold
  uint32_t n_cur_evals = 0;  // number of folded values carried into the current layer
  for (uint32_t l = 0; l < n_layers; l++) {
    VReader lr{words, n_words, st.layer_off[l], true};
    lr.take(8);
    const uint32_t n_fri = lr.u32();
    const uint32_t *fri = lr.take_items(n_fri, 4);
    if (!lr.ok) return stop(0, l, 0);
    const uint32_t d = D - l;
    VNode *out = leaves + (size_t)l * max_pos;
    uint32_t n_out = 0, wit = 0, qe = 0, n_next = 0;
    // layer queries = distinct(q >> l); groups = distinct(q >> (l + 1))
    uint32_t i = 0;
    while (i < nq) {
      const uint32_t g = queries[i] >> (l + 1);
      bool has0 = false, has1 = false;
      while (i < nq && (queries[i] >> (l + 1)) == g) {
        if ((queries[i] >> l) & 1u) has1 = true; else has0 = true;
        i++;
      }
      QM31 pair[2];
      for (uint32_t s = 0; s < 2; s++) {
        if (s == 0 ? has0 : has1) {
          if (l == 0) {
            if (qe >= n_evals) return stop(-1, l, 0);  // `query_evals.next().unwrap()` (src/proof.rs:166-173)
            pair[s] = v_q(ev + 4 * qe);
          } else {
            pair[s] = evals[qe];  // qe < n_cur_evals by construction
          }
          qe++;
        } else {
          if (wit >= n_fri) return stop(0, l, 0);  // InsufficientWitnessError
          pair[s] = v_q(fri + 4 * wit);
          wit++;
        }
        if (n_out >= max_pos) return stop(0, l, 0);
        out[n_out].idx = 2 * g + s;
        for (int c = 0; c < 4; c++) out[n_out].w[c] = pair[s].v[c];
        n_out++;
      }
      // fold this pair (the values feed the next layer's rebuild; evals[] is consumed front to back,
      // and n_next <= qe always holds, so writing in place is safe)
      uint32_t itw;
      if (l == 0) {
        // p = CircleDomain(half_odds(D-1)).at(brev(2g, D)); conjugate half for indices >= 2^(D-1)
        uint32_t bi = bit_reverse(2 * g, D);
        const uint32_t half = 1u << (D - 1);
        uint32_t pidx = bi < half ? half_odds_index(D - 1, bi)
                                  : (uint32_t)((0x80000000u - half_odds_index(D - 1, bi - half)) & 0x7fffffffu);
        itw = m31_inv(v_point_from_index(gp, pidx).y);
      } else {
        itw = m31_inv(v_point_from_index(gp, half_odds_index(d, bit_reverse(2 * g, d))).x);
      }
      evals[n_next++] = fri_fold_pair(pair[0], pair[1], itw, alphas[l]);
    }
    (void)n_cur_evals;
    if (wit != n_fri) return stop(0, l, 0);  // witness not fully consumed
    st.n_leaf[l] = n_out;
    n_cur_evals = n_next;
    if (l == 0 && n_layers == 1) return stop(-1, 0, 1);  // assert!(first_layer_columns.is_empty()), after Merkle
  }
  // ---- last layer: folded values against the last-layer polynomial
  {
    const uint32_t llog = D - n_layers;
    uint32_t plog = 0;
    while ((1u << plog) < n_last) plog++;
    uint32_t k = 0, i = 0;
    while (i < nq) {
      const uint32_t pos = queries[i] >> n_layers;
      while (i < nq && (queries[i] >> n_layers) == pos) i++;
      CPoint p = v_point_from_index(gp, half_odds_index(llog, bit_reverse(pos, llog)));
      // LinePoly::eval_at_point: fold over the doublings, coefficients in storage order
      QM31 dbl[21];
      QM31 x = {{p.x, 0, 0, 0}};
      const QM31 one = {{1, 0, 0, 0}};
      for (uint32_t b = 0; b < plog; b++) {
        dbl[b] = x;
        QM31 xx = qm31_mul(x, x);
        x = qm31_sub(qm31_add(xx, xx), one);
      }
      // iterative evaluation of fold(values, doublings): combine adjacent halves bottom-up
      // value(level plog) = coeffs; at each step v[j] = v[2j] ... uses the LAST doubling first
      QM31 acc = {{0, 0, 0, 0}};
      for (uint32_t c = 0; c < n_last; c++) {
        // term = coeff[c] * prod_{b : bit (plog-1-b) of c set} dbl[b]
        QM31 term = v_q(last + 4 * c);
        for (uint32_t b = 0; b < plog; b++)
          if ((c >> (plog - 1 - b)) & 1u) term = qm31_mul(term, dbl[b]);
        acc = qm31_add(acc, term);
      }
      if (!qm31_eq(evals[k], acc)) return stop(0, n_layers - 1, 1);  // LastLayerEvaluationsInvalid
      k++;
    }
  }
}

// Phase B: MerkleVerifier::verify of one layer, in place over its leaf list.
FR_HD int verify_phase_b(const uint32_t *words, uint32_t n_words, uint32_t layer_off, uint32_t d, VNode *nodes,
                         uint32_t n) {
  VReader lr{words, n_words, layer_off, true};
  const uint32_t *root = lr.take(8);
  const uint32_t n_fri = lr.u32();
  lr.take_items(n_fri, 4);
  const uint32_t n_hash = lr.u32();
  const uint32_t *hw = lr.take_items(n_hash, 8);
  const uint32_t n_colw = lr.u32();
  if (!lr.ok || n_colw != 0 || n == 0) return 0;
  for (uint32_t j = 0; j < n; j++) {
    uint32_t h[8];
    merkle_hash_leaf(nodes[j].w[0], nodes[j].w[1], nodes[j].w[2], nodes[j].w[3], h);
    for (int i = 0; i < 8; i++) nodes[j].w[i] = h[i];
  }
  uint32_t used = 0;
  for (int k = (int)d - 1; k >= 0; k--) {
    uint32_t i = 0, wpos = 0;
    while (i < n) {
      const uint32_t node = nodes[i].idx >> 1;
      uint32_t m[16];
      if (nodes[i].idx == 2 * node) {
        for (int t = 0; t < 8; t++) m[t] = nodes[i].w[t];
        i++;
      } else {
        if (used >= n_hash) return 0;  // WitnessTooShort
        for (int t = 0; t < 8; t++) m[t] = hw[8 * used + t];
        used++;
      }
      if (i < n && nodes[i].idx == 2 * node + 1) {
        for (int t = 0; t < 8; t++) m[8 + t] = nodes[i].w[t];
        i++;
      } else {
        if (used >= n_hash) return 0;
        for (int t = 0; t < 8; t++) m[8 + t] = hw[8 * used + t];
        used++;
      }
      uint32_t h[8];
      merkle_hash_node(m, h);
      nodes[wpos].idx = node;  // wpos < i: in place
      for (int t = 0; t < 8; t++) nodes[wpos].w[t] = h[t];
      wpos++;
    }
    n = wpos;
  }
  if (used != n_hash) return 0;  // WitnessTooLong
  if (n != 1) return 0;
  for (int t = 0; t < 8; t++)
    if (nodes[0].w[t] != root[t]) return 0;  // RootMismatch
  return 1;
}

// Reference order of checks over the two phases' outputs: 1 valid, 0 invalid, -1 reference panics.
FR_HD int verify_resolve(const VProofState &st) {
  const uint32_t stop_layer = st.a_kind == 1 ? st.n_layers : st.a_layer;
  for (uint32_t l = 0; l < st.n_layers; l++) {
    if (st.a_kind != 1 && l == stop_layer && !st.a_post) return st.a_kind;
    if (st.n_leaf[l] && !st.merkle_ok[l]) return 0;
    if (st.a_kind != 1 && l == stop_layer && st.a_post) return st.a_kind;
  }
  return st.a_kind == 1 ? 1 : st.a_kind;
}

}  // namespace frieda
