// split_proof.cpp -- merges the per-rank shares of a split blob's decommitment into the Proof (host only).
// A share (frieda_fri_split_decommit) carries the transcript results every rank agrees on (nonce, queries, layer roots,
// last-layer polynomial) and, in proof order, the values and hashes that rank holds; split_plan.hpp says which rank
// holds what, so merging is a walk over the plan with one cursor per share.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/frieda_b200.h"
#include "split_plan.hpp"

using namespace frieda;

namespace {
struct ShareView {
  SplitShareHeader h;
  const uint32_t *queries, *roots, *last, *qitems, *hitems;
  uint32_t n_q, n_h;
  uint32_t cur_q = 0, cur_h = 0;
};

bool parse_share(const uint8_t *bytes, size_t len, ShareView &v) {
  if (!bytes || (len & 3) || len < sizeof(SplitShareHeader) + 8) return false;
  std::memcpy(&v.h, bytes, sizeof v.h);
  const SplitShareHeader &h = v.h;
  if (h.magic != SPLIT_SHARE_MAGIC || h.n_layers == 0 || h.n_layers > 32 || h.D > 28 || h.log_last > 9 || h.nq_hi ||
      h.n_unique > h.nq_lo || h.nq_lo > 4096 || h.gl > 6 || h.world != (1u << h.gl) || h.rank >= h.world)
    return false;
  const size_t words = len / 4;
  size_t pos = sizeof(SplitShareHeader) / 4;
  const uint32_t *w = reinterpret_cast<const uint32_t *>(bytes);
  auto take = [&](size_t n, const uint32_t *&out) {
    if (n > words - pos) return false;
    out = w + pos;
    pos += n;
    return true;
  };
  const uint32_t *counts;
  if (!take(h.n_unique, v.queries) || !take((size_t)h.n_layers * 8, v.roots) || !take((size_t)4 << h.log_last, v.last) ||
      !take(2, counts))
    return false;
  v.n_q = counts[0];
  v.n_h = counts[1];
  return take((size_t)v.n_q * 4, v.qitems) && take((size_t)v.n_h * 8, v.hitems) && pos == words;
}
}  // namespace

extern "C" {

void frieda_buffer_free(uint8_t *p) { std::free(p); }

int frieda_fri_split_assemble(const uint8_t *const *shares, const size_t *share_lens, uint32_t world,
                              frieda_proof **proof_out) {
  if (!shares || !share_lens || !proof_out || world == 0 || world > 64) return FRIEDA_ERR_ARG;
  *proof_out = nullptr;
  std::vector<ShareView> sv(world);
  for (uint32_t r = 0; r < world; r++) {
    if (((reinterpret_cast<uintptr_t>(shares[r])) & 3) || !parse_share(shares[r], share_lens[r], sv[r])) return FRIEDA_ERR_ARG;
    const SplitShareHeader &a = sv[0].h, &b = sv[r].h;
    // every rank replayed the same transcript: anything else means the shares do not belong together
    if (b.rank != r || b.world != world || b.D != a.D || b.n_layers != a.n_layers || b.n_split != a.n_split ||
        b.nonce_lo != a.nonce_lo || b.nonce_hi != a.nonce_hi || b.n_unique != a.n_unique || b.nq_lo != a.nq_lo ||
        b.pow_bits != a.pow_bits || b.log_blowup != a.log_blowup || b.log_last != a.log_last ||
        std::memcmp(sv[r].queries, sv[0].queries, 4 * (size_t)a.n_unique) != 0 ||
        std::memcmp(sv[r].roots, sv[0].roots, 32 * (size_t)a.n_layers) != 0)
      return FRIEDA_ERR_ARG;
  }
  const SplitShareHeader &h = sv[0].h;
  std::vector<SplitItem> items;
  std::vector<uint32_t> n_fri, n_hash;
  split_plan(SplitShape{h.D, h.n_layers, h.n_split, h.gl}, sv[0].queries, h.n_unique, items, n_fri, n_hash);
  frieda_proof *p = (frieda_proof *)std::calloc(1, sizeof *p);
  if (!p) return FRIEDA_ERR_ALLOC;
  bool ok = true;
  p->pcs_config.log_blowup_factor = h.log_blowup;
  p->pcs_config.log_last_layer_degree_bound = h.log_last;
  p->pcs_config.n_queries = h.nq_lo;
  p->pcs_config.pow_bits = h.pow_bits;
  p->log_size_bound = h.log_size_bound;
  p->proof_of_work = (uint64_t)h.nonce_lo | ((uint64_t)h.nonce_hi << 32);
  p->n_inner_layers = h.n_layers - 1;
  p->inner_layers = (frieda_layer_proof *)std::calloc(h.n_layers, sizeof(frieda_layer_proof));
  p->n_last_layer_poly = 1u << h.log_last;
  p->last_layer_poly = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) << h.log_last);
  p->n_evaluations = h.n_unique;
  p->evaluations = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) * (h.n_unique ? h.n_unique : 1));
  ok = p->inner_layers && p->last_layer_poly && p->evaluations;
  if (ok) std::memcpy(p->last_layer_poly, sv[0].last, sizeof(frieda_qm31) << h.log_last);
  std::vector<frieda_layer_proof *> lp(h.n_layers);
  for (uint32_t l = 0; ok && l < h.n_layers; l++) {
    lp[l] = l == 0 ? &p->first_layer : &p->inner_layers[l - 1];
    std::memcpy(lp[l]->commitment, sv[0].roots + 8 * l, 32);
    lp[l]->fri_witness = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) * (n_fri[l] ? n_fri[l] : 1));
    lp[l]->hash_witness = (uint8_t *)std::malloc(32 * (size_t)(n_hash[l] ? n_hash[l] : 1));
    lp[l]->column_witness = (uint32_t *)std::malloc(4);
    ok = lp[l]->fri_witness && lp[l]->hash_witness && lp[l]->column_witness;
  }
  uint32_t at_eval = 0;
  for (size_t i = 0; ok && i < items.size(); i++) {
    const SplitItem &it = items[i];
    if (it.owner >= world) {
      ok = false;
      break;
    }
    ShareView &s = sv[it.owner];
    if (it.kind == SK_HASH) {
      if (s.cur_h >= s.n_h) { ok = false; break; }
      frieda_layer_proof *l = lp[it.layer];
      std::memcpy(l->hash_witness + 32 * (size_t)l->n_hash_witness++, s.hitems + 8 * (size_t)s.cur_h++, 32);
    } else {
      if (s.cur_q >= s.n_q) { ok = false; break; }
      const uint32_t *src = s.qitems + 4 * (size_t)s.cur_q++;
      if (it.kind == SK_EVAL) {
        std::memcpy(&p->evaluations[at_eval++], src, 16);
      } else {
        frieda_layer_proof *l = lp[it.layer];
        std::memcpy(&l->fri_witness[l->n_fri_witness++], src, 16);
      }
    }
  }
  // every share must be consumed exactly: a leftover item means the ranks disagree about who holds what
  for (uint32_t r = 0; ok && r < world; r++) ok = sv[r].cur_q == sv[r].n_q && sv[r].cur_h == sv[r].n_h;
  if (!ok) {
    frieda_proof_free(p);
    return FRIEDA_ERR_ARG;
  }
  *proof_out = p;
  return FRIEDA_OK;
}

}  // extern "C"
