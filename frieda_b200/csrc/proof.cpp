// proof.cpp -- ownership, cloning and flat encoding of frieda_proof, the C mirror of
// frieda::proof::Proof (src/proof.rs:19-26: FriProof + proof_of_work + pcs_config +
// log_size_bound + evaluations).  The reference derives serde traits but fixes no wire format;
// the encoding here is this library's own (documented in INTEGRATION.md).
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/frieda_b200.h"

namespace {

void *dup_mem(const void *src, size_t n) {
  void *d = std::malloc(n ? n : 1);
  if (d && n) std::memcpy(d, src, n);
  return d;
}
void layer_free(frieda_layer_proof *l) {
  std::free(l->fri_witness);
  std::free(l->hash_witness);
  std::free(l->column_witness);
}
bool layer_clone(frieda_layer_proof *d, const frieda_layer_proof *s) {
  *d = *s;
  d->fri_witness = (frieda_qm31 *)dup_mem(s->fri_witness, sizeof(frieda_qm31) * s->n_fri_witness);
  d->hash_witness = (uint8_t *)dup_mem(s->hash_witness, 32 * (size_t)s->n_hash_witness);
  d->column_witness = (uint32_t *)dup_mem(s->column_witness, 4 * (size_t)s->n_column_witness);
  return d->fri_witness && d->hash_witness && d->column_witness;
}

struct Writer {
  uint8_t *p;
  size_t cap, n;
  void bytes(const void *src, size_t k) {
    if (p && n + k <= cap) std::memcpy(p + n, src, k);
    n += k;
  }
  void u32(uint32_t v) {
    uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    bytes(b, 4);
  }
  void u64(uint64_t v) {
    u32((uint32_t)v);
    u32((uint32_t)(v >> 32));
  }
  void qm31s(const frieda_qm31 *q, uint32_t k) {
    u32(k);
    for (uint32_t i = 0; i < k; i++)
      for (int j = 0; j < 4; j++) u32(q[i].v[j]);
  }
  void layer(const frieda_layer_proof *l) {
    bytes(l->commitment, 32);
    qm31s(l->fri_witness, l->n_fri_witness);
    u32(l->n_hash_witness);
    bytes(l->hash_witness, 32 * (size_t)l->n_hash_witness);
    u32(l->n_column_witness);
    for (uint32_t i = 0; i < l->n_column_witness; i++) u32(l->column_witness[i]);
  }
};

struct Reader {
  const uint8_t *p;
  size_t len, n;
  bool ok;
  bool bytes(void *dst, size_t k) {
    if (!ok || n + k > len) return ok = false;
    std::memcpy(dst, p + n, k);
    n += k;
    return true;
  }
  uint32_t u32() {
    uint8_t b[4] = {0, 0, 0, 0};
    bytes(b, 4);
    return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
  }
  uint64_t u64() {
    uint64_t lo = u32();
    return lo | ((uint64_t)u32() << 32);
  }
  bool qm31s(frieda_qm31 **out, uint32_t *k) {
    *k = u32();
    if (!ok || (size_t)*k * 16 > len - n) return ok = false;
    *out = (frieda_qm31 *)std::malloc(sizeof(frieda_qm31) * (*k ? *k : 1));
    if (!*out) return ok = false;
    for (uint32_t i = 0; i < *k; i++)
      for (int j = 0; j < 4; j++) (*out)[i].v[j] = u32();
    return ok;
  }
  bool layer(frieda_layer_proof *l) {
    std::memset(l, 0, sizeof *l);
    bytes(l->commitment, 32);
    if (!qm31s(&l->fri_witness, &l->n_fri_witness)) return false;
    l->n_hash_witness = u32();
    if (!ok || (size_t)l->n_hash_witness * 32 > len - n) return ok = false;
    l->hash_witness = (uint8_t *)std::malloc(32 * (size_t)(l->n_hash_witness ? l->n_hash_witness : 1));
    if (!l->hash_witness) return ok = false;
    bytes(l->hash_witness, 32 * (size_t)l->n_hash_witness);
    l->n_column_witness = u32();
    if (!ok || (size_t)l->n_column_witness * 4 > len - n) return ok = false;
    l->column_witness = (uint32_t *)std::malloc(4 * (size_t)(l->n_column_witness ? l->n_column_witness : 1));
    if (!l->column_witness) return ok = false;
    for (uint32_t i = 0; i < l->n_column_witness; i++) l->column_witness[i] = u32();
    return ok;
  }
};

}  // namespace

extern "C" {

void frieda_proof_free(frieda_proof *p) {
  if (!p) return;
  layer_free(&p->first_layer);
  for (uint32_t i = 0; i < p->n_inner_layers; i++) layer_free(&p->inner_layers[i]);
  std::free(p->inner_layers);
  std::free(p->last_layer_poly);
  std::free(p->evaluations);
  std::free(p);
}

frieda_proof *frieda_proof_clone(const frieda_proof *p) {
  if (!p) return nullptr;
  frieda_proof *d = (frieda_proof *)std::calloc(1, sizeof *d);
  if (!d) return nullptr;
  *d = *p;
  d->inner_layers = nullptr;
  d->last_layer_poly = nullptr;
  d->evaluations = nullptr;
  std::memset(&d->first_layer, 0, sizeof d->first_layer);
  d->n_inner_layers = 0;
  bool ok = layer_clone(&d->first_layer, &p->first_layer);
  d->inner_layers = (frieda_layer_proof *)std::calloc(p->n_inner_layers ? p->n_inner_layers : 1, sizeof(frieda_layer_proof));
  ok = ok && d->inner_layers;
  for (uint32_t i = 0; ok && i < p->n_inner_layers; i++) {
    ok = layer_clone(&d->inner_layers[i], &p->inner_layers[i]);
    d->n_inner_layers = i + 1;
  }
  d->last_layer_poly = (frieda_qm31 *)dup_mem(p->last_layer_poly, sizeof(frieda_qm31) * p->n_last_layer_poly);
  d->evaluations = (frieda_qm31 *)dup_mem(p->evaluations, sizeof(frieda_qm31) * p->n_evaluations);
  if (!ok || !d->last_layer_poly || !d->evaluations) {
    frieda_proof_free(d);
    return nullptr;
  }
  return d;
}

size_t frieda_proof_serialize(const frieda_proof *p, uint8_t *out, size_t cap) {
  Writer w{out, cap, 0};
  w.bytes("FRDA", 4);
  w.u32(p->log_size_bound);
  w.u32(p->pcs_config.log_blowup_factor);
  w.u32(p->pcs_config.log_last_layer_degree_bound);
  w.u64(p->pcs_config.n_queries);
  w.u32(p->pcs_config.pow_bits);
  w.u64(p->proof_of_work);
  w.qm31s(p->evaluations, p->n_evaluations);
  w.qm31s(p->last_layer_poly, p->n_last_layer_poly);
  w.u32(1 + p->n_inner_layers);
  w.layer(&p->first_layer);
  for (uint32_t i = 0; i < p->n_inner_layers; i++) w.layer(&p->inner_layers[i]);
  return w.n;
}

// bincode 1.x default configuration (little-endian, fixed-width integers, u64 sequence lengths) of the
// reference's `#[derive(Serialize)] struct Proof` (src/proof.rs:19-26) with stwo's field order:
//   Proof { proof: FriProof { first_layer, inner_layers: Vec<FriLayerProof>, last_layer_poly: LinePoly { coeffs,
//   log_size } }, proof_of_work: u64, pcs_config: PcsConfig { pow_bits, fri_config: FriConfig { log_blowup_factor,
//   log_last_layer_degree_bound, n_queries: usize } }, log_size_bound: u32, evaluations: Vec<QM31> }
//   FriLayerProof { fri_witness: Vec<QM31>, decommitment: MerkleDecommitment { hash_witness: Vec<[u8; 32]>,
//   column_witness: Vec<M31> }, commitment: [u8; 32] }
// UNVALIDATED against the Rust crate (no Rust toolchain in this image).
size_t frieda_proof_serialize_bincode(const frieda_proof *p, uint8_t *out, size_t cap) {
  Writer w{out, cap, 0};
  auto seq_qm31 = [&](const frieda_qm31 *q, uint32_t k) {
    w.u64(k);
    for (uint32_t i = 0; i < k; i++)
      for (int j = 0; j < 4; j++) w.u32(q[i].v[j]);
  };
  auto layer = [&](const frieda_layer_proof *l) {
    seq_qm31(l->fri_witness, l->n_fri_witness);
    w.u64(l->n_hash_witness);
    w.bytes(l->hash_witness, 32 * (size_t)l->n_hash_witness);
    w.u64(l->n_column_witness);
    for (uint32_t i = 0; i < l->n_column_witness; i++) w.u32(l->column_witness[i]);
    w.bytes(l->commitment, 32);
  };
  layer(&p->first_layer);
  w.u64(p->n_inner_layers);
  for (uint32_t i = 0; i < p->n_inner_layers; i++) layer(&p->inner_layers[i]);
  seq_qm31(p->last_layer_poly, p->n_last_layer_poly);
  uint32_t log_size = 0;
  while ((1u << log_size) < p->n_last_layer_poly) log_size++;
  w.u32(log_size);
  w.u64(p->proof_of_work);
  w.u32(p->pcs_config.pow_bits);
  w.u32(p->pcs_config.log_blowup_factor);
  w.u32(p->pcs_config.log_last_layer_degree_bound);
  w.u64(p->pcs_config.n_queries);
  w.u32(p->log_size_bound);
  seq_qm31(p->evaluations, p->n_evaluations);
  return w.n;
}

int frieda_proof_deserialize(const uint8_t *bytes, size_t len, frieda_proof **proof_out) {
  if (!bytes || !proof_out) return FRIEDA_ERR_ARG;
  Reader r{bytes, len, 0, true};
  char magic[4];
  r.bytes(magic, 4);
  if (!r.ok || std::memcmp(magic, "FRDA", 4) != 0) return FRIEDA_ERR_ARG;
  frieda_proof *p = (frieda_proof *)std::calloc(1, sizeof *p);
  if (!p) return FRIEDA_ERR_ALLOC;
  p->log_size_bound = r.u32();
  p->pcs_config.log_blowup_factor = r.u32();
  p->pcs_config.log_last_layer_degree_bound = r.u32();
  p->pcs_config.n_queries = r.u64();
  p->pcs_config.pow_bits = r.u32();
  p->proof_of_work = r.u64();
  r.qm31s(&p->evaluations, &p->n_evaluations);
  r.qm31s(&p->last_layer_poly, &p->n_last_layer_poly);
  uint32_t n_layers = r.u32();
  if (!r.ok || n_layers == 0 || n_layers > 64) {
    frieda_proof_free(p);
    return FRIEDA_ERR_ARG;
  }
  r.layer(&p->first_layer);
  p->inner_layers = (frieda_layer_proof *)std::calloc(n_layers, sizeof(frieda_layer_proof));
  if (!p->inner_layers) {
    frieda_proof_free(p);
    return FRIEDA_ERR_ALLOC;
  }
  for (uint32_t i = 0; r.ok && i + 1 < n_layers; i++) {
    r.layer(&p->inner_layers[i]);
    p->n_inner_layers = i + 1;
  }
  if (!r.ok || r.n != len) {
    frieda_proof_free(p);
    return FRIEDA_ERR_ARG;
  }
  *proof_out = p;
  return FRIEDA_OK;
}

}  // extern "C"
