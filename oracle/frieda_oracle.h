/*
 * frieda_oracle.h -- CPU oracle for FRIEDA's commit path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a scalar, single-threaded C restatement of the algorithm that the
 * reference crate (keep-starknet-strange/frieda) runs through
 * stwo-prover 0.1.1 @ git 19d12d700b65b4d9ae45dcca82ee24ea00976542
 * (Cargo.toml:12, Cargo.lock:896-898; source NOT vendored in /root/reference)
 * with `CpuBackend`, `Blake2sMerkleHasher`, `Blake2sMerkleChannel`.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  Nothing under frieda_b200/ links,
 * imports or executes it.
 *
 * Parity status:
 *   - commitment root (packing + LDE + Merkle): PINNED by the reference's golden
 *     vector src/commit.rs:31-37 and reproduced by tests/test_oracle_golden.py.
 *   - packing: PINNED by src/utils.rs:40-66.
 *   - FRI inner-layer roots, alphas, last-layer poly, proof-of-work nonce, query
 *     positions, witness ordering: the reference holds no vectors for these, so the
 *     transcript conventions (channel byte layouts) are restated from the published
 *     stwo algorithm: "parity unpinned" for those details.  The mathematics is
 *     constrained by the prover's own degree assert and the verifier round trip.
 */
#ifndef FRIEDA_ORACLE_H
#define FRIEDA_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t v[4]; } fo_qm31; /* (a + b i) + (c + d i) u, memory order a,b,c,d */

typedef struct {
  uint32_t log_blowup_factor;
  uint32_t log_last_layer_degree_bound;
  uint64_t n_queries;
  uint32_t pow_bits;
} fo_pcs_config;

/* One FriLayerProof (stwo fri.rs): witness evals, Merkle decommitment, commitment. */
typedef struct {
  uint8_t commitment[32];
  uint32_t n_fri_witness;
  fo_qm31 *fri_witness;
  uint32_t n_hash_witness;
  uint8_t *hash_witness; /* n_hash_witness * 32 */
  uint32_t n_column_witness;
  uint32_t *column_witness; /* always empty on this path */
} fo_layer_proof;

/* Mirror of frieda::proof::Proof (src/proof.rs:19-26). */
typedef struct {
  fo_layer_proof first_layer;
  uint32_t n_inner_layers;
  fo_layer_proof *inner_layers;
  uint32_t n_last_layer_poly;
  fo_qm31 *last_layer_poly;
  uint64_t proof_of_work;
  fo_pcs_config pcs_config;
  uint32_t log_size_bound;
  uint32_t n_evaluations;
  fo_qm31 *evaluations;
} fo_proof;

/* Status codes.  The reference panics where these are negative. */
#define FO_OK 0
#define FO_ERR_PANIC (-1) /* a reference assert/unwrap/overflow would fire */
#define FO_ERR_ALLOC (-2)

/* ---- primitives (exposed for unit tests) -------------------------------- */
uint32_t fo_m31_mul(uint32_t a, uint32_t b);
uint32_t fo_m31_inv(uint32_t a);
fo_qm31 fo_qm31_mul(fo_qm31 a, fo_qm31 b);
void fo_circle_point(uint32_t index, uint32_t *x, uint32_t *y); /* G^index, index mod 2^31 */
void fo_blake2s_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t t1,
                         uint32_t f0, uint32_t f1);
void fo_blake2s_256(const uint8_t *data, size_t len, uint8_t out[32]);

/* src/utils.rs:10-19.  Returns the felt count; writes at most cap felts. */
size_t fo_bytes_to_felts(const uint8_t *data, size_t len, uint32_t *out, size_t cap);
/* src/utils.rs:21-33.  Returns poly_log (log2 of each of the 4 coordinate polys). */
uint32_t fo_poly_log(size_t len);
/* stwo CpuBackend::precompute_twiddles(Coset::half_odds(k)): 2^k forward + 2^k inverse. */
int fo_precompute_twiddles(uint32_t k, uint32_t *tw, uint32_t *itw);
/* stwo CpuBackend::evaluate on CircleDomain(half_odds(log_size-1)); in place, values has
 * 2^log_size entries (coefficients zero-extended); tw = tree of half_odds(log_size-1). */
int fo_circle_fft(uint32_t *values, uint32_t log_size, const uint32_t *tw);
/* Direct O(N^2) evaluation per SURVEY A.5's semantic definition; out in bit-reversed order. */
int fo_circle_eval_naive(const uint32_t *coeffs, uint32_t n_coeffs_log, uint32_t log_size,
                         uint32_t *out);

/* ---- the reference API (src/lib.rs:31-43) -------------------------------- */
int fo_commit(const uint8_t *data, size_t len, uint32_t log_blowup, uint8_t root_out[32]);

/* commit_and_generate_proof (src/proof.rs:32-77).  seed may be NULL. */
int fo_prove(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
             uint8_t root_out[32], fo_proof **proof_out);
/* verify_proof (src/proof.rs:79-101).  Returns 1/0, or FO_ERR_PANIC where the reference
 * panics (short `evaluations`, src/proof.rs:166-173). */
int fo_verify(const fo_proof *proof, const uint64_t *seed);
void fo_proof_free(fo_proof *p);
fo_proof *fo_proof_clone(const fo_proof *p);

/* Flat little-endian encoding shared with the product library's serializer, so that
 * two proofs can be compared byte for byte.  Returns bytes needed; writes if cap allows. */
size_t fo_proof_serialize(const fo_proof *p, uint8_t *out, size_t cap);

/* ---- traced prover: keeps every intermediate for parity tests ------------ */
typedef struct fo_trace fo_trace;
/* stop_after_fri != 0: run only the FRI commit phase (no grind / decommit). */
int fo_trace_run(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
                 int stop_after_fri, fo_trace **out);
void fo_trace_free(fo_trace *t);
uint32_t fo_trace_poly_log(const fo_trace *t);
uint32_t fo_trace_n_felts(const fo_trace *t);
const uint32_t *fo_trace_coeffs(const fo_trace *t);               /* 4 * 2^poly_log */
const uint32_t *fo_trace_twiddles(const fo_trace *t, int inverse); /* 2^(D-1) */
uint32_t fo_trace_n_layers(const fo_trace *t); /* 1 + inner layers */
/* layer 0 = circle evaluation (log D); layer i>=1 = i-th line layer (log D-i). */
uint32_t fo_trace_layer_log(const fo_trace *t, uint32_t layer);
const uint32_t *fo_trace_layer_column(const fo_trace *t, uint32_t layer, uint32_t coord);
/* Merkle layer `level` (0 = root) of the tree over `layer`; 32 bytes per node. */
const uint8_t *fo_trace_tree_level(const fo_trace *t, uint32_t layer, uint32_t level);
/* alpha drawn after committing `layer` (used to fold it). */
fo_qm31 fo_trace_alpha(const fo_trace *t, uint32_t layer);
/* last folded evaluation (size 2^(log_last+log_blowup)); one array per coordinate. */
const uint32_t *fo_trace_last_eval(const fo_trace *t, uint32_t *log_out);
const uint32_t *fo_trace_last_eval_col(const fo_trace *t, uint32_t coord);
const uint8_t *fo_trace_digest_after_fri(const fo_trace *t); /* channel digest, 32 B */
uint64_t fo_trace_nonce(const fo_trace *t);
uint32_t fo_trace_n_queries(const fo_trace *t);
const uint32_t *fo_trace_queries(const fo_trace *t);
const fo_proof *fo_trace_proof(const fo_trace *t);
const uint8_t *fo_trace_root(const fo_trace *t);

/* FRI commit phase only, for the CPU baseline: roots of all layers + last poly.
 * roots_out holds (1 + n_inner) * 32 bytes (capacity in layers given). */
int fo_fri_commit(const uint8_t *data, size_t len, const uint64_t *seed, const fo_pcs_config *cfg,
                  uint8_t *roots_out, uint32_t roots_cap, uint32_t *n_layers_out,
                  fo_qm31 *last_poly_out, uint32_t last_poly_cap);

#ifdef __cplusplus
}
#endif
#endif
