/*
 * frieda_b200.h -- C ABI of libfrieda_b200.so: FRIEDA's data-parallel commit path on B200.
 *
 * The reference crate (keep-starknet-strange/frieda) has no FFI boundary; its contract is the
 * Rust API at src/lib.rs:31-43 (api::commit / api::generate_proof / api::verify), forwarding to
 * src/commit.rs:11 (commit), src/proof.rs:28,32,79 (generate_proof, commit_and_generate_proof,
 * verify_proof) with the `Proof` struct of src/proof.rs:19-26.  Each entry point below names the
 * reference function it replaces; INTEGRATION.md shows the `extern "C"` block and the shim
 * `src/{commit,proof}.rs` a maintainer would add.
 *
 * Conventions crossing the ABI
 *   - a commitment / Merkle hash is 32 bytes = the 8 BLAKE2s state words little-endian
 *     (`Commitment = [u8; 32]`, src/commit.rs:9);
 *   - a QM31 is 4 x u32 little-endian (a, b, c, d) for (a + b i) + (c + d i) u;
 *   - inputs are borrowed for the call (`&[u8]`), roots are caller-allocated, proofs are
 *     library-allocated and released with frieda_proof_free (mirrors `Proof` being owned);
 *   - return 0 on success, negative FRIEDA_ERR_* otherwise.  The reference panics where
 *     FRIEDA_ERR_PANIC is returned (src/proof.rs:61, stwo asserts); the Rust shim panics on it.
 *   - a context owns one device, its streams, twiddle cache and workspace; calls on distinct
 *     contexts may run concurrently, one context is not internally locked.
 *   - there is NO CPU fallback: every compute entry point fails with FRIEDA_ERR_CUDA when no
 *     CUDA device is usable.  frieda_verify is host-only by design (sub-millisecond).
 */
#ifndef FRIEDA_B200_H
#define FRIEDA_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRIEDA_OK 0
#define FRIEDA_ERR_PANIC (-1)   /* the reference would panic on this input */
#define FRIEDA_ERR_ALLOC (-2)   /* host or device allocation failed */
#define FRIEDA_ERR_CUDA (-3)    /* CUDA runtime error / no device */
#define FRIEDA_ERR_ARG (-4)     /* invalid argument (null pointer, unsupported size) */

typedef struct { uint32_t v[4]; } frieda_qm31;

/* stwo PcsConfig { pow_bits, fri_config: FriConfig { log_blowup_factor,
 * log_last_layer_degree_bound, n_queries } } as built by literal in src/lib.rs:71-78,
 * src/proof.rs:109-116, benches/proof.rs:5-12. */
typedef struct {
  uint32_t log_blowup_factor;
  uint32_t log_last_layer_degree_bound;
  uint64_t n_queries;
  uint32_t pow_bits;
} frieda_pcs_config;

/* stwo FriLayerProof<Blake2sMerkleHasher>: fri_witness, decommitment {hash_witness,
 * column_witness}, commitment. */
typedef struct {
  uint8_t commitment[32];
  uint32_t n_fri_witness;
  frieda_qm31 *fri_witness;
  uint32_t n_hash_witness;
  uint8_t *hash_witness; /* n_hash_witness * 32 bytes */
  uint32_t n_column_witness;
  uint32_t *column_witness; /* always empty on this path */
} frieda_layer_proof;

/* frieda::proof::Proof (src/proof.rs:19-26); `proof: FriProof` is flattened into
 * first_layer / inner_layers / last_layer_poly. */
typedef struct frieda_proof {
  frieda_layer_proof first_layer;
  uint32_t n_inner_layers;
  frieda_layer_proof *inner_layers;
  uint32_t n_last_layer_poly;
  frieda_qm31 *last_layer_poly;
  uint64_t proof_of_work;
  frieda_pcs_config pcs_config;
  uint32_t log_size_bound;
  uint32_t n_evaluations;
  frieda_qm31 *evaluations;
} frieda_proof;

typedef struct frieda_ctx frieda_ctx;

/* ---- context ------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`.  FRIEDA_ERR_CUDA if there is none. */
int frieda_ctx_create(int device, frieda_ctx **out);
void frieda_ctx_destroy(frieda_ctx *ctx);
/* Last error text of this context (or of the failed create when ctx is NULL). */
const char *frieda_last_error(const frieda_ctx *ctx);
/* Caps the device workspace a batched call may use (bytes; 0 = 80% of free memory). */
int frieda_ctx_set_workspace_limit(frieda_ctx *ctx, size_t bytes);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t frieda_ctx_launch_count(const frieda_ctx *ctx);
/* Per-kernel timing: when on, every launch is bracketed by CUDA events on the context's stream.
 * frieda_ctx_profile_read synchronises and writes "name launches total_ms\n" per kernel name
 * (returns the text length, or a negative error); reset != 0 clears the records. */
int frieda_ctx_set_profiling(frieda_ctx *ctx, int on);
long frieda_ctx_profile_read(frieda_ctx *ctx, char *out, size_t cap, int reset);
/* The stream all of this context's kernels are launched on (a cudaStream_t). */
void *frieda_ctx_stream(const frieda_ctx *ctx);
/* The *_device entry points are asynchronous: where the reference would panic on the DATA (stwo's "invalid degree"
 * assert inside FriProver::commit, src/proof.rs:52) they cannot say so when they return.  This call synchronises the
 * stream and returns FRIEDA_ERR_PANIC if any *_device call since the last take hit that assert (FRIEDA_OK otherwise);
 * it clears the pending error.  The host-buffer entry points report it themselves. */
int frieda_ctx_take_error(frieda_ctx *ctx);

/* ---- commit: replaces frieda::commit::commit (src/commit.rs:11-23) -------------------- */
/* Host buffer in, 32-byte root out. */
int frieda_commit(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup,
                  uint8_t root_out[32]);
/* n blobs of blob_len bytes, blob i at blobs + i*blob_stride (host memory, ideally pinned);
 * roots_out = n * 32 bytes (host).  Host<->device copies are overlapped with compute. */
int frieda_commit_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride,
                        size_t n, uint32_t log_blowup, uint8_t *roots_out);
/* Same with DEVICE pointers for input and output (inputs already resident in HBM).
 * Asynchronous on frieda_ctx_stream(); d_roots_out = n * 32 bytes of device memory. */
int frieda_commit_batch_device(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride,
                               size_t n, uint32_t log_blowup, uint8_t *d_roots_out);

/* ---- FRI commit phase: the `FriProver::commit` call inside commit_and_generate_proof
 *      (src/proof.rs:38-57): channel init (+ optional mix_u64(seed)), LDE, first-layer tree,
 *      circle->line fold, inner layers with their Merkle commitments, last-layer polynomial.
 *      Outputs per blob: (1 + n_inner) layer roots (layer 0 == commit() root) and the
 *      2^log_last last-layer coefficients.  seeds may be NULL (no seed for any blob). ------ */
/* Number of inner FRI layers for a blob of blob_len bytes under cfg (or negative error). */
int frieda_fri_n_inner_layers(size_t blob_len, const frieda_pcs_config *cfg);
int frieda_fri_commit_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride, size_t n,
                            const uint64_t *seeds, const frieda_pcs_config *cfg, uint8_t *layer_roots_out,
                            frieda_qm31 *last_poly_out);
int frieda_fri_commit_batch_device(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride,
                                   size_t n, const uint64_t *d_seeds, const frieda_pcs_config *cfg,
                                   uint8_t *d_layer_roots_out, frieda_qm31 *d_last_poly_out);

/* ---- proof generation: replaces commit_and_generate_proof (src/proof.rs:32-77);
 *      generate_proof (src/proof.rs:28-30) is the same call ignoring root_out. ------------- */
int frieda_prove(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                 const frieda_pcs_config *cfg, uint8_t root_out[32], frieda_proof **proof_out);
/* seeds: n values or NULL; roots_out: n*32 bytes or NULL; proofs_out: n pointers. */
int frieda_prove_batch(frieda_ctx *ctx, const uint8_t *blobs, size_t blob_len, size_t blob_stride, size_t n,
                       const uint64_t *seeds, const frieda_pcs_config *cfg, uint8_t *roots_out,
                       frieda_proof **proofs_out);

/* ---- verification: replaces verify_proof (src/proof.rs:79-101).  Host only.
 *      Returns 1 (valid), 0 (invalid) or FRIEDA_ERR_PANIC where the reference panics
 *      (too few `evaluations`, src/proof.rs:166-173). */
int frieda_verify(const frieda_proof *proof, const uint64_t *seed_or_null);
/* The positions `proof.evaluations` belong to (src/proof.rs:62-66 gathers them in ascending query order, the
 * reference leaves them implicit): replays the Fiat-Shamir transcript of verify_proof (src/proof.rs:80-96) and
 * writes the sorted distinct query positions (indices into the bit-reversed evaluation domain of size
 * 2^(log_size_bound + log_blowup_factor)) to positions_out[0 .. min(count, cap)).  Returns the count, 0 when the
 * transcript is rejected before the queries are drawn (layer count, last layer, proof of work), or a negative
 * FRIEDA_ERR_*.  Host only; positions_out may be NULL to query the count. */
long long frieda_proof_query_positions(const frieda_proof *proof, const uint64_t *seed_or_null, uint32_t *positions_out,
                                       size_t cap);

/* GPU batch verification: results[i] = 1 / 0 / FRIEDA_ERR_PANIC per proof, same semantics as
 * frieda_verify.  seeds_or_null: one seed per proof, or NULL when no proof was seeded. */
int frieda_verify_batch(frieda_ctx *ctx, const frieda_proof *const *proofs, size_t n, const uint64_t *seeds_or_null,
                        int *results);
/* Same over SERIALISED proofs (frieda_proof_serialize encoding): proof i occupies bytes[byte_offsets[i] ..
 * byte_offsets[i+1]); `bytes` and all offsets 4-byte aligned, n + 1 non-decreasing offsets, none beyond bytes_len
 * (FRIEDA_ERR_ARG otherwise).  The proof bytes themselves are untrusted: malformed proofs give results[i] = 0.
 * Limits of the batch verifier (both entry points; frieda_verify has neither): n_queries <= 4096 and
 * log_last_layer_degree_bound <= 20 -- a proof beyond them makes the CALL fail with FRIEDA_ERR_ARG rather than
 * being reported as invalid. */
int frieda_verify_batch_bytes(frieda_ctx *ctx, const uint8_t *bytes, size_t bytes_len, const uint64_t *byte_offsets,
                              size_t n, const uint64_t *seeds_or_null, int *results);
/* The batch verifier's core run on the host for ONE proof (self-check of the kernels' logic on CPU). */
int frieda_verify_core_host(const frieda_proof *proof, const uint64_t *seed_or_null);
/* Same over ONE serialised, untrusted proof (len a multiple of 4; `bytes` 4-byte aligned): exactly the parser and
 * the two phases the kernels of frieda_verify_batch_bytes run, on the CPU.  1 / 0 / FRIEDA_ERR_PANIC, or
 * FRIEDA_ERR_ARG beyond the batch verifier's capacity (see frieda_verify_batch_bytes). */
int frieda_verify_core_host_bytes(const uint8_t *bytes, size_t len, const uint64_t *seed_or_null);

/* ---- proof objects --------------------------------------------------------------------- */
void frieda_proof_free(frieda_proof *proof);
frieda_proof *frieda_proof_clone(const frieda_proof *proof);
/* Flat little-endian encoding (the reference picks no wire format).
 * Returns the byte count; writes only if cap is large enough. */
size_t frieda_proof_serialize(const frieda_proof *proof, uint8_t *out, size_t cap);
int frieda_proof_deserialize(const uint8_t *bytes, size_t len, frieda_proof **proof_out);
/* bincode-1.x layout of the reference's serde-derived `Proof` (src/proof.rs:19-26), for interchange with
 * the Rust crate.  Unvalidated: no Rust toolchain exists in the build image. */
size_t frieda_proof_serialize_bincode(const frieda_proof *proof, uint8_t *out, size_t cap);

/* ---- one oversized blob split across GPUs (BASELINE config 5) -------------------------
 * Rank `rank` of `world` (a power of two) computes the LDE of its contiguous bit-reversed
 * index range [rank*N/world, (rank+1)*N/world) from the whole input and the root of the
 * Merkle subtree over it.  d_subroot_out: 32 bytes of device memory (the all-gather input). */
int frieda_commit_split_local(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup,
                              uint32_t rank, uint32_t world, uint8_t *d_subroot_out);
/* Same with the whole blob already resident in device memory (e.g. each rank uploaded 1/world of it
 * over its own PCIe link and the slices were all-gathered over NVLink). */
int frieda_commit_split_local_device(frieda_ctx *ctx, const uint8_t *d_data, size_t len, uint32_t log_blowup,
                                     uint32_t rank, uint32_t world, uint8_t *d_subroot_out);
/* Same over peer-mapped memory (NVLink / NVSwitch), with no all-gather of the input: slice r of the blob
 * (bytes [r * slice_len, (r + 1) * slice_len), slice_len a multiple of 16) lies at peer_slices[r], a
 * 16-byte aligned device pointer that is valid on this context's device (this GPU's own slice, or a peer's
 * symmetric / IPC-mapped buffer).  The packing kernel reads the slices where they lie.  ASYNCHRONOUS on the context's
 * stream: no host synchronisation, so the caller can order it between stream-ordered barriers (uploads
 * done before, roots visible after).  world <= 64. */
int frieda_commit_split_local_peers(frieda_ctx *ctx, const uint8_t *const *peer_slices, uint32_t world,
                                    size_t slice_len, size_t len, uint32_t log_blowup, uint32_t rank,
                                    uint8_t *d_subroot_out);
/* The whole split commit of rank `rank` as ONE call with the exchanges done by the library's own kernels over
 * peer-mapped memory (no NCCL on the data path): upload this rank's slice (host bytes
 * [rank * slice_len, ...)) into peer_slices[rank] -> barrier -> pack from all slices in place -> LDE + Merkle
 * subtree -> root into peer_roots[rank] -> barrier -> combine reading peer_roots[*] in place -> root_out (host).
 * peer_flags[r]: 512 u32 of zero-initialised peer-mapped memory per rank (signal words: two barriers and the
 * per-part "uploaded" flags -- slices of 1 MiB and more are uploaded in four parts on a second stream while the packing
 * kernel already reads the parts that have arrived);
 * data == NULL: this rank's slice already lies in peer_slices[rank] (inputs resident in HBM); nothing is uploaded.
 * epoch: a counter that every rank increases by one per call, starting at 1.  One host synchronisation.
 * Returns FRIEDA_ERR_CUDA with "peer barrier timed out" if a peer does not arrive within ~20 s. */
int frieda_commit_split_peers(frieda_ctx *ctx, const uint8_t *data, size_t len, uint32_t log_blowup, uint32_t rank,
                              uint32_t world, uint8_t *const *peer_slices, size_t slice_len,
                              uint8_t *const *peer_roots, uint32_t *const *peer_flags, uint32_t epoch,
                              uint8_t root_out[32]);
/* Hashes the top log2(world) Merkle levels over the gathered subtree roots
 * (d_subroots = world * 32 bytes of device memory, rank order) into root_out (host). */
int frieda_merkle_combine(frieda_ctx *ctx, const uint8_t *d_subroots, uint32_t world, uint8_t root_out[32]);
/* Same with subtree root r read in place from peer_roots[r] (32 bytes each, peer-mapped device pointers):
 * the exchange step of the split commit as loads over NVLink inside the combine kernel. */
int frieda_merkle_combine_peers(frieda_ctx *ctx, const uint8_t *const *peer_roots, uint32_t world,
                                uint8_t root_out[32]);

/* ---- FRI commit phase of ONE oversized blob split across GPUs (FriProver::commit of src/proof.rs:52-57
 *      for a blob that exceeds one GPU).  Rank r of world = 2^g owns the bit-reversed index range
 *      [r N_l / world, (r+1) N_l / world) of every layer: fold pairs are rank-local, a layer costs one all-gather of
 *      `world` 32-byte subtree roots, and every rank hashes the top g levels and runs mix_root / draw_felt itself.
 *      When a rank's share of a layer is down to 2^10 points the layer is all-gathered and the rest runs on every
 *      rank with the single-GPU kernels.  The exchanges are the CALLER's (NCCL all-gather over NVLink, or any other
 *      transport); every call but `finish` is asynchronous on frieda_ctx_stream().  Call order on every rank:
 *        begin; for l in 0 .. n_split-1: { layer(l); all-gather roots; combine(l) }; handoff; all-gather columns; finish
 *      No other call may use the context in between (the state lives in its workspace).  Results are bit-identical
 *      to frieda_fri_commit_batch on one GPU. ------------------------------------------------------------------- */
/* keep_trees != 0: every Merkle tree is kept whole (rank-local subtrees, the replicated top levels, the unsplit
 * layers) because a proof follows (frieda_fri_split_decommit). */
int frieda_fri_split_begin(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                           const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                           uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out);
/* Same with the whole blob already in device memory. */
int frieda_fri_split_begin_device(frieda_ctx *ctx, const uint8_t *d_data, size_t len, const uint64_t *seed_or_null,
                                  const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                                  uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out);
/* Same with the blob in `world` peer-mapped slices of slice_len bytes (a multiple of 16; the symmetric buffers of
 * frieda_commit_split_peers): this rank uploads slice `rank` of `data` (the whole blob in host memory; NULL = the slices
 * are already resident) over its own PCIe link, a barrier kernel (flag channel 7, `epoch`) orders the uploads, and the
 * packing kernel reads every slice in place over NVLink.  Asynchronous. */
int frieda_fri_split_begin_peers(frieda_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *seed_or_null,
                                 const frieda_pcs_config *cfg, uint32_t rank, uint32_t world, int keep_trees,
                                 uint8_t *const *peer_slices, size_t slice_len, uint32_t *const *peer_flags, uint32_t epoch,
                                 uint32_t *n_split_layers_out, uint32_t *n_layers_out, uint32_t *handoff_log_out);
/* Layer `layer` on this rank's range (fold of the previous layer fused into the leaf hashing) -> its subtree root,
 * 32 bytes of device memory (the all-gather's input). */
int frieda_fri_split_layer(frieda_ctx *ctx, uint32_t layer, uint8_t *d_subroot_out);
/* d_subroots: world * 32 bytes, rank order, device.  Top levels, root of the layer, mix_root, draw alpha. */
int frieda_fri_split_combine(frieda_ctx *ctx, uint32_t layer, const uint8_t *d_subroots);
/* ALL split layers in one call, the per-layer exchange of subtree roots done by the library's own kernels over
 * peer-mapped memory (NVLink / NVSwitch) -- replaces the per-layer {frieda_fri_split_layer, all-gather,
 * frieda_fri_split_combine} sequence.  peer_roots[r]: base of rank r's symmetric roots area (>= 64 slots of 32 bytes,
 * mapped on this GPU); peer_flags[r]: rank r's flag array as in frieda_commit_split_peers (channel 6 is used);
 * epoch: the caller's call counter over these flag arrays (same value on every rank, increasing).  Asynchronous; a
 * peer that never arrives is reported by frieda_fri_split_finish (FRIEDA_ERR_CUDA after ~20 s). */
int frieda_fri_split_layers_peers(frieda_ctx *ctx, uint8_t *const *peer_roots, uint32_t *const *peer_flags,
                                  uint32_t epoch);
/* Folds the last split layer into this rank's share of the next one, written to d_cols_local_out: 4 columns x
 * 2^handoff_log u32 of the caller's device memory (handoff_log from `begin`) -- the input of the all-gather of columns. */
int frieda_fri_split_handoff(frieda_ctx *ctx, uint32_t *d_cols_local_out);
/* d_cols_all: world x (4 x 2^handoff_log u32), rank order, device.  layer_roots_out: n_layers * 32 bytes,
 * last_poly_out: 2^log_last_layer_degree_bound QM31 (host).  Synchronises; FRIEDA_ERR_PANIC on "invalid degree". */
int frieda_fri_split_finish(frieda_ctx *ctx, const uint32_t *d_cols_all, uint8_t *layer_roots_out,
                            frieda_qm31 *last_poly_out);
/* Proof generation for the split blob: the rest of commit_and_generate_proof (src/proof.rs:58-66) after a commit
 * begun with keep_trees.  Every rank calls frieda_fri_split_decommit after `finish`: proof of work and query
 * sampling run on every rank (same channel, same nonce, same positions), then the rank gathers ITS share of the
 * decommitment -- the evaluations, sibling values and tree nodes it holds (owner-serves-path: a rank serves what lies
 * in its index range / subtree, rank 0 also the replicated top levels and the unsplit layers).  *share_out is a
 * malloc'ed, relocatable byte string (release with frieda_buffer_free).  frieda_fri_split_assemble (host only, any
 * rank or any other machine) merges the `world` shares, given in rank order, into the Proof -- byte-identical to the
 * one frieda_prove returns for the same blob on one GPU.  FRIEDA_ERR_ARG if the shares do not belong together. */
int frieda_fri_split_decommit(frieda_ctx *ctx, uint8_t **share_out, size_t *share_len_out);
int frieda_fri_split_assemble(const uint8_t *const *shares, const size_t *share_lens, uint32_t world,
                              frieda_proof **proof_out);
void frieda_buffer_free(uint8_t *buffer);

/* ---- erasure recovery (the reference's README.md:56-70 promises sampling/recovery, its code has none) --
 * The committed evaluation is a Reed-Solomon codeword of rate 2^-log_blowup made of 2^log_blowup coset blocks; ANY ONE
 * whole block determines the data.  block_evals (host): 4 columns x 2^poly_log u32 = entries
 * [block * 2^poly_log, (block + 1) * 2^poly_log) of each evaluation column (bit-reversed domain order, as produced
 * by the commit path).  Writes the original `len` bytes to data_out; FRIEDA_ERR_ARG if the block is not an encoding
 * of `len` bytes.  Any size the commit path accepts (inverse circle FFT: 2^15-point chunks in shared memory, one sweep
 * per layer above that).
 * NOT built: recovery from an arbitrary set of sampled POINTS (a general erasure decoder); a sampler has to collect at
 * least one whole coset block. */
int frieda_decode_block(frieda_ctx *ctx, const uint32_t *block_evals, size_t len, uint32_t log_blowup, uint32_t block,
                        uint8_t *data_out);
/* Several blocks of the same evaluation, block k of the list (index block_ids[k]) at block_evals + k * 4 * 2^poly_log:
 * decodes from the first one that is an encoding of `len` bytes (corrupted blocks are skipped); *used_out (may be
 * NULL) receives its position in the list.  FRIEDA_ERR_ARG when none decodes. */
int frieda_decode_blocks(frieda_ctx *ctx, const uint32_t *block_evals, const uint32_t *block_ids, size_t n_blocks,
                         size_t len, uint32_t log_blowup, uint8_t *data_out, uint32_t *used_out);

/* ---- standalone passes (bench.py's per-pass roofline; tests).  Device pointers. --------
 * LDE: d_coeffs = n * 4 * 2^poly_log u32 -> d_evals = n * 4 * 2^(poly_log+log_blowup) u32. */
int frieda_pass_pack(frieda_ctx *ctx, const uint8_t *d_blobs, size_t blob_len, size_t blob_stride, size_t n,
                     uint32_t *d_coeffs);
int frieda_pass_lde(frieda_ctx *ctx, const uint32_t *d_coeffs, uint32_t poly_log, uint32_t log_blowup, size_t n,
                    uint32_t n_felts, uint32_t *d_evals);
/* Merkle tree over 4 columns of 2^log each: d_tree = n * 2^(log+1) * 32 bytes, node (level k,
 * index i) at slot 2^k + i (slot 0 unused); pass d_tree = NULL to keep only roots.
 * d_roots = n * 32 bytes. */
int frieda_pass_merkle(frieda_ctx *ctx, const uint32_t *d_cols, uint32_t log, size_t n, uint8_t *d_tree,
                       uint8_t *d_roots);
/* FRI folds without hashing: circle (is_circle=1; src log = log) or line; d_alpha = n QM31. */
int frieda_pass_fold(frieda_ctx *ctx, const uint32_t *d_src, uint32_t log, int is_circle, size_t n,
                     const frieda_qm31 *d_alpha, uint32_t *d_dst);
/* Twiddle tree of Coset::half_odds(k): copies 2^k forward and 2^k inverse twiddles to host. */
int frieda_twiddles(frieda_ctx *ctx, uint32_t k, uint32_t *tw_out, uint32_t *itw_out);

/* ---- introspection of the last frieda_prove* / frieda_fri_commit* wave (tests only) ----
 * what: 0 coefficients (4*2^poly_log u32), 1 layer columns (layer, 4*2^log u32),
 *       2 tree level (layer, level; 2^level*32 bytes), 3 alpha (layer; 16 bytes),
 *       4 channel digest after the FRI commit phase (32 bytes), 5 nonce (8 bytes),
 *       6 query positions (u32 each).  Returns bytes written or negative error. */
long frieda_debug_fetch(frieda_ctx *ctx, int what, size_t blob, uint32_t layer, uint32_t level, void *out,
                        size_t cap);
/* Keep full per-layer state of the next wave resident for frieda_debug_fetch (default off). */
int frieda_ctx_set_debug_keep(frieda_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif
