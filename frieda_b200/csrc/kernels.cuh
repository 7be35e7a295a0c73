// kernels.cuh -- launcher declarations shared by the CUDA translation units and ctx.cu.
//
// HBM layouts (all per blob, blobs concatenated with the given strides):
//   coefficients : 4 columns x 2^poly_log u32           (polynomial_from_bytes, src/utils.rs:21-33)
//   layer columns: 4 columns x 2^log u32, SoA by coordinate, bit-reversed domain order
//                  (SecureEvaluation / LineEvaluation columns, src/proof.rs:48-52)
//   Merkle tree  : heap order, node (level k, index i) in 32-byte slot 2^k + i (slot 0 unused);
//                  truncated trees keep only slots below 2^(top_level+1)
//   channel      : one frieda::Channel (digest words + n_sent) per blob
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "blake2s.cuh"
#include "m31.cuh"

namespace frieda {

// Twiddle tree of Coset::half_odds(K), anchored from its end: the block of length s
// (line layer with s twiddles) starts at tw + (2^K - 2 s).  One table serves every
// smaller domain (half_odds(k).double() == half_odds(k-1)).
struct TwiddleTable {
  const uint32_t *tw;   // forward, 2^K
  const uint32_t *itw;  // element-wise inverses
  const uint32_t *tw2;  // forward, doubled (2t < 2^32): the operand form of m31_mul_t2
  uint32_t K;
  FR_HD const uint32_t *blk(uint32_t s) const { return tw + (((size_t)1 << K) - 2 * (size_t)s); }
  FR_HD const uint32_t *blk2(uint32_t s) const { return tw2 + (((size_t)1 << K) - 2 * (size_t)s); }
  FR_HD const uint32_t *iblk(uint32_t s) const { return itw + (((size_t)1 << K) - 2 * (size_t)s); }
};

struct GenPowers {
  CPoint g[31];  // G^(2^j)
};

enum MerkleSrc : int { SRC_COLS = 0, SRC_FOLD_CIRCLE = 1, SRC_FOLD_LINE = 2, SRC_NODES = 3 };

struct MerkleBottomParams {
  const uint32_t *src_cols;  // SRC_COLS: this layer; SRC_FOLD_*: previous layer (2^(log+1) per column)
  uint32_t *dst_cols;        // SRC_FOLD_*: this layer's columns (written)
  uint8_t *tree;             // this layer's tree
  const QM31 *alpha;         // SRC_FOLD_*: alpha per blob
  const uint32_t *itw_blk;   // SRC_FOLD_*: inverse twiddle block for the fold
  size_t src_stride;         // u32 elements between blobs in src_cols
  size_t dst_stride;         // u32 elements between blobs in dst_cols
  size_t tree_stride;        // 32-byte slots between blobs in tree
  size_t alpha_stride;       // QM31 elements between blobs in alpha
  uint32_t log;              // leaf count of this pass = 2^log (SRC_NODES: node count at src_level)
  uint32_t chunk_log;        // leaves handled per CTA
  uint32_t levels;           // levels reduced inside the CTA (<= chunk_log)
  uint32_t src_level;        // SRC_NODES: tree level the 2^log input nodes live on; others: == log
  int write_all;             // write every level (leaves included) to the tree, not only the tops
  int latency;               // latency form (merkle.cu): 128-thread CTAs, chunks of <= 256 leaves
  uint32_t one;              // runtime 1 for the IMAD adds of the compression (blake2s.cuh); set by the launcher
};

cudaError_t launch_pack(cudaStream_t st, const uint8_t *blobs, size_t len, size_t stride, size_t n_blobs,
                        uint32_t n_felts, uint32_t poly_log, uint32_t *coef);
// Packing of ONE blob whose bytes lie in `world` slices of slice_len bytes (a multiple of 16) behind
// separate device pointers (peer-mapped memory of the other GPUs): read in place over NVLink.
constexpr uint32_t MAX_PEERS = 64;
struct PeerPtrs {
  const uint8_t *p[MAX_PEERS];
};
// wait_flags != nullptr (this rank's flag array): the slices are uploaded in parts of part_len bytes while this kernel
// runs; part k of slice s may be read once flag (2 + k, s) >= epoch (launch_peer_signal by its owner).
cudaError_t launch_pack_peers(cudaStream_t st, const PeerPtrs &slices, uint32_t world, uint32_t rank, size_t slice_len,
                              size_t len, uint32_t n_felts, uint32_t poly_log, uint32_t *coef,
                              const uint32_t *wait_flags = nullptr, uint32_t part_len = 0, uint32_t epoch = 0,
                              int *timeout_flag = nullptr);
// tree slot (world + r) <- 32 bytes at roots.p[r]  (the leaves of the top tree of a split commit)
cudaError_t launch_gather_roots(cudaStream_t st, const PeerPtrs &roots, uint32_t world, uint8_t *tree);
// Barrier between the GPUs of a split commit, on the stream: every rank stores `epoch` into word
// (channel * 64 + rank) of every peer's flag array and waits until all `world` words of its own array reach
// it.  Stream order + the system-scope release/acquire make everything this rank wrote before the barrier
// visible to kernels its peers launch after theirs.  *timeout_flag is set if a peer does not arrive in ~20 s.
struct PeerFlags {
  uint32_t *p[MAX_PEERS];
};
// flag (channel, rank) of an array = word channel * MAX_PEERS + rank: channels 0 / 1 = the two barriers, 2 .. 5 = "part
// k of rank's slice is uploaded"; an array holds PEER_FLAG_CHANNELS * MAX_PEERS = 512 words
constexpr uint32_t PEER_FLAG_CHANNELS = 8, PEER_UPLOAD_PARTS = 4;
cudaError_t launch_peer_signal(cudaStream_t st, const PeerFlags &flags, uint32_t world, uint32_t rank, uint32_t channel,
                               uint32_t epoch);
cudaError_t launch_peer_barrier(cudaStream_t st, const PeerFlags &flags, uint32_t world, uint32_t rank, uint32_t channel,
                                uint32_t epoch, int *timeout_flag);
cudaError_t launch_twiddles(cudaStream_t st, const GenPowers &gp, uint32_t K, uint32_t *tw, uint32_t *itw,
                            uint32_t *tw2);
// Owned slice of the evaluation domain (bit-reversed order): [lo, lo + 2^log); lo is a
// multiple of 2^log.  The evaluation buffer then holds 4 x 2^log words per blob.
struct LdeRange {
  size_t lo;
  uint32_t log;
};
// Circle-FFT low-degree extension of n_blobs x 4 coefficient columns (range == nullptr: all).
cudaError_t launch_lde(cudaStream_t st, const uint32_t *coef, uint32_t *eval, uint32_t poly_log, uint32_t log_blowup,
                       size_t n_blobs, uint32_t n_felts, const TwiddleTable &tt, CPoint half_initial,
                       const LdeRange *range = nullptr);
// Erasure recovery: one coset block (4 columns x 2^p evaluations, block index hb) -> coefficients -> bytes.
cudaError_t launch_decode_block(cudaStream_t st, const uint32_t *block_evals, uint32_t *coef, uint32_t p,
                                uint32_t beta, uint32_t hb, const TwiddleTable &tt, size_t len, uint32_t n_felts,
                                uint8_t *out, int *flag);
cudaError_t launch_merkle_bottom(cudaStream_t st, int src, const MerkleBottomParams &p, size_t n_blobs);
// One CTA per blob: reduce 2^top_log nodes at tree level top_log to the root; optionally
// mix_root + draw the folding alpha on the blob's channel.
cudaError_t launch_merkle_top(cudaStream_t st, uint8_t *tree, size_t tree_stride, uint32_t top_log, int write_all,
                              uint8_t *roots, size_t roots_stride, Channel *chan, QM31 *alpha, size_t alpha_stride,
                              size_t n_blobs);
cudaError_t launch_channel_init(cudaStream_t st, Channel *chan, const uint64_t *seeds, size_t n_blobs);
cudaError_t launch_fold(cudaStream_t st, const uint32_t *src, size_t src_stride, uint32_t src_log, int is_circle,
                        const QM31 *alpha, size_t alpha_stride, const TwiddleTable &tt, uint32_t *dst,
                        size_t dst_stride, size_t n_blobs);

cudaError_t launch_fold_range(cudaStream_t st, const uint32_t *src_local, uint32_t src_log, uint32_t local_src_log,
                              size_t lo, int is_circle, const QM31 *alpha, const TwiddleTable &tt, uint32_t *dst_local);

struct TailParams {
  // layer `start_layer` (log = start_log) is resident in cols[start_layer]; the tail commits it,
  // folds on, and finishes the FRI commit phase inside one CTA per blob.
  uint32_t *cols[32];   // per layer: columns base (blob stride cols_stride[layer])
  size_t cols_stride[32];
  uint8_t *tree[32];    // per layer tree base
  size_t tree_stride[32];
  int write_all;
  uint32_t start_layer, start_log;
  uint32_t last_log;    // log size of the last (uncommitted) evaluation = log_last + log_blowup
  uint32_t log_last;    // log of the last-layer degree bound
  uint32_t inv_last_n;  // (2^last_log)^-1 mod P
  uint8_t *roots;       // [blob][n_layers][32]
  size_t roots_stride;  // bytes between blobs
  Channel *chan;
  QM31 *alpha;          // [blob][layer]
  size_t alpha_stride;
  QM31 *last_poly;      // [blob][2^log_last]
  int *error_flag;      // set to 1 on "invalid degree"
  TwiddleTable tt;
  uint32_t one;         // runtime 1 (blake2s.cuh); set by the launcher
};
constexpr uint32_t TAIL_LOG = 9;      // layers with log <= TAIL_LOG are finished by the tail kernel
constexpr uint32_t TAIL_LAST_MAX = 9; // largest supported log_last + log_blowup
cudaError_t launch_tail(cudaStream_t st, const TailParams &p, size_t n_blobs);

// Proof of work: best[b] = min nonce in [0, limit) whose raw-compress mix has >= pow_bits trailing
// zeros (atomicMin; initialise to ~0ull).  The warps of up to 128 CTAs per blob (more for a handful of blobs) take
// nonce chunks of their blob in increasing order from next[b] (zeroed here).
cudaError_t launch_grind(cudaStream_t st, const Channel *chan, uint32_t pow_bits, uint64_t limit,
                         unsigned long long *best, unsigned long long *next, size_t n_blobs);
// mix_u64(nonce) then Queries::generate: sorted unique positions per blob.
cudaError_t launch_queries(cudaStream_t st, Channel *chan, const unsigned long long *nonce, uint32_t log_domain,
                           uint32_t n_queries, uint32_t *queries, uint32_t *n_unique, size_t n_blobs);

struct DecommitParams {
  const uint32_t *queries;  // [blob][n_queries] sorted unique
  const uint32_t *n_unique; // [blob]
  uint32_t n_queries;       // stride of queries
  uint32_t n_layers;        // 1 + inner
  uint32_t D;               // log size of layer 0
  const uint32_t *cols[32];
  size_t cols_stride[32];
  const uint8_t *tree[32];
  size_t tree_stride[32];
  uint32_t *counts;         // [blob][n_layers][2] = (n_fri_witness, n_hash_witness)
  uint32_t *lvl;            // [blob][n_layers][lvl_stride]: per-walk counts, then prefixes inside the layer
  uint32_t lvl_stride;      // >= D (walk 0 = fri witness, walk j = tree level d-1-j)
  unsigned long long *offsets; // [blob][n_layers][2] element offsets into fri_out / hash_out
  QM31 *fri_out;
  uint8_t *hash_out;
  QM31 *evals_out;          // [blob][n_queries]
};
cudaError_t launch_decommit_count(cudaStream_t st, const DecommitParams &p, size_t n_blobs);
// Exclusive scan of counts into offsets; totals[0] = fri elements, totals[1] = hashes.
cudaError_t launch_decommit_scan(cudaStream_t st, const DecommitParams &p, size_t n_blobs, unsigned long long *totals);
cudaError_t launch_decommit_write(cudaStream_t st, const DecommitParams &p, size_t n_blobs);

// Gather by address (split-blob decommitment): a QM31 from four columns `stride` words apart, a 32-byte tree node.
struct GatherQ {
  const uint32_t *p;
  uint32_t stride;
};
struct GatherH {
  const uint4 *p;
};
cudaError_t launch_gather_items(cudaStream_t st, const GatherQ *dq, uint32_t n_q, const GatherH *dh, uint32_t n_h,
                                QM31 *out_q, uint8_t *out_h);

}  // namespace frieda
