// verify_batch.cu -- GPU batch verification (many proofs x a few thousand hashes each,
// the light-client fan-out case) of frieda::proof::verify_proof (src/proof.rs:79-101), plus a host
// self-check that runs the same core on the CPU.  frieda_verify (csrc/verify.cpp) stays the
// single-proof host entry point.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/frieda_b200.h"
#include "ctx_internal.h"
#include "host_math.hpp"
#include "verify_core.cuh"

using namespace frieda;

namespace {

__global__ void __launch_bounds__(64) verify_phase_a_kernel(const uint32_t *__restrict__ blob,
                                                            const unsigned long long *__restrict__ offs,
                                                            const unsigned long long *__restrict__ seeds,
                                                            const __grid_constant__ VGen gp, VProofState *st,
                                                            VNode *leaves, uint32_t n_layers_max, uint32_t max_pos,
                                                            uint32_t *queries, uint32_t max_q, QM31 *evals,
                                                            QM31 *alphas, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t seed = seeds ? seeds[i] : 0;
  verify_phase_a(blob + offs[i], (uint32_t)(offs[i + 1] - offs[i]), seeds ? &seed : nullptr, gp, st[i],
                 leaves + i * (size_t)n_layers_max * max_pos, max_pos, queries + i * (size_t)max_q, max_q,
                 evals + i * (size_t)max_q, alphas + i * (size_t)V_MAX_LAYERS, n_layers_max);
}

__global__ void __launch_bounds__(64) verify_phase_b_kernel(const uint32_t *__restrict__ blob,
                                                            const unsigned long long *__restrict__ offs,
                                                            VProofState *st, VNode *leaves, uint32_t n_layers_max,
                                                            uint32_t max_pos, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n_layers_max) return;
  size_t i = t / n_layers_max;
  uint32_t l = (uint32_t)(t % n_layers_max);
  VProofState &s = st[i];
  if (l >= s.n_layers || s.n_leaf[l] == 0) return;
  s.merkle_ok[l] = verify_phase_b(blob + offs[i], (uint32_t)(offs[i + 1] - offs[i]), s.layer_off[l], s.D - l,
                                  leaves + (i * (size_t)n_layers_max + l) * max_pos, s.n_leaf[l]);
}

VGen make_gen() {
  VGen g;
  CPoint cur = {host::GEN_X, host::GEN_Y};
  for (int j = 0; j < 31; j++) {
    g.g[j] = cur;
    cur = cpoint_add(cur, cur);
  }
  return g;
}

}  // namespace

extern "C" {

// Host self-check of the batch verifier's core (same code path as the kernels, run on the CPU).
int frieda_verify_core_host(const frieda_proof *proof, const uint64_t *seed_or_null) {
  if (!proof) return FRIEDA_ERR_ARG;
  size_t bytes = frieda_proof_serialize(proof, nullptr, 0);
  std::vector<uint32_t> words((bytes + 3) / 4);
  frieda_proof_serialize(proof, reinterpret_cast<uint8_t *>(words.data()), bytes);
  const uint32_t n_layers = 1 + proof->n_inner_layers;
  if (n_layers > V_MAX_LAYERS || proof->pcs_config.n_queries > (1u << 20)) return FRIEDA_ERR_ARG;
  const uint32_t max_q = (uint32_t)proof->pcs_config.n_queries ? (uint32_t)proof->pcs_config.n_queries : 1;
  const uint32_t max_pos = 2 * max_q;
  std::vector<VNode> leaves((size_t)n_layers * max_pos);
  std::vector<uint32_t> queries(max_q);
  std::vector<QM31> evals(max_q), alphas(V_MAX_LAYERS);
  VProofState st;
  VGen gp = make_gen();
  verify_phase_a(words.data(), (uint32_t)words.size(), seed_or_null, gp, st, leaves.data(), max_pos, queries.data(),
                 max_q, evals.data(), alphas.data());
  for (uint32_t l = 0; l < st.n_layers; l++)
    if (st.n_leaf[l])
      st.merkle_ok[l] = verify_phase_b(words.data(), (uint32_t)words.size(), st.layer_off[l], st.D - l,
                                       leaves.data() + (size_t)l * max_pos, st.n_leaf[l]);
  return verify_resolve(st);
}

// The same core over raw, untrusted words: what verify_phase_{a,b}_kernel run per proof, on the CPU.
int frieda_verify_core_host_bytes(const uint8_t *bytes, size_t len, const uint64_t *seed_or_null) {
  if (!bytes || (len & 3) || (reinterpret_cast<uintptr_t>(bytes) & 3) || len > 0xffffffffull * 4) return FRIEDA_ERR_ARG;
  const uint32_t *words = reinterpret_cast<const uint32_t *>(bytes);
  const uint32_t n_words = (uint32_t)(len / 4);
  uint32_t max_q = 1, n_layers = V_MAX_LAYERS;
  if (n_words > 10 && words[0] == 0x41445246u) {
    if (words[5] == 0 && words[4] > 4096) return FRIEDA_ERR_ARG;
    if (words[3] > 20 && words[3] < 32) return FRIEDA_ERR_ARG;
    max_q = words[5] ? 1u : (words[4] ? words[4] : 1u);
  }
  const uint32_t max_pos = 2 * max_q;
  std::vector<VNode> leaves((size_t)n_layers * max_pos);
  std::vector<uint32_t> queries(max_q);
  std::vector<QM31> evals(max_q), alphas(V_MAX_LAYERS);
  VProofState st;
  VGen gp = make_gen();
  verify_phase_a(words, n_words, seed_or_null, gp, st, leaves.data(), max_pos, queries.data(), max_q, evals.data(),
                 alphas.data());
  for (uint32_t l = 0; l < st.n_layers; l++)
    if (st.n_leaf[l])
      st.merkle_ok[l] = verify_phase_b(words, n_words, st.layer_off[l], st.D - l, leaves.data() + (size_t)l * max_pos,
                                       st.n_leaf[l]);
  return verify_resolve(st);
}

}  // extern "C"

// Common path: `words` = concatenated FRDA encodings (host memory, ideally pinned), offs = word offsets.
static int verify_batch_words(frieda_ctx *ctx, const uint32_t *words, const std::vector<unsigned long long> &offs,
                              size_t n, uint32_t n_layers_max, uint32_t max_q, const uint64_t *seeds_or_null,
                              int *results) {
  cudaError_t e = cudaSetDevice(frieda_ctx_device(ctx));
  if (e != cudaSuccess) return frieda_ctx_fail_cuda(ctx, (int)e, "cudaSetDevice");
  cudaStream_t st = (cudaStream_t)frieda_ctx_stream(ctx);
  const uint32_t max_pos = 2 * max_q;
  const size_t blob_bytes = offs[n] * 4;
  auto al = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o_blob = 0, o_offs = o_blob + al(blob_bytes), o_seeds = o_offs + al((n + 1) * 8);
  size_t o_st = o_seeds + al(n * 8), o_leaves = o_st + al(n * sizeof(VProofState));
  size_t o_q = o_leaves + al(n * (size_t)n_layers_max * max_pos * sizeof(VNode));
  size_t o_ev = o_q + al(n * (size_t)max_q * 4), o_al = o_ev + al(n * (size_t)max_q * sizeof(QM31));
  size_t total = o_al + al(n * (size_t)V_MAX_LAYERS * sizeof(QM31));
  uint8_t *d = nullptr;
  if ((e = cudaMalloc(&d, total)) != cudaSuccess) return frieda_ctx_fail_cuda(ctx, (int)e, "cudaMalloc(verify batch)");
  std::vector<VProofState> h_st(n);
  do {
    if ((e = cudaMemcpyAsync(d + o_blob, words, blob_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(d + o_offs, offs.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if (seeds_or_null &&
        (e = cudaMemcpyAsync(d + o_seeds, seeds_or_null, n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess)
      break;
    VGen gp = make_gen();
    frieda_ctx_prof_begin(ctx, "verify_phase_a");
    verify_phase_a_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(
        reinterpret_cast<const uint32_t *>(d + o_blob), reinterpret_cast<const unsigned long long *>(d + o_offs),
        seeds_or_null ? reinterpret_cast<const unsigned long long *>(d + o_seeds) : nullptr, gp,
        reinterpret_cast<VProofState *>(d + o_st), reinterpret_cast<VNode *>(d + o_leaves), n_layers_max, max_pos,
        reinterpret_cast<uint32_t *>(d + o_q), max_q, reinterpret_cast<QM31 *>(d + o_ev),
        reinterpret_cast<QM31 *>(d + o_al), n);
    frieda_ctx_prof_end(ctx);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    size_t nt = n * n_layers_max;
    frieda_ctx_prof_begin(ctx, "verify_phase_b");
    verify_phase_b_kernel<<<(unsigned)((nt + 63) / 64), 64, 0, st>>>(
        reinterpret_cast<const uint32_t *>(d + o_blob), reinterpret_cast<const unsigned long long *>(d + o_offs),
        reinterpret_cast<VProofState *>(d + o_st), reinterpret_cast<VNode *>(d + o_leaves), n_layers_max, max_pos, n);
    frieda_ctx_prof_end(ctx);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    frieda_ctx_count_launches(ctx, 2);
    if ((e = cudaMemcpyAsync(h_st.data(), d + o_st, n * sizeof(VProofState), cudaMemcpyDeviceToHost, st)) != cudaSuccess)
      break;
    e = cudaStreamSynchronize(st);
  } while (0);
  cudaFree(d);
  if (e != cudaSuccess) return frieda_ctx_fail_cuda(ctx, (int)e, "frieda_verify_batch");
  for (size_t i = 0; i < n; i++) results[i] = verify_resolve(h_st[i]);
  return FRIEDA_OK;
}

extern "C" int frieda_verify_batch_bytes(frieda_ctx *ctx, const uint8_t *bytes, size_t bytes_len,
                                         const uint64_t *byte_offsets, size_t n, const uint64_t *seeds_or_null,
                                         int *results);
// Serialised proofs (frieda_proof_serialize encoding), proof i at bytes + byte_offsets[i] .. byte_offsets[i+1];
// offsets must be multiples of 4 and end inside the buffer.  This is what a light client holds after receiving
// proofs: every byte is untrusted, the offsets and bytes_len are the caller's.
int frieda_verify_batch_bytes(frieda_ctx *ctx, const uint8_t *bytes, size_t bytes_len, const uint64_t *byte_offsets,
                              size_t n, const uint64_t *seeds_or_null, int *results) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!bytes || !byte_offsets || !results) return frieda_ctx_fail_arg(ctx, "null pointer");
  if (n == 0) return FRIEDA_OK;
  if ((reinterpret_cast<uintptr_t>(bytes) & 3) != 0) return frieda_ctx_fail_arg(ctx, "proof bytes must be 4-byte aligned");
  const uint32_t *words = reinterpret_cast<const uint32_t *>(bytes);
  std::vector<unsigned long long> offs(n + 1);
  uint32_t n_layers_max = 1, max_q = 1;
  for (size_t i = 0; i <= n; i++) {
    if (byte_offsets[i] & 3) return frieda_ctx_fail_arg(ctx, "proof offsets must be multiples of 4");
    if (byte_offsets[i] > bytes_len) return frieda_ctx_fail_arg(ctx, "proof offset beyond the end of the byte buffer");
    offs[i] = byte_offsets[i] / 4;
    if (i && offs[i] < offs[i - 1]) return frieda_ctx_fail_arg(ctx, "proof offsets must be non-decreasing");
    if (i && offs[i] - offs[i - 1] > 0xffffffffull) return frieda_ctx_fail_arg(ctx, "a proof longer than 16 GiB");
  }
  for (size_t i = 0; i < n; i++) {
    // header: magic, log_size_bound, log_blowup, log_last, n_queries (u64), pow_bits, pow (u64), n_evals, ...
    const uint32_t *w = words + offs[i];
    const size_t len = offs[i + 1] - offs[i];
    uint32_t nq = 1, nl = 1;
    if (len > 10) {
      // capacity of this verifier (frieda_verify_batch returns the same error; frieda_verify has neither limit):
      // a proof beyond it is UNSUPPORTED here, which must not be reported as "forged"
      if (w[0] == 0x41445246u && w[5] == 0 && w[4] > 4096)
        return frieda_ctx_fail_arg(ctx, "n_queries > 4096 is not supported by the batch verifier");
      if (w[0] == 0x41445246u && w[3] > 20 && w[3] < 32)
        return frieda_ctx_fail_arg(ctx, "log_last_layer_degree_bound > 20 is not supported by the batch verifier");
      nq = w[5] ? 1u : w[4];  // n_queries >= 2^32: phase A reports the reference's behaviour without drawing
      size_t pos = 10 + 4 * (size_t)w[9];
      if (pos < len) {
        pos += 1 + 4 * (size_t)w[pos];
        if (pos < len) nl = w[pos];
      }
    }
    if (nl > V_MAX_LAYERS) nl = V_MAX_LAYERS;
    max_q = nq > max_q ? nq : max_q;
    n_layers_max = nl > n_layers_max ? nl : n_layers_max;
  }
  return verify_batch_words(ctx, words, offs, n, n_layers_max, max_q, seeds_or_null, results);
}

extern "C" {

int frieda_verify_batch(frieda_ctx *ctx, const frieda_proof *const *proofs, size_t n, const uint64_t *seeds_or_null,
                        int *results) {
  if (!ctx) return FRIEDA_ERR_ARG;
  if (!proofs || !results) return frieda_ctx_fail_arg(ctx, "null pointer");
  if (n == 0) return FRIEDA_OK;
  std::vector<unsigned long long> offs(n + 1, 0);
  uint32_t n_layers_max = 1, max_q = 1;
  for (size_t i = 0; i < n; i++) {
    if (!proofs[i]) return frieda_ctx_fail_arg(ctx, "null proof");
    size_t bytes = frieda_proof_serialize(proofs[i], nullptr, 0);
    offs[i + 1] = offs[i] + (bytes + 3) / 4;
    uint32_t nl = 1 + proofs[i]->n_inner_layers;
    if (nl > V_MAX_LAYERS) return frieda_ctx_fail_arg(ctx, "too many FRI layers in a proof");
    if (proofs[i]->pcs_config.n_queries > 4096) return frieda_ctx_fail_arg(ctx, "n_queries > 4096 is not supported");
    n_layers_max = nl > n_layers_max ? nl : n_layers_max;
    uint32_t q = (uint32_t)proofs[i]->pcs_config.n_queries;
    max_q = q > max_q ? q : max_q;
  }
  const size_t blob_bytes = offs[n] * 4;
  // serialise straight into pinned memory, on a few host threads
  uint32_t *words = nullptr;
  cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&words), blob_bytes ? blob_bytes : 4, cudaHostAllocDefault);
  if (e != cudaSuccess) return frieda_ctx_fail_cuda(ctx, (int)e, "cudaHostAlloc(verify batch)");
  {
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    if (n < 64) nt = 1;
    auto work = [&](unsigned t) {
      for (size_t i = t; i < n; i += nt)
        frieda_proof_serialize(proofs[i], reinterpret_cast<uint8_t *>(words + offs[i]), (offs[i + 1] - offs[i]) * 4);
    };
    if (nt == 1) {
      work(0);
    } else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
      for (auto &x : th) x.join();
    }
  }
  int rc = verify_batch_words(ctx, words, offs, n, n_layers_max, max_q, seeds_or_null, results);
  cudaFreeHost(words);
  return rc;
}

}  // extern "C"
