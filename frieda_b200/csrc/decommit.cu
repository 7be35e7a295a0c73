// decommit.cu -- proof-of-work grind, query sampling and the batched query-gather kernels.
//
// Replaces, for src/proof.rs:58-66:
//   CpuBackend::grind(channel, pow_bits)                      -> grind_kernel (min-nonce search)
//   channel.mix_u64(nonce); Queries::generate(...)            -> queries_kernel
//   FriProver::decommit -> compute_decommitment_positions_and_witness_evals +
//   MerkleProver::decommit per layer, and the `evaluations` gather (src/proof.rs:62-66)
//                                                             -> decommit_{count,scan,write}
// "parity unpinned" conventions are confined to blake2s.cuh and this file.
#include "kernels.cuh"

namespace frieda {

// ---------------------------------------------------------------- grind
// CpuBackend::grind returns the SMALLEST nonce whose mix has >= pow_bits trailing zeros.  Each warp
// takes the next chunk of GR_CHUNK nonces of its blob from an atomic counter, so chunks are handed
// out in strictly increasing order no matter when a CTA was scheduled or how fast it runs; a warp
// stops when the chunk it was handed starts above the best nonce found so far.  Every nonce below
// the answer is examined (the minimum is exact), little above it is, and there is no host round trip.
constexpr int GR_THREADS = 256;
constexpr uint32_t GR_CHUNK = 128;  // nonces per grab = 4 per lane

__global__ void __launch_bounds__(GR_THREADS) grind_kernel(const Channel *__restrict__ chan, uint32_t pow_bits,
                                                           uint64_t limit, unsigned long long *best,
                                                           unsigned long long *next, uint32_t one) {
  const size_t blob = blockIdx.y;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = chan[blob].digest[i];
  for (;;) {
    unsigned long long start = 0, cur = 0;
    if (lane == 0) {
      start = atomicAdd(next + blob, (unsigned long long)GR_CHUNK);
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(best + blob));
    }
    start = __shfl_sync(0xffffffffu, start, 0);
    cur = __shfl_sync(0xffffffffu, cur, 0);
    if (start > cur || start >= limit) return;
#pragma unroll 1
    for (uint32_t k = 0; k < GR_CHUNK / 32; k++) {
      const uint64_t nonce = start + k * 32 + lane;
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; i++) h[i] = d[i];
      uint32_t m[16] = {(uint32_t)nonce, (uint32_t)(nonce >> 32), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      blake2s_compress_t<0x0003u>(h, m, 0, 0, 0, 0, one);  // only the two nonce words are non-zero
      if (digest_trailing_zeros(h) >= pow_bits) atomicMin(best + blob, (unsigned long long)nonce);
    }
  }
}

cudaError_t launch_grind(cudaStream_t st, const Channel *chan, uint32_t pow_bits, uint64_t limit, uint32_t ctas_per_blob,
                         unsigned long long *best, unsigned long long *next, size_t n_blobs) {
  if (ctas_per_blob == 0) ctas_per_blob = 1;
  cudaError_t e = cudaMemsetAsync(next, 0, n_blobs * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  for (size_t b0 = 0; b0 < n_blobs; b0 += 32768) {
    size_t nb = n_blobs - b0 < 32768 ? n_blobs - b0 : 32768;
    grind_kernel<<<dim3(ctas_per_blob, (unsigned)nb), GR_THREADS, 0, st>>>(chan + b0, pow_bits, limit, best + b0,
                                                                          next + b0, 1u);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- queries
// One thread per blob: mix_u64(nonce), then exactly n_queries draws of 4-byte chunks masked to
// the domain, inserted into an ordered set (insertion sort + dedup).
__global__ void queries_kernel(Channel *chan, const unsigned long long *nonce, uint32_t log_domain,
                               uint32_t n_queries, uint32_t *queries, uint32_t *n_unique, size_t n) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  Channel c = chan[b];
  channel_mix_u64(c, nonce[b]);  // src/proof.rs:59
  uint32_t *q = queries + b * n_queries;
  const uint32_t mask = log_domain >= 32 ? 0xffffffffu : ((1u << log_domain) - 1);
  uint32_t cnt = 0, m = 0;
  while (cnt < n_queries) {
    uint32_t w[8];
    channel_draw_random_words(c, w);
    for (int i = 0; i < 8 && cnt < n_queries; i++, cnt++) {
      uint32_t v = w[i] & mask;
      // sorted insert, skipping duplicates
      uint32_t lo = 0, hi = m;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (q[mid] < v) lo = mid + 1; else hi = mid;
      }
      if (lo < m && q[lo] == v) continue;
      for (uint32_t j = m; j > lo; j--) q[j] = q[j - 1];
      q[lo] = v;
      m++;
    }
  }
  n_unique[b] = m;
  chan[b] = c;
}
cudaError_t launch_queries(cudaStream_t st, Channel *chan, const unsigned long long *nonce, uint32_t log_domain,
                           uint32_t n_queries, uint32_t *queries, uint32_t *n_unique, size_t n_blobs) {
  queries_kernel<<<(unsigned)((n_blobs + 63) / 64), 64, 0, st>>>(chan, nonce, log_domain, n_queries, queries, n_unique,
                                                                n_blobs);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- decommit
// Layer l (log d = D - l) is queried at distinct(q >> l).  Its decommitment positions are the
// sibling pairs {2g, 2g+1}, g = distinct(q >> (l+1)); positions that are not queries contribute
// their QM31 value to fri_witness.  The Merkle multi-proof walks the levels k = d-2 .. 0: node n
// of level k is on the path iff some query has q >> (l + d - k) == n; each of its children
// (level k+1) that is not on the path contributes its hash to hash_witness (left, then right).
// All lists are shifts of the one sorted query array, so the walk needs no scratch storage.
template <bool WRITE>
__device__ void decommit_layer(const DecommitParams &p, size_t blob, uint32_t layer, uint32_t *n_fri_out,
                               uint32_t *n_hash_out) {
  const uint32_t *q = p.queries + blob * p.n_queries;
  const uint32_t nq = p.n_unique[blob];
  const uint32_t d = p.D - layer;
  const uint32_t *cols = p.cols[layer] + blob * p.cols_stride[layer];
  const uint4 *tree = reinterpret_cast<const uint4 *>(p.tree[layer]) + 2 * blob * p.tree_stride[layer];
  QM31 *fri = nullptr;
  uint4 *hw = nullptr;
  if (WRITE) {
    const unsigned long long *off = p.offsets + (blob * p.n_layers + layer) * 2;
    fri = p.fri_out + off[0];
    hw = reinterpret_cast<uint4 *>(p.hash_out) + 2 * off[1];
  }
  uint32_t n_fri = 0, n_hash = 0;
  const size_t n = (size_t)1 << d;
  // fri witness: sibling positions that are not themselves queried
  {
    uint32_t i = 0;
    while (i < nq) {
      uint32_t g = q[i] >> (layer + 1);
      bool has0 = false, has1 = false;
      while (i < nq && (q[i] >> (layer + 1)) == g) {
        if ((q[i] >> layer) & 1u) has1 = true; else has0 = true;
        i++;
      }
      for (uint32_t s = 0; s < 2; s++) {
        if (s == 0 ? has0 : has1) continue;
        if (WRITE) {
          size_t pos = 2 * (size_t)g + s;
          fri[n_fri] = {{cols[pos], cols[n + pos], cols[2 * n + pos], cols[3 * n + pos]}};
        }
        n_fri++;
      }
    }
  }
  // hash witness
  for (int k = (int)d - 2; k >= 0; k--) {
    const uint32_t sh = layer + d - (uint32_t)k;  // node = q >> sh ; child = q >> (sh - 1)
    uint32_t i = 0;
    while (i < nq) {
      uint32_t node = q[i] >> sh;
      bool has0 = false, has1 = false;
      while (i < nq && (q[i] >> sh) == node) {
        if ((q[i] >> (sh - 1)) & 1u) has1 = true; else has0 = true;
        i++;
      }
      for (uint32_t s = 0; s < 2; s++) {
        if (s == 0 ? has0 : has1) continue;
        if (WRITE) {
          size_t slot = ((size_t)1 << (k + 1)) + 2 * (size_t)node + s;
          hw[2 * (size_t)n_hash] = tree[2 * slot];
          hw[2 * (size_t)n_hash + 1] = tree[2 * slot + 1];
        }
        n_hash++;
      }
    }
  }
  *n_fri_out = n_fri;
  *n_hash_out = n_hash;
}

__global__ void decommit_count_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_blobs * p.n_layers) return;
  size_t blob = g / p.n_layers;
  uint32_t layer = (uint32_t)(g % p.n_layers);
  uint32_t nf, nh;
  decommit_layer<false>(p, blob, layer, &nf, &nh);
  p.counts[g * 2] = nf;
  p.counts[g * 2 + 1] = nh;
}
__global__ void decommit_scan_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs,
                                     unsigned long long *totals) {
  // tiny: a single thread walks blob-major, layer-minor
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned long long f = 0, h = 0;
  for (size_t g = 0; g < n_blobs * p.n_layers; g++) {
    p.offsets[g * 2] = f;
    p.offsets[g * 2 + 1] = h;
    f += p.counts[g * 2];
    h += p.counts[g * 2 + 1];
  }
  totals[0] = f;
  totals[1] = h;
}
__global__ void decommit_write_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_blobs * p.n_layers) return;
  size_t blob = g / p.n_layers;
  uint32_t layer = (uint32_t)(g % p.n_layers);
  uint32_t nf, nh;
  decommit_layer<true>(p, blob, layer, &nf, &nh);
}
// evaluations at the query positions, ascending (src/proof.rs:62-66)
__global__ void evaluations_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_blobs * p.n_queries) return;
  size_t blob = g / p.n_queries;
  uint32_t i = (uint32_t)(g % p.n_queries);
  if (i >= p.n_unique[blob]) return;
  const uint32_t *cols = p.cols[0] + blob * p.cols_stride[0];
  size_t n = (size_t)1 << p.D;
  uint32_t pos = p.queries[blob * p.n_queries + i];
  p.evals_out[g] = {{cols[pos], cols[n + pos], cols[2 * n + pos], cols[3 * n + pos]}};
}

cudaError_t launch_decommit_count(cudaStream_t st, const DecommitParams &p, size_t n_blobs) {
  size_t n = n_blobs * p.n_layers;
  decommit_count_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(p, n_blobs);
  return cudaGetLastError();
}
cudaError_t launch_decommit_scan(cudaStream_t st, const DecommitParams &p, size_t n_blobs,
                                 unsigned long long *totals) {
  decommit_scan_kernel<<<1, 32, 0, st>>>(p, n_blobs, totals);
  return cudaGetLastError();
}
cudaError_t launch_decommit_write(cudaStream_t st, const DecommitParams &p, size_t n_blobs) {
  size_t n = n_blobs * p.n_layers;
  decommit_write_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(p, n_blobs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  size_t ne = n_blobs * p.n_queries;
  evaluations_kernel<<<(unsigned)((ne + 127) / 128), 128, 0, st>>>(p, n_blobs);
  return cudaGetLastError();
}

}  // namespace frieda
