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
// Grid: x = blob, y = one of the blob's CTAs.  CTAs are scheduled in x-major order, so every blob of the wave starts
// at once with one or two CTAs, a CTA whose blob is already solved exits at its first grab, and the hardware
// scheduler keeps handing the freed slots to the blobs that are still open: the unlucky blobs of a wave (the search
// length is geometric) end up with up to 128 CTAs each.  With y = blob (round 1: blobs solved 27 at a time, 32 CTAs
// each) the last blobs of every wave ran on a nearly empty GPU: 22.7 -> 20.7 ms per 512 proofs of work at pow 20
// (28.8 G compressions/s; the hot loop is 614 ALU-pipe instructions, i.e. a roof of 30.3 G/s).
constexpr int GR_THREADS = 256;
// nonces per grab: 4 per lane for batches; 1 per lane when a few blobs have the whole GPU to themselves
// (every resident warp examines at least one chunk, and with 128-nonce chunks that alone is more work
// than the expected 2^20-nonce search of the reference's benches)
constexpr uint32_t GR_CHUNK_BATCH = 128, GR_CHUNK_SINGLE = 32;

template <uint32_t GR_CHUNK>
__global__ void __launch_bounds__(GR_THREADS) grind_kernel(const Channel *__restrict__ chan, uint32_t pow_bits,
                                                           uint64_t limit, unsigned long long *best,
                                                           unsigned long long *next, uint32_t one) {
  const size_t blob = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) d[i] = chan[blob].digest[i];
  const uint32_t low_mask = pow_bits >= 32 ? 0xffffffffu : ((1u << pow_bits) - 1u);
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
      // The screen reads digest word 0 only, so everything of the last round that does not feed
      // v0 / v8 (two whole G functions and the tails of four more) is dead code in this loop.
      if ((h[0] & low_mask) == 0) {
        // rare (2^-min(pow_bits, 32) per nonce): the whole digest for the exact trailing-zero count;
        // the nonce goes through an opaque asm so this copy is not merged with the screen above
        uint32_t lo = (uint32_t)nonce, hi = (uint32_t)(nonce >> 32);
        asm volatile("" : "+r"(lo), "+r"(hi));
        uint32_t g[8];
#pragma unroll
        for (int i = 0; i < 8; i++) g[i] = d[i];
        uint32_t m2[16] = {lo, hi, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        blake2s_compress_t<0x0003u>(g, m2, 0, 0, 0, 0, one);
        if (digest_trailing_zeros(g) >= pow_bits) atomicMin(best + blob, (unsigned long long)nonce);
      }
    }
  }
}

cudaError_t launch_grind(cudaStream_t st, const Channel *chan, uint32_t pow_bits, uint64_t limit,
                         unsigned long long *best, unsigned long long *next, size_t n_blobs) {
  if (n_blobs == 0) return cudaSuccess;
  if (n_blobs > 0x7fffffffu) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(next, 0, n_blobs * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  // up to 128 CTAs per blob; more when the wave alone would not fill the GPU (148 SMs x 6 CTAs x ~2.7)
  uint32_t ctas = 128;
  while ((size_t)ctas * n_blobs < 2368 && ctas < 2048) ctas <<= 1;
  const dim3 grid((unsigned)n_blobs, ctas);
  if (n_blobs < 16)
    grind_kernel<GR_CHUNK_SINGLE><<<grid, GR_THREADS, 0, st>>>(chan, pow_bits, limit, best, next, 1u);
  else
    grind_kernel<GR_CHUNK_BATCH><<<grid, GR_THREADS, 0, st>>>(chan, pow_bits, limit, best, next, 1u);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- queries
// mix_u64(nonce), then exactly n_queries draws of 4-byte chunks masked to the domain, collected into an ordered set.
// One CTA per blob: draw k of the channel hashes (digest, counter k), so the ceil(n_queries / 8) draws are
// independent compressions spread over the threads; the values are then sorted (bitonic, shared memory) and
// deduplicated with a scan.  (Round 1 did all of it in one thread per blob: 0.65 ms per 512 proofs.)
constexpr int Q_THREADS = 128;
constexpr uint32_t Q_MAX = 4096;  // n_queries limit of the proof path (ctx.cu)

__global__ void __launch_bounds__(Q_THREADS) queries_kernel(Channel *chan, const unsigned long long *nonce,
                                                            uint32_t log_domain, uint32_t n_queries, uint32_t *queries,
                                                            uint32_t *n_unique) {
  __shared__ uint32_t vals[Q_MAX];
  __shared__ uint32_t part[Q_THREADS];
  __shared__ Channel s_ch;
  const size_t b = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    Channel c = chan[b];
    channel_mix_u64(c, nonce[b]);  // src/proof.rs:59
    s_ch = c;
  }
  __syncthreads();
  const uint32_t mask = log_domain >= 32 ? 0xffffffffu : ((1u << log_domain) - 1);
  const uint32_t n_draws = (n_queries + 7) / 8;
  uint32_t n_pad = 1;
  while (n_pad < n_queries) n_pad <<= 1;
  for (uint32_t i = n_queries + tid; i < n_pad; i += Q_THREADS) vals[i] = 0xffffffffu;  // sorts behind every position
  for (uint32_t k = tid; k < n_draws; k += Q_THREADS) {
    Channel c = s_ch;
    c.n_sent += k;  // draw k: H(digest || counter k)
    uint32_t w[8];
    channel_draw_random_words(c, w);
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (8 * k + i < n_queries) vals[8 * k + i] = w[i] & mask;
  }
  __syncthreads();
  // bitonic sort of n_pad values, ascending
  for (uint32_t size = 2; size <= n_pad; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = tid; t < (n_pad >> 1); t += Q_THREADS) {
        const uint32_t lo = ((t / stride) * stride << 1) + (t % stride), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint32_t x = vals[lo], y = vals[hi];
        if ((x > y) == up) {
          vals[lo] = y;
          vals[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  // keep the first of every run of equal values: per-thread contiguous piece, scan of the piece counts
  const uint32_t per = (n_pad + Q_THREADS - 1) / Q_THREADS;
  const uint32_t lo = tid * per, hi = lo + per < n_queries ? lo + per : n_queries;
  uint32_t cnt = 0;
  for (uint32_t i = lo; i < hi; i++) cnt += (i == 0 || vals[i] != vals[i - 1]) ? 1u : 0u;
  part[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int t = 0; t < Q_THREADS; t++) {
      const uint32_t c = part[t];
      part[t] = run;
      run += c;
    }
    n_unique[b] = run;
    Channel c = s_ch;
    c.n_sent += n_draws;
    chan[b] = c;
  }
  __syncthreads();
  uint32_t *q = queries + b * n_queries;
  uint32_t at = part[tid];
  for (uint32_t i = lo; i < hi; i++)
    if (i == 0 || vals[i] != vals[i - 1]) q[at++] = vals[i];
}
cudaError_t launch_queries(cudaStream_t st, Channel *chan, const unsigned long long *nonce, uint32_t log_domain,
                           uint32_t n_queries, uint32_t *queries, uint32_t *n_unique, size_t n_blobs) {
  if (n_queries == 0 || n_queries > Q_MAX) return cudaErrorInvalidValue;
  queries_kernel<<<(unsigned)n_blobs, Q_THREADS, 0, st>>>(chan, nonce, log_domain, n_queries, queries, n_unique);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- decommit
// Layer l (log d = D - l) is queried at distinct(q >> l).  Its decommitment positions are the
// sibling pairs {2g, 2g+1}, g = distinct(q >> (l+1)); positions that are not queries contribute
// their QM31 value to fri_witness.  The Merkle multi-proof walks the levels k = d-2 .. 0: node n
// of level k is on the path iff some query has q >> (l + d - k) == n; each of its children
// (level k+1) that is not on the path contributes its hash to hash_witness (left, then right).
// All lists are shifts of the one sorted query array, so every (blob, layer, walk) is independent:
// one thread per walk, walk 0 = the fri witness, walk j >= 1 = tree level k = d - 1 - j.  A count
// pass, a scan (per layer, then over all layers of the wave) and a write pass that repeats the walk
// with its output offset known.
// Runs of equal (q >> sh): every run yields the children (bit sh-1) that no query covers.
template <class Emit>
__device__ __forceinline__ uint32_t walk_runs(const uint32_t *__restrict__ q, uint32_t nq, uint32_t sh, Emit emit) {
  uint32_t cnt = 0, i = 0;
  while (i < nq) {
    const uint32_t node = sh >= 32 ? 0u : q[i] >> sh;
    bool has0 = false, has1 = false;
    while (i < nq && (sh >= 32 ? 0u : q[i] >> sh) == node) {
      if ((q[i] >> (sh - 1)) & 1u) has1 = true; else has0 = true;
      i++;
    }
    if (!has0) { emit(cnt, node, 0u); cnt++; }
    if (!has1) { emit(cnt, node, 1u); cnt++; }
  }
  return cnt;
}

// lvl[(blob * n_layers + layer) * lvl_stride + walk]: count of the walk, turned into the exclusive
// prefix inside the layer's hash witness by the scan (walk 0, the fri witness, starts its own list).
__global__ void __launch_bounds__(128) decommit_count_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = t / p.lvl_stride;  // (blob, layer)
  const uint32_t walk = (uint32_t)(t % p.lvl_stride);
  if (g >= n_blobs * p.n_layers) return;
  const size_t blob = g / p.n_layers;
  const uint32_t layer = (uint32_t)(g % p.n_layers);
  const uint32_t d = p.D - layer;
  uint32_t c = 0;
  if (walk < d) {
    const uint32_t sh = walk == 0 ? layer + 1 : layer + 1 + walk;  // walk j: node = q >> (layer + d - k), k = d-1-j
    c = walk_runs(p.queries + blob * p.n_queries, p.n_unique[blob], sh, [](uint32_t, uint32_t, uint32_t) {});
  }
  p.lvl[t] = c;
}

// One thread per (blob, layer) sums its walks; one CTA then scans the wave.
__global__ void __launch_bounds__(128) decommit_layer_sum_kernel(const __grid_constant__ DecommitParams p,
                                                                 size_t n_blobs) {
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_blobs * p.n_layers) return;
  uint32_t *lv = p.lvl + g * p.lvl_stride;
  const uint32_t d = p.D - (uint32_t)(g % p.n_layers);
  uint32_t run = 0;
  for (uint32_t j = 1; j < d; j++) {
    uint32_t c = lv[j];
    lv[j] = run;
    run += c;
  }
  p.counts[g * 2] = lv[0];
  p.counts[g * 2 + 1] = run;
}
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) decommit_scan_kernel(const __grid_constant__ DecommitParams p,
                                                                     size_t n_blobs, unsigned long long *totals) {
  // blob-major, layer-minor exclusive scan of both counters: contiguous chunk per thread,
  // Hillis-Steele over the chunk sums in shared memory
  __shared__ unsigned long long sf[SCAN_THREADS], sh[SCAN_THREADS];
  const size_t N = n_blobs * p.n_layers;
  const size_t chunk = (N + SCAN_THREADS - 1) / SCAN_THREADS;
  const size_t lo = (size_t)threadIdx.x * chunk, hi = lo + chunk < N ? lo + chunk : N;
  unsigned long long f = 0, h = 0;
  for (size_t g = lo; g < hi; g++) {
    f += p.counts[g * 2];
    h += p.counts[g * 2 + 1];
  }
  sf[threadIdx.x] = f;
  sh[threadIdx.x] = h;
  __syncthreads();
  for (int off = 1; off < SCAN_THREADS; off <<= 1) {
    unsigned long long af = 0, ah = 0;
    if ((int)threadIdx.x >= off) {
      af = sf[threadIdx.x - off];
      ah = sh[threadIdx.x - off];
    }
    __syncthreads();
    sf[threadIdx.x] += af;
    sh[threadIdx.x] += ah;
    __syncthreads();
  }
  unsigned long long bf = sf[threadIdx.x] - f, bh = sh[threadIdx.x] - h;  // exclusive
  for (size_t g = lo; g < hi; g++) {
    p.offsets[g * 2] = bf;
    p.offsets[g * 2 + 1] = bh;
    bf += p.counts[g * 2];
    bh += p.counts[g * 2 + 1];
  }
  if (threadIdx.x == SCAN_THREADS - 1) {
    totals[0] = sf[threadIdx.x];
    totals[1] = sh[threadIdx.x];
  }
}

__global__ void __launch_bounds__(128) decommit_write_kernel(const __grid_constant__ DecommitParams p, size_t n_blobs) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = t / p.lvl_stride;
  const uint32_t walk = (uint32_t)(t % p.lvl_stride);
  if (g >= n_blobs * p.n_layers) return;
  const size_t blob = g / p.n_layers;
  const uint32_t layer = (uint32_t)(g % p.n_layers);
  const uint32_t d = p.D - layer;
  if (walk >= d) return;
  const uint32_t *q = p.queries + blob * p.n_queries;
  const uint32_t nq = p.n_unique[blob];
  const unsigned long long *off = p.offsets + g * 2;
  if (walk == 0) {
    const uint32_t *cols = p.cols[layer] + blob * p.cols_stride[layer];
    const size_t n = (size_t)1 << d;
    QM31 *fri = p.fri_out + off[0];
    walk_runs(q, nq, layer + 1, [&](uint32_t at, uint32_t node, uint32_t s) {
      const size_t pos = 2 * (size_t)node + s;
      fri[at] = {{cols[pos], cols[n + pos], cols[2 * n + pos], cols[3 * n + pos]}};
    });
  } else {
    const uint32_t k = d - 1 - walk;
    const uint4 *tree = reinterpret_cast<const uint4 *>(p.tree[layer]) + 2 * blob * p.tree_stride[layer];
    uint4 *hw = reinterpret_cast<uint4 *>(p.hash_out) + 2 * (off[1] + p.lvl[t]);
    walk_runs(q, nq, layer + 1 + walk, [&](uint32_t at, uint32_t node, uint32_t s) {
      const size_t slot = ((size_t)1 << (k + 1)) + 2 * (size_t)node + s;
      hw[2 * (size_t)at] = tree[2 * slot];
      hw[2 * (size_t)at + 1] = tree[2 * slot + 1];
    });
  }
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

// Split blob (frieda_fri_split_decommit): the values and hashes THIS rank holds, picked out by address.
__global__ void __launch_bounds__(128) gather_items_kernel(const GatherQ *__restrict__ dq, uint32_t n_q,
                                                           const GatherH *__restrict__ dh, uint32_t n_h,
                                                           QM31 *__restrict__ out_q, uint4 *__restrict__ out_h) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_q) {
    const GatherQ d = dq[t];
    out_q[t] = {{d.p[0], d.p[d.stride], d.p[2 * (size_t)d.stride], d.p[3 * (size_t)d.stride]}};
  } else if (t < n_q + n_h) {
    const uint4 *src = dh[t - n_q].p;
    out_h[2 * (size_t)(t - n_q)] = src[0];
    out_h[2 * (size_t)(t - n_q) + 1] = src[1];
  }
}
cudaError_t launch_gather_items(cudaStream_t st, const GatherQ *dq, uint32_t n_q, const GatherH *dh, uint32_t n_h,
                                QM31 *out_q, uint8_t *out_h) {
  const uint32_t n = n_q + n_h;
  if (n == 0) return cudaSuccess;
  gather_items_kernel<<<(n + 127) / 128, 128, 0, st>>>(dq, n_q, dh, n_h, out_q, reinterpret_cast<uint4 *>(out_h));
  return cudaGetLastError();
}

cudaError_t launch_decommit_count(cudaStream_t st, const DecommitParams &p, size_t n_blobs) {
  if (p.lvl_stride < p.D) return cudaErrorInvalidValue;
  size_t n = n_blobs * p.n_layers * p.lvl_stride;
  decommit_count_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, n_blobs);
  return cudaGetLastError();
}
cudaError_t launch_decommit_scan(cudaStream_t st, const DecommitParams &p, size_t n_blobs,
                                 unsigned long long *totals) {
  size_t n = n_blobs * p.n_layers;
  decommit_layer_sum_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, n_blobs);
  decommit_scan_kernel<<<1, SCAN_THREADS, 0, st>>>(p, n_blobs, totals);
  return cudaGetLastError();
}
cudaError_t launch_decommit_write(cudaStream_t st, const DecommitParams &p, size_t n_blobs) {
  size_t n = n_blobs * p.n_layers * p.lvl_stride;
  decommit_write_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(p, n_blobs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  size_t ne = n_blobs * p.n_queries;
  evaluations_kernel<<<(unsigned)((ne + 127) / 128), 128, 0, st>>>(p, n_blobs);
  return cudaGetLastError();
}

}  // namespace frieda
