// fri_small.cu -- latency path of the FRI commit phase: the layers above the tail of a FEW blobs as ONE
// cooperative launch (BASELINE config 2: a single 128 KiB blob; the criterion benches of benches/proof.rs).
//
// FriProver::commit (src/proof.rs:52-57) is a chain: layer l+1 cannot be folded before the root of layer l went into
// the channel (Fiat-Shamir).  For one blob that chain -- about 220 dependent compressions -- is the whole cost, and
// the throughput path (merkle.cu: one launch per tree pass, 19 launches for nine layers) pays a launch boundary per
// link and runs two warps per scheduler where one would finish a compression in half the time.  Here:
//   * one CTA per (blob, chunk of 2^CL leaves), all co-resident (cooperative launch); 2^(CL-1) threads: one node
//     per thread on every tree level;
//   * per layer: fold (layers >= 1) + leaf hashes + CL levels in shared memory -> chunk root -> release on a
//     per-(layer, blob) counter.  CTAs that own a chunk of the NEXT layer acquire the counter, read all chunk roots,
//     and each of them reduces the top of the tree and runs mix_root / draw_felt itself (same bytes, same alpha):
//     one point of synchronisation per layer and no grid-wide barrier; the others exit;
//   * the tree level buffer is the parity-split structure of arrays of merkle.cu, reduced in place.
// Everything written during the kernel by other CTAs (chunk roots, the previous layer's columns) is read with
// ld.global.cg after the acquire -- never through the non-coherent path.
#include "kernels.cuh"

namespace frieda {

namespace {

struct alignas(16) FsHash {
  uint4 lo, hi;
};
__device__ __forceinline__ void fs_store(FsHash *dst, const uint32_t h[8]) {
  dst->lo = make_uint4(h[0], h[1], h[2], h[3]);
  dst->hi = make_uint4(h[4], h[5], h[6], h[7]);
}

// word s of node n of the level buffer: [parity][s][n >> 1], the odd half shifted by 16 banks
template <uint32_t HALF>
__device__ __forceinline__ uint32_t *fs_word(uint32_t *sm, uint32_t node, uint32_t s) {
  const uint32_t par = node & 1u;
  return sm + (par * 8 + s) * HALF + (node >> 1) + par * 16;
}
template <uint32_t HALF>
struct FsPairMsg {
  const uint32_t *e, *o;  // word 0 of the left (even) and of the right (odd) child
  __device__ __forceinline__ uint32_t operator[](int s) const { return s < 8 ? e[s * HALF] : o[(s - 8) * HALF]; }
};
template <uint32_t HALF>
__device__ __forceinline__ void fs_put(uint32_t *sm, uint32_t node, const uint32_t h[8]) {
#pragma unroll
  for (int s = 0; s < 8; s++) *fs_word<HALF>(sm, node, s) = h[s];
}

__device__ __forceinline__ uint32_t fs_circle_itw(const uint32_t *iblk, uint32_t i) {
  const uint32_t q = i >> 2, e = i & 3;
  const uint32_t v = __ldg(iblk + 2 * q + (e < 2 ? 1 : 0));
  return (e == 1 || e == 2) ? m31_neg(v) : v;
}

// Code footprint matters here: a lone warp runs straight-line compressions (about 18 KB of SASS each), and with a copy
// inlined at every call site the kernel outgrew the instruction cache -- the first version of this file spent 42 us
// per layer where the dependent chain accounts for 22 (bench_micro/chain.cu: 0.93 us per link).  So there is ONE
// leaf-compression site, ONE node-compression site (fs_reduce, not inlined, one node per thread per level) and ONE
// generic compression for the channel (fs_compress, not inlined).
__device__ __noinline__ void fs_compress(uint32_t h[8], const uint32_t m[16], uint32_t t0, uint32_t f0) {
  blake2s_compress_t<0xFFFFu>(h, m, t0, 0u, f0, 0u, 1u);
}
// Blake2sChannel steps over fs_compress (same bytes as blake2s.cuh's channel_*)
__device__ __forceinline__ void fs_b2s_block64(const uint32_t m[16], uint32_t out[8]) {
  b2s256_init(out);
  fs_compress(out, m, 64u, 0xFFFFFFFFu);
}
__device__ __forceinline__ void fs_mix_u64(Channel &c, uint64_t v) {
  uint32_t m[16] = {(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  fs_compress(c.digest, m, 0u, 0u);
  c.n_sent = 0;
}
__device__ __forceinline__ QM31 fs_mix_root_draw_felt(Channel &c, const uint32_t root[8]) {
  uint32_t m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    m[i] = c.digest[i];
    m[8 + i] = root[i];
  }
  fs_b2s_block64(m, c.digest);  // mix_root
  c.n_sent = 0;
  for (;;) {                    // draw_felt: retry until all 8 words < 2P
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = c.digest[i];
    m[8] = c.n_sent;
#pragma unroll
    for (int i = 9; i < 16; i++) m[i] = 0;
    c.n_sent += 1;
    uint32_t u[8];
    fs_b2s_block64(m, u);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) ok = ok && (u[i] < 2u * P31);
    if (!ok) continue;
    QM31 r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = u[i] >= P31 ? u[i] - P31 : u[i];
    return r;
  }
}

// Reduces the `cnt` (<= 2 x THREADS) nodes in the level buffer (tree level `level`, first node index idx0) by
// `levels` levels, in place, one node per thread per level.  Stores every level when store_mid and the last one when
// store_last; `top_out` receives the digest of node 0 of the last level on thread 0.
template <uint32_t HALF, int THREADS>
__device__ __noinline__ void fs_reduce(uint32_t *sm, uint32_t cnt, uint32_t levels, FsHash *tree, uint32_t level,
                                       uint32_t idx0, bool store_mid, bool store_last, uint32_t top_out[8]) {
#pragma unroll 1
  for (uint32_t l = 0; l < levels; l++) {
    cnt >>= 1;
    level -= 1;
    idx0 >>= 1;
    const bool last = l + 1 == levels;
    const uint32_t j = threadIdx.x;
    uint32_t h[8];
    if (j < cnt) merkle_hash_node_msg(FsPairMsg<HALF>{sm + j, sm + 8 * HALF + 16 + j}, h, 1u);
    __syncthreads();  // every compression of this level has read its children
    if (j < cnt) {
      if (!last) fs_put<HALF>(sm, j, h);
      if (last ? store_last : store_mid) fs_store(tree + ((size_t)1 << level) + idx0 + j, h);
    }
    if (last && j == 0) {
#pragma unroll
      for (int s = 0; s < 8; s++) top_out[s] = h[s];
    }
    if (!last) __syncthreads();
  }
}

__device__ __forceinline__ void fs_stamp(const FriSmallParams &p, uint32_t layer, int k) {
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[layer * 6 + k] = t;
  }
}

template <int CL>
__global__ void __launch_bounds__(1 << (CL - 1), CL == 9 ? 4 : 2) fri_small_kernel(const __grid_constant__ FriSmallParams p) {
  constexpr int THREADS = 1 << (CL - 1);
  constexpr uint32_t HALF = 1u << (CL - 1), CHUNK = 1u << CL;
  __shared__ uint32_t sm[2 * 8 * HALF + 16];
  __shared__ QM31Mat s_amat;  // alpha of the previous layer as a 4x4 matrix (kept out of the register file)
  const uint32_t tid = threadIdx.x;
  const uint32_t blob = blockIdx.x % p.n_blobs, chunk = blockIdx.x / p.n_blobs;
  Channel ch;
  if (tid == 0) {
    channel_init(ch);
    if (p.seeds) fs_mix_u64(ch, p.seeds[blob]);  // src/proof.rs:40-42
  }
#pragma unroll 1
  for (uint32_t layer = 0; layer < p.n_big; layer++) {
    const uint32_t d = p.D - layer;
    const uint32_t n_chunks = 1u << (d - CL);
    const size_t n = (size_t)1 << d;
    FsHash *tree = reinterpret_cast<FsHash *>(p.tree[layer]) + (size_t)blob * p.tree_stride[layer];
    const uint32_t leaf0 = chunk << CL;
    fs_stamp(p, layer, 0);
    if (p.trace && layer == 1 && tid == 0) {  // TEMP: per-CTA placement and timing of layer 1
      unsigned long long t; uint32_t smid;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.trace[256 + 3 * blockIdx.x] = smid; p.trace[256 + 3 * blockIdx.x + 1] = t;
    }
    // ---- leaves of my chunk (layers >= 1: fold of the previous layer first), one leaf-compression site
    {
      const uint32_t *s0 = p.cols[0] + (size_t)blob * p.cols_stride[0] + leaf0;
      const uint32_t lp = layer ? layer - 1 : 0;
      // pairs (2 i, 2 i + 1) of the previous layer, written during this kernel by the CTAs of chunks 2c and 2c + 1
      const uint2 *s = reinterpret_cast<const uint2 *>(p.cols[lp] + (size_t)blob * p.cols_stride[lp]) + leaf0;
      uint32_t *dcol = p.cols[layer] + (size_t)blob * p.cols_stride[layer] + leaf0;
      const uint32_t *iblk = layer == 1 ? p.tt.iblk(1u << (p.D - 2)) : p.tt.iblk(1u << d);
#pragma unroll 1
      for (uint32_t j = tid; j < CHUNK; j += THREADS) {
        uint32_t c0, c1, c2, c3;
        if (layer == 0) {
          c0 = __ldg(s0 + j);
          c1 = __ldg(s0 + n + j);
          c2 = __ldg(s0 + 2 * n + j);
          c3 = __ldg(s0 + 3 * n + j);
        } else {
          const uint2 e0 = __ldcg(s + j), e1 = __ldcg(s + n + j), e2 = __ldcg(s + 2 * n + j), e3 = __ldcg(s + 3 * n + j);
          const QM31 a = {{e0.x, e1.x, e2.x, e3.x}}, b = {{e0.y, e1.y, e2.y, e3.y}};
          const uint32_t itw = layer == 1 ? fs_circle_itw(iblk, leaf0 + j) : __ldg(iblk + leaf0 + j);
          const QM31 f = fri_fold_pair_mat(a, b, itw, s_amat);
          c0 = f.v[0];
          c1 = f.v[1];
          c2 = f.v[2];
          c3 = f.v[3];
          dcol[j] = c0;
          dcol[n + j] = c1;
          dcol[2 * n + j] = c2;
          dcol[3 * n + j] = c3;
        }
        uint32_t h[8];
        merkle_hash_leaf(c0, c1, c2, c3, h, 1u);
        fs_put<HALF>(sm, j, h);
        if (p.write_all) fs_store(tree + n + leaf0 + j, h);
      }
    }
    __syncthreads();
    fs_stamp(p, layer, 1);
    uint32_t top[8];
    fs_reduce<HALF, THREADS>(sm, CHUNK, CL, tree, d, leaf0, p.write_all != 0, true, top);
    fs_stamp(p, layer, 2);
    if (p.trace && layer == 1 && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.trace[256 + 3 * blockIdx.x + 2] = t;
    }
    // ---- publish my chunk root (and my slice of this layer's columns): release on the layer's counter
    __syncthreads();
    uint32_t *counter = p.counters + (size_t)layer * p.n_blobs + blob;
    if (tid == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
    // ---- who goes on: the CTAs that own a chunk of the next layer; after the last layer here, chunk 0 only
    const bool last_big = layer + 1 == p.n_big;
    if (last_big ? chunk != 0 : chunk >= (n_chunks >> 1)) return;
    if (tid == 0) {
      uint32_t seen;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        if (seen >= n_chunks) break;
        __nanosleep(20);
      }
    }
    __syncthreads();
    fs_stamp(p, layer, 3);
    // ---- top of the tree from the n_chunks chunk roots (every surviving CTA of the blob computes the same bytes)
    {
      const FsHash *src = tree + n_chunks;
      for (uint32_t j = tid; j < n_chunks; j += THREADS) {
        const uint4 lo = __ldcg(&src[j].lo), hi = __ldcg(&src[j].hi);
        const uint32_t h[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        fs_put<HALF>(sm, j, h);
        if (n_chunks == 1 && tid == 0) {
#pragma unroll
          for (int s = 0; s < 8; s++) top[s] = h[s];
        }
      }
      __syncthreads();
      fs_reduce<HALF, THREADS>(sm, n_chunks, d - CL, tree, d - CL, 0, p.write_all != 0 && chunk == 0, chunk == 0, top);
    }
    fs_stamp(p, layer, 4);
    // ---- Fiat-Shamir step of FriProver::commit for this layer: mix_root, draw the folding alpha
    if (tid == 0) {
      const QM31 a = fs_mix_root_draw_felt(ch, top);
      s_amat = qm31_mat(a);
      if (chunk == 0) {
        fs_store(reinterpret_cast<FsHash *>(p.roots + (size_t)blob * p.roots_stride + 32 * (size_t)layer), top);
        p.alpha[(size_t)blob * p.alpha_stride + layer] = a;
        if (last_big) p.chan[blob] = ch;
      }
    }
    __syncthreads();
    fs_stamp(p, layer, 5);
  }
}

}  // namespace

// Largest number of CTAs of fri_small_kernel<CL> that can be resident at once on the current device (0 on error).
static unsigned fs_capacity(int cl) {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return 0;
  cudaError_t e = cl == 9 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fri_small_kernel<9>, 256, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fri_small_kernel<10>, 512, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (unsigned)(sms * per_sm);
}

int fri_small_chunk_log(uint32_t D, uint32_t n_big, size_t n_blobs) {
  if (n_big == 0 || n_blobs == 0 || n_blobs > 64) return 0;
  static unsigned cap[2] = {fs_capacity(9), fs_capacity(10)};
  for (int cl = 9; cl <= 10; cl++) {
    // the smallest layer handled here has 2^(D - n_big + 1) >= 2^cl leaves; the top of layer 0 has 2^(D - cl) <= 2^cl
    // nodes (two per thread on its first level); every CTA is resident
    if (D - (n_big - 1) < (uint32_t)cl || D > 2u * (uint32_t)cl) continue;
    if ((n_blobs << (D - cl)) <= cap[cl - 9]) return cl;
  }
  return 0;
}

cudaError_t launch_fri_small(cudaStream_t st, const FriSmallParams &p, int chunk_log) {
  if (chunk_log != 9 && chunk_log != 10) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.counters, 0, (size_t)p.n_big * p.n_blobs * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  FriSmallParams q = p;
  q.one = 1u;
  void *args[] = {&q};
  const unsigned grid = (unsigned)(p.n_blobs << (p.D - chunk_log));
  if (chunk_log == 9) return cudaLaunchCooperativeKernel((const void *)fri_small_kernel<9>, dim3(grid), dim3(256), args, 0, st);
  return cudaLaunchCooperativeKernel((const void *)fri_small_kernel<10>, dim3(grid), dim3(512), args, 0, st);
}

}  // namespace frieda
