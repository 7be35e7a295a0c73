// merkle.cu -- Merkle commitment kernels (one BLAKE2s compression per thread, state in
// registers, level-by-level reduction in shared memory) with the FRI fold fused in front of
// the leaf hashing.
//
// Replaces MerkleProver::<CpuBackend, Blake2sMerkleHasher>::commit (src/commit.rs:17-22) and,
// for the FRI layers, FriOps::fold_circle_into_line / fold_line followed by
// FriInnerLayerProver::new's MerkleProver::commit (stwo fri.rs, reached from src/proof.rs:52-57).
//
// merkle_bottom_kernel<SRC>: each CTA owns 2^chunk_log consecutive leaves of one blob.
//   SRC_COLS         leaf i = H(col0[i], col1[i], col2[i], col3[i])
//   SRC_FOLD_CIRCLE  first folds the pair (2i, 2i+1) of the circle evaluation with alpha
//                    (f0 + alpha f1, f1 = (a - b) / y), stores the new line layer, hashes it
//   SRC_FOLD_LINE    same with the line twiddle 1/x
//   SRC_NODES        leaves are 2^log existing nodes of tree level src_level (middle pass)
// then reduces `levels` tree levels in shared memory and writes the surviving tops (and, when
// write_all, every level on the way) into the heap-ordered tree.
#include "kernels.cuh"

namespace frieda {

constexpr int MB_THREADS = 256;      // throughput form: batches, big trees, and 512+ leaves per CTA
constexpr int MB_THREADS_LAT = 128;  // latency form (see MbOccupancy)
constexpr uint32_t MB_CHUNK_LOG_MAX = 10;  // 1024 leaves -> 32 KiB of shared memory (levels are reduced in place)

struct alignas(16) Hash32 {
  uint4 lo, hi;
};

__device__ __forceinline__ void store_hash(Hash32 *dst, const uint32_t h[8]) {
  dst->lo = make_uint4(h[0], h[1], h[2], h[3]);
  dst->hi = make_uint4(h[4], h[5], h[6], h[7]);
}
__device__ __forceinline__ void load_pair(const Hash32 *src, uint32_t m[16]) {
  uint4 a = src[0].lo, b = src[0].hi, c = src[1].lo, d = src[1].hi;
  m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
  m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
  m[8] = c.x; m[9] = c.y; m[10] = c.z; m[11] = c.w;
  m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
}

// inverse twiddle of the circle->line fold for pair i: from the (x, y)
// pairs of the largest line block as [1/y, -1/y, -1/x, 1/x].
__device__ __forceinline__ uint32_t circle_fold_itw(const uint32_t *iblk, size_t i) {
  size_t q = i >> 2;
  uint32_t e = (uint32_t)(i & 3);
  uint32_t v = __ldg(iblk + 2 * q + (e < 2 ? 1 : 0));
  return (e == 1 || e == 2) ? m31_neg(v) : v;
}

// ---------------------------------------------------------------- the pass
// The 16 message words of a node compression are not loaded into registers: they are fetched
// from shared memory at the point of use (10 rounds x 16 words = 160 LDS per compression; the LSU pipe is idle in
// these kernels), which frees ~16 registers and lets more CTAs stay resident.  The level buffer is a structure of
// arrays split by node parity -- word s of node n at E/O[n & 1][s][n >> 1], the odd half shifted by 16 banks -- so
// that "word s of my left / right child" and "word s of my own digest" are both conflict-free across a warp.
// Levels are reduced in place: all threads finish hashing a level before anyone overwrites it.
// Resident CTAs per SM (measured, one FRI-commit wave of 1024 blobs): leaves-from-columns 20.09 / 19.50 / 18.84 ms
// at 4 / 5 / 6; the fold variants 11.1 / 10.3 / 10.6 ms (the fold's 4x4 matrix wants the registers).
// Latency form (T = 128, chunks of <= 256 leaves): one blob (or a few) has the GPU to itself, so a tree is a chain of
// dependent compressions (0.85 - 0.93 us per link with one warp per scheduler, 1.76 us with two:
// bench_micro/chain.cu), not a throughput problem.  The host picks chunks small enough that about one CTA lands on
// every SM (ctx.cu: chunk_log_for); ptxas places the adds itself (8 % faster per dependent compression than the
// all-IMAD form) and is not held to a register budget.  Measured on one 128 KiB blob (FRI commit, 9 trees):
// 0.489 -> 0.439 ms.  Letting the CTA that finishes last (a ticket per blob) also reduce the tops and run the
// transcript step, instead of launching merkle_top_kernel, was built and measured: no gain (the fences and the
// ticket cost what the launch gap does), so the top stays a kernel of its own.
template <int SRC, int T>
struct MbOccupancy {
  static constexpr int min_blocks = T == MB_THREADS_LAT ? 1 : (SRC == SRC_COLS ? 6 : 5);
};
constexpr uint32_t MB2_HALF = 1u << (MB_CHUNK_LOG_MAX - 1);
__device__ __forceinline__ uint32_t *mb2_word(uint32_t *sm, uint32_t node, uint32_t s) {
  const uint32_t par = node & 1u;
  return sm + (par * 8 + s) * MB2_HALF + (node >> 1) + par * 16;
}
struct Mb2PairMsg {
  const uint32_t *e, *o;  // word 0 of the left (even) and of the right (odd) child
  __device__ __forceinline__ uint32_t operator[](int s) const { return s < 8 ? e[s * MB2_HALF] : o[(s - 8) * MB2_HALF]; }
};
__device__ __forceinline__ void mb2_put(uint32_t *sm, uint32_t node, const uint32_t h[8]) {
#pragma unroll
  for (int s = 0; s < 8; s++) *mb2_word(sm, node, s) = h[s];
}

// Reduces `levels` levels of the cnt nodes held in the level buffer (in place); the nodes sit on tree level `level`
// at index idx0 + j.  Stores every level (write_all) or only the last one into the tree.
template <int T>
__device__ __forceinline__ void mb2_reduce(uint32_t *sm, uint32_t cnt, uint32_t levels, uint32_t level, size_t idx0,
                                           Hash32 *tree, int write_all, uint32_t one) {
  for (uint32_t l = 0; l < levels; l++) {
    cnt >>= 1;
    level -= 1;
    idx0 >>= 1;
    const bool top = (l + 1 == levels);
    Hash32 *out = tree + ((size_t)1 << level) + idx0;
    if (T == MB_THREADS) {
      // cnt <= 512: at most two nodes per thread, both digests held across ONE barrier (every compression of the
      // level has read its children before anyone overwrites them).  Holding only the first and hashing the second
      // after the barrier removes the spills of this form but measured 4.5 % slower end to end.
      const uint32_t j0 = threadIdx.x, j1 = threadIdx.x + T;
      uint32_t h0[8], h1[8];
      if (j0 < cnt) merkle_hash_node_msg(Mb2PairMsg{sm + j0, sm + 8 * MB2_HALF + 16 + j0}, h0, one);
      if (j1 < cnt) merkle_hash_node_msg(Mb2PairMsg{sm + j1, sm + 8 * MB2_HALF + 16 + j1}, h1, one);
      __syncthreads();
      if (j0 < cnt) {
        if (!top) mb2_put(sm, j0, h0);
        if (write_all || top) store_hash(out + j0, h0);
      }
      if (j1 < cnt) {
        if (!top) mb2_put(sm, j1, h1);
        if (write_all || top) store_hash(out + j1, h1);
      }
    } else {
      // rounds of T nodes: node j reads slots j of the even / odd halves and is stored to slot j >> 1.  An EVEN round
      // holds its digest (one per thread) across a barrier, because its stores land on slots that the round itself
      // (round 0) or the previous round still reads; an ODD round r stores at once: its slots [rT/2, (r+1)T/2) were
      // read in rounds < r, which every warp has left at the last barrier.
      for (uint32_t r = 0, j0 = 0; j0 < cnt; r++, j0 += T) {
        const uint32_t j = j0 + threadIdx.x;
        uint32_t h[8];
        if (j < cnt) merkle_hash_node_msg(Mb2PairMsg{sm + j, sm + 8 * MB2_HALF + 16 + j}, h, one);
        if ((r & 1u) == 0) __syncthreads();
        if (j < cnt) {
          if (!top) mb2_put(sm, j, h);
          if (write_all || top) store_hash(out + j, h);
        }
      }
    }
    if (!top) __syncthreads();
  }
}

template <int SRC, int T>
__global__ void __launch_bounds__(T, MbOccupancy<SRC, T>::min_blocks) merkle_bottom_kernel(const MerkleBottomParams p) {
  __shared__ uint32_t sm[2 * 8 * MB2_HALF + 16];
  const size_t blob = blockIdx.y;
  const uint32_t chunk = blockIdx.x;
  const uint32_t n_chunk = 1u << p.chunk_log;
  const size_t leaf0 = (size_t)chunk << p.chunk_log;
  Hash32 *tree = reinterpret_cast<Hash32 *>(p.tree) + blob * p.tree_stride;
  // the throughput form forces the adds onto the FMA pipe through the opaque runtime 1 (blake2s.cuh)
  const uint32_t one = T == MB_THREADS_LAT ? 1u : p.one;

  if (SRC == SRC_NODES) {
    const Hash32 *src = tree + ((size_t)1 << p.src_level) + leaf0;
    for (uint32_t j = threadIdx.x; j < n_chunk; j += T) {
      const Hash32 v = src[j];
      const uint32_t h[8] = {v.lo.x, v.lo.y, v.lo.z, v.lo.w, v.hi.x, v.hi.y, v.hi.z, v.hi.w};
      mb2_put(sm, j, h);
    }
  } else {
    const size_t n = (size_t)1 << p.log;
    QM31Mat amat;
    if (SRC == SRC_FOLD_CIRCLE || SRC == SRC_FOLD_LINE) amat = qm31_mat(p.alpha[blob * p.alpha_stride]);
    for (uint32_t j = threadIdx.x; j < n_chunk; j += T) {
      const size_t i = leaf0 + j;
      uint32_t c0, c1, c2, c3;
      if (SRC == SRC_COLS) {
        const uint32_t *s = p.src_cols + blob * p.src_stride + i;
        c0 = __ldg(s);
        c1 = __ldg(s + n);
        c2 = __ldg(s + 2 * n);
        c3 = __ldg(s + 3 * n);
      } else {
        const uint2 *s = reinterpret_cast<const uint2 *>(p.src_cols + blob * p.src_stride) + i;
        uint2 e0 = __ldg(s), e1 = __ldg(s + n), e2 = __ldg(s + 2 * n), e3 = __ldg(s + 3 * n);
        QM31 a = {{e0.x, e1.x, e2.x, e3.x}}, b = {{e0.y, e1.y, e2.y, e3.y}};
        uint32_t itw = SRC == SRC_FOLD_CIRCLE ? circle_fold_itw(p.itw_blk, i) : __ldg(p.itw_blk + i);
        QM31 f = fri_fold_pair_mat(a, b, itw, amat);
        c0 = f.v[0];
        c1 = f.v[1];
        c2 = f.v[2];
        c3 = f.v[3];
        uint32_t *d = p.dst_cols + blob * p.dst_stride + i;
        d[0] = c0;
        d[n] = c1;
        d[2 * n] = c2;
        d[3 * n] = c3;
      }
      uint32_t h[8];
      merkle_hash_leaf(c0, c1, c2, c3, h, one);
      mb2_put(sm, j, h);
      if (p.write_all) store_hash(tree + n + i, h);
    }
  }
  __syncthreads();
  const uint32_t level = (SRC == SRC_NODES ? p.src_level : p.log);
  mb2_reduce<T>(sm, n_chunk, p.levels, level, leaf0, tree, p.write_all, one);
  if (p.levels == 0 && !p.write_all && SRC != SRC_NODES) {
    const size_t n = (size_t)1 << p.log;
    for (uint32_t j = threadIdx.x; j < n_chunk; j += T) {
      uint32_t h[8];
#pragma unroll
      for (int s = 0; s < 8; s++) h[s] = *mb2_word(sm, j, s);
      store_hash(tree + n + leaf0 + j, h);
    }
  }
}

cudaError_t launch_merkle_bottom(cudaStream_t st, int src, const MerkleBottomParams &p, size_t n_blobs) {
  if (p.chunk_log > MB_CHUNK_LOG_MAX || p.levels > p.chunk_log || p.chunk_log > p.log) return cudaErrorInvalidValue;
  unsigned chunks = 1u << (p.log - p.chunk_log);
  for (size_t b0 = 0; b0 < n_blobs; b0 += 32768) {
    size_t nb = n_blobs - b0 < 32768 ? n_blobs - b0 : 32768;
    MerkleBottomParams q = p;
    q.one = 1u;
    if (q.src_cols) q.src_cols += b0 * p.src_stride;
    if (q.dst_cols) q.dst_cols += b0 * p.dst_stride;
    q.tree += b0 * p.tree_stride * 32;
    if (q.alpha) q.alpha += b0 * p.alpha_stride;
    dim3 grid(chunks, (unsigned)nb);
    if (p.latency) {
      constexpr int T = MB_THREADS_LAT;
      switch (src) {
        case SRC_COLS: merkle_bottom_kernel<SRC_COLS, T><<<grid, T, 0, st>>>(q); break;
        case SRC_FOLD_CIRCLE: merkle_bottom_kernel<SRC_FOLD_CIRCLE, T><<<grid, T, 0, st>>>(q); break;
        case SRC_FOLD_LINE: merkle_bottom_kernel<SRC_FOLD_LINE, T><<<grid, T, 0, st>>>(q); break;
        case SRC_NODES: merkle_bottom_kernel<SRC_NODES, T><<<grid, T, 0, st>>>(q); break;
        default: return cudaErrorInvalidValue;
      }
    } else {
      constexpr int T = MB_THREADS;
      switch (src) {
        case SRC_COLS: merkle_bottom_kernel<SRC_COLS, T><<<grid, T, 0, st>>>(q); break;
        case SRC_FOLD_CIRCLE: merkle_bottom_kernel<SRC_FOLD_CIRCLE, T><<<grid, T, 0, st>>>(q); break;
        case SRC_FOLD_LINE: merkle_bottom_kernel<SRC_FOLD_LINE, T><<<grid, T, 0, st>>>(q); break;
        case SRC_NODES: merkle_bottom_kernel<SRC_NODES, T><<<grid, T, 0, st>>>(q); break;
        default: return cudaErrorInvalidValue;
      }
    }
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- top of the tree + channel
// One CTA per blob.  Reduces the 2^top_log nodes of level top_log to the root, stores the root,
// then (chan != nullptr) runs the Fiat-Shamir step of FriProver::commit for this layer:
// MC::mix_root(channel, root); folding_alpha = channel.draw_felt().
constexpr int MT_THREADS = 128;
constexpr uint32_t MT_TOP_LOG_MAX = 10;

__global__ void __launch_bounds__(MT_THREADS) merkle_top_kernel(uint8_t *tree_, size_t tree_stride, uint32_t top_log,
                                                                 int write_all, uint8_t *roots, size_t roots_stride,
                                                                 Channel *chan, QM31 *alpha, size_t alpha_stride,
                                                                 uint32_t one) {
  __shared__ Hash32 sm_a[1u << MT_TOP_LOG_MAX];
  __shared__ Hash32 sm_b[1u << (MT_TOP_LOG_MAX - 1)];
  const size_t blob = blockIdx.x;
  Hash32 *tree = reinterpret_cast<Hash32 *>(tree_) + blob * tree_stride;
  uint32_t cnt = 1u << top_log;
  for (uint32_t j = threadIdx.x; j < cnt; j += MT_THREADS) sm_a[j] = tree[cnt + j];
  __syncthreads();
  Hash32 *cur = sm_a, *nxt = sm_b;
  for (uint32_t level = top_log; level > 0; level--) {
    cnt >>= 1;
    for (uint32_t j = threadIdx.x; j < cnt; j += MT_THREADS) {
      uint32_t m[16], h[8];
      load_pair(cur + 2 * j, m);
      // latency-bound here (one CTA per blob, a few warps): the literal 1 lets ptxas place the adds itself, which is
      // 8 % faster per dependent compression than the all-IMAD form of the throughput kernels (bench_micro/chain.cu)
      (void)one;
      merkle_hash_node(m, h, 1u);
      store_hash(&nxt[j], h);
      if (write_all || level == 1) store_hash(tree + cnt + j, h);
    }
    __syncthreads();
    Hash32 *t = cur;
    cur = nxt;
    nxt = t;
  }
  if (threadIdx.x == 0) {
    Hash32 r = cur[0];
    uint32_t root[8] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, r.hi.z, r.hi.w};
    if (roots) store_hash(reinterpret_cast<Hash32 *>(roots + blob * roots_stride), root);
    if (chan) {
      Channel c = chan[blob];
      channel_mix_root(c, root);
      QM31 a = channel_draw_felt(c);
      chan[blob] = c;
      alpha[blob * alpha_stride] = a;
    }
  }
}

cudaError_t launch_merkle_top(cudaStream_t st, uint8_t *tree, size_t tree_stride, uint32_t top_log, int write_all,
                              uint8_t *roots, size_t roots_stride, Channel *chan, QM31 *alpha, size_t alpha_stride,
                              size_t n_blobs) {
  if (top_log > MT_TOP_LOG_MAX) return cudaErrorInvalidValue;
  merkle_top_kernel<<<(unsigned)n_blobs, MT_THREADS, 0, st>>>(tree, tree_stride, top_log, write_all, roots,
                                                              roots_stride, chan, alpha, alpha_stride, 1u);
  return cudaGetLastError();
}

__global__ void channel_init_kernel(Channel *chan, const uint64_t *seeds, size_t n) {
  size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  Channel c;
  channel_init(c);
  if (seeds) channel_mix_u64(c, seeds[b]);  // src/proof.rs:40-42
  chan[b] = c;
}
cudaError_t launch_channel_init(cudaStream_t st, Channel *chan, const uint64_t *seeds, size_t n_blobs) {
  channel_init_kernel<<<(unsigned)((n_blobs + 127) / 128), 128, 0, st>>>(chan, seeds, n_blobs);
  return cudaGetLastError();
}

}  // namespace frieda
