// fri.cu -- standalone FRI fold pass and the per-blob tail kernel.
//
// fold_kernel: FriOps::fold_circle_into_line / fold_line (stwo fri.rs; reached from
// src/proof.rs:52-57) as a pure streaming pass: read 2 secure points, write 1.
//
// fri_tail_kernel: once a layer has at most 2^TAIL_LOG points the rest of
// FriProver::commit runs inside ONE CTA per blob: per layer Merkle tree, mix_root,
// draw alpha, fold; then the last layer's interpolation (LineEvaluation::interpolate), the
// degree check ("invalid degree"), and mix_felts(last_layer_poly).
#include "kernels.cuh"

namespace frieda {

__device__ __forceinline__ uint32_t circle_fold_itw_(const uint32_t *iblk, size_t i) {
  size_t q = i >> 2;
  uint32_t e = (uint32_t)(i & 3);
  uint32_t v = __ldg(iblk + 2 * q + (e < 2 ? 1 : 0));
  return (e == 1 || e == 2) ? m31_neg(v) : v;
}

__global__ void __launch_bounds__(256) fold_kernel(const uint32_t *__restrict__ src, size_t src_stride,
                                                    uint32_t dst_log, int is_circle, const QM31 *__restrict__ alpha,
                                                    size_t alpha_stride, const uint32_t *__restrict__ iblk,
                                                    uint32_t *__restrict__ dst, size_t dst_stride) {
  const size_t blob = blockIdx.y;
  const size_t n = (size_t)1 << dst_log;
  const QM31Mat amat = qm31_mat(alpha[blob * alpha_stride]);
  const uint2 *s = reinterpret_cast<const uint2 *>(src + blob * src_stride);
  uint32_t *d = dst + blob * dst_stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint2 e0 = __ldg(s + i), e1 = __ldg(s + n + i), e2 = __ldg(s + 2 * n + i), e3 = __ldg(s + 3 * n + i);
    QM31 a = {{e0.x, e1.x, e2.x, e3.x}}, b = {{e0.y, e1.y, e2.y, e3.y}};
    uint32_t itw = is_circle ? circle_fold_itw_(iblk, i) : __ldg(iblk + i);
    QM31 f = fri_fold_pair_mat(a, b, itw, amat);
    d[i] = f.v[0];
    d[n + i] = f.v[1];
    d[2 * n + i] = f.v[2];
    d[3 * n + i] = f.v[3];
  }
}

// Fold of the slice [lo, lo + 2^local_src_log) of a layer with 2^src_log points (split blob: a rank folds its own
// index range; pairs (2i, 2i+1) never cross a rank boundary).  lo is a multiple of 2^local_src_log >= 8.
cudaError_t launch_fold_range(cudaStream_t st, const uint32_t *src_local, uint32_t src_log, uint32_t local_src_log,
                              size_t lo, int is_circle, const QM31 *alpha, const TwiddleTable &tt, uint32_t *dst_local) {
  if (local_src_log < 3 || local_src_log > src_log || (lo & (((size_t)1 << local_src_log) - 1))) return cudaErrorInvalidValue;
  if (is_circle && src_log < 3) return cudaErrorInvalidValue;
  // pair i of the whole layer uses iblk[i] (line) or the (x, y) pair iblk[2 (i >> 2) ..] (circle): offset by lo / 2 pairs
  const size_t pair0 = lo >> 1;
  const uint32_t *iblk = is_circle ? tt.iblk(1u << (src_log - 2)) + 2 * (pair0 >> 2) : tt.iblk(1u << (src_log - 1)) + pair0;
  const uint32_t dst_log = local_src_log - 1;
  const size_t n = (size_t)1 << dst_log;
  unsigned bx = (unsigned)((n + 255) / 256);
  if (bx > 4096) bx = 4096;
  fold_kernel<<<dim3(bx, 1), 256, 0, st>>>(src_local, (size_t)4 << local_src_log, dst_log, is_circle, alpha, 1, iblk, dst_local,
                                           (size_t)4 << dst_log);
  return cudaGetLastError();
}

cudaError_t launch_fold(cudaStream_t st, const uint32_t *src, size_t src_stride, uint32_t src_log, int is_circle,
                        const QM31 *alpha, size_t alpha_stride, const TwiddleTable &tt, uint32_t *dst,
                        size_t dst_stride, size_t n_blobs) {
  if (src_log == 0) return cudaErrorInvalidValue;
  const uint32_t dst_log = src_log - 1;
  // circle fold of a log-D evaluation uses the largest line block of half_odds(D-1) (2^(D-2) pairs);
  // line fold of a log-k layer uses the block of length 2^(k-1).
  const uint32_t *iblk;
  CPoint dummy;
  (void)dummy;
  if (is_circle) {
    if (src_log < 3) return cudaErrorInvalidValue;  // handled by the caller (tiny domains)
    iblk = tt.iblk(1u << (src_log - 2));
  } else {
    iblk = tt.iblk(1u << (src_log - 1));
  }
  size_t n = (size_t)1 << dst_log;
  unsigned bx = (unsigned)((n + 255) / 256);
  if (bx > 4096) bx = 4096;
  for (size_t b0 = 0; b0 < n_blobs; b0 += 32768) {
    size_t nb = n_blobs - b0 < 32768 ? n_blobs - b0 : 32768;
    fold_kernel<<<dim3(bx, (unsigned)nb), 256, 0, st>>>(src + b0 * src_stride, src_stride, dst_log, is_circle,
                                                        alpha + b0 * alpha_stride, alpha_stride, iblk,
                                                        dst + b0 * dst_stride, dst_stride);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- tail
constexpr int TAIL_THREADS = 256;

struct alignas(16) THash {
  uint4 lo, hi;
};
__device__ __forceinline__ void t_store(THash *dst, const uint32_t h[8]) {
  dst->lo = make_uint4(h[0], h[1], h[2], h[3]);
  dst->hi = make_uint4(h[4], h[5], h[6], h[7]);
}
__device__ __forceinline__ void t_load_pair(const THash *src, uint32_t m[16]) {
  uint4 a = src[0].lo, b = src[0].hi, c = src[1].lo, d = src[1].hi;
  m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
  m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
  m[8] = c.x; m[9] = c.y; m[10] = c.z; m[11] = c.w;
  m[12] = d.x; m[13] = d.y; m[14] = d.z; m[15] = d.w;
}

__global__ void __launch_bounds__(TAIL_THREADS) fri_tail_kernel(const __grid_constant__ TailParams p) {
  constexpr uint32_t NMAX = 1u << TAIL_LOG;
  __shared__ uint32_t s_cols[2][4][NMAX];  // ping-pong layer columns
  __shared__ THash s_ha[NMAX];
  __shared__ THash s_hb[NMAX / 2];
  __shared__ QM31 s_alpha;
  const size_t blob = blockIdx.x;
  const uint32_t tid = threadIdx.x;
  uint32_t layer = p.start_layer, log = p.start_log;
  int cur = 0;
  {
    const uint32_t n = 1u << log;
    const uint32_t *src = p.cols[layer] + blob * p.cols_stride[layer];
    for (uint32_t e = tid; e < 4 * n; e += TAIL_THREADS) s_cols[0][e >> log][e & (n - 1)] = src[e];
  }
  __syncthreads();
  Channel ch;
  if (tid == 0) ch = p.chan[blob];
  // commit + fold until the evaluation reaches the last-layer domain size (layer 0 is always
  // committed: the caller guarantees start_log - 1 >= last_log when start_layer == 0)
  while (layer == 0 || log > p.last_log) {
    const uint32_t n = 1u << log;
    THash *tree = reinterpret_cast<THash *>(p.tree[layer]) + blob * p.tree_stride[layer];
    // leaves
    for (uint32_t j = tid; j < n; j += TAIL_THREADS) {
      uint32_t h[8];
      merkle_hash_leaf(s_cols[cur][0][j], s_cols[cur][1][j], s_cols[cur][2][j], s_cols[cur][3][j], h, 1u);  // latency-bound: see merkle_top_kernel
      t_store(&s_ha[j], h);
      if (p.write_all || log == 0) t_store(tree + n + j, h);
    }
    __syncthreads();
    THash *hc = s_ha, *hn = s_hb;
    uint32_t cnt = n;
    for (uint32_t level = log; level > 0; level--) {
      cnt >>= 1;
      for (uint32_t j = tid; j < cnt; j += TAIL_THREADS) {
        uint32_t m[16], h[8];
        t_load_pair(hc + 2 * j, m);
        merkle_hash_node(m, h, 1u);
        t_store(&hn[j], h);
        if (p.write_all || level == 1) t_store(tree + cnt + j, h);
      }
      __syncthreads();
      THash *t = hc;
      hc = hn;
      hn = t;
    }
    if (tid == 0) {
      THash r = hc[0];
      uint32_t root[8] = {r.lo.x, r.lo.y, r.lo.z, r.lo.w, r.hi.x, r.hi.y, r.hi.z, r.hi.w};
      t_store(reinterpret_cast<THash *>(p.roots + blob * p.roots_stride + 32 * (size_t)layer), root);
      channel_mix_root(ch, root);
      QM31 a = channel_draw_felt(ch);
      p.alpha[blob * p.alpha_stride + layer] = a;
      s_alpha = a;
    }
    __syncthreads();
    const QM31Mat amat = qm31_mat(s_alpha);
    // fold into the next layer
    const uint32_t nn = n >> 1;
    uint32_t *dst = p.cols[layer + 1] + blob * p.cols_stride[layer + 1];
    // (circle domains below 8 points are rejected by the host: make_geom_fri)
    const uint32_t *iblk = layer == 0 ? p.tt.iblk(1u << (log - 2)) : p.tt.iblk(1u << (log - 1));
    for (uint32_t i = tid; i < nn; i += TAIL_THREADS) {
      QM31 a = {{s_cols[cur][0][2 * i], s_cols[cur][1][2 * i], s_cols[cur][2][2 * i], s_cols[cur][3][2 * i]}};
      QM31 b = {{s_cols[cur][0][2 * i + 1], s_cols[cur][1][2 * i + 1], s_cols[cur][2][2 * i + 1],
                 s_cols[cur][3][2 * i + 1]}};
      uint32_t itw = layer == 0 ? circle_fold_itw_(iblk, i) : __ldg(iblk + i);
      QM31 f = fri_fold_pair_mat(a, b, itw, amat);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        s_cols[cur ^ 1][c][i] = f.v[c];
        dst[((size_t)c << (log - 1)) + i] = f.v[c];
      }
    }
    __syncthreads();
    cur ^= 1;
    layer += 1;
    log -= 1;
  }
  // ---- last layer: interpolate the 2^last_log evaluations (bit-reversed order) ----
  // reuse the hash buffer as QM31 storage
  QM31 *vals = reinterpret_cast<QM31 *>(s_ha);
  const uint32_t n = 1u << log;
  for (uint32_t i = tid; i < n; i += TAIL_THREADS) {
    // bit_reverse_column: natural[i] = stored[brev(i)]
    uint32_t j = bit_reverse(i, log);
    vals[i] = {{s_cols[cur][0][j], s_cols[cur][1][j], s_cols[cur][2][j], s_cols[cur][3][j]}};
  }
  __syncthreads();
  // line_ifft: domain half_odds(dlog); x_i = natural-order at(i).x = tw[blk(size/2) + brev(i)]
  for (uint32_t dlog = log; dlog >= 1; dlog--) {
    const uint32_t size = 1u << dlog, hs = size >> 1;
    const uint32_t *iblk = p.tt.iblk(hs);
    for (uint32_t bf = tid; bf < (n >> 1); bf += TAIL_THREADS) {
      uint32_t chunk = bf >> (dlog - 1), i = bf & (hs - 1);
      QM31 *l = vals + (size_t)chunk * size + i, *r = l + hs;
      uint32_t itw = __ldg(iblk + bit_reverse(i, dlog - 1));
      QM31 a = *l, b = *r;
      *l = qm31_add(a, b);
      *r = qm31_mul_m31(qm31_sub(a, b), itw);
    }
    __syncthreads();
  }
  // scale by 1/n; coefficients are now in LinePoly storage (bit-reversed) order.
  // ordered[j] = storage[brev(j, log)]; require ordered[j] == 0 for j >= 2^log_last;
  // last_layer_poly (storage order of the truncated poly)[i] = ordered[brev(i, log_last)].
  const uint32_t bound = 1u << p.log_last;
  for (uint32_t j = tid; j < n; j += TAIL_THREADS) {
    QM31 v = qm31_mul_m31(vals[bit_reverse(j, log)], p.inv_last_n);
    if (j >= bound) {
      if (!qm31_is_zero(v)) atomicExch(p.error_flag, 1);
    } else {
      p.last_poly[blob * bound + bit_reverse(j, p.log_last)] = v;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // mix_felts(last_layer_poly); the coefficients were just written by this CTA
    channel_mix_felts(ch, p.last_poly + blob * bound, bound);
    p.chan[blob] = ch;
  }
}

cudaError_t launch_tail(cudaStream_t st, const TailParams &p, size_t n_blobs) {
  if (p.start_log > TAIL_LOG || p.last_log > TAIL_LAST_MAX) return cudaErrorInvalidValue;
  for (size_t b0 = 0; b0 < n_blobs; b0 += 65535) {
    size_t nb = n_blobs - b0 < 65535 ? n_blobs - b0 : 65535;
    TailParams q = p;
    q.one = 1u;
    for (int l = 0; l < 32; l++) {
      if (q.cols[l]) q.cols[l] += b0 * p.cols_stride[l];
      if (q.tree[l]) q.tree[l] += b0 * p.tree_stride[l] * 32;
    }
    q.roots += b0 * p.roots_stride;
    q.chan += b0;
    q.alpha += b0 * p.alpha_stride;
    q.last_poly += b0 * ((size_t)1 << p.log_last);
    fri_tail_kernel<<<(unsigned)nb, TAIL_THREADS, 0, st>>>(q);
  }
  return cudaGetLastError();
}

}  // namespace frieda
