// lde.cu -- byte packing, twiddle tree and the circle-FFT low-degree extension.
//
// Replaces, for the path of src/commit.rs:12-16 and src/proof.rs:38,44-50:
//   utils::bytes_to_felt_le / polynomial_from_bytes   (src/utils.rs:10-33)   -> pack_kernel
//   CpuBackend::precompute_twiddles(half_odds(K))     (src/commit.rs:15)     -> twiddle_kernel
//   SecureCirclePoly::evaluate_with_twiddles          (src/commit.rs:16)     -> lde_warp_kernel (blocks of 2^10 ..
//                                                        2^15 points), lde_block_kernel (smaller), lde_strided_r16_kernel
//                                                        (the layers above 2^15-point blocks)
//
// LDE structure: coefficients are zero-padded at the high indices and the
// circle FFT runs its largest strides first, so the first log_blowup layers only replicate the
// coefficient vector: the extension is 2^log_blowup independent FFTs of 2^poly_log points, block
// hb using the twiddles whose index has hb as its high bits.  Each (blob, column, block) is one
// CTA working in shared memory; HBM sees the coefficients once (L2 for re-reads) and the
// evaluations once.
#include <cstdlib>

#include "kernels.cuh"

namespace frieda {

// ---------------------------------------------------------------- pack
// Input viewed as one little-endian bit string cut into 30-bit limbs (src/utils.rs:10-19),
// zero-padded to 4 * 2^poly_log coefficients (src/utils.rs:23-24).
__device__ __forceinline__ uint32_t load_word_le(const uint8_t *base, size_t len, size_t w, bool aligned) {
  size_t byte = 4 * w;
  if (aligned && byte + 4 <= len) return *reinterpret_cast<const uint32_t *>(base + byte);
  uint32_t v = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (byte + j < len) v |= (uint32_t)base[byte + j] << (8 * j);
  return v;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ blobs, size_t len, size_t stride,
                                                    uint32_t n_felts, uint32_t n_coef, uint32_t *__restrict__ coef) {
  const size_t blob = blockIdx.y;
  const uint8_t *base = blobs + blob * stride;
  const bool aligned = ((reinterpret_cast<uintptr_t>(base)) & 3) == 0;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_coef; k += gridDim.x * blockDim.x) {
    uint32_t val = 0;
    if (k < n_felts) {
      uint64_t bit = 30ull * k;
      size_t w = (size_t)(bit >> 5);
      uint32_t off = (uint32_t)(bit & 31);
      uint32_t w0 = load_word_le(base, len, w, aligned);
      uint32_t w1 = off > 2 ? load_word_le(base, len, w + 1, aligned) : 0u;
      val = __funnelshift_r(w0, w1, off) & 0x3fffffffu;
    }
    coef[blob * n_coef + k] = val;
  }
}

constexpr uint32_t PP_LIMBS = 4096, PP_BYTES = 15360;  // 4 x (1024 limbs = 3840 bytes) per CTA iteration
struct PackWait {  // pipelined upload (frieda_commit_split_peers): part k of slice s is ready once flag (2 + k, s) >= epoch
  const uint32_t *my_flags;  // this rank's flag array, or nullptr: no waiting
  uint32_t part_len, epoch;
  int *timeout_flag;
};
__global__ void pack_peers_kernel(const __grid_constant__ PeerPtrs sl, uint32_t slice_len, uint32_t len, uint32_t n_felts,
                                  uint32_t n_coef, uint32_t first_chunk, size_t blob_stride, uint32_t *__restrict__ coef_all,
                                  PackWait pw);

cudaError_t launch_pack(cudaStream_t st, const uint8_t *blobs, size_t len, size_t stride, size_t n_blobs,
                        uint32_t n_felts, uint32_t poly_log, uint32_t *coef) {
  uint32_t n_coef = 4u << poly_log;
  // 16-byte aligned batches of at least 4096 coefficients: the staged 128-bit kernel (the whole blob is one "slice")
  if (n_coef >= PP_LIMBS && len > 0 && len < ((size_t)1 << 31) && (reinterpret_cast<uintptr_t>(blobs) & 15) == 0 &&
      (stride & 15) == 0) {
    PeerPtrs one;
    for (uint32_t r = 0; r < MAX_PEERS; r++) one.p[r] = blobs;
    const uint32_t slice = (uint32_t)((len + 15) / 16 * 16);
    for (size_t b0 = 0; b0 < n_blobs; b0 += 65535) {
      size_t nb = n_blobs - b0 < 65535 ? n_blobs - b0 : 65535;
      one.p[0] = blobs + b0 * stride;
      pack_peers_kernel<<<dim3(n_coef / PP_LIMBS, (unsigned)nb), 256, 0, st>>>(one, slice, (uint32_t)len, n_felts, n_coef, 0,
                                                                              stride, coef + b0 * n_coef, PackWait{});
    }
    return cudaGetLastError();
  }
  uint32_t bx = (n_coef + 255) / 256;
  if (bx > 1024) bx = 1024;
  for (size_t b0 = 0; b0 < n_blobs; b0 += 65535) {
    size_t nb = n_blobs - b0 < 65535 ? n_blobs - b0 : 65535;
    pack_kernel<<<dim3(bx, (unsigned)nb), 256, 0, st>>>(blobs + b0 * stride, len, stride, n_felts, n_coef,
                                                        coef + b0 * n_coef);
  }
  return cudaGetLastError();
}

// One blob in `world` slices behind separate (peer-mapped) pointers.  1024 limbs are exactly 960 words =
// 240 x 16 bytes, so a CTA stages one such chunk in shared memory with 240 coalesced 128-bit loads (slices are
// multiples of 16 bytes: a 16-byte piece never straddles two) -- over NVLink that is the access pattern of a
// bulk copy, where 4-byte loads reached a quarter of the link rate -- and cuts it into limbs from there.
// blockIdx.y = blob of a batch (blob_stride bytes apart behind every slice pointer; a split blob is a batch of one).
__global__ void __launch_bounds__(256) pack_peers_kernel(const __grid_constant__ PeerPtrs sl, uint32_t slice_len,
                                                          uint32_t len, uint32_t n_felts, uint32_t n_coef,
                                                          uint32_t first_chunk, size_t blob_stride,
                                                          uint32_t *__restrict__ coef_all, PackWait pw) {
  __shared__ __align__(16) uint32_t w[PP_BYTES / 4 + 4];
  const uint32_t t = threadIdx.x, n_chunks = n_coef / PP_LIMBS;
  const size_t in_off = (size_t)blockIdx.y * blob_stride;
  uint32_t *__restrict__ coef = coef_all + (size_t)blockIdx.y * n_coef;
  for (uint32_t ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    // every rank starts at its own slice, so the GPUs do not all read the same peer at the same time
    uint32_t c = ci + first_chunk;
    if (c >= n_chunks) c -= n_chunks;
    const uint32_t k0 = c * PP_LIMBS;
    if (k0 >= n_felts) {  // pure zero padding
#pragma unroll
      for (int i = 0; i < 4; i++) reinterpret_cast<uint4 *>(coef + k0)[t + 256 * i] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    if (pw.my_flags) {
      // The slices are still being uploaded, part by part: wait for the (at most two) parts this chunk's bytes lie in.
      // A part's owner raises flag (2 + part, owner) in every rank's array after its copy (system-scope release).
      if (t < 2) {
        uint32_t byte = c * PP_BYTES + (t ? PP_BYTES - 1 : 0);
        if (byte >= len) byte = len - 1;
        const uint32_t s = byte / slice_len, part = (byte - s * slice_len) / pw.part_len;
        const uint32_t *flag = pw.my_flags + (2 + part) * MAX_PEERS + s;
        const long long t0 = clock64();
        for (;;) {
          uint32_t v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if ((int32_t)(v - pw.epoch) >= 0) break;
          if (clock64() - t0 > 40000000000ll) {  // ~20 s: the owner of that slice is gone
            atomicExch(pw.timeout_flag, 1);
            break;
          }
          __nanosleep(200);
        }
      }
      __syncthreads();
    }
    uint4 v[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {  // 960 pieces of 16 bytes: all loads of a thread are in flight together
      const uint32_t piece = t + 256 * i;
      v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (piece < PP_BYTES / 16) {
        const uint32_t byte = c * PP_BYTES + 16 * piece;  // < 2^31 + 16 KiB: inputs are below 2^31 bytes
        if (byte < len) {
          const uint32_t s = byte / slice_len;
          const uint8_t *src = sl.p[s] + in_off + (byte - s * slice_len);
          if (byte + 16 <= len) {
            v[i] = *reinterpret_cast<const uint4 *>(src);
          } else {
            uint32_t q[4] = {0u, 0u, 0u, 0u};
            for (uint32_t j = 0; byte + j < len; j++) q[j >> 2] |= (uint32_t)src[j] << (8 * (j & 3));
            v[i] = make_uint4(q[0], q[1], q[2], q[3]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (t + 256 * i < PP_BYTES / 16) reinterpret_cast<uint4 *>(w)[t + 256 * i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint32_t out[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t kl = 4 * (t + 256 * i) + j, bit = 30u * kl, wi = bit >> 5, off = bit & 31;
        const uint32_t x = __funnelshift_r(w[wi], off > 2 ? w[wi + 1] : 0u, off) & 0x3fffffffu;
        out[j] = k0 + kl < n_felts ? x : 0u;
      }
      reinterpret_cast<uint4 *>(coef + k0)[t + 256 * i] = make_uint4(out[0], out[1], out[2], out[3]);
    }
    __syncthreads();
  }
}

// polynomials below 4096 coefficients: one thread per limb
__global__ void pack_peers_small_kernel(const __grid_constant__ PeerPtrs sl, uint32_t slice_len, uint32_t len,
                                        uint32_t n_felts, uint32_t n_coef, uint32_t *__restrict__ coef) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_coef) return;
  auto word = [&](uint32_t byte) -> uint32_t {
    uint32_t v = 0;
    for (uint32_t j = 0; j < 4 && byte + j < len; j++) {
      const uint32_t b = byte + j, s = b / slice_len;
      v |= (uint32_t)sl.p[s][b - s * slice_len] << (8 * j);
    }
    return v;
  };
  uint32_t val = 0;
  if (k < n_felts) {
    const uint32_t bit = 30u * k, byte = (bit >> 5) * 4, off = bit & 31;
    val = __funnelshift_r(word(byte), off > 2 ? word(byte + 4) : 0u, off) & 0x3fffffffu;
  }
  coef[k] = val;
}

// Raises flag (channel, rank) = epoch in every rank's flag array (system-scope release): "everything this rank put
// on the stream before this kernel -- the upload of one part of its slice -- is done".
__global__ void peer_signal_kernel(const __grid_constant__ PeerFlags flags, uint32_t world, uint32_t rank, uint32_t channel,
                                   uint32_t epoch) {
  const uint32_t j = threadIdx.x;
  if (j >= world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[j] + channel * MAX_PEERS + rank), "r"(epoch) : "memory");
}
cudaError_t launch_peer_signal(cudaStream_t st, const PeerFlags &flags, uint32_t world, uint32_t rank, uint32_t channel,
                               uint32_t epoch) {
  if (world == 0 || world > MAX_PEERS || rank >= world || channel >= PEER_FLAG_CHANNELS) return cudaErrorInvalidValue;
  peer_signal_kernel<<<1, MAX_PEERS, 0, st>>>(flags, world, rank, channel, epoch);
  return cudaGetLastError();
}

cudaError_t launch_pack_peers(cudaStream_t st, const PeerPtrs &slices, uint32_t world, uint32_t rank, size_t slice_len,
                              size_t len, uint32_t n_felts, uint32_t poly_log, uint32_t *coef, const uint32_t *wait_flags,
                              uint32_t part_len, uint32_t epoch, int *timeout_flag) {
  if (world == 0 || world > MAX_PEERS || slice_len == 0 || (slice_len & 15) || len >= ((size_t)1 << 31) ||
      slice_len >= ((size_t)1 << 31))
    return cudaErrorInvalidValue;
  for (uint32_t r = 0; r < world; r++)
    if ((reinterpret_cast<uintptr_t>(slices.p[r]) & 15) != 0) return cudaErrorInvalidValue;
  const uint32_t n_coef = 4u << poly_log;
  if (n_coef < PP_LIMBS) {
    if (wait_flags) return cudaErrorInvalidValue;  // (callers use a full barrier for inputs this small)
    pack_peers_small_kernel<<<(n_coef + 127) / 128, 128, 0, st>>>(slices, (uint32_t)slice_len, (uint32_t)len, n_felts,
                                                                 n_coef, coef);
    return cudaGetLastError();
  }
  uint32_t bx = n_coef / PP_LIMBS;
  if (bx > 148 * 64) bx = 148 * 64;
  if (wait_flags) {
    // CTAs of this kernel spin until the part they need is uploaded, and the kernels that announce the parts (on the
    // copy stream, this GPU's own among them) need an SM to run on: never let the waiting grid fill the machine
    // (2 CTAs of 256 threads x 64 registers per SM leave half of every SM's threads, registers and shared memory free;
    // with 4 the register file was full and the signal kernels starved -- every rank then waited out the time limit)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (bx > (uint32_t)sms * 2) bx = (uint32_t)sms * 2;
  }
  // chunk holding the first byte of this rank's slice
  const uint32_t first_chunk = (uint32_t)(((uint64_t)rank * slice_len / PP_BYTES) % (n_coef / PP_LIMBS));
  pack_peers_kernel<<<bx, 256, 0, st>>>(slices, (uint32_t)slice_len, (uint32_t)len, n_felts, n_coef, first_chunk, 0, coef,
                                        PackWait{wait_flags, part_len, epoch, timeout_flag});
  return cudaGetLastError();
}

__global__ void gather_roots_kernel(const __grid_constant__ PeerPtrs roots, uint32_t world, uint8_t *tree) {
  // 8 threads per root, 4 bytes each (the peer buffers are only guaranteed 4-byte aligned)
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, r = t >> 3, w = t & 7;
  if (r >= world) return;
  reinterpret_cast<uint32_t *>(tree)[(size_t)(world + r) * 8 + w] = reinterpret_cast<const uint32_t *>(roots.p[r])[w];
}
cudaError_t launch_gather_roots(cudaStream_t st, const PeerPtrs &roots, uint32_t world, uint8_t *tree) {
  if (world == 0 || world > MAX_PEERS) return cudaErrorInvalidValue;
  for (uint32_t r = 0; r < world; r++)
    if ((reinterpret_cast<uintptr_t>(roots.p[r]) & 3) != 0) return cudaErrorInvalidValue;
  gather_roots_kernel<<<(world * 8 + 127) / 128, 128, 0, st>>>(roots, world, tree);
  return cudaGetLastError();
}

__global__ void peer_barrier_kernel(const __grid_constant__ PeerFlags flags, uint32_t world, uint32_t rank,
                                    uint32_t channel, uint32_t epoch, int *timeout_flag) {
  const uint32_t j = threadIdx.x;
  if (j >= world) return;
  __threadfence_system();
  uint32_t *theirs = flags.p[j] + channel * MAX_PEERS + rank;  // my arrival, in peer j's array
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t *mine = flags.p[rank] + channel * MAX_PEERS + j;  // peer j's arrival, in my array
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    if (clock64() - t0 > 40000000000ll) {  // ~20 s at 1.9 GHz (host-side skew between ranks is allowed for): a peer is gone
      atomicExch(timeout_flag, 1);
      break;
    }
  }
}
cudaError_t launch_peer_barrier(cudaStream_t st, const PeerFlags &flags, uint32_t world, uint32_t rank, uint32_t channel,
                                uint32_t epoch, int *timeout_flag) {
  if (world == 0 || world > MAX_PEERS || rank >= world || channel >= PEER_FLAG_CHANNELS) return cudaErrorInvalidValue;
  peer_barrier_kernel<<<1, MAX_PEERS, 0, st>>>(flags, world, rank, channel, epoch, timeout_flag);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- twiddles
// T (length 2^K): for level l = 0..K-1 the x-coordinates of the first half of half_odds(K-l),
// bit-reversed; then a trailing 1 (stwo slow_precompute_twiddles).
__device__ __forceinline__ CPoint point_from_index(const GenPowers &gp, uint32_t idx) {
  CPoint r = {1u, 0u};
#pragma unroll 1
  for (int j = 0; j < 31; j++)
    if ((idx >> j) & 1u) r = cpoint_add(r, gp.g[j]);
  return r;
}

__global__ void __launch_bounds__(256) twiddle_kernel(const __grid_constant__ GenPowers gp, uint32_t K,
                                                       uint32_t *__restrict__ tw, uint32_t *__restrict__ itw,
                                                       uint32_t *__restrict__ tw2) {
  size_t len = (size_t)1 << K;
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= len) return;
  size_t rem = len - g;  // distance from the end, >= 1
  if (rem == 1) {
    tw[g] = 1;
    itw[g] = 1;
    tw2[g] = 2;
    return;
  }
  // the block of length s occupies [len - 2s, len - s): s = largest power of two < rem
  uint32_t ls = 63 - __clzll((long long)(rem - 1));
  size_t s = (size_t)1 << ls;
  uint32_t i = (uint32_t)(g - (len - 2 * s));
  uint32_t idx = half_odds_index(ls + 1, bit_reverse(i, ls));
  uint32_t x = point_from_index(gp, idx).x;
  tw[g] = x;
  itw[g] = m31_inv(x);
  tw2[g] = x + x;
}

cudaError_t launch_twiddles(cudaStream_t st, const GenPowers &gp, uint32_t K, uint32_t *tw, uint32_t *itw,
                            uint32_t *tw2) {
  size_t len = (size_t)1 << K;
  unsigned blocks = (unsigned)((len + 255) / 256);
  twiddle_kernel<<<blocks, 256, 0, st>>>(gp, K, tw, itw, tw2);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- LDE, blocks below 2^10 points
// (larger blocks: lde_warp_kernel further down)
// One CTA = one (block hb, column, blob): a 2^p-point FFT, layers p-1 .. 0, done as radix-16
// passes held in registers (4 layers per shared-memory round trip) and a last pass of
// R_last = ((p-1) mod 4) + 1 layers over 2^R_last consecutive points whose results go straight to
// HBM as 128-bit stores.  The first pass reads the coefficients straight from global memory
// (coalesced: consecutive threads take consecutive low index bits).
// Layer i >= 1 (line): twiddle = blk(2^(K-i))[idx >> (i+1)];  layer 0 (circle): from the pairs
// (x, y) of blk(2^(K-1)) as [y, -y, -x, x] (stwo circle_twiddles_from_line_twiddles).
// A column whose coefficients are all zero evaluates to zero (an arithmetic identity, not a
// special case of a config).  Shared memory carries 4 pad words per 64 so that the pass whose
// stride is below a warp's width stays bank-conflict free.
// `rg` restricts the output to the owned index range [lo, lo + 2^log) (split-blob path): whole
// blocks when the range is at least a block, otherwise the in-range part of one block.
__device__ __forceinline__ uint32_t padi(uint32_t i) { return i + ((i >> 6) << 2); }

__device__ __forceinline__ void bfly_t2(uint32_t &a, uint32_t &b, uint32_t t2) {
  uint32_t tmp = m31_mul_t2(b, t2), va = a;
  a = m31_add(va, tmp);
  b = m31_sub(va, tmp);
}

// R line layers top, top-1, .. top-R+1 over the 2^R register values v[j]; element j has global
// index base + (j << lo) with lo = top - R + 1, and `tw_row(i)` = twiddle row of layer i already
// offset to this group (index of element 0 >> (i+1)).
template <int R, int NSTAGES>
__device__ __forceinline__ void line_stages(uint32_t (&v)[1 << R], const TwiddleTable &tt, uint32_t K, uint32_t p,
                                            uint32_t hb, uint32_t top, uint32_t grp_hi) {
#pragma unroll
  for (int s = 0; s < NSTAGES; s++) {
    const uint32_t i = top - s;
    const uint32_t *tw = tt.blk(1u << (K - i)) + ((size_t)hb << (p - i - 1)) + ((size_t)grp_hi << s);
    constexpr int N = 1 << R;
    const int half = N >> (s + 1);
#pragma unroll
    for (int g = 0; g < (1 << s); g++) {
      uint32_t t = __ldg(tw + g);
      uint32_t t2 = t + t;
#pragma unroll
      for (int k = 0; k < half; k++) bfly_t2(v[g * 2 * half + k], v[g * 2 * half + k + half], t2);
    }
  }
}

template <int R>
__device__ __forceinline__ void last_pass(const uint32_t *__restrict__ src, bool src_is_smem, uint32_t *__restrict__ out,
                                          uint32_t g, const TwiddleTable &tt, uint32_t K, uint32_t p, uint32_t hb,
                                          uint32_t w_lo, uint32_t w_n, bool full) {
  constexpr int N = 1 << R;
  uint32_t v[N];
  const uint32_t idx0 = g << R;
  if (src_is_smem) {
    if (R >= 2) {
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        uint4 q = *reinterpret_cast<const uint4 *>(src + padi(idx0 + j));
        v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
      }
    } else {
      uint2 q = *reinterpret_cast<const uint2 *>(src + padi(idx0));
      v[0] = q.x; v[1] = q.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; j++) v[j] = __ldg(src + idx0 + j);
  }
  // line layers R-1 .. 1
  line_stages<R, R - 1>(v, tt, K, p, hb, R - 1, g);
  // circle layer
  if (R >= 3) {
    // N/2 consecutive circle twiddles starting at a multiple of 4: the [y, -y, -x, x] pattern is static
    const uint32_t *tw0 = tt.blk(1u << (K - 1));
    const uint32_t h_base = (hb << (p - 1)) + (g << (R - 1));
    const uint2 *xy = reinterpret_cast<const uint2 *>(tw0) + (h_base >> 2);
#pragma unroll
    for (int m = 0; m < N / 8; m++) {
      const uint2 q = __ldg(xy + m);  // (x, y)
      const uint32_t y2 = q.y + q.y, ny2 = 2 * m31_neg(q.y), nx2 = 2 * m31_neg(q.x), x2 = q.x + q.x;
      bfly_t2(v[8 * m + 0], v[8 * m + 1], y2);
      bfly_t2(v[8 * m + 2], v[8 * m + 3], ny2);
      bfly_t2(v[8 * m + 4], v[8 * m + 5], nx2);
      bfly_t2(v[8 * m + 6], v[8 * m + 7], x2);
    }
  } else {
    const uint32_t *tw0 = tt.blk(1u << (K - 1));
    const uint32_t h_base = (hb << (p - 1)) + (g << (R - 1));
#pragma unroll
    for (int jj = 0; jj < N / 2; jj++) {
      uint32_t h = h_base + jj;
      uint32_t q = h >> 2, e = h & 3;
      uint32_t x = __ldg(tw0 + 2 * q), y = __ldg(tw0 + 2 * q + 1);
      uint32_t t = e == 0 ? y : e == 1 ? m31_neg(y) : e == 2 ? m31_neg(x) : x;
      bfly_t2(v[2 * jj], v[2 * jj + 1], t + t);
    }
  }
  if (full) {
    if (R >= 2) {
#pragma unroll
      for (int j = 0; j < N; j += 4)
        *reinterpret_cast<uint4 *>(out + idx0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      *reinterpret_cast<uint2 *>(out + idx0) = make_uint2(v[0], v[1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; j++) {
      uint32_t i = idx0 + j;
      if (i >= w_lo && i < w_lo + w_n) out[i] = v[j];
    }
  }
}

// One radix-2^R pass over layers LO+R-1 .. LO with every index compile-time: the 2^R shared-memory
// slots of a group are one computed address plus immediates.
template <int P, int LO, int R, bool FROM_GLOBAL, int THREADS>
__device__ __forceinline__ void rn_pass(const uint32_t *__restrict__ c, uint32_t *sm, const TwiddleTable &tt,
                                        uint32_t K, uint32_t hb) {
  constexpr int N = 1 << R;
  // twiddle rows of the R layers, offset to this block (warp-uniform)
  const uint32_t *row[R];
#pragma unroll
  for (int s = 0; s < R; s++) {
    const int i = LO + R - 1 - s;
    row[s] = tt.blk(1u << (K - i)) + ((size_t)hb << (P - i - 1));
  }
  // padded offset of element j from element 0: the group's low bits never carry into bit 6
  auto off = [](int j) -> uint32_t { return ((uint32_t)j << LO) + ((((uint32_t)j << LO) >> 6) << 2); };
#pragma unroll 1
  for (uint32_t g = threadIdx.x; g < (1u << (P - R)); g += THREADS) {
    const uint32_t lower = g & ((1u << LO) - 1), upper = g >> LO;
    const uint32_t base = (upper << (LO + R)) | lower;
    uint32_t *slot = sm + padi(base);
    uint32_t v[N];
    if (FROM_GLOBAL) {
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = __ldg(c + base + ((uint32_t)j << LO));
    } else {
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = slot[off(j)];
    }
#pragma unroll
    for (int s = 0; s < R; s++) {
      const uint32_t *tw = row[s] + ((size_t)upper << s);
      const int half = N >> (s + 1);
#pragma unroll
      for (int q = 0; q < (1 << s); q++) {
        uint32_t t = __ldg(tw + q);
        uint32_t t2 = t + t;
#pragma unroll
        for (int k = 0; k < half; k++) bfly_t2(v[q * 2 * half + k], v[q * 2 * half + k + half], t2);
      }
    }
#pragma unroll
    for (int j = 0; j < N; j++) slot[off(j)] = v[j];
  }
}

// Pass schedule for a 2^P-point block: a last pass of R_LAST layers on consecutive points, and the
// P - R_LAST layers above it split as evenly as possible into passes of at most MAXR layers.
template <int P, int MAXR>
struct LdeSched {
  static constexpr int R_LAST = P >= 4 ? 4 : P;
  static constexpr int REM = P - R_LAST;
  static constexpr int NP = (REM + MAXR - 1) / MAXR;
  static constexpr int r(int k) { return NP == 0 ? 0 : REM / NP + (k < REM % NP ? 1 : 0); }
  static constexpr int lo(int k) {  // lowest layer of pass k
    int top = P;
    for (int i = 0; i <= k; i++) top -= r(i);
    return top;
  }
};

// INPLACE: the block is a 2^P chunk of an evaluation array that already went through the layers
// above P (large polynomials); it is read from and written back to `eval`, and `p_full` is the
// true poly_log (for the zero-column test and the twiddle geometry D = p_full + beta).
template <int P, int THREADS, bool INPLACE, int MAXR = 4>
__global__ void __launch_bounds__(THREADS) lde_block_kernel(const uint32_t *coef, uint32_t *eval, uint32_t beta,
                                                            uint32_t n_felts, TwiddleTable tt, LdeRange rg,
                                                            uint32_t p_full) {
  extern __shared__ __align__(16) uint32_t sm[];
  constexpr uint32_t p = P;
  const uint32_t hb = blockIdx.x + (uint32_t)(rg.lo >> p), col = blockIdx.y;
  const size_t blob = blockIdx.z;
  constexpr uint32_t n4 = 1u << P;
  const uint32_t D = (INPLACE ? p_full : p) + beta, K = D - 1;
  // local window of this block inside the owned range
  const uint32_t w_lo = rg.log >= p ? 0u : (uint32_t)(rg.lo & (n4 - 1));
  const uint32_t w_n = rg.log >= p ? n4 : (1u << rg.log);
  const bool full = w_n == n4;
  // out[i], i = index inside the block, addresses the owned-range buffer (global index - rg.lo)
  uint32_t *out = eval + ((blob * 4 + col) << rg.log) + ((ptrdiff_t)((size_t)hb << p) - (ptrdiff_t)rg.lo);
  const uint32_t *c = INPLACE ? out : coef + (blob * 4 + col) * (size_t)n4;
  if ((uint64_t)n_felts <= ((uint64_t)col << (INPLACE ? p_full : p))) {  // all-zero column
    if (INPLACE) return;  // the first strided pass already wrote the zeros
    for (uint32_t i = threadIdx.x; i < w_n; i += THREADS) out[w_lo + i] = 0u;
    return;
  }
  using S = LdeSched<P, MAXR>;
  constexpr int R_LAST = S::R_LAST;
  if constexpr (S::NP >= 1) {
    rn_pass<P, S::lo(0), S::r(0), true, THREADS>(c, sm, tt, K, hb);
    __syncthreads();
  }
  if constexpr (S::NP >= 2) {
    rn_pass<P, S::lo(1), S::r(1), false, THREADS>(c, sm, tt, K, hb);
    __syncthreads();
  }
  if constexpr (S::NP >= 3) {
    rn_pass<P, S::lo(2), S::r(2), false, THREADS>(c, sm, tt, K, hb);
    __syncthreads();
  }
  static_assert(S::NP <= 3, "schedule needs more passes than the kernel unrolls");
  const uint32_t *src = S::NP ? sm : c;
#pragma unroll 1
  for (uint32_t g = threadIdx.x; g < (n4 >> R_LAST); g += THREADS)
    last_pass<R_LAST>(src, S::NP != 0, out, g, tt, K, p, hb, w_lo, w_n, full);
}

// ---------------------------------------------------------------- LDE, warp-tiled (2^10 .. 2^15-point blocks)
// Same (block hb, column, blob) decomposition as lde_block_kernel, scheduled so that only ONE
// CTA-wide barrier remains:
//   A  layers P-1 .. 10 (radix 2^(P-10), twiddles uniform over the CTA, held in registers):
//      thread-task = one of the 1024 low index positions, coefficients straight from HBM/L2;
//   -- __syncthreads --
//   the block is now 2^(P-10) independent 1024-point sub-FFTs; each is done by ONE warp:
//   B  layers 9 .. 5, lane = index bits 4..0, registers = bits 9..5 (twiddles uniform over the
//      warp, fetched as 128-bit loads);
//   -- __syncwarp (transpose through the warp's own 32 rows of 36 words) --
//   C  layers 4 .. 1 and the circle layer on 32 consecutive points per lane (128-bit shared loads,
//      per-lane twiddles as 128-bit loads, results staged through the warp's rows and written to HBM as
//      fully coalesced 512-byte pieces).
// Warps drift apart after the barrier, so the load/store phases of one warp overlap the arithmetic
// of the others.  Rows of 32 words are padded to 36: lane-consecutive scalar accesses and 128-bit
// row accesses are both bank-conflict free.  Twiddles come pre-doubled (TwiddleTable::tw2); a
// butterfly with twiddle -t is the butterfly with t and its outputs swapped, so the circle layer
// [y, -y, -x, x] needs no negation.
__device__ __forceinline__ void bfly_t2_swapped(uint32_t &a, uint32_t &b, uint32_t t2) {
  uint32_t tmp = m31_mul_t2(b, t2), va = a;
  a = m31_sub(va, tmp);
  b = m31_add(va, tmp);
}

template <int N>
__device__ __forceinline__ void load_tw_vec(const uint32_t *p, uint32_t (&t)[N]) {
  if constexpr (N == 1) {
    t[0] = __ldg(p);
  } else if constexpr (N == 2) {
    uint2 q = __ldg(reinterpret_cast<const uint2 *>(p));
    t[0] = q.x; t[1] = q.y;
  } else {
#pragma unroll
    for (int j = 0; j < N / 4; j++) {
      uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + j);
      t[4 * j] = q.x; t[4 * j + 1] = q.y; t[4 * j + 2] = q.z; t[4 * j + 3] = q.w;
    }
  }
}

// stage S of a radix-N register pass: 2^S groups of N >> S values, one (doubled) twiddle per group
template <int N, int S>
__device__ __forceinline__ void reg_stage(uint32_t (&v)[N], const uint32_t *t2) {
  constexpr int half = N >> (S + 1);
#pragma unroll
  for (int q = 0; q < (1 << S); q++)
#pragma unroll
    for (int k = 0; k < half; k++) bfly_t2(v[q * 2 * half + k], v[q * 2 * half + k + half], t2[q]);
}

// stages S .. NST-1 with all twiddles in one register array (stage s at t[2^s - 1 ..]).  The first
// SKIP stages have a structurally zero second operand (zero-padded coefficients): a + 0 t = a - 0 t.
template <int N, int NST, int SKIP, int S = 0>
__device__ __forceinline__ void reg_stages_held(uint32_t (&v)[N], const uint32_t *t) {
  if constexpr (S < NST) {
    if constexpr (S < SKIP) {
      constexpr int half = N >> (S + 1);
#pragma unroll
      for (int q = 0; q < (1 << S); q++)
#pragma unroll
        for (int k = 0; k < half; k++) v[q * 2 * half + k + half] = v[q * 2 * half + k];
    } else {
      reg_stage<N, S>(v, t + ((1 << S) - 1));
    }
    reg_stages_held<N, NST, SKIP, S + 1>(v, t);
  }
}

// Phase A of lde_warp_kernel: layers P-1 .. 10 for the 1024 low positions, results to shared memory.
// The loads of task it+1 are issued before the arithmetic of task it (register budget permitting).
template <int P, int THREADS, int SKIP>
__device__ __forceinline__ void lde_phase_a(const uint32_t *__restrict__ c, uint32_t *sm, const uint32_t *ta) {
  constexpr int RA = P - 10, NSUB = 1 << RA, NLOAD = NSUB >> SKIP;  // rows >= NLOAD are zero padding
  constexpr int LDE_SUB_ = 32 * 36;
  constexpr bool PREFETCH = NSUB <= 16;
  uint32_t nxt[PREFETCH ? NLOAD : 1];
  if constexpr (PREFETCH) {
#pragma unroll
    for (int j = 0; j < NLOAD; j++) nxt[j] = __ldg(c + ((uint32_t)j << 10) + threadIdx.x);
  }
#pragma unroll 1
  for (uint32_t low = threadIdx.x; low < 1024; low += THREADS) {
    uint32_t v[NSUB];
#pragma unroll
    for (int j = NLOAD; j < NSUB; j++) v[j] = 0u;
    if constexpr (PREFETCH) {
#pragma unroll
      for (int j = 0; j < NLOAD; j++) v[j] = nxt[j];
      if (low + THREADS < 1024) {
#pragma unroll
        for (int j = 0; j < NLOAD; j++) nxt[j] = __ldg(c + ((uint32_t)j << 10) + low + THREADS);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NLOAD; j++) v[j] = __ldg(c + ((uint32_t)j << 10) + low);
    }
    reg_stages_held<NSUB, RA, SKIP>(v, ta);
    uint32_t *dst = sm + low + ((low >> 5) << 2);
#pragma unroll
    for (int j = 0; j < NSUB; j++) dst[j * LDE_SUB_] = v[j];
  }
}

// stages S .. NST-1 for line layers top, top-1, ..; row(i) = doubled-twiddle row of layer i offset
// to element 0 of this group (global index >> (i+1)); stage s reads 2^s consecutive words there.
template <int N, int NST, int S = 0, typename RowFn>
__device__ __forceinline__ void reg_stages_vec(uint32_t (&v)[N], int top, RowFn row) {
  if constexpr (S < NST) {
    uint32_t t[1 << S];
    load_tw_vec<(1 << S)>(row(top - S), t);
    reg_stage<N, S>(v, t);
    reg_stages_vec<N, NST, S + 1>(v, top, row);
  }
}

constexpr int LDE_ROW = 36;            // padded row of 32 words
constexpr int LDE_SUB = 32 * LDE_ROW;  // one 1024-point sub-FFT in shared memory

// Epilogue variants of lde_warp_kernel (round 2): how a lane's 32 finished points (one 128-byte row
// of the warp's tile) reach HBM.
//   BULK = false: rows re-read across lanes, 8 x (LDS.128 + STG.128) per lane, every store a coalesced 512 bytes;
//   BULK = true : each lane hands its own row to the copy engine: fence.proxy.async + ONE cp.async.bulk
//                 shared::cta -> global of 128 bytes (SASS: UBLKCP), no __syncwarp, no LDS / STG in the issue stream.
//                 A warp's tile is never reused inside the kernel, so the only wait is before the CTA exits.
__device__ __forceinline__ void bulk_store_row(uint32_t *gdst, const uint32_t *srow) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(srow))
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_drain() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <int P, int THREADS, bool INPLACE, bool BULK = false>
__global__ void __launch_bounds__(THREADS) lde_warp_kernel(const uint32_t *coef, uint32_t *eval, uint32_t beta,
                                                           uint32_t n_felts, TwiddleTable tt, LdeRange rg,
                                                           uint32_t p_full, int vec_ok) {
  extern __shared__ __align__(16) uint32_t sm[];
  static_assert(P >= 10 && P <= 15, "block size");
  constexpr int RA = P - 10, NSUB = 1 << RA, NWARPS = THREADS / 32;
  constexpr uint32_t p = P, n4 = 1u << P;
  const uint32_t hb = blockIdx.x + (uint32_t)(rg.lo >> p), col = blockIdx.y;
  const size_t blob = blockIdx.z;
  const uint32_t D = (INPLACE ? p_full : p) + beta, K = D - 1;
  const uint32_t w_lo = rg.log >= p ? 0u : (uint32_t)(rg.lo & (n4 - 1));
  const uint32_t w_n = rg.log >= p ? n4 : (1u << rg.log);
  const bool full = w_n == n4 && vec_ok;
  uint32_t *out = eval + ((blob * 4 + col) << rg.log) + ((ptrdiff_t)((size_t)hb << p) - (ptrdiff_t)rg.lo);
  const uint32_t *c = INPLACE ? out : coef + (blob * 4 + col) * (size_t)n4;
  if ((uint64_t)n_felts <= ((uint64_t)col << (INPLACE ? p_full : p))) {  // all-zero column
    if (INPLACE) return;  // the first strided pass already wrote the zeros
    if (full) {
      for (uint32_t i = threadIdx.x * 4; i < n4; i += THREADS * 4)
        *reinterpret_cast<uint4 *>(out + i) = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (uint32_t i = threadIdx.x; i < w_n; i += THREADS) out[w_lo + i] = 0u;
    }
    return;
  }
  // Coefficients at index >= nz are zero padding (src/utils.rs:23-24), so the second operand of the
  // first P - ceil(log2 nz) layers is zero: those layers only replicate (a function of the input
  // length, not of the data).
  uint32_t skip = 0;
  if constexpr (!INPLACE && RA > 0) {
    const uint64_t rest = (uint64_t)n_felts - ((uint64_t)col << p);
    if (rest < n4) {
      const uint32_t lg = rest <= 1 ? 0u : 32u - __clz((uint32_t)rest - 1u);  // ceil(log2 rest)
      skip = p - lg < (uint32_t)RA ? p - lg : (uint32_t)RA;
    }
  }
  // doubled-twiddle row of line layer i, offset to this block
  auto blk_row = [&](int i) -> const uint32_t * {
    return tt.blk2(1u << (K - i)) + ((size_t)hb << (P - i - 1));
  };
  if constexpr (RA > 0) {
    uint32_t ta[NSUB - 1];
#pragma unroll
    for (int s = 0; s < RA; s++)
#pragma unroll
      for (int q = 0; q < (1 << s); q++) ta[(1 << s) - 1 + q] = __ldg(blk_row(P - 1 - s) + q);
    switch (skip) {
      case 0: lde_phase_a<P, THREADS, 0>(c, sm, ta); break;
      case 1: lde_phase_a<P, THREADS, (RA >= 1 ? 1 : 0)>(c, sm, ta); break;
      case 2: lde_phase_a<P, THREADS, (RA >= 2 ? 2 : 0)>(c, sm, ta); break;
      case 3: lde_phase_a<P, THREADS, (RA >= 3 ? 3 : 0)>(c, sm, ta); break;
      case 4: lde_phase_a<P, THREADS, (RA >= 4 ? 4 : 0)>(c, sm, ta); break;
      default: lde_phase_a<P, THREADS, (RA >= 5 ? 5 : 0)>(c, sm, ta); break;
    }
    __syncthreads();
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (uint32_t sub = warp; sub < (uint32_t)NSUB; sub += NWARPS) {
    uint32_t *base = sm + sub * LDE_SUB;
    uint32_t v[32];
    // ---- B: layers 9..5; register k of this lane is local index (k << 5) | lane
    if constexpr (RA > 0) {
#pragma unroll
      for (int k = 0; k < 32; k++) v[k] = base[k * LDE_ROW + lane];
    } else {
#pragma unroll
      for (int k = 0; k < 32; k++) v[k] = __ldg(c + (k << 5) + lane);
    }
    reg_stages_vec<32, 5>(v, 9, [&](int i) { return blk_row(i) + ((size_t)sub << (9 - i)); });
#pragma unroll
    for (int k = 0; k < 32; k++) base[k * LDE_ROW + lane] = v[k];
    __syncwarp();
    // ---- C: layers 4..1 and the circle layer; register m of this lane is local index (lane << 5) | m
    const uint32_t *rowp = base + lane * LDE_ROW;
#pragma unroll
    for (int m = 0; m < 32; m += 4) {
      uint4 q = *reinterpret_cast<const uint4 *>(rowp + m);
      v[m] = q.x; v[m + 1] = q.y; v[m + 2] = q.z; v[m + 3] = q.w;
    }
    reg_stages_vec<32, 4>(v, 4, [&](int i) {
      return blk_row(i) + ((size_t)sub << (9 - i)) + ((size_t)lane << (4 - i));
    });
    {
      // 16 circle butterflies: pair index h = (global index) >> 1 starts at a multiple of 16, so
      // the [y, -y, -x, x] pattern over 4 consecutive (x, y) pairs of the first line row is static
      const uint32_t *cp = tt.blk2(1u << (K - 1)) + 2 * (((size_t)hb << (P - 3)) + ((size_t)sub << 7) + (lane << 2));
      uint32_t xy[8];
      load_tw_vec<8>(cp, xy);
#pragma unroll
      for (int m = 0; m < 4; m++) {
        const uint32_t x2 = xy[2 * m], y2 = xy[2 * m + 1];
        bfly_t2(v[8 * m + 0], v[8 * m + 1], y2);
        bfly_t2_swapped(v[8 * m + 2], v[8 * m + 3], y2);
        bfly_t2_swapped(v[8 * m + 4], v[8 * m + 5], x2);
        bfly_t2(v[8 * m + 6], v[8 * m + 7], x2);
      }
    }
    const uint32_t idx0 = (sub << 10) | (lane << 5);
    if (full) {
      // back through this lane's row, then 8 fully coalesced 512-byte pieces (a lane-strided
      // 32-byte store costs the LSU data pipe 8x the wavefronts of a coalesced one)
      uint32_t *roww = base + lane * LDE_ROW;
#pragma unroll
      for (int m = 0; m < 32; m += 4) *reinterpret_cast<uint4 *>(roww + m) = make_uint4(v[m], v[m + 1], v[m + 2], v[m + 3]);
      if constexpr (BULK) {
        bulk_store_row(out + idx0, roww);
      } else {
        __syncwarp();
        uint32_t *o = out + (sub << 10) + (lane << 2);
        const uint32_t *src = base + (lane >> 3) * LDE_ROW + ((lane & 7) << 2);
#pragma unroll
        for (int r = 0; r < 8; r++)
          *reinterpret_cast<uint4 *>(o + 128 * r) = *reinterpret_cast<const uint4 *>(src + 4 * r * LDE_ROW);
      }
    } else {
#pragma unroll
      for (int m = 0; m < 32; m++) {
        uint32_t i = idx0 + m;
        if (i >= w_lo && i < w_lo + w_n) out[i] = v[m];
      }
    }
  }
  if constexpr (BULK) bulk_store_drain();  // the copy engine has read this CTA's shared memory
}

// poly_log 0: one coefficient per column; every layer is a replication.
__global__ void lde_copy_kernel(const uint32_t *__restrict__ coef, uint32_t *__restrict__ eval, uint32_t beta,
                                uint32_t n_felts, LdeRange rg) {
  const uint32_t hb = blockIdx.x + (uint32_t)rg.lo, col = blockIdx.y;
  const size_t blob = blockIdx.z;
  (void)beta;
  eval[((blob * 4 + col) << rg.log) + (hb - rg.lo)] = n_felts > col ? coef[blob * 4 + col] : 0u;
}

// D == 1 and D == 2: stwo hard-codes these; one thread per (blob, column).
__global__ void lde_tiny_kernel(const uint32_t *__restrict__ coef, uint32_t *__restrict__ eval, uint32_t p,
                                uint32_t D, size_t n_cols, CPoint init) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_cols) return;
  const uint32_t *c = coef + (g << p);
  uint32_t *o = eval + (g << D);
  uint32_t v[4] = {0, 0, 0, 0};
  for (uint32_t i = 0; i < (1u << p); i++) v[i] = c[i];
  auto bfly = [](uint32_t &a, uint32_t &b, uint32_t t) {
    uint32_t tmp = m31_mul(b, t), va = a;
    a = m31_add(va, tmp);
    b = m31_sub(va, tmp);
  };
  if (D == 1) {
    bfly(v[0], v[1], init.y);
  } else {
    bfly(v[0], v[2], init.x);
    bfly(v[1], v[3], init.x);
    bfly(v[0], v[1], init.y);
    bfly(v[2], v[3], m31_neg(init.y));
  }
  for (uint32_t i = 0; i < (1u << D); i++) o[i] = v[i];
}

// Large polynomials (2^p points do not fit one CTA's shared memory): the layers above the final
// 2^c chunks are done 4 at a time by a register-only radix-16 pass.  A thread owns one column
// position and the 16 rows that are 2^(top-4) apart; a warp reads/writes 32 consecutive words of
// each row (128 B), so every access is coalesced and nothing goes through shared memory.  FIRST
// reads the replicated coefficients (index mod 2^p) and may compute rows outside the owned range
// (only in-range rows are stored); later passes work in place inside the owned range.
template <bool FIRST>
__global__ void __launch_bounds__(256) lde_strided_r16_kernel(const uint32_t *__restrict__ coef, uint32_t *eval,
                                                              uint32_t p, uint32_t beta, uint32_t n_felts, uint32_t top,
                                                              TwiddleTable tt, LdeRange rg) {
  const uint32_t D = p + beta, K = D - 1;
  const uint32_t col = blockIdx.y;
  const size_t blob = blockIdx.z;
  const uint32_t rs = top - 4;  // log2 of the row stride
  // indices are below 2^28 (D <= 28): 32-bit index arithmetic throughout
  const uint32_t lo32 = (uint32_t)rg.lo, span = 1u << rg.log;
  uint32_t *ev = eval + ((blob * 4 + col) << rg.log) - rg.lo;  // local address = global index - rg.lo
  const bool zero_col = (uint64_t)n_felts <= ((uint64_t)col << p);
  if (zero_col && !FIRST) return;
  // one thread per (sub-FFT, column position); sub-FFTs touching the owned range only
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t sub = (t >> rs) + (lo32 >> top);
  const uint32_t pos = t & ((1u << rs) - 1);
  const uint32_t base = (sub << top) + pos;
  const uint32_t rstep = 1u << rs;
  uint32_t v[16];
  if (FIRST) {
    if (zero_col) {
      uint32_t idx = base;
#pragma unroll
      for (int j = 0; j < 16; j++, idx += rstep)
        if (idx - lo32 < span) ev[idx] = 0u;
      return;
    }
    // top == p: the coefficient index is the index inside the block (the block index only selects twiddles)
    const uint32_t *c = coef + ((blob * 4 + col) << p);
    uint32_t ci = pos;
#pragma unroll
    for (int j = 0; j < 16; j++, ci += rstep) v[j] = __ldg(c + ci);
  } else {
    uint32_t idx = base;
#pragma unroll
    for (int j = 0; j < 16; j++, idx += rstep) v[j] = ev[idx];
  }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const uint32_t i = top - 1 - s;
    // twiddle index of element j: (global index) >> (i + 1) = (sub << s) | (j >> (4 - s))
    const uint32_t *tw = tt.blk2(1u << (K - i)) + (sub << s);  // pre-doubled twiddles
    const int half = 8 >> s;
#pragma unroll
    for (int q = 0; q < (1 << s); q++) {
      uint32_t t2 = __ldg(tw + q);
#pragma unroll
      for (int k = 0; k < half; k++) bfly_t2(v[q * 2 * half + k], v[q * 2 * half + k + half], t2);
    }
  }
  uint32_t idx = base;
#pragma unroll
  for (int j = 0; j < 16; j++, idx += rstep)
    if (!FIRST || idx - lo32 < span) ev[idx] = v[j];
}

// Eight or nine layers in ONE sweep (round 2): top-1 .. top-R over 2^R rows that are 2^(top-R) apart, R = A + 4, as a
// radix-2^A and a radix-16 register step with a transpose through shared memory in between.  A CTA owns a tile of
// 2^R rows x 32 consecutive columns (32 / 64 KiB): every global access is a full 128-byte line per warp, as in
// lde_strided_r16_kernel, but the evaluation array crosses HBM once for R layers instead of twice.
//   step 1  thread (r, col): rows {16 j + r}, row-index bits R-1..4, layers top-1 .. top-A; twiddles CTA-uniform
//   step 2  thread (h, col): rows {16 h + r}, row-index bits 3..0, layers top-A-1 .. top-R; twiddles per warp
// The tile is stored [row][col]: both steps access it with lane = column, conflict-free without padding.
// Every index is below 2^28 (D <= 28): 32-bit index arithmetic (the 64-bit form of the same code executed twice the
// instructions, profiles/r02_ncu_summary.md).
template <bool FIRST, int A>
__global__ void __launch_bounds__(512) lde_strided_tile_kernel(const uint32_t *__restrict__ coef, uint32_t *eval,
                                                               uint32_t p, uint32_t beta, uint32_t n_felts, uint32_t top,
                                                               TwiddleTable tt, LdeRange rg) {
  extern __shared__ uint32_t tile[];  // [2^R][32]
  constexpr int R = A + 4, NA = 1 << A;
  const uint32_t D = p + beta, K = D - 1;
  const uint32_t col = blockIdx.y;
  const size_t blob = blockIdx.z;
  const uint32_t rs = top - R;  // log2 of the row stride
  const uint32_t lo32 = (uint32_t)rg.lo, span = 1u << rg.log;
  uint32_t *ev = eval + ((blob * 4 + col) << rg.log) - rg.lo;  // local address = global index - rg.lo
  const bool zero_col = (uint64_t)n_felts <= ((uint64_t)col << p);
  if (zero_col && !FIRST) return;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;  // 16 warps: w = r in step 1
  const uint32_t cg_log = rs - 5;                                // column groups of 32 per sub-FFT
  const uint32_t sub = (blockIdx.x >> cg_log) + (lo32 >> top);
  const uint32_t col0 = ((blockIdx.x & ((1u << cg_log) - 1)) << 5) + lane;
  const uint32_t base = (sub << top) + col0;
  const uint32_t rstep = 16u << rs;
  if (FIRST && zero_col) {
    for (uint32_t h = w; h < (uint32_t)NA; h += 16) {
      uint32_t idx = base + ((16 * h) << rs);
#pragma unroll
      for (int k = 0; k < 16; k++, idx += 1u << rs)
        if (idx - lo32 < span) ev[idx] = 0u;
    }
    return;
  }
  {
    // ---- step 1: rows 16 j + w, j = 0 .. 2^A - 1
    uint32_t v[NA];
    if (FIRST) {
      // top == p: the block index `sub` only selects twiddles; the coefficient index is the index inside the block
      const uint32_t *c = coef + ((blob * 4 + col) << p);
      uint32_t ci = col0 + (w << rs);
#pragma unroll
      for (int j = 0; j < NA; j++, ci += rstep) v[j] = __ldg(c + ci);
    } else {
      uint32_t idx = base + (w << rs);
#pragma unroll
      for (int j = 0; j < NA; j++, idx += rstep) v[j] = ev[idx];
    }
#pragma unroll
    for (int s = 0; s < A; s++) {
      const uint32_t i = top - 1 - s;
      const uint32_t *tw = tt.blk2(1u << (K - i)) + (sub << s);  // index of row-group j: (sub << s) | (j >> (A - s))
      const int half = NA >> (s + 1);
#pragma unroll
      for (int q = 0; q < (1 << s); q++) {
        const uint32_t t2 = __ldg(tw + q);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_t2(v[q * 2 * half + k], v[q * 2 * half + k + half], t2);
      }
    }
#pragma unroll
    for (int j = 0; j < NA; j++) tile[(16 * j + w) * 32 + lane] = v[j];
  }
  __syncthreads();
  // ---- step 2: rows 16 h + r, h = w (and w + 16 when A = 5)
#pragma unroll 1
  for (uint32_t h = w; h < (uint32_t)NA; h += 16) {
    const uint32_t idx0 = base + ((16 * h) << rs);
    // partial range (split blob, FIRST only): 16 rows that all lie outside the owned range have nothing to store
    if (FIRST && span < (1u << top) && (idx0 + (15u << rs) < lo32 || idx0 >= lo32 + span)) continue;
    uint32_t v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = tile[(16 * h + r) * 32 + lane];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const uint32_t i = top - A - 1 - s;
      // index of row 16 h + r: (sub << (A + s)) | (h << s) | (r >> (4 - s)); 2^s consecutive words per warp
      const uint32_t *tw = tt.blk2(1u << (K - i)) + (((sub << A) | h) << s);
      const int half = 8 >> s;
      uint32_t t2[8];
      if (s == 0) t2[0] = __ldg(tw);
      if (s == 1) { const uint2 q = __ldg(reinterpret_cast<const uint2 *>(tw)); t2[0] = q.x; t2[1] = q.y; }
      if (s == 2) { const uint4 q = __ldg(reinterpret_cast<const uint4 *>(tw)); t2[0] = q.x; t2[1] = q.y; t2[2] = q.z; t2[3] = q.w; }
      if (s == 3) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(tw)), q1 = __ldg(reinterpret_cast<const uint4 *>(tw) + 1);
        t2[0] = q0.x; t2[1] = q0.y; t2[2] = q0.z; t2[3] = q0.w; t2[4] = q1.x; t2[5] = q1.y; t2[6] = q1.z; t2[7] = q1.w;
      }
#pragma unroll
      for (int q = 0; q < (1 << s); q++)
#pragma unroll
        for (int k = 0; k < half; k++) bfly_t2(v[q * 2 * half + k], v[q * 2 * half + k + half], t2[q]);
    }
    uint32_t idx = idx0;
#pragma unroll
    for (int r = 0; r < 16; r++, idx += 1u << rs)
      if (!FIRST || idx - lo32 < span) ev[idx] = v[r];
  }
}

// ---------------------------------------------------------------- decode (erasure recovery)
// Any one of the 2^log_blowup coset blocks of an evaluation column determines the polynomial: the
// block is the 2^p-point FFT of the coefficients with block-specific twiddles, so running the
// layers backwards (circle layer first, then the line layers with growing stride) with the inverse
// twiddles and scaling by 2^-p recovers the coefficients.  ibutterfly: (x, y) -> (x + y, (x - y) / t).
// This is the decode side of the Reed-Solomon code the commit path encodes (the
// reference's README describes sampling/recovery but its code has no counterpart).
// Layers 0 .. c-1 of the inverse transform inside 2^c-point chunks (c = min(p, 15)); a chunk is one CTA in shared
// memory.  The result is scaled by inv_n when the chunk is the whole block (p == c).
template <int THREADS>
__global__ void __launch_bounds__(THREADS) lde_block_inverse_kernel(const uint32_t *__restrict__ block_evals,
                                                                    uint32_t *__restrict__ coef, uint32_t p, uint32_t c,
                                                                    uint32_t beta, uint32_t hb, uint32_t inv_n,
                                                                    TwiddleTable tt) {
  extern __shared__ uint32_t smi[];
  const uint32_t col = blockIdx.y, ch = blockIdx.x;
  const uint32_t n4 = 1u << c;
  const uint32_t D = p + beta, K = D - 1;
  const uint32_t *src = block_evals + ((size_t)col << p) + ((size_t)ch << c);
  for (uint32_t i = threadIdx.x; i < n4; i += THREADS) smi[i] = src[i];
  __syncthreads();
  const uint32_t half = n4 >> 1;
  if (c >= 1) {
    // circle layer (pairs of neighbours), inverse twiddles [1/y, -1/y, -1/x, 1/x]
    const uint32_t *itw0 = tt.iblk(1u << (K - 1));
    for (uint32_t bf = threadIdx.x; bf < half; bf += THREADS) {
      uint32_t h = (hb << (p - 1)) | (ch << (c - 1)) | bf;
      uint32_t q = h >> 2, e = h & 3;
      uint32_t ix = __ldg(itw0 + 2 * q), iy = __ldg(itw0 + 2 * q + 1);
      uint32_t it = e == 0 ? iy : e == 1 ? m31_neg(iy) : e == 2 ? m31_neg(ix) : ix;
      uint32_t x = smi[2 * bf], y = smi[2 * bf + 1];
      smi[2 * bf] = m31_add(x, y);
      smi[2 * bf + 1] = m31_mul(m31_sub(x, y), it);
    }
    __syncthreads();
  }
  for (uint32_t i = 1; i < c; i++) {
    // twiddle index of a butterfly = (index of its first point inside the evaluation domain) >> (i + 1)
    const uint32_t *itw = tt.iblk(1u << (K - i)) + ((size_t)hb << (p - i - 1)) + ((size_t)ch << (c - i - 1));
    for (uint32_t bf = threadIdx.x; bf < half; bf += THREADS) {
      uint32_t lo = bf & ((1u << i) - 1), hi = bf >> i;
      uint32_t a = (hi << (i + 1)) | lo, b = a + (1u << i);
      uint32_t it = __ldg(itw + hi);
      uint32_t x = smi[a], y = smi[b];
      smi[a] = m31_add(x, y);
      smi[b] = m31_mul(m31_sub(x, y), it);
    }
    __syncthreads();
  }
  uint32_t *dst = coef + ((size_t)col << p) + ((size_t)ch << c);
  const uint32_t scale = p == c ? inv_n : 1u;
  for (uint32_t i = threadIdx.x; i < n4; i += THREADS) dst[i] = m31_mul(smi[i], scale);
}

// One inverse layer i >= 15 over the whole block in HBM (polynomials above 2^15 coefficients per column): a thread
// per butterfly, both accesses coalesced (the stride is at least 2^15).  `scale` (2^-p on the last layer, else 1)
// multiplies both outputs.  Decoding is not on the commit path: one sweep per layer is good enough here.
__global__ void __launch_bounds__(256) lde_inverse_layer_kernel(uint32_t *__restrict__ coef, uint32_t p, uint32_t i,
                                                                uint32_t beta, uint32_t hb, uint32_t scale,
                                                                TwiddleTable tt) {
  const uint32_t D = p + beta, K = D - 1;
  const uint32_t col = blockIdx.y;
  uint32_t *v = coef + ((size_t)col << p);
  const uint32_t *itw = tt.iblk(1u << (K - i)) + ((size_t)hb << (p - i - 1));
  const uint32_t n_bf = 1u << (p - 1);
  for (uint32_t bf = blockIdx.x * blockDim.x + threadIdx.x; bf < n_bf; bf += gridDim.x * blockDim.x) {
    const uint32_t lo = bf & ((1u << i) - 1), hi = bf >> i;
    const uint32_t a = (hi << (i + 1)) | lo, b = a + (1u << i);
    const uint32_t it = __ldg(itw + hi);
    const uint32_t x = v[a], y = v[b];
    v[a] = m31_mul(m31_add(x, y), scale);
    v[b] = m31_mul(m31_mul(m31_sub(x, y), it), scale);
  }
}

// Inverse of pack_kernel: 30-bit limbs back to bytes.  flag[0] is set when the coefficients are not
// a packing of `len` bytes (a limb >= 2^30, or non-zero padding), i.e. the block was not a codeword.
__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t *__restrict__ coef, uint32_t n_coef,
                                                      uint32_t n_felts, size_t len, uint8_t *__restrict__ out,
                                                      int *flag) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = gid; i < len; i += stride) {
    const uint64_t bit = 8ull * i;
    const uint32_t k = (uint32_t)(bit / 30), off = (uint32_t)(bit % 30);
    uint64_t v = coef[k];
    if (off > 22 && k + 1 < n_coef) v |= (uint64_t)coef[k + 1] << 30;
    out[i] = (uint8_t)(v >> off);
  }
  for (size_t k = gid; k < n_coef; k += stride) {
    uint32_t c = coef[k];
    bool bad = c >= (1u << 30) || (k >= n_felts && c != 0);
    if (k + 1 == n_felts) {
      // the last limb only carries the remaining bits
      uint32_t used = (uint32_t)(8ull * len - 30ull * k);
      if (used < 30 && (c >> used)) bad = true;
    }
    if (bad) atomicExch(flag, 1);
  }
}

cudaError_t launch_decode_block(cudaStream_t st, const uint32_t *block_evals, uint32_t *coef, uint32_t p,
                                uint32_t beta, uint32_t hb, const TwiddleTable &tt, size_t len, uint32_t n_felts,
                                uint8_t *out, int *flag) {
  if (p > 26 || p + beta < 3 || p + beta > 28) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(lde_block_inverse_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 << 15);
    attr = true;
  }
  const uint32_t inv_n = m31_inv(1u << p);
  const uint32_t c = p < 15 ? p : 15;
  lde_block_inverse_kernel<1024><<<dim3(1u << (p - c), 4), 1024, (size_t)4 << c, st>>>(block_evals, coef, p, c, beta, hb,
                                                                                      inv_n, tt);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  for (uint32_t i = c; i < p; i++) {
    lde_inverse_layer_kernel<<<dim3(2048, 4), 256, 0, st>>>(coef, p, i, beta, hb, i + 1 == p ? inv_n : 1u, tt);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  const uint32_t n_coef = 4u << p;
  size_t work = len > n_coef ? len : n_coef;
  unsigned bx = (unsigned)((work + 255) / 256);
  if (bx > 4096) bx = 4096;
  if (bx == 0) bx = 1;
  unpack_kernel<<<bx, 256, 0, st>>>(coef, n_coef, n_felts, len, out, flag);
  return cudaGetLastError();
}

constexpr bool LDE_BULK_DEFAULT = false;  // set from the measurement in profiles/r02_lde_probes.txt
constexpr uint32_t LDE_SMEM_LOG_MAX = 15;  // 2^15 u32 = 128 KiB of shared memory per CTA

cudaError_t launch_lde(cudaStream_t st, const uint32_t *coef, uint32_t *eval, uint32_t p, uint32_t beta,
                       size_t n_blobs, uint32_t n_felts, const TwiddleTable &tt, CPoint half_initial,
                       const LdeRange *range) {
  const uint32_t D = p + beta;
  if (D == 0) return cudaErrorInvalidValue;
  LdeRange rg = range ? *range : LdeRange{0, D};
  if (rg.log > D || (rg.lo & (((size_t)1 << rg.log) - 1)) || rg.lo >= ((size_t)1 << D)) return cudaErrorInvalidValue;
  if (D <= 2) {
    if (rg.log != D) return cudaErrorInvalidValue;
    size_t n_cols = n_blobs * 4;
    lde_tiny_kernel<<<(unsigned)((n_cols + 127) / 128), 128, 0, st>>>(coef, eval, p, D, n_cols, half_initial);
    return cudaGetLastError();
  }
  static const bool attr_set = [] {  // blocks of 2^14 / 2^15 points need more than the default 48 KiB
    cudaFuncSetAttribute(lde_warp_kernel<14, 256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<14, 256, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<15, 512, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<15, 512, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<14, 256, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<14, 256, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<15, 512, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * LDE_SUB * 4);
    cudaFuncSetAttribute(lde_warp_kernel<15, 512, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * LDE_SUB * 4);
    return true;
  }();
  (void)attr_set;
  // epilogue through the copy engine (cp.async.bulk) or through LDS/STG; FRIEDA_LDE_BULK=0/1 overrides (A/B probes)
  static const bool bulk = [] {
    const char *e = std::getenv("FRIEDA_LDE_BULK");
    return e ? e[0] != '0' : LDE_BULK_DEFAULT;
  }();
  const int vec_ok = (reinterpret_cast<uintptr_t>(eval) & 31) == 0;
  for (size_t b0 = 0; b0 < n_blobs; b0 += 32768) {
    size_t nb = n_blobs - b0 < 32768 ? n_blobs - b0 : 32768;
    const uint32_t *cf = coef + b0 * ((size_t)4 << p);
    uint32_t *ev = eval + b0 * ((size_t)4 << rg.log);
    if (p >= 10 && p <= LDE_SMEM_LOG_MAX) {
      dim3 grid(rg.log >= p ? 1u << (rg.log - p) : 1u, 4, (unsigned)nb);
#define FR_LDE_WARP(PP, TT)                                                                                           \
  case PP:                                                                                                            \
    if (bulk)                                                                                                         \
      lde_warp_kernel<PP, TT, false, true><<<grid, TT, ((size_t)LDE_SUB * 4) << (PP - 10), st>>>(cf, ev, beta, n_felts, \
                                                                                                 tt, rg, p, vec_ok);  \
    else                                                                                                              \
      lde_warp_kernel<PP, TT, false, false><<<grid, TT, ((size_t)LDE_SUB * 4) << (PP - 10), st>>>(cf, ev, beta, n_felts, \
                                                                                                  tt, rg, p, vec_ok); \
    break;
      switch (p) {
        FR_LDE_WARP(10, 32) FR_LDE_WARP(11, 64) FR_LDE_WARP(12, 128) FR_LDE_WARP(13, 256) FR_LDE_WARP(14, 256)
        FR_LDE_WARP(15, 512)
      }
#undef FR_LDE_WARP
    } else if (p < 10) {
      dim3 grid(rg.log >= p ? 1u << (rg.log - p) : 1u, 4, (unsigned)nb);
      size_t smem = ((size_t)4 << p) + (((size_t)4 << p) >> 4) + 64;  // + 4 pad words per 64
#define FR_LDE_CASE(PP, TT) \
  case PP: lde_block_kernel<PP, TT, false><<<grid, TT, smem, st>>>(cf, ev, beta, n_felts, tt, rg, p); break;
      switch (p) {
        case 0: lde_copy_kernel<<<grid, 1, 0, st>>>(cf, ev, beta, n_felts, rg); break;
        FR_LDE_CASE(1, 32) FR_LDE_CASE(2, 32) FR_LDE_CASE(3, 32) FR_LDE_CASE(4, 32) FR_LDE_CASE(5, 32)
        FR_LDE_CASE(6, 32) FR_LDE_CASE(7, 64) FR_LDE_CASE(8, 64) FR_LDE_CASE(9, 128)
        default: return cudaErrorInvalidValue;
      }
#undef FR_LDE_CASE
    } else {
      // Strided passes over the layers above the shared-memory chunks: tile sweeps of nine or eight layers (one HBM
      // round trip each; nine when that leaves 2^14-point chunks, the size at which lde_warp_kernel keeps three CTAs
      // per SM) while at least 2^12 points remain per chunk, otherwise a register-only radix-16 pass; then 2^c
      // chunks in shared memory, c in 12..15.  After the first pass everything stays inside the owned range.
      // FRIEDA_LDE_TILE=0 keeps to radix-16 passes (the round-1 schedule), =8 to eight-layer tiles (A/B probes).
      static const int tile_mode = [] {
        const char *e = std::getenv("FRIEDA_LDE_TILE");
        return e ? std::atoi(e) : 9;
      }();
      static const bool tile_attr = [] {
        cudaFuncSetAttribute(lde_strided_tile_kernel<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(lde_strided_tile_kernel<false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        return true;
      }();
      (void)tile_attr;
      uint32_t c = p, n_pass = 0, radix_log[8];
      while (c > LDE_SMEM_LOG_MAX) {
        uint32_t r = 4;
        if (tile_mode >= 9 && c - 14 == 9) r = 9;
        else if (tile_mode >= 8 && c >= 20) r = 8;
        radix_log[n_pass++] = r;
        c -= r;
      }
      if (rg.log < c || (rg.log < p && p - rg.log > radix_log[0])) return cudaErrorInvalidValue;
      uint32_t top = p;
      for (uint32_t pass = 0; pass < n_pass; top -= radix_log[pass], pass++) {
        const size_t subs = rg.log >= top ? (size_t)1 << (rg.log - top) : 1;
        const uint32_t r = radix_log[pass];
        if (r >= 8) {
          // CTAs = (sub-FFTs touching the range) x (2^(top-r) / 32 column groups); 512 threads
          dim3 grid((unsigned)(subs << (top - r - 5)), 4, (unsigned)nb);
          const size_t smem = (size_t)128 << r;
          if (r == 8 && pass == 0) lde_strided_tile_kernel<true, 4><<<grid, 512, smem, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
          if (r == 8 && pass != 0) lde_strided_tile_kernel<false, 4><<<grid, 512, smem, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
          if (r == 9 && pass == 0) lde_strided_tile_kernel<true, 5><<<grid, 512, smem, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
          if (r == 9 && pass != 0) lde_strided_tile_kernel<false, 5><<<grid, 512, smem, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
          continue;
        }
        // threads = (sub-FFTs touching the range) x 2^(top-4) column positions
        size_t threads = subs << (top - 4);
        dim3 grid((unsigned)(threads / 256), 4, (unsigned)nb);
        if (pass == 0)
          lde_strided_r16_kernel<true><<<grid, 256, 0, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
        else
          lde_strided_r16_kernel<false><<<grid, 256, 0, st>>>(cf, ev, p, beta, n_felts, top, tt, rg);
      }
      dim3 grid(1u << (rg.log - c), 4, (unsigned)nb);
#define FR_LDE_WARP(PP, TT)                                                                                          \
  case PP:                                                                                                           \
    if (bulk)                                                                                                        \
      lde_warp_kernel<PP, TT, true, true><<<grid, TT, ((size_t)LDE_SUB * 4) << (PP - 10), st>>>(cf, ev, beta, n_felts, \
                                                                                                tt, rg, p, vec_ok);  \
    else                                                                                                             \
      lde_warp_kernel<PP, TT, true, false><<<grid, TT, ((size_t)LDE_SUB * 4) << (PP - 10), st>>>(cf, ev, beta, n_felts, \
                                                                                                 tt, rg, p, vec_ok); \
    break;
      switch (c) { FR_LDE_WARP(12, 128) FR_LDE_WARP(13, 256) FR_LDE_WARP(14, 256) FR_LDE_WARP(15, 512) }
#undef FR_LDE_WARP
    }
  }
  return cudaGetLastError();
}

}  // namespace frieda
