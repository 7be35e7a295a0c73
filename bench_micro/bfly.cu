// Microbenchmark: the M31 radix-2 butterfly (a, b, t) -> (a + b t, a - b t) as the LDE kernels
// issue it, in registers only (radix-32 passes, 16 independent butterflies per stage), in several
// formulations.  Prints clocks per warp-butterfly per SM sub-partition: the arithmetic floor of the
// circle-FFT pass.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../frieda_b200/csrc/m31.cuh"
using namespace frieda;

// V0: IMAD.WIDE + LEA.HI + min, add + min, sub + min (what lde.cu uses)
__device__ __forceinline__ void bf0(uint32_t &a, uint32_t &b, uint32_t t2) {
  uint32_t tmp = m31_mul_t2(b, t2), va = a;
  a = m31_add(va, tmp);
  b = m31_sub(va, tmp);
}
// V1: Shoup: q = hi(b * t'), r = b * t - q * P (all IMAD), r in [0, 2P)
__device__ __forceinline__ void bf1(uint32_t &a, uint32_t &b, uint32_t t, uint32_t tp) {
  uint32_t q = __umulhi(b, tp);
  uint32_t r = b * t - q * P31;
  r = umin32(r, r - P31);
  uint32_t va = a;
  a = m31_add(va, r);
  b = m31_sub(va, r);
}
// V2: as V0 but the two adds forced onto the ALU pipe (IADD3 via inline asm add)
__device__ __forceinline__ void bf2(uint32_t &a, uint32_t &b, uint32_t t2) {
  uint32_t tmp = m31_mul_t2(b, t2), va = a, s, d;
  asm("add.u32 %0, %1, %2;" : "=r"(s) : "r"(va), "r"(tmp));
  asm("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(va), "r"(tmp));
  a = umin32(s, s - P31);
  b = umin32(d, d + P31);
}
// V3: the product reduced by AND/shift instead of the doubled twiddle (undoubled t)
__device__ __forceinline__ void bf3(uint32_t &a, uint32_t &b, uint32_t t) {
  uint32_t tmp = m31_mul(b, t), va = a;
  a = m31_add(va, tmp);
  b = m31_sub(va, tmp);
}

// ---- lazy forms (round 2): values live in [0, 2^32) as arbitrary residues, reduced only where an overflow looms.
// L1 (Shoup, any 32-bit b): q = hi(b * t'), r = b t - q P in [0, 2P)  [t' = floor(t 2^32 / P)]
//   Tr = min(r, r - P) in [0, P);  aP = max(a, a + P) in [P, 2P + 1]  (a + P wraps exactly when a >= P + 2)
//   a' = aP + Tr - P in [0, 2P];   b' = aP - Tr in [1, 2P + 1]
// 3 FMA-pipe instructions for the product, 3 ALU-pipe (2 VIADDMNMX + IADD3), and the subtraction on either pipe.
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
template <bool SUB_ON_FMA>
__device__ __forceinline__ void bfL1(uint32_t &a, uint32_t &b, uint32_t t, uint32_t tp, uint32_t neg1) {
  uint32_t q = __umulhi(b, tp);
  uint32_t r = b * t - q * P31;
  uint32_t tr = umin32(r, r - P31);
  uint32_t ap = umax32(a, a + P31);
  a = ap + tr - P31;
  if (SUB_ON_FMA) {
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(tr), "r"(neg1), "r"(ap));  // ap - tr as IMAD (neg1 = runtime -1)
  } else {
    // neg1 + 1 is an opaque zero: a three-register add can only be an ALU-pipe IADD3 (ptxas turns a plain
    // two-input subtraction into an FMA-pipe IMAD.IADD on its own)
    b = ap - tr + (neg1 + 1u);
  }
}
// L0 (doubled twiddle, needs b <= 2^31): b reduced first, the rest as L1
__device__ __forceinline__ void bfL0(uint32_t &a, uint32_t &b, uint32_t t2, uint32_t neg1) {
  uint32_t br = umin32(b, b - P31);                 // [0, P + 1]
  uint64_t p = (uint64_t)br * t2;
  uint32_t r = (uint32_t)(p >> 32) + ((uint32_t)p >> 1);  // <= 2P
  uint32_t tr = umin32(r, r - P31);
  uint32_t ap = umax32(a, a + P31);
  a = ap + tr - P31;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(b) : "r"(tr), "r"(neg1), "r"(ap));
}
__device__ __forceinline__ uint32_t canon(uint32_t x) {
  x = umin32(x, x - P31);
  return umin32(x, x - P31);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, const uint32_t *tw, int iters) {
  uint32_t v[32], t[31], tp[31], t2l[MODE == 10 ? 31 : 1];
  const uint32_t neg1 = tw[93];  // runtime -1: keeps `ap - tr` an IMAD where asked
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = (threadIdx.x * 2654435761u + j * 40503u + blockIdx.x) & 0x7ffffffeu;
  if (MODE >= 7) {
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = canon(v[j]);
  }
#pragma unroll
  for (int j = 0; j < 31; j++) {
    t[j] = tw[(MODE >= 7 ? 62 : 0) + j];  // modes >= 7: canonical twiddle; below: the value as round 1 timed it
    tp[j] = tw[31 + j];
    if (MODE == 10) t2l[j] = 2 * tw[62 + j];
  }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
      const int half = 16 >> s;
#pragma unroll
      for (int q = 0; q < (1 << s); q++)
#pragma unroll
        for (int kk = 0; kk < half; kk++) {
          uint32_t &a = v[q * 2 * half + kk], &b = v[q * 2 * half + kk + half];
          const int ti = (1 << s) - 1 + q;
          if (MODE == 0) bf0(a, b, t[ti]);
          if (MODE == 1) bf1(a, b, t[ti], tp[ti]);
          if (MODE == 2) bf2(a, b, t[ti]);
          if (MODE == 3) bf3(a, b, t[ti]);
          if (MODE == 4) { if (kk % 3 == 2) bf1(a, b, t[ti], tp[ti]); else bf0(a, b, t[ti]); }
          if (MODE == 5) { if (kk % 4 == 3) bf1(a, b, t[ti], tp[ti]); else bf0(a, b, t[ti]); }
          if (MODE == 6) { if (kk % 2 == 1) bf1(a, b, t[ti], tp[ti]); else bf0(a, b, t[ti]); }
          if (MODE == 7) bfL1<false>(a, b, t[ti], tp[ti], neg1);
          if (MODE == 8) bfL1<true>(a, b, t[ti], tp[ti], neg1);
          if (MODE == 9) { if (kk % 2 == 1) bfL1<true>(a, b, t[ti], tp[ti], neg1); else bfL1<false>(a, b, t[ti], tp[ti], neg1); }
          if (MODE == 10) { if (kk % 4 == 3) bfL0(a, b, t2l[ti], neg1); else bfL1<false>(a, b, t[ti], tp[ti], neg1); }
          if (MODE == 11) { if (kk % 4 == 3) bfL1<true>(a, b, t[ti], tp[ti], neg1); else bfL1<false>(a, b, t[ti], tp[ti], neg1); }
          if (MODE == 12) bf3(a, b, t[ti]);  // canonical reference with the undoubled twiddle (checker for 7..11)
        }
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) x ^= (MODE >= 7 ? canon(v[j]) : v[j]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int MODE>
void run(const char *name, uint32_t *out, const uint32_t *tw, int clock_khz, int n_sm, int ctas_per_sm) {
  const int iters = 400, blocks = n_sm * ctas_per_sm, threads = 256;
  k<MODE><<<blocks, threads>>>(out, tw, 4);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, tw, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double warp_bf = (double)blocks * (threads / 32) * iters * 80.0;
  double clk = ms * 1e-3 * clock_khz * 1e3;
  printf("%-44s %d CTA/SM  %8.3f ms  %6.2f clk per warp-butterfly per SMSP\n", name, ctas_per_sm, ms,
         clk * (n_sm * 4) / warp_bf);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  printf("%s  SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clock_khz);
  int n = p.multiProcessorCount;
  uint32_t *out, *tw;
  cudaMalloc(&out, 4 * 256 * n * 8);
  uint32_t h[94];
  for (int j = 0; j < 62; j++) h[j] = (0x12345u * (j + 3) * 2654435761u) & 0x7ffffffeu;
  for (int j = 0; j < 31; j++) {  // canonical twiddles and their Shoup companions floor(t 2^32 / P)
    uint32_t t = h[j] % P31;
    h[62 + j] = t;
    h[31 + j] = (uint32_t)(((uint64_t)t << 32) / P31);
  }
  h[93] = 0xffffffffu;
  cudaMalloc(&tw, sizeof h);
  cudaMemcpy(tw, h, sizeof h, cudaMemcpyHostToDevice);
  {
    // the lazy forms against the canonical butterfly on the same inputs (xor of the canonicalised outputs per thread)
    const int nthr = 256 * n;
    uint32_t *ref = new uint32_t[nthr], *got = new uint32_t[nthr];
    k<12><<<n, 256>>>(out, tw, 3);
    cudaMemcpy(ref, out, 4 * nthr, cudaMemcpyDeviceToHost);
    auto check = [&](const char *name) {
      cudaMemcpy(got, out, 4 * nthr, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < nthr; i++) bad += got[i] != ref[i];
      printf("check %-10s %s (%d of %d threads differ)\n", name, bad ? "MISMATCH" : "ok", bad, nthr);
    };
    k<7><<<n, 256>>>(out, tw, 3); check("L1 alu-sub");
    k<8><<<n, 256>>>(out, tw, 3); check("L1 fma-sub");
    k<9><<<n, 256>>>(out, tw, 3); check("L1 mix");
    k<10><<<n, 256>>>(out, tw, 3); check("L1:L0 3:1");
    k<11><<<n, 256>>>(out, tw, 3); check("L1 mix 3:1");
    delete[] ref;
    delete[] got;
  }
  for (int c : {3, 6}) {
    run<0>("V0 WIDE+LEA.HI, IMAD adds (lde.cu)", out, tw, clock_khz, n, c);
    run<1>("V1 Shoup (IMAD.HI + 2 IMAD)", out, tw, clock_khz, n, c);
    run<2>("V2 adds on the ALU pipe", out, tw, clock_khz, n, c);
    run<3>("V3 undoubled twiddle (AND + shift reduce)", out, tw, clock_khz, n, c);
    run<4>("V4 mix V0:V1 2:1", out, tw, clock_khz, n, c);
    run<5>("V5 mix V0:V1 3:1", out, tw, clock_khz, n, c);
    run<6>("V6 mix V0:V1 1:1", out, tw, clock_khz, n, c);
    run<7>("L1 lazy Shoup, sub on ALU (6 instr)", out, tw, clock_khz, n, c);
    run<8>("L1 lazy Shoup, sub on FMA", out, tw, clock_khz, n, c);
    run<9>("L1 lazy Shoup, sub alternating", out, tw, clock_khz, n, c);
    run<10>("L1:L0 3:1 (L0 = lazy doubled twiddle)", out, tw, clock_khz, n, c);
    run<11>("L1 sub ALU:FMA 3:1", out, tw, clock_khz, n, c);
  }
  return 0;
}
