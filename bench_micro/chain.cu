// Microbenchmark: latency of a CHAIN of dependent BLAKE2s compressions (the Fiat-Shamir / tree-depth chain that bounds
// single-blob latency, fri_small.cu) for 1, 2 and 4 hashing warps per scheduler, message in registers.
// Prints microseconds per link.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../frieda_b200/csrc/blake2s.cuh"
using namespace frieda;

template <bool IMAD_ADDS>
__global__ void chain(uint32_t *out, int links, uint32_t one) {
  uint32_t m[16], h[8];
#pragma unroll
  for (int i = 0; i < 16; i++) m[i] = threadIdx.x * 16 + i + blockIdx.x;
#pragma unroll 1
  for (int l = 0; l < links; l++) {
    if (IMAD_ADDS) merkle_hash_node(m, h, one); else merkle_hash_node(m, h, 1u);
#pragma unroll
    for (int i = 0; i < 8; i++) { m[i] = h[i]; m[8 + i] ^= h[i]; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = h[0] ^ h[7];
}

template <bool IMAD_ADDS>
void run(const char *name, int threads, int blocks, uint32_t *out) {
  const int links = 2000;
  chain<IMAD_ADDS><<<blocks, threads>>>(out, 10, 1u);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain<IMAD_ADDS><<<blocks, threads>>>(out, links, 1u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s %4d threads x %4d CTAs: %7.3f us per link\n", name, threads, blocks, ms * 1e3 / links);
}

int main() {
  uint32_t *out; cudaMalloc(&out, 4 * 1024 * 1024);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock %d kHz\n", clk);
  for (int rep = 0; rep < 2; rep++) {
    run<true>("IMAD adds (product kernels)", 32, 1, out);
    run<false>("compiler's add placement", 32, 1, out);
    run<true>("IMAD adds", 128, 1, out);
    run<true>("IMAD adds", 256, 1, out);
    run<true>("IMAD adds", 512, 1, out);
    run<true>("IMAD adds", 128, 148, out);
    run<true>("IMAD adds", 128, 592, out);
    run<false>("compiler's add placement", 128, 592, out);
  }
  return 0;
}
