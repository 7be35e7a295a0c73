// Microbenchmark: issue rates of the integer instructions BLAKE2s is made of, alone and mixed,
// on sm_100a.  Prints warp-instructions per clock per SM sub-partition (SMSP).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <string>

#define REP8(x) x x x x x x x x
#define REP64(x) REP8(REP8(x))

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t k1, uint32_t k2, int iters) {
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  uint32_t b = blockIdx.x | 1;
  unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {  // IADD3 (3-input add)
      REP8(asm volatile("{add.u32 %0, %0, %8; add.u32 %0, %0, %9; add.u32 %1, %1, %8; add.u32 %1, %1, %9; add.u32 %2, %2, %8; add.u32 %2, %2, %9; add.u32 %3, %3, %8; add.u32 %3, %3, %9;"
                        "add.u32 %4, %4, %8; add.u32 %4, %4, %9; add.u32 %5, %5, %8; add.u32 %5, %5, %9; add.u32 %6, %6, %8; add.u32 %6, %6, %9; add.u32 %7, %7, %8; add.u32 %7, %7, %9;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 1) {  // LOP3
      REP8(asm volatile("{lop3.b32 %0, %0, %8, %9, 0x96; lop3.b32 %1, %1, %8, %9, 0x96; lop3.b32 %2, %2, %8, %9, 0x96; lop3.b32 %3, %3, %8, %9, 0x96;"
                        "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 2) {  // SHF
      REP8(asm volatile("{shf.r.wrap.b32 %0, %0, %0, 12; shf.r.wrap.b32 %1, %1, %1, 12; shf.r.wrap.b32 %2, %2, %2, 12; shf.r.wrap.b32 %3, %3, %3, 12;"
                        "shf.r.wrap.b32 %4, %4, %4, 7; shf.r.wrap.b32 %5, %5, %5, 7; shf.r.wrap.b32 %6, %6, %6, 7; shf.r.wrap.b32 %7, %7, %7, 7;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7));)
    } else if (MODE == 3) {  // PRMT
      REP8(asm volatile("{prmt.b32 %0, %0, %0, 0x1032; prmt.b32 %1, %1, %1, 0x1032; prmt.b32 %2, %2, %2, 0x1032; prmt.b32 %3, %3, %3, 0x1032;"
                        "prmt.b32 %4, %4, %4, 0x0321; prmt.b32 %5, %5, %5, 0x0321; prmt.b32 %6, %6, %6, 0x0321; prmt.b32 %7, %7, %7, 0x0321;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7));)
    } else if (MODE == 4) {  // IMAD (lo) with a runtime multiplier
      REP8(asm volatile("{mad.lo.u32 %0, %0, %9, %8; mad.lo.u32 %1, %1, %9, %8; mad.lo.u32 %2, %2, %9, %8; mad.lo.u32 %3, %3, %9, %8;"
                        "mad.lo.u32 %4, %4, %9, %8; mad.lo.u32 %5, %5, %9, %8; mad.lo.u32 %6, %6, %9, %8; mad.lo.u32 %7, %7, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 5) {  // IMAD.WIDE (both halves fed back so nothing is dead or loop-invariant)
      REP8(asm volatile("{.reg .b64 t; mul.wide.u32 t, %0, %8; mov.b64 {%0, %1}, t; mul.wide.u32 t, %2, %9; mov.b64 {%2, %3}, t; mul.wide.u32 t, %4, %8; mov.b64 {%4, %5}, t; mul.wide.u32 t, %6, %9; mov.b64 {%6, %7}, t; mul.wide.u32 t, %1, %8; mov.b64 {%1, %0}, t; mul.wide.u32 t, %3, %9; mov.b64 {%3, %2}, t; mul.wide.u32 t, %5, %8; mov.b64 {%5, %4}, t; mul.wide.u32 t, %7, %9; mov.b64 {%7, %6}, t; }"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(k1), "r"(k2));)
    } else if (MODE == 6) {  // 1:1 LOP3 : IMAD
      REP8(asm volatile("{lop3.b32 %0, %0, %8, %9, 0x96; mad.lo.u32 %4, %4, %9, %8; lop3.b32 %1, %1, %8, %9, 0x96; mad.lo.u32 %5, %5, %9, %8;"
                        "lop3.b32 %2, %2, %8, %9, 0x96; mad.lo.u32 %6, %6, %9, %8; lop3.b32 %3, %3, %8, %9, 0x96; mad.lo.u32 %7, %7, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 7) {  // 2:1 ALU (LOP3, PRMT) : IMAD
      REP8(asm volatile("{lop3.b32 %0, %0, %8, %9, 0x96; prmt.b32 %1, %1, %1, 0x1032; mad.lo.u32 %4, %4, %9, %8; lop3.b32 %2, %2, %8, %9, 0x96; prmt.b32 %3, %3, %3, 0x1032; mad.lo.u32 %5, %5, %9, %8;"
                        "lop3.b32 %0, %0, %8, %9, 0x96; prmt.b32 %1, %1, %1, 0x0321; mad.lo.u32 %6, %6, %9, %8; lop3.b32 %2, %2, %8, %9, 0x96; prmt.b32 %3, %3, %3, 0x0321; mad.lo.u32 %7, %7, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 8) {  // 2:1 LOP3 : IMAD.WIDE
      REP8(asm volatile("{.reg .b64 t; lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; mul.wide.u32 t, %0, %8; mov.b64 {%0, %1}, t; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; mul.wide.u32 t, %2, %9; mov.b64 {%2, %3}, t; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mul.wide.u32 t, %1, %8; mov.b64 {%1, %0}, t; lop3.b32 %7, %7, %8, %9, 0x96; lop3.b32 %4, %4, %8, %9, 0x96; mul.wide.u32 t, %3, %9; mov.b64 {%3, %2}, t; }"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(k1), "r"(k2));)
    } else if (MODE == 9) {  // 1:1 LOP3 : IMAD.WIDE
      REP8(asm volatile("{.reg .b64 t; lop3.b32 %4, %4, %8, %9, 0x96; mul.wide.u32 t, %0, %8; mov.b64 {%0, %1}, t; lop3.b32 %5, %5, %8, %9, 0x96; mul.wide.u32 t, %2, %9; mov.b64 {%2, %3}, t; lop3.b32 %6, %6, %8, %9, 0x96; mul.wide.u32 t, %1, %8; mov.b64 {%1, %0}, t; lop3.b32 %7, %7, %8, %9, 0x96; mul.wide.u32 t, %3, %9; mov.b64 {%3, %2}, t; }"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(k1), "r"(k2));)
    } else if (MODE == 10) {  // IMAD.HI
      REP8(asm volatile("{mad.hi.u32 %0, %0, %9, %8; mad.hi.u32 %1, %1, %9, %8; mad.hi.u32 %2, %2, %9, %8; mad.hi.u32 %3, %3, %9, %8;"
                        "mad.hi.u32 %4, %4, %9, %8; mad.hi.u32 %5, %5, %9, %8; mad.hi.u32 %6, %6, %9, %8; mad.hi.u32 %7, %7, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 12) {  // 2:1 LOP3 : IMAD.HI
      REP8(asm volatile("{lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; mad.hi.u32 %0, %0, %9, %8; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %1, %1, %9, %8;"
                        "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; mad.hi.u32 %2, %2, %9, %8; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %3, %3, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 13) {  // 4:1 LOP3 : IMAD.HI
      REP8(asm volatile("{lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %0, %0, %9, %8;"
                        "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %1, %1, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 14) {  // 4:1 LOP3 : IMAD.WIDE
      REP8(asm volatile("{.reg .b64 t; lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mul.wide.u32 t, %0, %8; mov.b64 {%0, %1}, t;"
                        "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96; mul.wide.u32 t, %2, %9; mov.b64 {%2, %3}, t;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(k1), "r"(k2));)
    } else if (MODE == 15) {  // 4:2:1 LOP3 : IMAD : IMAD.HI
      REP8(asm volatile("{lop3.b32 %4, %4, %8, %9, 0x96; mad.lo.u32 %2, %2, %9, %8; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; mad.lo.u32 %3, %3, %9, %8; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %0, %0, %9, %8;"
                        "lop3.b32 %4, %4, %8, %9, 0x96; mad.lo.u32 %2, %2, %9, %8; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; mad.lo.u32 %3, %3, %9, %8; lop3.b32 %7, %7, %8, %9, 0x96; mad.hi.u32 %1, %1, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    } else if (MODE == 11) {  // 3:1 ALU : IMAD
      REP8(asm volatile("{lop3.b32 %0, %0, %8, %9, 0x96; prmt.b32 %1, %1, %1, 0x1032; shf.r.wrap.b32 %2, %2, %2, 12; mad.lo.u32 %4, %4, %9, %8; lop3.b32 %3, %3, %8, %9, 0x96; prmt.b32 %0, %0, %0, 0x1032; shf.r.wrap.b32 %1, %1, %1, 7; mad.lo.u32 %5, %5, %9, %8;"
                        "lop3.b32 %2, %2, %8, %9, 0x96; prmt.b32 %3, %3, %3, 0x0321; shf.r.wrap.b32 %0, %0, %0, 12; mad.lo.u32 %6, %6, %9, %8; lop3.b32 %1, %1, %8, %9, 0x96; prmt.b32 %2, %2, %2, 0x0321; shf.r.wrap.b32 %3, %3, %3, 7; mad.lo.u32 %7, %7, %9, %8;}"
                        : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(b), "r"(k1));)
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ (uint32_t)w0 ^ (uint32_t)w1 ^ (uint32_t)w2 ^ (uint32_t)w3 ^ (uint32_t)(w0 >> 32) ^ (uint32_t)(w1 >> 32) ^ (uint32_t)(w2 >> 32) ^ (uint32_t)(w3 >> 32);
}

template <int MODE>
void run(const char *name, int per_iter, uint32_t *out, int clock_khz, int n_sm) {
  const int iters = 2000, blocks = n_sm * 8, threads = 256;
  k<MODE><<<blocks, threads>>>(out, 1u << 20, 1u << 25, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, 1u << 20, 1u << 25, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double warp_instr = (double)blocks * (threads / 32) * iters * 8.0 * per_iter;
  double clk = ms * 1e-3 * clock_khz * 1e3;
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SMSP\n", name, ms, warp_instr / clk / (n_sm * 4));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clock_khz = 0; cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  printf("%s  SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clock_khz);
  uint32_t *out; cudaMalloc(&out, 4 * 256 * p.multiProcessorCount * 8);
  int n = p.multiProcessorCount;
  run<0>("IADD3", 8, out, clock_khz, n);
  run<1>("LOP3", 8, out, clock_khz, n);
  run<2>("SHF", 8, out, clock_khz, n);
  run<3>("PRMT", 8, out, clock_khz, n);
  run<4>("IMAD.lo", 8, out, clock_khz, n);
  run<5>("IMAD.WIDE", 8, out, clock_khz, n);
  run<10>("IMAD.HI", 8, out, clock_khz, n);
  run<6>("LOP3:IMAD 1:1", 8, out, clock_khz, n);
  run<7>("ALU:IMAD 2:1", 12, out, clock_khz, n);
  run<11>("ALU:IMAD 3:1", 16, out, clock_khz, n);
  run<8>("LOP3:IMAD.WIDE 2:1", 12, out, clock_khz, n);
  run<9>("LOP3:IMAD.WIDE 1:1", 8, out, clock_khz, n);
  run<14>("LOP3:IMAD.WIDE 4:1", 10, out, clock_khz, n);
  run<12>("LOP3:IMAD.HI 2:1", 12, out, clock_khz, n);
  run<13>("LOP3:IMAD.HI 4:1", 10, out, clock_khz, n);
  run<15>("LOP3:IMAD:IMAD.HI 4:2:1", 14, out, clock_khz, n);
  return 0;
}
