// Issue cost of the packed float FMA of sm_100a (fma.rn.f32x2, SASS FFMA2) beside the scalar FFMA:
//   A  8 independent FFMA chains per thread                      -> scalar FMA rate
//   B  8 independent FFMA2 chains per thread                     -> packed FMA rate (lane-level FMAs per second)
//   C  per round: 8 FFMA  + 8 integer adds (LOP3/IADD3 on the ALU pipe)
//   D  per round: 4 FFMA2 + 8 integer adds                       -> same lane-level FMAs as C in half the FMA instructions
// If a packed instruction costs ONE issue slot, D needs 12 slots per round against C's 16 and runs faster; if it holds the
// scheduler for two, C and D take the same time.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, unsigned* iout, int rounds, float a, float b) {
  float f[8]; float2 p[8]; unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x + i; p[i] = make_float2(f[i], f[i] + 1); u[i] = threadIdx.x * 7 + i; }
  const float2 aa = make_float2(a, a), bb = make_float2(b, b);
  for (int r = 0; r < rounds; ++r) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a, b);
    }
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], aa, bb);
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = __ffma2_rn(p[i], aa, bb);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (unsigned)r) + 0x9e3779b9u;
    }
  }
  float s = 0; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += f[i] + p[i].x + p[i].y; t ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
static float run(float* out, unsigned* iout, int rounds) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, iout, rounds, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iout, rounds, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out; unsigned* iout;
  cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&iout, 148 * 8 * 256 * 4);
  const int rounds = 20000;
  const double thr = 148.0 * 8 * 256;
  const float a = run<0>(out, iout, rounds), b = run<1>(out, iout, rounds), c = run<2>(out, iout, rounds), d = run<3>(out, iout, rounds);
  printf("A  8 FFMA            per round: %.3f ms  %.1f TFLOP/s\n", a, 2 * 8 * thr * rounds / a * 1e-9);
  printf("B  8 FFMA2           per round: %.3f ms  %.1f TFLOP/s\n", b, 2 * 16 * thr * rounds / b * 1e-9);
  printf("C  8 FFMA  + 8 int   per round: %.3f ms\n", c);
  printf("D  4 FFMA2 + 8 int   per round: %.3f ms   (D / C = %.3f; 0.75 = one issue slot per FFMA2, 1.0 = two)\n", d, d / c);
  return cudaDeviceSynchronize() != cudaSuccess;
}
