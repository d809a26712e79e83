// Development probe (not part of libbpv): does a TILED tensor-map load (cp.async.bulk.tensor, UTMALDG) with
// CU_TENSOR_MAP_L2_PROMOTION_NONE fill L2 in 32-byte sectors, i.e. can it cut F1's DRAM over-fetch below the 64-byte
// granule floor that ld/cp.async .L2::64B reach?  And how fast does TMA move F1's access pattern (288 B x 65 rows at a
// 5760 B pitch, one ROI per CTA, 9 CTAs per SM)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_probe tools/tma_probe.cu
//   ./tma_probe <promotion 0|1|2|3> <frames> <iters> <mode: 0 = tma, 1 = cp.async L2::64B> [hint 0|1]
// Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` for the traffic.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int H = 1080, W3 = 5760;
constexpr int BW = 96 * 3, BH = 65;          // forehead box of config 2
constexpr int SPLIT = 256;                   // first strip 256 B, second 32 B
constexpr int STAGE = 24 * 1024;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void box_of(long long roi, int& x, int& y, int& f) {
  unsigned h = (unsigned)roi * 2654435761u;
  f = (int)(roi >> 1);
  const int which = (int)(roi & 1);
  x = (which ? 1270 : 864) * 3 + 3 * (int)(h % 5u);          // unaligned byte offsets, jittered
  y = (which ? 770 : 260) + (int)((h >> 8) % 5u);
}

template <bool HINT>
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                                                 unsigned long long* out, int xalign) {
  extern __shared__ __align__(1024) uint8_t stage_raw[];
  uint8_t* stage = stage_raw + ((1024u - (s32(stage_raw) & 1023u)) & 1023u);     // TMA destination: 128-byte aligned
  unsigned long long* barp = reinterpret_cast<unsigned long long*>(stage + STAGE);
  int x, y, f;
  box_of(blockIdx.x, x, y, f);
  if (xalign) x &= ~15;
  const uint32_t b = s32(barp);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"((SPLIT + 32) * BH) : "memory");
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (HINT) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                   :: "r"(s32(stage)), "l"(&mA), "r"(x), "r"(y), "r"(f), "r"(b), "l"(pol) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                   :: "r"(s32(stage) + SPLIT * BH), "l"(&mB), "r"(x + SPLIT), "r"(y), "r"(f), "r"(b), "l"(pol) : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(s32(stage)), "l"(&mA), "r"(x), "r"(y), "r"(f), "r"(b) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(s32(stage) + SPLIT * BH), "l"(&mB), "r"(x + SPLIT), "r"(y), "r"(f), "r"(b) : "memory");
    }
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b) : "memory");
  uint32_t s = 0;
  for (int v = threadIdx.x; v < (SPLIT + 32) * BH / 16; v += 128) {
    const uint4 d = *reinterpret_cast<const uint4*>(stage + 16 * v);
    s = __dp4a(d.x, 0x01010101u, __dp4a(d.y, 0x01010101u, __dp4a(d.z, 0x01010101u, __dp4a(d.w, 0x01010101u, s))));
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.x, (unsigned long long)s);
}

__global__ void __launch_bounds__(128) cpasync_kernel(const uint8_t* frames, unsigned long long* out, int xalign) {
  extern __shared__ __align__(128) uint8_t stage[];
  int x, y, f;
  box_of(blockIdx.x, x, y, f);
  if (xalign) x &= ~15;
  const uint8_t* base = frames + (long long)f * H * W3 + (long long)y * W3 + x;
  const int off = (int)((uintptr_t)base & 15);
  const int vpr = (off + BW + 15) >> 4;                       // 19 or 20
  const int rps = 128 / vpr, r0 = threadIdx.x / vpr, v0 = threadIdx.x - r0 * vpr;
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint32_t s = 0;
  if (r0 < rps) {
    const uint32_t slot = s32(stage) + 16u * threadIdx.x;
    int k = 0;
    for (int r = r0; r < BH; r += rps, ++k)
      asm volatile("cp.async.cg.shared.global.L2::cache_hint.L2::64B [%0], [%1], 16, %2;"
                   :: "r"(slot + (uint32_t)k * 2048u), "l"(base - off + (long long)r * W3 + 16 * v0), "l"(pol) : "memory");
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    k = 0;
    for (int r = r0; r < BH; r += rps, ++k) {
      const uint4 d = *reinterpret_cast<const uint4*>(stage + 16 * threadIdx.x + k * 2048);
      s = __dp4a(d.x, 0x01010101u, __dp4a(d.y, 0x01010101u, __dp4a(d.z, 0x01010101u, __dp4a(d.w, 0x01010101u, s))));
    }
  }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.x, (unsigned long long)s);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int promo = argc > 1 ? atoi(argv[1]) : 0;
  const long long NF = argc > 2 ? atoll(argv[2]) : 2048;
  const int iters = argc > 3 ? atoi(argv[3]) : 10;
  const int mode = argc > 4 ? atoi(argv[4]) : 0;
  const int hint = argc > 5 ? atoi(argv[5]) : 1;
  const int xalign = argc > 6 ? atoi(argv[6]) : 0;
  uint8_t* frames;
  CK(cudaMalloc(&frames, NF * H * W3));
  CK(cudaMemset(frames, 1, NF * H * W3));
  unsigned long long* out;
  CK(cudaMalloc(&out, NF * 2 * 8));
  CK(cudaMemset(out, 0, NF * 2 * 8));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap mA, mB;
  cuuint64_t dims[3] = {(cuuint64_t)W3, (cuuint64_t)H, (cuuint64_t)NF};
  cuuint64_t strides[2] = {(cuuint64_t)W3, (cuuint64_t)H * W3};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t boxA[3] = {SPLIT, BH, 1}, boxB[3] = {32, BH, 1};
  CUresult r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, frames, dims, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, frames, dims, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  CK(cudaFuncSetAttribute(tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE + 2048));
  CK(cudaFuncSetAttribute(tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE + 2048));
  CK(cudaFuncSetAttribute(cpasync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f, sum = 0;
  for (int i = 0; i < iters + 2; ++i) {
    CK(cudaEventRecord(e0));
    if (mode == 0) { if (hint) tma_kernel<true><<<(unsigned)(NF * 2), 128, STAGE + 1024 + 16>>>(mA, mB, out, xalign); else tma_kernel<false><<<(unsigned)(NF * 2), 128, STAGE + 1024 + 16>>>(mA, mB, out, xalign); }
    else cpasync_kernel<<<(unsigned)(NF * 2), 128, STAGE>>>(frames, out, xalign);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (i >= 2) { best = ms < best ? ms : best; sum += ms; }
  }
  CK(cudaGetLastError());
  unsigned long long h[2];
  CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
  const double bytes = (double)NF * 2 * BW * BH;
  printf("mode %d promo %d hint %d frames %lld: alg %.1f MB  mean %.1f us  best %.1f us -> %.0f GB/s (best)   check %llu %llu (expect %d per launch)\n",
         mode, promo, hint, NF, bytes / 1e6, sum / iters * 1e3, best * 1e3, bytes / best / 1e6, h[0], h[1], BW * BH);
  return 0;
}
