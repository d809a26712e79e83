// Development probe (not part of libbpv): how fast does the copy engine move ROI boxes out of PINNED HOST frames with one
// pitched 3-D copy per (stream, ROI) — extent (box bytes, box rows, T frames), slice pitch = one frame — compared with the
// zero-copy ROI kernel's 33 GB/s of ROI bytes (49 GB/s on the wire at 128-byte requests)?
//   nvcc -O3 -o tools/bin/dma_probe tools/dma_probe.cu && tools/bin/dma_probe [streams] [T] [iters]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 256, T = argc > 2 ? atoi(argv[2]) : 8, iters = argc > 3 ? atoi(argv[3]) : 10;
  const int H = 1080, W3 = 5760;
  const size_t frame = (size_t)H * W3;
  uint8_t* host;
  CK(cudaHostAlloc(&host, (size_t)S * T * frame, cudaHostAllocDefault));
  memset(host, 7, (size_t)S * T * frame);
  // union boxes: forehead 100 x 69 px, palm 81 x 75 px (config-2 boxes + the +-2 px jitter)
  const int bw[2] = {100 * 3, 81 * 3}, bh[2] = {69, 75}, bx[2] = {862 * 3, 1268 * 3}, by[2] = {258, 768};
  size_t per_stream = 0;
  for (int r = 0; r < 2; ++r) per_stream += (size_t)((bw[r] + 15) / 16 * 16) * bh[r] * T;
  uint8_t* dev;
  CK(cudaMalloc(&dev, per_stream * S));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double bytes = 0;
  for (int r = 0; r < 2; ++r) bytes += (double)bw[r] * bh[r] * T * S;
  for (int it = 0; it < iters + 1; ++it) {
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(e0, st));
    size_t off = 0;
    for (int s = 0; s < S; ++s)
      for (int r = 0; r < 2; ++r) {
        const int pitch = (bw[r] + 15) / 16 * 16;
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(host + (size_t)s * T * frame, W3, W3, H);
        p.srcPos = make_cudaPos(bx[r] + 3 * (s % 5), by[r] + (s % 3), 0);
        p.dstPtr = make_cudaPitchedPtr(dev + off, pitch, pitch, bh[r]);
        p.dstPos = make_cudaPos(0, 0, 0);
        p.extent = make_cudaExtent(bw[r], bh[r], T);
        p.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3DAsync(&p, st));
        off += (size_t)pitch * bh[r] * T;
      }
    CK(cudaEventRecord(e1, st));
    auto t1 = std::chrono::steady_clock::now();
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double enq = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (it) printf("S=%d T=%d: %d copies, %.1f MB of ROI bytes: enqueue %.3f ms (%.2f us per call), device %.3f ms -> %.1f GB/s, %.0f frames/s\n",
                   S, T, 2 * S, bytes / 1e6, enq, enq * 1e3 / (2 * S), ms, bytes / ms / 1e6, (double)S * T / ms * 1e3);
  }
  // reference: one contiguous pinned copy of the same number of bytes
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0, st));
    CK(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it) printf("contiguous pinned copy of %.1f MB: %.3f ms -> %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
  }
  return 0;
}
