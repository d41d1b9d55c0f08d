// FP64 red.global.add throughput on scattered addresses (the digestion's L2 atomic roofline).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atomic_bench tools/atomic_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_red(double* f, size_t n, int per_thread, unsigned seed, int same) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
  unsigned wbase = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 40503u + seed;
  for (int i = 0; i < per_thread; ++i) {
    x = x * 1664525u + 1013904223u;
    size_t a = (size_t)(x >> 4) % n;
    if (same && (i % 4 == 0)) { wbase = wbase * 1664525u + 1013904223u; a = (size_t)(wbase >> 4) % n; }  // warp-uniform address
    atomicAdd(f + a, 1.0);
  }
}
__global__ void k_ld(const double* __restrict__ f, size_t n, int per_thread, unsigned seed, double* out) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + seed;
  double s = 0;
  for (int i = 0; i < per_thread; ++i) {
    x = x * 1664525u + 1013904223u;
    s += __ldg(f + (size_t)(x >> 4) % n);
  }
  if (s == 1.2345) out[0] = s;
}
int main() {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (size_t mb : {14, 55, 220}) {
    size_t n = mb * 1000000 / 8;
    double* f; cudaMalloc(&f, n * 8); cudaMemset(f, 0, n * 8);
    for (int same = 0; same < 2; ++same) {
      int nb = 148 * 16, nt = 256, per = 256;
      k_red<<<nb, nt>>>(f, n, per, 1, same);
      cudaEventRecord(e0);
      k_red<<<nb, nt>>>(f, n, per, 7, same);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("red.f64 %3zu MB array, %s: %.1f G atomics/s\n", mb, same ? "1 of 4 warp-uniform" : "all scattered", (double)nb * nt * per / ms / 1e6);
    }
    int nb = 148 * 16, nt = 256, per = 256;
    k_ld<<<nb, nt>>>(f, n, per, 1, f);
    cudaEventRecord(e0);
    k_ld<<<nb, nt>>>(f, n, per, 7, f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg.f64 %3zu MB array scattered: %.1f G loads/s\n", mb, (double)nb * nt * per / ms / 1e6);
    cudaFree(f);
  }
  return 0;
}
