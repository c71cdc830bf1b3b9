// ubench.cu — FP64 micro-benchmarks that the land-kernel design rests on (development aid, not part of the product):
// dependent-issue latency of DFMA/DADD/DMUL, of a shared-memory table look-up, and FP64-pipe throughput versus the
// number of resident warps per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double* out, long long* cycles, int iters, int mode) {
  __shared__ double tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
  __syncthreads();
  double a = 1.0 + threadIdx.x * 1e-12, m = 1.0000000001, c = 1e-12;
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < iters; ++i) a = fma(a, m, c);
  else if (mode == 1) for (int i = 0; i < iters; ++i) a = a + c;
  else if (mode == 2) for (int i = 0; i < iters; ++i) a = a * m;
  else if (mode == 3) for (int i = 0; i < iters; ++i) a = tab[__double2loint(a) & 63] + c;   // LDS + DADD
  else if (mode == 4) for (int i = 0; i < iters; ++i) a = (a > c) ? a * m : c;                // DSETP + DMUL + SEL
  else if (mode == 5) for (int i = 0; i < iters; ++i) a = 1.0 / a + c;                        // full division
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (a == 123.456) out[0] = a;
}

__global__ void thr_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) { a0 = fma(a0, m, c); a1 = fma(a1, m, c); }
  if (a0 + a1 == 123.456) out[0] = a0;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 1024);
  const char* names[] = {"DFMA", "DADD", "DMUL", "LDS+DADD", "DSETP+DMUL+SEL", "1/x + c"};
  const int iters = 4096;
  for (int mode = 0; mode < 6; ++mode) {
    lat_kernel<<<1, 32>>>(out, cyc, iters, mode);
    lat_kernel<<<1, 32>>>(out, cyc, iters, mode);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent latency %-16s : %.2f cycles per iteration\n", names[mode], (double)h / iters);
  }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 1; warps <= 16; warps *= 2) {   // warps per scheduler, 2 independent DFMA chains per thread
    const int threads = warps * 4 * 32 > 1024 ? 1024 : warps * 4 * 32;
    const int blocks_per_sm = warps * 4 * 32 / threads;
    const int it = 1 << 16;
    thr_kernel<<<sms * blocks_per_sm, threads>>>(out, it);
    cudaEventRecord(e0);
    thr_kernel<<<sms * blocks_per_sm, threads>>>(out, it);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma_s = 2.0 * it * (double)threads * blocks_per_sm * sms / (ms * 1e-3);
    printf("throughput %2d warps/scheduler x 2 chains: %.2f T DFMA/s (%.1f TFLOP/s)\n", warps, fma_s / 1e12, 2 * fma_s / 1e12);
  }
  return 0;
}
