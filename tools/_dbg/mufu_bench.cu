// microbenchmark: MUFU.EX2 throughput per SM vs warps per SM sub-partition, alone and inside a softmax-like mix
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float scale) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
  float l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  unsigned pk = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) {
        a[i] = ex2(a[i]); a[i + 1] = ex2(a[i + 1]);
      } else {
        float p0 = ex2(fmaf(a[i], scale, -0.5f)), p1 = ex2(fmaf(a[i + 1], scale, -0.5f));
        if ((i & 6) == 0) l0 += p0 + p1; else if ((i & 6) == 2) l1 += p0 + p1; else if ((i & 6) == 4) l2 += p0 + p1; else l3 += p0 + p1;
        unsigned h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(p1), "f"(p0));
        pk ^= h;
        a[i] = p0 * 0.5f; a[i + 1] = p1 * 0.5f;
      }
    }
  }
  long long t1 = clock64();
  float s = l0 + l1 + l2 + l3 + __uint_as_float(pk);
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 2048;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 12, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, clk, iters, 0.1f); else k<1><<<148, warps * 32>>>(out, clk, iters, 0.1f);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double c = h[0];
      double n_ex2 = double(iters) * 16 * warps * 32;
      printf("mode %d warps/SM %2d (%d per SMSP): %.0f clk -> %.2f ex2/clk/SM, %.1f clk per warp-MUFU per SMSP\n", mode, warps, warps / 4, c,
             n_ex2 / c, c / (double(iters) * 16 * (warps / 4)));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
