// Warp-level mma.sync.m16n8k16 (f16 x f16 -> f32) issue rate on B200: the tensor path the single-pass GN/CG operator uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hmma_bench tools/hmma_bench.cu && tools/hmma_bench
// Prints cycles per instruction per SM for 4..16 warps per SM with 1, 2, 5 independent accumulator chains per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(int iters, float *out, long long *cyc) {
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int u = 0; u < 4; ++u) d[c][u] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) for (int u = 0; u < 4; ++u) s += d[c][u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
void run(int warps) {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<CHAINS><<<148, warps * 32>>>(iters, out, cyc);
  k<CHAINS><<<148, warps * 32>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_sm = (double)h / ((double)iters * CHAINS * warps);
  printf("warps/SM %2d chains %d: %.2f cycles per mma per SM  -> %.0f MAC/clk/SM (%s)\n", warps, CHAINS, per_sm, 2048.0 / per_sm,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) { run<1>(w); run<2>(w); run<5>(w); }
  return 0;
}
