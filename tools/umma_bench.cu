// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands in 128B-swizzled smem) as a function of N,
// of the A operand's major-ness and of whether consecutive MMAs share an accumulator.  nvcc -arch=sm_100a -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../frtm_vos_b200/csrc/tc_ptx.cuh"
using namespace frtm;

__device__ __forceinline__ uint64_t desc_ls(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int N, bool MN_MAJOR, int NACC, int NCOMMIT = 0>
__global__ void __launch_bounds__(128) bench(long long *out, int reps) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t *gen = raw + (base - smem_u32(raw));
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t *>(gen)[i] = 0;
  __shared__ uint32_t tslot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t dummy[4];
  const int warp = uniform_warp_idx();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[i]), 1000000);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc(128, N) | (MN_MAJOR ? (1u << 15) : 0u);
    const uint64_t a = MN_MAJOR ? desc_ls(base, 16384, 1024) : desc_ls(base, 0, 1024);
    const uint64_t b = desc_ls(base + 32768, 0, 1024);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem + ((r * 4 + k) % NACC) * N % 512, a + (MN_MAJOR ? (uint64_t)(k * 2048 >> 4) : (uint64_t)(k * 32 >> 4)), b + (uint64_t)(k * 32 >> 4), idesc, 1u);
          }
#pragma unroll
          for (int cidx = 0; cidx < NCOMMIT; ++cidx) umma_commit(smem_u32(&dummy[cidx]));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int N, bool MN, int NACC, int NCOMMIT = 0>
void run(const char *name, int blocks) {
  long long *d; cudaMalloc(&d, 8 * blocks);
  const int reps = 256, smem = 66 * 1024;
  cudaFuncSetAttribute(bench<N, MN, NACC, NCOMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N, MN, NACC, NCOMMIT><<<blocks, 128, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, d, 8 * blocks, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-34s N=%3d acc=%d ctas=%3d : %7.1f cycles per MMA (%s)\n", name, N, NACC, blocks, (double)mx / (reps * 4), cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16, false, 1>("K-major A, same accumulator", 1);
  run<32, false, 1>("K-major A, same accumulator", 1);
  run<64, false, 1>("K-major A, same accumulator", 1);
  run<128, false, 1>("K-major A, same accumulator", 1);
  run<256, false, 1>("K-major A, same accumulator", 1);
  run<16, false, 4>("K-major A, 4 accumulators", 1);
  run<64, false, 4>("K-major A, 4 accumulators", 1);
  run<128, false, 2>("K-major A, 2 accumulators", 1);
  run<16, true, 1>("MN-major A, same accumulator", 1);
  run<32, true, 1>("MN-major A, same accumulator", 1);
  run<64, true, 1>("MN-major A, same accumulator", 1);
  run<64, false, 1>("K-major A, all SMs", 148);
  run<128, false, 1>("K-major A, all SMs", 148);
  run<256, false, 1>("K-major A, all SMs", 148);
  run<64, false, 1>("K-major A, 2 CTAs/SM", 296);
  run<32, false, 1, 1>("4 MMAs + 1 commit per group", 1);
  run<32, false, 1, 2>("4 MMAs + 2 commits per group", 1);
  run<32, false, 1, 3>("4 MMAs + 3 commits per group", 1);
  run<64, false, 1, 1>("4 MMAs + 1 commit per group", 1);
  return 0;
}
