// Micro-benchmark of the MMA *issue path*: descriptors change every MMA (as in conv.cu, where the
// activation descriptor is stepped per tap / K step).  Variant 0: one thread in a divergent branch
// (values in vector registers -> R2UR before every UTCHMMA).  Variant 1: the whole warp runs the loop,
// loop-invariant bases are made provably warp-uniform with redux.sync, and the election + MMA live
// inside one asm statement, so the stepping can stay on the uniform datapath.
// usage: umma_issue_bench variant [grid] [iters]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../muzero_b200/csrc/umma.cuh"
using namespace mz::umma;

__device__ __forceinline__ void bench_mma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void bench_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

template <int kVariant>
__global__ void __launch_bounds__(128) bench(int iters, int a_rows, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16 * a_rows * 16;
  for (int i = tid; i < (16 * a_rows + 8 * 128) * 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = instr_desc_f16(128, 128);
  const uint64_t a_t = smem_desc(0, a_rows * 16, 128), b_t = smem_desc(0, 128 * 16, 128);
  const uint32_t a_hi = (uint32_t)(a_t >> 32), b_hi = (uint32_t)(b_t >> 32);
  auto d64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
  if (warp == 1) {
    if (kVariant == 0) {
      if (lane == 0) {
        const uint32_t tm = tmem_base;
        const uint32_t a0 = (uint32_t)a_t + (smem_u32(sA) >> 4) + 11, b0 = (uint32_t)b_t + (smem_u32(sB) >> 4);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t a_lo = a0 + (uint32_t)((tap / 3 - 1) * 10 + (tap % 3 - 1));
            uint32_t b_lo = b0;
            for (int ks = 0; ks < 4; ++ks) {
              mma_f16(tm, d64(a_lo, a_hi), d64(b_lo, b_hi), idesc, 1);
              mma_f16(tm + 128, d64(a_lo + 128, a_hi), d64(b_lo, b_hi), idesc, 1);
              a_lo += 2 * a_rows; b_lo += 256;
            }
          }
        }
        commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
      }
    } else {
      // whole warp, provably uniform bases
      const uint32_t tm = __reduce_max_sync(0xffffffffu, tmem_base);
      const uint32_t a0 = __reduce_max_sync(0xffffffffu, (uint32_t)a_t + (smem_u32(sA) >> 4) + 11);
      const uint32_t b0 = __reduce_max_sync(0xffffffffu, (uint32_t)b_t + (smem_u32(sB) >> 4));
      const uint32_t arows2 = __reduce_max_sync(0xffffffffu, (uint32_t)(2 * a_rows));
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int tap = 0; tap < 9; ++tap) {
          uint32_t a_lo = a0 + (uint32_t)((tap / 3 - 1) * 10 + (tap % 3 - 1));
          uint32_t b_lo = b0;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            bench_mma_elect(tm, d64(a_lo, a_hi), d64(b_lo, b_hi), idesc, 1);
            bench_mma_elect(tm + 128, d64(a_lo + 128, a_hi), d64(b_lo, b_hi), idesc, 1);
            a_lo += arows2; b_lo += 256;
          }
        }
      }
      bench_commit_elect(&bar);
      mbar_wait(&bar, 0);
      if (lane == 0) out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0, grid = argc > 2 ? atoi(argv[2]) : 1, iters = argc > 3 ? atoi(argv[3]) : 2000;
  const int a_rows = 279;
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 16 * a_rows * 16 + 8 * 128 * 16 + 1024;
  if (variant == 0) { cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); bench<0><<<grid, 128, smem>>>(iters, a_rows, d); }
  else { cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); bench<1><<<grid, 128, smem>>>(iters, a_rows, d); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[1024];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("issue variant %d grid=%d: %.1f cycles/MMA (tensor floor 64)\n", variant, grid, mx / ((double)iters * 72));
  return 0;
}
