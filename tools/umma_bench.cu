// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, fp16) issued back to back from shared
// memory operands in the no-swizzle K-major layout, as a function of N and of the K-direction
// strides (LBO) of A and B.  Answers: is the operand fetch bank-conflict bound?
// usage: umma_bench N a_rows b_rows [grid] [iters] [extra_smem_writer]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../muzero_b200/csrc/umma.cuh"

using namespace mz::umma;

__global__ void __launch_bounds__(128) bench(int N, int a_rows, int b_rows, int iters, long long* out, int writer) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16 * a_rows * 16;
  for (int i = tid; i < (16 * a_rows + 8 * b_rows) * 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = instr_desc_f16(128, N);
    const uint64_t at = smem_desc(smem_u32(sA) + 11 * 16, a_rows * 16, 128);
    const uint64_t bt = smem_desc(smem_u32(sB), b_rows * 16, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        mma_bf16(tm, at + (uint64_t)(2 * ks * a_rows), bt + (uint64_t)(2 * ks * b_rows), idesc, 1);
        if (N <= 128) mma_bf16(tm + 256, at + 128 + (uint64_t)(2 * ks * a_rows), bt + (uint64_t)(2 * ks * b_rows), idesc, 1);
      }
    }
    commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (writer && warp >= 2) {
    // competing shared-memory store traffic (what cp.async / TMA fills do in the real kernel)
    const uint32_t w = smem_u32(sB + 8 * b_rows * 16);
    for (int it = 0; it < iters * writer; ++it)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(w + (uint32_t)((tid - 64) + 64 * (it & 31)) * 16), "r"(it)
                   : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 128, a_rows = argc > 2 ? atoi(argv[2]) : 279, b_rows = argc > 3 ? atoi(argv[3]) : N;
  const int grid = argc > 4 ? atoi(argv[4]) : 1, iters = argc > 5 ? atoi(argv[5]) : 2000, writer = argc > 6 ? atoi(argv[6]) : 0;
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 16 * a_rows * 16 + 8 * b_rows * 16 + 64 * 32 * 16 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<<<grid, 128, smem>>>(N, a_rows, b_rows, iters, d, writer);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long h[256];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  const int per_it = (N <= 128) ? 8 : 4;
  double mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc = mx / ((double)iters * per_it);
  printf("N=%d a_rows=%d (LBO %d B, mod128=%d) b_rows=%d (LBO %d B, mod128=%d) grid=%d writer=%d: %.1f cycles/MMA, "
         "%.0f flop/cycle/SM (ideal 8192)\n", N, a_rows, a_rows * 16, (a_rows * 16) % 128, b_rows, b_rows * 16,
         (b_rows * 16) % 128, grid, writer, cyc, 2.0 * 128 * N * 16 / cyc);
  return 0;
}
