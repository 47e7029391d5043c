// Stand-alone check of the tcgen05 building blocks conv.cu relies on (run on a B200):
//   C[128 x 128] = A[row + shift][0..K) . B[n][0..K)^T, bf16 in, fp32 accumulate in TMEM,
// operands in the no-swizzle K-major core-matrix layout of umma.cuh, A written by
// threads (generic proxy + fence.proxy.async), B brought in by one 1-D bulk copy.
// usage: umma_probe [mode] [shift]    mode bit0: swap LBO/SBO meaning, bit1: B by threads too
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../muzero_b200/csrc/umma.cuh"

using namespace mz::umma;

constexpr int K = 64, KG = K / 8, AROWS = 160, M = 128, N = 128;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* Bpacked, float* C, int mode,
                                             int shift) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                               // [KG][AROWS][16 B]
  unsigned char* sB = smem + KG * AROWS * 16;             // [KG][N][16 B]
  __shared__ __align__(8) uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 128);
  for (int i = tid; i < KG * AROWS; i += 128) {
    const int g = i / AROWS, r = i % AROWS;
    *reinterpret_cast<int4*>(sA + (size_t)(g * AROWS + r) * 16) = *reinterpret_cast<const int4*>(A + (size_t)r * K + g * 8);
  }
  if (mode & 2)
    for (int i = tid; i < KG * N; i += 128)
      *reinterpret_cast<int4*>(sB + (size_t)i * 16) = reinterpret_cast<const int4*>(Bpacked)[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    if (!(mode & 2)) {
      mbar_arrive_expect_tx(&bar_b, KG * N * 16);
      bulk_g2s(sB, Bpacked, KG * N * 16, &bar_b);
      mbar_wait(&bar_b, 0);
    }
    const uint32_t idesc = instr_desc_bf16(M, N);
    for (int k = 0; k < K / 16; ++k) {
      const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 16 + (uint32_t)(2 * k) * AROWS * 16;
      const uint32_t b0 = smem_u32(sB) + (uint32_t)(2 * k) * N * 16;
      uint64_t da, db;
      if (mode & 1) { da = smem_desc(a0, 128, AROWS * 16); db = smem_desc(b0, 128, N * 16); }
      else          { da = smem_desc(a0, AROWS * 16, 128); db = smem_desc(b0, N * 16, 128); }
      mma_f16(tm, da, db, idesc, k > 0);
    }
    commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, shift = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<__nv_bfloat16> A(AROWS * K), B(N * K), Bp(N * K);
  std::vector<float> Af(AROWS * K), Bf(N * K);
  srand(1);
  for (int i = 0; i < AROWS * K; ++i) { A[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Af[i] = __bfloat162float(A[i]); }
  for (int i = 0; i < N * K; ++i) { B[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Bf[i] = __bfloat162float(B[i]); }
  for (int g = 0; g < KG; ++g) for (int n = 0; n < N; ++n) for (int e = 0; e < 8; ++e) Bp[(g * N + n) * 8 + e] = B[n * K + g * 8 + e];
  __nv_bfloat16 *dA, *dB; float* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, Bp.size() * 2); cudaMalloc(&dC, M * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dC, 0, M * N * 4);
  const int smem = KG * AROWS * 16 + KG * N * 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dC, mode, shift);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("umma_probe mode=%d shift=%d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 2; }
  std::vector<float> C(M * N);
  cudaMemcpy(C.data(), dC, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)Af[(m + shift) * K + k] * Bf[n * K + k];
    const double err = fabs(ref - C[m * N + n]);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3) ++bad;
  }
  printf("umma_probe mode=%d shift=%d: max abs err %.3g, %d/%d wrong -> %s\n", mode, shift, maxerr, bad, M * N,
         bad ? "FAIL" : "PASS");
  return bad ? 1 : 0;
}
