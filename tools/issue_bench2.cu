// Which code shape lets ONE warp issue M128 N128 K16 MMAs at the tensor pipe's 64 cycles per MMA while it also runs the
// weight-ring protocol (wait full -> MMAs -> commit empty)?  Compile-time variants, steady state (7200 MMAs).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../muzero_b200/csrc/umma.cuh"
using namespace mz::umma;

constexpr int kStages = 4;
__device__ __forceinline__ uint64_t d64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// kUniform: bases through redux.sync (provably warp-uniform).  kPer: MMAs per ring stage.  kSingle: one MMA per asm statement.
template <bool kUniform, int kPer, bool kSingle, bool kRing>
__global__ void __launch_bounds__(192) k(int total, int a_rows, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16 * a_rows + 16 * 128 * kStages) * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int nst = total / kPer;
  if (warp == 0) {
    if (lane == 0 && kRing) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < nst; ++it) {
        if (it >= kStages) mbar_wait(&empty[s], ph);
        mbar_arrive(&full[s]);
        if (++s == kStages) { s = 0; if (it >= kStages) ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t sA = smem_u32(smem), sW = sA + 16 * a_rows * 16;
    const uint64_t a_t = smem_desc(sA, a_rows * 16, 128), b_t = smem_desc(sW, 128 * 16, 128);
    uint32_t tm = tmem_base, a_hi = (uint32_t)(a_t >> 32), b_hi = (uint32_t)(b_t >> 32), a_base = (uint32_t)a_t + 11, b_base = (uint32_t)b_t;
    uint32_t a_step = (uint32_t)(2 * a_rows);
    uint32_t idesc = instr_desc_f16(128, 128);
    if (kUniform) {
      tm = __reduce_max_sync(0xffffffffu, tm); a_hi = __reduce_max_sync(0xffffffffu, a_hi); b_hi = __reduce_max_sync(0xffffffffu, b_hi);
      a_base = __reduce_max_sync(0xffffffffu, a_base); b_base = __reduce_max_sync(0xffffffffu, b_base);
      a_step = __reduce_max_sync(0xffffffffu, a_step);
    }
    constexpr uint32_t b_step = 256u, stage_units = 16u * 128;
    const long long t0 = clock64();
    uint32_t st = 0, ph = 0, b_lo = b_base;
    for (int it = 0; it < nst; ++it) {
      if (kRing) { mbar_wait(&full[st], ph); tc_fence_after(); }
      uint32_t a_lo = a_base + (uint32_t)(it & 7);
#pragma unroll
      for (int k = 0; k < kPer; k += 4) {
        if (kSingle) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_f16_elect(tm, d64(a_lo + (k + j) * a_step, a_hi), d64(b_lo + (k + j) * b_step, b_hi), idesc, 1);
        } else {
          mma4_f16_elect(tm, d64(a_lo + k * a_step, a_hi), d64(a_lo + (k + 1) * a_step, a_hi), d64(a_lo + (k + 2) * a_step, a_hi),
                         d64(a_lo + (k + 3) * a_step, a_hi), d64(b_lo + k * b_step, b_hi), d64(b_lo + (k + 1) * b_step, b_hi),
                         d64(b_lo + (k + 2) * b_step, b_hi), d64(b_lo + (k + 3) * b_step, b_hi), idesc, 1);
        }
      }
      if (kRing) commit_elect(&empty[st]);
      b_lo += stage_units;
      if (++st == kStages) { st = 0; ph ^= 1u; b_lo = b_base; }
    }
    const long long t1 = clock64();
    commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

template <bool U, int P, bool S, bool R>
void run(const char* name, long long* d, int grid) {
  const int a_rows = 151, total = 7200;
  const int smem = 16 * a_rows * 16 + kStages * 16 * 128 * 16 + 1024;
  cudaFuncSetAttribute(k<U, P, S, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) k<U, P, S, R><<<grid, 192, smem>>>(total, a_rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  static long long h[4096];
  cudaMemcpy(h, d, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double md = 0;
  for (int i = 0; i < grid; ++i) md += h[2 * i + 1];
  printf("%-64s %.1f cycles per MMA (floor 64)\n", name, md / grid / total);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 100;
  long long* d;
  cudaMalloc(&d, 2 * grid * sizeof(long long));
  run<false, 4, false, false>("plain, mma4, no ring", d, grid);
  run<true, 4, false, false>("uniform, mma4, no ring", d, grid);
  run<false, 4, true, false>("plain, single-MMA asm, no ring", d, grid);
  run<true, 4, true, false>("uniform, single-MMA asm, no ring", d, grid);
  run<false, 4, false, true>("plain, mma4, ring x 4 MMAs", d, grid);
  run<true, 4, false, true>("uniform, mma4, ring x 4 MMAs", d, grid);
  run<true, 4, true, true>("uniform, single-MMA asm, ring x 4 MMAs", d, grid);
  run<false, 8, false, true>("plain, mma4, ring x 8 MMAs", d, grid);
  run<true, 8, false, true>("uniform, mma4, ring x 8 MMAs", d, grid);
  run<true, 8, true, true>("uniform, single-MMA asm, ring x 8 MMAs", d, grid);
  return 0;
}
