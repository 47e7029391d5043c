// Micro-benchmark of the training conv kernel's MMA loop: one producer lane, one MMA warp, a ring of weight stages with
// full / empty mbarriers, tcgen05.commit per stage.  No data is copied (the producer just arrives): what is measured is
// the hand-over protocol itself against the 64 cycles an M128 N128 K16 MMA occupies the tensor pipe.
// usage: ring_bench [grid]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../muzero_b200/csrc/umma.cuh"
using namespace mz::umma;

constexpr int kMaxStages = 8;

// four K steps under one election, accumulators d0, d1, d2, d3 (may coincide)
__device__ __forceinline__ void mma4_acc(uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, uint64_t a0, uint64_t a1, uint64_t a2,
                                         uint64_t a3, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q, t;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %12, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%1], %5, %9, %12, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%2], %6, %10, %12, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%3], %7, %11, %12, t;\n\t}"
      ::"r"(d0), "r"(d1), "r"(d2), "r"(d3), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc)
      : "memory");
}

// variant: 0 ring, commit per stage via commit_elect; 1 ring, commit from lane 0 only (branch); 2 no ring (MMAs back to back,
// one commit at the end); 3 ring, MMAs skipped (protocol only); 4 ring, the MMA warp does not wait for `full`
struct Args { int variant, mmas_per_stage, total_mmas, stages, a_rows, naccum, N; long long* out; };

__global__ void __launch_bounds__(192) ring_kernel(const Args p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (16 * p.a_rows + 8 * 128 * kMaxStages) * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int nst = p.total_mmas / p.mmas_per_stage;
  const uint32_t idesc = instr_desc_f16(128, (uint32_t)p.N);
  if (warp == 0) {
    if (lane == 0 && p.variant != 2 && p.variant != 7 && p.variant != 9) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        if (it >= p.stages) mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive(&full[s]);
      }
    }
  } else if (warp == 1 && p.variant >= 6) {
    // the lean loop: stage index / phase kept incrementally, descriptor high words hoisted, low words stepped by adds
    const bool uni = p.variant >= 8;           // variants 8 / 9: operands made provably warp-uniform with redux.sync
    auto U = [&](uint32_t x) { return uni ? __reduce_max_sync(0xffffffffu, x) : x; };
    const uint32_t tm = U(tmem_base);
    const uint32_t sA = smem_u32(smem), sW = sA + 16 * p.a_rows * 16;
    const uint64_t a_t = smem_desc(sA, p.a_rows * 16, 128), b_t = smem_desc(sW, 128 * 16, 128);
    const uint32_t a_hi = U((uint32_t)(a_t >> 32)), b_hi = U((uint32_t)(b_t >> 32));
    const uint32_t a_base = U((uint32_t)a_t + 11), b_base = U((uint32_t)b_t);
    const uint32_t a_step = U((uint32_t)(2 * p.a_rows)), b_step = 256u, stage_units = 8u * 128;
    auto d64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    const long long t0 = clock64();
    uint32_t st = 0, ph = 0, b_lo = b_base;
    const int per = p.mmas_per_stage;
    for (int it = 0; it < nst; ++it) {
      if (p.variant != 7 && p.variant != 9) { mbar_wait(&full[st], ph); tc_fence_after(); }
      uint32_t a_lo = a_base + (uint32_t)(it & 7);
      for (int k = 0; k < per; k += 4) {
        mma4_f16_elect(tm, d64(a_lo, a_hi), d64(a_lo + a_step, a_hi), d64(a_lo + 2 * a_step, a_hi), d64(a_lo + 3 * a_step, a_hi),
                       d64(b_lo, b_hi), d64(b_lo + b_step, b_hi), d64(b_lo + 2 * b_step, b_hi), d64(b_lo + 3 * b_step, b_hi), idesc, 1);
        a_lo += 4 * a_step;
      }
      if (p.variant != 7 && p.variant != 9) commit_elect(&empty[st]);
      b_lo += stage_units;
      if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1u; b_lo = b_base; }
    }
    const long long t1 = clock64();
    commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (lane == 0) { p.out[2 * blockIdx.x] = t1 - t0; p.out[2 * blockIdx.x + 1] = t2 - t0; }
  } else if (warp == 1) {
    const uint32_t tm = tmem_base;
    const uint32_t sA = smem_u32(smem), sW = sA + 16 * p.a_rows * 16;
    const long long t0 = clock64();
    for (int it = 0; it < nst; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
      if (p.variant != 2 && p.variant != 4) { mbar_wait(&full[s], ph); tc_fence_after(); }
      const uint32_t a0 = sA + (uint32_t)((11 + (it % 9)) * 16), b0 = sW + (uint32_t)s * (8 * 128 * 16);
      const uint32_t a_step = (uint32_t)(2 * p.a_rows * 16), b_step = 2u * 128 * 16;
      if (p.variant != 3) {
        for (int k = 0; k < p.mmas_per_stage; k += 4) {
          const uint32_t a = a0 + (uint32_t)(k % 8) * a_step, b = b0 + (uint32_t)(k % 4) * b_step;
          if (p.naccum == 1) {
            mma4_f16_elect(tm, smem_desc(a, p.a_rows * 16, 128), smem_desc(a + a_step, p.a_rows * 16, 128),
                           smem_desc(a + 2 * a_step, p.a_rows * 16, 128), smem_desc(a + 3 * a_step, p.a_rows * 16, 128),
                           smem_desc(b, 128 * 16, 128), smem_desc(b + b_step, 128 * 16, 128), smem_desc(b + 2 * b_step, 128 * 16, 128),
                           smem_desc(b + 3 * b_step, 128 * 16, 128), idesc, 1);
          } else if (p.naccum >= 12) {
            const uint32_t d1 = tm + 128, d2 = p.naccum == 14 ? tm + 256 : tm, d3 = p.naccum == 14 ? tm + 384 : tm + 128;
            mma4_acc(tm, d1, d2, d3, smem_desc(a, p.a_rows * 16, 128), smem_desc(a + a_step, p.a_rows * 16, 128),
                     smem_desc(a + 2 * a_step, p.a_rows * 16, 128), smem_desc(a + 3 * a_step, p.a_rows * 16, 128),
                     smem_desc(b, 128 * 16, 128), smem_desc(b + b_step, 128 * 16, 128), smem_desc(b + 2 * b_step, 128 * 16, 128),
                     smem_desc(b + 3 * b_step, 128 * 16, 128), idesc);
          } else {
            for (int j = 0; j < 4; ++j)
              mma_f16_elect(tm + (uint32_t)((p.naccum == 5 ? 0 : (j % p.naccum)) * 128), smem_desc(a + j * a_step, p.a_rows * 16, 128),
                            smem_desc(b + j * b_step, 128 * 16, 128), idesc, 1);
          }
        }
      }
      if (p.variant == 1) { if (lane == 0) commit(&empty[s]); __syncwarp(); }
      else if (p.variant != 2) commit_elect(&empty[s]);
    }
    const long long t1 = clock64();
    commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (lane == 0) { p.out[2 * blockIdx.x] = t1 - t0; p.out[2 * blockIdx.x + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 100;
  const int a_rows = 151;
  long long* d;
  cudaMalloc(&d, 2 * grid * sizeof(long long));
  const int smem = 16 * a_rows * 16 + kMaxStages * 8 * 128 * 16 + 1024;
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct V { const char* name; int variant, per_stage, stages, naccum, N; };
  const V vs[] = {
    {"no ring: 72 MMAs back to back, one commit", 2, 4, 8, 1, 128},
    {"ring of 8, 18 stages x 4 MMAs, commit_elect per stage", 0, 4, 8, 1, 128},
    {"ring of 8, 18 stages x 4 MMAs, commit from lane 0", 1, 4, 8, 1, 128},
    {"ring of 4, 9 stages x 8 MMAs, commit_elect per stage", 0, 8, 4, 1, 128},
    {"ring of 8, 18 stages, protocol only (no MMAs)", 3, 4, 8, 1, 128},
    {"ring of 8, 18 stages x 4 MMAs, MMA warp does not wait for full", 4, 4, 8, 1, 128},
    {"ring of 8, 72 stages x 1.. (4 MMAs each, 288 MMAs)", 0, 4, 8, 1, 128},
    {"no ring, single MMAs (not mma4), 1 accumulator", 2, 4, 8, 5, 128},
    {"no ring, 2 accumulators alternating", 2, 4, 8, 2, 128},
    {"no ring, 4 accumulators alternating", 2, 4, 8, 4, 128},
    {"no ring, 1 accumulator, N = 64", 2, 4, 8, 1, 64},
    {"no ring, 2 accumulators, N = 64", 2, 4, 8, 2, 64},
    {"ring of 8, 18 stages x 4, 2 accumulators", 0, 4, 8, 2, 128},
    {"ring of 8, 18 stages x 4, 4 accumulators", 0, 4, 8, 4, 128},
    {"ring of 4, 9 stages x 8, 4 accumulators", 0, 8, 4, 4, 128},
    {"no ring, mma4 over 2 accumulators", 2, 4, 8, 12, 128},
    {"no ring, mma4 over 4 accumulators", 2, 4, 8, 14, 128},
    {"ring of 8 x 4, mma4 over 2 accumulators", 0, 4, 8, 12, 128},
    {"ring of 8 x 4, mma4 over 4 accumulators", 0, 4, 8, 14, 128},
    {"ring of 4 x 8, mma4 over 4 accumulators", 0, 8, 4, 14, 128},
    {"lean loop, no ring", 7, 4, 8, 1, 128},
    {"lean loop, ring of 8 x 4 MMAs", 6, 4, 8, 1, 128},
    {"lean loop, ring of 4 x 8 MMAs", 6, 8, 4, 1, 128},
    {"uniform lean loop, no ring, 7200 MMAs", 9, 4, 8, 1, 128},
    {"uniform lean loop, ring of 8 x 4 MMAs, 7200 MMAs", 8, 4, 8, 1, 128},
    {"uniform lean loop, ring of 4 x 8 MMAs, 7200 MMAs", 8, 8, 4, 1, 128},
    {"lean loop, no ring, 7200 MMAs", 7, 4, 8, 1, 128},
    {"lean loop, ring of 8 x 4 MMAs, 7200 MMAs", 6, 4, 8, 1, 128},
    {"lean loop, ring of 4 x 8 MMAs, 7200 MMAs", 6, 8, 4, 1, 128},
    {"lean loop, ring of 2 x 16 MMAs, 7200 MMAs", 6, 16, 2, 1, 128},
    {"no ring, 1 accumulator, 720 MMAs", 2, 4, 8, 1, 128},
    {"no ring, 1 accumulator, 7200 MMAs", 2, 4, 8, 1, 128},
    {"no ring, 1 accumulator, 72000 MMAs", 2, 4, 8, 1, 128},
    {"ring of 8 x 4 MMAs, 72000 MMAs", 0, 4, 8, 1, 128},
  };
  int idx = 0;
  for (const V& v : vs) {
    const int total = idx == 6 ? 288 : (idx >= 23 && idx <= 29 ? 7200 : (idx == 30 ? 720 : (idx == 31 ? 7200 : (idx >= 32 ? 72000 : 72))));
    Args a{v.variant, v.per_stage, total, v.stages, a_rows, v.naccum == 5 ? 1 : v.naccum, v.N, d};
    if (v.naccum == 5) a.naccum = 5;
    for (int rep = 0; rep < 3; ++rep) ring_kernel<<<grid, 192, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    static long long h[4096];
    cudaMemcpy(h, d, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mi = 0, md = 0;
    for (int i = 0; i < grid; ++i) { mi += h[2 * i]; md += h[2 * i + 1]; }
    printf("%-70s issue loop %7.0f cycles, until done %7.0f cycles = %.1f per MMA (floor 64)\n", v.name, mi / grid, md / grid, md / grid / total);
    ++idx;
  }
  return 0;
}
