// Which resource keeps a small kernel from starting beside the persistent conv tower?  Probe kernels of a chosen
// footprint (registers via a live accumulator array, dynamic shared memory) stamp %globaltimer / %smid per block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/bin/libcoresident_probe.so tools/coresident_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int R>
__global__ void __launch_bounds__(128) probe_kernel(unsigned long long* out, int iters) {
  extern __shared__ float sm[];
  float acc[R];
#pragma unroll
  for (int i = 0; i < R; ++i) acc[i] = (float)(threadIdx.x + i);
  const unsigned long long t0 = gtime();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = fmaf(acc[i], 1.0001f, acc[(i + 1) % R]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) s += acc[i];
  if (s == 12345.678f) sm[0] = s;      // keeps acc[] live
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    out[blockIdx.x * 4 + 0] = t0;
    out[blockIdx.x * 4 + 1] = gtime();
    out[blockIdx.x * 4 + 2] = smid;
  }
}

// dependent global loads (what a tree descent is): 64 hops through a random permutation, one warp per block
__global__ void __launch_bounds__(128) chase_kernel(unsigned long long* out, const int* __restrict__ next, int hops) {
  int i = (blockIdx.x * 977 + threadIdx.x) & 0xfffff;
  const unsigned long long t0 = gtime();
  for (int h = 0; h < hops; ++h) i = *((volatile const int*)next + i);
  if (threadIdx.x == 0 || i == -1) {
    out[blockIdx.x * 4 + 0] = t0;
    out[blockIdx.x * 4 + 1] = gtime();
    out[blockIdx.x * 4 + 2] = 0;
  }
}
extern "C" int chase_launch(int blocks, int hops, unsigned long long* out, const int* next, void* stream) {
  cudaFuncSetAttribute(chase_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  chase_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(out, next, hops);
  return (int)cudaGetLastError();
}

// dependent chains of one instruction class each (what slows a tree warp down beside the tower?)
//   0: DFMA   1: REDUX + VOTE + SHFL   2: LDS (dependent)   3: LDG.CONSTANT (__ldg, dependent, L1-resident table)
//   4: LDG.128 of 1.3 KB rows spread over 256 MB (dependent)   5: F2F/DMUL/FADD mix
__global__ void __launch_bounds__(128) opchain_kernel(unsigned long long* out, int mode, int iters, const int* __restrict__ tab,
                                                      const int4* __restrict__ big) {
  extern __shared__ int lds[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lds[i] = (i * 37 + 11) & 255;
  __syncthreads();
  double d = 1.0 + threadIdx.x;
  unsigned u = threadIdx.x * 2654435761u;
  int idx = threadIdx.x & 255;
  const unsigned long long t0 = gtime();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) d = __fma_rn(d, 1.0000001, 0.5);
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) {
      const unsigned m = __reduce_max_sync(0xffffffffu, u);
      const unsigned b = __ballot_sync(0xffffffffu, u == m);
      u = __shfl_sync(0xffffffffu, u + b, (lane + 1) & 31) * 1664525u + 1013904223u;
    }
  } else if (mode == 2) {
    for (int it = 0; it < iters; ++it) idx = lds[idx];
  } else if (mode == 3) {
    for (int it = 0; it < iters; ++it) idx = __ldg(tab + idx) & 255;
  } else if (mode == 4) {
    size_t row = (size_t)(blockIdx.x * 4 + (threadIdx.x >> 5)) * 4099u;
    for (int it = 0; it < iters; ++it) {
      const int4 v = big[(row % 200000u) * 82 + lane];
      row = row * 31u + (unsigned)__shfl_sync(0xffffffffu, v.x, 0) + 7u;
    }
    idx = (int)row;
  } else {
    float f = (float)threadIdx.x;
    for (int it = 0; it < iters; ++it) {
      d = __dmul_rn(d, 1.0000001);
      f = __fadd_rn(f, __double2float_rn(d));
      d = __dadd_rn((double)f, 1.0);
    }
    u = __float_as_uint(f);
  }
  const unsigned long long t1 = gtime();
  if (threadIdx.x == 0 || (d == 3.25 && u == 77u && idx == -5)) {
    out[blockIdx.x * 4 + 0] = t0;
    out[blockIdx.x * 4 + 1] = t1;
    out[blockIdx.x * 4 + 2] = (unsigned long long)(d + u + idx);
  }
}
extern "C" int opchain_launch(int mode, int blocks, int iters, unsigned long long* out, const int* tab, const void* big,
                              void* stream) {
  cudaFuncSetAttribute(opchain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  opchain_kernel<<<blocks, 128, 1024, (cudaStream_t)stream>>>(out, mode, iters, tab, (const int4*)big);
  return (int)cudaGetLastError();
}

__global__ void stamp_kernel(unsigned long long* out) { out[0] = gtime(); }

extern "C" int probe_launch(int regs_class, int blocks, int smem, int iters, unsigned long long* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (regs_class == 0) {
    cudaFuncSetAttribute(probe_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    probe_kernel<8><<<blocks, 128, smem, st>>>(out, iters);
  } else if (regs_class == 1) {
    cudaFuncSetAttribute(probe_kernel<48>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    probe_kernel<48><<<blocks, 128, smem, st>>>(out, iters);
  } else {
    cudaFuncSetAttribute(probe_kernel<80>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    probe_kernel<80><<<blocks, 128, smem, st>>>(out, iters);
  }
  return (int)cudaGetLastError();
}
extern "C" int stamp_launch(unsigned long long* out, void* stream) {
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out);
  return (int)cudaGetLastError();
}
