// Per-kernel cost of a DEPENDENT chain of small kernels inside one CUDA graph on this GPU: the floor under the
// training towers' conv -> BatchNorm -> conv chains (each link waits for the one before it).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/chain_probe tools/chain_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// mode bit 0: allocate / free 128 TMEM columns; bit 1: touch data (read src, write dst: `rows` int4 per thread)
struct P { const int4* src; int4* dst; int n; int mode; int pdl; };

__global__ void __launch_bounds__(256) link_kernel(const __grid_constant__ P p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t holder;
  if (p.pdl) pdl_trigger();
  if (p.mode & 1) {
    if (threadIdx.x < 32) {
      uint32_t a = (uint32_t)__cvta_generic_to_shared(&holder);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(a) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
  }
  if (p.pdl) pdl_wait();
  if (p.mode & 2) {
    const int stride = gridDim.x * blockDim.x;
    int4 v[4];
    int i0 = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = i0 + k * stride; v[k] = i < p.n ? __ldcg(p.src + i) : make_int4(0, 0, 0, 0); }
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int i = i0 + k * stride; if (i < p.n) { v[k].x += 1; p.dst[i] = v[k]; } }
  }
  if (p.mode & 1) {
    __syncthreads();
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(holder) : "memory");
    }
  }
  if (smem[0] == 77 && p.n == -1) p.dst[0] = make_int4(1, 1, 1, 1);
}

struct Variant { const char* name; int ctas_a, thr_a, smem_a, mode_a; int ctas_b, thr_b, smem_b, mode_b; int pdl; int carve; };

int main() {
  const int n = 12800 * 16;           // int4 elements of one 16-plane activation tensor (3.3 MB)
  int4 *a, *b;
  CK(cudaMalloc(&a, (size_t)n * 16)); CK(cudaMalloc(&b, (size_t)n * 16));
  CK(cudaMemset(a, 0, (size_t)n * 16)); CK(cudaMemset(b, 0, (size_t)n * 16));
  CK(cudaFuncSetAttribute(link_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  const Variant vs[] = {
    {"empty 100x192, no smem", 100, 192, 0, 0, 100, 192, 0, 0, 0, 0},
    {"empty 100x192, 200 KB smem", 100, 192, 200 * 1024, 0, 100, 192, 200 * 1024, 0, 0, 0},
    {"empty 100x192, 200 KB smem, TMEM alloc", 100, 192, 200 * 1024, 1, 100, 192, 200 * 1024, 1, 0, 0},
    {"alternating 200 KB+TMEM / 208x256 no smem", 100, 192, 200 * 1024, 1, 208, 256, 0, 0, 0, 0},
    {"alternating, max-shared carve-out on both", 100, 192, 200 * 1024, 1, 208, 256, 0, 0, 0, 1},
    {"alternating, carve-out, PDL", 100, 192, 200 * 1024, 1, 208, 256, 0, 0, 1, 1},
    {"alternating, carve-out, PDL, second kernel streams 3.3 MB in / out", 100, 192, 200 * 1024, 1, 208, 256, 0, 2, 1, 1},
    {"alternating, carve-out, no PDL, second kernel streams 3.3 MB in / out", 100, 192, 200 * 1024, 1, 208, 256, 0, 2, 0, 1},
    {"stream kernel alone 208x256 (3.3 MB in / out), no PDL", 208, 256, 0, 2, 208, 256, 0, 2, 0, 0},
    {"stream kernel alone 800x256 (1 int4 per thread x4 strided), no PDL", 800, 256, 0, 2, 800, 256, 0, 2, 0, 0},
    {"empty 1x32", 1, 32, 0, 0, 1, 32, 0, 0, 0, 0},
    {"empty 1x32 PDL", 1, 32, 0, 0, 1, 32, 0, 0, 1, 0},
  };
  for (const Variant& v : vs) {
    CK(cudaFuncSetAttribute(link_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            v.carve ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault));
    const int links = 200;
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < links; ++i) {
      const bool A = (i & 1) == 0;
      P p{(i & 1) ? a : b, (i & 1) ? b : a, n, A ? v.mode_a : v.mode_b, v.pdl};
      cudaLaunchConfig_t lc = {};
      lc.gridDim = dim3(A ? v.ctas_a : v.ctas_b); lc.blockDim = dim3(A ? v.thr_a : v.thr_b);
      lc.dynamicSmemBytes = A ? v.smem_a : v.smem_b; lc.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = v.pdl;
      lc.attrs = at; lc.numAttrs = 1;
      CK(cudaLaunchKernelEx(&lc, link_kernel, p));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int w = 0; w < 3; ++w) CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 10;
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-75s %6.2f us per kernel\n", v.name, ms * 1e3 / (reps * links));
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  }
  return 0;
}
