// Error reporting, launch accounting and device checks of the C ABI.
#include "common.cuh"

#include <atomic>
#include <cstdlib>

namespace mz {

static thread_local char g_error[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("MZ_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

}  // namespace mz

extern "C" const char* mz_last_error(void) { return mz::g_error; }

extern "C" int mz_version(void) { return 100; }

extern "C" int mz_set_pdl(int enable) { mz::set_pdl(enable); return MZ_OK; }

extern "C" uint64_t mz_launch_count(void) { return mz::g_launches.load(std::memory_order_relaxed); }

extern "C" int mz_device_check(int device, int* sm_count, int* cc) {
  cudaDeviceProp prop;
  MZ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc) *cc = prop.major * 10 + prop.minor;
  if (prop.major != 10) {
    mz::set_error("muzero_b200 is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major,
                  prop.minor, prop.name);
    return MZ_EINVAL;
  }
  return MZ_OK;
}
