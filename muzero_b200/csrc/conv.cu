#include "net.cuh"
namespace mz {
int conv_hidden_bytes(const mz_net_config&, int32_t*) { set_error("conv nets not built yet"); return MZ_EINVAL; }
int conv_arena_bytes(const mz_net_config&, int, size_t*) { set_error("conv nets not built yet"); return MZ_EINVAL; }
int conv_create(const mz_net_config&, const float* const*, int, int, void*, size_t, NetImpl**) { set_error("conv nets not built yet"); return MZ_EINVAL; }
}
