#include "vfa_common.cuh"
namespace vfa {
size_t bwd_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*) { return 0; }
int launch_bwd(AggParams, const float* const*, const float*, float* const*, float* const*, float* const*, void*, cudaStream_t) { set_error("bwd not built"); return VFA_ERR_UNSUPPORTED; }
}
