#include "vfa_common.cuh"
namespace vfa {
size_t umma_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*, uint32_t) { return 0; }
bool umma_supported(const vfa_geometry_t*, const vfa_shape_t*, uint32_t) { return false; }
int launch_fwd_umma(AggParams, const float* const*, void*, uint32_t, cudaStream_t) { set_error("umma not built"); return VFA_ERR_UNSUPPORTED; }
size_t bwd_workspace_bytes(const vfa_geometry_t*, const vfa_shape_t*) { return 0; }
int launch_bwd(AggParams, const float* const*, const float*, float* const*, float* const*, float* const*, void*, cudaStream_t) { set_error("bwd not built"); return VFA_ERR_UNSUPPORTED; }
}
