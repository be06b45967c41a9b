// forward kernels, float32 raw
#include "isp_fwd_tu.cuh"
namespace r2l {
int launch_forward_f32(const FwdArgs& a, bool stats, cudaStream_t st, int* grid_used, bool* fused_tail) {
    return launch_forward_impl<float>(a, stats, st, grid_used, fused_tail);
}
}  // namespace r2l
