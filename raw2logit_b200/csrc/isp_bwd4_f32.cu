// fourth-generation backward kernels, float32 raw
#include "isp_bwd4_tu.cuh"
namespace r2l {
int launch_backward4_f32(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    return launch_backward4_impl<float>(a, st, grid_used);
}
}  // namespace r2l
