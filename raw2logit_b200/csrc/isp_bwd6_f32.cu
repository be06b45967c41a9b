// sixth-generation backward kernels, float32 raw
#include "isp_bwd6_tu.cuh"
namespace r2l {
int launch_backward6_f32(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    return launch_backward6_impl<float>(a, st, grid_used);
}
}  // namespace r2l
