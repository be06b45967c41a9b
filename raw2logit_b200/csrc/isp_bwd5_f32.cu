// fifth-generation backward kernels, float32 raw
#include "isp_bwd5_tu.cuh"
namespace r2l {
int launch_backward5_f32(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    return launch_backward5_impl<float>(a, st, grid_used);
}
}  // namespace r2l
