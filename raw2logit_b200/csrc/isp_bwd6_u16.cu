// sixth-generation backward kernels, uint16 raw
#include "isp_bwd6_tu.cuh"
namespace r2l {
int launch_backward6_u16(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    return launch_backward6_impl<uint16_t>(a, st, grid_used);
}
}  // namespace r2l
