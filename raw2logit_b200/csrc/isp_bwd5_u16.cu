// fifth-generation backward kernels, uint16 raw
#include "isp_bwd5_tu.cuh"
namespace r2l {
int launch_backward5_u16(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    return launch_backward5_impl<uint16_t>(a, st, grid_used);
}
}  // namespace r2l
