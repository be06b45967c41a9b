// forward kernels, uint16 raw (value = u / raw_denominator, dataset.py:87)
#include "isp_fwd_tu.cuh"
namespace r2l {
int launch_forward_u16(const FwdArgs& a, bool stats, cudaStream_t st, int* grid_used, bool* fused_tail) {
    return launch_forward_impl<uint16_t>(a, stats, st, grid_used, fused_tail);
}
}  // namespace r2l
