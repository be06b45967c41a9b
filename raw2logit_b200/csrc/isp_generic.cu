// Generic scalar kernels (first generation, isp_core.cuh): one site per thread-iteration, every border rule
// evaluated per site.  They serve the shapes and alignments the vectorised kernels do not (W % 4 != 0, tiny
// images, a last tile row/column of <= 4 sites, pointers that are not 16-byte aligned).
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT>
__global__ void __launch_bounds__(Cfg::NT, 1) isp_backward_generic_kernel(BwdArgs a, TileGrid grid) {
    extern __shared__ __align__(128) float smem[];
    bwd_cta<Cfg, RawT>(blockIdx.x, gridDim.x, a, grid, smem);
}

template <class Cfg, typename RawT>
static int launch_backward_generic_t(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid(a.B, a.H, a.W, Cfg::TH, Cfg::TW);
    int g = 0;
    int rc = persistent_grid(isp_backward_generic_kernel<Cfg, RawT>, Cfg::NT, Cfg::kSmemBytes, grid.n, &g);
    if (rc != R2L_OK) return rc;
    isp_backward_generic_kernel<Cfg, RawT><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, grid);
    if (grid_used) *grid_used = g;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

int launch_backward_generic(const BwdArgs& a, int raw_dtype, cudaStream_t st, int* grid_used) {
    if (a.graw) {
        return raw_dtype == R2L_F32 ? launch_backward_generic_t<BwdWithRaw, float>(a, st, grid_used)
                                    : launch_backward_generic_t<BwdWithRaw, uint16_t>(a, st, grid_used);
    }
    return raw_dtype == R2L_F32 ? launch_backward_generic_t<BwdNoRaw, float>(a, st, grid_used)
                                : launch_backward_generic_t<BwdNoRaw, uint16_t>(a, st, grid_used);
}

}  // namespace r2l
