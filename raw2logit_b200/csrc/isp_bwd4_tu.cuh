// isp_bwd4_tu.cuh -- fourth-generation backward kernels + launcher for one raw element type
#pragma once
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT, int CPS>
__global__ void __launch_bounds__(Cfg::NT, CPS) isp_backward4_kernel(BwdArgs a, TileGrid grid) {
    extern __shared__ __align__(128) float smem[];
    bwd4_cta<Cfg, RawT>(blockIdx.x, gridDim.x, a, grid, smem);
}

template <class Cfg, typename RawT, int CPS>
static int launch_backward4_t(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);     // tiles of image pairs
    int g = 0;
    int rc = persistent_grid(isp_backward4_kernel<Cfg, RawT, CPS>, Cfg::NT, Cfg::kSmemBytes, grid.n, &g);
    if (rc != R2L_OK) return rc;
    if (a.ticket) {
        cudaError_t e0 = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned), st);
        if (e0 != cudaSuccess) return cuda_fail(e0);
    }
    isp_backward4_kernel<Cfg, RawT, CPS><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, grid);
    if (grid_used) *grid_used = g;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

// kNotServed when the shape or an alignment rule sends the call to an older generation
template <typename RawT>
static int launch_backward4_impl(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    if (!a.out || !a.luma || !bwd4_shape_ok(a.H, a.W)) return kNotServed;
    if (!aligned(a.gout, 16) || !aligned(a.graw, 16) || !aligned(a.additive, 16) || !aligned(a.out, 16) ||
        !aligned(a.luma, 16) || !aligned(a.raw, 4 * sizeof(RawT)))
        return kNotServed;                                                          // 128-bit rows
    const bool tail = a.gtail != nullptr;
    constexpr int CPS = kBwd4CtasPerSm;
    if (a.graw) return tail ? launch_backward4_t<Bwd4<true, true>, RawT, CPS>(a, st, grid_used)
                            : launch_backward4_t<Bwd4<true, false>, RawT, CPS>(a, st, grid_used);
    return tail ? launch_backward4_t<Bwd4<false, true>, RawT, CPS>(a, st, grid_used)
                : launch_backward4_t<Bwd4<false, false>, RawT, CPS>(a, st, grid_used);
}

}  // namespace r2l
