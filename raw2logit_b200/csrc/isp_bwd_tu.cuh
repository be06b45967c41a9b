// isp_bwd_tu.cuh -- third-generation backward kernels + launcher for one raw element type
#pragma once
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT>
__global__ void __launch_bounds__(Cfg::NT, 1) isp_backward_kernel(BwdArgs a, TileGrid grid,
                                                                  const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smem[];
    bwd3_cta<Cfg, RawT, true>(blockIdx.x, gridDim.x, a, grid, smem, &tmap);
}

template <class Cfg, typename RawT>
static int launch_backward3_t(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);     // tiles of image pairs
    CUtensorMap tmap;
    if (!make_raw_tensor_map(&tmap, a.raw, (int)sizeof(RawT), a.B, a.H, a.W, sizeof(RawT) == 2 ? Cfg::PW + 8 : Cfg::PW, Cfg::RH))
        return kNotServed;
    int g = 0;
    int rc = persistent_grid(isp_backward_kernel<Cfg, RawT>, Cfg::NT, Cfg::kSmemBytesTma, grid.n, &g);
    if (rc != R2L_OK) return rc;
    isp_backward_kernel<Cfg, RawT><<<g, Cfg::NT, Cfg::kSmemBytesTma, st>>>(a, grid, tmap);
    if (grid_used) *grid_used = g;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

// kNotServed when the shape or an alignment rule sends the call to the generic kernel
template <typename RawT>
static int launch_backward3_impl(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    using Probe = Bwd3<false, false>;
    if (!bwd3_shape_ok(a.H, a.W, Probe::TH, Probe::TW)) return kNotServed;
    if (!aligned(a.gout, 16) || !aligned(a.graw, 16) || !aligned(a.additive, 16)) return kNotServed;   // 128-bit rows
    const bool tail = a.gtail != nullptr;
    if (a.out && aligned(a.out, 16)) {                  // forward output at hand: no Gaussian / colour-tail recompute
        if (a.graw) return tail ? launch_backward3_t<Bwd3<true, true, true>, RawT>(a, st, grid_used)
                                : launch_backward3_t<Bwd3<true, false, true>, RawT>(a, st, grid_used);
        return tail ? launch_backward3_t<Bwd3<false, true, true>, RawT>(a, st, grid_used)
                    : launch_backward3_t<Bwd3<false, false, true>, RawT>(a, st, grid_used);
    }
    if (a.graw) return tail ? launch_backward3_t<Bwd3<true, true>, RawT>(a, st, grid_used)
                            : launch_backward3_t<Bwd3<true, false>, RawT>(a, st, grid_used);
    return tail ? launch_backward3_t<Bwd3<false, true>, RawT>(a, st, grid_used)
                : launch_backward3_t<Bwd3<false, false>, RawT>(a, st, grid_used);
}

}  // namespace r2l
