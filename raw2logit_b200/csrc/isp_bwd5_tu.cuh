// isp_bwd5_tu.cuh -- fifth-generation backward kernels + launcher for one raw element type
#pragma once
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT, int CPS>
__global__ void __launch_bounds__(Cfg::NT, CPS) isp_backward5_kernel(BwdArgs a, TileGrid grid, const __grid_constant__ Bwd5Maps maps) {
    extern __shared__ __align__(128) float smem[];
    bwd5_cta<Cfg, RawT>(blockIdx.x, gridDim.x, a, grid, smem, &maps);
}

template <class Cfg, typename RawT, int CPS>
static int launch_backward5_t(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);     // tiles of image pairs
    int g = 0;
    int rc = tmem_ctas_per_device(reinterpret_cast<const void*>(isp_backward5_kernel<Cfg, RawT, CPS>), Cfg::NT, Cfg::kSmemBytes,
                                  Cfg::kTmemCols, CPS, &g);
    if (rc != R2L_OK) return rc;
    if (g > grid.n) g = grid.n;
    if (g > kFinishMaxCtas) g = kFinishMaxCtas;                 // the last rows of the workspace hold the finish's group rows
    if (rc != R2L_OK) return rc;
    Bwd5Maps maps;
    // default 0x33: both groups of boxes at the start of B7 (in-call A/B, profiles/r02_experiments.md: 89.3 us without,
    // 87.3 us with; at the start of the tile -- 23 us ahead -- the boxes are evicted again before use: 93.8 us)
    static const int mode = [] { const char* v = getenv("R2L_ISP_BWD_PREFETCH"); return v ? (int)strtol(v, nullptr, 16) : 0x33; }();
    maps.on = (mode && make_bwd5_prefetch_maps(&maps, a, sizeof(RawT), Cfg::TH, Cfg::TW)) ? mode : 0;
    BwdArgs a2 = a;
    a2.ticket_gen = next_ticket_generation();                   // no memset in front of the kernel (take_ticket, isp_bwd5.cuh)
    cudaError_t e = launch_pdl(pdl_enabled_backward(), isp_backward5_kernel<Cfg, RawT, CPS>, g, Cfg::NT, Cfg::kSmemBytes, st, a2, grid, maps);
    if (grid_used) *grid_used = g;
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

// kNotServed when the shape or an alignment rule sends the call to an older generation
template <typename RawT>
static int launch_backward5_impl(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    if (!a.out || !a.luma || !bwd5_shape_ok(a.H, a.W)) return kNotServed;
    if (!aligned(a.gout, 16) || !aligned(a.graw, 16) || !aligned(a.additive, 16) || !aligned(a.out, 16) ||
        !aligned(a.luma, 32) || !aligned(a.raw, 4 * sizeof(RawT)))
        return kNotServed;                                                          // 128-bit rows
    const bool tail = a.gtail != nullptr;
    constexpr int CPS = kBwd5CtasPerSm;
    if (a.graw) return tail ? launch_backward5_t<Bwd5<true, true>, RawT, CPS>(a, st, grid_used)
                            : launch_backward5_t<Bwd5<true, false>, RawT, CPS>(a, st, grid_used);
    return tail ? launch_backward5_t<Bwd5<false, true>, RawT, CPS>(a, st, grid_used)
                : launch_backward5_t<Bwd5<false, false>, RawT, CPS>(a, st, grid_used);
}

}  // namespace r2l
