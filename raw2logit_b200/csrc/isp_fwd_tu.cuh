// isp_fwd_tu.cuh -- forward kernels + launcher for one raw element type (included by isp_fwd_f32.cu / isp_fwd_u16.cu)
#pragma once
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT, bool STATS>
__global__ void __launch_bounds__(Cfg::NT, 2) isp_forward_kernel(FwdArgs a, TileGrid grid) {
    extern __shared__ __align__(128) float smem[];
    fwd2_cta<Cfg, RawT, STATS, false>(blockIdx.x, gridDim.x, a, grid, smem);
}
// same kernel, raw window delivered by TMA (tensor map over raw as (W, H, B))
template <class Cfg, typename RawT, bool STATS>
__global__ void __launch_bounds__(Cfg::NT, 2) isp_forward_tma_kernel(FwdArgs a, TileGrid grid,
                                                                     const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smem[];
    fwd2_cta<Cfg, RawT, STATS, true>(blockIdx.x, gridDim.x, a, grid, smem, &tmap);
}

template <class Cfg, typename RawT, bool STATS>
static int launch_forward_t(const FwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);     // tiles of image pairs
    int g = 0;
    CUtensorMap tmap;
    if (make_raw_tensor_map(&tmap, a.raw, (int)sizeof(RawT), a.B, a.H, a.W, Cfg::P, Cfg::RH)) {
        int rc = persistent_grid(isp_forward_tma_kernel<Cfg, RawT, STATS>, Cfg::NT, Cfg::kSmemBytesTma, grid.n, &g);
        if (rc != R2L_OK) return rc;
        isp_forward_tma_kernel<Cfg, RawT, STATS><<<g, Cfg::NT, Cfg::kSmemBytesTma, st>>>(a, grid, tmap);
    } else {
        int rc = persistent_grid(isp_forward_kernel<Cfg, RawT, STATS>, Cfg::NT, Cfg::kSmemBytes, grid.n, &g);
        if (rc != R2L_OK) return rc;
        isp_forward_kernel<Cfg, RawT, STATS><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, grid);
    }
    if (grid_used) *grid_used = g;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

// third generation (TMA-fed, W % 4 == 0)
template <class Cfg, typename RawT, bool STATS, bool TAIL>
__global__ void __launch_bounds__(Cfg::NT, 2) isp_forward3_kernel(FwdArgs a, TileGrid grid,
                                                                  const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float smem[];
    fwd3_cta<Cfg, RawT, STATS, TAIL, true>(blockIdx.x, gridDim.x, a, grid, smem, &tmap);
}

template <class Cfg, typename RawT, bool STATS, bool TAIL>
static int launch_forward3_t(const FwdArgs& a, cudaStream_t st, int* grid_used) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);     // tiles of image pairs
    CUtensorMap tmap;
    if (!make_raw_tensor_map(&tmap, a.raw, (int)sizeof(RawT), a.B, a.H, a.W, Cfg::LW, Cfg::RH)) return kNotServed;
    int g = 0;
    int rc = persistent_grid(isp_forward3_kernel<Cfg, RawT, STATS, TAIL>, Cfg::NT, Cfg::kSmemBytesTma, grid.n, &g);
    if (rc != R2L_OK) return rc;
    cudaError_t e = launch_pdl(pdl_enabled(), isp_forward3_kernel<Cfg, RawT, STATS, TAIL>, g, Cfg::NT, Cfg::kSmemBytesTma, st, a, grid, tmap);
    if (grid_used) *grid_used = g;
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

template <typename RawT>
static int launch_forward_impl(const FwdArgs& a, bool stats, cudaStream_t st, int* grid_used, bool* fused_tail) {
    if (fused_tail) *fused_tail = false;
    const char* force = getenv("R2L_ISP_FORCE_GENERIC");            // debugging knob: second-generation kernel
    if (!(force && force[0] == '1') && fwd3_shape_ok(a.H, a.W) && aligned(a.out, 16) && aligned(a.additive, 16)) {
        int rc;
        if (stats) rc = launch_forward3_t<Fwd3Default, RawT, true, false>(a, st, grid_used);
        else if (a.additive || a.affine) rc = launch_forward3_t<Fwd3Default, RawT, false, true>(a, st, grid_used);
        else rc = launch_forward3_t<Fwd3Default, RawT, false, false>(a, st, grid_used);
        if (rc != kNotServed) {
            if (fused_tail) *fused_tail = rc == R2L_OK && stats && a.bn_sync != nullptr;
            return rc;
        }
    }
    return stats ? launch_forward_t<Fwd2Default, RawT, true>(a, st, grid_used)
                 : launch_forward_t<Fwd2Default, RawT, false>(a, st, grid_used);
}

}  // namespace r2l
