// isp_bwd6_tu.cuh -- sixth-generation backward kernels + launcher for one raw element type
#pragma once
#include "isp_launch.h"

namespace r2l {

template <class Cfg, typename RawT>
__global__ void __launch_bounds__(Cfg::NT, 1) isp_backward6_kernel(BwdArgs a, const __grid_constant__ CUtensorMap tmap_g,
                                                                   const __grid_constant__ CUtensorMap tmap_o) {
    extern __shared__ __align__(128) float smem[];
    bwd6_cta<Cfg, RawT>(a, smem, &tmap_g, &tmap_o);
}

template <class Cfg, typename RawT>
static int launch_backward6_t(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    CUtensorMap tg, to;
    if (!make_plane_tensor_map(&tg, a.gout, a.B, a.H, a.W, kB6SW, 2) || !make_plane_tensor_map(&to, a.out, a.B, a.H, a.W, kB6SW, 2))
        return kNotServed;
    int g = 0;
    int rc = tmem_ctas_per_device(reinterpret_cast<const void*>(isp_backward6_kernel<Cfg, RawT>), Cfg::NT, Cfg::kSmemBytes,
                                  Cfg::kTmemCols, 1, &g);
    if (rc != R2L_OK) return rc;
    const long long units = (long long)((a.B + 1) / 2) * ((a.W + kB6SW - 1) / kB6SW) * ((a.H + 1) / 2);
    if (g > units) g = (int)units;
    if (g > kMaxCtas) g = kMaxCtas;
    if (a.ticket) {
        cudaError_t e0 = cudaMemsetAsync(a.ticket, 0, sizeof(unsigned), st);
        if (e0 != cudaSuccess) return cuda_fail(e0);
    }
    isp_backward6_kernel<Cfg, RawT><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, tg, to);
    if (grid_used) *grid_used = g;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

// kNotServed when the shape or an alignment rule sends the call to an older generation
template <typename RawT>
static int launch_backward6_impl(const BwdArgs& a, cudaStream_t st, int* grid_used) {
    if (!a.out || !a.luma || !a.ticket || !bwd6_shape_ok(a.H, a.W)) return kNotServed;
    if (!aligned(a.gout, 16) || !aligned(a.graw, 16) || !aligned(a.additive, 16) || !aligned(a.out, 16) ||
        !aligned(a.luma, 16) || !aligned(a.raw, 4 * sizeof(RawT)))
        return kNotServed;                                                          // 128-bit rows, TMA base address
    if ((long long)a.B * 3 * a.H * a.W >= (1ll << 31)) return kNotServed;          // 32-bit element offsets inside the kernel
    const bool tail = a.gtail != nullptr;
    if (a.graw) return tail ? launch_backward6_t<Bwd6<true, true>, RawT>(a, st, grid_used)
                            : launch_backward6_t<Bwd6<true, false>, RawT>(a, st, grid_used);
    return tail ? launch_backward6_t<Bwd6<false, true>, RawT>(a, st, grid_used)
                : launch_backward6_t<Bwd6<false, false>, RawT>(a, st, grid_used);
}

}  // namespace r2l
