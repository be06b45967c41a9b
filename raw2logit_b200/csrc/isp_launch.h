// isp_launch.h -- host-side helpers shared by the translation units of libr2l_isp.so (not part of the C ABI).
#pragma once
#include <cuda.h>            // CUtensorMap types only; the encoder is fetched through the runtime (no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/r2l_isp.h"
#include "isp_config.h"

namespace r2l {

int cuda_fail(cudaError_t e);                      // records the error for r2l_isp_last_cuda_error(), returns R2L_ERR_CUDA

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// persistent grid: one wave of resident CTAs (or fewer when the job is small).  The occupancy query and the
// dynamic-shared-memory opt-in are done once per (kernel, device) and remembered (isp_host.cu).
int cached_ctas_per_device(const void* kernel, int threads, size_t smem, int* out);   // fills *out, returns R2L_* code
// The same for a kernel that allocates tensor memory.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every
// kernel that contains tcgen05.alloc, although the hardware co-schedules CTAs as long as their column allocations fit
// the SM's 512 (profiles/microbench/tmem_occupancy.cu), so the limit is taken from the kernel's own resource use:
// registers, shared memory, TMEM columns and the CTAs per SM it was compiled for (want_per_sm).
int tmem_ctas_per_device(const void* kernel, int threads, size_t smem, int tmem_cols, int want_per_sm, int* out);

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int n_tiles, int* grid_out) {
    int g = 0;
    int rc = cached_ctas_per_device(reinterpret_cast<const void*>(kernel), threads, smem, &g);
    if (rc != R2L_OK) return rc;
    if (g > n_tiles) g = n_tiles;
    if (g > kMaxCtas) g = kMaxCtas;
    *grid_out = g;
    return R2L_OK;
}

// launch tag of the fused finish's ticket word: 1, 2, 3, ... (never 0), process-wide
unsigned next_ticket_generation();

// <<<grid, threads, smem, st>>> with programmatic stream serialization allowed: the kernel may become resident while the
// kernel before it on the stream drains; it orders itself behind that kernel with griddepcontrol.wait (pdl_wait()).
// R2L_ISP_NO_PDL=1 launches plainly (debugging knob).
bool pdl_enabled();
bool pdl_enabled_backward();
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool allow, void (*kernel)(KArgs...), int grid, int threads, size_t smem, cudaStream_t st,
                              Args... args) {
    if (!allow) {
        kernel<<<grid, threads, smem, st>>>(KArgs(args)...);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// tensor map over the raw batch: dims (W, H, B), box (box_w, box_h, 2), zero fill outside.
// false when the shape/pointer does not meet TMA's 16-byte rules (then a non-TMA kernel runs)
bool make_raw_tensor_map(CUtensorMap* map, const void* raw, int elem_bytes, int B, int H, int W, int box_w, int box_h);

// tensor maps for the fifth-generation backward's L2 prefetch of the next tile (isp_bwd5.cuh); false: no prefetch
struct Bwd5Maps;
struct BwdArgs;
bool make_bwd5_prefetch_maps(Bwd5Maps* maps, const BwdArgs& a, int raw_elem_bytes, int th, int tw);

// launchers, one translation unit each (compiled in parallel by _build.py)
// *fused_tail (may be null): the launch ran the fused train-mode BatchNorm tail itself (FwdArgs::bn_sync, third generation)
int launch_forward_f32(const FwdArgs& a, bool stats, cudaStream_t st, int* grid_used, bool* fused_tail = nullptr);
int launch_forward_u16(const FwdArgs& a, bool stats, cudaStream_t st, int* grid_used, bool* fused_tail = nullptr);
// third-generation backward (TMA-fed); returns kNotServed when the shape / alignment is not served
int launch_backward3_f32(const BwdArgs& a, cudaStream_t st, int* grid_used);
int launch_backward3_u16(const BwdArgs& a, cudaStream_t st, int* grid_used);
// fifth-generation backward (forward output + saved luma planes, running sums in tensor memory, fused finish)
int launch_backward5_f32(const BwdArgs& a, cudaStream_t st, int* grid_used);
int launch_backward5_u16(const BwdArgs& a, cudaStream_t st, int* grid_used);
// generic scalar kernels: any shape, any alignment
int launch_backward_generic(const BwdArgs& a, int raw_dtype, cudaStream_t st, int* grid_used);
constexpr int kNotServed = 1;

}  // namespace r2l
