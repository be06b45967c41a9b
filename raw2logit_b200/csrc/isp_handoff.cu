// isp_handoff.cu -- augmentation + hand-off of the processed RGB batch to the task model in ONE pass.
//
// Reference: utils/augmentation.py:70-74 (`augmentation_weak` = RandomHorizontalFlip, RandomVerticalFlip, RandomRotate90
// :8-11), applied by LitModel.forward right after the processor (model.py:79-82), followed by whatever layout / dtype
// conversion the task model wants (channels_last, bf16 for the ResNet stem).  Stock PyTorch runs up to five full
// read+write passes over the (B, 3, H, W) tensor for that (flip, flip, rot90 = flip + transposed copy, .contiguous(
// channels_last), .to(bf16)).  The three augmentations are elements of the dihedral group acting on the last two
// dimensions, so their composition is ONE index map  (y_in, x_in) = (a0 + a1 y + a2 x, b0 + b1 y + b2 x):  this kernel
// reads the source once and writes the augmented tensor once, in the requested strides (NCHW or channels_last) and dtype
// (fp32 or bf16).  Its adjoint is the same kernel with the inverse map, source = the incoming gradient (any strides /
// dtype), destination = contiguous fp32.  Bit-exact against the stock ops (pure permutation; bf16 by round-to-nearest-even
// like Tensor.to).  HBM-bound: 12 B/px read + 12 (6) B/px written.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "isp_launch.h"

namespace r2l {

struct DihedralMap { int a0, a1, a2, b0, b1, b2; };          // source (y, x) of destination (y, x)
struct Strides4 { long long b, c, y, x; };                    // element strides

template <typename T> __device__ __forceinline__ float load_as_float(const T* p);
template <> __device__ __forceinline__ float load_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void store_from_float(T* p, float v);
template <> __device__ __forceinline__ void store_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// One CTA = one 32 x 32 destination tile of one image, all channels.  The source footprint of the tile is a 32 x 32
// square as well (flipped and / or transposed); it is read along the SOURCE rows (coalesced whatever the map) into
// shared memory and written along the destination rows.
template <typename InT, typename OutT, int CMAX>
__global__ void __launch_bounds__(256) dihedral_copy_kernel(const InT* __restrict__ src, Strides4 ss, OutT* __restrict__ dst,
                                                            Strides4 ds, int C, int Hd, int Wd, int Hs, int Ws, DihedralMap m,
                                                            int tiles_x, int tiles_per_image) {
    __shared__ float tile[CMAX][32][33];
    const int b = blockIdx.x / tiles_per_image, t = blockIdx.x - b * tiles_per_image;
    const int y0 = (t / tiles_x) * 32, x0 = (t % tiles_x) * 32;
    const bool transposed = m.a1 == 0;                       // the destination row index drives the source column
    // source corner of the tile: the source coordinates of the four destination corners bound a 32 x 32 square
    const int ya = m.a0 + m.a1 * y0 + m.a2 * x0, yb = m.a0 + m.a1 * (y0 + 31) + m.a2 * (x0 + 31);
    const int xa = m.b0 + m.b1 * y0 + m.b2 * x0, xb = m.b0 + m.b1 * (y0 + 31) + m.b2 * (x0 + 31);
    const int sy0 = ya < yb ? ya : yb, sx0 = xa < xb ? xa : xb;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int c = 0; c < C; ++c)
        for (int r = ty; r < 32; r += 8) {
            const int sy = sy0 + r, sx = sx0 + tx;
            float v = 0.f;
            if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) v = load_as_float<InT>(src + b * ss.b + c * ss.c + sy * ss.y + sx * ss.x);
            tile[c][r][tx] = v;
        }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int y = y0 + r, x = x0 + tx;
        if (y >= Hd || x >= Wd) continue;
        const int sy = m.a0 + m.a1 * y + m.a2 * x - sy0, sx = m.b0 + m.b1 * y + m.b2 * x - sx0;
        OutT* o = dst + b * ds.b + y * ds.y + x * ds.x;
        for (int c = 0; c < C; ++c) store_from_float<OutT>(o + c * ds.c, tile[c][sy][sx]);
    }
    (void)transposed;
}

template <typename InT, typename OutT>
static int launch_dihedral(const void* src, Strides4 ss, void* dst, Strides4 ds, int B, int C, int Hd, int Wd, int Hs, int Ws,
                           DihedralMap m, cudaStream_t st) {
    const int tiles_x = (Wd + 31) / 32, tiles_y = (Hd + 31) / 32;
    const long long blocks = (long long)B * tiles_x * tiles_y;
    if (blocks > 0x7fffffffLL) return R2L_ERR_BAD_SHAPE;
    dihedral_copy_kernel<InT, OutT, 4><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const InT*>(src), ss, static_cast<OutT*>(dst), ds,
                                                                        C, Hd, Wd, Hs, Ws, m, tiles_x, tiles_x * tiles_y);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

}  // namespace r2l

using namespace r2l;

extern "C" int r2l_isp_dihedral_copy(const void* src, int src_dtype, const long long* src_strides, void* dst, int dst_dtype,
                                     const long long* dst_strides, int B, int C, int H_dst, int W_dst, int H_src, int W_src,
                                     const int* map6, void* stream) {
    if (B < 0 || C < 1 || C > 4 || H_dst < 1 || W_dst < 1 || H_src < 1 || W_src < 1) return R2L_ERR_BAD_SHAPE;
    if (B == 0) return R2L_OK;
    if (!src || !dst || !src_strides || !dst_strides || !map6) return R2L_ERR_NULL_POINTER;
    if ((src_dtype != 0 && src_dtype != 2) || (dst_dtype != 0 && dst_dtype != 2)) return R2L_ERR_BAD_DTYPE;   // 0 = f32, 2 = bf16
    const DihedralMap m{map6[0], map6[1], map6[2], map6[3], map6[4], map6[5]};
    // a dihedral map: one of (a1, a2) and one of (b1, b2) is +-1, the others 0, and rows / columns do not share a driver
    const bool plain = (m.a1 == 1 || m.a1 == -1) && m.a2 == 0 && m.b1 == 0 && (m.b2 == 1 || m.b2 == -1);
    const bool trans = m.a1 == 0 && (m.a2 == 1 || m.a2 == -1) && (m.b1 == 1 || m.b1 == -1) && m.b2 == 0;
    if (!plain && !trans) return R2L_ERR_BAD_ARGUMENT;
    const Strides4 ss{src_strides[0], src_strides[1], src_strides[2], src_strides[3]};
    const Strides4 ds{dst_strides[0], dst_strides[1], dst_strides[2], dst_strides[3]};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (src_dtype == 0 && dst_dtype == 0) return launch_dihedral<float, float>(src, ss, dst, ds, B, C, H_dst, W_dst, H_src, W_src, m, st);
    if (src_dtype == 0) return launch_dihedral<float, __nv_bfloat16>(src, ss, dst, ds, B, C, H_dst, W_dst, H_src, W_src, m, st);
    if (dst_dtype == 0) return launch_dihedral<__nv_bfloat16, float>(src, ss, dst, ds, B, C, H_dst, W_dst, H_src, W_src, m, st);
    return launch_dihedral<__nv_bfloat16, __nv_bfloat16>(src, ss, dst, ds, B, C, H_dst, W_dst, H_src, W_src, m, st);
}
