// r2l_torch.cpp -- torch operator shim over the C ABI (include/r2l_isp.h): TORCH_LIBRARY(raw2logit_isp) + the autograd
// node of the fused ISP in C++.  Built in-tree into raw2logit_b200/libr2l_torch.so by _build.py and loaded
// with torch.ops.load_library; no Python object crosses into libr2l_isp.so.
//
// What it replaces in the reference: the 79-node autograd graph ParametrizedProcessing.forward records per call
// (pipeline_torch.py:175-225) -- here one dispatcher call, one C++ autograd node, one kernel launch each way.
// Ops are registered for the CUDA key only: CPU tensors fail in the dispatcher; there is no fallback.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <cstdlib>
#include <map>
#include <mutex>

#include "r2l_isp.h"

namespace {

using at::Tensor;
using c10::optional;

constexpr int64_t kParamSizes[9] = {4, 3, 9, 1, 81, 9, 25, 9, 9};
const char* const kParamNames[9] = {"black_level", "white_balance", "colour_correction", "gamma_correct",
                                    "debayer_weight", "sharpen_weight", "gauss_weight", "rgb2yuv", "yuv2rgb"};
struct GradSlot { int64_t off, n; std::vector<int64_t> shape; };
const GradSlot kGradLayout[7] = {{R2L_G_BLACK_LEVEL, 4, {4}},       {R2L_G_WHITE_BALANCE, 3, {1, 3}},
                                 {R2L_G_COLOUR, 9, {3, 3}},         {R2L_G_GAMMA, 1, {1}},
                                 {R2L_G_DEBAYER, 81, {3, 3, 3, 3}}, {R2L_G_SHARPEN, 9, {1, 1, 3, 3}},
                                 {R2L_G_GAUSS, 25, {1, 1, 5, 5}}};

void check_rc(int rc, const char* what) {
    if (rc == R2L_OK) return;
    if (rc == R2L_ERR_CUDA)
        TORCH_CHECK(false, what, " failed: ", r2l_isp_error_string(rc), " (cudaError ", r2l_isp_last_cuda_error(), ")");
    TORCH_CHECK(false, what, " failed: ", r2l_isp_error_string(rc));
}

bool env_is(const char* name, char c) {
    const char* v = std::getenv(name);
    return v && v[0] == c;
}

Tensor f32c(const Tensor& t, int64_t n, const char* name) {
    TORCH_CHECK_TYPE(t.scalar_type() == at::kFloat && t.is_cuda(), name, " must be a float32 CUDA tensor, got ",
                     t.scalar_type(), " on ", t.device());
    TORCH_CHECK_VALUE(t.numel() == n, name, " must have ", n, " elements, got ", t.sizes());
    return t.contiguous();
}

struct Packed {
    r2l_isp_params p;
    Tensor keep[9];
};
void pack_params(const Tensor* const (&ts)[9], Packed& out) {
    const float** slots[9] = {&out.p.black_level, &out.p.white_balance, &out.p.colour_correction, &out.p.gamma_correct,
                              &out.p.debayer_weight, &out.p.sharpen_weight, &out.p.gauss_weight, &out.p.rgb2yuv,
                              &out.p.yuv2rgb};
    for (int i = 0; i < 9; ++i) {
        out.keep[i] = f32c(*ts[i], kParamSizes[i], kParamNames[i]);
        *slots[i] = out.keep[i].data_ptr<float>();
    }
}

// (tensor, dtype code): uint16 is ingested natively, every other dtype is computed in fp32 like the reference, whose
// output buffer is always fp32 (pipeline_torch.py:272)
Tensor raw_input(const Tensor& raw, int& code) {
    if (raw.scalar_type() == at::kUInt16) { code = R2L_U16; return raw.contiguous(); }
    code = R2L_F32;
    return (raw.scalar_type() == at::kFloat ? raw : raw.to(at::kFloat)).contiguous();
}

void check_shape(const Tensor& raw, int& b, int& h, int& w) {
    TORCH_CHECK(raw.dim() == 3, "needs dims (B, H, W), got ", raw.sizes());    // pipeline_torch.py:176 (asserted in Python)
    b = (int)raw.size(0); h = (int)raw.size(1); w = (int)raw.size(2);
    // the reference fails inside the Gaussian's reflect pad (pipeline_torch.py:165, 202)
    TORCH_CHECK(h >= 3 && w >= 3, "Padding size should be less than the corresponding input dimension, got H=", h, ", W=", w);
}

void* cur_stream(const Tensor& t) { return (void*)c10::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

// one scratch buffer per (device, stream), kept for the life of the process: r2l_isp_workspace_bytes is a fixed upper
// bound and calls on one stream run in order, so they can share it
std::mutex g_ws_mutex;
std::map<std::pair<int, void*>, Tensor> g_workspaces;
Tensor workspace(const Tensor& like, int b, int h, int w, size_t& nbytes) {
    nbytes = r2l_isp_workspace_bytes(b, h, w);
    const auto key = std::make_pair((int)like.get_device(), cur_stream(like));
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto it = g_workspaces.find(key);
    if (it == g_workspaces.end() || (size_t)it->second.numel() * 4 < nbytes) {
        Tensor buf = at::empty({(int64_t)(nbytes / 4)}, like.options().dtype(at::kFloat));
        g_workspaces[key] = buf;
        return buf;
    }
    return it->second;
}

// data-parallel gradient exchange fused into the backward kernel (parallel.enable_fused_gradient_exchange)
struct Exchange {
    bool on = false;
    int world = 1, rank = 0;
    int64_t peers = 0;
    float scale = 1.f;
};
std::mutex g_xch_mutex;
Exchange g_xch;

Tensor luma_buffer(const Tensor& raw, int code, int b, int h, int w, const Tensor& out, const float* add, bool save_luma) {
    if (save_luma && b > 0 && r2l_isp_luma_supported(raw.data_ptr(), code, b, h, w, out.data_ptr<float>(), add))
        return at::empty({2, (b + 1) / 2, h, w, 2}, out.options());
    return at::empty({0}, out.options());
}

// ---------------------------------------------------------------------------------------------------------------------
// plain ops (CUDA key)
// ---------------------------------------------------------------------------------------------------------------------
std::tuple<Tensor, Tensor> forward_cuda(const Tensor& raw_, const Tensor& bl, const Tensor& wb, const Tensor& ccm,
                                        const Tensor& gamma, const Tensor& wd, const Tensor& ws, const Tensor& wg,
                                        const Tensor& m1, const Tensor& m2, const optional<Tensor>& additive,
                                        const optional<Tensor>& affine, double raw_denominator, bool save_luma) {
    int b, h, w, code;
    check_shape(raw_, b, h, w);
    Tensor raw = raw_input(raw_, code);
    c10::cuda::CUDAGuard guard(raw.device());
    Packed pk;
    const Tensor* const ts[9] = {&bl, &wb, &ccm, &gamma, &wd, &ws, &wg, &m1, &m2};
    pack_params(ts, pk);
    Tensor add, aff;
    if (additive.has_value() && additive->defined()) add = f32c(*additive, (int64_t)3 * h * w, "additive");
    if (affine.has_value() && affine->defined()) aff = f32c(*affine, 6, "affine");
    r2l_isp_tail tail{add.defined() ? add.data_ptr<float>() : nullptr, aff.defined() ? aff.data_ptr<float>() : nullptr};
    Tensor out = at::empty({b, 3, h, w}, raw.options().dtype(at::kFloat));
    Tensor luma = luma_buffer(raw, code, b, h, w, out, tail.additive, save_luma);
    check_rc(r2l_isp_forward(raw.data_ptr(), code, (float)raw_denominator, b, h, w, &pk.p, &tail, out.data_ptr<float>(),
                             luma.numel() ? luma.data_ptr<float>() : nullptr, cur_stream(raw)),
             "r2l_isp_forward");
    return {out, luma};
}

std::tuple<Tensor, Tensor, Tensor> forward_bn_train_cuda(
    const Tensor& raw_, const Tensor& bl, const Tensor& wb, const Tensor& ccm, const Tensor& gamma, const Tensor& wd,
    const Tensor& ws, const Tensor& wg, const Tensor& m1, const Tensor& m2, const optional<Tensor>& additive,
    const optional<Tensor>& running_mean, const optional<Tensor>& running_var, const optional<Tensor>& num_batches_tracked,
    double momentum, double eps, double raw_denominator, bool save_luma) {
    int b, h, w, code;
    check_shape(raw_, b, h, w);
    TORCH_CHECK_VALUE((int64_t)b * h * w >= 2, "Expected more than 1 value per channel when training");
    Tensor raw = raw_input(raw_, code);
    c10::cuda::CUDAGuard guard(raw.device());
    Packed pk;
    const Tensor* const ts[9] = {&bl, &wb, &ccm, &gamma, &wd, &ws, &wg, &m1, &m2};
    pack_params(ts, pk);
    Tensor add;
    if (additive.has_value() && additive->defined()) add = f32c(*additive, (int64_t)3 * h * w, "additive");
    float *rm = nullptr, *rv = nullptr;
    auto stat_ptr = [](const optional<Tensor>& t, const char* name) -> float* {
        if (!t.has_value() || !t->defined()) return nullptr;
        TORCH_CHECK_TYPE(t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() == 3, name,
                         " must be a contiguous float32 tensor with 3 elements");
        return t->data_ptr<float>();
    };
    rm = stat_ptr(running_mean, "running_mean");
    rv = stat_ptr(running_var, "running_var");
    long long* nbt = nullptr;
    if (num_batches_tracked.has_value() && num_batches_tracked->defined()) {
        TORCH_CHECK_TYPE(num_batches_tracked->scalar_type() == at::kLong && num_batches_tracked->numel() == 1 &&
                         num_batches_tracked->is_cuda(), "num_batches_tracked must be one int64 on the device");
        nbt = reinterpret_cast<long long*>(num_batches_tracked->data_ptr<int64_t>());
    }
    Tensor out = at::empty({b, 3, h, w}, raw.options().dtype(at::kFloat));
    Tensor saved = at::empty({6}, out.options());
    size_t nbytes;
    Tensor wsb = workspace(out, b, h, w, nbytes);
    Tensor luma = luma_buffer(raw, code, b, h, w, out, add.defined() ? add.data_ptr<float>() : nullptr, save_luma);
    check_rc(r2l_isp_forward_bn_train(raw.data_ptr(), code, (float)raw_denominator, b, h, w, &pk.p,
                                      add.defined() ? add.data_ptr<float>() : nullptr, out.data_ptr<float>(), rm, rv, nbt,
                                      (float)momentum, (float)eps, saved.data_ptr<float>(),
                                      luma.numel() ? luma.data_ptr<float>() : nullptr, wsb.data_ptr(), nbytes,
                                      cur_stream(raw)),
             "r2l_isp_forward_bn_train");
    return {out, saved, luma};
}

// complete = false: with the full workspace the C ABI defers c1 / c2 of the tail to the backward kernel's prologue (no
// finish launch); complete = true asks for the finished 15 floats (the additive layer's gradient reads them on this side)
Tensor bn_backward_prepare_cuda(const Tensor& grad_out, const Tensor& out, const Tensor& saved_affine, bool complete) {
    TORCH_CHECK(out.dim() == 4, "out must be (B, 3, H, W)");
    const int b = (int)out.size(0), h = (int)out.size(2), w = (int)out.size(3);
    c10::cuda::CUDAGuard guard(out.device());
    Tensor g = f32c(grad_out, out.numel(), "grad_out"), y = f32c(out, out.numel(), "out");
    Tensor sa = f32c(saved_affine, 6, "saved_affine");
    Tensor tail = at::empty({15}, y.options());
    size_t nbytes;
    Tensor wsb = workspace(y, b, h, w, nbytes);
    if (complete || env_is("R2L_ISP_BN_TAIL_COMPLETE", '1'))                   // (the variable: debugging knob, tests compare the two)
        nbytes = (size_t)3 * 296 * 2 * sizeof(float);                          // the minimal workspace: separate finish kernel
    check_rc(r2l_isp_bn_backward_prepare(g.data_ptr<float>(), y.data_ptr<float>(), sa.data_ptr<float>(), b, h, w,
                                         tail.data_ptr<float>(), wsb.data_ptr(), nbytes, cur_stream(y)),
             "r2l_isp_bn_backward_prepare");
    return tail;
}

std::tuple<Tensor, Tensor> backward_cuda(const Tensor& raw_, const Tensor& bl, const Tensor& wb, const Tensor& ccm,
                                         const Tensor& gamma, const Tensor& wd, const Tensor& ws, const Tensor& wg,
                                         const Tensor& m1, const Tensor& m2, const Tensor& grad_out,
                                         const optional<Tensor>& grad_tail, const optional<Tensor>& additive,
                                         const optional<Tensor>& out, const optional<Tensor>& luma, bool need_raw_grad,
                                         double raw_denominator) {
    int b, h, w, code;
    check_shape(raw_, b, h, w);
    Tensor raw = raw_input(raw_, code);
    c10::cuda::CUDAGuard guard(raw.device());
    Packed pk;
    const Tensor* const ts[9] = {&bl, &wb, &ccm, &gamma, &wd, &ws, &wg, &m1, &m2};
    pack_params(ts, pk);
    const int64_t n_out = (int64_t)b * 3 * h * w;
    Tensor g = f32c(grad_out, n_out, "grad_out");
    Tensor gs, add, y, lum;
    if (grad_tail.has_value() && grad_tail->defined()) gs = f32c(*grad_tail, 15, "grad_tail");
    if (additive.has_value() && additive->defined()) add = f32c(*additive, (int64_t)3 * h * w, "additive");
    if (out.has_value() && out->defined()) y = f32c(*out, n_out, "out");
    if (luma.has_value() && luma->defined() && luma->numel())
        lum = f32c(*luma, (int64_t)r2l_isp_saved_luma_floats(b, h, w), "luma");
    auto opt = raw.options().dtype(at::kFloat);
    Tensor graw = need_raw_grad ? at::empty({b, h, w}, opt) : at::empty({0}, opt);
    Tensor gpar = at::empty({R2L_NUM_PARAM_GRADS}, opt);
    size_t nbytes;
    Tensor wsb = workspace(gpar, b, h, w, nbytes);
    auto fp = [](const Tensor& t) -> float* { return t.defined() ? t.data_ptr<float>() : nullptr; };
    Exchange x;
    {
        std::lock_guard<std::mutex> lock(g_xch_mutex);
        x = g_xch;
    }
    int rc;
    if (x.on) {
        // data-parallel: the 132 gradients leave the kernel already reduced over the ranks (parallel.PeerExchange)
        TORCH_CHECK(y.defined() && lum.defined(), "the fused gradient exchange needs the saved output and luma planes "
                    "(unset R2L_ISP_RECOMPUTE / R2L_ISP_NO_LUMA, shapes with W % 4 == 0)");
        // the kernel keeps the epoch in the exchange buffer: nothing per call comes from the host (graph replays)
        r2l_isp_allreduce d{x.world, x.rank, reinterpret_cast<float* const*>(x.peers), R2L_EPOCH_DEVICE, x.scale};
        rc = r2l_isp_backward_dp(raw.data_ptr(), code, (float)raw_denominator, b, h, w, &pk.p, g.data_ptr<float>(), fp(gs),
                                 fp(add), fp(y), fp(lum), need_raw_grad ? graw.data_ptr<float>() : nullptr,
                                 gpar.data_ptr<float>(), wsb.data_ptr(), nbytes, &d, cur_stream(raw));
    } else {
        rc = r2l_isp_backward(raw.data_ptr(), code, (float)raw_denominator, b, h, w, &pk.p, g.data_ptr<float>(), fp(gs),
                              fp(add), fp(y), fp(lum), need_raw_grad ? graw.data_ptr<float>() : nullptr,
                              gpar.data_ptr<float>(), wsb.data_ptr(), nbytes, cur_stream(raw));
    }
    check_rc(rc, "r2l_isp_backward");
    return {graw, gpar};
}

Tensor mosaic_cuda(const Tensor& raw_, const optional<Tensor>& black_level, bool reduce_size, int64_t out_channels,
                   double raw_denominator) {
    TORCH_CHECK(out_channels == 3 || out_channels == 4, "out_channels must be 3 or 4");   // :252 (asserted in Python)
    TORCH_CHECK_VALUE(raw_.dim() == 3, "needs dims (B, H, W), got ", raw_.sizes());
    const int b = (int)raw_.size(0), h = (int)raw_.size(1), w = (int)raw_.size(2);
    // reference: assigning ceil(H/2) rows into an H//2 buffer raises (pipeline_torch.py:261-265)
    TORCH_CHECK(!(reduce_size && (h % 2 || w % 2)), "The expanded size of the tensor must match the existing size: odd H=",
                h, " or W=", w, " with reduce_size=True");
    int code;
    Tensor raw = raw_input(raw_, code);
    c10::cuda::CUDAGuard guard(raw.device());
    Tensor bl;
    if (black_level.has_value() && black_level->defined()) bl = f32c(*black_level, 4, "black_level");
    Tensor out = reduce_size ? at::empty({b, out_channels, h / 2, w / 2}, raw.options().dtype(at::kFloat))
                             : at::empty({b, out_channels, h, w}, raw.options().dtype(at::kFloat));
    check_rc(r2l_isp_mosaic(raw.data_ptr(), code, (float)raw_denominator, b, h, w, bl.defined() ? bl.data_ptr<float>() : nullptr,
                            (int)reduce_size, (int)out_channels, out.data_ptr<float>(), cur_stream(raw)),
             "r2l_isp_mosaic");
    return out;
}

Tensor mosaic_backward_cuda(const Tensor& grad_out, int64_t h, int64_t w, bool reduce_size, int64_t out_channels) {
    const int b = (int)grad_out.size(0);
    c10::cuda::CUDAGuard guard(grad_out.device());
    Tensor g = f32c(grad_out, grad_out.numel(), "grad_out");
    Tensor graw = at::empty({b, h, w}, g.options());
    check_rc(r2l_isp_mosaic_backward(g.data_ptr<float>(), b, (int)h, (int)w, (int)reduce_size, (int)out_channels,
                                     graw.data_ptr<float>(), cur_stream(g)),
             "r2l_isp_mosaic_backward");
    return graw;
}

Tensor batch_sum_cuda(const Tensor& x, const optional<Tensor>& scale) {
    const int b = (int)x.size(0), c = (int)x.size(1);
    const int64_t hw = x.numel() / std::max<int64_t>((int64_t)b * c, 1);
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), sc;
    if (scale.has_value() && scale->defined()) sc = f32c(*scale, c, "scale");
    std::vector<int64_t> shape(x.sizes().begin(), x.sizes().end());
    shape[0] = 1;
    Tensor out = at::empty(shape, xc.options());
    check_rc(r2l_isp_batch_sum(xc.data_ptr<float>(), sc.defined() ? sc.data_ptr<float>() : nullptr, b, c, (int)hw,
                               out.data_ptr<float>(), cur_stream(xc)),
             "r2l_isp_batch_sum");
    return out;
}

// ---- augmentation + hand-off in one pass (utils/augmentation.py:70-74, model.py:79-82): dihedral index map, strides, dtype
Tensor dihedral_copy_cuda(const Tensor& src, std::vector<int64_t> map6, int64_t h_dst, int64_t w_dst, bool channels_last,
                          bool to_bf16) {
    TORCH_CHECK(src.dim() == 4 && src.is_cuda(), "dihedral_copy needs a (B, C, H, W) CUDA tensor");
    TORCH_CHECK(src.scalar_type() == at::kFloat || src.scalar_type() == at::kBFloat16, "float32 or bfloat16 source, got ",
                src.scalar_type());
    TORCH_CHECK(map6.size() == 6, "map6 = {a0, a1, a2, b0, b1, b2}");
    c10::cuda::CUDAGuard guard(src.device());
    const int b = (int)src.size(0), c = (int)src.size(1);
    Tensor dst = at::empty({b, c, h_dst, w_dst}, src.options().dtype(to_bf16 ? at::kBFloat16 : at::kFloat),
                           channels_last ? at::MemoryFormat::ChannelsLast : at::MemoryFormat::Contiguous);
    long long ss[4], ds[4];
    int m[6];
    for (int i = 0; i < 4; ++i) { ss[i] = src.stride(i); ds[i] = dst.stride(i); }
    for (int i = 0; i < 6; ++i) m[i] = (int)map6[i];
    check_rc(r2l_isp_dihedral_copy(src.data_ptr(), src.scalar_type() == at::kFloat ? 0 : 2, ss, dst.data_ptr(), to_bf16 ? 2 : 0, ds,
                                   b, c, (int)h_dst, (int)w_dst, (int)src.size(2), (int)src.size(3), m, cur_stream(src)),
             "r2l_isp_dihedral_copy");
    return dst;
}

// ---- numpy-compatible static pipeline (processing/pipeline_numpy.py:70-141), forward only ------------------------------
Tensor numpy_forward_cuda(const Tensor& raw_, std::vector<double> black_level, std::vector<double> white_balance,
                          std::vector<double> colour_matrix, bool sharpening_filter, bool gaussian_denoising,
                          double gaussian_sigma, double gamma, double raw_denominator) {
    TORCH_CHECK(raw_.dim() == 3, "needs dims (B, H, W), got ", raw_.sizes());
    TORCH_CHECK(black_level.size() == 4 && white_balance.size() == 3 && colour_matrix.size() == 9,
                "camera parameters: black_level[4], white_balance[3], colour_matrix[9]");
    const int b = (int)raw_.size(0), h = (int)raw_.size(1), w = (int)raw_.size(2);
    int code;
    Tensor raw = raw_input(raw_, code);
    c10::cuda::CUDAGuard guard(raw.device());
    float bl[4], wb[3], ccm[9];
    for (int i = 0; i < 4; ++i) bl[i] = (float)black_level[i];
    for (int i = 0; i < 3; ++i) wb[i] = (float)white_balance[i];
    for (int i = 0; i < 9; ++i) ccm[i] = (float)colour_matrix[i];
    Tensor out = at::empty({b, 3, h, w}, raw.options().dtype(at::kFloat));
    check_rc(r2l_isp_numpy_forward(raw.data_ptr(), code, (float)raw_denominator, b, h, w, bl, wb, ccm, (int)sharpening_filter,
                                   (int)gaussian_denoising, (float)gaussian_sigma, (float)gamma, out.data_ptr<float>(),
                                   cur_stream(raw)),
             "r2l_isp_numpy_forward");
    return out;
}

// ---- SSIM (utils/ssim.py:19-39): value from the per-tile partial sums, image gradients from the fused backward ----------
Tensor ssim_forward_cuda(const Tensor& img1, const Tensor& img2, int64_t window_size, bool size_average) {
    TORCH_CHECK(img1.dim() == 4 && img1.sizes() == img2.sizes(), "ssim needs two (B, C, H, W) tensors of one shape, got ",
                img1.sizes(), " and ", img2.sizes());
    const int b = (int)img1.size(0), c = (int)img1.size(1), h = (int)img1.size(2), w = (int)img1.size(3);
    c10::cuda::CUDAGuard guard(img1.device());
    Tensor x1 = f32c(img1, img1.numel(), "img1"), x2 = f32c(img2, img2.numel(), "img2");
    const int64_t tiles = b > 0 ? (int64_t)r2l_isp_ssim_partial_count(b, c, h, w) / ((int64_t)b * c) : 0;
    Tensor partial = at::empty({b, c * tiles}, x1.options().dtype(at::kDouble));
    check_rc(r2l_isp_ssim_forward(x1.data_ptr<float>(), x2.data_ptr<float>(), b, c, h, w, (int)window_size,
                                  partial.data_ptr<double>(), cur_stream(x1)),
             "r2l_isp_ssim_forward");
    const double per_image = (double)c * h * w;
    if (size_average) return (partial.sum() / (per_image * b)).to(at::kFloat);                    // ssim_map.mean()  :36
    return (partial.sum(1) / per_image).to(at::kFloat);                                            // per image       :38
}

std::tuple<Tensor, Tensor> ssim_backward_cuda(const Tensor& grad, const Tensor& img1, const Tensor& img2,
                                              int64_t window_size, bool size_average, bool need1, bool need2) {
    const int b = (int)img1.size(0), c = (int)img1.size(1), h = (int)img1.size(2), w = (int)img1.size(3);
    c10::cuda::CUDAGuard guard(img1.device());
    Tensor x1 = f32c(img1, img1.numel(), "img1"), x2 = f32c(img2, img2.numel(), "img2");
    const double per_image = (double)c * h * w;
    Tensor scale = size_average ? (grad.to(at::kFloat).reshape({1}) / (per_image * b)).expand({b}).contiguous()
                                : (grad.to(at::kFloat).reshape({b}) / per_image).contiguous();
    Tensor g1 = need1 ? at::empty_like(x1) : Tensor(), g2 = need2 ? at::empty_like(x2) : Tensor();
    check_rc(r2l_isp_ssim_backward(x1.data_ptr<float>(), x2.data_ptr<float>(), scale.data_ptr<float>(), b, c, h, w,
                                   (int)window_size, need1 ? g1.data_ptr<float>() : nullptr,
                                   need2 ? g2.data_ptr<float>() : nullptr, cur_stream(x1)),
             "r2l_isp_ssim_backward");
    return {need1 ? g1 : at::empty({0}, x1.options()), need2 ? g2 : at::empty({0}, x1.options())};
}

struct SsimFn : public torch::autograd::Function<SsimFn> {
    static Tensor forward(torch::autograd::AutogradContext* ctx, const Tensor& img1, const Tensor& img2, int64_t window_size,
                          bool size_average) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::ssim_forward", "")
                             .typed<decltype(ssim_forward_cuda)>();
        ctx->save_for_backward({img1, img2});
        ctx->saved_data["window_size"] = window_size;
        ctx->saved_data["size_average"] = size_average;
        ctx->saved_data["need1"] = img1.requires_grad();
        ctx->saved_data["need2"] = img2.requires_grad();
        return op.call(img1, img2, window_size, size_average);
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::ssim_backward", "")
                             .typed<decltype(ssim_backward_cuda)>();
        auto sv = ctx->get_saved_variables();
        const bool need1 = ctx->saved_data["need1"].toBool(), need2 = ctx->saved_data["need2"].toBool();
        torch::autograd::variable_list grads(4);
        if (need1 || need2) {
            auto r = op.call(grad_outputs[0], sv[0], sv[1], ctx->saved_data["window_size"].toInt(),
                             ctx->saved_data["size_average"].toBool(), need1, need2);
            if (need1) grads[0] = std::get<0>(r);
            if (need2) grads[1] = std::get<1>(r);
        }
        return grads;
    }
};
Tensor ssim_autograd(const Tensor& img1, const Tensor& img2, int64_t window_size, bool size_average) {
    return SsimFn::apply(img1, img2, window_size, size_average);
}

// ---- staged mode (track_stages=True, pipeline_torch.py:183-221): one repo kernel per stage (csrc/isp_stages.cu); the
// autograd nodes that string them together are Python (raw2logit_b200/staged.py) -- an inspection path, a few images per epoch
void stage_dims(const Tensor& x, const Tensor& weight, int& b, int& h, int& w, int& k) {
    TORCH_CHECK(x.dim() == 4 && x.size(1) == 3, "stage_conv needs a (B, 3, H, W) tensor, got ", x.sizes());
    TORCH_CHECK(weight.dim() == 4 && weight.size(0) == 3 && weight.size(1) == 3 && weight.size(2) == weight.size(3),
                "stage_conv needs a (3, 3, K, K) weight, got ", weight.sizes());
    b = (int)x.size(0); h = (int)x.size(2); w = (int)x.size(3); k = (int)weight.size(2);
}
Tensor stage_scratch(const Tensor& like, int k, size_t& nbytes) {
    nbytes = r2l_isp_stage_workspace_bytes(k);
    return at::empty({(int64_t)(nbytes / 8)}, like.options().dtype(at::kDouble));
}
Tensor stage_conv_cuda(const Tensor& x, const Tensor& weight, bool reflect) {
    int b, h, w, k;
    stage_dims(x, weight, b, h, w, k);
    // the reference fails inside F.pad / the conv's reflect padding for frames this small
    TORCH_CHECK(!reflect || (h > k / 2 && w > k / 2), "Padding size should be less than the corresponding input dimension, got H=",
                h, ", W=", w);
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), wc = f32c(weight, 9 * k * k, "weight");
    Tensor y = at::empty_like(xc);
    check_rc(r2l_isp_stage_conv(xc.data_ptr<float>(), wc.data_ptr<float>(), b, h, w, k, reflect ? 1 : 0, y.data_ptr<float>(),
                                cur_stream(xc)),
             "r2l_isp_stage_conv");
    return y;
}
std::tuple<Tensor, Tensor> stage_conv_backward_cuda(const Tensor& x, const Tensor& weight, const Tensor& grad_y, bool reflect,
                                                    bool need_x, bool need_w) {
    int b, h, w, k;
    stage_dims(x, weight, b, h, w, k);
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), wc = f32c(weight, 9 * k * k, "weight"), gc = f32c(grad_y, x.numel(), "grad_y");
    Tensor gx = need_x ? at::empty_like(xc) : at::empty({0}, xc.options());
    Tensor gw = need_w ? at::empty({3, 3, k, k}, xc.options()) : at::empty({0}, xc.options());
    size_t nbytes = 0;
    Tensor ws = stage_scratch(xc, k, nbytes);
    check_rc(r2l_isp_stage_conv_backward(xc.data_ptr<float>(), wc.data_ptr<float>(), gc.data_ptr<float>(), b, h, w, k,
                                         reflect ? 1 : 0, need_x ? gx.data_ptr<float>() : nullptr,
                                         need_w ? gw.data_ptr<float>() : nullptr, ws.data_ptr(), nbytes, cur_stream(xc)),
             "r2l_isp_stage_conv_backward");
    return {gx, gw};
}
Tensor stage_clip_cuda(const Tensor& x, double lo, double hi) {
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x");
    Tensor y = at::empty_like(xc);
    check_rc(r2l_isp_stage_clip(xc.data_ptr<float>(), (long long)xc.numel(), (float)lo, (float)hi, y.data_ptr<float>(),
                                cur_stream(xc)),
             "r2l_isp_stage_clip");
    return y;
}
Tensor stage_clip_backward_cuda(const Tensor& x, const Tensor& grad_y, double lo, double hi) {
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), gc = f32c(grad_y, x.numel(), "grad_y");
    Tensor gx = at::empty_like(xc);
    check_rc(r2l_isp_stage_clip_backward(xc.data_ptr<float>(), gc.data_ptr<float>(), (long long)xc.numel(), (float)lo, (float)hi,
                                         gx.data_ptr<float>(), cur_stream(xc)),
             "r2l_isp_stage_clip_backward");
    return gx;
}
Tensor stage_gamma_cuda(const Tensor& x, const Tensor& gamma) {
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), gm = f32c(gamma, 1, "gamma");
    Tensor y = at::empty_like(xc);
    check_rc(r2l_isp_stage_gamma(xc.data_ptr<float>(), gm.data_ptr<float>(), (long long)xc.numel(), y.data_ptr<float>(),
                                 cur_stream(xc)),
             "r2l_isp_stage_gamma");
    return y;
}
std::tuple<Tensor, Tensor> stage_gamma_backward_cuda(const Tensor& x, const Tensor& y, const Tensor& grad_y, const Tensor& gamma,
                                                     bool need_x) {
    c10::cuda::CUDAGuard guard(x.device());
    Tensor xc = f32c(x, x.numel(), "x"), yc = f32c(y, x.numel(), "y"), gc = f32c(grad_y, x.numel(), "grad_y");
    Tensor gm = f32c(gamma, 1, "gamma");
    Tensor gx = need_x ? at::empty_like(xc) : at::empty({0}, xc.options());
    Tensor gg = at::empty({1}, xc.options());
    size_t nbytes = 0;
    Tensor ws = stage_scratch(xc, 3, nbytes);
    check_rc(r2l_isp_stage_gamma_backward(xc.data_ptr<float>(), yc.data_ptr<float>(), gc.data_ptr<float>(), gm.data_ptr<float>(),
                                          (long long)xc.numel(), need_x ? gx.data_ptr<float>() : nullptr, gg.data_ptr<float>(),
                                          ws.data_ptr(), nbytes, cur_stream(xc)),
             "r2l_isp_stage_gamma_backward");
    return {gx, gg};
}

void set_exchange(int64_t world, int64_t rank, int64_t peers, double scale, bool on) {
    std::lock_guard<std::mutex> lock(g_xch_mutex);
    g_xch.on = on; g_xch.world = (int)world; g_xch.rank = (int)rank; g_xch.peers = peers; g_xch.scale = (float)scale;
}

// ---------------------------------------------------------------------------------------------------------------------
// the autograd node of the fused ISP: raw (B,H,W) + 7 parameter tensors + 2 buffers [+ additive] [+ BatchNorm tail]
// -> (B,3,H,W).  Saves raw, the (tiny) parameters, the output (which its consumer keeps alive anyway) and the Y0 / Y1
// planes the forward kernel computes on the way; the backward kernel recomputes nothing.
// bn_mode: 0 = no tail, 1 = eval (affine from the running statistics), 2 = train (batch statistics; running statistics
// updated in place by the kernel).
// ---------------------------------------------------------------------------------------------------------------------
const auto& op_forward() {
    static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::forward", "")
                         .typed<decltype(forward_cuda)>();
    return op;
}
const auto& op_forward_bn() {
    static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::forward_bn_train", "")
                         .typed<decltype(forward_bn_train_cuda)>();
    return op;
}
const auto& op_bn_prepare() {
    static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::bn_backward_prepare", "")
                         .typed<decltype(bn_backward_prepare_cuda)>();
    return op;
}
const auto& op_backward() {
    static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::backward", "")
                         .typed<decltype(backward_cuda)>();
    return op;
}
const auto& op_batch_sum() {
    static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::batch_sum", "")
                         .typed<decltype(batch_sum_cuda)>();
    return op;
}

struct FusedISPFn : public torch::autograd::Function<FusedISPFn> {
    static Tensor forward(torch::autograd::AutogradContext* ctx, const Tensor& raw, const Tensor& bl, const Tensor& wb,
                          const Tensor& ccm, const Tensor& gamma, const Tensor& wd, const Tensor& ws, const Tensor& wg,
                          const Tensor& m1, const Tensor& m2, const optional<Tensor>& additive, int64_t bn_mode,
                          const optional<Tensor>& running_mean, const optional<Tensor>& running_var,
                          const optional<Tensor>& num_batches_tracked, double momentum, double eps, double raw_denominator) {
        at::AutoDispatchBelowADInplaceOrView guard;
        const Tensor* const ps[7] = {&bl, &wb, &ccm, &gamma, &wd, &ws, &wg};
        bool any_param = false;
        for (auto* p : ps) any_param = any_param || p->requires_grad();
        const bool need_raw = raw.requires_grad();
        const bool need_add = additive.has_value() && additive->defined() && additive->requires_grad();
        // the backward reads the luma planes only together with the saved output (R2L_ISP_RECOMPUTE=1: neither)
        const bool save_luma = !env_is("R2L_ISP_RECOMPUTE", '1') && !env_is("R2L_ISP_NO_LUMA", '1') && (need_raw || any_param);
        Tensor out, luma, saved_affine;
        if (bn_mode == 2) {
            auto r = op_forward_bn().call(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, running_mean, running_var,
                                          num_batches_tracked, momentum, eps, raw_denominator, save_luma);
            out = std::get<0>(r); saved_affine = std::get<1>(r); luma = std::get<2>(r);
        } else {
            optional<Tensor> aff;
            if (bn_mode == 1) {
                TORCH_CHECK(running_mean.has_value() && running_var.has_value(), "eval-mode BatchNorm needs running statistics");
                Tensor scale = at::rsqrt(*running_var + eps);
                saved_affine = at::cat({scale, -(*running_mean) * scale});
                aff = saved_affine;
            }
            auto r = op_forward().call(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, aff, raw_denominator, save_luma);
            out = std::get<0>(r); luma = std::get<1>(r);
        }
        ctx->save_for_backward({raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2,
                                (additive.has_value() && additive->defined()) ? *additive : Tensor(),
                                saved_affine, out, luma});
        ctx->saved_data["bn_mode"] = bn_mode;
        ctx->saved_data["raw_denominator"] = raw_denominator;
        ctx->saved_data["need_raw"] = need_raw;
        ctx->saved_data["need_add"] = need_add;
        int64_t np = 0;
        for (int i = 0; i < 7; ++i) np |= (int64_t)(ps[i]->requires_grad() ? 1 : 0) << i;
        ctx->saved_data["need_params"] = np;
        return out;
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        at::AutoDispatchBelowADInplaceOrView guard;
        auto sv = ctx->get_saved_variables();
        const Tensor &raw = sv[0], &bl = sv[1], &wb = sv[2], &ccm = sv[3], &gamma = sv[4], &wd = sv[5], &ws = sv[6],
                     &wg = sv[7], &m1 = sv[8], &m2 = sv[9], &additive = sv[10], &saved_affine = sv[11], &out_saved = sv[12],
                     &luma = sv[13];
        const int64_t bn_mode = ctx->saved_data["bn_mode"].toInt();
        const double raw_denominator = ctx->saved_data["raw_denominator"].toDouble();
        const bool need_raw = ctx->saved_data["need_raw"].toBool(), need_add = ctx->saved_data["need_add"].toBool();
        const int64_t np = ctx->saved_data["need_params"].toInt();
        const bool any_param = np != 0;
        Tensor grad_out = grad_outputs[0].contiguous();
        torch::autograd::variable_list grads(18);
        // 15-float description of the tail the forward applied: {gs, c1, c2, ysc, ysh} x 3 channels (r2l_isp.h)
        Tensor tail;
        if (bn_mode == 2) {
            tail = op_bn_prepare().call(grad_out, out_saved, saved_affine, additive.defined() && need_add);
        } else if (bn_mode == 1) {
            Tensor zeros = at::zeros({6}, saved_affine.options());
            tail = at::cat({saved_affine.narrow(0, 0, 3), zeros, saved_affine});    // gs = ysc = scale, c1 = c2 = 0, ysh = shift
        } else if (additive.defined()) {
            Tensor one = at::ones({3}, grad_out.options()), zero = at::zeros({3}, grad_out.options());
            tail = at::cat({one, zero, zero, one, zero});                           // identity affine around the additive layer
        }
        const bool use_out = !env_is("R2L_ISP_RECOMPUTE", '1');
        if (need_raw || any_param) {
            optional<Tensor> o_tail, o_add, o_out, o_luma;
            if (tail.defined()) { o_tail = tail; if (additive.defined()) o_add = additive; }
            if (use_out) { o_out = out_saved; o_luma = luma; }
            auto r = op_backward().call(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, grad_out, o_tail, o_add, o_out, o_luma,
                                        need_raw, raw_denominator);
            const Tensor& graw = std::get<0>(r);
            const Tensor& gpar = std::get<1>(r);
            if (need_raw) grads[0] = raw.scalar_type() == at::kFloat ? graw : graw.to(raw.scalar_type());
            for (int i = 0; i < 7; ++i)
                if ((np >> i) & 1) grads[1 + i] = gpar.narrow(0, kGradLayout[i].off, kGradLayout[i].n).view(kGradLayout[i].shape);
        }
        if (additive.defined() && need_add) {
            if (bn_mode == 2) {
                // d/d(additive) = sum_b gs*(G - c1 - c2*yhat); yhat is the saved output
                Tensor gs = tail.narrow(0, 0, 3).view({1, 3, 1, 1}), c1 = tail.narrow(0, 3, 3).view({1, 3, 1, 1}),
                       c2 = tail.narrow(0, 6, 3).view({1, 3, 1, 1});
                Tensor geff = gs * (grad_out - c1 - c2 * out_saved);
                grads[10] = op_batch_sum().call(geff, c10::nullopt).view(additive.sizes());
            } else {
                optional<Tensor> scale;
                if (saved_affine.defined()) scale = saved_affine.narrow(0, 0, 3);
                grads[10] = op_batch_sum().call(grad_out, scale).view(additive.sizes());
            }
        }
        return grads;
    }
};

Tensor fused_autograd(const Tensor& raw, const Tensor& bl, const Tensor& wb, const Tensor& ccm, const Tensor& gamma,
                      const Tensor& wd, const Tensor& ws, const Tensor& wg, const Tensor& m1, const Tensor& m2,
                      const optional<Tensor>& additive, int64_t bn_mode, const optional<Tensor>& running_mean,
                      const optional<Tensor>& running_var, const optional<Tensor>& num_batches_tracked, double momentum,
                      double eps, double raw_denominator) {
    return FusedISPFn::apply(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, bn_mode, running_mean, running_var,
                             num_batches_tracked, momentum, eps, raw_denominator);
}

// the same forward without a graph: torch.inference_mode() skips the Autograd key
Tensor fused_cuda(const Tensor& raw, const Tensor& bl, const Tensor& wb, const Tensor& ccm, const Tensor& gamma,
                  const Tensor& wd, const Tensor& ws, const Tensor& wg, const Tensor& m1, const Tensor& m2,
                  const optional<Tensor>& additive, int64_t bn_mode, const optional<Tensor>& running_mean,
                  const optional<Tensor>& running_var, const optional<Tensor>& num_batches_tracked, double momentum,
                  double eps, double raw_denominator) {
    if (bn_mode == 2)
        return std::get<0>(forward_bn_train_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, running_mean,
                                                 running_var, num_batches_tracked, momentum, eps, raw_denominator, false));
    optional<Tensor> aff;
    if (bn_mode == 1) {
        TORCH_CHECK(running_mean.has_value() && running_var.has_value(), "eval-mode BatchNorm needs running statistics");
        Tensor scale = at::rsqrt(*running_var + eps);
        aff = at::cat({scale, -(*running_mean) * scale});
    }
    return std::get<0>(forward_cuda(raw, bl, wb, ccm, gamma, wd, ws, wg, m1, m2, additive, aff, raw_denominator, false));
}

// CFA split (raw2rgb, pipeline_torch.py:240-283), differentiable in raw and black_level like the reference
struct MosaicFn : public torch::autograd::Function<MosaicFn> {
    static Tensor forward(torch::autograd::AutogradContext* ctx, const Tensor& raw, const optional<Tensor>& black_level,
                          bool reduce_size, int64_t out_channels, double raw_denominator) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::mosaic", "").typed<decltype(mosaic_cuda)>();
        ctx->saved_data["h"] = raw.size(1);
        ctx->saved_data["w"] = raw.size(2);
        ctx->saved_data["reduce"] = reduce_size;
        ctx->saved_data["channels"] = out_channels;
        ctx->saved_data["raw_dtype"] = (int64_t)raw.scalar_type();
        ctx->saved_data["need_raw"] = raw.requires_grad();
        ctx->saved_data["need_bl"] = black_level.has_value() && black_level->defined() && black_level->requires_grad();
        return op.call(raw, black_level, reduce_size, out_channels, raw_denominator);
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        at::AutoDispatchBelowADInplaceOrView guard;
        static auto op = c10::Dispatcher::singleton().findSchemaOrThrow("raw2logit_isp::mosaic_backward", "")
                             .typed<decltype(mosaic_backward_cuda)>();
        const int64_t h = ctx->saved_data["h"].toInt(), w = ctx->saved_data["w"].toInt();
        Tensor g = op.call(grad_outputs[0].contiguous(), h, w, ctx->saved_data["reduce"].toBool(),
                           ctx->saved_data["channels"].toInt());
        torch::autograd::variable_list grads(5);
        if (ctx->saved_data["need_raw"].toBool()) {
            const auto dt = (at::ScalarType)ctx->saved_data["raw_dtype"].toInt();
            grads[0] = dt == at::kFloat ? g : g.to(dt);
        }
        if (ctx->saved_data["need_bl"].toBool()) {
            // d/d(black_level[par]) = -sum over the sites of CFA phase par of d/d(raw)   (raw2rgb :256-259)
            Tensor gb = at::empty({4}, g.options());
            for (int par = 0; par < 4; ++par)
                gb.select(0, par).copy_(-g.slice(1, par >> 1, c10::nullopt, 2).slice(2, par & 1, c10::nullopt, 2).sum());
            grads[1] = gb;
        }
        return grads;
    }
};

Tensor mosaic_autograd(const Tensor& raw, const optional<Tensor>& black_level, bool reduce_size, int64_t out_channels,
                       double raw_denominator) {
    return MosaicFn::apply(raw, black_level, reduce_size, out_channels, raw_denominator);
}

}  // namespace

#define R2L_PARAMS_SCHEMA                                                                                              \
    "Tensor black_level, Tensor white_balance, Tensor colour_correction, Tensor gamma_correct, Tensor debayer_weight, " \
    "Tensor sharpen_weight, Tensor gauss_weight, Tensor rgb2yuv, Tensor yuv2rgb"

TORCH_LIBRARY(raw2logit_isp, m) {
    m.def("forward(Tensor raw, " R2L_PARAMS_SCHEMA ", Tensor? additive, Tensor? affine, float raw_denominator, "
          "bool save_luma=False) -> (Tensor, Tensor)");
    m.def("forward_bn_train(Tensor raw, " R2L_PARAMS_SCHEMA ", Tensor? additive, Tensor(a!)? running_mean, "
          "Tensor(b!)? running_var, Tensor(c!)? num_batches_tracked, float momentum, float eps, float raw_denominator, "
          "bool save_luma=False) "
          "-> (Tensor, Tensor, Tensor)");
    m.def("bn_backward_prepare(Tensor grad_out, Tensor out, Tensor saved_affine, bool complete=False) -> Tensor");
    m.def("backward(Tensor raw, " R2L_PARAMS_SCHEMA ", Tensor grad_out, Tensor? grad_tail, Tensor? additive, "
          "Tensor? out, Tensor? luma, bool need_raw_grad, float raw_denominator) -> (Tensor, Tensor)");
    m.def("mosaic(Tensor raw, Tensor? black_level, bool reduce_size, int out_channels, float raw_denominator) -> Tensor");
    m.def("mosaic_backward(Tensor grad_out, int H, int W, bool reduce_size, int out_channels) -> Tensor");
    m.def("batch_sum(Tensor x, Tensor? scale) -> Tensor");
    // differentiable entry points (C++ autograd nodes)
    m.def("fused(Tensor raw, " R2L_PARAMS_SCHEMA ", Tensor? additive, int bn_mode, Tensor(a!)? running_mean, "
          "Tensor(b!)? running_var, Tensor(c!)? num_batches_tracked, float momentum, float eps, float raw_denominator) -> Tensor");
    m.def("mosaic_ad(Tensor raw, Tensor? black_level, bool reduce_size, int out_channels, float raw_denominator) -> Tensor");
    m.def("dihedral_copy(Tensor src, int[] map6, int h_dst, int w_dst, bool channels_last, bool to_bf16) -> Tensor");
    m.def("numpy_forward(Tensor raw, float[] black_level, float[] white_balance, float[] colour_matrix, bool sharpening_filter, "
          "bool gaussian_denoising, float gaussian_sigma, float gamma, float raw_denominator) -> Tensor");
    // SSIM regulariser (utils/ssim.py): plain ops + the differentiable entry point
    m.def("ssim_forward(Tensor img1, Tensor img2, int window_size, bool size_average) -> Tensor");
    m.def("ssim_backward(Tensor grad, Tensor img1, Tensor img2, int window_size, bool size_average, bool need1, "
          "bool need2) -> (Tensor, Tensor)");
    m.def("ssim(Tensor img1, Tensor img2, int window_size, bool size_average) -> Tensor");
    // staged mode (track_stages=True): one kernel per stage, csrc/isp_stages.cu; autograd nodes in raw2logit_b200/staged.py
    m.def("stage_conv(Tensor x, Tensor weight, bool reflect) -> Tensor");
    m.def("stage_conv_backward(Tensor x, Tensor weight, Tensor grad_y, bool reflect, bool need_x, bool need_w) -> (Tensor, Tensor)");
    m.def("stage_clip(Tensor x, float lo, float hi) -> Tensor");
    m.def("stage_clip_backward(Tensor x, Tensor grad_y, float lo, float hi) -> Tensor");
    m.def("stage_gamma(Tensor x, Tensor gamma) -> Tensor");
    m.def("stage_gamma_backward(Tensor x, Tensor y, Tensor grad_y, Tensor gamma, bool need_x) -> (Tensor, Tensor)");
    // process-wide state of the fused data-parallel exchange (parallel.enable_fused_gradient_exchange)
    m.def("set_exchange(int world, int rank, int peers, float scale, bool on) -> ()", &set_exchange);
}

TORCH_LIBRARY_IMPL(raw2logit_isp, CUDA, m) {
    m.impl("forward", &forward_cuda);
    m.impl("forward_bn_train", &forward_bn_train_cuda);
    m.impl("bn_backward_prepare", &bn_backward_prepare_cuda);
    m.impl("backward", &backward_cuda);
    m.impl("mosaic", &mosaic_cuda);
    m.impl("mosaic_backward", &mosaic_backward_cuda);
    m.impl("batch_sum", &batch_sum_cuda);
    m.impl("fused", &fused_cuda);
    m.impl("dihedral_copy", &dihedral_copy_cuda);
    m.impl("numpy_forward", &numpy_forward_cuda);
    m.impl("ssim_forward", &ssim_forward_cuda);
    m.impl("ssim_backward", &ssim_backward_cuda);
    m.impl("ssim", &ssim_forward_cuda);
    m.impl("mosaic_ad", &mosaic_cuda);
    m.impl("stage_conv", &stage_conv_cuda);
    m.impl("stage_conv_backward", &stage_conv_backward_cuda);
    m.impl("stage_clip", &stage_clip_cuda);
    m.impl("stage_clip_backward", &stage_clip_backward_cuda);
    m.impl("stage_gamma", &stage_gamma_cuda);
    m.impl("stage_gamma_backward", &stage_gamma_backward_cuda);
}

TORCH_LIBRARY_IMPL(raw2logit_isp, Autograd, m) {
    m.impl("fused", &fused_autograd);
    m.impl("mosaic_ad", &mosaic_autograd);
    m.impl("ssim", &ssim_autograd);
}
