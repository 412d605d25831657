// Python bindings (torch extension) for the sm_100a kernels and the native runtime.
#include <torch/extension.h>
#include <pybind11/stl.h>
#include <sstream>
#include <stdexcept>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "kernels/kernels.h"

namespace py = pybind11;
using torch::Tensor;

namespace ssb {
void bind_runtime(py::module_& m);   // csrc/runtime/bindings_runtime.cpp
}

static void check_mat(const Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be fp32");
    TORCH_CHECK(t.dim() == 2, name, " must be 2-D");
    TORCH_CHECK(t.size(0) > 0 && t.size(1) > 0, name, " is empty");
    TORCH_CHECK(t.stride(1) == 1 || t.size(1) == 1, name, " must have unit inner stride");
}
static int ld_of(const Tensor& t) { return (int)t.stride(0); }
static void cuda_ok(cudaError_t e, const char* what) {
    TORCH_CHECK(e == cudaSuccess, what, ": ", cudaGetErrorString(e));
}
static cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// y[rows, out] = relu?(x @ W^T + b)
static ssb::GemmLo make_lo(const c10::optional<Tensor>& a, const c10::optional<Tensor>& b, const c10::optional<Tensor>& out) {
    ssb::GemmLo lo;
    if (a.has_value() && b.has_value()) {
        lo.A = a->data_ptr<float>();
        lo.B = b->data_ptr<float>();
        lo.out = out.has_value() ? out->data_ptr<float>() : nullptr;
    }
    return lo;
}

static void split_lo(const Tensor& x, Tensor& lo) {
    TORCH_CHECK(x.is_cuda() && lo.is_cuda() && x.scalar_type() == torch::kFloat32 && lo.scalar_type() == torch::kFloat32);
    TORCH_CHECK(x.storage().nbytes() > 0 && x.numel() == lo.numel());
    c10::cuda::CUDAGuard guard(x.device());
    // operates on the padded storage of 2-D views (same strides for x and lo)
    const int64_t n = x.dim() == 2 ? x.size(0) * x.stride(0) : x.numel();
    cuda_ok(ssb::launch_split_lo(x.data_ptr<float>(), lo.data_ptr<float>(), n, cur_stream()), "split_lo");
}

// k_splits: 0 / 1 = plain kernel, -1 = let the planner decide, k >= 2 = force k (must not leave a split empty).
// The workspace and the tile counters are torch tensors owned by the caller's scope (stream-ordered reuse).
static void enable_splitk(ssb::GemmPlan& plan, int64_t k_splits, const Tensor& like, Tensor& ws, Tensor& counters) {
    if (k_splits == 0 || k_splits == 1) return;
    int splits = (int)k_splits;
    if (k_splits < 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        splits = ssb::gemm_splitk_choice(plan, sms);
        if (splits < 2) return;
    }
    ws = torch::empty({(int64_t)ssb::gemm_splitk_workspace_floats(plan, splits)}, like.options());
    counters = torch::zeros({(int64_t)plan.grid.x * plan.grid.y}, like.options().dtype(torch::kInt32));
    const char* err = ssb::gemm_plan_enable_splitk(&plan, splits, ws.data_ptr<float>(),
                                                   reinterpret_cast<unsigned int*>(counters.data_ptr<int32_t>()));
    TORCH_CHECK(err == nullptr, "split-K: ", err ? err : "");
}

static void linear_fwd(const Tensor& x, const Tensor& W, const c10::optional<Tensor>& bias, int64_t bias_stride,
                       bool relu, Tensor& y, const c10::optional<Tensor>& W_lo, const c10::optional<Tensor>& x_lo,
                       const c10::optional<Tensor>& y_lo, int64_t k_splits) {
    check_mat(x, "x"); check_mat(W, "W"); check_mat(y, "y");
    const int rows = x.size(0), in = x.size(1), out = W.size(0);
    TORCH_CHECK(W.size(1) == in && y.size(0) == rows && y.size(1) == out, "linear_fwd: shape mismatch");
    c10::cuda::CUDAGuard guard(x.device());
    ssb::GemmPlan plan;
    const char* err = ssb::gemm_plan_fwd(&plan, W.data_ptr<float>(), ld_of(W), x.data_ptr<float>(), ld_of(x),
                                         y.data_ptr<float>(), ld_of(y), rows, in, out,
                                         bias.has_value() ? bias->data_ptr<float>() : nullptr, (int)bias_stride, relu,
                                         make_lo(W_lo, x_lo, y_lo));
    TORCH_CHECK(err == nullptr, "linear_fwd: ", err ? err : "");
    Tensor ws, counters;
    enable_splitk(plan, k_splits, x, ws, counters);
    cuda_ok(ssb::gemm_launch(plan, cur_stream()), "linear_fwd launch");
}

// dx[rows, in] = (dz @ W) * (mask > 0)
static void linear_dgrad(const Tensor& dz, const Tensor& W, const c10::optional<Tensor>& mask, Tensor& dx,
                         const c10::optional<Tensor>& W_lo, const c10::optional<Tensor>& dz_lo, const c10::optional<Tensor>& dx_lo,
                         int64_t k_splits) {
    check_mat(dz, "dz"); check_mat(W, "W"); check_mat(dx, "dx");
    const int rows = dz.size(0), out = dz.size(1), in = W.size(1);
    TORCH_CHECK(W.size(0) == out && dx.size(0) == rows && dx.size(1) == in, "linear_dgrad: shape mismatch");
    if (mask.has_value()) { check_mat(*mask, "mask"); TORCH_CHECK(mask->size(0) == rows && mask->size(1) == in, "mask shape"); }
    c10::cuda::CUDAGuard guard(dz.device());
    ssb::GemmPlan plan;
    const char* err = ssb::gemm_plan_dgrad(&plan, W.data_ptr<float>(), ld_of(W), dz.data_ptr<float>(), ld_of(dz),
                                           dx.data_ptr<float>(), ld_of(dx), rows, in, out,
                                           mask.has_value() ? mask->data_ptr<float>() : nullptr,
                                           mask.has_value() ? ld_of(*mask) : 0, make_lo(W_lo, dz_lo, dx_lo));
    TORCH_CHECK(err == nullptr, "linear_dgrad: ", err ? err : "");
    Tensor ws, counters;
    enable_splitk(plan, k_splits, dz, ws, counters);
    cuda_ok(ssb::gemm_launch(plan, cur_stream()), "linear_dgrad launch");
}

// G[out, in] (+)= dz^T @ x ; db[out] (+)= colsum(dz) ; optionally W -= lr * (G + ...)
static void linear_wgrad(const Tensor& dz, const Tensor& x, Tensor& G, bool accumulate, const c10::optional<Tensor>& db,
                         int64_t db_stride, const c10::optional<Tensor>& W, double lr, bool fuse_sgd,
                         const c10::optional<Tensor>& dz_lo, const c10::optional<Tensor>& x_lo) {
    check_mat(dz, "dz"); check_mat(x, "x"); check_mat(G, "G");
    const int rows = dz.size(0), out = dz.size(1), in = x.size(1);
    TORCH_CHECK(x.size(0) == rows && G.size(0) == out && G.size(1) == in, "linear_wgrad: shape mismatch");
    c10::cuda::CUDAGuard guard(dz.device());
    ssb::GemmPlan plan;
    const char* err = ssb::gemm_plan_wgrad(&plan, dz.data_ptr<float>(), ld_of(dz), x.data_ptr<float>(), ld_of(x),
                                           G.data_ptr<float>(), ld_of(G), rows, in, out, accumulate,
                                           db.has_value() ? db->data_ptr<float>() : nullptr, (int)db_stride,
                                           W.has_value() ? W->data_ptr<float>() : nullptr,
                                           W.has_value() ? ld_of(*W) : 0, (float)lr, fuse_sgd, make_lo(dz_lo, x_lo, c10::nullopt));
    TORCH_CHECK(err == nullptr, "linear_wgrad: ", err ? err : "");
    cuda_ok(ssb::gemm_launch(plan, cur_stream()), "linear_wgrad launch");
}

static void loss_head(const Tensor& logits, const c10::optional<Tensor>& target, const c10::optional<Tensor>& probs,
                      const c10::optional<Tensor>& dlogits, const c10::optional<Tensor>& loss_out, double inv_batch) {
    check_mat(logits, "logits");
    c10::cuda::CUDAGuard guard(logits.device());
    const int rows = logits.size(0), cols = logits.size(1);
    cuda_ok(ssb::launch_loss_head(logits.data_ptr<float>(), ld_of(logits),
                                  target.has_value() ? target->data_ptr<float>() : nullptr, target.has_value() ? ld_of(*target) : 0,
                                  probs.has_value() ? probs->data_ptr<float>() : nullptr, probs.has_value() ? ld_of(*probs) : 0,
                                  dlogits.has_value() ? dlogits->data_ptr<float>() : nullptr, dlogits.has_value() ? ld_of(*dlogits) : 0,
                                  loss_out.has_value() ? loss_out->data_ptr<float>() : nullptr, rows, cols, (float)inv_batch,
                                  cur_stream()), "loss_head");
}

static void softmax_grad(const Tensor& logits, const Tensor& upstream, Tensor& dlogits) {
    check_mat(logits, "logits"); check_mat(upstream, "upstream"); check_mat(dlogits, "dlogits");
    c10::cuda::CUDAGuard guard(logits.device());
    cuda_ok(ssb::launch_softmax_grad(logits.data_ptr<float>(), ld_of(logits), upstream.data_ptr<float>(), ld_of(upstream),
                                     dlogits.data_ptr<float>(), ld_of(dlogits), logits.size(0), logits.size(1), cur_stream()),
            "softmax_grad");
}

static void relu_mask_(Tensor& g, const Tensor& y) {
    check_mat(g, "g"); check_mat(y, "y");
    c10::cuda::CUDAGuard guard(g.device());
    cuda_ok(ssb::launch_relu_mask(g.data_ptr<float>(), ld_of(g), y.data_ptr<float>(), ld_of(y), g.size(0), g.size(1), cur_stream()), "relu_mask");
}
static void relu_fwd(const Tensor& x, Tensor& y) {
    TORCH_CHECK(x.is_contiguous() && y.is_contiguous() && x.numel() == y.numel());
    c10::cuda::CUDAGuard guard(x.device());
    cuda_ok(ssb::launch_relu_fwd(x.data_ptr<float>(), y.data_ptr<float>(), x.numel(), cur_stream()), "relu_fwd");
}
static void axpby(const Tensor& x, const Tensor& t, Tensor& y, double a, double b) {
    TORCH_CHECK(x.is_contiguous() && t.is_contiguous() && y.is_contiguous() && x.numel() == y.numel() && t.numel() == y.numel());
    c10::cuda::CUDAGuard guard(x.device());
    cuda_ok(ssb::launch_axpby(x.data_ptr<float>(), t.data_ptr<float>(), y.data_ptr<float>(), (float)a, (float)b, x.numel(), cur_stream()), "axpby");
}
static void sgd_(Tensor& w, const Tensor& g, double lr) {
    TORCH_CHECK(w.is_contiguous() && g.is_contiguous() && w.numel() == g.numel());
    c10::cuda::CUDAGuard guard(w.device());
    cuda_ok(ssb::launch_sgd(w.data_ptr<float>(), g.data_ptr<float>(), (float)lr, w.numel(), cur_stream()), "sgd");
}
static void argmax_correct(const Tensor& pred, const Tensor& target, Tensor& correct) {
    check_mat(pred, "pred"); check_mat(target, "target");
    TORCH_CHECK(correct.scalar_type() == torch::kInt32 && correct.is_cuda());
    c10::cuda::CUDAGuard guard(pred.device());
    cuda_ok(ssb::launch_argmax_correct(pred.data_ptr<float>(), ld_of(pred), target.data_ptr<float>(), ld_of(target),
                                       pred.size(0), pred.size(1), correct.data_ptr<int>(), cur_stream()), "argmax_correct");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "shallowspeed_b200 native module: sm_100a kernels + C++ pipeline runtime";
    m.def("split_lo", &split_lo);
    m.def("linear_fwd", &linear_fwd, py::arg("x"), py::arg("W"), py::arg("bias"), py::arg("bias_stride"), py::arg("relu"),
          py::arg("y"), py::arg("W_lo") = py::none(), py::arg("x_lo") = py::none(), py::arg("y_lo") = py::none(),
          py::arg("k_splits") = 0);
    m.def("linear_dgrad", &linear_dgrad, py::arg("dz"), py::arg("W"), py::arg("mask"), py::arg("dx"),
          py::arg("W_lo") = py::none(), py::arg("dz_lo") = py::none(), py::arg("dx_lo") = py::none(), py::arg("k_splits") = 0);
    m.def("linear_wgrad", &linear_wgrad, py::arg("dz"), py::arg("x"), py::arg("G"), py::arg("accumulate"), py::arg("db"),
          py::arg("db_stride"), py::arg("W"), py::arg("lr"), py::arg("fuse_sgd"), py::arg("dz_lo") = py::none(),
          py::arg("x_lo") = py::none());
    m.def("loss_head", &loss_head);
    m.def("softmax_grad", &softmax_grad);
    m.def("relu_mask_", &relu_mask_);
    m.def("relu_fwd", &relu_fwd);
    m.def("axpby", &axpby);
    m.def("sgd_", &sgd_);
    m.def("argmax_correct", &argmax_correct);
    m.def("gemm_kernel_count", &ssb::gemm_kernel_count);
    m.def("chain_budget", [](int mb_rows, bool split) {
        int kps = 0, stages = 0, smem = 0;
        const bool ok = ssb::chain_budget(mb_rows, split, &kps, &stages, &smem);
        return std::make_tuple(ok, kps, stages, smem);
    });
    m.def("chain_eligible", [](const std::vector<std::pair<int, int>>& in_out, int mb_rows, int out_dim, bool has_loss, bool split) {
        std::vector<ssb::ChainLayer> ls(in_out.size());
        for (size_t i = 0; i < in_out.size(); ++i) { ls[i] = ssb::ChainLayer{}; ls[i].in = in_out[i].first; ls[i].out = in_out[i].second; }
        return ssb::chain_eligible(ls.data(), (int)ls.size(), mb_rows, out_dim, has_loss, split);
    });
    m.def("dp_layer_geometry", [](int in, int out, int dp, bool one_shot) {
        int block_n = 0, tm = 0, tn = 0;
        int64_t slots = 0, slot_floats = 0;
        ssb::dp_layer_geometry(in, out, dp, &block_n, &tm, &tn, &slots, &slot_floats, one_shot ? 1 : 0);
        return std::make_tuple(block_n, tm, tn, slots, slot_floats);
    });
    // host-only planning logic, callable without a GPU (tests/test_splitk_planning.py):
    // (k_splits chosen, k-blocks per split, error string of enable() with dummy buffers or "")
    m.def("splitk_plan", [](int m_total, int n_rows, int k_total, int num_sms, int force) {
        ssb::GemmPlan plan{};
        plan.mode = ssb::GEMM_FWD;
        plan.p.m_total = m_total; plan.p.n_total = n_rows; plan.p.k_total = k_total;
        plan.p.block_n = n_rows >= 256 ? 256 : (n_rows + 15) / 16 * 16;
        plan.grid = dim3((m_total + 127) / 128, (n_rows + plan.p.block_n - 1) / plan.p.block_n, 1);
        int splits = force > 0 ? force : ssb::gemm_splitk_choice(plan, num_sms);
        const int num_kb = (k_total + 31) / 32;
        std::string err;
        if (splits > 1) {
            float dummy_ws;
            unsigned int dummy_ctr;
            const char* e = ssb::gemm_plan_enable_splitk(&plan, splits, &dummy_ws, &dummy_ctr);
            if (e) err = e;
        }
        const int per = splits > 0 ? (num_kb + splits - 1) / splits : num_kb;
        return std::make_tuple(splits, per, (int)plan.grid.z, plan.p.stages, plan.smem_bytes, err);
    });
    // LL data-parallel kernel: host-side geometry (tiles per layer, lines per landing zone), callable without a GPU
    m.def("dp_ll_tiles", [](int in, int out) { return ssb::dp_ll_tiles(in, out); });
    m.def("dp_ll_zone_lines", [](int dp, int n_tiles) { return (int64_t)ssb::dp_ll_zone_lines(dp, n_tiles); });
    // Exercises the C++ runtime features that break when libstdc++ is linked statically next to torch's dynamic
    // copy (the round-1 GPU-suite segfault): locale-dependent integer / float formatting and an exception that
    // crosses a function boundary.  Callable without a GPU (tests/test_native_linkage.py).
    m.def("selftest_format", [](int v) {
        std::ostringstream os;
        os << "v=" << v << " f=" << 1.5 << " h=" << std::hex << 255;
        try {
            throw std::runtime_error(os.str());
        } catch (const std::exception& e) {
            return std::string(e.what());
        }
    });
    ssb::bind_runtime(m);
}
