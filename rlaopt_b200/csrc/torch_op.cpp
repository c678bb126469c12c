// torch.library front of the C ABI: `torch.ops.rlaopt.kernel_matmat`.
//
// Registration pattern of the reference's own ops (rlaopt/csrc/cpp/csc_matmat.cpp:83-87: schema in a
// TORCH_LIBRARY_FRAGMENT(rlaopt, m), implementation per dispatch key); the op is the whole of
// _KernelLinOp's matvec / rmatvec / row_oracle / blk_oracle (rlaopt/kernels/base.py:43-47, 104-128):
//
//   kernel_matmat(A1, A2, V, kernel_id, lengthscale, lengthscale_vec?, const_scaling, transpose, row_idx?, col_idx?)
//       -> const_scaling * K(A1[row_idx], A2[col_idx]) @ V       (transpose: K^T @ V)
//
// It only validates, allocates (outputs and workspaces come from the caching allocator, on the current stream) and
// calls the extern "C" entry points of include/rlaopt_b200.h -- no arithmetic happens here.  Checks mirror
// _check_inputs (rlaopt/kernels/base.py:75-86) and fail with TORCH_CHECK -> RuntimeError like the reference's ops
// (rlaopt/csrc/cpp/input_checks.cpp:9-69).  The CPU key raises: the path has no CPU implementation.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/rlaopt_b200.h"

namespace rlaopt_b200_op {
namespace {

constexpr double kTcEpsD = 3.0e-7, kTcTol = 1.0e-5;  // accuracy guard of the tensor-core layout

double tc_sensitivity(int64_t kid) {
    switch (kid) {
        case RLAOPT_B200_KERNEL_MATERN32: return 1.5;
        case RLAOPT_B200_KERNEL_MATERN52: return 5.0 / 6.0;
        default: return 0.5;
    }
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, "rlaopt::kernel_matmat: ", what, " failed (code ", rc, "): ", rlaopt_b200_last_error());
}

at::Tensor bytes(size_t n, const at::Tensor& like) {
    return at::empty({(int64_t)(n > 0 ? n : 1)}, like.options().dtype(at::kByte));
}

struct Pack {
    at::Tensor buf;
    int64_t n;
};

template <typename T>
Pack pack(const at::Tensor& X, const c10::optional<at::Tensor>& idx, T inv_ls, const c10::optional<at::Tensor>& inv_vec,
          const c10::optional<at::Tensor>& center, int layout, void* stream) {
    const int64_t n_src = X.size(0), d = X.size(1);
    const int64_t n = idx.has_value() ? idx->numel() : n_src;
    Pack p{bytes(rlaopt_b200_packed_bytes(n, d, (int)sizeof(T), layout), X), n};
    const int64_t* ip = idx.has_value() ? idx->data_ptr<int64_t>() : nullptr;
    const T* vp = inv_vec.has_value() ? inv_vec->data_ptr<T>() : nullptr;
    const T* cp = center.has_value() ? center->data_ptr<T>() : nullptr;
    const int64_t ldx = n_src > 1 ? X.stride(0) : (d > 0 ? d : 1);
    int rc;
    if constexpr (sizeof(T) == 4)
        rc = rlaopt_b200_pack_points_f32(X.data_ptr<float>(), n, n_src, d, ldx, ip, inv_ls, vp, cp, layout, p.buf.data_ptr(), stream);
    else
        rc = rlaopt_b200_pack_points_f64(X.data_ptr<double>(), n, n_src, d, ldx, ip, inv_ls, vp, cp, layout, p.buf.data_ptr(), stream);
    check_rc(rc, "pack_points");
    return p;
}

template <typename T>
at::Tensor run(const at::Tensor& A1, const at::Tensor& A2, const at::Tensor& V2, int64_t kid, double lengthscale,
               const c10::optional<at::Tensor>& ls_vec, double scale, bool transpose,
               const c10::optional<at::Tensor>& row_idx, const c10::optional<at::Tensor>& col_idx) {
    void* stream = c10::cuda::getCurrentCUDAStream(A1.device().index()).stream();
    const int64_t d = A1.size(1), k = V2.size(1);
    c10::optional<at::Tensor> inv_vec;
    T inv_ls = (T)(1.0 / lengthscale);
    if (ls_vec.has_value()) {
        TORCH_CHECK(ls_vec->dim() == 1 && ls_vec->size(0) == d, "lengthscale tensor must have shape (", d, ",)");
        inv_vec = ls_vec->to(A1.device(), at::kDouble).reciprocal().to(A1.scalar_type()).contiguous();
        inv_ls = (T)1;
    }
    auto index = [&](const c10::optional<at::Tensor>& idx, int64_t n_src) -> c10::optional<at::Tensor> {
        if (!idx.has_value()) return c10::nullopt;
        TORCH_CHECK(idx->dim() == 1, "index tensor must be 1-D");
        at::Tensor t = idx->to(A1.device(), at::kLong).contiguous();
        if (t.numel() > 0)  // A1[blk] semantics: a device-side assert, like advanced indexing on CUDA tensors
            at::_assert_async(((t >= -n_src) & (t < n_src)).all(), "rlaopt::kernel_matmat: gather index out of bounds");
        return t;
    };
    const c10::optional<at::Tensor> ridx = index(row_idx, A1.size(0)), cidx = index(col_idx, A2.size(0));
    int layout = RLAOPT_B200_LAYOUT_SIMT;
    if (sizeof(T) == 4 && rlaopt_b200_layout_supported((int)kid, 4, d, k, RLAOPT_B200_LAYOUT_TC)) layout = RLAOPT_B200_LAYOUT_TC;
    Pack P1, P2;
    if (layout == RLAOPT_B200_LAYOUT_TC) {
        if constexpr (sizeof(T) == 4) {
            // both operands shifted by the column means of A2[col_idx] (K is a function of x - y)
            const int64_t m_eff = cidx.has_value() ? cidx->numel() : A2.size(0);
            at::Tensor center = at::zeros({d}, A2.options());
            if (m_eff > 0) {
                const size_t wsb = rlaopt_b200_column_mean_workspace_bytes(m_eff, d);
                at::Tensor ws = bytes(wsb, A2);
                check_rc(rlaopt_b200_column_mean_f32(A2.data_ptr<float>(), m_eff, A2.size(0), d,
                                                     A2.size(0) > 1 ? A2.stride(0) : d,
                                                     cidx.has_value() ? cidx->data_ptr<int64_t>() : nullptr,
                                                     center.data_ptr<float>(), ws.data_ptr(), wsb, stream),
                         "column_mean");
            }
            P1 = pack<T>(A1, ridx, inv_ls, inv_vec, center, layout, stream);
            P2 = pack<T>(A2, cidx, inv_ls, inv_vec, center, layout, stream);
            float r1 = 0.f, r2 = 0.f;  // accuracy guard: centred norms beyond the budget run on direct differences
            if (P1.n > 0) check_rc(rlaopt_b200_packed_stats_host(P1.buf.data_ptr(), layout, &r1, nullptr, stream), "packed_stats");
            if (P2.n > 0) check_rc(rlaopt_b200_packed_stats_host(P2.buf.data_ptr(), layout, &r2, nullptr, stream), "packed_stats");
            if (kTcEpsD * tc_sensitivity(kid) * ((double)r1 + (double)r2) > kTcTol) layout = RLAOPT_B200_LAYOUT_SIMT;
        }
    }
    if (layout == RLAOPT_B200_LAYOUT_SIMT) {
        P1 = pack<T>(A1, ridx, inv_ls, inv_vec, c10::nullopt, layout, stream);
        P2 = pack<T>(A2, cidx, inv_ls, inv_vec, c10::nullopt, layout, stream);
    }
    const Pack& rows = transpose ? P2 : P1;
    const Pack& cols = transpose ? P1 : P2;
    TORCH_CHECK(V2.size(0) == cols.n, "dimension mismatch: operator has ", cols.n, " columns, V has ", V2.size(0), " rows");
    at::Tensor Y = at::empty({rows.n, k}, V2.options());
    const size_t wsb = rlaopt_b200_matmat_workspace_bytes(rows.n, cols.n, d, k, (int)sizeof(T), layout);
    at::Tensor ws = bytes(wsb, V2);
    const int64_t ldv = V2.size(0) > 1 ? V2.stride(0) : k;
    int rc;
    if constexpr (sizeof(T) == 4)
        rc = rlaopt_b200_matmat_packed_f32(rows.buf.data_ptr(), rows.n, cols.buf.data_ptr(), cols.n, d, V2.data_ptr<float>(), k,
                                           ldv, Y.data_ptr<float>(), k, (int)kid, (float)scale, layout, ws.data_ptr(), wsb, stream);
    else
        rc = rlaopt_b200_matmat_packed_f64(rows.buf.data_ptr(), rows.n, cols.buf.data_ptr(), cols.n, d, V2.data_ptr<double>(), k,
                                           ldv, Y.data_ptr<double>(), k, (int)kid, scale, layout, ws.data_ptr(), wsb, stream);
    check_rc(rc, "matmat_packed");
    return Y;
}

at::Tensor kernel_matmat_cuda(const at::Tensor& A1, const at::Tensor& A2, const at::Tensor& V, int64_t kernel_id,
                              double lengthscale, const c10::optional<at::Tensor>& lengthscale_vec, double const_scaling,
                              bool transpose, const c10::optional<at::Tensor>& row_idx,
                              const c10::optional<at::Tensor>& col_idx) {
    TORCH_CHECK(A1.dim() == 2 && A2.dim() == 2, "A1 and A2 must be 2D tensors");
    TORCH_CHECK(A1.size(1) == A2.size(1), "A1 and A2 must have the same number of features, got ", A1.size(1), " and ", A2.size(1));
    TORCH_CHECK(A1.device() == A2.device() && A1.device() == V.device(), "A1, A2 and V must be on the same device.");
    TORCH_CHECK(A1.scalar_type() == A2.scalar_type() && A1.scalar_type() == V.scalar_type(), "A1, A2 and V must have the same dtype.");
    TORCH_CHECK(A1.scalar_type() == at::kFloat || A1.scalar_type() == at::kDouble, "dtype must be float32 or float64");
    TORCH_CHECK(V.dim() == 1 || V.dim() == 2, "x must be a 1D or 2D tensor. Received ", V.dim(), "D tensor.");
    TORCH_CHECK(kernel_id >= 0 && kernel_id <= 4, "unknown kernel id ", kernel_id);
    c10::cuda::CUDAGuard guard(A1.device());
    const bool vec = V.dim() == 1;
    at::Tensor A1c = A1.stride(1) == 1 || A1.size(1) <= 1 ? A1 : A1.contiguous();
    at::Tensor A2c = A2.stride(1) == 1 || A2.size(1) <= 1 ? A2 : A2.contiguous();
    at::Tensor V2 = (vec ? V.unsqueeze(1) : V);
    if ((V2.size(1) > 1 && V2.stride(1) != 1) || (V2.size(0) > 1 && V2.stride(0) < V2.size(1))) V2 = V2.contiguous();
    at::Tensor Y = A1.scalar_type() == at::kFloat
                       ? run<float>(A1c, A2c, V2, kernel_id, lengthscale, lengthscale_vec, const_scaling, transpose, row_idx, col_idx)
                       : run<double>(A1c, A2c, V2, kernel_id, lengthscale, lengthscale_vec, const_scaling, transpose, row_idx, col_idx);
    return vec ? Y.select(1, 0) : Y;
}

at::Tensor kernel_matmat_cpu(const at::Tensor&, const at::Tensor&, const at::Tensor&, int64_t, double,
                             const c10::optional<at::Tensor>&, double, bool, const c10::optional<at::Tensor>&,
                             const c10::optional<at::Tensor>&) {
    TORCH_CHECK(false, "rlaopt::kernel_matmat has no CPU implementation: the kernel-matmat path is CUDA-only (sm_100a)");
    return at::Tensor();
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(rlaopt, m) {
    m.def(
        "kernel_matmat(Tensor A1, Tensor A2, Tensor V, int kernel_id, float lengthscale, Tensor? lengthscale_vec, "
        "float const_scaling, bool transpose=False, Tensor? row_idx=None, Tensor? col_idx=None) -> Tensor");
}
TORCH_LIBRARY_IMPL(rlaopt, CUDA, m) { m.impl("kernel_matmat", &kernel_matmat_cuda); }
TORCH_LIBRARY_IMPL(rlaopt, CPU, m) { m.impl("kernel_matmat", &kernel_matmat_cpu); }

}  // namespace rlaopt_b200_op
