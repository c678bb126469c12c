// extern "C" entry points declared in include/rlaopt_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rlaopt_b200.h"
#include "kmm_common.cuh"
#include "kmm_launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t err, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(err), cudaGetErrorString(err));
    return (int)err;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

bool valid_kernel(int kid) { return kid >= 0 && kid < kmm::KID_COUNT; }

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <typename T>
size_t packed_bytes_t(int64_t n, int64_t d, int layout) {
    if (n < 0 || d < 0) return 0;
    if (layout == RLAOPT_B200_LAYOUT_SIMT)
        return (size_t)kmm::round_up(n, kmm::PACK_ROWS) * (size_t)kmm::round_up(d, kmm::PACK_FEATS) * sizeof(T);
    if (layout == RLAOPT_B200_LAYOUT_TC && sizeof(T) == 4) return kmm::tc_packed_bytes(n, d);
    return 0;
}

template <typename T>
int pack_points_t(const T* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx, T inv_ls,
                  const T* inv_ls_vec, const T* center, int layout, void* packed, void* stream) {
    if (n < 0 || d <= 0 || ldx < d) return fail(RLAOPT_B200_EINVAL, "pack_points: bad shape n=%lld d=%lld ldx=%lld",
                                                (long long)n, (long long)d, (long long)ldx);
    if (n_src < 0 || (!idx && n > n_src))
        return fail(RLAOPT_B200_EINVAL, "pack_points: n=%lld points requested from n_src=%lld rows", (long long)n,
                    (long long)n_src);
    if (n == 0) return 0;
    if (!X || !packed) return fail(RLAOPT_B200_EINVAL, "pack_points: null pointer");
    cudaError_t err;
    if (layout == RLAOPT_B200_LAYOUT_SIMT) {
        // direct differences are exact under translation: the shift is not applied on this layout
        err = kmm::launch_pack<T>(X, n, n_src, d, ldx, idx, inv_ls, inv_ls_vec, static_cast<T*>(packed),
                                  static_cast<cudaStream_t>(stream));
    } else if (layout == RLAOPT_B200_LAYOUT_TC) {
        if constexpr (sizeof(T) == 4) {
            if (!kmm::tc_supported_d(d))
                return fail(RLAOPT_B200_EUNSUPPORTED, "pack_points: d=%lld not supported by the tensor-core layout",
                            (long long)d);
            err = kmm::launch_tc_pack(X, n, n_src, d, ldx, idx, inv_ls, inv_ls_vec, center, packed,
                                      static_cast<cudaStream_t>(stream));
        } else {
            return fail(RLAOPT_B200_EUNSUPPORTED, "pack_points: tensor-core layout is fp32 only");
        }
    } else {
        return fail(RLAOPT_B200_EINVAL, "pack_points: unknown layout %d", layout);
    }
    return err == cudaSuccess ? 0 : cuda_fail(err, "pack_points");
}

template <typename T>
size_t matmat_workspace_t(int64_t n, int64_t m, int64_t d, int64_t k, int layout) {
    if (n <= 0 || m <= 0 || k <= 0) return 0;
    if (layout == RLAOPT_B200_LAYOUT_SIMT) return kmm::simt_workspace_bytes<T>(n, m, k, sm_count());
    if (layout == RLAOPT_B200_LAYOUT_TC && sizeof(T) == 4) return kmm::tc_workspace_bytes(n, m, d, k, sm_count());
    return 0;
}

// workspace of the fused form: [main kernel workspace incl. the kept partial sums | reduction partials]
template <typename T>
size_t fused_main_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int layout) {
    if (layout == RLAOPT_B200_LAYOUT_SIMT) return align256(kmm::simt_workspace_bytes<T>(n, m, k, sm_count(), true));
    if (layout == RLAOPT_B200_LAYOUT_TC && sizeof(T) == 4)
        return align256(kmm::tc_workspace_bytes(n, m, d, k, sm_count(), true));
    return 0;
}

template <typename T>
size_t fused_workspace_t(int64_t n, int64_t m, int64_t d, int64_t k, int layout, int64_t gram_cols, int want_sqnorm) {
    if (n <= 0 || k <= 0) return 0;
    const size_t main_bytes = m > 0 ? fused_main_bytes<T>(n, m, d, k, layout) : 0;
    return main_bytes + align256(kmm::fuse_workspace_bytes<T>(n, k, gram_cols, want_sqnorm));
}

template <typename T>
int matmat_packed_t(const void* rows, int64_t n, const void* cols, int64_t m, int64_t d, const T* V, int64_t k,
                    int64_t ldv, T* Y, int64_t ldy, int kid, T scale, int layout, void* ws, size_t ws_bytes,
                    void* stream, const kmm::FuseArgs<T>* fuse = nullptr);

template <typename T, typename E>
int matmat_fused_t(const void* rows, int64_t n, const void* cols, int64_t m, int64_t d, const T* V, int64_t k,
                   int64_t ldv, T* Y, int64_t ldy, int kid, T scale, int layout, const E* e, void* ws, size_t ws_bytes,
                   void* stream) {
    if (!e) return fail(RLAOPT_B200_EINVAL, "matmat_fused: null epilogue");
    if ((e->addend && (e->ld_addend < k || e->addend_rows < 0)) || (e->rhs && (e->ld_rhs < k || e->rhs_rows < 0)))
        return fail(RLAOPT_B200_EINVAL, "matmat_fused: bad addend / rhs stride");
    if (!e->addend_idx && e->addend && e->addend_rows < n)
        return fail(RLAOPT_B200_EINVAL, "matmat_fused: addend has %lld rows, product has %lld", (long long)e->addend_rows,
                    (long long)n);
    if (!e->rhs_idx && e->rhs && e->rhs_rows < n)
        return fail(RLAOPT_B200_EINVAL, "matmat_fused: rhs has %lld rows, product has %lld", (long long)e->rhs_rows,
                    (long long)n);
    const bool want_gram = e->gram_lhs != nullptr && e->gram_cols > 0;
    if (want_gram && (!e->gram_out || e->ld_gram_lhs < e->gram_cols))
        return fail(RLAOPT_B200_EINVAL, "matmat_fused: Gram requested without output / with a bad stride");
    if ((want_gram || e->sqnorm_out) && !kmm::fuse_reductions_supported<T>(k, want_gram ? e->gram_cols : 0))
        return fail(RLAOPT_B200_EUNSUPPORTED, "matmat_fused: the fused reductions cover k <= 64 and gram_cols <= 64 (k=%lld)",
                    (long long)k);
    if (!Y && !want_gram && !e->sqnorm_out) return fail(RLAOPT_B200_EINVAL, "matmat_fused: nothing to compute");
    kmm::FuseArgs<T> f;
    f.alpha = e->alpha;
    f.beta = e->beta;
    f.addend = e->addend;
    f.ld_addend = e->ld_addend;
    f.addend_rows = e->addend_rows;
    f.addend_idx = e->addend_idx;
    f.gamma = e->gamma;
    f.rhs = e->rhs;
    f.ld_rhs = e->ld_rhs;
    f.rhs_rows = e->rhs_rows;
    f.rhs_idx = e->rhs_idx;
    f.Y = Y;
    f.ldy = ldy;
    f.gram_lhs = want_gram ? e->gram_lhs : nullptr;
    f.ld_gram_lhs = e->ld_gram_lhs;
    f.gram_cols = want_gram ? (int)e->gram_cols : 0;
    f.gram_out = e->gram_out;
    f.want_sqnorm = e->sqnorm_out != nullptr;
    f.sqnorm_out = e->sqnorm_out;
    return matmat_packed_t<T>(rows, n, cols, m, d, V, k, ldv, Y, ldy, kid, scale, layout, ws, ws_bytes, stream, &f);
}

template <typename T>
int matmat_packed_t(const void* rows, int64_t n, const void* cols, int64_t m, int64_t d, const T* V, int64_t k,
                    int64_t ldv, T* Y, int64_t ldy, int kid, T scale, int layout, void* ws, size_t ws_bytes,
                    void* stream, const kmm::FuseArgs<T>* fuse) {
    if (!valid_kernel(kid)) return fail(RLAOPT_B200_EINVAL, "matmat: unknown kernel id %d", kid);
    if (n < 0 || m < 0 || d <= 0 || k < 0 || ldv < k || (Y && ldy < k))
        return fail(RLAOPT_B200_EINVAL, "matmat: bad shape n=%lld m=%lld d=%lld k=%lld ldv=%lld ldy=%lld", (long long)n,
                    (long long)m, (long long)d, (long long)k, (long long)ldv, (long long)ldy);
    if (n == 0 || k == 0) return 0;
    if (!Y && !fuse) return fail(RLAOPT_B200_EINVAL, "matmat: null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t red_bytes = fuse ? align256(kmm::fuse_workspace_bytes<T>(n, k, fuse->gram_cols, fuse->want_sqnorm)) : 0;
    if (m == 0) {  // empty sum
        if (fuse) {  // the product is zero: only the addend / rhs terms remain (no partial sums: splits = 0)
            if (red_bytes > 0 && (!ws || ws_bytes < red_bytes))
                return fail(RLAOPT_B200_EWORKSPACE, "matmat_fused: workspace %zu < required %zu bytes", ws_bytes, red_bytes);
            cudaError_t err = kmm::launch_fuse<T>(*fuse, nullptr, 0, n, k, T(0), ws, st);
            return err == cudaSuccess ? 0 : cuda_fail(err, "matmat_fused(empty)");
        }
        cudaError_t err = cudaMemset2DAsync(Y, ldy * sizeof(T), 0, k * sizeof(T), n, st);
        return err == cudaSuccess ? 0 : cuda_fail(err, "matmat(memset)");
    }
    if (!rows || !cols || !V) return fail(RLAOPT_B200_EINVAL, "matmat: null pointer");
    const int sms = sm_count();
    if (sms <= 0) return fail(RLAOPT_B200_EINVAL, "matmat: no CUDA device");
    const bool keep = fuse != nullptr;
    const size_t main_bytes = keep ? fused_main_bytes<T>(n, m, d, k, layout) : 0;
    if (keep && (ws == nullptr || ws_bytes < main_bytes + red_bytes))
        return fail(RLAOPT_B200_EWORKSPACE, "matmat_fused: workspace %zu < required %zu bytes", ws_bytes,
                    main_bytes + red_bytes);
    int splits = 1;
    const T* part = static_cast<const T*>(ws);
    cudaError_t err;
    if (layout == RLAOPT_B200_LAYOUT_SIMT) {
        kmm::SimtArgs<T> a;
        a.Rt = static_cast<const T*>(rows);
        a.n = n;
        a.n_pad = kmm::round_up(n, kmm::PACK_ROWS);
        a.Ct = static_cast<const T*>(cols);
        a.m = m;
        a.m_pad = kmm::round_up(m, kmm::PACK_ROWS);
        a.d_pad = kmm::round_up(d, kmm::PACK_FEATS);
        a.V = V;
        a.ldv = ldv;
        a.k = k;
        a.Y = Y;
        a.ldy = ldy;
        a.scale = scale;
        a.kid = kid;
        a.stream = st;
        const size_t need = kmm::simt_workspace_bytes<T>(n, m, k, sms, keep);
        if (need > 0 && (ws == nullptr || ws_bytes < need))
            return fail(RLAOPT_B200_EWORKSPACE, "matmat: workspace %zu < required %zu bytes", ws_bytes, need);
        err = kmm::launch_simt<T>(a, sms, ws, ws_bytes, keep, &splits);
    } else if (layout == RLAOPT_B200_LAYOUT_TC) {
        if constexpr (sizeof(T) == 4) {
            if (kid == kmm::KID_LAPLACE)
                return fail(RLAOPT_B200_EUNSUPPORTED, "matmat: Laplace (L1) has no tensor-core path");
            if (!kmm::tc_supported_d(d))
                return fail(RLAOPT_B200_EUNSUPPORTED, "matmat: d=%lld not supported by the tensor-core path", (long long)d);
            const size_t need = kmm::tc_workspace_bytes(n, m, d, k, sms, keep);
            if (need > 0 && (ws == nullptr || ws_bytes < need))
                return fail(RLAOPT_B200_EWORKSPACE, "matmat: workspace %zu < required %zu bytes", ws_bytes, need);
            err = kmm::launch_tc(rows, n, cols, m, d, V, k, ldv, Y, ldy, kid, scale, sms, ws, ws_bytes, st, keep, &splits,
                                 &part);
        } else {
            return fail(RLAOPT_B200_EUNSUPPORTED, "matmat: tensor-core path is fp32 only");
        }
    } else {
        return fail(RLAOPT_B200_EINVAL, "matmat: unknown layout %d", layout);
    }
    if (err != cudaSuccess) return cuda_fail(err, "matmat");
    if (keep) {
        err = kmm::launch_fuse<T>(*fuse, part, splits, n, k, fuse->alpha * scale,
                                  static_cast<unsigned char*>(ws) + main_bytes, st);
        if (err != cudaSuccess) return cuda_fail(err, "matmat_fused(output stage)");
    }
    return 0;
}

// one-shot entry, tensor-core layout: d floats for the center + the column-mean partial sums
size_t center_bytes(int64_t m_cols, int64_t d, int layout) {
    if (layout != RLAOPT_B200_LAYOUT_TC) return 0;
    return align256((size_t)d * sizeof(float)) + align256(kmm::column_mean_workspace_bytes(m_cols, d));
}

template <typename T>
size_t oneshot_workspace_t(int64_t n_rows, int64_t m_cols, int64_t d, int64_t k, int layout) {
    return align256(packed_bytes_t<T>(n_rows, d, layout)) + align256(packed_bytes_t<T>(m_cols, d, layout)) +
           align256(matmat_workspace_t<T>(n_rows, m_cols, d, k, layout)) + center_bytes(m_cols, d, layout);
}

template <typename T>
int oneshot_t(const T* A1, int64_t n, int64_t lda1, const T* A2, int64_t m, int64_t lda2, int64_t d, const T* V,
              int64_t k, int64_t ldv, T* Y, int64_t ldy, int kid, T inv_ls, const T* inv_ls_vec, T scale,
              int transpose, const int64_t* row_idx, int64_t n_idx, const int64_t* col_idx, int64_t m_idx, int layout,
              void* ws, size_t ws_bytes, void* stream) {
    const int64_t n_eff = row_idx ? n_idx : n;
    const int64_t m_eff = col_idx ? m_idx : m;
    // rows of the product are the "row operand"; the transpose swaps the operands' roles
    const int64_t out_rows = transpose ? m_eff : n_eff;
    const int64_t red_cols = transpose ? n_eff : m_eff;
    const size_t pb1 = align256(packed_bytes_t<T>(n_eff, d, layout));
    const size_t pb2 = align256(packed_bytes_t<T>(m_eff, d, layout));
    const size_t mm = align256(matmat_workspace_t<T>(out_rows, red_cols, d, k, layout));
    const size_t cb = center_bytes(m_eff, d, layout);
    if (pb1 + pb2 + mm + cb > ws_bytes || (pb1 + pb2 + mm + cb > 0 && !ws))
        return fail(RLAOPT_B200_EWORKSPACE, "kernel_matmat: workspace %zu < required %zu bytes", ws_bytes,
                    pb1 + pb2 + mm + cb);
    if (n_eff == 0 || m_eff == 0)  // empty operand: nothing to pack; the product is an empty sum
        return matmat_packed_t<T>(nullptr, out_rows, nullptr, red_cols, d, V, k, ldv, Y, ldy, kid, scale, layout,
                                  nullptr, 0, stream);
    unsigned char* base = static_cast<unsigned char*>(ws);
    void* p1 = base;
    void* p2 = base + pb1;
    void* pw = base + pb1 + pb2;
    const T* center = nullptr;
    if constexpr (sizeof(T) == 4) {
        if (cb > 0) {  // tensor-core layout: both operands are shifted by the column means of A2[col_idx]
            float* c = reinterpret_cast<float*>(base + pb1 + pb2 + mm);
            cudaError_t err = kmm::launch_column_mean(A2, m_eff, m, d, lda2, col_idx, c,
                                                      base + pb1 + pb2 + mm + align256((size_t)d * sizeof(float)),
                                                      static_cast<cudaStream_t>(stream));
            if (err != cudaSuccess) return cuda_fail(err, "kernel_matmat(column_mean)");
            center = c;
        }
    }
    int rc = pack_points_t<T>(A1, n_eff, n, d, lda1, row_idx, inv_ls, inv_ls_vec, center, layout, p1, stream);
    if (rc) return rc;
    rc = pack_points_t<T>(A2, m_eff, m, d, lda2, col_idx, inv_ls, inv_ls_vec, center, layout, p2, stream);
    if (rc) return rc;
    if (transpose)
        return matmat_packed_t<T>(p2, m_eff, p1, n_eff, d, V, k, ldv, Y, ldy, kid, scale, layout, pw, mm, stream);
    return matmat_packed_t<T>(p1, n_eff, p2, m_eff, d, V, k, ldv, Y, ldy, kid, scale, layout, pw, mm, stream);
}

}  // namespace

extern "C" {

int rlaopt_b200_abi_version(void) { return RLAOPT_B200_ABI_VERSION; }
const char* rlaopt_b200_last_error(void) { return g_err; }
int rlaopt_b200_device_sm_count(void) { return sm_count(); }

int rlaopt_b200_layout_supported(int kernel_id, int elem_bytes, int64_t d, int64_t k, int layout) {
    if (!valid_kernel(kernel_id) || d <= 0 || k <= 0 || (elem_bytes != 4 && elem_bytes != 8)) return 0;
    if (layout == RLAOPT_B200_LAYOUT_SIMT) return 1;
    if (layout == RLAOPT_B200_LAYOUT_TC)
        return elem_bytes == 4 && kernel_id != kmm::KID_LAPLACE && kmm::tc_supported_d(d) && kmm::tc_supported_k(k);
    return 0;
}

size_t rlaopt_b200_packed_bytes(int64_t n, int64_t d, int elem_bytes, int layout) {
    return elem_bytes == 8 ? packed_bytes_t<double>(n, d, layout) : packed_bytes_t<float>(n, d, layout);
}

int rlaopt_b200_pack_points_f32(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                float inv_lengthscale, const float* inv_lengthscale_vec, const float* center,
                                int layout, void* packed, void* stream) {
    return pack_points_t<float>(X, n, n_src, d, ldx, idx, inv_lengthscale, inv_lengthscale_vec, center, layout, packed,
                                stream);
}
int rlaopt_b200_pack_points_f64(const double* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                double inv_lengthscale, const double* inv_lengthscale_vec, const double* center,
                                int layout, void* packed, void* stream) {
    return pack_points_t<double>(X, n, n_src, d, ldx, idx, inv_lengthscale, inv_lengthscale_vec, center, layout,
                                 packed, stream);
}

size_t rlaopt_b200_column_mean_workspace_bytes(int64_t n, int64_t d) { return kmm::column_mean_workspace_bytes(n, d); }

int rlaopt_b200_column_mean_f32(const float* X, int64_t n, int64_t n_src, int64_t d, int64_t ldx, const int64_t* idx,
                                float* center, void* workspace, size_t workspace_bytes, void* stream) {
    if (n <= 0 || d <= 0 || ldx < d || n_src < 0 || (!idx && n > n_src))
        return fail(RLAOPT_B200_EINVAL, "column_mean: bad shape n=%lld n_src=%lld d=%lld ldx=%lld", (long long)n,
                    (long long)n_src, (long long)d, (long long)ldx);
    if (!X || !center) return fail(RLAOPT_B200_EINVAL, "column_mean: null pointer");
    const size_t need = kmm::column_mean_workspace_bytes(n, d);
    if (!workspace || workspace_bytes < need)
        return fail(RLAOPT_B200_EWORKSPACE, "column_mean: workspace %zu < required %zu bytes", workspace_bytes, need);
    cudaError_t err = kmm::launch_column_mean(X, n, n_src, d, ldx, idx, center, workspace, static_cast<cudaStream_t>(stream));
    return err == cudaSuccess ? 0 : cuda_fail(err, "column_mean");
}

int rlaopt_b200_packed_stats_host(const void* packed, int layout, float* max_sqnorm_host, int64_t* bad_index_host,
                                  void* stream) {
    if (layout != RLAOPT_B200_LAYOUT_TC)
        return fail(RLAOPT_B200_EUNSUPPORTED, "packed_stats: only the tensor-core layout carries statistics");
    if (!packed) return fail(RLAOPT_B200_EINVAL, "packed_stats: null pointer");
    unsigned int words[2] = {0u, 0u};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t err = cudaMemcpyAsync(words, static_cast<const unsigned char*>(packed) + kmm::tc_stats_offset(),
                                      sizeof(words), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) return cuda_fail(err, "packed_stats");
    if (max_sqnorm_host) memcpy(max_sqnorm_host, &words[0], sizeof(float));
    if (bad_index_host) *bad_index_host = (int64_t)words[1];
    return 0;
}

size_t rlaopt_b200_matmat_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int elem_bytes, int layout) {
    return elem_bytes == 8 ? matmat_workspace_t<double>(n, m, d, k, layout) : matmat_workspace_t<float>(n, m, d, k, layout);
}

int rlaopt_b200_matmat_packed_f32(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                                  const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy, int kernel_id,
                                  float const_scaling, int layout, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    return matmat_packed_t<float>(rows_packed, n, cols_packed, m, d, V, k, ldv, Y, ldy, kernel_id, const_scaling,
                                  layout, workspace, workspace_bytes, stream);
}
int rlaopt_b200_matmat_packed_f64(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m, int64_t d,
                                  const double* V, int64_t k, int64_t ldv, double* Y, int64_t ldy, int kernel_id,
                                  double const_scaling, int layout, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    return matmat_packed_t<double>(rows_packed, n, cols_packed, m, d, V, k, ldv, Y, ldy, kernel_id, const_scaling,
                                   layout, workspace, workspace_bytes, stream);
}

size_t rlaopt_b200_matmat_fused_workspace_bytes(int64_t n, int64_t m, int64_t d, int64_t k, int elem_bytes, int layout,
                                                int64_t gram_cols, int want_sqnorm) {
    return elem_bytes == 8 ? fused_workspace_t<double>(n, m, d, k, layout, gram_cols, want_sqnorm)
                           : fused_workspace_t<float>(n, m, d, k, layout, gram_cols, want_sqnorm);
}

int rlaopt_b200_matmat_packed_fused_f32(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m,
                                        int64_t d, const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy,
                                        int kernel_id, float const_scaling, int layout,
                                        const rlaopt_b200_epilogue_f32* epilogue, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    return matmat_fused_t<float>(rows_packed, n, cols_packed, m, d, V, k, ldv, Y, ldy, kernel_id, const_scaling, layout,
                                 epilogue, workspace, workspace_bytes, stream);
}
int rlaopt_b200_matmat_packed_fused_f64(const void* rows_packed, int64_t n, const void* cols_packed, int64_t m,
                                        int64_t d, const double* V, int64_t k, int64_t ldv, double* Y, int64_t ldy,
                                        int kernel_id, double const_scaling, int layout,
                                        const rlaopt_b200_epilogue_f64* epilogue, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    return matmat_fused_t<double>(rows_packed, n, cols_packed, m, d, V, k, ldv, Y, ldy, kernel_id, const_scaling,
                                  layout, epilogue, workspace, workspace_bytes, stream);
}

size_t rlaopt_b200_kernel_matmat_workspace_bytes(int64_t n_rows, int64_t m_cols, int64_t d, int64_t k, int elem_bytes,
                                                 int layout) {
    // symmetric in (n_rows, m_cols) up to the split workspace; take the larger of both orientations
    const size_t a = elem_bytes == 8 ? oneshot_workspace_t<double>(n_rows, m_cols, d, k, layout)
                                     : oneshot_workspace_t<float>(n_rows, m_cols, d, k, layout);
    const size_t b = elem_bytes == 8 ? oneshot_workspace_t<double>(m_cols, n_rows, d, k, layout)
                                     : oneshot_workspace_t<float>(m_cols, n_rows, d, k, layout);
    return a > b ? a : b;
}

int rlaopt_b200_kernel_matmat_f32(const float* A1, int64_t n, int64_t lda1, const float* A2, int64_t m, int64_t lda2,
                                  int64_t d, const float* V, int64_t k, int64_t ldv, float* Y, int64_t ldy,
                                  int kernel_id, float inv_lengthscale, const float* inv_lengthscale_vec,
                                  float const_scaling, int transpose, const int64_t* row_idx, int64_t n_idx,
                                  const int64_t* col_idx, int64_t m_idx, int layout, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    return oneshot_t<float>(A1, n, lda1, A2, m, lda2, d, V, k, ldv, Y, ldy, kernel_id, inv_lengthscale,
                            inv_lengthscale_vec, const_scaling, transpose, row_idx, n_idx, col_idx, m_idx, layout,
                            workspace, workspace_bytes, stream);
}
int rlaopt_b200_kernel_matmat_f64(const double* A1, int64_t n, int64_t lda1, const double* A2, int64_t m,
                                  int64_t lda2, int64_t d, const double* V, int64_t k, int64_t ldv, double* Y,
                                  int64_t ldy, int kernel_id, double inv_lengthscale,
                                  const double* inv_lengthscale_vec, double const_scaling, int transpose,
                                  const int64_t* row_idx, int64_t n_idx, const int64_t* col_idx, int64_t m_idx,
                                  int layout, void* workspace, size_t workspace_bytes, void* stream) {
    return oneshot_t<double>(A1, n, lda1, A2, m, lda2, d, V, k, ldv, Y, ldy, kernel_id, inv_lengthscale,
                             inv_lengthscale_vec, const_scaling, transpose, row_idx, n_idx, col_idx, m_idx, layout,
                             workspace, workspace_bytes, stream);
}

int rlaopt_b200_kernel_matmat_host_f32(const float* A1_host, int64_t n, const float* A2_host, int64_t m, int64_t d,
                                       const float* V_host, int64_t k, float* Y_host, int kernel_id,
                                       float inv_lengthscale, float const_scaling, int transpose, int layout) {
    if (n <= 0 || m <= 0 || d <= 0 || k <= 0 || !A1_host || !A2_host || !V_host || !Y_host)
        return fail(RLAOPT_B200_EINVAL, "kernel_matmat_host: bad argument");
    const int64_t v_rows = transpose ? n : m, y_rows = transpose ? m : n;
    const size_t ws_bytes = rlaopt_b200_kernel_matmat_workspace_bytes(n, m, d, k, 4, layout);
    const size_t bA1 = align256((size_t)n * d * 4), bA2 = align256((size_t)m * d * 4);
    const size_t bV = align256((size_t)v_rows * k * 4), bY = align256((size_t)y_rows * k * 4);
    unsigned char* dev = nullptr;
    cudaError_t err = cudaMalloc(&dev, bA1 + bA2 + bV + bY + ws_bytes + 256);
    if (err != cudaSuccess) return cuda_fail(err, "kernel_matmat_host(cudaMalloc)");
    float* dA1 = reinterpret_cast<float*>(dev);
    float* dA2 = reinterpret_cast<float*>(dev + bA1);
    float* dV = reinterpret_cast<float*>(dev + bA1 + bA2);
    float* dY = reinterpret_cast<float*>(dev + bA1 + bA2 + bV);
    void* dW = dev + bA1 + bA2 + bV + bY;
    cudaStream_t st = nullptr;
    int rc = 0;
    if ((err = cudaMemcpyAsync(dA1, A1_host, (size_t)n * d * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (err = cudaMemcpyAsync(dA2, A2_host, (size_t)m * d * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (err = cudaMemcpyAsync(dV, V_host, (size_t)v_rows * k * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) {
        rc = cuda_fail(err, "kernel_matmat_host(H2D)");
    } else {
        rc = rlaopt_b200_kernel_matmat_f32(dA1, n, d, dA2, m, d, d, dV, k, k, dY, k, kernel_id, inv_lengthscale,
                                           nullptr, const_scaling, transpose, nullptr, 0, nullptr, 0, layout, dW,
                                           ws_bytes, st);
        if (rc == 0) {
            err = cudaMemcpyAsync(Y_host, dY, (size_t)y_rows * k * 4, cudaMemcpyDeviceToHost, st);
            if (err == cudaSuccess) err = cudaStreamSynchronize(st);
            if (err != cudaSuccess) rc = cuda_fail(err, "kernel_matmat_host(D2H)");
        }
    }
    cudaFree(dev);
    return rc;
}

}  // extern "C"
