// extern "C" entry points for PnP-RANSAC (include/nclt_b200.h).
#include "common.cuh"
#include "scratch.cuh"

static int check_pnp(nclt_ctx* c, const void* obj, const void* img, const void* n, int P, int Nmax,
                     const nclt_pnp_params* prm) {
    if (!c) return NCLT_ERR_ARG;
    if (P < 0 || Nmax <= 0 || !prm) return nclt_fail(c, NCLT_ERR_ARG, "pnp: bad P/Nmax/params");
    if (P > 0 && (!obj || !img || !n)) return nclt_fail(c, NCLT_ERR_ARG, "pnp: null input");
    if (prm->iterations < 1 || prm->iterations > 100000) return nclt_fail(c, NCLT_ERR_ARG, "pnp: bad iteration count");
    if (!(prm->fx != 0.0) || !(prm->fy != 0.0)) return nclt_fail(c, NCLT_ERR_ARG, "pnp: zero focal length");
    cudaSetDevice(c->device);
    return NCLT_OK;
}

static size_t pnp_buf_bytes(int P, int iters) {
    size_t h = (size_t)P * iters;
    return pad256(h * 5 * sizeof(int)) + pad256(h * 6 * sizeof(double)) + pad256(h * sizeof(int)) + pad256((size_t)P * 16);
}
static PnpBuffers carve_pnp(Carver& cv, int P, int iters) {
    size_t h = (size_t)P * iters;
    PnpBuffers b;
    b.sets = cv.take<int>(h * 5);
    b.models = cv.take<double>(h * 6);
    b.counts = cv.take<int>(h);
    b.state = cv.take<int>((size_t)P * 4);
    return b;
}

extern "C" int nclt_pnp_ransac_dev(nclt_ctx* c, const float* obj, const float* img, const int32_t* n, int P,
                                   int Nmax, const nclt_pnp_params* prm, uint8_t* out_ok, double* out_rvec,
                                   double* out_tvec, int32_t* out_n_inliers, uint8_t* out_mask,
                                   float* out_mean_err) {
    int rc = check_pnp(c, obj, img, n, P, Nmax, prm);
    if (rc) return rc;
    if (!out_ok || !out_rvec || !out_tvec || !out_n_inliers) return nclt_fail(c, NCLT_ERR_ARG, "pnp: null output");
    if (P == 0) return NCLT_OK;
    ScratchScope scope(c);
    if ((rc = nclt_scratch_reserve(c, pnp_buf_bytes(P, prm->iterations) + pad256((size_t)P * Nmax)))) return rc;
    Carver cv(c);
    PnpBuffers buf = carve_pnp(cv, P, prm->iterations);
    uint8_t* mask = out_mask ? out_mask : cv.take<uint8_t>((size_t)P * Nmax);
    return launch_pnp(c, obj, img, n, P, nullptr, Nmax, prm, buf, nullptr, out_ok, out_rvec, out_tvec, out_n_inliers, mask,
                      out_mean_err, nullptr, nullptr, false);
}

extern "C" int nclt_pnp_ransac(nclt_ctx* c, const float* obj, const float* img, const int32_t* n, int P, int Nmax,
                               const nclt_pnp_params* prm, uint8_t* out_ok, double* out_rvec, double* out_tvec,
                               int32_t* out_n_inliers, uint8_t* out_mask, float* out_mean_err, int32_t* out_sets,
                               double* out_models, int32_t* out_counts, int32_t* out_best_iter,
                               int32_t* out_niters) {
    int rc = check_pnp(c, obj, img, n, P, Nmax, prm);
    if (rc) return rc;
    if (!out_ok || !out_rvec || !out_tvec || !out_n_inliers) return nclt_fail(c, NCLT_ERR_ARG, "pnp: null output");
    if (P == 0) return NCLT_OK;
    const int iters = prm->iterations;
    ScratchScope scope(c);
    size_t pts = (size_t)P * Nmax;
    size_t need = pnp_buf_bytes(P, iters) + pad256(pts * 12) + pad256(pts * 8) + pad256((size_t)P * 4) + pad256(pts) +
                  pad256((size_t)P) + 2 * pad256((size_t)P * 24) + 4 * pad256((size_t)P * 4);
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    PnpBuffers buf = carve_pnp(cv, P, iters);
    float* d_obj = cv.take<float>(pts * 3);
    float* d_img = cv.take<float>(pts * 2);
    int* d_n = cv.take<int>(P);
    uint8_t* d_mask = cv.take<uint8_t>(pts);
    uint8_t* d_ok = cv.take<uint8_t>(P);
    double* d_r = cv.take<double>((size_t)P * 3);
    double* d_t = cv.take<double>((size_t)P * 3);
    int* d_ninl = cv.take<int>(P);
    float* d_err = cv.take<float>(P);
    int* d_best = cv.take<int>(P);
    int* d_nit = cv.take<int>(P);
    cudaStream_t s = c->stream;
    CU_TRY(c, cudaMemcpyAsync(d_obj, obj, pts * 12, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(d_img, img, pts * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(d_n, n, (size_t)P * 4, cudaMemcpyHostToDevice, s));
    if ((rc = launch_pnp(c, d_obj, d_img, d_n, P, nullptr, Nmax, prm, buf, nullptr, d_ok, d_r, d_t, d_ninl, d_mask, d_err,
                         d_best, d_nit, false)))
        return rc;
    size_t h = (size_t)P * iters;
    CU_TRY(c, cudaMemcpyAsync(out_ok, d_ok, P, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemcpyAsync(out_rvec, d_r, (size_t)P * 24, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemcpyAsync(out_tvec, d_t, (size_t)P * 24, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaMemcpyAsync(out_n_inliers, d_ninl, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    if (out_mask) CU_TRY(c, cudaMemcpyAsync(out_mask, d_mask, pts, cudaMemcpyDeviceToHost, s));
    if (out_mean_err) CU_TRY(c, cudaMemcpyAsync(out_mean_err, d_err, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    if (out_sets) CU_TRY(c, cudaMemcpyAsync(out_sets, buf.sets, h * 20, cudaMemcpyDeviceToHost, s));
    if (out_models) CU_TRY(c, cudaMemcpyAsync(out_models, buf.models, h * 48, cudaMemcpyDeviceToHost, s));
    if (out_counts) CU_TRY(c, cudaMemcpyAsync(out_counts, buf.counts, h * 4, cudaMemcpyDeviceToHost, s));
    if (out_best_iter) CU_TRY(c, cudaMemcpyAsync(out_best_iter, d_best, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    if (out_niters) CU_TRY(c, cudaMemcpyAsync(out_niters, d_nit, (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    return NCLT_OK;
}

extern "C" int nclt_pnp_score(nclt_ctx* c, const float* obj, const float* img, const int32_t* n, int P, int Nmax,
                              const nclt_pnp_params* prm, const double* models, int32_t* out_counts) {
    int rc = check_pnp(c, obj, img, n, P, Nmax, prm);
    if (rc) return rc;
    if (!models || !out_counts) return nclt_fail(c, NCLT_ERR_ARG, "pnp_score: null models/output");
    if (P == 0) return NCLT_OK;
    const int iters = prm->iterations;
    ScratchScope scope(c);
    size_t pts = (size_t)P * Nmax, h = (size_t)P * iters;
    if ((rc = nclt_scratch_reserve(c, pad256(pts * 12) + pad256(pts * 8) + pad256((size_t)P * 4) + pad256(h * 48) +
                                          pad256(h * 4))))
        return rc;
    Carver cv(c);
    float* d_obj = cv.take<float>(pts * 3);
    float* d_img = cv.take<float>(pts * 2);
    int* d_n = cv.take<int>(P);
    double* d_models = cv.take<double>(h * 6);
    PnpBuffers buf{nullptr, nullptr, cv.take<int>(h), nullptr};
    cudaStream_t s = c->stream;
    CU_TRY(c, cudaMemcpyAsync(d_obj, obj, pts * 12, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(d_img, img, pts * 8, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(d_n, n, (size_t)P * 4, cudaMemcpyHostToDevice, s));
    CU_TRY(c, cudaMemcpyAsync(d_models, models, h * 48, cudaMemcpyHostToDevice, s));
    if ((rc = launch_pnp(c, d_obj, d_img, d_n, P, nullptr, Nmax, prm, buf, d_models, nullptr, nullptr, nullptr, nullptr,
                         nullptr, nullptr, nullptr, nullptr, true)))
        return rc;
    CU_TRY(c, cudaMemcpyAsync(out_counts, buf.counts, h * 4, cudaMemcpyDeviceToHost, s));
    CU_TRY(c, cudaStreamSynchronize(s));
    return NCLT_OK;
}

extern "C" int nclt_project_points(nclt_ctx* c, const float* obj, int n, const double* rvec, const double* tvec,
                                   double fx, double fy, double cx, double cy, float* out) {
    if (!c) return NCLT_ERR_ARG;
    if (n < 0 || (n > 0 && (!obj || !out)) || !rvec || !tvec) return nclt_fail(c, NCLT_ERR_ARG, "project_points args");
    if (n == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    ScratchScope scope(c);
    int rc;
    if ((rc = nclt_scratch_reserve(c, pad256((size_t)n * 12) + pad256((size_t)n * 8)))) return rc;
    Carver cv(c);
    float* d_obj = cv.take<float>((size_t)n * 3);
    float* d_out = cv.take<float>((size_t)n * 2);
    CU_TRY(c, cudaMemcpyAsync(d_obj, obj, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_project_points(c, d_obj, n, rvec, tvec, fx, fy, cx, cy, d_out))) return rc;
    CU_TRY(c, cudaMemcpyAsync(out, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}
