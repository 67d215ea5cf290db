/*
 * TEST INFRASTRUCTURE (CPU oracle) - not product code.
 *
 * Restatement of cv2.solvePnPRansac(obj, img, K, dist=0, iterationsCount, reprojectionError,
 * confidence=0.99, flags=SOLVEPNP_ITERATIVE) as called at visual_landmark_matcher.py:342-346
 * and checkpoint_a_selftest.py:78-82.  OpenCV is not vendored in /root/reference; this follows
 * its published implementation (modules/calib3d/src/ptsetreg.cpp RANSACPointSetRegistrator,
 * solvepnp.cpp PnPRansacCallback, calibration.cpp projectPoints / findExtrinsicCameraParams2)
 * and SURVEY.md Appendix A.  Pinned against cv2 4.13.0 by tests/test_oracle_pnp.py: identical
 * ok/inlier sets and bit-identical rvec/tvec of the RANSAC model; refined pose to 1e-9.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

int orc_solvepnp_epnp(const float* obj, const float* img, int n, double fx, double fy, double cx, double cy,
                      double* rvec, double* tvec);
void orc_rodrigues_v2m(const double* r, double* R);
void orc_solve_svd(const double* A, int m, int n, const double* b, double* x);

/* cv::RNG: multiply-with-carry, state = (uint64)-1 at the start of every solvePnPRansac call */
static unsigned rng_next(uint64_t* st) {
    *st = (uint64_t)(unsigned)(*st) * 4164903690U + (unsigned)(*st >> 32);
    return (unsigned)(*st);
}

/* minimal sets: for each of `iters` iterations 5 distinct indices, duplicates re-drawn */
void orc_ransac_sets(int n, int iters, int* sets) {
    uint64_t st = (uint64_t)-1;
    for (int it = 0; it < iters; it++) {
        int* idx = sets + it * 5;
        for (int i = 0; i < 5; i++) {
            for (;;) {
                int v = (int)(rng_next(&st) % (unsigned)n);
                int dup = 0;
                for (int j = 0; j < i; j++) dup |= (idx[j] == v);
                idx[i] = v;
                if (!dup) break;
            }
        }
    }
}

/* PnPRansacCallback::computeError: projectPoints in double, rounded to float32, then the
 * squared distance in float32 */
void orc_reproj_err(const float* obj, const float* img, int n, const double* rvec, const double* tvec, double fx,
                    double fy, double cx, double cy, float* err, float* proj_out) {
    double R[9];
    orc_rodrigues_v2m(rvec, R);
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
        z = z ? 1. / z : 1;
        x *= z;
        y *= z;
        float pu = (float)(x * fx + cx);
        float pv = (float)(y * fy + cy);
        if (proj_out) {
            proj_out[2 * i] = pu;
            proj_out[2 * i + 1] = pv;
        }
        if (err) {
            float dx = img[2 * i] - pu, dy = img[2 * i + 1] - pv;
            float s = 0;
            s += dx * dx;
            s += dy * dy;
            err[i] = s;
        }
    }
}

static int cv_round(double v) { return (int)nearbyint(v); }   /* round-half-even */

int orc_ransac_update_niters(double p, double ep, int model_points, int max_iters) {
    p = p > 0. ? p : 0.;
    p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.;
    ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : cv_round(num / denom);
}

/* Levenberg-Marquardt refinement = solvePnP(ITERATIVE, useExtrinsicGuess=true): CvLevMarq(6,
 * 2n, 20 iterations, FLT_EPSILON) around projectPoints with analytic Jacobians. obj/img double. */
static void proj_jac(const double* obj, const double* img, int n, const double* param, double fx, double fy,
                     double cx, double cy, double* err /*2n*/, double* JtJ /*36 or NULL*/, double* JtErr /*6*/) {
    double R[9], dRdr[27];
    double rx = param[0], ry = param[1], rz = param[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    orc_rodrigues_v2m(param, R);
    if (JtJ) {
        memset(JtJ, 0, 36 * sizeof(double));
        memset(JtErr, 0, 6 * sizeof(double));
        if (theta < DBL_EPSILON) {
            memset(dRdr, 0, sizeof(dRdr));
            dRdr[5] = dRdr[15] = dRdr[19] = -1;
            dRdr[7] = dRdr[11] = dRdr[21] = 1;
        } else {
            double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
            double ux = rx * itheta, uy = ry * itheta, uz = rz * itheta;
            double rrt[9] = {ux * ux, ux * uy, ux * uz, ux * uy, uy * uy, uy * uz, ux * uz, uy * uz, uz * uz};
            double r_x[9] = {0, -uz, uy, uz, 0, -ux, -uy, ux, 0};
            static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            double drrt[27] = {ux + ux, uy, uz, uy, 0, 0, uz, 0, 0, 0, ux, 0, ux, uy + uy, uz, 0, uz, 0,
                               0, 0, ux, 0, 0, uy, ux, uy, uz + uz};
            static const double d_r_x_[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0,
                                              0, -1, 0, 1, 0, 0, 0, 0, 0};
            for (int i = 0; i < 3; i++) {
                double ri = i == 0 ? ux : i == 1 ? uy : uz;
                double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
                double a3 = (c - s * itheta) * ri, a4 = s * itheta;
                for (int k = 0; k < 9; k++)
                    dRdr[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x_[i * 9 + k];
            }
        }
    }
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + param[3];
        double y = R[3] * X + R[4] * Y + R[5] * Z + param[4];
        double z = R[6] * X + R[7] * Y + R[8] * Z + param[5];
        double iz = z ? 1. / z : 1;
        double xn = x * iz, yn = y * iz;
        double eu = xn * fx + cx - img[2 * i];
        double ev = yn * fy + cy - img[2 * i + 1];
        err[2 * i] = eu;
        err[2 * i + 1] = ev;
        if (JtJ) {
            double Ju[6], Jv[6];
            /* d/dt */
            Ju[3] = fx * iz; Ju[4] = 0; Ju[5] = -fx * xn * iz;
            Jv[3] = 0; Jv[4] = fy * iz; Jv[5] = -fy * yn * iz;
            for (int k = 0; k < 3; k++) {
                const double* D = dRdr + 9 * k;
                double dx = D[0] * X + D[1] * Y + D[2] * Z;
                double dy = D[3] * X + D[4] * Y + D[5] * Z;
                double dz = D[6] * X + D[7] * Y + D[8] * Z;
                Ju[k] = fx * (dx * iz - xn * iz * dz);
                Jv[k] = fy * (dy * iz - yn * iz * dz);
            }
            for (int a = 0; a < 6; a++) {
                JtErr[a] += Ju[a] * eu + Jv[a] * ev;
                for (int b = 0; b < 6; b++) JtJ[a * 6 + b] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
            }
        }
    }
}

static double norm2(const double* v, int n) {
    double s = 0;
    for (int i = 0; i < n; i++) s += v[i] * v[i];
    return sqrt(s);
}

void orc_lm_refine(const double* obj, const double* img, int n, double fx, double fy, double cx, double cy,
                   double* rvec, double* tvec, double* work /* 2n doubles */) {
    double param[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]}, prev[6];
    double JtJ[36], JtErr[6], A[36], delta[6];
    int lambdaLg10 = -3, iters = 0;
    const int max_iter = 20;
    const double epsilon = FLT_EPSILON, LOG10 = log(10.);
    double prevErrNorm = DBL_MAX, errNorm;
    for (;;) {
        /* CALC_J */
        proj_jac(obj, img, n, param, fx, fy, cx, cy, work, JtJ, JtErr);
        if (iters == 0) prevErrNorm = norm2(work, 2 * n);
        memcpy(prev, param, sizeof(prev));
        for (;;) {
            /* step(): (JtJ with diagonal * (1+lambda)) x = JtErr ; param = prev - x */
            double lambda = exp(lambdaLg10 * LOG10);
            memcpy(A, JtJ, sizeof(A));
            for (int i = 0; i < 6; i++) A[i * 6 + i] *= 1. + lambda;
            orc_solve_svd(A, 6, 6, JtErr, delta);
            for (int i = 0; i < 6; i++) param[i] = prev[i] - delta[i];
            /* CHECK_ERR */
            proj_jac(obj, img, n, param, fx, fy, cx, cy, work, NULL, NULL);
            errNorm = norm2(work, 2 * n);
            if (errNorm > prevErrNorm && ++lambdaLg10 <= 16) continue;
            break;
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double d[6];
        for (int i = 0; i < 6; i++) d[i] = param[i] - prev[i];
        if (++iters >= max_iter || norm2(d, 6) / norm2(prev, 6) < epsilon) break;
        prevErrNorm = errNorm;
    }
    rvec[0] = param[0]; rvec[1] = param[1]; rvec[2] = param[2];
    tvec[0] = param[3]; tvec[1] = param[4]; tvec[2] = param[5];
}

/*
 * Full solvePnPRansac.  Debug outputs (any may be NULL): sets[max_iters*5], counts[max_iters]
 * (-1 for iterations never reached), models[max_iters*6] (rvec,tvec per iteration).
 * Returns 1 ok / 0 not ok / -1 unsupported (n < 5; the reference never gets there because of
 * MIN_MATCHES=10, visual_landmark_matcher.py:330).
 */
int orc_pnp_ransac(const float* obj, const float* img, int n, double fx, double fy, double cx, double cy,
                   int max_iters, float reproj_thr, double confidence, int refine, int* sets, int* counts,
                   double* models, int* best_iter_out, int* niters_out, unsigned char* mask, double* rvec,
                   double* tvec, float* err_buf /* n */, unsigned char* mask_buf /* n */,
                   double* work /* 7n doubles */) {
    const int MP = 5;
    if (n < MP) return -1;
    int niters = max_iters > 1 ? max_iters : 1, max_good = 0, best_iter = -1;
    double best_r[3] = {0, 0, 0}, best_t[3] = {0, 0, 0};
    const float t2 = reproj_thr * reproj_thr;
    if (counts) for (int i = 0; i < max_iters; i++) counts[i] = -1;
    memset(mask, 0, n);
    if (n == MP) {
        orc_solvepnp_epnp(obj, img, n, fx, fy, cx, cy, best_r, best_t);
        memset(mask, 1, n);
        max_good = n;
        best_iter = 0;
        refine = 0;   /* npoints == model_points: solvePnPRansac returns the kernel result as is */
    } else {
        uint64_t st = (uint64_t)-1;
        for (int it = 0; it < niters; it++) {
            int idx[5];
            float so[15], si[10];
            for (int i = 0; i < MP; i++) {
                for (;;) {
                    int v = (int)(rng_next(&st) % (unsigned)n);
                    int dup = 0;
                    for (int j = 0; j < i; j++) dup |= (idx[j] == v);
                    idx[i] = v;
                    if (!dup) break;
                }
                memcpy(so + 3 * i, obj + 3 * idx[i], 12);
                memcpy(si + 2 * i, img + 2 * idx[i], 8);
            }
            if (sets) memcpy(sets + it * 5, idx, sizeof(idx));
            double r[3], t[3];
            orc_solvepnp_epnp(so, si, MP, fx, fy, cx, cy, r, t);
            if (models) { memcpy(models + it * 6, r, 24); memcpy(models + it * 6 + 3, t, 24); }
            orc_reproj_err(obj, img, n, r, t, fx, fy, cx, cy, err_buf, NULL);
            int good = 0;
            for (int i = 0; i < n; i++) {
                mask_buf[i] = err_buf[i] <= t2;
                good += mask_buf[i];
            }
            if (counts) counts[it] = good;
            if (good > (max_good > MP - 1 ? max_good : MP - 1)) {
                memcpy(mask, mask_buf, n);
                memcpy(best_r, r, 24);
                memcpy(best_t, t, 24);
                max_good = good;
                best_iter = it;
                niters = orc_ransac_update_niters(confidence, (double)(n - good) / n, MP, niters);
            }
        }
    }
    if (best_iter_out) *best_iter_out = best_iter;
    if (niters_out) *niters_out = niters;
    memcpy(rvec, best_r, 24);
    memcpy(tvec, best_t, 24);
    if (max_good <= 0) return 0;
    if (refine) {
        double* o = work;
        double* im = work + 3 * n;
        int m = 0;
        for (int i = 0; i < n; i++)
            if (mask[i]) {
                o[3 * m] = obj[3 * i]; o[3 * m + 1] = obj[3 * i + 1]; o[3 * m + 2] = obj[3 * i + 2];
                im[2 * m] = img[2 * i]; im[2 * m + 1] = img[2 * i + 1];
                m++;
            }
        orc_lm_refine(o, im, m, fx, fy, cx, cy, rvec, tvec, work + 5 * n);
    }
    return 1;
}
