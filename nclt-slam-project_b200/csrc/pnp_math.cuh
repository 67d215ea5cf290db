// FP64 device math shared by K3 (5-point EPnP hypotheses, pnp_epnp_sm.cuh) and K5 (LM refinement, pnp.cu) - the GPU
// twin of OpenCV's numerics that cv2.solvePnPRansac runs under visual_landmark_matcher.py:342-346.
//
// 5-point EPnP is numerically chaotic: the 12x12 MtM has rank 10 and OpenCV's one-sided Jacobi
// SVD returns normalised rounding noise for the two null left vectors, which EPnP then uses
// (SURVEY.md App. A.4).  Reproducing cv2's hypotheses therefore needs the same operations in
// the same order: this file is compiled with -fmad=false, uses only IEEE +,-,*,/,sqrt in the
// solver, and follows the published algorithms step by step (Jacobi SVD of modules/core
// lapack.cpp incl. its private hypot, SVD back-substitution, EPnP of modules/calib3d epnp.cpp,
// Rodrigues).  cos/sin/acos appear only where a 1-ulp libm difference is harmless.
#pragma once
#include <cfloat>
#include <cstdint>

namespace pnpm {

constexpr int MAXN = 12;

__device__ __forceinline__ double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) {
        b /= a;
        return a * sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * sqrt(1 + a * a);
    }
    return 0;
}

// One-sided Jacobi on At (n rows of length m, row stride m). Rows become unit left singular
// vectors, W singular values (descending), Vt (n x n, may be null) right singular vectors.
__device__ void jacobi_svd(double* At, double* Wout, double* Vt, int m, int n) {
    const double eps = DBL_EPSILON * 10;
    const double minval = DBL_MIN;
    double W[MAXN];
    int i, j, k, iter;
    const int max_iter = m > 30 ? m : 30;
    double c, s, sd;

    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) {
            double t = At[i * m + k];
            sd += t * t;
        }
        W[i] = sd;
        if (Vt) {
            for (k = 0; k < n; k++) Vt[i * n + k] = 0;
            Vt[i * n + i] = 1;
        }
    }
    for (iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (i = 0; i < n - 1; i++)
            for (j = i + 1; j < n; j++) {
                double *Ai = At + i * m, *Aj = At + j * m;
                double a = W[i], p = 0, b = W[j];
                for (k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = cv_hypot(p, beta);
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (k = 0; k < m; k++) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0;
                    Aj[k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                if (Vt) {
                    double *Vi = Vt + i * n, *Vj = Vt + j * n;
                    for (k = 0; k < n; k++) {
                        double t0 = c * Vi[k] + s * Vj[k];
                        double t1 = -s * Vi[k] + c * Vj[k];
                        Vi[k] = t0;
                        Vj[k] = t1;
                    }
                }
            }
        if (!changed) break;
    }
    for (i = 0; i < n; i++) {
        for (k = 0, sd = 0; k < m; k++) {
            double t = At[i * m + k];
            sd += t * t;
        }
        W[i] = sqrt(sd);
    }
    for (i = 0; i < n - 1; i++) {
        j = i;
        for (k = i + 1; k < n; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            if (Vt) {
                for (k = 0; k < m; k++) { t = At[i * m + k]; At[i * m + k] = At[j * m + k]; At[j * m + k] = t; }
                for (k = 0; k < n; k++) { t = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = t; }
            }
        }
    }
    for (i = 0; i < n; i++) Wout[i] = W[i];
    if (!Vt) return;

    uint64_t rng = 0x12345678ull;
    for (i = 0; i < n; i++) {
        sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / m;
            for (k = 0; k < m; k++) {
                rng = (uint64_t)(uint32_t)rng * 4164903690ull + (uint32_t)(rng >> 32);
                uint32_t r = (uint32_t)rng;
                At[i * m + k] = (r & 256) != 0 ? val0 : -val0;
            }
            for (iter = 0; iter < 2; iter++)
                for (j = 0; j < i; j++) {
                    sd = 0;
                    for (k = 0; k < m; k++) sd += At[i * m + k] * At[j * m + k];
                    double asum = 0;
                    for (k = 0; k < m; k++) {
                        double t = At[i * m + k] - sd * At[j * m + k];
                        At[i * m + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (k = 0; k < m; k++) At[i * m + k] *= asum;
                }
            sd = 0;
            for (k = 0; k < m; k++) {
                double t = At[i * m + k];
                sd += t * t;
            }
            sd = sqrt(sd);
        }
        s = sd > minval ? 1 / sd : 0.;
        for (k = 0; k < m; k++) At[i * m + k] *= s;
    }
}

// cv::solve(A[m x n], b, DECOMP_SVD): At/Vt are caller scratch (n*m and n*n doubles)
__device__ void solve_svd(const double* A, int m, int n, const double* b, double* x, double* At, double* Vt) {
    double w[MAXN];
    int i, j;
    for (i = 0; i < m; i++)
        for (j = 0; j < n; j++) At[j * m + i] = A[i * n + j];
    jacobi_svd(At, w, Vt, m, n);
    double threshold = 0;
    for (i = 0; i < n; i++) x[i] = 0;
    for (i = 0; i < n; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (j = 0; j < m; j++) s += At[i * m + j] * b[j];
        s *= wi;
        for (j = 0; j < n; j++) x[j] = x[j] + s * Vt[i * n + j];
    }
}

// cv::Rodrigues vector -> matrix
__device__ void rodrigues_v2m(const double* r, double* R) {
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = theta ? 1. / theta : 0.;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = c * eye[k] + c1 * rrt[k] + s * r_x[k];
}

__device__ __forceinline__ double dot3(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ double dist2(const double* p1, const double* p2) {
    return (p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]) + (p1[2] - p2[2]) * (p1[2] - p2[2]);
}

// The 5-point EPnP itself (K3) lives in pnp_epnp_sm.cuh: same arithmetic, its arrays in shared memory.

}  // namespace pnpm
