// FP64 device math for K3 (5-point EPnP hypotheses) - the GPU twin of OpenCV's numerics that
// cv2.solvePnPRansac runs under visual_landmark_matcher.py:342-346.
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

// 3x3 SVD in cv::SVD::compute layout: w[3], u[9] (row-major, columns = left vectors), vt[9]
__device__ void svd3(const double* A, double* w, double* u, double* vt) {
    double At[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) At[j * 3 + i] = A[i * 3 + j];
    jacobi_svd(At, w, vt, 3, 3);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) u[i * 3 + j] = At[j * 3 + i];
}

__device__ void invert3_svd(const double* A, double* Ainv) {
    double w[3], u[9], vt[9];
    svd3(A, w, u, vt);
    double threshold = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
    for (int i = 0; i < 9; i++) Ainv[i] = 0;
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buffer[3];
        for (int j = 0; j < 3; j++) buffer[j] = u[j * 3 + i] * wi;
        for (int k = 0; k < 3; k++) {
            double s = vt[i * 3 + k];
            for (int j = 0; j < 3; j++) Ainv[k * 3 + j] = Ainv[k * 3 + j] + s * buffer[j];
        }
    }
}

// cv::Rodrigues matrix -> vector
__device__ void rodrigues_m2v(const double* Rin, double* rvec) {
    double w[3], U[9], Vt[9], R[9];
    svd3(Rin, w, U, Vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += U[i * 3 + k] * Vt[k * 3 + j];
            R[i * 3 + j] = s;
        }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        double t;
        if (c > 0)
            rx = ry = rz = 0;
        else {
            t = (R[0] + 1) * 0.5;
            rx = sqrt(t > 0. ? t : 0.);
            t = (R[4] + 1) * 0.5;
            ry = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5;
            rz = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        rx *= vth; ry *= vth; rz *= vth;
    }
    rvec[0] = rx; rvec[1] = ry; rvec[2] = rz;
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

// Householder QR least squares of EPnP's Gauss-Newton step (6x4), incl. its row-range quirk.
__device__ void qr_solve_6x4(double* pA, double* pb, double* pX) {
    const int nr = 6, nc = 4;
    double A1[4], A2[4];
    double* ppAkk = pA;
    for (int k = 0; k < nc; k++) {
        double *ppAik1 = ppAkk, eta = fabs(*ppAik1);
        for (int i = k + 1; i < nr; i++) {
            double elt = fabs(*ppAik1);
            if (eta < elt) eta = elt;
            ppAik1 += nc;
        }
        if (eta == 0) {
            A1[k] = A2[k] = 0.0;
            return;
        } else {
            double *ppAik2 = ppAkk, sum2 = 0.0, inv_eta = 1. / eta;
            for (int i = k; i < nr; i++) {
                *ppAik2 *= inv_eta;
                sum2 += *ppAik2 * *ppAik2;
                ppAik2 += nc;
            }
            double sigma = sqrt(sum2);
            if (*ppAkk < 0) sigma = -sigma;
            *ppAkk += sigma;
            A1[k] = sigma * *ppAkk;
            A2[k] = -eta * sigma;
            for (int j = k + 1; j < nc; j++) {
                double *ppAik = ppAkk, sum = 0;
                for (int i = k; i < nr; i++) {
                    sum += *ppAik * ppAik[j - k];
                    ppAik += nc;
                }
                double tau = sum / A1[k];
                ppAik = ppAkk;
                for (int i = k; i < nr; i++) {
                    ppAik[j - k] -= tau * *ppAik;
                    ppAik += nc;
                }
            }
        }
        ppAkk += nc + 1;
    }
    double* ppAjj = pA;
    for (int j = 0; j < nc; j++) {
        double *ppAij = ppAjj, tau = 0;
        for (int i = j; i < nr; i++) {
            tau += *ppAij * pb[i];
            ppAij += nc;
        }
        tau /= A1[j];
        ppAij = ppAjj;
        for (int i = j; i < nr; i++) {
            pb[i] -= tau * *ppAij;
            ppAij += nc;
        }
        ppAjj += nc + 1;
    }
    pX[nc - 1] = pb[nc - 1] / A2[nc - 1];
    for (int i = nc - 2; i >= 0; i--) {
        double *ppAij = pA + i * nc + (i + 1), sum = 0;
        for (int j = i + 1; j < nc; j++) {
            sum += *ppAij * pX[j];
            ppAij++;
        }
        pX[i] = (pb[i] - sum) / A2[i];
    }
}

__device__ void gauss_newton(const double* L, const double* rho, double* betas) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int k = 0; k < 5; k++) {
        for (int i = 0; i < 6; i++) {
            const double* rowL = L + i * 10;
            double* rowA = a + i * 4;
            rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
            rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
            rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
            rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
            b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] +
                             rowL[2] * betas[1] * betas[1] + rowL[3] * betas[0] * betas[2] +
                             rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                             rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] +
                             rowL[8] * betas[2] * betas[3] + rowL[9] * betas[3] * betas[3]);
        }
        qr_solve_6x4(a, b, x);
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

// EPnP state for the 5-point minimal problem
struct Epnp5 {
    double pws[15], us[10], alphas[20], pcs[15];
    double cws[4][3], ccs[4][3];
    double fu, fv, uc, vc;
};

__device__ double compute_R_and_t(Epnp5& e, const double* ut, const double* betas, double* R /*9*/, double* t) {
    const int n = 5;
    int i, j, k;
    for (i = 0; i < 4; i++) e.ccs[i][0] = e.ccs[i][1] = e.ccs[i][2] = 0.0;
    for (i = 0; i < 4; i++) {
        const double* v = ut + 12 * (11 - i);
        for (j = 0; j < 4; j++)
            for (k = 0; k < 3; k++) e.ccs[j][k] += betas[i] * v[3 * j + k];
    }
    for (i = 0; i < n; i++) {
        const double* a = e.alphas + 4 * i;
        double* pc = e.pcs + 3 * i;
        for (j = 0; j < 3; j++)
            pc[j] = a[0] * e.ccs[0][j] + a[1] * e.ccs[1][j] + a[2] * e.ccs[2][j] + a[3] * e.ccs[3][j];
    }
    if (e.pcs[2] < 0.0) {
        for (i = 0; i < 4; i++)
            for (j = 0; j < 3; j++) e.ccs[i][j] = -e.ccs[i][j];
        for (i = 0; i < n; i++) {
            e.pcs[3 * i] = -e.pcs[3 * i];
            e.pcs[3 * i + 1] = -e.pcs[3 * i + 1];
            e.pcs[3 * i + 2] = -e.pcs[3 * i + 2];
        }
    }
    // estimate_R_and_t
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (i = 0; i < n; i++)
        for (j = 0; j < 3; j++) {
            pc0[j] += e.pcs[3 * i + j];
            pw0[j] += e.pws[3 * i + j];
        }
    for (j = 0; j < 3; j++) {
        pc0[j] /= n;
        pw0[j] /= n;
    }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, abt_d[3], abt_u[9], abt_vt[9], abt_v[9];
    for (i = 0; i < n; i++) {
        const double* pc = e.pcs + 3 * i;
        const double* pw = e.pws + 3 * i;
        for (j = 0; j < 3; j++) {
            abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    svd3(abt, abt_d, abt_u, abt_vt);
    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) abt_v[i * 3 + j] = abt_vt[j * 3 + i];
    for (i = 0; i < 3; i++)
        for (j = 0; j < 3; j++) R[i * 3 + j] = dot3(abt_u + 3 * i, abt_v + 3 * j);
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] -
                       R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
    if (det < 0) {
        R[6] = -R[6];
        R[7] = -R[7];
        R[8] = -R[8];
    }
    t[0] = pc0[0] - dot3(R, pw0);
    t[1] = pc0[1] - dot3(R + 3, pw0);
    t[2] = pc0[2] - dot3(R + 6, pw0);
    // reprojection error
    double sum2 = 0.0;
    for (i = 0; i < n; i++) {
        const double* pw = e.pws + 3 * i;
        double Xc = dot3(R, pw) + t[0];
        double Yc = dot3(R + 3, pw) + t[1];
        double inv_Zc = 1.0 / (dot3(R + 6, pw) + t[2]);
        double ue = e.uc + e.fu * Xc * inv_Zc;
        double ve = e.vc + e.fv * Yc * inv_Zc;
        double u = e.us[2 * i], v = e.us[2 * i + 1];
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

// cv2.solvePnP(5 points f32, K, dist=0, flags=EPNP) -> rvec[3], tvec[3]
// scratch: 2*144 + 60 doubles of thread-private memory supplied by the caller
__device__ void solvepnp_epnp5(const float* obj /*15*/, const float* img /*10*/, double fx, double fy, double cx,
                               double cy, double* rvec, double* tvec, double* scratch) {
    const int n = 5;
    Epnp5 e;
    e.fu = fx; e.fv = fy; e.uc = cx; e.vc = cy;
    const double ifx = 1. / fx, ify = 1. / fy;
    int i, j, k;
    for (i = 0; i < n; i++) {
        e.pws[3 * i] = obj[3 * i];
        e.pws[3 * i + 1] = obj[3 * i + 1];
        e.pws[3 * i + 2] = obj[3 * i + 2];
        // undistortPoints writes normalised coordinates as float32; epnp maps them back in double
        double x = ((double)img[2 * i] - cx) * ifx;
        double y = ((double)img[2 * i + 1] - cy) * ify;
        e.us[2 * i] = (double)(float)x * fx + cx;
        e.us[2 * i + 1] = (double)(float)y * fy + cy;
    }
    // choose_control_points
    e.cws[0][0] = e.cws[0][1] = e.cws[0][2] = 0;
    for (i = 0; i < n; i++)
        for (j = 0; j < 3; j++) e.cws[0][j] += e.pws[3 * i + j];
    for (j = 0; j < 3; j++) e.cws[0][j] /= n;
    {
        double pw0[15], pw0tpw0[9], dc[3], u[9], vt[9];
        for (i = 0; i < n; i++)
            for (j = 0; j < 3; j++) pw0[3 * i + j] = e.pws[3 * i + j] - e.cws[0][j];
        for (i = 0; i < 3; i++)
            for (j = i; j < 3; j++) {
                double s = 0;
                for (k = 0; k < n; k++) s += pw0[3 * k + i] * pw0[3 * k + j];
                pw0tpw0[i * 3 + j] = s;
                pw0tpw0[j * 3 + i] = s;
            }
        svd3(pw0tpw0, dc, u, vt);
        for (i = 1; i < 4; i++) {
            double kk = sqrt(dc[i - 1] / n);
            // row (i-1) of U^T = column (i-1) of U
            for (j = 0; j < 3; j++) e.cws[i][j] = e.cws[0][j] + kk * u[j * 3 + (i - 1)];
        }
    }
    // compute_barycentric_coordinates
    {
        double cc[9], ci[9];
        for (i = 0; i < 3; i++)
            for (j = 1; j < 4; j++) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
        invert3_svd(cc, ci);
        for (i = 0; i < n; i++) {
            const double* pi = e.pws + 3 * i;
            double* a = e.alphas + 4 * i;
            for (j = 0; j < 3; j++)
                a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) +
                           ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
            a[0] = 1.0 - a[1] - a[2] - a[3];
        }
    }
    double* M = scratch;          // 10 x 12
    double* At = scratch + 120;   // 12 x 12 : MtM, then U^T rows after the SVD
    double* tmp = scratch + 264;  // solve scratch (At 6x5=30 + Vt 25)
    for (i = 0; i < n; i++) {
        const double* as = e.alphas + 4 * i;
        double u = e.us[2 * i], v = e.us[2 * i + 1];
        double* M1 = M + (2 * i) * 12;
        double* M2 = M1 + 12;
        for (j = 0; j < 4; j++) {
            M1[3 * j] = as[j] * e.fu;
            M1[3 * j + 1] = 0.0;
            M1[3 * j + 2] = as[j] * (e.uc - u);
            M2[3 * j] = 0.0;
            M2[3 * j + 1] = as[j] * e.fv;
            M2[3 * j + 2] = as[j] * (e.vc - v);
        }
    }
    // MtM (symmetric) -> its transpose is itself: At = MtM, rows become U^T after the SVD
    for (i = 0; i < 12; i++)
        for (j = i; j < 12; j++) {
            double s = 0;
            for (k = 0; k < 2 * n; k++) s += M[k * 12 + i] * M[k * 12 + j];
            At[i * 12 + j] = s;
            At[j * 12 + i] = s;
        }
    double d[12];
    // Vt is needed: the normalisation of the left vectors only happens when V is requested
    double* Vt12 = scratch;   // M is dead now (120 < 144: extend into tmp? no - use separate area)
    Vt12 = scratch + 320;     // 144 doubles
    jacobi_svd(At, d, Vt12, 12, 12);
    const double* ut = At;

    double l_6x10[60], rho[6];
    {
        const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
        double dv[4][6][3];
        for (i = 0; i < 4; i++) {
            int a = 0, b = 1;
            for (j = 0; j < 6; j++) {
                dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
                dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
                dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
                b++;
                if (b > 3) { a++; b = a + 1; }
            }
        }
        for (i = 0; i < 6; i++) {
            double* row = l_6x10 + 10 * i;
            row[0] = dot3(dv[0][i], dv[0][i]);
            row[1] = 2.0 * dot3(dv[0][i], dv[1][i]);
            row[2] = dot3(dv[1][i], dv[1][i]);
            row[3] = 2.0 * dot3(dv[0][i], dv[2][i]);
            row[4] = 2.0 * dot3(dv[1][i], dv[2][i]);
            row[5] = dot3(dv[2][i], dv[2][i]);
            row[6] = 2.0 * dot3(dv[0][i], dv[3][i]);
            row[7] = 2.0 * dot3(dv[1][i], dv[3][i]);
            row[8] = 2.0 * dot3(dv[2][i], dv[3][i]);
            row[9] = dot3(dv[3][i], dv[3][i]);
        }
    }
    rho[0] = dist2(e.cws[0], e.cws[1]);
    rho[1] = dist2(e.cws[0], e.cws[2]);
    rho[2] = dist2(e.cws[0], e.cws[3]);
    rho[3] = dist2(e.cws[1], e.cws[2]);
    rho[4] = dist2(e.cws[1], e.cws[3]);
    rho[5] = dist2(e.cws[2], e.cws[3]);

    double Betas[4][4], rep_errors[4], Rs[4][9], ts[4][3];
    for (i = 0; i < 4; i++)
        for (j = 0; j < 4; j++) Betas[i][j] = 0;
    // N = 1
    {
        double l[24], b4[4];
        for (i = 0; i < 6; i++) {
            l[i * 4 + 0] = l_6x10[i * 10 + 0]; l[i * 4 + 1] = l_6x10[i * 10 + 1];
            l[i * 4 + 2] = l_6x10[i * 10 + 3]; l[i * 4 + 3] = l_6x10[i * 10 + 6];
        }
        solve_svd(l, 6, 4, rho, b4, tmp, tmp + 30);
        double* betas = Betas[1];
        if (b4[0] < 0) {
            betas[0] = sqrt(-b4[0]);
            betas[1] = -b4[1] / betas[0];
            betas[2] = -b4[2] / betas[0];
            betas[3] = -b4[3] / betas[0];
        } else {
            betas[0] = sqrt(b4[0]);
            betas[1] = b4[1] / betas[0];
            betas[2] = b4[2] / betas[0];
            betas[3] = b4[3] / betas[0];
        }
        gauss_newton(l_6x10, rho, betas);
        rep_errors[1] = compute_R_and_t(e, ut, betas, Rs[1], ts[1]);
    }
    // N = 2
    {
        double l[18], b3[3];
        for (i = 0; i < 6; i++) {
            l[i * 3 + 0] = l_6x10[i * 10 + 0]; l[i * 3 + 1] = l_6x10[i * 10 + 1]; l[i * 3 + 2] = l_6x10[i * 10 + 2];
        }
        solve_svd(l, 6, 3, rho, b3, tmp, tmp + 30);
        double* betas = Betas[2];
        if (b3[0] < 0) {
            betas[0] = sqrt(-b3[0]);
            betas[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0;
        } else {
            betas[0] = sqrt(b3[0]);
            betas[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0;
        }
        if (b3[1] < 0) betas[0] = -betas[0];
        betas[2] = 0.0;
        betas[3] = 0.0;
        gauss_newton(l_6x10, rho, betas);
        rep_errors[2] = compute_R_and_t(e, ut, betas, Rs[2], ts[2]);
    }
    // N = 3
    {
        double l[30], b5[5];
        for (i = 0; i < 6; i++)
            for (j = 0; j < 5; j++) l[i * 5 + j] = l_6x10[i * 10 + j];
        solve_svd(l, 6, 5, rho, b5, tmp, tmp + 30);
        double* betas = Betas[3];
        if (b5[0] < 0) {
            betas[0] = sqrt(-b5[0]);
            betas[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0;
        } else {
            betas[0] = sqrt(b5[0]);
            betas[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0;
        }
        if (b5[1] < 0) betas[0] = -betas[0];
        betas[2] = b5[3] / betas[0];
        betas[3] = 0.0;
        gauss_newton(l_6x10, rho, betas);
        rep_errors[3] = compute_R_and_t(e, ut, betas, Rs[3], ts[3]);
    }
    int N = 1;
    if (rep_errors[2] < rep_errors[1]) N = 2;
    if (rep_errors[3] < rep_errors[N]) N = 3;
    tvec[0] = ts[N][0]; tvec[1] = ts[N][1]; tvec[2] = ts[N][2];
    rodrigues_m2v(Rs[N], rvec);
}

constexpr int EPNP5_SCRATCH_DOUBLES = 320 + 144;

}  // namespace pnpm
