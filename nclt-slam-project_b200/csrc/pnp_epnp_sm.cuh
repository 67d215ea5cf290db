// K3, second version: 5-point EPnP (cv2.solvePnP(flags=EPNP) on a RANSAC minimal set, the model kernel of
// cv2.solvePnPRansac at visual_landmark_matcher.py:342-346) with every large or dynamically indexed array in SHARED
// memory instead of thread-local (stack) memory.
//
// The first version ran one thread per hypothesis with 464 doubles of scratch plus the EPnP state on the local
// stack: 255 registers, 29 M local loads/stores per 32 768 hypotheses, 227 MB of DRAM traffic (ncu, round 1).
// The arithmetic cannot be spread over the lanes of a warp without changing it: OpenCV accumulates every dot product
// and every column norm of the one-sided Jacobi SVD SEQUENTIALLY (sum += a[k] * b[k], k ascending), 5-point EPnP is
// chaotic, and a tree- or shuffle-reduction rounds differently.  A warp per hypothesis would have
// to walk those sums with 12 shuffles + 12 adds per dot product - all 32 lanes issuing for one hypothesis - which is
// ~18x the issue slots of "one hypothesis per lane", where every instruction works for 32 hypotheses.  So the mapping
// stays one hypothesis per lane, and what changes is WHERE the data lives:
//   * per lane a column of doubles in shared memory, element e of lane l at base[e * 32 + l] (bank-conflict free,
//     one 256-byte wavefront pair per access): the 12 x 12 MtM / U^T (144), then - rows 0..7 are dead after the SVD -
//     the 6 x 10 L system (60) and the scratch of the small SVDs (30); a second column of 50: barycentric
//     coordinates (20), singular values, V^T of the small solves;
//   * M (10 x 12) is never stored (MtM is accumulated from M's closed-form entries, same values, same order);
//   * V^T of the 12 x 12 SVD is never computed: OpenCV only needs "V was requested" to normalise the left vectors,
//     and U^T does not depend on V^T;
//   * the per-hypothesis inputs, rho, the Gauss-Newton 6 x 4 system and the best (R, t) so far stay in registers
//     (fully unrolled loops, static indices).
// Every floating-point operation and its order are those of pnp_math.cuh / oracle/cvmath.c (-fmad=false); the
// staged parity tests (tests/test_pnp_gpu.py) hold the result to the CPU restatement bit for bit.
#pragma once
#include "pnp_math.cuh"

namespace pnpm {

constexpr int SM_LANES = 32;                 // hypotheses per warp = stride of a lane's column
constexpr int SM_BIG = 144;                  // doubles per lane: MtM / U^T, later L + small-SVD rows
constexpr int SM_AUX = 50;                   // doubles per lane: alphas, W, small V^T
constexpr int EPNP5_SM_DOUBLES_PER_LANE = SM_BIG + SM_AUX;
constexpr int EPNP5_SM_BYTES_PER_WARP = EPNP5_SM_DOUBLES_PER_LANE * SM_LANES * 8;   // 49 664

struct SmCol {       // one lane's column
    double* p;
    __device__ __forceinline__ double& operator[](int e) const { return p[e * SM_LANES]; }
    __device__ __forceinline__ SmCol operator+(int o) const { return SmCol{p + o * SM_LANES}; }
};

// One-sided Jacobi of modules/core/src/lapack.cpp (JacobiSVDImpl_), At = N rows of length M.
// WANT_V: V was requested (row sort + normalisation of the left vectors happen); STORE_V: V^T is also accumulated.
template <int M, int N, bool WANT_V, bool STORE_V>
__device__ void jacobi_svd_sm(SmCol At, SmCol W, SmCol Vt) {
    const double eps = DBL_EPSILON * 10;
    const double minval = DBL_MIN;
    constexpr int max_iter = M > 30 ? M : 30;
    for (int i = 0; i < N; i++) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < M; k++) {
            const double t = At[i * M + k];
            sd += t * t;
        }
        W[i] = sd;
        if (STORE_V) {
            for (int k = 0; k < N; k++) Vt[i * N + k] = 0;
            Vt[i * N + i] = 1;
        }
    }
    // Row i stays in registers for the whole j loop (pairs (i, i+1) ... (i, N-1) all rotate it) and is written back once.
    // (Accumulating the NEXT pair's dot product inside the rotation loop - a third independent chain beside the two new
    // norms - was measured slower: 10.5 vs 8.9 ms per 345 k hypotheses; the extra live row costs more than it hides.)
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < N - 1; i++) {
            double ai[M];
#pragma unroll
            for (int k = 0; k < M; k++) ai[k] = At[i * M + k];
            bool row_dirty = false;
            for (int j = i + 1; j < N; j++) {
                double aj[M];
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < M; k++) {
                    aj[k] = At[j * M + k];
                    p += ai[k] * aj[k];
                }
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cv_hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
#pragma unroll
                for (int k = 0; k < M; k++) {
                    const double t0 = c * ai[k] + s * aj[k];
                    const double t1 = -s * ai[k] + c * aj[k];
                    ai[k] = t0;
                    At[j * M + k] = t1;
                    a += t0 * t0;
                    b += t1 * t1;
                }
                W[i] = a;
                W[j] = b;
                changed = true;
                row_dirty = true;
                if (STORE_V) {
#pragma unroll
                    for (int k = 0; k < N; k++) {
                        const double vi = Vt[i * N + k], vj = Vt[j * N + k];
                        const double t0 = c * vi + s * vj;
                        const double t1 = -s * vi + c * vj;
                        Vt[i * N + k] = t0;
                        Vt[j * N + k] = t1;
                    }
                }
            }
            if (row_dirty) {
#pragma unroll
                for (int k = 0; k < M; k++) At[i * M + k] = ai[k];
            }
        }
        if (!changed) break;
    }
    for (int i = 0; i < N; i++) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < M; k++) {
            const double t = At[i * M + k];
            sd += t * t;
        }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < N - 1; i++) {
        int j = i;
        for (int k = i + 1; k < N; k++)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            if (WANT_V) {
                for (int k = 0; k < M; k++) { t = At[i * M + k]; At[i * M + k] = At[j * M + k]; At[j * M + k] = t; }
                if (STORE_V)
                    for (int k = 0; k < N; k++) { t = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = t; }
            }
        }
    }
    if (!WANT_V) return;
    uint64_t rng = 0x12345678ull;
    for (int i = 0; i < N; i++) {
        double sd = W[i];
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            const double val0 = 1. / M;
            for (int k = 0; k < M; k++) {
                rng = (uint64_t)(uint32_t)rng * 4164903690ull + (uint32_t)(rng >> 32);
                const uint32_t r = (uint32_t)rng;
                At[i * M + k] = (r & 256) != 0 ? val0 : -val0;
            }
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0;
                    for (int k = 0; k < M; k++) sd += At[i * M + k] * At[j * M + k];
                    double asum = 0;
                    for (int k = 0; k < M; k++) {
                        const double t = At[i * M + k] - sd * At[j * M + k];
                        At[i * M + k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < M; k++) At[i * M + k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < M; k++) {
                const double t = At[i * M + k];
                sd += t * t;
            }
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
#pragma unroll
        for (int k = 0; k < M; k++) At[i * M + k] *= s;
    }
}

// shared-memory scratch of the small decompositions: At (<= 30) in the big column, W (<= 6) and V^T (<= 25) in the aux one
struct SmScratch {
    SmCol at, w, vt;
};

// 3x3 SVD in cv::SVD::compute layout from a register matrix A (row-major): w[3], u[9] (columns = left vectors), vt[9]
__device__ __forceinline__ void svd3_sm(const double (&A)[9], double (&w)[3], double (&u)[9], double (&vt)[9], const SmScratch& sc) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) sc.at[j * 3 + i] = A[i * 3 + j];
    jacobi_svd_sm<3, 3, true, true>(sc.at, sc.w, sc.vt);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        w[i] = sc.w[i];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            u[i * 3 + j] = sc.at[j * 3 + i];
            vt[i * 3 + j] = sc.vt[i * 3 + j];
        }
    }
}

__device__ __forceinline__ void invert3_svd_sm(const double (&A)[9], double (&Ainv)[9], const SmScratch& sc) {
    double w[3], u[9], vt[9];
    svd3_sm(A, w, u, vt, sc);
    const double threshold = (w[0] + w[1] + w[2]) * (DBL_EPSILON * 2);
#pragma unroll
    for (int i = 0; i < 9; i++) Ainv[i] = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double buffer[3];
#pragma unroll
        for (int j = 0; j < 3; j++) buffer[j] = u[j * 3 + i] * wi;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double s = vt[i * 3 + k];
#pragma unroll
            for (int j = 0; j < 3; j++) Ainv[k * 3 + j] = Ainv[k * 3 + j] + s * buffer[j];
        }
    }
}

// cv::solve(A[6 x N], rho, DECOMP_SVD) where A's columns are columns COLS[] of the 6 x 10 system L (shared memory)
template <int N>
__device__ __forceinline__ void solve_svd_L(SmCol L, const int (&cols)[N], const double (&b)[6], double (&x)[N], const SmScratch& sc) {
    constexpr int M = 6;
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
        for (int j = 0; j < N; j++) sc.at[j * M + i] = L[i * 10 + cols[j]];
    jacobi_svd_sm<M, N, true, true>(sc.at, sc.w, sc.vt);
    double threshold = 0;
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 0;
    for (int i = 0; i < N; i++) threshold += sc.w[i];
    threshold *= DBL_EPSILON * 2;
    for (int i = 0; i < N; i++) {
        double wi = sc.w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
#pragma unroll
        for (int j = 0; j < M; j++) s += sc.at[i * M + j] * b[j];
        s *= wi;
#pragma unroll
        for (int j = 0; j < N; j++) x[j] = x[j] + s * sc.vt[i * N + j];
    }
}

// Householder QR least squares of EPnP's Gauss-Newton step (6 x 4) with static indices (registers).  The first
// loop of every column looks at rows k .. nr-2 (the pointer is advanced AFTER the read): OpenCV's quirk, kept.
__device__ __forceinline__ void qr_solve_6x4_reg(double (&A)[24], double (&b)[6], double (&X)[4]) {
    constexpr int nr = 6, nc = 4;
    double A1[4], A2[4];
#pragma unroll
    for (int k = 0; k < nc; k++) {
        double eta = fabs(A[k * nc + k]);
#pragma unroll
        for (int i = k + 1; i < nr; i++) {
            const double elt = fabs(A[(i - 1) * nc + k]);
            if (eta < elt) eta = elt;
        }
        if (eta == 0) {
            A1[k] = A2[k] = 0.0;
            return;
        }
        double sum2 = 0.0;
        const double inv_eta = 1. / eta;
#pragma unroll
        for (int i = k; i < nr; i++) {
            A[i * nc + k] *= inv_eta;
            sum2 += A[i * nc + k] * A[i * nc + k];
        }
        double sigma = sqrt(sum2);
        if (A[k * nc + k] < 0) sigma = -sigma;
        A[k * nc + k] += sigma;
        A1[k] = sigma * A[k * nc + k];
        A2[k] = -eta * sigma;
#pragma unroll
        for (int j = k + 1; j < nc; j++) {
            double sum = 0;
#pragma unroll
            for (int i = k; i < nr; i++) sum += A[i * nc + k] * A[i * nc + j];
            const double tau = sum / A1[k];
#pragma unroll
            for (int i = k; i < nr; i++) A[i * nc + j] -= tau * A[i * nc + k];
        }
    }
#pragma unroll
    for (int j = 0; j < nc; j++) {
        double tau = 0;
#pragma unroll
        for (int i = j; i < nr; i++) tau += A[i * nc + j] * b[i];
        tau /= A1[j];
#pragma unroll
        for (int i = j; i < nr; i++) b[i] -= tau * A[i * nc + j];
    }
    X[nc - 1] = b[nc - 1] / A2[nc - 1];
#pragma unroll
    for (int i = nc - 2; i >= 0; i--) {
        double sum = 0;
#pragma unroll
        for (int j = i + 1; j < nc; j++) sum += A[i * nc + j] * X[j];
        X[i] = (b[i] - sum) / A2[i];
    }
}

__device__ __forceinline__ void gauss_newton_sm(SmCol L, const double (&rho)[6], double (&betas)[4]) {
    double a[24], b[6], x[4] = {0, 0, 0, 0};
    for (int k = 0; k < 5; k++) {
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double rowL[10];
#pragma unroll
            for (int q = 0; q < 10; q++) rowL[q] = L[i * 10 + q];
            a[i * 4 + 0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
            a[i * 4 + 1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
            a[i * 4 + 2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
            a[i * 4 + 3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
            b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] +
                             rowL[2] * betas[1] * betas[1] + rowL[3] * betas[0] * betas[2] +
                             rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                             rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] +
                             rowL[8] * betas[2] * betas[3] + rowL[9] * betas[3] * betas[3]);
        }
        qr_solve_6x4_reg(a, b, x);
#pragma unroll
        for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
}

struct Epnp5In {           // registers
    float obj[15], img[10];
    double fu, fv, uc, vc;
};
// image point i as EPnP sees it: undistortPoints wrote the normalised coordinates as float32
__device__ __forceinline__ void epnp_uv(const Epnp5In& in, int i, double& u, double& v) {
    const double ifx = 1. / in.fu, ify = 1. / in.fv;
    const double x = ((double)in.img[2 * i] - in.uc) * ifx;
    const double y = ((double)in.img[2 * i + 1] - in.vc) * ify;
    u = (double)(float)x * in.fu + in.uc;
    v = (double)(float)y * in.fv + in.vc;
}

// epnp::compute_R_and_t for one beta set: ut = U^T rows of the 12 x 12 SVD (shared), alphas (shared)
__device__ double compute_R_and_t_sm(const Epnp5In& in, SmCol ut, SmCol alphas, const double (&betas)[4], double (&R)[9],
                                     double (&t)[3], const SmScratch& sc) {
    constexpr int n = 5;
    double ccs[4][3], pcs[15];
#pragma unroll
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const SmCol v = ut + 12 * (11 - i);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
        const double a0 = alphas[4 * i], a1 = alphas[4 * i + 1], a2 = alphas[4 * i + 2], a3 = alphas[4 * i + 3];
#pragma unroll
        for (int j = 0; j < 3; j++) pcs[3 * i + j] = a0 * ccs[0][j] + a1 * ccs[1][j] + a2 * ccs[2][j] + a3 * ccs[3][j];
    }
    if (pcs[2] < 0.0) {
        // (the control points' sign flip has no reader afterwards; the points' flip does)
#pragma unroll
        for (int i = 0; i < 15; i++) pcs[i] = -pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            pc0[j] += pcs[3 * i + j];
            pw0[j] += (double)in.obj[3 * i + j];
        }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        pc0[j] /= n;
        pw0[j] /= n;
    }
    double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, abt_d[3], abt_u[9], abt_vt[9], abt_v[9];
#pragma unroll
    for (int i = 0; i < n; i++) {
        const double pw[3] = {(double)in.obj[3 * i], (double)in.obj[3 * i + 1], (double)in.obj[3 * i + 2]};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            abt[3 * j] += (pcs[3 * i + j] - pc0[j]) * (pw[0] - pw0[0]);
            abt[3 * j + 1] += (pcs[3 * i + j] - pc0[j]) * (pw[1] - pw0[1]);
            abt[3 * j + 2] += (pcs[3 * i + j] - pc0[j]) * (pw[2] - pw0[2]);
        }
    }
    svd3_sm(abt, abt_d, abt_u, abt_vt, sc);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) abt_v[i * 3 + j] = abt_vt[j * 3 + i];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            R[i * 3 + j] = abt_u[3 * i] * abt_v[3 * j] + abt_u[3 * i + 1] * abt_v[3 * j + 1] + abt_u[3 * i + 2] * abt_v[3 * j + 2];
    const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] -
                       R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
    if (det < 0) {
        R[6] = -R[6];
        R[7] = -R[7];
        R[8] = -R[8];
    }
    t[0] = pc0[0] - (R[0] * pw0[0] + R[1] * pw0[1] + R[2] * pw0[2]);
    t[1] = pc0[1] - (R[3] * pw0[0] + R[4] * pw0[1] + R[5] * pw0[2]);
    t[2] = pc0[2] - (R[6] * pw0[0] + R[7] * pw0[1] + R[8] * pw0[2]);
    double sum2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) {
        const double pw[3] = {(double)in.obj[3 * i], (double)in.obj[3 * i + 1], (double)in.obj[3 * i + 2]};
        const double Xc = (R[0] * pw[0] + R[1] * pw[1] + R[2] * pw[2]) + t[0];
        const double Yc = (R[3] * pw[0] + R[4] * pw[1] + R[5] * pw[2]) + t[1];
        const double inv_Zc = 1.0 / ((R[6] * pw[0] + R[7] * pw[1] + R[8] * pw[2]) + t[2]);
        const double ue = in.uc + in.fu * Xc * inv_Zc;
        const double ve = in.vc + in.fv * Yc * inv_Zc;
        double u, v;
        epnp_uv(in, i, u, v);
        sum2 += sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
}

// cv::Rodrigues matrix -> vector with the 3x3 SVD in shared memory
__device__ __forceinline__ void rodrigues_m2v_sm(const double (&Rin)[9], double* rvec, const SmScratch& sc) {
    double w[3], U[9], Vt[9], R[9];
    svd3_sm(Rin, w, U, Vt, sc);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += U[i * 3 + k] * Vt[k * 3 + j];
            R[i * 3 + j] = s;
        }
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
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

// entry (row k, column c) of EPnP's 10 x 12 matrix M, exactly as epnp::fill_M writes it
__device__ __forceinline__ double epnp_M_entry(const Epnp5In& in, SmCol alphas, int pt, int second, int c, double u, double v) {
    const double as = alphas[4 * pt + c / 3];
    const int comp = c % 3;
    if (second == 0) return comp == 0 ? as * in.fu : comp == 1 ? 0.0 : as * (in.uc - u);
    return comp == 0 ? 0.0 : comp == 1 ? as * in.fv : as * (in.vc - v);
}

// cv2.solvePnP(5 points f32, K, dist = 0, flags = EPNP) -> rvec[3], tvec[3].  big / aux: this lane's shared-memory
// columns (SM_BIG / SM_AUX doubles at stride 32).
__device__ void solvepnp_epnp5_sm(const Epnp5In& in, double* rvec, double* tvec, SmCol big, SmCol aux) {
    constexpr int n = 5;
    const SmCol alphas = aux;                      // [0, 20)
    const SmScratch sc{big + 60, aux + 20, aux + 25};   // small SVDs: At big[60, 90), W aux[20, 25), V^T aux[25, 50)
    double rho[6];
    {
        // choose_control_points
        double cws[4][3];
        cws[0][0] = cws[0][1] = cws[0][2] = 0;
#pragma unroll
        for (int i = 0; i < n; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) cws[0][j] += (double)in.obj[3 * i + j];
#pragma unroll
        for (int j = 0; j < 3; j++) cws[0][j] /= n;
        {
            double pw0[15], pw0tpw0[9], dc[3], u[9], vt[9];
#pragma unroll
            for (int i = 0; i < n; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) pw0[3 * i + j] = (double)in.obj[3 * i + j] - cws[0][j];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = i; j < 3; j++) {
                    double s = 0;
#pragma unroll
                    for (int k = 0; k < n; k++) s += pw0[3 * k + i] * pw0[3 * k + j];
                    pw0tpw0[i * 3 + j] = s;
                    pw0tpw0[j * 3 + i] = s;
                }
            svd3_sm(pw0tpw0, dc, u, vt, sc);
#pragma unroll
            for (int i = 1; i < 4; i++) {
                const double kk = sqrt(dc[i - 1] / n);
#pragma unroll
                for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + kk * u[j * 3 + (i - 1)];
            }
        }
        // compute_barycentric_coordinates
        {
            double cc[9], ci[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
            invert3_svd_sm(cc, ci, sc);
#pragma unroll
            for (int i = 0; i < n; i++) {
                const double p0 = (double)in.obj[3 * i], p1 = (double)in.obj[3 * i + 1], p2 = (double)in.obj[3 * i + 2];
                double a[4];
#pragma unroll
                for (int j = 0; j < 3; j++)
                    a[1 + j] = ci[3 * j] * (p0 - cws[0][0]) + ci[3 * j + 1] * (p1 - cws[0][1]) + ci[3 * j + 2] * (p2 - cws[0][2]);
                a[0] = 1.0 - a[1] - a[2] - a[3];
#pragma unroll
                for (int j = 0; j < 4; j++) alphas[4 * i + j] = a[j];
            }
        }
        rho[0] = dist2(cws[0], cws[1]);
        rho[1] = dist2(cws[0], cws[2]);
        rho[2] = dist2(cws[0], cws[3]);
        rho[3] = dist2(cws[1], cws[2]);
        rho[4] = dist2(cws[1], cws[3]);
        rho[5] = dist2(cws[2], cws[3]);
    }
    // MtM (symmetric; its transpose is itself) accumulated over the 10 rows of M in order, M never stored
    {
        double us[10];
#pragma unroll
        for (int i = 0; i < n; i++) epnp_uv(in, i, us[2 * i], us[2 * i + 1]);
        for (int i = 0; i < 12; i++)
            for (int j = i; j < 12; j++) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < 2 * n; k++)
                    s += epnp_M_entry(in, alphas, k >> 1, k & 1, i, us[k & ~1], us[k | 1]) *
                         epnp_M_entry(in, alphas, k >> 1, k & 1, j, us[k & ~1], us[k | 1]);
                big[i * 12 + j] = s;
                big[j * 12 + i] = s;
            }
    }
    // V is "requested" (the left vectors get normalised) but never stored; W in aux[20, 32)
    jacobi_svd_sm<12, 12, true, false>(big, aux + 20, aux + 20);
    const SmCol ut = big;            // rows 8..11 = big[96, 144) are read from here on; rows 0..7 are dead
    const SmCol L = big;             // the 6 x 10 system lives in big[0, 60)
    {
        // row i of L belongs to the control-point pair (a, b) = (0,1) (0,2) (0,3) (1,2) (1,3) (2,3); dv[q] is that
        // difference in null-space vector q (12 live doubles instead of a 4 x 6 x 3 table)
        constexpr int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double dv[4][3];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const SmCol v = ut + 12 * (11 - q);
                dv[q][0] = v[3 * PA[i]] - v[3 * PB[i]];
                dv[q][1] = v[3 * PA[i] + 1] - v[3 * PB[i] + 1];
                dv[q][2] = v[3 * PA[i] + 2] - v[3 * PB[i] + 2];
            }
            L[10 * i + 0] = dot3(dv[0], dv[0]);
            L[10 * i + 1] = 2.0 * dot3(dv[0], dv[1]);
            L[10 * i + 2] = dot3(dv[1], dv[1]);
            L[10 * i + 3] = 2.0 * dot3(dv[0], dv[2]);
            L[10 * i + 4] = 2.0 * dot3(dv[1], dv[2]);
            L[10 * i + 5] = dot3(dv[2], dv[2]);
            L[10 * i + 6] = 2.0 * dot3(dv[0], dv[3]);
            L[10 * i + 7] = 2.0 * dot3(dv[1], dv[3]);
            L[10 * i + 8] = 2.0 * dot3(dv[2], dv[3]);
            L[10 * i + 9] = dot3(dv[3], dv[3]);
        }
    }
    double bestR[9], bestT[3], best_err;
    // N = 1
    {
        const int cols[4] = {0, 1, 3, 6};
        double b4[4], betas[4];
        solve_svd_L<4>(L, cols, rho, b4, sc);
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
        gauss_newton_sm(L, rho, betas);
        best_err = compute_R_and_t_sm(in, ut, alphas, betas, bestR, bestT, sc);
    }
    // N = 2
    {
        const int cols[3] = {0, 1, 2};
        double b3[3], betas[4], R[9], t[3];
        solve_svd_L<3>(L, cols, rho, b3, sc);
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
        gauss_newton_sm(L, rho, betas);
        const double err = compute_R_and_t_sm(in, ut, alphas, betas, R, t, sc);
        if (err < best_err) {          // rep_errors[2] < rep_errors[1]
            best_err = err;
#pragma unroll
            for (int k = 0; k < 9; k++) bestR[k] = R[k];
            bestT[0] = t[0]; bestT[1] = t[1]; bestT[2] = t[2];
        }
    }
    // N = 3
    {
        const int cols[5] = {0, 1, 2, 3, 4};
        double b5[5], betas[4], R[9], t[3];
        solve_svd_L<5>(L, cols, rho, b5, sc);
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
        gauss_newton_sm(L, rho, betas);
        const double err = compute_R_and_t_sm(in, ut, alphas, betas, R, t, sc);
        if (err < best_err) {          // rep_errors[3] < rep_errors[N]
#pragma unroll
            for (int k = 0; k < 9; k++) bestR[k] = R[k];
            bestT[0] = t[0]; bestT[1] = t[1]; bestT[2] = t[2];
        }
    }
    tvec[0] = bestT[0]; tvec[1] = bestT[1]; tvec[2] = bestT[2];
    rodrigues_m2v_sm(bestR, rvec, sc);
}

}  // namespace pnpm
