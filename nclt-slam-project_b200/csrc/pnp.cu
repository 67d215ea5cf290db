// K3-K5: PnP-RANSAC on the GPU, replacing cv2.solvePnPRansac(obj, img, K, DIST,
// iterationsCount=200, reprojectionError=3.0, flags=SOLVEPNP_ITERATIVE) and the
// cv2.projectPoints mean-error gate (visual_landmark_matcher.py:342-356,
// checkpoint_a_selftest.py:78-88).  SURVEY.md section 8a rows a5/a6, Appendix A.
//
//   k_pnp_sets      one thread per problem: the MWC minimal-set sequence (sequential RNG)
//   k_pnp_hypo      K3, one lane per (problem, iteration): 5-point EPnP in FP64 -> rvec,tvec; the 12 x 12 Jacobi SVD, the
//                   6 x 10 system and the small decompositions live in shared memory, one column per lane
//                   (pnp_epnp_sm.cuh - no local-memory traffic; why not one warp per hypothesis is explained there)
//   k_pnp_score     K4, one warp per (problem, iteration): project all points in FP64, round to
//                   float32, float32 squared error <= thr^2, ballot/popc inlier count
//   k_pnp_finish    K5, one warp per problem: replay OpenCV's sequential best/early-stop rule over
//                   the counts, rebuild the winning mask, Levenberg-Marquardt refinement on the
//                   inliers, mean reprojection error of the refined pose
// All hypotheses of all problems are evaluated in parallel; the sequential semantics (strict >
// on max(best,4), niters shrinking in place) are restored by the replay.
// Compiled with -fmad=false: operation order is part of the contract (pnp_math.cuh).
#include "common.cuh"
#include "pnp_math.cuh"
#include "pnp_epnp_sm.cuh"

namespace {

struct PnpView {
    const float* obj;   // [P][Nmax][3]
    const float* img;   // [P][Nmax][2]
    const int* n;       // [P]
    int P, Nmax;          // P = problems the buffers are sized for
    const int* P_dev;     // optional device-side problem count (async pipeline): effective P = min(*P_dev, P)
    double fx, fy, cx, cy;
    int iters;
    float thr2;
    double conf;
    int refine;
};

__device__ __forceinline__ int eff_P(const PnpView& v) { return v.P_dev ? min(*v.P_dev, v.P) : v.P; }

__device__ __forceinline__ uint32_t mwc_next(uint64_t& st) {
    st = (uint64_t)(uint32_t)st * 4164903690ull + (uint32_t)(st >> 32);
    return (uint32_t)st;
}

// Minimal sets of all iterations, one WARP per problem.  cv::RNG((uint64)-1) produces the same raw stream for every call;
// what differs per problem is `% n` and the re-draws after a duplicate, which shift the rest of the stream.  The raw
// outputs come from a table (`raw`, computed once per context); a round lets 32 lanes take 32 consecutive iterations
// assuming no re-draws before them (offset = base + 5 * lane): every lane up to and including the first one that
// re-drew started from the right offset, those are committed, and the next round starts behind them.  Re-draws are rare
// (~3 % of the iterations at n = 334), so 200 iterations take ~13 rounds instead of a 1000-step sequential chain
// (66 -> 7 us per call).  A problem that would run past the table falls back to the sequential recurrence.
__device__ void pnp_sets_sequential(int n, int iters, int* out) {
    uint64_t st = ~0ull;   // cv::RNG((uint64)-1), fresh for every solvePnPRansac call
    for (int it = 0; it < iters; it++) {
        int idx[5];
        for (int i = 0; i < 5; i++) {
            for (;;) {
                int x = (int)(mwc_next(st) % (uint32_t)n);
                bool dup = false;
                for (int j = 0; j < i; j++) dup |= (idx[j] == x);
                idx[i] = x;
                if (!dup) break;
            }
            out[it * 5 + i] = idx[i];
        }
    }
}

__global__ void __launch_bounds__(128) k_pnp_sets(PnpView v, int* sets /*[P][iters][5]*/, const uint32_t* __restrict__ raw, int raw_n) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= eff_P(v)) return;
    const int n = v.n[p];
    int* out = sets + (size_t)p * v.iters * 5;
    if (n <= 5) {
        for (int i = lane; i < v.iters * 5; i += 32) out[i] = n == 5 ? i % 5 : -1;
        return;
    }
    int base = 0;                 // position in the raw stream where iteration `it0` starts
    bool overrun = false;
    for (int it0 = 0; it0 < v.iters && !overrun;) {
        const int it = it0 + lane;
        int pos = base + 5 * lane, idx[5] = {0, 0, 0, 0, 0};
        bool bad = false;         // ran past the table
        if (it < v.iters) {
#pragma unroll
            for (int i = 0; i < 5; i++) {
                for (;;) {
                    if (pos >= raw_n) { bad = true; break; }
                    const int x = (int)(raw[pos++] % (uint32_t)n);
                    bool dup = false;
#pragma unroll
                    for (int j = 0; j < i; j++) dup |= (idx[j] == x);
                    idx[i] = x;
                    if (!dup) break;
                }
                if (bad) break;
            }
        }
        const int used = pos - (base + 5 * lane);
        // lanes below the first one that re-drew (or ran out of iterations) started from the right offset; so did that lane
        const unsigned redraw = __ballot_sync(0xFFFFFFFFu, it < v.iters && used != 5);
        const int last = redraw ? __ffs(redraw) - 1 : 31;           // last lane whose result stands
        if (__shfl_sync(0xFFFFFFFFu, (int)bad, last) || __ballot_sync(0xFFFFFFFFu, bad && lane <= last)) {
            overrun = true;
            break;
        }
        if (lane <= last && it < v.iters) {
#pragma unroll
            for (int i = 0; i < 5; i++) out[it * 5 + i] = idx[i];
        }
        base = __shfl_sync(0xFFFFFFFFu, pos, last);                 // the stream continues behind lane `last`
        it0 += last + 1;
    }
    if (overrun) {
        __syncwarp();
        if (lane == 0) pnp_sets_sequential(n, v.iters, out);
    }
}

// Per-problem RANSAC replay state: {cursor, best, max_good, niters}. Hypotheses are evaluated in
// rounds of `it_cnt` iterations; iterations OpenCV's loop would never reach (it >= niters, which
// shrinks as soon as a good model appears) are skipped.
#ifndef NCLT_CORESIDENT
#define NCLT_CORESIDENT 0
#endif
#if NCLT_CORESIDENT
#define PNP_WARP_KERNEL __maxnreg__(200)
#define PNP_SCORE_THREADS 64
#else
#define PNP_WARP_KERNEL __launch_bounds__(32)
#define PNP_SCORE_THREADS 128
#endif
__global__ void PNP_WARP_KERNEL k_pnp_hypo(PnpView v, const int* __restrict__ sets, double* models /*[P][iters][6]*/,
                                                 int it_lo, int it_cnt, const int* __restrict__ state) {
    extern __shared__ __align__(16) double sm_cols[];      // [EPNP5_SM_DOUBLES_PER_LANE][32]
    int gg = blockIdx.x * blockDim.x + threadIdx.x;
    if (gg >= eff_P(v) * it_cnt) return;
    int p = gg / it_cnt;
    int it = it_lo + gg % it_cnt;
    if (it >= v.iters) return;
    if (state && it >= state[4 * p + 3]) return;
    int g = p * v.iters + it;
    int n = v.n[p];
    double* out = models + (size_t)g * 6;
    if (n < 5) {
        for (int k = 0; k < 6; k++) out[k] = 0;
        return;
    }
    const int* idx = sets + (size_t)g * 5;
    const float* obj = v.obj + (size_t)p * v.Nmax * 3;
    const float* img = v.img + (size_t)p * v.Nmax * 2;
    pnpm::Epnp5In in;
    in.fu = v.fx; in.fv = v.fy; in.uc = v.cx; in.vc = v.cy;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        int j = idx[i];
        in.obj[3 * i] = obj[3 * j]; in.obj[3 * i + 1] = obj[3 * j + 1]; in.obj[3 * i + 2] = obj[3 * j + 2];
        in.img[2 * i] = img[2 * j]; in.img[2 * i + 1] = img[2 * j + 1];
    }
    double r[3], t[3];
    const pnpm::SmCol big{sm_cols + threadIdx.x};
    const pnpm::SmCol aux{sm_cols + pnpm::SM_BIG * pnpm::SM_LANES + threadIdx.x};
    pnpm::solvepnp_epnp5_sm(in, r, t, big, aux);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
    out[3] = t[0]; out[4] = t[1]; out[5] = t[2];
}

// PnPRansacCallback::computeError for one point: projectPoints in double -> float32, then the
// float32 squared distance.
__device__ __forceinline__ float reproj_err2(const double* R, const double* t, double fx, double fy, double cx,
                                             double cy, float X, float Y, float Z, float u, float v, float* pu_out,
                                             float* pv_out) {
    double Xd = X, Yd = Y, Zd = Z;
    double x = R[0] * Xd + R[1] * Yd + R[2] * Zd + t[0];
    double y = R[3] * Xd + R[4] * Yd + R[5] * Zd + t[1];
    double z = R[6] * Xd + R[7] * Yd + R[8] * Zd + t[2];
    z = z ? 1. / z : 1;
    x *= z;
    y *= z;
    float pu = (float)(x * fx + cx);
    float pv = (float)(y * fy + cy);
    if (pu_out) { *pu_out = pu; *pv_out = pv; }
    float dx = u - pu, dy = v - pv;
    float s = 0.f;
    s += dx * dx;
    s += dy * dy;
    return s;
}

__global__ void __launch_bounds__(128) k_pnp_score(PnpView v, const double* __restrict__ models, int* counts, int it_lo,
                                                   int it_cnt, const int* __restrict__ state) {
    int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (gw >= eff_P(v) * it_cnt) return;
    int p = gw / it_cnt;
    int it = it_lo + gw % it_cnt;
    if (it >= v.iters) return;
    if (state && it >= state[4 * p + 3]) return;
    int warp = p * v.iters + it;
    int n = v.n[p];
    if (n < 5) {
        if (lane == 0) counts[warp] = 0;
        return;
    }
    const double* m = models + (size_t)warp * 6;
    double R[9], t[3] = {m[3], m[4], m[5]};
    pnpm::rodrigues_v2m(m, R);
    const float* obj = v.obj + (size_t)p * v.Nmax * 3;
    const float* img = v.img + (size_t)p * v.Nmax * 2;
    int good = 0;
    for (int i = lane; i < n; i += 32) {
        float e = reproj_err2(R, t, v.fx, v.fy, v.cx, v.cy, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i],
                              img[2 * i + 1], nullptr, nullptr);
        good += (e <= v.thr2) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xFFFFFFFFu, good, o);
    if (lane == 0) counts[warp] = good;
}

__device__ int ransac_update_niters(double p, double ep, int model_points, int max_iters) {
    p = p > 0. ? p : 0.;
    p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.;
    ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// Jacobian pieces of projectPoints w.r.t. (rvec, tvec); accumulates JtJ (upper+lower) and JtErr
struct LmAcc {
    double JtJ[21];   // packed upper triangle, row-major
    double JtE[6];
    double e2;
};

__device__ void drdr_from_rvec(const double* r, double* dRdr /*27*/) {
    double rx = r[0], ry = r[1], rz = r[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    for (int i = 0; i < 27; i++) dRdr[i] = 0;
    if (theta < DBL_EPSILON) {
        dRdr[5] = dRdr[15] = dRdr[19] = -1;
        dRdr[7] = dRdr[11] = dRdr[21] = 1;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    double ux = rx * itheta, uy = ry * itheta, uz = rz * itheta;
    const double rrt[9] = {ux * ux, ux * uy, ux * uz, ux * uy, uy * uy, uy * uz, ux * uz, uy * uz, uz * uz};
    const double r_x[9] = {0, -uz, uy, uz, 0, -ux, -uy, ux, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double drrt[27] = {ux + ux, uy, uz, uy, 0, 0, uz, 0, 0, 0, ux, 0, ux, uy + uy, uz, 0, uz, 0,
                             0, 0, ux, 0, 0, uy, ux, uy, uz + uz};
    const double d_r_x_[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; i++) {
        double ri = i == 0 ? ux : i == 1 ? uy : uz;
        double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
        double a3 = (c - s * itheta) * ri, a4 = s * itheta;
        for (int k = 0; k < 9; k++)
            dRdr[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x_[i * 9 + k];
    }
}

__device__ __forceinline__ double warp_sum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
    return x;
}

// residual norm^2 (and optionally the normal equations) over the inliers, warp-cooperative
__device__ double lm_eval(const PnpView& v, const float* obj, const float* img, const unsigned char* mask, int n,
                          const double* param, bool want_jac, double* JtJ /*36*/, double* JtE /*6*/, int lane) {
    double R[9], dRdr[27];
    pnpm::rodrigues_v2m(param, R);
    if (want_jac) drdr_from_rvec(param, dRdr);
    double acc[27];
    for (int i = 0; i < 27; i++) acc[i] = 0;
    double e2 = 0;
    for (int i = lane; i < n; i += 32) {
        if (!mask[i]) continue;
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + param[3];
        double y = R[3] * X + R[4] * Y + R[5] * Z + param[4];
        double z = R[6] * X + R[7] * Y + R[8] * Z + param[5];
        double iz = z ? 1. / z : 1;
        double xn = x * iz, yn = y * iz;
        double eu = xn * v.fx + v.cx - (double)img[2 * i];
        double ev = yn * v.fy + v.cy - (double)img[2 * i + 1];
        e2 += eu * eu + ev * ev;
        if (want_jac) {
            double Ju[6], Jv[6];
            Ju[3] = v.fx * iz; Ju[4] = 0; Ju[5] = -v.fx * xn * iz;
            Jv[3] = 0; Jv[4] = v.fy * iz; Jv[5] = -v.fy * yn * iz;
            for (int k = 0; k < 3; k++) {
                const double* D = dRdr + 9 * k;
                double dx = D[0] * X + D[1] * Y + D[2] * Z;
                double dy = D[3] * X + D[4] * Y + D[5] * Z;
                double dz = D[6] * X + D[7] * Y + D[8] * Z;
                Ju[k] = v.fx * (dx * iz - xn * iz * dz);
                Jv[k] = v.fy * (dy * iz - yn * iz * dz);
            }
            int q = 0;
            for (int a = 0; a < 6; a++) {
                acc[21 + a] += Ju[a] * eu + Jv[a] * ev;
                for (int b = a; b < 6; b++) acc[q++] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
            }
        }
    }
    e2 = warp_sum(e2);
    if (want_jac) {
        for (int i = 0; i < 27; i++) acc[i] = warp_sum(acc[i]);
        int q = 0;
        for (int a = 0; a < 6; a++) {
            JtE[a] = acc[21 + a];
            for (int b = a; b < 6; b++) {
                JtJ[a * 6 + b] = acc[q];
                JtJ[b * 6 + a] = acc[q];
                q++;
            }
        }
    }
    return e2;
}

struct PnpOut {
    unsigned char* ok;       // [P]
    double* rvec;            // [P][3]
    double* tvec;            // [P][3]
    int* n_inliers;          // [P]
    unsigned char* mask;     // [P][Nmax] (required: used as workspace)
    float* mean_err;         // [P] or null
    int* best_iter;          // [P] or null
    int* niters;             // [P] or null
};

// sequential part of RANSACPointSetRegistrator::run over the counts of one round: strict > on
// max(best, 4), niters shrinking in place. One thread per problem.
__global__ void k_pnp_replay(PnpView v, const int* __restrict__ counts, int* state, int it_hi) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= eff_P(v)) return;
    int n = v.n[p];
    int cursor = state[4 * p], best = state[4 * p + 1], max_good = state[4 * p + 2], niters = state[4 * p + 3];
    if (n == 5) {
        best = 0; max_good = 5; cursor = niters;
    } else if (n > 5) {
        const int* cnt = counts + (size_t)p * v.iters;
        int hi = it_hi < niters ? it_hi : niters;
        for (; cursor < hi; cursor++) {
            int good = cnt[cursor];
            if (good > (max_good > 4 ? max_good : 4)) {
                max_good = good;
                best = cursor;
                niters = ransac_update_niters(v.conf, (double)(n - good) / n, 5, niters);
                hi = it_hi < niters ? it_hi : niters;
            }
        }
    } else {
        cursor = niters;
    }
    state[4 * p] = cursor; state[4 * p + 1] = best; state[4 * p + 2] = max_good; state[4 * p + 3] = niters;
}

__global__ void k_pnp_state_init(int P, int iters, int* state) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    state[4 * p] = 0; state[4 * p + 1] = -1; state[4 * p + 2] = 0; state[4 * p + 3] = iters > 1 ? iters : 1;
}

__global__ void PNP_WARP_KERNEL k_pnp_finish(PnpView v, const double* __restrict__ models,
                                                    const int* __restrict__ state, PnpOut o) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= eff_P(v)) return;
    const int n = v.n[p];
    const float* obj = v.obj + (size_t)p * v.Nmax * 3;
    const float* img = v.img + (size_t)p * v.Nmax * 2;
    unsigned char* mask = o.mask + (size_t)p * v.Nmax;
    const int best = state[4 * p + 1], max_good = state[4 * p + 2], niters = state[4 * p + 3];
    if (lane == 0) {
        if (o.best_iter) o.best_iter[p] = best;
        if (o.niters) o.niters[p] = niters;
    }
    double param[6] = {0, 0, 0, 0, 0, 0};
    if (best < 0) {
        for (int i = lane; i < v.Nmax; i += 32) mask[i] = 0;
        if (lane == 0) {
            o.ok[p] = 0;
            o.n_inliers[p] = 0;
            for (int k = 0; k < 3; k++) { o.rvec[p * 3 + k] = 0; o.tvec[p * 3 + k] = 0; }
            if (o.mean_err) o.mean_err[p] = 0.f;
        }
        return;
    }
    const double* m = models + ((size_t)p * v.iters + best) * 6;
    for (int k = 0; k < 6; k++) param[k] = m[k];

    // --- winning mask ------------------------------------------------------------------
    {
        double R[9];
        pnpm::rodrigues_v2m(param, R);
        for (int i = lane; i < v.Nmax; i += 32) {
            unsigned char f = 0;
            if (i < n) {
                if (n == 5) f = 1;
                else {
                    float e = reproj_err2(R, param + 3, v.fx, v.fy, v.cx, v.cy, obj[3 * i], obj[3 * i + 1],
                                          obj[3 * i + 2], img[2 * i], img[2 * i + 1], nullptr, nullptr);
                    f = e <= v.thr2;
                }
            }
            mask[i] = f;
        }
        __syncwarp();
    }

    // --- Levenberg-Marquardt refinement (CvLevMarq: 20 iterations, FLT_EPSILON) -----------
    if (v.refine && n > 5) {
        double prev[6], JtJ[36], JtE[6], A[36], delta[6];
        double At[36], Vt[36];
        int lambdaLg10 = -3, iters = 0;
        const double LOG10 = log(10.);
        double prevErr2 = DBL_MAX, err2 = 0;
        for (;;) {
            double e0 = lm_eval(v, obj, img, mask, n, param, true, JtJ, JtE, lane);
            if (iters == 0) prevErr2 = e0;
            for (int k = 0; k < 6; k++) prev[k] = param[k];
            for (;;) {
                double lambda = exp(lambdaLg10 * LOG10);
                for (int k = 0; k < 36; k++) A[k] = JtJ[k];
                for (int k = 0; k < 6; k++) A[k * 6 + k] *= 1. + lambda;
                pnpm::solve_svd(A, 6, 6, JtE, delta, At, Vt);   // every lane solves the same system
                for (int k = 0; k < 6; k++) param[k] = prev[k] - delta[k];
                err2 = lm_eval(v, obj, img, mask, n, param, false, nullptr, nullptr, lane);
                // compare norms like cvNorm (sqrt is monotone: compare the squares' roots)
                if (sqrt(err2) > sqrt(prevErr2) && ++lambdaLg10 <= 16) continue;
                break;
            }
            lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
            double dn = 0, pn = 0;
            for (int k = 0; k < 6; k++) {
                double d = param[k] - prev[k];
                dn += d * d;
                pn += prev[k] * prev[k];
            }
            if (++iters >= 20 || sqrt(dn) / sqrt(pn) < (double)FLT_EPSILON) break;
            prevErr2 = err2;
        }
    }

    // --- a6 gate input: mean L2 reprojection error of the refined pose over the inliers ----
    float mean_err = 0.f;
    {
        double R[9];
        pnpm::rodrigues_v2m(param, R);
        double s = 0;
        for (int i = lane; i < n; i += 32) {
            if (!mask[i]) continue;
            float pu, pv;
            reproj_err2(R, param + 3, v.fx, v.fy, v.cx, v.cy, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], img[2 * i],
                        img[2 * i + 1], &pu, &pv);
            float dx = pu - img[2 * i], dy = pv - img[2 * i + 1];
            s += (double)sqrtf(dx * dx + dy * dy);
        }
        s = warp_sum(s);
        mean_err = max_good > 0 ? (float)(s / max_good) : 0.f;
    }
    if (lane == 0) {
        o.ok[p] = 1;
        o.n_inliers[p] = max_good;
        for (int k = 0; k < 3; k++) { o.rvec[p * 3 + k] = param[k]; o.tvec[p * 3 + k] = param[3 + k]; }
        if (o.mean_err) o.mean_err[p] = mean_err;
    }
}

// cv2.projectPoints(obj, rvec, tvec, K, DIST=0) -> float32 pixels
__global__ void k_project_points(const float* __restrict__ obj, int n, double r0, double r1, double r2, double t0,
                                 double t1, double t2, double fx, double fy, double cx, double cy, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[3] = {r0, r1, r2}, t[3] = {t0, t1, t2}, R[9];
    pnpm::rodrigues_v2m(r, R);
    float pu, pv;
    reproj_err2(R, t, fx, fy, cx, cy, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], 0.f, 0.f, &pu, &pv);
    out[2 * i] = pu;
    out[2 * i + 1] = pv;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
int launch_pnp(nclt_ctx* c, const float* obj, const float* img, const int* n, int P, const int* P_dev, int Nmax,
               const nclt_pnp_params* prm, const PnpBuffers& buf, const double* models_override,
               unsigned char* ok, double* rvec, double* tvec, int* n_inl, unsigned char* mask, float* mean_err,
               int* best_iter, int* niters, bool score_only) {
    if (P <= 0) return NCLT_OK;
    // 48.5 KB of dynamic shared memory per 32-lane block (4 blocks per SM); per device, so set on every call
    CU_TRY(c, cudaFuncSetAttribute(k_pnp_hypo, cudaFuncAttributeMaxDynamicSharedMemorySize, pnpm::EPNP5_SM_BYTES_PER_WARP));
    PnpView v;
    v.obj = obj; v.img = img; v.n = n; v.P = P; v.P_dev = P_dev; v.Nmax = Nmax;
    v.fx = prm->fx; v.fy = prm->fy; v.cx = prm->cx; v.cy = prm->cy;
    v.iters = prm->iterations;
    v.thr2 = prm->reproj_error * prm->reproj_error;
    v.conf = prm->confidence;
    v.refine = prm->refine;
    const double* models = models_override ? models_override : buf.models;
    if (score_only) {
        // staged parity (ii): score every caller-supplied hypothesis
        long long threads = (long long)P * v.iters * 32;
        k_pnp_score<<<(unsigned)((threads + PNP_SCORE_THREADS - 1) / PNP_SCORE_THREADS), PNP_SCORE_THREADS, 0, c->stream>>>(v, models, buf.counts, 0, v.iters, nullptr);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        return NCLT_OK;
    }
    // counts of iterations that are never reached stay -1 (debug output)
    CU_TRY(c, cudaMemsetAsync(buf.counts, 0xFF, (size_t)P * v.iters * sizeof(int), c->stream));
    k_pnp_sets<<<(P + 3) / 4, 128, 0, c->stream>>>(v, buf.sets, c->d_mwc, NCLT_MWC_N);
    k_pnp_state_init<<<(P + 127) / 128, 128, 0, c->stream>>>(P, v.iters, buf.state);
    c->launches += 2;
    // growing rounds of 32, 64, then 128 hypotheses (200 iterations = 3 rounds): OpenCV's loop stops at iteration niters,
    // which collapses after the first good model, and hypotheses past it are never evaluated - later rounds exit at
    // once for the easy problems, and the hard ones (wrong candidates of a production tick run all 200) pay one launch
    // less than with equal rounds of 64.  Measured per 512-problem replay step: 32/64/128 1.11 ms, 64/64/64/8 1.09 ms,
    // 8/24/64/128 1.62 ms (a round costs at least one EPnP latency, ~0.5 ms).
    for (int lo = 0, round = 0; lo < v.iters; ++round) {
        const int size = round == 0 ? 32 : round == 1 ? 64 : 128;
        const int cnt = v.iters - lo < size ? v.iters - lo : size;
        const int total = P * cnt;
        nclt_prof_mark_tag(c, 3);
        k_pnp_hypo<<<(total + 31) / 32, 32, pnpm::EPNP5_SM_BYTES_PER_WARP, c->stream>>>(v, buf.sets, buf.models, lo, cnt, buf.state);
        nclt_prof_mark_tag(c, 3);
        long long threads = (long long)total * 32;
        nclt_prof_mark_tag(c, 4);
        k_pnp_score<<<(unsigned)((threads + PNP_SCORE_THREADS - 1) / PNP_SCORE_THREADS), PNP_SCORE_THREADS, 0, c->stream>>>(v, models, buf.counts, lo, cnt, buf.state);
        nclt_prof_mark_tag(c, 4);
        k_pnp_replay<<<(P + 127) / 128, 128, 0, c->stream>>>(v, buf.counts, buf.state, lo + cnt);
        c->launches += 3;
        lo += cnt;
    }
    {
        PnpOut o{ok, rvec, tvec, n_inl, mask, mean_err, best_iter, niters};
        long long threads = (long long)P * 32;
        nclt_prof_mark_tag(c, 5);
        k_pnp_finish<<<(unsigned)((threads + 31) / 32), 32, 0, c->stream>>>(v, models, buf.state, o);
        nclt_prof_mark_tag(c, 5);
        c->launches++;
    }
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

int launch_project_points(nclt_ctx* c, const float* obj, int n, const double* rvec, const double* tvec, double fx,
                          double fy, double cx, double cy, float* out) {
    if (n <= 0) return NCLT_OK;
    k_project_points<<<(n + 255) / 256, 256, 0, c->stream>>>(obj, n, rvec[0], rvec[1], rvec[2], tvec[0], tvec[1],
                                                             tvec[2], fx, fy, cx, cy, out);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}
