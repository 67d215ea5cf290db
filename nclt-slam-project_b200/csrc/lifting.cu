// Teach-time keypoint lifting (scripts/common/visual_landmark_recorder.py:247-291): ORB keypoints +
// the aligned 16-bit depth image -> the kept keypoint indices and their optical-frame 3-D points, i.e. the
// `keypoints_3d_cam` / `keypoints_2d` / `descriptors` rows of one landmarks.pkl record.
//
// One CTA per frame, one thread per keypoint, ordered block compaction (the record keeps ORB's keypoint order).
// Every floating-point operation is the reference's, in its order and precision: np.round (half to even), float32
// depth / 1000, the float32 std of the non-zero 3x3 neighbourhood with NumPy's pairwise summation order
// (oracle/lifting.py::np_sum_f32), float32 gates against float32 constants, float64 back-projection rounded to float32.
// HBM bound by construction: 9 depth samples (sector-scattered) + 8 B per keypoint in, <= 16 B out.
#include "common.cuh"
#include "scratch.cuh"

namespace {

// NumPy float32 add-reduction of n <= 9 values: sequential below 8, else 8 accumulators + remainder
__device__ __forceinline__ float np_sum9(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), __fadd_rn(a[2], a[3])),
                          __fadd_rn(__fadd_rn(a[4], a[5]), __fadd_rn(a[6], a[7])));
    for (int i = 8; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

struct LiftView {
    const uint16_t* depth;   // [F][H][W] millimetres
    const float* kpts;       // [F][Nmax][2] (x, y)
    const int* n_kpts;       // [F]
    int H, W, Nmax;
    double fx, fy, cx, cy;
    int ground_y;
    float dmin, dmax, std_max;
    int* out_keep;           // [F][Nmax] indices into the frame's keypoints, ascending
    float* out_pts3d;        // [F][Nmax][3]
    int* out_n;              // [F]
};

__global__ void __launch_bounds__(256) k_lift_keypoints(LiftView v) {
    const int f = blockIdx.x;
    const int n = min(v.n_kpts[f], v.Nmax);
    const uint16_t* depth = v.depth + (size_t)f * v.H * v.W;
    const float* kp = v.kpts + (size_t)f * v.Nmax * 2;
    int* keep = v.out_keep + (size_t)f * v.Nmax;
    float* pts = v.out_pts3d + (size_t)f * v.Nmax * 3;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        bool ok = false;
        float X = 0.f, Y = 0.f, Z = 0.f;
        if (i < n) {
            const float fxk = kp[2 * i], fyk = kp[2 * i + 1];
            const float ru = rintf(fxk), rv = rintf(fyk);                      // np.round: half to even (recorder:250-251)
            if (fabsf(ru) < 1e9f && fabsf(rv) < 1e9f) {                        // also rejects NaN / inf
                const int u = (int)ru, w = (int)rv;
                if (u >= 1 && u < v.W - 1 && w >= 1 && w < v.H - 1 && w > v.ground_y) {      // recorder:252-253
                    const float d = __fdiv_rn((float)depth[(size_t)w * v.W + u], 1000.0f);   // recorder:260
                    float vals[9];
                    int cnt = 0;
#pragma unroll
                    for (int dv = -1; dv <= 1; ++dv)
#pragma unroll
                        for (int du = -1; du <= 1; ++du) {                      // recorder:264-267, row-major patch
                            const float p = __fdiv_rn((float)depth[(size_t)(w + dv) * v.W + (u + du)], 1000.0f);
                            if (p > 0.01f) vals[cnt++] = p;
                        }
                    float sd = 999.0f;
                    if (cnt >= 3) {                                             // ndarray.std(), float32, ddof 0
                        const float mean = __fdiv_rn(np_sum9(vals, cnt), (float)cnt);
                        float sq[9];
                        for (int k = 0; k < cnt; ++k) {
                            const float e = __fsub_rn(vals[k], mean);
                            sq[k] = __fmul_rn(e, e);
                        }
                        sd = __fsqrt_rn(__fdiv_rn(np_sum9(sq, cnt), (float)cnt));
                    }
                    if (d > v.dmin && d < v.dmax && sd < v.std_max) {           // recorder:268-270
                        ok = true;
                        X = (float)(__ddiv_rn(__dmul_rn((double)u - v.cx, (double)d), v.fx));     // recorder:283-286
                        Y = (float)(__ddiv_rn(__dmul_rn((double)w - v.cy, (double)d), v.fy));
                        Z = d;
                    }
                }
            }
        }
        // ordered compaction of the 256 flags
        const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) {
            const int c = s_warp[w2];
            if (w2 < warp) before += c;
            total += c;
        }
        const int base = s_base;
        if (ok) {
            const int o = base + before + __popc(m & ((1u << lane) - 1u));
            keep[o] = i;
            pts[3 * o] = X;
            pts[3 * o + 1] = Y;
            pts[3 * o + 2] = Z;
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0) v.out_n[f] = s_base;
}

int launch_lift(nclt_ctx* c, const LiftView& v, int F) {
    if (F <= 0) return NCLT_OK;
    k_lift_keypoints<<<F, 256, 0, c->stream>>>(v);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

int check_lift_args(nclt_ctx* c, const void* depth, int F, int H, int W, const void* kpts, const void* n_kpts, int Nmax,
                    const nclt_lift_params* prm, const void* keep, const void* pts, const void* out_n) {
    if (!c) return NCLT_ERR_ARG;
    if (!prm) return nclt_fail(c, NCLT_ERR_ARG, "lift: null params");
    if (F < 0 || H < 3 || W < 3 || Nmax <= 0) return nclt_fail(c, NCLT_ERR_ARG, "lift: bad F/H/W/Nmax");
    if (F > 0 && (!depth || !kpts || !n_kpts || !keep || !pts || !out_n)) return nclt_fail(c, NCLT_ERR_ARG, "lift: null pointer");
    return NCLT_OK;
}

LiftView make_view(const uint16_t* depth, int H, int W, const float* kpts, const int* n_kpts, int Nmax,
                   const nclt_lift_params* prm, int* keep, float* pts, int* out_n) {
    LiftView v;
    v.depth = depth; v.kpts = kpts; v.n_kpts = n_kpts; v.H = H; v.W = W; v.Nmax = Nmax;
    v.fx = prm->fx; v.fy = prm->fy; v.cx = prm->cx; v.cy = prm->cy;
    v.ground_y = prm->ground_y;
    v.dmin = prm->depth_min_m; v.dmax = prm->depth_max_m; v.std_max = prm->depth_std_max_m;
    v.out_keep = keep; v.out_pts3d = pts; v.out_n = out_n;
    return v;
}

}  // namespace

extern "C" int nclt_lift_keypoints_dev(nclt_ctx* c, const uint16_t* depth_mm, int F, int H, int W, const float* kpts_xy,
                                       const int32_t* n_kpts, int Nmax, const nclt_lift_params* prm, int32_t* out_keep,
                                       float* out_pts3d, int32_t* out_n) {
    int rc = check_lift_args(c, depth_mm, F, H, W, kpts_xy, n_kpts, Nmax, prm, out_keep, out_pts3d, out_n);
    if (rc) return rc;
    cudaSetDevice(c->device);
    return launch_lift(c, make_view(depth_mm, H, W, kpts_xy, n_kpts, Nmax, prm, out_keep, out_pts3d, out_n), F);
}

extern "C" int nclt_lift_keypoints(nclt_ctx* c, const uint16_t* depth_mm, int F, int H, int W, const float* kpts_xy,
                                   const int32_t* n_kpts, int Nmax, const nclt_lift_params* prm, int32_t* out_keep,
                                   float* out_pts3d, int32_t* out_n) {
    int rc = check_lift_args(c, depth_mm, F, H, W, kpts_xy, n_kpts, Nmax, prm, out_keep, out_pts3d, out_n);
    if (rc) return rc;
    if (F == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    const size_t db = (size_t)F * H * W * 2, kb = (size_t)F * Nmax * 8, nb = (size_t)F * 4;
    const size_t keepb = (size_t)F * Nmax * 4, ptsb = (size_t)F * Nmax * 12;
    ScratchScope scope(c);
    if ((rc = nclt_scratch_reserve(c, pad256(db) + pad256(kb) + 2 * pad256(nb) + pad256(keepb) + pad256(ptsb) + 1536))) return rc;
    Carver cv(c);
    uint16_t* d_depth = cv.take<uint16_t>((size_t)F * H * W);
    float* d_k = cv.take<float>((size_t)F * Nmax * 2);
    int* d_n = cv.take<int>((size_t)F);
    int* d_keep = cv.take<int>((size_t)F * Nmax);
    float* d_pts = cv.take<float>((size_t)F * Nmax * 3);
    int* d_on = cv.take<int>((size_t)F);
    CU_TRY(c, cudaMemcpyAsync(d_depth, depth_mm, db, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_k, kpts_xy, kb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_n, n_kpts, nb, cudaMemcpyHostToDevice, c->stream));
    rc = launch_lift(c, make_view(d_depth, H, W, d_k, d_n, Nmax, prm, d_keep, d_pts, d_on), F);
    if (rc == NCLT_OK) {
        CU_TRY(c, cudaMemcpyAsync(out_keep, d_keep, keepb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaMemcpyAsync(out_pts3d, d_pts, ptsb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaMemcpyAsync(out_n, d_on, nb, cudaMemcpyDeviceToHost, c->stream));
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return rc;
}
