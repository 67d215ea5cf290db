// The whole repeat-time hot path for a batch of frames, device resident end to end:
//   match (K1/K1'/K2) -> MIN_MATCHES gate -> gather 3D/2D correspondences (a4) -> PnP-RANSAC
//   (K3-K5) -> MIN_INLIERS / REPROJ_MAX_PX gates (a6) -> best candidate per frame (a7).
// Mirrors the loop bodies of checkpoint_a_selftest.py:62-103 (mode 0: knn2 + Lowe ratio) and
// visual_landmark_matcher.py:318-380 (mode 1: crossCheck).  The pose composition that follows
// (matcher:361-378) is 3x3 host arithmetic and stays in Python.
#include "common.cuh"
#include "scratch.cuh"

#include <algorithm>

namespace {

// items with enough matches become PnP problems (order irrelevant: problems are independent)
// `cap`: problem capacity of the buffers (asynchronous mode); what does not fit is counted in
// `overflow` (the caller re-runs such a batch synchronously - parity is never silently lost).
// A candidate keyframe with fewer than MIN_MATCHES descriptors is skipped BEFORE matching by the reference
// (`len(desc_t) < MIN_MATCHES: continue`, checkpoint_a_selftest.py:64-65, visual_landmark_matcher.py:321-322):
// in ratio mode a 2..9-row keyframe can collect >= MIN_MATCHES many-to-one matches, so the gate on the match
// count alone is not enough.  Such items (and empty slots) report 0 matches.
__global__ void k_select_problems(int* __restrict__ n_pairs, int n_items, int min_matches, const int* __restrict__ cand,
                                  int C, const int* __restrict__ kf_count, int* prob_item, int* item_prob, int* count,
                                  int cap, int* overflow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int slot = -1;
    const int kf = cand ? cand[i] : i % C;
    if (kf < 0 || kf_count[kf] < min_matches) n_pairs[i] = 0;
    if (kf >= 0 && n_pairs[i] >= min_matches) {
        slot = atomicAdd(count, 1);
        if (slot < cap) prob_item[slot] = i;
        else { slot = -1; if (overflow) atomicAdd(overflow, 1); }
    }
    item_prob[i] = slot;
}

// ---- pruned candidate loop (candidate lists, no per-item outputs requested) -------------------------------------------
// The reference tries every candidate of a tick in turn and keeps the FIRST one with the MOST inliers among those that pass
// the gates (matcher:318-380).  A candidate cannot have more inliers than matches, so once one candidate has been accepted
// with X inliers, a candidate with fewer than X matches - or with exactly X from a later slot - cannot become the result
// and its 200 RANSAC iterations need not run.  Pass A solves, per frame, the eligible candidate with the most matches
// (lowest slot on ties); pass B whatever can still win.  Same best candidate / pose / inliers / error as the full loop
// (tests/test_localize_gpu.py::test_pruned_candidate_loop_equals_full_loop); four of five candidates of a production
// tick are wrong keyframes that run all 200 iterations, so this is most of the tick's PnP work.
__global__ void k_select_top(int* __restrict__ n_pairs, int B, int C, int min_matches, const int* __restrict__ cand,
                             const int* __restrict__ kf_count, int* prob_item, int* item_prob, int* top_c, int* count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int best = -1, bn = -1;
    for (int c = 0; c < C; ++c) {
        const int i = b * C + c, kf = cand[i];
        if (kf < 0 || kf_count[kf] < min_matches) n_pairs[i] = 0;
        item_prob[i] = -1;
        const int np = n_pairs[i];
        if (kf >= 0 && np >= min_matches && np > bn) { bn = np; best = c; }
    }
    top_c[b] = best;
    if (best >= 0) {
        const int slot = atomicAdd(count, 1);        // <= B: always fits
        prob_item[slot] = b * C + best;
        item_prob[b * C + best] = slot;
    }
}

__global__ void k_select_rest(const int* __restrict__ n_pairs, int B, int C, int min_matches, const int* __restrict__ cand,
                              const int* __restrict__ top_c, const unsigned char* __restrict__ ok, const int* __restrict__ n_inl,
                              const float* __restrict__ mean_err, int min_inliers, float reproj_max, int base, int* prob_item,
                              int* item_prob, int* count, int cap, int* overflow) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int ca = top_c[b];
    if (ca < 0) return;
    const int pa = item_prob[b * C + ca];
    int X = -1;                                      // inliers of the accepted pass-A result, -1: not accepted
    if (ok[pa] && n_inl[pa] >= min_inliers && mean_err[pa] <= reproj_max) X = n_inl[pa];
    for (int c = 0; c < C; ++c) {
        if (c == ca) continue;
        const int i = b * C + c, np = n_pairs[i];    // 0 for empty slots and short keyframes (k_select_top)
        if (cand[i] < 0 || np < min_matches) continue;
        if (np > X || (np == X && c < ca)) {
            const int slot = atomicAdd(count, 1);
            if (slot < cap) { prob_item[slot] = i; item_prob[i] = base + slot; }
            else if (overflow) atomicAdd(overflow, 1);
        }
    }
}

#ifndef NCLT_CORESIDENT
#define NCLT_CORESIDENT 0
#endif
#define GATHER_THREADS (NCLT_CORESIDENT ? 128 : 256)
// a4: obj_pts = keypoints_3d_cam[teach row], img_pts = pts_curr_2d[frame row]
__global__ void __launch_bounds__(256) k_gather_problems(const int* __restrict__ prob_item, const int2* __restrict__ pairs,
                                                         const int* __restrict__ n_pairs, int pair_stride, int mode,
                                                         const int* __restrict__ cand, int C,
                                                         const int* __restrict__ kf_start,
                                                         const float* __restrict__ pts3d,
                                                         const float* __restrict__ pts2d, int Nq, float* obj,
                                                         float* img, int* n_out, int Nmax,
                                                         const int* __restrict__ count) {
    const int p = blockIdx.x;
    if (p >= *count) return;
    const int item = prob_item[p];
    const int b = item / C;
    const int kf = cand ? cand[item] : item % C;
    const int n = n_pairs[item];
    const int base3 = kf_start[kf];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int2 pr = pairs[(size_t)item * pair_stride + i];
        int frame_row = mode == 0 ? pr.x : pr.y;
        int teach_row = mode == 0 ? pr.y : pr.x;
        const float* s3 = pts3d + (size_t)(base3 + teach_row) * 3;
        const float* s2 = pts2d + ((size_t)b * Nq + frame_row) * 2;
        float* o = obj + ((size_t)p * Nmax + i) * 3;
        float* im = img + ((size_t)p * Nmax + i) * 2;
        o[0] = s3[0]; o[1] = s3[1]; o[2] = s3[2];
        im[0] = s2[0]; im[1] = s2[1];
    }
    if (threadIdx.x == 0) n_out[p] = n;
}

// a6 + a7: gates and "first candidate with the most inliers wins" (matcher:349-359,379-380).  One warp per frame: the
// lanes stride over the C candidates (400 in the all-keyframes replay: a single thread walking them took 63 us per
// 512 frames), then a shuffle reduction on (most inliers, earliest slot).
__global__ void __launch_bounds__(128) k_reduce_frames(const int* __restrict__ item_prob, int B, int C,
                                                       const unsigned char* __restrict__ ok, const int* __restrict__ n_inl,
                                                       const float* __restrict__ mean_err, const double* __restrict__ rvec,
                                                       const double* __restrict__ tvec, int min_inliers, float reproj_max,
                                                       int* best_cand, int* best_inl, float* best_err, double* best_rvec,
                                                       double* best_tvec) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    int bc = -1, bi = 0, bp = -1;
    for (int c = lane; c < C; c += 32) {             // a lane sees its slots in increasing order: strict > keeps the earliest
        const int p = item_prob[b * C + c];
        if (p < 0 || !ok[p]) continue;
        const int ni = n_inl[p];
        if (ni < min_inliers) continue;
        if (mean_err[p] > reproj_max) continue;
        if (bc < 0 || ni > bi) { bc = c; bi = ni; bp = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int oc = __shfl_xor_sync(0xFFFFFFFFu, bc, o), oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o),
                  op = __shfl_xor_sync(0xFFFFFFFFu, bp, o);
        if (oc >= 0 && (bc < 0 || oi > bi || (oi == bi && oc < bc))) { bc = oc; bi = oi; bp = op; }
    }
    if (lane == 0) {
        best_cand[b] = bc;
        best_inl[b] = bi;
        best_err[b] = bp >= 0 ? mean_err[bp] : 0.f;
    }
    if (lane < 3) {
        best_rvec[b * 3 + lane] = bp >= 0 ? rvec[bp * 3 + lane] : 0.0;
        best_tvec[b * 3 + lane] = bp >= 0 ? tvec[bp * 3 + lane] : 0.0;
    }
}

// optional per-item view of the PnP results (parity tests / CSV logging)
__global__ void k_scatter_items(const int* __restrict__ item_prob, int n_items, const unsigned char* __restrict__ ok,
                                const int* __restrict__ n_inl, const float* __restrict__ mean_err,
                                const double* __restrict__ rvec, const double* __restrict__ tvec,
                                unsigned char* it_ok, int* it_inl, float* it_err, double* it_rvec, double* it_tvec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int p = item_prob[i];
    bool has = p >= 0;
    if (it_ok) it_ok[i] = has ? ok[p] : 0;
    if (it_inl) it_inl[i] = has ? n_inl[p] : 0;
    if (it_err) it_err[i] = has ? mean_err[p] : 0.f;
    for (int k = 0; k < 3; k++) {
        if (it_rvec) it_rvec[i * 3 + k] = has ? rvec[p * 3 + k] : 0.0;
        if (it_tvec) it_tvec[i * 3 + k] = has ? tvec[p * 3 + k] : 0.0;
    }
}

}  // namespace

extern "C" int nclt_localize_batch_dev(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const float* q_pts2d,
                                       const int32_t* q_n, int B, int Nq, const int32_t* cand, int C,
                                       const nclt_localize_params* prm, int32_t* out_best_cand,
                                       int32_t* out_n_inliers, float* out_reproj, double* out_rvec,
                                       double* out_tvec, int32_t* out_n_problems_host, int32_t* out_item_nmatch,
                                       uint8_t* out_item_ok, int32_t* out_item_ninl, float* out_item_err,
                                       double* out_item_rvec, double* out_item_tvec) {
    if (!c || !L || !prm) return nclt_fail(c, NCLT_ERR_ARG, "localize: null handle/params");
    if (B < 0 || Nq <= 0 || C <= 0) return nclt_fail(c, NCLT_ERR_ARG, "localize: bad B/Nq/C");
    if (B > 0 && (!q || !q_pts2d)) return nclt_fail(c, NCLT_ERR_ARG, "localize: null q/q_pts2d");
    if (!out_best_cand || !out_n_inliers || !out_reproj || !out_rvec || !out_tvec)
        return nclt_fail(c, NCLT_ERR_ARG, "localize: null output");
    if (prm->mode != 0 && prm->mode != 1) return nclt_fail(c, NCLT_ERR_ARG, "localize: mode must be 0 or 1");
    if (out_n_problems_host) *out_n_problems_host = 0;
    if (B == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    ScratchScope scope(c);
    const size_t items = (size_t)B * C;
    // pair rows per item: frame rows in ratio mode, teach rows in crossCheck mode
    const int Nrow = prm->mode == 0 ? Nq : (L->max_count > 0 ? L->max_count : 1);
    int rc;
    // stage 1 scratch: pairs + counts + problem tables (the match entry points reserve their own)
    size_t need1 = pad256(items * Nrow * 8) + 4 * pad256(items * 4) + pad256((size_t)B * 4) + 4096;
    if ((rc = nclt_scratch_reserve(c, need1))) return rc;
    Carver cv(c);
    int2* pairs = cv.take<int2>(items * Nrow);
    int* n_pairs = out_item_nmatch ? out_item_nmatch : cv.take<int>(items);
    int* prob_item = cv.take<int>(items);
    int* item_prob = cv.take<int>(items);
    int* d_count = cv.take<int>(1);
    if (prm->mode == 0) {
        rc = nclt_match_ratio_dev(c, L, q, q_n, B, Nq, cand, C, prm->ratio_num, prm->ratio_den,
                                  reinterpret_cast<int32_t*>(pairs), n_pairs);
    } else {
        rc = nclt_match_cross_dev(c, L, q, q_n, B, Nq, cand, C, Nrow, reinterpret_cast<int32_t*>(pairs), nullptr,
                                  n_pairs);
    }
    if (rc) return rc;
    if (out_n_problems_host && prm->mode == 0 && c->engine == 2 && !cand && nclt_overflow_take(c) > 0) {
        // synchronous call and the fused fp4 matcher ran out of candidate capacity: redo the matching on the integer
        // engine (same results).  The asynchronous mode only counts (nclt_ctx_overflow), as documented.
        c->engine = 0;
        rc = nclt_match_ratio_dev(c, L, q, q_n, B, Nq, cand, C, prm->ratio_num, prm->ratio_den,
                                  reinterpret_cast<int32_t*>(pairs), n_pairs);
        c->engine = 2;
        if (rc) return rc;
    }
    const bool async_mode = out_n_problems_host == nullptr;
    if (cand && C > 1 && !out_item_ok && !out_item_ninl && !out_item_err && !out_item_rvec && !out_item_tvec) {
        // ---- pruned candidate loop: pass A (<= B problems, capacity known: no read-back), pass B (what can still win)
        const int PA = B;
        const int capB = (int)std::min<size_t>(items - (size_t)B, std::max<size_t>((size_t)B * 4, 1024));
        const int iters = prm->pnp.iterations;
        int* prob_item2 = cv.take<int>(items);
        int* top_c = cv.take<int>(B);
        int* d_count2 = cv.take<int>(1);
        CU_TRY(c, cudaMemsetAsync(d_count, 0, 4, c->stream));
        CU_TRY(c, cudaMemsetAsync(d_count2, 0, 4, c->stream));
        k_select_top<<<(B + 127) / 128, 128, 0, c->stream>>>(n_pairs, B, C, prm->min_matches, cand, L->d_count, prob_item, item_prob,
                                                            top_c, d_count);
        c->launches++;
        const int Pm = std::max(PA, capB), Pt = PA + capB;
        size_t h = (size_t)Pm * iters;
        size_t need2 = pad256((size_t)Pm * Nrow * 12) + pad256((size_t)Pm * Nrow * 8) + pad256((size_t)Pm * 4) +
                       pad256((size_t)Pm * Nrow) + pad256((size_t)Pt) + 2 * pad256((size_t)Pt * 24) + 2 * pad256((size_t)Pt * 4) +
                       pad256(h * 20) + pad256(h * 48) + pad256(h * 4) + pad256((size_t)Pm * 16);
        if ((rc = nclt_scratch_reserve(c, need2))) return rc;   // never moves what is already carved
        Carver cv2(c);
        float* obj = cv2.take<float>((size_t)Pm * Nrow * 3);
        float* img = cv2.take<float>((size_t)Pm * Nrow * 2);
        int* pn = cv2.take<int>(Pm);
        unsigned char* mask = cv2.take<unsigned char>((size_t)Pm * Nrow);
        unsigned char* p_ok = cv2.take<unsigned char>(Pt);
        double* p_r = cv2.take<double>((size_t)Pt * 3);
        double* p_t = cv2.take<double>((size_t)Pt * 3);
        int* p_inl = cv2.take<int>(Pt);
        float* p_err = cv2.take<float>(Pt);
        PnpBuffers buf;
        buf.sets = cv2.take<int>(h * 5);
        buf.models = cv2.take<double>(h * 6);
        buf.counts = cv2.take<int>(h);
        buf.state = cv2.take<int>((size_t)Pm * 4);
        k_gather_problems<<<PA, GATHER_THREADS, 0, c->stream>>>(prob_item, pairs, n_pairs, Nrow, prm->mode, cand, C, L->d_start, L->d_pts3d,
                                                     q_pts2d, Nq, obj, img, pn, Nrow, d_count);
        c->launches++;
        if ((rc = launch_pnp(c, obj, img, pn, PA, d_count, Nrow, &prm->pnp, buf, nullptr, p_ok, p_r, p_t, p_inl, mask, p_err,
                             nullptr, nullptr, false)))
            return rc;
        k_select_rest<<<(B + 127) / 128, 128, 0, c->stream>>>(n_pairs, B, C, prm->min_matches, cand, top_c, p_ok, p_inl, p_err,
                                                             prm->min_inliers, prm->reproj_max_px, PA, prob_item2, item_prob,
                                                             d_count2, capB, async_mode ? c->d_overflow : nullptr);
        c->launches++;
        int PB = capB;
        if (!async_mode) {
            if ((rc = nclt_pinned_reserve(c, 64))) return rc;
            int* hp = static_cast<int*>(c->pinned);
            CU_TRY(c, cudaMemcpyAsync(hp, d_count, 4, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(c, cudaMemcpyAsync(hp + 1, d_count2, 4, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            PB = std::min(hp[1], capB);
            if (hp[1] > capB) return nclt_fail(c, NCLT_ERR_STATE, "localize: pass-B problem capacity exceeded");
            *out_n_problems_host = hp[0] + PB;
        }
        if (PB > 0) {
            k_gather_problems<<<PB, GATHER_THREADS, 0, c->stream>>>(prob_item2, pairs, n_pairs, Nrow, prm->mode, cand, C, L->d_start,
                                                         L->d_pts3d, q_pts2d, Nq, obj, img, pn, Nrow, d_count2);
            c->launches++;
            if ((rc = launch_pnp(c, obj, img, pn, PB, d_count2, Nrow, &prm->pnp, buf, nullptr, p_ok + PA, p_r + 3 * (size_t)PA,
                                 p_t + 3 * (size_t)PA, p_inl + PA, mask, p_err + PA, nullptr, nullptr, false)))
                return rc;
        }
        k_reduce_frames<<<(B + 3) / 4, 128, 0, c->stream>>>(item_prob, B, C, p_ok, p_inl, p_err, p_r, p_t, prm->min_inliers,
                                                                prm->reproj_max_px, out_best_cand, out_n_inliers, out_reproj,
                                                                out_rvec, out_tvec);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        return NCLT_OK;
    }
    CU_TRY(c, cudaMemsetAsync(d_count, 0, 4, c->stream));
    // The number of PnP problems is data dependent.  Synchronous mode (out_n_problems given): one
    // 4-byte read-back, buffers sized exactly.  Asynchronous mode (out_n_problems == NULL): no host
    // sync at all - buffers sized for `cap` problems, kernels read the count on the device, and a
    // per-context overflow counter (nclt_ctx_overflow) reports batches that needed more.
    int P = 0;
    if (async_mode) {
        P = (int)std::min<size_t>(items, std::max<size_t>((size_t)B * 4, 1024));
    }
    k_select_problems<<<(unsigned)((items + 255) / 256), 256, 0, c->stream>>>(
        n_pairs, (int)items, prm->min_matches, cand, C, L->d_count, prob_item, item_prob, d_count,
        async_mode ? P : (int)items, async_mode ? c->d_overflow : nullptr);
    c->launches++;
    if (!async_mode) {
        if ((rc = nclt_pinned_reserve(c, 64))) return rc;
        CU_TRY(c, cudaMemcpyAsync(c->pinned, d_count, 4, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        P = *static_cast<int*>(c->pinned);
        *out_n_problems_host = P;
    }

    unsigned char* p_ok = nullptr;
    int* p_inl = nullptr;
    float* p_err = nullptr;
    double *p_r = nullptr, *p_t = nullptr;
    if (P > 0) {
        const int iters = prm->pnp.iterations;
        size_t h = (size_t)P * iters;
        size_t need2 = pad256((size_t)P * Nrow * 12) + pad256((size_t)P * Nrow * 8) + 3 * pad256((size_t)P * 4) +
                       pad256((size_t)P * Nrow) + pad256((size_t)P) + 2 * pad256((size_t)P * 24) +
                       pad256(h * 20) + pad256(h * 48) + pad256(h * 4) + pad256((size_t)P * 16);
        if ((rc = nclt_scratch_reserve(c, need2))) return rc;   // never moves what is already carved
        Carver cv2(c);
        float* obj = cv2.take<float>((size_t)P * Nrow * 3);
        float* img = cv2.take<float>((size_t)P * Nrow * 2);
        int* pn = cv2.take<int>(P);
        unsigned char* mask = cv2.take<unsigned char>((size_t)P * Nrow);
        p_ok = cv2.take<unsigned char>(P);
        p_r = cv2.take<double>((size_t)P * 3);
        p_t = cv2.take<double>((size_t)P * 3);
        p_inl = cv2.take<int>(P);
        p_err = cv2.take<float>(P);
        PnpBuffers buf;
        buf.sets = cv2.take<int>(h * 5);
        buf.models = cv2.take<double>(h * 6);
        buf.counts = cv2.take<int>(h);
        buf.state = cv2.take<int>((size_t)P * 4);
        k_gather_problems<<<P, GATHER_THREADS, 0, c->stream>>>(prob_item, pairs, n_pairs, Nrow, prm->mode, cand, C, L->d_start,
                                                    L->d_pts3d, q_pts2d, Nq, obj, img, pn, Nrow, d_count);
        c->launches++;
        if ((rc = launch_pnp(c, obj, img, pn, P, async_mode ? d_count : nullptr, Nrow, &prm->pnp, buf, nullptr, p_ok,
                             p_r, p_t, p_inl, mask, p_err, nullptr, nullptr, false)))
            return rc;
    }
    k_reduce_frames<<<(B + 3) / 4, 128, 0, c->stream>>>(item_prob, B, C, p_ok, p_inl, p_err, p_r, p_t,
                                                            prm->min_inliers, prm->reproj_max_px, out_best_cand,
                                                            out_n_inliers, out_reproj, out_rvec, out_tvec);
    c->launches++;
    if (out_item_ok || out_item_ninl || out_item_err || out_item_rvec || out_item_tvec) {
        k_scatter_items<<<(unsigned)((items + 255) / 256), 256, 0, c->stream>>>(
            item_prob, (int)items, p_ok, p_inl, p_err, p_r, p_t, out_item_ok, out_item_ninl, out_item_err,
            out_item_rvec, out_item_tvec);
        c->launches++;
    }
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

extern "C" int nclt_localize_batch(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const float* q_pts2d,
                                   const int32_t* q_n, int B, int Nq, const int32_t* cand, int C,
                                   const nclt_localize_params* prm, int32_t* out_best_cand, int32_t* out_n_inliers,
                                   float* out_reproj, double* out_rvec, double* out_tvec, int32_t* out_n_problems,
                                   int32_t* out_item_nmatch, uint8_t* out_item_ok, int32_t* out_item_ninl,
                                   float* out_item_err, double* out_item_rvec, double* out_item_tvec) {
    if (!c || !L || !prm) return nclt_fail(c, NCLT_ERR_ARG, "localize: null handle/params");
    if (B < 0 || Nq <= 0 || C <= 0) return nclt_fail(c, NCLT_ERR_ARG, "localize: bad B/Nq/C");
    if (B > 0 && (!q || !q_pts2d)) return nclt_fail(c, NCLT_ERR_ARG, "localize: null q/q_pts2d");
    if (!out_best_cand || !out_n_inliers || !out_reproj || !out_rvec || !out_tvec)
        return nclt_fail(c, NCLT_ERR_ARG, "localize: null output");
    if (B == 0) { if (out_n_problems) *out_n_problems = 0; return NCLT_OK; }
    {
        int rc0 = nclt_check_host_lists(c, L, q_n, B, Nq, cand, C);      // host arrays: reject bad ids before they index device tables
        if (rc0) return rc0;
        if (!cand && C > L->n_kf) return nclt_fail(c, NCLT_ERR_ARG, "localize: C exceeds keyframe count with cand == NULL");
    }
    cudaSetDevice(c->device);
    // host staging lives in its own allocations so that the device pipeline can grow scratch
    const size_t items = (size_t)B * C;
    size_t in_bytes = pad256((size_t)B * Nq * 32) + pad256((size_t)B * Nq * 8) + pad256((size_t)B * 4) + pad256(items * 4);
    size_t out_bytes = 2 * pad256((size_t)B * 4) + pad256((size_t)B * 4) + 2 * pad256((size_t)B * 24) +
                       2 * pad256(items * 4) + pad256(items) + pad256(items * 4) + 2 * pad256(items * 24);
    // staging comes from the context scratch (chunked: the nested device call can still grow it);
    // no cudaMalloc/cudaFree per call - both synchronise the device
    ScratchScope scope(c);
    {
        int rc0 = nclt_scratch_reserve(c, in_bytes + out_bytes + 8192);
        if (rc0) return rc0;
    }
    Carver cvs(c);
    char* stage = cvs.take<char>(in_bytes + out_bytes + 4096);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = stage + off; off += pad256(bytes); return p; };
    uint8_t* d_q = (uint8_t*)take((size_t)B * Nq * 32);
    float* d_p2 = (float*)take((size_t)B * Nq * 8);
    int* d_qn = q_n ? (int*)take((size_t)B * 4) : nullptr;
    int* d_cand = cand ? (int*)take(items * 4) : nullptr;
    int* d_bc = (int*)take((size_t)B * 4);
    int* d_bi = (int*)take((size_t)B * 4);
    float* d_be = (float*)take((size_t)B * 4);
    double* d_br = (double*)take((size_t)B * 24);
    double* d_bt = (double*)take((size_t)B * 24);
    int* d_inm = out_item_nmatch ? (int*)take(items * 4) : nullptr;
    uint8_t* d_iok = out_item_ok ? (uint8_t*)take(items) : nullptr;
    int* d_iin = out_item_ninl ? (int*)take(items * 4) : nullptr;
    float* d_ier = out_item_err ? (float*)take(items * 4) : nullptr;
    double* d_irv = out_item_rvec ? (double*)take(items * 24) : nullptr;
    double* d_itv = out_item_tvec ? (double*)take(items * 24) : nullptr;
    cudaStream_t s = c->stream;
    int rc = NCLT_OK;
#define LOC_TRY(call)                                                       \
    do {                                                                    \
        cudaError_t _e = (call);                                            \
        if (_e != cudaSuccess) { rc = nclt_fail(c, NCLT_ERR_CUDA, #call, _e); goto done; } \
    } while (0)
    LOC_TRY(cudaMemcpyAsync(d_q, q, (size_t)B * Nq * 32, cudaMemcpyHostToDevice, s));
    LOC_TRY(cudaMemcpyAsync(d_p2, q_pts2d, (size_t)B * Nq * 8, cudaMemcpyHostToDevice, s));
    if (q_n) LOC_TRY(cudaMemcpyAsync(d_qn, q_n, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    if (cand) LOC_TRY(cudaMemcpyAsync(d_cand, cand, items * 4, cudaMemcpyHostToDevice, s));
    rc = nclt_localize_batch_dev(c, L, d_q, d_p2, d_qn, B, Nq, d_cand, C, prm, d_bc, d_bi, d_be, d_br, d_bt,
                                 out_n_problems, d_inm, d_iok, d_iin, d_ier, d_irv, d_itv);
    if (rc) goto done;
    LOC_TRY(cudaMemcpyAsync(out_best_cand, d_bc, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    LOC_TRY(cudaMemcpyAsync(out_n_inliers, d_bi, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    LOC_TRY(cudaMemcpyAsync(out_reproj, d_be, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    LOC_TRY(cudaMemcpyAsync(out_rvec, d_br, (size_t)B * 24, cudaMemcpyDeviceToHost, s));
    LOC_TRY(cudaMemcpyAsync(out_tvec, d_bt, (size_t)B * 24, cudaMemcpyDeviceToHost, s));
    if (d_inm) LOC_TRY(cudaMemcpyAsync(out_item_nmatch, d_inm, items * 4, cudaMemcpyDeviceToHost, s));
    if (d_iok) LOC_TRY(cudaMemcpyAsync(out_item_ok, d_iok, items, cudaMemcpyDeviceToHost, s));
    if (d_iin) LOC_TRY(cudaMemcpyAsync(out_item_ninl, d_iin, items * 4, cudaMemcpyDeviceToHost, s));
    if (d_ier) LOC_TRY(cudaMemcpyAsync(out_item_err, d_ier, items * 4, cudaMemcpyDeviceToHost, s));
    if (d_irv) LOC_TRY(cudaMemcpyAsync(out_item_rvec, d_irv, items * 24, cudaMemcpyDeviceToHost, s));
    if (d_itv) LOC_TRY(cudaMemcpyAsync(out_item_tvec, d_itv, items * 24, cudaMemcpyDeviceToHost, s));
    // out_n_problems == NULL: asynchronous - everything is enqueued on the context stream (input copies, the
    // device pipeline with its problem count kept on the device, result copies); the caller keeps its (pinned)
    // host buffers alive and calls nclt_ctx_sync() before reading them.  The staging scratch released below is
    // only re-used by later calls on the same stream, i.e. after this batch.
    if (out_n_problems) LOC_TRY(cudaStreamSynchronize(s));
done:
    if (out_n_problems || rc) cudaStreamSynchronize(s);
    return rc;
#undef LOC_TRY
}
