/*
 * nclt_b200.h - C ABI of the B200 (sm_100a) landmark matcher / teach-map builder.
 *
 * This is the drop-in boundary for the ONE hot path named in BASELINE.json (SURVEY.md
 * section 8).  The reference has no plugin/FFI layer for this path: the seam is three OpenCV
 * call sites and one Python class.  Every entry point below names the reference
 * interface it replaces (paths relative to /root/reference/simulation/isaac).
 *
 * Conventions: every function returns 0 on success or a negative NCLT_ERR_* code; the text
 * of the last failure on a context is nclt_last_error(ctx).  Nothing throws, no allocation
 * ownership crosses the boundary except the opaque handles.  One context = one GPU + one CUDA
 * stream + one caller at a time (the reference nodes are single-threaded executors,
 * scripts/common/visual_landmark_matcher.py:513, teach_run_depth_mapper.py:256).
 * Functions without a suffix take HOST pointers and copy in/out on the context's stream;
 * *_dev variants take DEVICE pointers, enqueue on the stream and return without waiting.
 * There is no CPU fallback: nclt_ctx_create fails if no CUDA device is usable.
 */
#ifndef NCLT_B200_H
#define NCLT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCLT_OK 0
#define NCLT_ERR_CUDA (-1)
#define NCLT_ERR_ARG (-2)
#define NCLT_ERR_NOMEM (-3)
#define NCLT_ERR_STATE (-4)

typedef struct nclt_ctx nclt_ctx;
typedef struct nclt_lib nclt_lib;
typedef struct nclt_occ nclt_occ;

/* ---- context ---------------------------------------------------------------------- */
/* stream: a cudaStream_t created by the caller (e.g. torch's current stream), or NULL to let
 * the context own a non-blocking stream. */
int nclt_ctx_create(int device, void* stream, nclt_ctx** out);
int nclt_ctx_destroy(nclt_ctx* ctx);
int nclt_ctx_sync(nclt_ctx* ctx);
const char* nclt_last_error(nclt_ctx* ctx);
/* kernels launched through this context so far (bench.py's gpu_launches) */
unsigned long long nclt_ctx_launches(nclt_ctx* ctx);
/* Generation of the device allocations behind this context: bumped whenever memory that enqueued work may point
 * into is freed or moved - the scratch arena grows, a tensor-engine library image or work-split table is rebuilt
 * (different batch size, nclt_lib_append), the library itself grows.  A CUDA graph captured from *_dev calls on this
 * context holds raw pointers into that memory: record the generation at capture time and re-capture (never replay)
 * once it has changed (pipeline.py::DeviceLocalizer.capture / replay do exactly that). */
unsigned long long nclt_ctx_alloc_generation(nclt_ctx* ctx);
/* asynchronous nclt_localize_batch_dev calls (out_n_problems == NULL) size their PnP buffers for
 * max(4*B, 1024) problems per batch; returns how many problems were dropped since the last reset
 * (>= 0; synchronises the stream). A caller that sees > 0 re-runs those batches synchronously. */
int nclt_ctx_overflow(nclt_ctx* ctx, int reset);
/* matching engine for the "every frame against every keyframe" ratio mode (cand == NULL):
 * 0 = integer pipe (LOP3+POPC, the default), 1 = tcgen05 tensor cores with fp8 +-1 operands and fp16
 * accumulators in TMEM, 2 = tcgen05 block-scaled fp4 (kind::mxf4) +-1 operands with f32 accumulators
 * (the fastest); all three produce identical results (exact index recovery on the integer pipe).
 * Used by nclt_match_ratio[_dev] / nclt_localize_batch[_dev] with cand == NULL and by
 * nclt_match_flat2_dev; with engine 2 also by nclt_match_cross[_dev] with cand == NULL (crossCheck of every
 * frame against every keyframe: both directions on tcgen05 with index-carrying cells, no verification pass).
 * Candidate-list matching (cand != NULL) always uses the integer pipe; crossCheck there is ONE pass over each item's
 * distance matrix (both directions: k_hamming_cross, csrc/hamming.cu). */
int nclt_ctx_set_engine(nclt_ctx* ctx, int engine);
/* The tensor-engine matching kernel is persistent (one CTA per SM) and fills every SM it runs on.  When two contexts
 * take batches alternately, leaving n SMs free lets the short tail kernels (candidate verification, PnP-RANSAC) of one
 * batch run beside the matching kernel of the next instead of after it.  Default 0. */
int nclt_ctx_set_tail_sms(nclt_ctx* ctx, int n);
/* enable/disable CUDA-event timing of the dominant kernel (the Hamming top-2 launches) on this
 * context; nclt_ctx_profile_read synchronises, returns the summed device time and launch count
 * since the last read, and resets. Used by bench.py for the live roofline figure. */
int nclt_ctx_profile(nclt_ctx* ctx, int enable);
int nclt_ctx_profile_read(nclt_ctx* ctx, double* ms_total, int* n_launches);
/* the same per kernel family (8 entries each): 0 Hamming top-2, 1 k_occ_frame (map stage A), 2 k_occ_apply (map stage B),
 * 3 k_pnp_hypo, 4 k_pnp_score, 5 k_pnp_finish, 6 candidate verification (k_tc_verify), 7 unused; resets like the above */
int nclt_ctx_profile_read_tags(nclt_ctx* ctx, double* ms_by_tag, int* n_by_tag);
/* library ABI version, bumped on any signature change */
int nclt_abi_version(void);
/* measured POPC32 op/s of a register-only kernel: roofline denominator for the matcher */
double nclt_popc_peak(nclt_ctx* ctx, int iters, float* ms_out);

/* ---- teach library ---------------------------------------------------------------- */
/* Device copy of landmarks.pkl's per-keyframe 'descriptors' u8[n,32] and 'keypoints_3d_cam'
 * f32[n,3] (scripts/common/visual_landmark_recorder.py:290-297; loaded at
 * visual_landmark_matcher.py:179-187).  kf_offsets[n_kf+1] are row offsets into desc/pts3d. */
int nclt_lib_create(nclt_ctx* ctx, int n_kf, const int32_t* kf_offsets, const uint8_t* desc,
                    const float* pts3d, nclt_lib** out);
/* append one keyframe (visual_landmark_matcher.py:492-496 _maybe_accumulate) */
int nclt_lib_append(nclt_ctx* ctx, nclt_lib* lib, const uint8_t* desc, const float* pts3d, int n);
int nclt_lib_destroy(nclt_ctx* ctx, nclt_lib* lib);
int nclt_lib_size(const nclt_lib* lib, int* n_kf, int* n_desc, int* max_kf_rows);

/* ---- descriptor matching ---------------------------------------------------------- */
/* Common arguments: q u8[B,Nq,32] live-frame descriptors (row stride Nq per frame), q_n i32[B]
 * valid rows per frame (NULL = Nq), cand i32[B,C] keyframe ids per frame, -1 = empty slot
 * (NULL = keyframes 0..C-1 for every frame). */

/* replaces cv2.BFMatcher(NORM_HAMMING).knnMatch(desc_curr, desc_t, k=2)
 * (routes/03_south/teach/scripts/checkpoint_a_selftest.py:46,68) for every (frame, candidate).
 * out_idx i32[B,C,Nq,2] trainIdx (-1 = missing), out_dist u16[B,C,Nq,2] (65535 = missing);
 * ascending distance, ties -> lowest trainIdx. */
int nclt_match_knn2(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                    int Nq, const int32_t* cand, int C, int32_t* out_idx, uint16_t* out_dist);
int nclt_match_knn2_dev(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                        int Nq, const int32_t* cand, int C, int32_t* out_idx, uint16_t* out_dist);

/* knnMatch(k=2) + Lowe ratio `m.distance < LOWE_RATIO * n.distance` (checkpoint_a_selftest.py:71)
 * evaluated exactly as den*d1 < num*d2 (0.80 -> num=4, den=5).  out_pairs i32[B,C,Nq,2] =
 * (queryIdx, trainIdx) in increasing queryIdx, out_n i32[B,C].  A keyframe with < 2 rows yields 0. */
int nclt_match_ratio(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                     int Nq, const int32_t* cand, int C, int num, int den, int32_t* out_pairs,
                     int32_t* out_n);
int nclt_match_ratio_dev(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                         int Nq, const int32_t* cand, int C, int num, int den, int32_t* out_pairs,
                         int32_t* out_n);

/* replaces cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(desc_t, desc_curr)
 * (scripts/common/visual_landmark_matcher.py:211,327).  out_pairs i32[B,C,Nmax,2] =
 * (queryIdx = teach row, trainIdx = frame row) in increasing queryIdx, out_dist u16[B,C,Nmax],
 * out_n i32[B,C]; Nmax >= the largest candidate keyframe.  cand == NULL, C == keyframe count: every keyframe (exp 63's
 * whole-library ranking, experiments/63_global_reloc/scripts/visual_landmark_matcher.py:314-345). */
int nclt_match_cross(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                     int Nq, const int32_t* cand, int C, int Nmax, int32_t* out_pairs, uint16_t* out_dist,
                     int32_t* out_n);
int nclt_match_cross_dev(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                         int Nq, const int32_t* cand, int C, int Nmax, int32_t* out_pairs,
                         uint16_t* out_dist, int32_t* out_n);

/* BASELINE config 5 (no reference analogue; closest: experiments/63_global_reloc/scripts/
 * visual_landmark_matcher.py:314-345): flat global top-2 of every query row over ALL library
 * rows of this rank.  out_keys u32[B,Nq,2]: key = dist<<23 | (idx_offset + library row),
 * 0xFFFFFFFF = missing; keys from several ranks merge with nclt_merge_top2_dev.  With a tensor engine
 * selected (nclt_ctx_set_engine) the keyframe structure of the library is used: per-keyframe top-2 on
 * tcgen05, then an exact re-scan of the two keyframes that can hold the global top-2; same keys. */
int nclt_match_flat2_dev(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const int32_t* q_n, int B,
                         int Nq, uint32_t idx_offset, uint32_t* out_keys);
/* parts u32[nparts,B*Nq,2] -> out_keys u32[B*Nq,2], out_idx i32[B*Nq,2], out_dist u16[B*Nq,2]
 * (any output may be NULL); tie -> lowest global index. */
int nclt_merge_top2_dev(nclt_ctx* ctx, const uint32_t* parts, int nparts, int rows, uint32_t* out_keys,
                        int32_t* out_idx, uint16_t* out_dist);

/* ---- PnP-RANSAC -------------------------------------------------------------------- */
/* K = [[fx,0,cx],[0,fy,cy],[0,0,1]] (visual_landmark_matcher.py:49-51), distortion is zero
 * (matcher:52).  iterations/reproj_error/confidence = iterationsCount, reprojectionError,
 * confidence of cv2.solvePnPRansac; refine != 0 runs the SOLVEPNP_ITERATIVE refinement. */
typedef struct nclt_pnp_params {
    double fx, fy, cx, cy;
    int iterations;       /* RANSAC_ITERATIONS = 200 (matcher:69) */
    float reproj_error;   /* RANSAC_REPROJ_PX = 3.0 (matcher:68) */
    double confidence;    /* cv2 default 0.99 */
    int refine;           /* 1 = flags=SOLVEPNP_ITERATIVE */
} nclt_pnp_params;

/* replaces cv2.solvePnPRansac(obj_pts, img_pts, K, DIST, iterationsCount=200,
 * reprojectionError=3.0, flags=cv2.SOLVEPNP_ITERATIVE) (visual_landmark_matcher.py:342-346,
 * checkpoint_a_selftest.py:78-82) and the cv2.projectPoints mean-error computation of
 * matcher:353-355, for P independent problems.
 * obj f32[P,Nmax,3], img f32[P,Nmax,2], n i32[P] valid correspondences per problem (n < 5 -> ok=0;
 * the reference gates at >= 10, matcher:330).
 * out_ok u8[P], out_rvec f64[P,3], out_tvec f64[P,3], out_n_inliers i32[P] (= len(inliers)),
 * out_mask u8[P,Nmax] inlier mask (NULL ok), out_mean_err f32[P] (NULL ok).
 * Debug outputs (NULL ok): out_sets i32[P,iters,5] minimal sets, out_models f64[P,iters,6]
 * (rvec,tvec) per hypothesis, out_counts i32[P,iters], out_best_iter i32[P], out_niters i32[P]. */
int nclt_pnp_ransac(nclt_ctx* ctx, const float* obj, const float* img, const int32_t* n, int P, int Nmax,
                    const nclt_pnp_params* prm, uint8_t* out_ok, double* out_rvec, double* out_tvec,
                    int32_t* out_n_inliers, uint8_t* out_mask, float* out_mean_err, int32_t* out_sets,
                    double* out_models, int32_t* out_counts, int32_t* out_best_iter, int32_t* out_niters);
/* device-pointer variant; debug outputs are not available here */
int nclt_pnp_ransac_dev(nclt_ctx* ctx, const float* obj, const float* img, const int32_t* n, int P, int Nmax,
                        const nclt_pnp_params* prm, uint8_t* out_ok, double* out_rvec, double* out_tvec,
                        int32_t* out_n_inliers, uint8_t* out_mask, float* out_mean_err);
/* K4 alone: inlier counts of caller-supplied hypotheses models f64[P,iters,6] (staged parity,
 * SURVEY.md section 7 (ii)); out_counts i32[P,iters]. Host pointers. */
int nclt_pnp_score(nclt_ctx* ctx, const float* obj, const float* img, const int32_t* n, int P, int Nmax,
                   const nclt_pnp_params* prm, const double* models, int32_t* out_counts);
/* replaces cv2.projectPoints(obj, rvec, tvec, K, DIST) (matcher:353): out f32[n,2]. Host pointers. */
int nclt_project_points(nclt_ctx* ctx, const float* obj, int n, const double* rvec, const double* tvec,
                        double fx, double fy, double cx, double cy, float* out);

/* ---- the whole repeat-time path for a batch of frames ------------------------------------ */
typedef struct nclt_localize_params {
    int mode;              /* 0: knnMatch(k=2)+Lowe ratio (checkpoint_a_selftest.py:68-71);
                              1: crossCheck match(desc_t, desc_curr) (visual_landmark_matcher.py:327) */
    int ratio_num, ratio_den; /* LOWE_RATIO = 0.80 -> 4, 5 (matcher:66) */
    int min_matches;       /* MIN_MATCHES = 10 (matcher:65; gate at matcher:330, selftest:72) */
    int min_inliers;       /* MIN_INLIERS = 10 (matcher:70; gate at matcher:349) */
    float reproj_max_px;   /* REPROJ_MAX_PX = 2.0 (matcher:67; gate at matcher:356) */
    nclt_pnp_params pnp;
} nclt_localize_params;

/* replaces the per-candidate loop body of visual_landmark_matcher.py:318-380 and
 * checkpoint_a_selftest.py:62-103 for B frames x C candidates: the `len(desc_t) < MIN_MATCHES -> continue` skip
 * (matcher:321-322, selftest:64-65: a candidate keyframe with fewer than min_matches rows is never matched and
 * reports 0 matches), match, MIN_MATCHES gate on the match count (matcher:330, selftest:72), gather
 * obj/img points, solvePnPRansac, MIN_INLIERS and mean-reprojection gates, and the "most inliers,
 * earliest candidate on ties" selection (matcher:379-380).
 * q u8[B,Nq,32], q_pts2d f32[B,Nq,2] keypoint pixel coordinates, q_n/cand as for nclt_match_*.
 * Per frame: out_best_cand i32[B] winning candidate SLOT (-1 = none accepted), out_n_inliers i32[B],
 * out_reproj f32[B], out_rvec/out_tvec f64[B,3] (teach camera in the current camera frame, as
 * solvePnPRansac returns it).  out_n_problems: HOST int, PnP problems solved; passing NULL selects the
 * fully asynchronous mode (no host synchronisation, see nclt_ctx_overflow): the _dev variant just enqueues
 * its kernels; the host-pointer variant enqueues input copies + kernels + result copies and returns -
 * its host buffers (page-locked for real overlap) must stay valid, and hold the results, only after
 * nclt_ctx_sync().  Two contexts used alternately overlap one batch's copies with the other's kernels.
 * Optional per (frame, candidate) outputs (NULL ok): out_item_nmatch i32[B,C] matches after the
 * ratio / crossCheck filter, out_item_ok u8[B,C], out_item_ninl i32[B,C], out_item_err f32[B,C],
 * out_item_rvec/out_item_tvec f64[B,C,3] (valid where nmatch >= min_matches).
 * With a candidate list (cand != NULL) and none of out_item_ok / _ninl / _err / _rvec / _tvec requested, candidates
 * that cannot become the frame's result are not solved: PnP runs first on the candidate with the most matches, then
 * only on those with more matches than its accepted inlier count (or as many, from an earlier slot).  The per-frame
 * outputs are those of the full loop; out_n_problems counts the problems actually solved. */
int nclt_localize_batch(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const float* q_pts2d,
                        const int32_t* q_n, int B, int Nq, const int32_t* cand, int C,
                        const nclt_localize_params* prm, int32_t* out_best_cand, int32_t* out_n_inliers,
                        float* out_reproj, double* out_rvec, double* out_tvec, int32_t* out_n_problems,
                        int32_t* out_item_nmatch, uint8_t* out_item_ok, int32_t* out_item_ninl,
                        float* out_item_err, double* out_item_rvec, double* out_item_tvec);
/* device pointers for every array argument; out_n_problems stays a HOST int (the problem count is
 * the one value read back inside the call). Returns with the remaining work enqueued. */
int nclt_localize_batch_dev(nclt_ctx* ctx, const nclt_lib* lib, const uint8_t* q, const float* q_pts2d,
                            const int32_t* q_n, int B, int Nq, const int32_t* cand, int C,
                            const nclt_localize_params* prm, int32_t* out_best_cand, int32_t* out_n_inliers,
                            float* out_reproj, double* out_rvec, double* out_tvec, int32_t* out_n_problems,
                            int32_t* out_item_nmatch, uint8_t* out_item_ok, int32_t* out_item_ninl,
                            float* out_item_err, double* out_item_rvec, double* out_item_tvec);

/* ---- teach-time map builder ------------------------------------------------------------ */
/* Log-odds occupancy grid of scripts/common/teach_run_depth_mapper.py (class TeachDepthMapper,
 * :83-100): W = int(width_m/res) columns, H = int(height_m/res) rows, origin (origin_x, origin_y).
 * The device grid holds exact integers in units of 0.2 (L_FREE -0.4 = -2, L_OCC +1.4 = +7,
 * L_MIN/L_MAX +-5 = +-25); updates are applied in the reference's order (frame, ray, cell). */
int nclt_occ_create(nclt_ctx* ctx, double origin_x, double origin_y, double res, int W, int H, nclt_occ** out);
int nclt_occ_destroy(nclt_ctx* ctx, nclt_occ* occ);
int nclt_occ_reset(nclt_ctx* ctx, nclt_occ* occ);

/* replaces TFRelay.depth_cb (tf_wall_clock_relay.py:868-903) followed by TeachDepthMapper.cb
 * (teach_run_depth_mapper.py:125-170) for F depth frames, processed in order.
 * depth f32[F,Hd,Wd] metres (32FC1) or, with is_u16 != 0, u16[F,Hd,Wd] millimetres (16UC1);
 * T f64[F,16] row-major map<-camera_link matrices (what _tf_to_matrix builds, mapper:64-80);
 * fx,fy,cx,cy the relay intrinsics (relay:80-81). */
int nclt_occ_integrate_depth(nclt_ctx* ctx, nclt_occ* occ, const void* depth, int is_u16, int F, int Hd, int Wd,
                             const double* T, double fx, double fy, double cx, double cy);
/* device pointers; optionally also returns the relay's point clouds: out_pts f32[F,pts_cap,3],
 * out_pts_n i32[F] (NULL ok) */
int nclt_occ_integrate_depth_dev(nclt_ctx* ctx, nclt_occ* occ, const void* depth, int is_u16, int F, int Hd,
                                 int Wd, const double* T, double fx, double fy, double cx, double cy,
                                 float* out_pts, int32_t* out_pts_n, int pts_cap);
/* replaces TeachDepthMapper.cb(PointCloud2) alone: pts f32[F,Nmax,3] camera_link points
 * (x,y,z FLOAT32 at offsets 0/4/8, point_step 12), n i32[F] points per cloud. */
int nclt_occ_integrate_points(nclt_ctx* ctx, nclt_occ* occ, const float* pts, const int32_t* n, int F, int Nmax,
                              const double* T);
int nclt_occ_integrate_points_dev(nclt_ctx* ctx, nclt_occ* occ, const float* pts, const int32_t* n, int F,
                                  int Nmax, const double* T);
/* replaces TFRelay.depth_cb alone (the PointCloud2 payload, also consumed by Nav2's obstacle
 * layer): out_pts f32[F,pts_cap,3] (z, -px, -py), out_n i32[F]. Host pointers. */
int nclt_depth_to_points(nclt_ctx* ctx, const void* depth, int is_u16, int F, int Hd, int Wd, double fx,
                         double fy, double cx, double cy, float* out_pts, int32_t* out_n, int pts_cap);
/* read-back (any pointer may be NULL), host pointers: out_logodds f32[H,W] = TeachDepthMapper.grid
 * (0.2 * units); out_pgm u8[H,W] = the P5 payload of save() (mapper:208-216: 205 unknown, 0
 * occupied, 254 free, flipud); out_units i32[H,W]; out_counters i64[3] = frames_integrated,
 * total_points_integrated, frames_skipped_empty (mapper:94-97). */
int nclt_occ_read(nclt_ctx* ctx, nclt_occ* occ, float* out_logodds, uint8_t* out_pgm, int32_t* out_units,
                  int64_t* out_counters);

/* ---- teach-time keypoint lifting --------------------------------------
 * Replaces the per-keypoint Python loop of scripts/common/visual_landmark_recorder.py:247-291:
 * ORB keypoints (kp.pt, float32) + the aligned depth image (u16 millimetres, recorder:36-50) ->
 * the keypoints the recorder keeps and their optical-frame 3-D points.  Gates, in the reference's
 * arithmetic: np.round to the pixel, 1 <= u < W-1, 1 <= v < H-1, v > ground_y (recorder:250-253);
 * depth_min < d < depth_max with d = float32(mm)/1000 (:260,:268); float32 std of the > 0.01 m
 * values of the 3x3 neighbourhood < depth_std_max, 999 when fewer than 3 (:261-267,:269);
 * X = (u - cx) * d / fx in float64, rounded to float32 (:283-286).  F frames per call.
 * kpts_xy f32[F,Nmax,2], n_kpts i32[F] -> out_keep i32[F,Nmax] (indices into the frame's
 * keypoints, ascending = ORB order), out_pts3d f32[F,Nmax,3], out_n i32[F].  The frame-level
 * gate "fewer than 30 points -> no landmark" (:271-279) is the caller's (recorder.py). */
typedef struct nclt_lift_params {
    double fx, fy, cx, cy;        /* recorder:53-54 */
    int32_t ground_y;             /* GROUND_Y_THRESHOLD = 180 (recorder:72) */
    float depth_min_m;            /* DEPTH_MIN_M = 0.5 */
    float depth_max_m;            /* DEPTH_MAX_M = 15.0 */
    float depth_std_max_m;        /* DEPTH_VAR_MAX_M = 0.30 */
} nclt_lift_params;
int nclt_lift_keypoints(nclt_ctx* ctx, const uint16_t* depth_mm, int F, int H, int W, const float* kpts_xy,
                        const int32_t* n_kpts, int Nmax, const nclt_lift_params* prm, int32_t* out_keep,
                        float* out_pts3d, int32_t* out_n);
int nclt_lift_keypoints_dev(nclt_ctx* ctx, const uint16_t* depth_mm, int F, int H, int W, const float* kpts_xy,
                            const int32_t* n_kpts, int Nmax, const nclt_lift_params* prm, int32_t* out_keep,
                            float* out_pts3d, int32_t* out_n);

/* ---- hit-count occupancy ----------------------------------------------
 * datasets/rover/scripts/occupancy_astar.py:142-187 `build_occupancy`: points f64[N,3] (world, HOST),
 * labels i8[N] (0 floor, 1 obstacle, -1 ignore) -> per X-Z cell floor / obstacle hit counts
 * (np.add.at) and occupancy i8 (-1 unknown, 0 free, 1 occupied; known = total >= min_total,
 * occupied = obstacle >= min_obstacle).  Grid: origin = min of the classified points - 0.5 m,
 * n = int((max + 0.5 - origin) / grid_res) + 1 (lines 155-161), cell = clip(int((p - origin) /
 * grid_res), 0, n - 1), row-major [nz][nx].  out_origin f64[2] = (x_min, z_min), out_dims i32[2] =
 * (nx, nz); out_occ i8[cell_cap], out_floor / out_obs i32[cell_cap] (NULL ok).  Fails with
 * NCLT_ERR_ARG (out_dims filled in) when nx * nz > cell_cap, or when no label is >= 0. */
int nclt_hitcount_occupancy(nclt_ctx* ctx, const double* points, const int8_t* labels, long long N, double grid_res,
                            int min_total, int min_obstacle, long long cell_cap, double* out_origin,
                            int32_t* out_dims, int8_t* out_occ, int32_t* out_floor, int32_t* out_obs);

/* ---- ORB feature extraction --------------------------------------------
 * Replaces `cv2.ORB_create(nfeatures=500)` + `cv2.cvtColor(BGR2GRAY)` + `orb.detectAndCompute(gray, None)`
 * at scripts/common/visual_landmark_matcher.py:207,305-306 and visual_landmark_recorder.py:159,240-241
 * (cv2 defaults: scaleFactor 1.2, 8 levels, edgeThreshold 31, WTA_K 2, HARRIS_SCORE, patch 31, FAST
 * threshold 20), for F frames of W x H per call.  Keypoints, their ORDER and descriptors are bit-identical
 * to cv2 4.13.0.  img: u8[F,H,W] (channels = 1) or u8[F,H,W,3] BGR (channels = 3).
 * out_kp f32[F,out_cap,6] = (pt.x, pt.y, size, angle, response, octave) per keypoint, out_desc
 * u8[F,out_cap,32], out_n i32[F].  out_cap >= 500: a level keeps every keypoint that ties with its last
 * retained Harris response, so a frame can exceed nfeatures; NCLT_ERR_STATE if it exceeds out_cap.
 * The handle owns its device memory: pyramid, score map, blurred pyramid, an input staging plane and
 * worst-case candidate lists, ~11 MB per frame slot at 640 x 480 (max_frames slots) - one handle, one
 * caller at a time, like nclt_ctx.  The call returns after its results have been read back (it
 * synchronises the context's stream once; a second, internal stream runs the resize chain and the blur
 * beside the other kernels). */
typedef struct nclt_orb nclt_orb;
int nclt_orb_create(nclt_ctx* ctx, int W, int H, int max_frames, int out_cap, nclt_orb** out);
int nclt_orb_destroy(nclt_ctx* ctx, nclt_orb* orb);
/* per-level geometry of the handle: width, height, features retained, scale (8 entries each; NULL ok) */
int nclt_orb_levels(const nclt_orb* orb, int32_t* out_w, int32_t* out_h, int32_t* out_n, float* out_scale);
int nclt_orb_detect_and_compute(nclt_ctx* ctx, nclt_orb* orb, const uint8_t* img, int channels, int F,
                                float* out_kp, uint8_t* out_desc, int32_t* out_n);
/* Where the per-level selection (OpenCV's KeyPointsFilter::retainBest, twice) runs: 0 = on the device (default; the
 * libstdc++ nth_element / partition steps restated for one GPU thread per level, csrc/orb_select.cuh - the whole call
 * is then one stream of kernels and one read-back), 1 = on the host with the std:: algorithms themselves (one extra
 * round trip).  Mode 0 falls back to mode 1 by itself in the one case it does not restate (introselect's heap-select
 * branch); nclt_orb_host_fallbacks counts those calls.  Results are identical in both modes.  Mode 2 is a
 * diagnostic: device selection followed by a forced hand-over (tests the fall-back path). */
int nclt_orb_set_select(nclt_ctx* ctx, nclt_orb* orb, int mode);
long long nclt_orb_host_fallbacks(const nclt_orb* orb);
/* The same call in two halves, so that one host thread can keep two handles busy (the PCIe copy of one batch under
 * the kernels of the other; one context + handle per pipeline): nclt_orb_submit enqueues the input copy, every
 * kernel and the result copies on the context's stream and returns; nclt_orb_wait synchronises, checks the
 * selection flags (falls back to the host selection if asked to) and fills out_n.  The host buffers (page-locked for
 * real overlap) must stay valid, and hold the results only after nclt_orb_wait returned 0.  A handle takes one
 * submitted call at a time (NCLT_ERR_STATE otherwise); with host selection (mode 1) submit is the whole call. */
int nclt_orb_submit(nclt_ctx* ctx, nclt_orb* orb, const uint8_t* img, int channels, int F, float* out_kp,
                    uint8_t* out_desc, int32_t* out_n);
int nclt_orb_wait(nclt_ctx* ctx, nclt_orb* orb);
/* nclt_orb_submit with DEVICE pointers (img, out_kp, out_desc, out_n): nothing but kernels and one device-to-device
 * copy are enqueued, the call returns at once and work queued behind it on the context's stream (nclt_localize_batch_dev
 * on the descriptors) may consume the outputs.  nclt_orb_wait later synchronises and checks the selection flags; when
 * it had to hand the selection over to the host (nclt_orb_host_fallbacks grew - never seen on camera images) it
 * rewrites the outputs, and whatever consumed them meanwhile has to be run again (pipeline.py does). */
int nclt_orb_submit_dev(nclt_ctx* ctx, nclt_orb* orb, const uint8_t* img, int channels, int F, float* out_kp,
                        uint8_t* out_desc, int32_t* out_n);
/* img, out_kp, out_desc, out_n are DEVICE pointers */
int nclt_orb_detect_and_compute_dev(nclt_ctx* ctx, nclt_orb* orb, const uint8_t* img, int channels, int F,
                                    float* out_kp, uint8_t* out_desc, int32_t* out_n);

#ifdef __cplusplus
}
#endif
#endif /* NCLT_B200_H */
