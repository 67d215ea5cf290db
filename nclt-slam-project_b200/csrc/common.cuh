// Shared declarations for the sm_100a kernels behind the C ABI (include/nclt_b200.h).
#pragma once
#include <cuda_runtime.h>
#include "../../include/nclt_b200.h"
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>


// top-2 key: dist (9 bits, 0..256) << 23 | row index (23 bits). min() on the key is the
// reference's tie rule (lowest train index wins on equal distance).
#define NCLT_KEY_SHIFT 23
#define NCLT_KEY_IDX_MASK 0x7FFFFFu
#define NCLT_KEY_INVALID 0xFFFFFFFFu
#define NCLT_MWC_N 8192

struct nclt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    std::string err;
    // grow-only device scratch
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t scratch_off = 0;      // stack pointer inside the current chunk
    std::vector<std::pair<void*, size_t>> scratch_chunks;   // `scratch` = chunks[scratch_chunk].first
    int scratch_chunk = -1;
    // pinned host staging for small result reads
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // counts of kernel launches issued through this context (bench.py gpu_launches)
    unsigned long long launches = 0;
    // optional timing of the dominant kernel (bench.py roofline): event pairs around every
    // Hamming top-2 launch, summed by nclt_ctx_profile_read
    int engine = 0;             // 0 = integer pipe (LOP3+POPC), 1 / 2 = tensor cores (fp8 / block-scaled fp4) for all-keyframe ratio matching
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_tag;  // tag of the pair that starts at event 2i (nclt_prof_mark_tag)
    size_t prof_used = 0;
    // async pipeline: PnP problems dropped because a batch produced more than its problem capacity
    int* d_overflow = nullptr;
    uint32_t* d_mwc = nullptr;      // first NCLT_MWC_N outputs of cv::RNG((uint64)-1): the same stream for every solvePnPRansac call

    unsigned long long* d_tc_clk = nullptr;   // tensor-kernel clock diagnostics (profile mode), 64 x u64
    int tc_clk_launch = 0;
    // bumped whenever device memory a captured CUDA graph may point into is freed or moved: scratch chunks,
    // tensor-engine library images / split tables, library growth (nclt_ctx_alloc_generation)
    unsigned long long alloc_gen = 0;
    int tail_sms = 0;           // SMs the persistent matching kernel leaves free for the tail kernels of another context
};

struct nclt_lib {
    int device = 0;
    int n_kf = 0;
    int n_desc = 0;          // total rows in use
    int cap_desc = 0;        // allocated rows
    int cap_kf = 0;
    uint4* d_desc = nullptr;     // [cap_desc][2] uint4 = 32 B per descriptor
    float* d_pts3d = nullptr;    // [cap_desc][3]
    int* d_start = nullptr;      // [cap_kf] first row of keyframe
    int* d_count = nullptr;      // [cap_kf] rows in keyframe
    std::vector<int> h_start, h_count;
    int max_count = 0;
    void* tc_cache = nullptr;   // tensor-core operand images + tile table (tc_hamming.cu), built lazily
    void* tc4_cache = nullptr;  // same for the block-scaled fp4 flavour
    void* tc4x_cache = nullptr; // same for the crossCheck flavour (index-carrying cells, 192-byte rows)
};

// A ragged set of 32-byte descriptors on the device.
struct SegView {
    const uint4* base;   // 2 uint4 per row
    const int* start;    // per-segment first row, or nullptr -> seg * stride
    const int* count;    // per-segment rows, or nullptr -> stride
    int stride;
};

static inline int nclt_fail(nclt_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess) {
    if (c) {
        c->err = what;
        if (e != cudaSuccess) {
            c->err += ": ";
            c->err += cudaGetErrorString(e);
        }
    }
    return code;
}

#define CU_TRY(ctx, call)                                                         \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return nclt_fail((ctx), NCLT_ERR_CUDA, #call, _e); \
    } while (0)

int nclt_scratch_reserve(nclt_ctx* c, size_t bytes);
// Profile mode: event pairs around selected launches, summed per tag by nclt_ctx_profile_read[_tags].
// Tags: 0 Hamming top-2 (the dominant matching kernel), 1 k_occ_frame, 2 k_occ_apply, 3 k_pnp_hypo, 4 k_pnp_score,
// 5 k_pnp_finish, 6 candidate verification, 7 free.
#define NCLT_PROF_TAGS 8
static inline void nclt_prof_mark_tag(nclt_ctx* c, int tag) {
    if (!c->prof) return;
    if (c->prof_used == c->prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->prof_ev.push_back(e);
    }
    if ((c->prof_used & 1) == 0) {
        if (c->prof_tag.size() <= c->prof_used / 2) c->prof_tag.resize(c->prof_used / 2 + 1);
        c->prof_tag[c->prof_used / 2] = tag;
    }
    cudaEventRecord(c->prof_ev[c->prof_used++], c->stream);
}
static inline void nclt_prof_mark(nclt_ctx* c) { nclt_prof_mark_tag(c, 0); }
int nclt_pinned_reserve(nclt_ctx* c, size_t bytes);
int nclt_overflow_take(nclt_ctx* c);
int nclt_check_host_lists(nclt_ctx* c, const nclt_lib* L, const int32_t* q_n, int B, int Nq, const int32_t* cand, int C);

// ---- hamming.cu ----
struct MatchLaunch {
    SegView A, B;            // rows of A are matched against rows of B
    const int* cand;         // [n_outer, C] B-segment (or A-segment when swap) ids, -1 = skip; nullptr -> c
    int n_outer;             // frames
    int C;                   // candidates per frame
    int swap;                // 0: A-seg = frame, B-seg = cand ; 1: A-seg = cand, B-seg = frame
    int a_rows_max;          // output row stride per item
    int nsplit;              // split B rows over this many CTAs (flat mode); 1 otherwise
    int b_seg_fixed;         // >=0: every item uses this B segment (flat mode)
    uint2* out_keys;         // [n_items*nsplit, a_rows_max] (best, second) keys, or nullptr
    int2* out_idx;           // [n_items, a_rows_max] or nullptr   (only when nsplit==1)
    ushort2* out_dist;       // [n_items, a_rows_max] or nullptr
};
int launch_hamming_top2(nclt_ctx* c, const MatchLaunch& m, uint32_t idx_offset);
int launch_hamming_cross(nclt_ctx* c, const SegView& frames, const SegView& lib, const int* cand, int n_outer, int C, int Nq,
                         int fwd_rows_max, uint2* fwd, uint2* bwd);
int launch_merge_top2(nclt_ctx* c, const uint2* parts, int n_items, int nparts, int rows, long long part_stride,
                      long long item_stride, uint2* out_keys, int2* out_idx, ushort2* out_dist);
int launch_cross_combine(nclt_ctx* c, const uint2* fwd_keys, const uint2* bwd_keys, const SegView& Lib,
                         const int* cand, int n_outer, int C, int fwd_rows_max, int bwd_rows_max, int2* out_pairs,
                         unsigned short* out_dist, int* out_n, int out_stride);
int launch_ratio_compact(nclt_ctx* c, const uint2* keys, const int* a_count, int a_stride_cnt, const int* cand,
                         int n_outer, int C, int a_rows_max, int num, int den, int2* out_pairs, int* out_n);
double run_popc_peak(nclt_ctx* c, int iters, float* ms_out);

// ---- tc_hamming.cu ----
void nclt_tc_release(nclt_lib* L);
int tc_match_ratio_all(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, int num, int den,
                       int32_t* out_pairs, int32_t* out_n, bool fp4);
int tc4_match_cross_all(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, int Nmax,
                        int32_t* out_pairs, uint16_t* out_dist, int32_t* out_n);
int tc_match_flat2(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, uint32_t idx_offset,
                   uint32_t* out_keys, bool fp4);

// ---- pnp.cu ----
struct PnpBuffers {
    int* sets;         // [P][iters][5]
    double* models;    // [P][iters][6]
    int* counts;       // [P][iters]
    int* state;        // [P][4] replay state: cursor, best, max_good, niters
};
int launch_pnp(nclt_ctx* c, const float* obj, const float* img, const int* n, int P, const int* P_dev, int Nmax,
               const nclt_pnp_params* prm, const PnpBuffers& buf, const double* models_override,
               unsigned char* ok, double* rvec, double* tvec, int* n_inl, unsigned char* mask, float* mean_err,
               int* best_iter, int* niters, bool score_only);
int launch_project_points(nclt_ctx* c, const float* obj, int n, const double* rvec, const double* tvec, double fx,
                          double fy, double cx, double cy, float* out);
