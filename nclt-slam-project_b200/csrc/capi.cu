// extern "C" entry points declared in include/nclt_b200.h: context, teach library, matching.
#include "../../include/nclt_b200.h"
#include "common.cuh"
#include "scratch.cuh"

#include <algorithm>
#include <cstring>

// Guarantee `bytes` of carve room after the current stack position. Nothing already carved moves:
// if the current chunk is too small the stack continues in another (new or idle) chunk.
int nclt_scratch_reserve(nclt_ctx* c, size_t bytes) {
    bytes += 4096;
    if (c->scratch_chunk >= 0 && c->scratch_off + bytes <= c->scratch_bytes) return NCLT_OK;
    const bool idle = c->scratch_chunk < 0 || (c->scratch_chunk == 0 && c->scratch_off == 0);
    if (idle) {
        c->alloc_gen++;  // frees every chunk: captured graphs are stale from here on
        // top-level call and nothing in use: consolidate into one chunk that fits
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        size_t total = 0;      // the sum: what used to need several chunks then fits one
        for (auto& ch : c->scratch_chunks) {
            total += ch.second;
            cudaFree(ch.first);
        }
        c->scratch_chunks.clear();
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        c->scratch_chunk = -1;
        c->scratch_off = 0;
        size_t want = std::max(total, bytes + bytes / 4);
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) return nclt_fail(c, NCLT_ERR_NOMEM, "cudaMalloc scratch", e);
        }
        c->scratch_chunks.emplace_back(p, want);
        c->scratch_chunk = 0;
        c->scratch = p;
        c->scratch_bytes = want;
        return NCLT_OK;
    }
    // in use: continue in a later chunk (reuse an idle one if it is large enough)
    for (int i = c->scratch_chunk + 1; i < (int)c->scratch_chunks.size(); ++i) {
        if (c->scratch_chunks[i].second >= bytes) {
            if (i != c->scratch_chunk + 1) std::swap(c->scratch_chunks[i], c->scratch_chunks[c->scratch_chunk + 1]);
            c->scratch_chunk += 1;
            c->scratch = c->scratch_chunks[c->scratch_chunk].first;
            c->scratch_bytes = c->scratch_chunks[c->scratch_chunk].second;
            c->scratch_off = 0;
            return NCLT_OK;
        }
    }
    // (re-using an idle chunk above moves nothing; a NEW chunk does not invalidate earlier pointers either, but a
    // graph captured before it existed may be replayed next to calls that now carve from it - stay conservative)
    c->alloc_gen++;
    void* p = nullptr;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return nclt_fail(c, NCLT_ERR_NOMEM, "cudaMalloc scratch chunk", e);
    c->scratch_chunks.insert(c->scratch_chunks.begin() + c->scratch_chunk + 1, std::make_pair(p, want));
    c->scratch_chunk += 1;
    c->scratch = p;
    c->scratch_bytes = want;
    c->scratch_off = 0;
    return NCLT_OK;
}

int nclt_pinned_reserve(nclt_ctx* c, size_t bytes) {
    if (bytes <= c->pinned_bytes) return NCLT_OK;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->pinned) cudaFreeHost(c->pinned);
    c->pinned = nullptr;
    c->pinned_bytes = 0;
    CU_TRY(c, cudaMallocHost(&c->pinned, bytes));
    c->pinned_bytes = bytes;
    return NCLT_OK;
}

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int nclt_abi_version(void) { return 1; }

extern "C" int nclt_ctx_create(int device, void* stream, nclt_ctx** out) {
    if (!out) return NCLT_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return NCLT_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return NCLT_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NCLT_ERR_CUDA;
    if (prop.major < 10) return NCLT_ERR_CUDA;   // sm_100a-only binary
    nclt_ctx* c = new nclt_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (stream) {
        c->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            return NCLT_ERR_CUDA;
        }
        c->own_stream = true;
    }
    // the capacity-overflow counter of the asynchronous paths (PnP problems, tensor-engine candidates) exists from the
    // start: allocating it later would not be capturable into a CUDA graph
    // [0] problems dropped by asynchronous calls, [1] item counter of the persistent fp4 matching kernel
    if (cudaMalloc(&c->d_overflow, 8) != cudaSuccess || cudaMemset(c->d_overflow, 0, 8) != cudaSuccess) {
        if (c->own_stream) cudaStreamDestroy(c->stream);
        delete c;
        return NCLT_ERR_NOMEM;
    }
    {   // raw outputs of OpenCV's multiply-with-carry generator from the seed every solvePnPRansac call starts with
        // (k_pnp_sets walks this table instead of the sequential 64-bit recurrence)
        std::vector<uint32_t> raw(NCLT_MWC_N);
        uint64_t st = ~0ull;
        for (int i = 0; i < NCLT_MWC_N; ++i) {
            st = (uint64_t)(uint32_t)st * 4164903690ull + (uint32_t)(st >> 32);
            raw[i] = (uint32_t)st;
        }
        if (cudaMalloc(&c->d_mwc, NCLT_MWC_N * 4) != cudaSuccess ||
            cudaMemcpy(c->d_mwc, raw.data(), NCLT_MWC_N * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(c->d_overflow);
            if (c->own_stream) cudaStreamDestroy(c->stream);
            delete c;
            return NCLT_ERR_NOMEM;
        }
    }
    *out = c;
    return NCLT_OK;
}

extern "C" int nclt_ctx_destroy(nclt_ctx* c) {
    if (!c) return NCLT_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& ch : c->scratch_chunks) cudaFree(ch.first);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    if (c->d_overflow) cudaFree(c->d_overflow);
    if (c->d_mwc) cudaFree(c->d_mwc);
    if (c->d_tc_clk) cudaFree(c->d_tc_clk);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return NCLT_OK;
}

extern "C" int nclt_ctx_sync(nclt_ctx* c) {
    if (!c) return NCLT_ERR_ARG;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}

extern "C" const char* nclt_last_error(nclt_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" unsigned long long nclt_ctx_launches(nclt_ctx* c) { return c ? c->launches : 0ull; }
extern "C" unsigned long long nclt_ctx_alloc_generation(nclt_ctx* c) { return c ? c->alloc_gen : 0ull; }

extern "C" int nclt_ctx_overflow(nclt_ctx* c, int reset) {
    if (!c) return NCLT_ERR_ARG;
    if (!c->d_overflow) return 0;
    cudaSetDevice(c->device);
    int v = 0;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return NCLT_ERR_CUDA;
    if (cudaMemcpy(&v, c->d_overflow, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return NCLT_ERR_CUDA;
    if (reset) cudaMemset(c->d_overflow, 0, 4);
    return v;
}

// synchronises, returns the overflow counter and clears it (synchronous entry points that can repair an overflow)
int nclt_overflow_take(nclt_ctx* c) {
    if (!c->d_overflow) return 0;
    int v = 0;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return 0;
    if (cudaMemcpy(&v, c->d_overflow, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    if (v) cudaMemset(c->d_overflow, 0, 4);
    return v;
}

extern "C" int nclt_ctx_set_engine(nclt_ctx* c, int engine) {
    if (!c || engine < 0 || engine > 2) return nclt_fail(c, NCLT_ERR_ARG, "engine must be 0 (integer pipe), 1 (tensor cores, fp8) or 2 (tensor cores, block-scaled fp4)");
    c->engine = engine;
    return NCLT_OK;
}

extern "C" int nclt_ctx_set_tail_sms(nclt_ctx* c, int n) {
    if (!c || n < 0 || n >= c->sm_count) return nclt_fail(c, NCLT_ERR_ARG, "tail_sms must be in [0, SM count)");
    c->tail_sms = n;
    return NCLT_OK;
}

extern "C" int nclt_ctx_profile(nclt_ctx* c, int enable) {
    if (!c) return NCLT_ERR_ARG;
    c->prof = enable != 0;
    c->prof_used = 0;
    return NCLT_OK;
}

extern "C" int nclt_ctx_profile_read(nclt_ctx* c, double* ms_total, int* n_launches) {
    if (!c) return NCLT_ERR_ARG;
    cudaSetDevice(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    double ms = 0;
    int n = 0;
    for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
        float t = 0;
        if (c->prof_tag.size() > i / 2 && c->prof_tag[i / 2] != 0) continue;
        if (cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]) == cudaSuccess) { ms += t; n++; }
    }
    if (ms_total) *ms_total = ms;
    if (n_launches) *n_launches = n;
    c->prof_used = 0;
    return NCLT_OK;
}

extern "C" int nclt_ctx_profile_read_tags(nclt_ctx* c, double* ms_by_tag, int* n_by_tag) {
    if (!c || !ms_by_tag || !n_by_tag) return NCLT_ERR_ARG;
    cudaSetDevice(c->device);
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (int t = 0; t < NCLT_PROF_TAGS; ++t) { ms_by_tag[t] = 0; n_by_tag[t] = 0; }
    for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
        float t = 0;
        int tag = c->prof_tag.size() > i / 2 ? c->prof_tag[i / 2] : 0;
        if (tag < 0 || tag >= NCLT_PROF_TAGS) continue;
        if (cudaEventElapsedTime(&t, c->prof_ev[i], c->prof_ev[i + 1]) == cudaSuccess) { ms_by_tag[tag] += t; n_by_tag[tag]++; }
    }
    c->prof_used = 0;
    return NCLT_OK;
}

extern "C" double nclt_popc_peak(nclt_ctx* c, int iters, float* ms_out) {
    if (!c) return -1.0;
    cudaSetDevice(c->device);
    return run_popc_peak(c, iters, ms_out);
}

// ---------------------------------------------------------------------------------------
// teach library
// ---------------------------------------------------------------------------------------
static int lib_grow(nclt_ctx* c, nclt_lib* L, int need_desc, int need_kf) {
    if (need_desc > L->cap_desc) {
        int cap = std::max(need_desc, L->cap_desc + L->cap_desc / 2);
        uint4* nd = nullptr;
        float* np = nullptr;
        cudaError_t e = cudaMalloc(&nd, (size_t)cap * 32);
        if (e == cudaSuccess) e = cudaMalloc(&np, (size_t)cap * 12);
        if (e == cudaSuccess && L->n_desc) {
            e = cudaMemcpyAsync(nd, L->d_desc, (size_t)L->n_desc * 32, cudaMemcpyDeviceToDevice, c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(np, L->d_pts3d, (size_t)L->n_desc * 12, cudaMemcpyDeviceToDevice, c->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {      // nothing of the library has been touched yet: drop the new buffers
            if (nd) cudaFree(nd);
            if (np) cudaFree(np);
            return nclt_fail(c, e == cudaErrorMemoryAllocation ? NCLT_ERR_NOMEM : NCLT_ERR_CUDA, "lib_grow (descriptors)", e);
        }
        c->alloc_gen++;
        if (L->d_desc) cudaFree(L->d_desc);
        if (L->d_pts3d) cudaFree(L->d_pts3d);
        L->d_desc = nd;
        L->d_pts3d = np;
        L->cap_desc = cap;
    }
    if (need_kf > L->cap_kf) {
        int cap = std::max(need_kf, L->cap_kf + L->cap_kf / 2 + 16);
        int *ns = nullptr, *nc = nullptr;
        cudaError_t e = cudaMalloc(&ns, (size_t)cap * 4);
        if (e == cudaSuccess) e = cudaMalloc(&nc, (size_t)cap * 4);
        if (e == cudaSuccess && L->n_kf) {
            e = cudaMemcpyAsync(ns, L->d_start, (size_t)L->n_kf * 4, cudaMemcpyDeviceToDevice, c->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(nc, L->d_count, (size_t)L->n_kf * 4, cudaMemcpyDeviceToDevice, c->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            if (ns) cudaFree(ns);
            if (nc) cudaFree(nc);
            return nclt_fail(c, e == cudaErrorMemoryAllocation ? NCLT_ERR_NOMEM : NCLT_ERR_CUDA, "lib_grow (keyframe table)", e);
        }
        c->alloc_gen++;
        if (L->d_start) cudaFree(L->d_start);
        if (L->d_count) cudaFree(L->d_count);
        L->d_start = ns;
        L->d_count = nc;
        L->cap_kf = cap;
    }
    return NCLT_OK;
}

extern "C" int nclt_lib_create(nclt_ctx* c, int n_kf, const int32_t* kf_offsets, const uint8_t* desc,
                               const float* pts3d, nclt_lib** out) {
    if (!c || !out || n_kf < 0 || (n_kf > 0 && !kf_offsets)) return nclt_fail(c, NCLT_ERR_ARG, "lib_create args");
    *out = nullptr;
    cudaSetDevice(c->device);
    const int N = n_kf ? kf_offsets[n_kf] : 0;
    if (N < 0 || N > (int)NCLT_KEY_IDX_MASK) return nclt_fail(c, NCLT_ERR_ARG, "library too large for 23-bit index");
    if (N > 0 && !desc) return nclt_fail(c, NCLT_ERR_ARG, "lib_create: desc is null");
    nclt_lib* L = new nclt_lib();
    L->device = c->device;
    int rc = lib_grow(c, L, std::max(N, 1), std::max(n_kf, 1));
    if (rc) { delete L; return rc; }
    L->h_start.resize(n_kf);
    L->h_count.resize(n_kf);
    for (int k = 0; k < n_kf; ++k) {
        int cnt = kf_offsets[k + 1] - kf_offsets[k];
        if (cnt < 0) { nclt_lib_destroy(c, L); return nclt_fail(c, NCLT_ERR_ARG, "kf_offsets not monotone"); }
        L->h_start[k] = kf_offsets[k];
        L->h_count[k] = cnt;
        L->max_count = std::max(L->max_count, cnt);
    }
    cudaError_t e = cudaSuccess;
    if (N) {
        e = cudaMemcpyAsync(L->d_desc, desc, (size_t)N * 32, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess)
            e = pts3d ? cudaMemcpyAsync(L->d_pts3d, pts3d, (size_t)N * 12, cudaMemcpyHostToDevice, c->stream)
                      : cudaMemsetAsync(L->d_pts3d, 0, (size_t)N * 12, c->stream);
    }
    if (e == cudaSuccess && n_kf) {
        e = cudaMemcpyAsync(L->d_start, L->h_start.data(), (size_t)n_kf * 4, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(L->d_count, L->h_count.data(), (size_t)n_kf * 4, cudaMemcpyHostToDevice, c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        cudaStreamSynchronize(c->stream);
        nclt_lib_destroy(c, L);
        return nclt_fail(c, NCLT_ERR_CUDA, "lib_create: upload", e);
    }
    L->n_kf = n_kf;
    L->n_desc = N;
    *out = L;
    return NCLT_OK;
}

extern "C" int nclt_lib_append(nclt_ctx* c, nclt_lib* L, const uint8_t* desc, const float* pts3d, int n) {
    if (!c || !L || n < 0 || (n > 0 && !desc)) return nclt_fail(c, NCLT_ERR_ARG, "lib_append args");
    cudaSetDevice(c->device);
    if ((long long)L->n_desc + n > (long long)NCLT_KEY_IDX_MASK) return nclt_fail(c, NCLT_ERR_ARG, "library too large");
    int rc = lib_grow(c, L, L->n_desc + n, L->n_kf + 1);
    if (rc) return rc;
    int start = L->n_desc;
    if (n) {
        CU_TRY(c, cudaMemcpyAsync(L->d_desc + (size_t)start * 2, desc, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
        if (pts3d)
            CU_TRY(c, cudaMemcpyAsync(L->d_pts3d + (size_t)start * 3, pts3d, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        else
            CU_TRY(c, cudaMemsetAsync(L->d_pts3d + (size_t)start * 3, 0, (size_t)n * 12, c->stream));
    }
    L->h_start.push_back(start);
    L->h_count.push_back(n);
    CU_TRY(c, cudaMemcpyAsync(L->d_start + L->n_kf, &L->h_start[L->n_kf], 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(L->d_count + L->n_kf, &L->h_count[L->n_kf], 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    L->n_kf += 1;
    L->n_desc += n;
    L->max_count = std::max(L->max_count, n);
    return NCLT_OK;
}

extern "C" int nclt_lib_destroy(nclt_ctx* c, nclt_lib* L) {
    if (!L) return NCLT_OK;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    nclt_tc_release(L);
    if (L->d_desc) cudaFree(L->d_desc);
    if (L->d_pts3d) cudaFree(L->d_pts3d);
    if (L->d_start) cudaFree(L->d_start);
    if (L->d_count) cudaFree(L->d_count);
    delete L;
    return NCLT_OK;
}

extern "C" int nclt_lib_size(const nclt_lib* L, int* n_kf, int* n_desc, int* max_kf_rows) {
    if (!L) return NCLT_ERR_ARG;
    if (n_kf) *n_kf = L->n_kf;
    if (n_desc) *n_desc = L->n_desc;
    if (max_kf_rows) *max_kf_rows = L->max_count;
    return NCLT_OK;
}

// ---------------------------------------------------------------------------------------
// matching
// ---------------------------------------------------------------------------------------
static SegView lib_view(const nclt_lib* L) { return SegView{L->d_desc, L->d_start, L->d_count, 0}; }
static SegView frame_view(const uint8_t* q, const int32_t* q_n, int Nq) {
    return SegView{reinterpret_cast<const uint4*>(q), nullptr, q_n, Nq};
}

static int check_match_args(nclt_ctx* c, const nclt_lib* L, const void* q, int B, int Nq, int C) {
    if (!c || !L) return nclt_fail(c, NCLT_ERR_ARG, "null handle");
    if (L->device != c->device) return nclt_fail(c, NCLT_ERR_ARG, "library lives on another device");
    if (B < 0 || Nq <= 0 || C <= 0) return nclt_fail(c, NCLT_ERR_ARG, "bad B/Nq/C");
    if (B > 0 && !q) return nclt_fail(c, NCLT_ERR_ARG, "q is null");
    if ((reinterpret_cast<uintptr_t>(q) & 15) != 0) return nclt_fail(c, NCLT_ERR_ARG, "q must be 16-byte aligned");
    if (Nq > (int)NCLT_KEY_IDX_MASK) return nclt_fail(c, NCLT_ERR_ARG, "Nq too large");
    cudaSetDevice(c->device);
    return NCLT_OK;
}
// when cand == NULL the candidates are keyframes 0..C-1: they must exist
static int check_cand_null(nclt_ctx* c, const nclt_lib* L, const int32_t* cand, int C) {
    if (!cand && C > L->n_kf) return nclt_fail(c, NCLT_ERR_ARG, "C exceeds keyframe count with cand == NULL");
    return NCLT_OK;
}

extern "C" int nclt_match_knn2_dev(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B,
                                   int Nq, const int32_t* cand, int C, int32_t* out_idx, uint16_t* out_dist) {
    int rc = check_match_args(c, L, q, B, Nq, C);
    if (rc) return rc;
    if ((rc = check_cand_null(c, L, cand, C))) return rc;
    if (B == 0) return NCLT_OK;
    MatchLaunch m{};
    m.A = frame_view(q, q_n, Nq);
    m.B = lib_view(L);
    m.cand = cand; m.n_outer = B; m.C = C; m.swap = 0; m.a_rows_max = Nq; m.nsplit = 1; m.b_seg_fixed = -1;
    m.out_idx = reinterpret_cast<int2*>(out_idx);
    m.out_dist = reinterpret_cast<ushort2*>(out_dist);
    return launch_hamming_top2(c, m, 0u);
}

extern "C" int nclt_match_ratio_dev(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B,
                                    int Nq, const int32_t* cand, int C, int num, int den, int32_t* out_pairs,
                                    int32_t* out_n) {
    int rc = check_match_args(c, L, q, B, Nq, C);
    if (rc) return rc;
    if ((rc = check_cand_null(c, L, cand, C))) return rc;
    if (num <= 0 || den <= 0 || !out_pairs || !out_n) return nclt_fail(c, NCLT_ERR_ARG, "ratio args");
    if (B == 0) return NCLT_OK;
    // every frame against every keyframe: the tensor-core path (tc_hamming.cu), same outputs
    // (the tensor path's verification packs (distance << 16 | row): keyframes of 65 536 rows or more stay on the integer pipe)
    if (c->engine >= 1 && !cand && C == L->n_kf && (long long)B * Nq < (1LL << 30) && L->max_count < 65536)
        return tc_match_ratio_all(c, const_cast<nclt_lib*>(L), q, q_n, B, Nq, num, den, out_pairs, out_n, c->engine == 2);
    ScratchScope scope(c);
    size_t items = (size_t)B * C;
    if ((rc = nclt_scratch_reserve(c, pad256(items * Nq * sizeof(uint2))))) return rc;
    Carver cv(c);
    uint2* keys = cv.take<uint2>(items * Nq);
    MatchLaunch m{};
    m.A = frame_view(q, q_n, Nq);
    m.B = lib_view(L);
    m.cand = cand; m.n_outer = B; m.C = C; m.swap = 0; m.a_rows_max = Nq; m.nsplit = 1; m.b_seg_fixed = -1;
    m.out_keys = keys;
    if ((rc = launch_hamming_top2(c, m, 0u))) return rc;
    return launch_ratio_compact(c, keys, q_n, Nq, cand, B, C, Nq, num, den, reinterpret_cast<int2*>(out_pairs),
                                out_n);
}

extern "C" int nclt_match_cross_dev(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B,
                                    int Nq, const int32_t* cand, int C, int Nmax, int32_t* out_pairs,
                                    uint16_t* out_dist, int32_t* out_n) {
    int rc = check_match_args(c, L, q, B, Nq, C);
    if (rc) return rc;
    if ((rc = check_cand_null(c, L, cand, C))) return rc;
    if (!out_pairs || !out_n) return nclt_fail(c, NCLT_ERR_ARG, "cross outputs null");
    if (Nmax < L->max_count) return nclt_fail(c, NCLT_ERR_ARG, "Nmax smaller than the largest keyframe");
    if (B == 0) return NCLT_OK;
    // every frame against every keyframe (exp 63's whole-library ranking): both directions on the tensor cores
    // (tc_hamming.cu, index-carrying cells), same outputs
    if (c->engine == 2 && !cand && C == L->n_kf && (long long)B * Nq < (1LL << 30) && L->n_desc < (1 << 23) && Nq < (1 << 23))
        return tc4_match_cross_all(c, const_cast<nclt_lib*>(L), q, q_n, B, Nq, Nmax, out_pairs, out_dist, out_n);
    ScratchScope scope(c);
    size_t items = (size_t)B * C;
    if ((rc = nclt_scratch_reserve(c, pad256(items * Nmax * sizeof(uint2)) + pad256(items * Nq * sizeof(uint2)))))
        return rc;
    Carver cv(c);
    uint2* fwd = cv.take<uint2>(items * Nmax);
    uint2* bwd = cv.take<uint2>(items * Nq);
    // both directions from one pass over each item's distance matrix (hamming.cu, k_hamming_cross): fwd = teach rows
    // (the query side of match(desc_t, desc_curr)) -> nearest frame row, bwd = frame rows -> nearest teach row
    if ((rc = launch_hamming_cross(c, frame_view(q, q_n, Nq), lib_view(L), cand, B, C, Nq, Nmax, fwd, bwd))) return rc;
    return launch_cross_combine(c, fwd, bwd, lib_view(L), cand, B, C, Nmax, Nq,
                                reinterpret_cast<int2*>(out_pairs), out_dist, out_n, Nmax);
}

extern "C" int nclt_match_flat2_dev(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B,
                                    int Nq, uint32_t idx_offset, uint32_t* out_keys) {
    int rc = check_match_args(c, L, q, B, Nq, 1);
    if (rc) return rc;
    if (!out_keys) return nclt_fail(c, NCLT_ERR_ARG, "out_keys null");
    if ((unsigned long long)idx_offset + (unsigned)L->n_desc > NCLT_KEY_IDX_MASK)
        return nclt_fail(c, NCLT_ERR_ARG, "global index exceeds 23 bits");
    if (B == 0) return NCLT_OK;
    // tensor engines: per-keyframe (d1, d2 bound) on tcgen05, then an exact re-scan of the two keyframes that can
    // hold the global top-2 (tc_hamming.cu); identical keys
    if (c->engine >= 1 && L->n_kf > 0 && L->n_kf < 65536 && (long long)B * Nq < (1LL << 30) && L->max_count < 65536)
        return tc_match_flat2(c, const_cast<nclt_lib*>(L), q, q_n, B, Nq, idx_offset, out_keys, c->engine == 2);
    // the whole library is one segment; split its rows so that the grid fills the GPU
    int want = (c->sm_count * 8 + B - 1) / B;
    int max_split = std::max(1, (L->n_desc + 2047) / 2048);
    int nsplit = std::max(1, std::min(want, max_split));
    ScratchScope scope(c);
    size_t part = (size_t)B * Nq;
    if ((rc = nclt_scratch_reserve(c, pad256(part * nsplit * sizeof(uint2)) + 512))) return rc;
    Carver cv(c);
    uint2* parts = cv.take<uint2>(part * nsplit);
    int* seg = cv.take<int>(2);
    int h[2] = {0, L->n_desc};
    CU_TRY(c, cudaMemcpyAsync(seg, h, 8, cudaMemcpyHostToDevice, c->stream));
    MatchLaunch m{};
    m.A = frame_view(q, q_n, Nq);
    m.B = SegView{L->d_desc, seg, seg + 1, 0};
    m.cand = nullptr; m.n_outer = B; m.C = 1; m.swap = 0; m.a_rows_max = Nq; m.nsplit = nsplit; m.b_seg_fixed = 0;
    m.out_keys = parts;
    if ((rc = launch_hamming_top2(c, m, idx_offset))) return rc;
    // parts layout: [(item*nsplit + split), row]
    return launch_merge_top2(c, parts, B, nsplit, Nq, (long long)Nq, (long long)nsplit * Nq,
                             reinterpret_cast<uint2*>(out_keys), nullptr, nullptr);
}

extern "C" int nclt_merge_top2_dev(nclt_ctx* c, const uint32_t* parts, int nparts, int rows, uint32_t* out_keys,
                                   int32_t* out_idx, uint16_t* out_dist) {
    if (!c || !parts || nparts <= 0 || rows < 0) return nclt_fail(c, NCLT_ERR_ARG, "merge args");
    cudaSetDevice(c->device);
    return launch_merge_top2(c, reinterpret_cast<const uint2*>(parts), 1, nparts, rows, (long long)rows, 0LL,
                             reinterpret_cast<uint2*>(out_keys), reinterpret_cast<int2*>(out_idx),
                             reinterpret_cast<ushort2*>(out_dist));
}

// ---- host-pointer variants: stage through device scratch -------------------------------
struct HostStage {
    const uint8_t* d_q = nullptr;
    const int32_t* d_qn = nullptr;
    const int32_t* d_cand = nullptr;
};
static int stage_inputs(nclt_ctx* c, Carver& cv, const uint8_t* q, const int32_t* q_n, int B, int Nq,
                        const int32_t* cand, int C, HostStage& s) {
    uint8_t* dq = cv.take<uint8_t>((size_t)B * Nq * 32);
    CU_TRY(c, cudaMemcpyAsync(dq, q, (size_t)B * Nq * 32, cudaMemcpyHostToDevice, c->stream));
    s.d_q = dq;
    if (q_n) {
        int32_t* d = cv.take<int32_t>(B);
        CU_TRY(c, cudaMemcpyAsync(d, q_n, (size_t)B * 4, cudaMemcpyHostToDevice, c->stream));
        s.d_qn = d;
    }
    if (cand) {
        int32_t* d = cv.take<int32_t>((size_t)B * C);
        CU_TRY(c, cudaMemcpyAsync(d, cand, (size_t)B * C * 4, cudaMemcpyHostToDevice, c->stream));
        s.d_cand = d;
    }
    return NCLT_OK;
}
static size_t stage_bytes(int B, int Nq, int C) {
    return pad256((size_t)B * Nq * 32) + pad256((size_t)B * 4) + pad256((size_t)B * C * 4);
}
// host-pointer entry points: q_n and cand are host arrays, so bad values are rejected here instead of indexing
// kf_start / kf_count out of bounds on the device
int nclt_check_host_lists(nclt_ctx* c, const nclt_lib* L, const int32_t* q_n, int B, int Nq, const int32_t* cand, int C) {
    if (q_n)
        for (int b = 0; b < B; ++b)
            if (q_n[b] < 0 || q_n[b] > Nq) return nclt_fail(c, NCLT_ERR_ARG, "q_n entry outside [0, Nq]");
    if (cand) {
        const long long n = (long long)B * C;
        for (long long i = 0; i < n; ++i)
            if (cand[i] < -1 || cand[i] >= L->n_kf) return nclt_fail(c, NCLT_ERR_ARG, "cand entry outside [-1, n_kf)");
    }
    return NCLT_OK;
}
static int check_host(nclt_ctx* c, const nclt_lib* L, const void* q, int B, int Nq, int C, const int32_t* q_n = nullptr,
                      const int32_t* cand = nullptr) {
    if (!c || !L) return nclt_fail(c, NCLT_ERR_ARG, "null handle");
    if (B < 0 || Nq <= 0 || C <= 0 || (B > 0 && !q)) return nclt_fail(c, NCLT_ERR_ARG, "bad B/Nq/C/q");
    int rc = nclt_check_host_lists(c, L, q_n, B, Nq, cand, C);
    if (rc) return rc;
    cudaSetDevice(c->device);
    return NCLT_OK;
}

extern "C" int nclt_match_knn2(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq,
                               const int32_t* cand, int C, int32_t* out_idx, uint16_t* out_dist) {
    int rc = check_host(c, L, q, B, Nq, C, q_n, cand);
    if (rc) return rc;
    if (B == 0) return NCLT_OK;
    ScratchScope scope(c);
    size_t rows = (size_t)B * C * Nq;
    if ((rc = nclt_scratch_reserve(c, stage_bytes(B, Nq, C) + pad256(rows * 8) + pad256(rows * 4)))) return rc;
    Carver cv(c);
    HostStage s;
    if ((rc = stage_inputs(c, cv, q, q_n, B, Nq, cand, C, s))) return rc;
    int32_t* d_idx = cv.take<int32_t>(rows * 2);
    uint16_t* d_dist = cv.take<uint16_t>(rows * 2);
    if ((rc = nclt_match_knn2_dev(c, L, s.d_q, s.d_qn, B, Nq, s.d_cand, C, d_idx, d_dist))) return rc;
    if (out_idx) CU_TRY(c, cudaMemcpyAsync(out_idx, d_idx, rows * 8, cudaMemcpyDeviceToHost, c->stream));
    if (out_dist) CU_TRY(c, cudaMemcpyAsync(out_dist, d_dist, rows * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}

extern "C" int nclt_match_ratio(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq,
                                const int32_t* cand, int C, int num, int den, int32_t* out_pairs, int32_t* out_n) {
    int rc = check_host(c, L, q, B, Nq, C, q_n, cand);
    if (rc) return rc;
    if (!out_pairs || !out_n) return nclt_fail(c, NCLT_ERR_ARG, "ratio outputs null");
    if (B == 0) return NCLT_OK;
    ScratchScope scope(c);
    size_t items = (size_t)B * C, rows = items * Nq;
    if ((rc = nclt_scratch_reserve(c, stage_bytes(B, Nq, C) + pad256(rows * 8) + pad256(items * 4) +
                                          pad256(rows * sizeof(uint2)))))
        return rc;
    Carver cv(c);
    HostStage s;
    if ((rc = stage_inputs(c, cv, q, q_n, B, Nq, cand, C, s))) return rc;
    int32_t* d_pairs = cv.take<int32_t>(rows * 2);
    int32_t* d_n = cv.take<int32_t>(items);
    if ((rc = nclt_match_ratio_dev(c, L, s.d_q, s.d_qn, B, Nq, s.d_cand, C, num, den, d_pairs, d_n))) return rc;
    if (c->engine == 2 && !cand && nclt_overflow_take(c) > 0) {
        // the fused fp4 path ran out of candidate capacity (pathological input: most pairs pass the bound test):
        // this entry point is synchronous, so redo the batch on the integer engine - same results
        c->engine = 0;
        rc = nclt_match_ratio_dev(c, L, s.d_q, s.d_qn, B, Nq, s.d_cand, C, num, den, d_pairs, d_n);
        c->engine = 2;
        if (rc) return rc;
    }
    CU_TRY(c, cudaMemcpyAsync(out_pairs, d_pairs, rows * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(out_n, d_n, items * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}

extern "C" int nclt_match_cross(nclt_ctx* c, const nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq,
                                const int32_t* cand, int C, int Nmax, int32_t* out_pairs, uint16_t* out_dist,
                                int32_t* out_n) {
    int rc = check_host(c, L, q, B, Nq, C, q_n, cand);
    if (rc) return rc;
    if (!out_pairs || !out_n || Nmax <= 0) return nclt_fail(c, NCLT_ERR_ARG, "cross outputs null");
    if (B == 0) return NCLT_OK;
    ScratchScope scope(c);
    size_t items = (size_t)B * C, rows = items * Nmax;
    if ((rc = nclt_scratch_reserve(c, stage_bytes(B, Nq, C) + pad256(rows * 8) + pad256(rows * 2) +
                                          pad256(items * 4) + pad256(rows * sizeof(uint2)) +
                                          pad256(items * Nq * sizeof(uint2)))))
        return rc;
    Carver cv(c);
    HostStage s;
    if ((rc = stage_inputs(c, cv, q, q_n, B, Nq, cand, C, s))) return rc;
    int32_t* d_pairs = cv.take<int32_t>(rows * 2);
    uint16_t* d_dist = cv.take<uint16_t>(rows);
    int32_t* d_n = cv.take<int32_t>(items);
    if ((rc = nclt_match_cross_dev(c, L, s.d_q, s.d_qn, B, Nq, s.d_cand, C, Nmax, d_pairs, d_dist, d_n))) return rc;
    CU_TRY(c, cudaMemcpyAsync(out_pairs, d_pairs, rows * 8, cudaMemcpyDeviceToHost, c->stream));
    if (out_dist) CU_TRY(c, cudaMemcpyAsync(out_dist, d_dist, rows * 2, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(out_n, d_n, items * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}
