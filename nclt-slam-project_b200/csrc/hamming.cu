// K1 / K1' / K2: 256-bit Hamming brute-force matching on the integer pipe (LOP3 + POPC),
// with a packed (dist, index) top-2 so that min() on the key is the reference's tie rule.
//
// Replaces:
//   a1  cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2)      checkpoint_a_selftest.py:46,68
//   a2  Lowe ratio  m.distance < 0.80 * n.distance            checkpoint_a_selftest.py:71
//   a3  cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match    visual_landmark_matcher.py:211,327
//
// Layout: a descriptor is 32 B = 2 x uint4. The "B" rows (train side) of one work item are
// staged into shared memory by TMA bulk copies (cp.async.bulk, completion on an mbarrier,
// two stages); each thread keeps R "A" rows (query side) in registers, loaded with 128-bit
// coalesced loads, and streams the staged rows with broadcast LDS.128.
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int TB = 256;      // threads per CTA
constexpr int CHUNK = 512;   // B rows per stage (16 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct KParams {
    SegView A, B;
    const int* cand;
    int C, swap, a_rows_max, nsplit, b_seg_fixed;
    uint32_t idx_offset;     // added to every B row index (shard offset in flat mode)
    uint2* out_keys;
    int2* out_idx;
    ushort2* out_dist;
};

template <int R>
__global__ void __launch_bounds__(TB) k_hamming_top2(KParams p) {
    __shared__ __align__(128) uint4 sB[2][CHUNK * 2];
    __shared__ __align__(8) uint64_t mbar[2];

    const int item = blockIdx.x;
    const int split = blockIdx.y;
    const int achunk = blockIdx.z;
    const int tid = threadIdx.x;
    const int b = item / p.C;
    const int cand = p.b_seg_fixed >= 0 ? p.b_seg_fixed : (p.cand ? p.cand[item] : item % p.C);

    const int row0 = achunk * (TB * R);
    const size_t out_base = (size_t)(item * p.nsplit + split) * p.a_rows_max;

    uint32_t m1[R], m2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) m1[r] = m2[r] = NCLT_KEY_INVALID;

    int a_cnt = 0;
    if (cand >= 0) {
        const int a_seg = p.swap ? cand : b;
        const int b_seg = p.swap ? b : cand;
        const int a_start = p.A.start ? p.A.start[a_seg] : a_seg * p.A.stride;
        a_cnt = p.A.count ? p.A.count[a_seg] : p.A.stride;
        const int b_start = p.B.start ? p.B.start[b_seg] : b_seg * p.B.stride;
        const int b_cnt = p.B.count ? p.B.count[b_seg] : p.B.stride;

        // this CTA's slice of the B rows
        int per = (b_cnt + p.nsplit - 1) / p.nsplit;
        int lo = split * per;
        int hi = min(b_cnt, lo + per);
        const int nrows = max(0, hi - lo);
        // register rows of this CTA that hold A rows (a frame may have far fewer rows than the stride the launch was
        // sized for: ORB returns <= 500 keypoints in a 640-row buffer); the pass below runs with those only
        const int live = min(R, (a_cnt - row0 + TB - 1) / TB);
        const int nch = live > 0 ? (nrows + CHUNK - 1) / CHUNK : 0;
        const uint4* bsrc = p.B.base + (size_t)(b_start + lo) * 2;

        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && nch > 0) {
            uint32_t bytes = (uint32_t)min(CHUNK, nrows) * 32u;
            mbar_expect_tx(&mbar[0], bytes);
            tma_bulk_g2s(&sB[0][0], bsrc, bytes, &mbar[0]);
        }

        // A rows of this thread: 2 x 128-bit loads, consecutive threads -> consecutive rows
        uint4 q0[R], q1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int row = row0 + r * TB + tid;
            if (row < a_cnt) {
                const uint4* src = p.A.base + (size_t)(a_start + row) * 2;
                q0[r] = __ldg(src);
                q1[r] = __ldg(src + 1);
            } else {
                q0[r] = make_uint4(0, 0, 0, 0);
                q1[r] = make_uint4(0, 0, 0, 0);
            }
        }

        auto pass = [&](auto rl) {
            constexpr int RL = decltype(rl)::value;
            for (int ch = 0; ch < nch; ++ch) {
                const int st = ch & 1;
                if (tid == 0 && ch + 1 < nch) {
                    const int nst = st ^ 1;
                    uint32_t bytes = (uint32_t)min(CHUNK, nrows - (ch + 1) * CHUNK) * 32u;
                    mbar_expect_tx(&mbar[nst], bytes);
                    tma_bulk_g2s(&sB[nst][0], bsrc + (size_t)(ch + 1) * CHUNK * 2, bytes, &mbar[nst]);
                }
                mbar_wait(&mbar[st], (ch >> 1) & 1);
                const int n = min(CHUNK, nrows - ch * CHUNK);
                const uint4* sb = &sB[st][0];
                uint32_t jkey = p.idx_offset + (uint32_t)(lo + ch * CHUNK);
    #pragma unroll 2
                for (int j = 0; j < n; ++j, ++jkey) {
                    const uint4 t0 = sb[2 * j];
                    const uint4 t1 = sb[2 * j + 1];
    #pragma unroll
                    for (int r = 0; r < RL; ++r) {
                        uint32_t d = __popc(q0[r].x ^ t0.x) + __popc(q0[r].y ^ t0.y) + __popc(q0[r].z ^ t0.z) +
                                     __popc(q0[r].w ^ t0.w) + __popc(q1[r].x ^ t1.x) + __popc(q1[r].y ^ t1.y) +
                                     __popc(q1[r].z ^ t1.z) + __popc(q1[r].w ^ t1.w);
                        uint32_t key = (d << NCLT_KEY_SHIFT) + jkey;
                        uint32_t mx = max(m1[r], key);
                        m1[r] = min(m1[r], key);
                        m2[r] = min(m2[r], mx);
                    }
                }
                __syncthreads();   // everyone is done with stage st before it is refilled
            }
        };
        if (live >= R) pass(std::integral_constant<int, R>{});
        else if (R > 3 && live == 3) pass(std::integral_constant<int, (R > 3 ? 3 : R)>{});
        else if (R > 2 && live == 2) pass(std::integral_constant<int, (R > 2 ? 2 : R)>{});
        else if (live == 1) pass(std::integral_constant<int, 1>{});
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
        int row = row0 + r * TB + tid;
        if (row >= p.a_rows_max) continue;
        bool live = row < a_cnt;
        uint32_t k1 = live ? m1[r] : NCLT_KEY_INVALID;
        uint32_t k2 = live ? m2[r] : NCLT_KEY_INVALID;
        if (p.out_keys) p.out_keys[out_base + row] = make_uint2(k1, k2);
        if (p.out_idx) {
            int i1 = k1 == NCLT_KEY_INVALID ? -1 : (int)(k1 & NCLT_KEY_IDX_MASK);
            int i2 = k2 == NCLT_KEY_INVALID ? -1 : (int)(k2 & NCLT_KEY_IDX_MASK);
            p.out_idx[out_base + row] = make_int2(i1, i2);
        }
        if (p.out_dist) {
            unsigned short d1 = k1 == NCLT_KEY_INVALID ? 65535 : (unsigned short)(k1 >> NCLT_KEY_SHIFT);
            unsigned short d2 = k2 == NCLT_KEY_INVALID ? 65535 : (unsigned short)(k2 >> NCLT_KEY_SHIFT);
            p.out_dist[out_base + row] = make_ushort2(d1, d2);
        }
    }
}

// K1' fused: both directions of crossCheck from ONE pass over the distance matrix of an item (frame x candidate keyframe).
// A = frame rows in registers (R per thread), B = the keyframe's rows streamed through shared memory as above.
//   bwd[item][frame row].x = min over teach rows of (d << 23 | teach row)   - the thread's running minimum
//   fwd[item][teach row].x = min over frame rows of (d << 23 | frame row)   - per teach row: minimum over the thread's R
//                            rows, REDUX.MIN over the warp, one shared-memory slot per (warp, teach row), the CTA's 8
//                            slots folded when the chunk is through, atomicMin into fwd (preset to INVALID; several
//                            CTAs when the frame has more than TB * R rows)
// min() on the packed key is the reference's tie rule in both directions (lowest index).  Half the POPC work of the
// two directional launches it replaces (visual_landmark_matcher.py:327 runs this against <= 5 candidates per tick).
constexpr int XCHUNK = 256;
struct CrossShared {
    uint4 sB[2][XCHUNK * 2];
    uint32_t s_col[TB / 32][XCHUNK];
    uint64_t mbar[2];
};

// the pass over the keyframe's rows with RL of the thread's R register rows in use (the frame may have far fewer rows
// than the row stride Nq the launch was sized for: ORB returns <= 500 keypoints in a 512-row buffer)
template <int R, int RL>
__device__ __forceinline__ void cross_pass(CrossShared& sh, const uint4* __restrict__ bsrc, int nrows, int nch, const uint4 (&q0)[R],
                                           const uint4 (&q1)[R], const uint32_t (&rowkey)[R], uint32_t (&m1)[R], uint2* fwd_item) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int ch = 0; ch < nch; ++ch) {
        const int st = ch & 1;
        if (tid == 0 && ch + 1 < nch) {
            const int nst = st ^ 1;
            uint32_t bytes = (uint32_t)min(XCHUNK, nrows - (ch + 1) * XCHUNK) * 32u;
            mbar_expect_tx(&sh.mbar[nst], bytes);
            tma_bulk_g2s(&sh.sB[nst][0], bsrc + (size_t)(ch + 1) * XCHUNK * 2, bytes, &sh.mbar[nst]);
        }
        mbar_wait(&sh.mbar[st], (ch >> 1) & 1);
        const int n = min(XCHUNK, nrows - ch * XCHUNK);
        const uint4* sb = &sh.sB[st][0];
        uint32_t jkey = (uint32_t)(ch * XCHUNK);
#pragma unroll 2
        for (int j = 0; j < n; ++j, ++jkey) {
            const uint4 t0 = sb[2 * j];
            const uint4 t1 = sb[2 * j + 1];
            uint32_t col = NCLT_KEY_INVALID;
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                const uint32_t d = __popc(q0[r].x ^ t0.x) + __popc(q0[r].y ^ t0.y) + __popc(q0[r].z ^ t0.z) +
                                   __popc(q0[r].w ^ t0.w) + __popc(q1[r].x ^ t1.x) + __popc(q1[r].y ^ t1.y) +
                                   __popc(q1[r].z ^ t1.z) + __popc(q1[r].w ^ t1.w);
                const uint32_t dk = d << NCLT_KEY_SHIFT;
                m1[r] = min(m1[r], dk + jkey);
                col = min(col, dk | rowkey[r]);
            }
            col = __reduce_min_sync(0xFFFFFFFFu, col);
            if (lane == 0) sh.s_col[warp][j] = col;
        }
        __syncthreads();
        if (tid < n) {
            uint32_t v = sh.s_col[0][tid];
#pragma unroll
            for (int w = 1; w < TB / 32; ++w) v = min(v, sh.s_col[w][tid]);
            if (v != NCLT_KEY_INVALID) atomicMin(&fwd_item[ch * XCHUNK + tid].x, v);
        }
        __syncthreads();   // stage st and s_col are free again
    }
}

template <int R>
__global__ void __launch_bounds__(TB) k_hamming_cross(KParams p, uint2* __restrict__ fwd, int fwd_rows_max) {
    __shared__ __align__(128) CrossShared sh;
    const int item = blockIdx.x, achunk = blockIdx.z, tid = threadIdx.x;
    const int b = item / p.C;
    const int cand = p.cand ? p.cand[item] : item % p.C;
    const int row0 = achunk * (TB * R);
    const size_t out_base = (size_t)item * p.a_rows_max;
    uint32_t m1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) m1[r] = NCLT_KEY_INVALID;
    int a_cnt = 0;
    if (cand >= 0) {
        const int a_start = p.A.start ? p.A.start[b] : b * p.A.stride;
        a_cnt = p.A.count ? p.A.count[b] : p.A.stride;
        const int b_start = p.B.start ? p.B.start[cand] : cand * p.B.stride;
        const int nrows = p.B.count ? p.B.count[cand] : p.B.stride;
        const int live = min(R, (a_cnt - row0 + TB - 1) / TB);       // register rows of this CTA that hold frame rows
        const int nch = live > 0 ? (nrows + XCHUNK - 1) / XCHUNK : 0;
        const uint4* bsrc = p.B.base + (size_t)b_start * 2;
        if (tid == 0) {
            mbar_init(&sh.mbar[0], 1);
            mbar_init(&sh.mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && nch > 0) {
            uint32_t bytes = (uint32_t)min(XCHUNK, nrows) * 32u;
            mbar_expect_tx(&sh.mbar[0], bytes);
            tma_bulk_g2s(&sh.sB[0][0], bsrc, bytes, &sh.mbar[0]);
        }
        uint4 q0[R], q1[R];
        uint32_t rowkey[R];          // the frame row, or all ones for rows past the frame's end (their keys stay INVALID)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r * TB + tid;
            if (row < a_cnt) {
                const uint4* src = p.A.base + (size_t)(a_start + row) * 2;
                q0[r] = __ldg(src);
                q1[r] = __ldg(src + 1);
                rowkey[r] = (uint32_t)row;
            } else {
                q0[r] = make_uint4(0, 0, 0, 0);
                q1[r] = make_uint4(0, 0, 0, 0);
                rowkey[r] = NCLT_KEY_INVALID;
            }
        }
        uint2* fwd_item = fwd + (size_t)item * fwd_rows_max;
        if (R >= 4 && live == 4) cross_pass<R, (R >= 4 ? 4 : R)>(sh, bsrc, nrows, nch, q0, q1, rowkey, m1, fwd_item);
        else if (R >= 3 && live == 3) cross_pass<R, (R >= 3 ? 3 : R)>(sh, bsrc, nrows, nch, q0, q1, rowkey, m1, fwd_item);
        else if (R >= 2 && live == 2) cross_pass<R, (R >= 2 ? 2 : R)>(sh, bsrc, nrows, nch, q0, q1, rowkey, m1, fwd_item);
        else if (live == 1) cross_pass<R, 1>(sh, bsrc, nrows, nch, q0, q1, rowkey, m1, fwd_item);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * TB + tid;
        if (row >= p.a_rows_max) continue;
        p.out_keys[out_base + row] = make_uint2(row < a_cnt ? m1[r] : NCLT_KEY_INVALID, NCLT_KEY_INVALID);
    }
}

// Merge `nparts` partial top-2 lists per (item,row) into one (flat mode: B-row splits on one
// GPU, then library shards gathered from all ranks).
__global__ void k_merge_top2(const uint2* __restrict__ parts, int n_items, int nparts, int rows,
                             long long part_stride, long long item_stride, uint2* out_keys, int2* out_idx,
                             ushort2* out_dist) {
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_items * rows;
    if (g >= total) return;
    int item = (int)(g / rows), row = (int)(g % rows);
    uint32_t m1 = NCLT_KEY_INVALID, m2 = NCLT_KEY_INVALID;
    for (int s = 0; s < nparts; ++s) {
        uint2 k = parts[(long long)s * part_stride + (long long)item * item_stride + row];
        uint32_t mx = max(m1, k.x);
        m1 = min(m1, k.x);
        m2 = min(m2, mx);
        mx = max(m1, k.y);
        m1 = min(m1, k.y);
        m2 = min(m2, mx);
    }
    if (out_keys) out_keys[g] = make_uint2(m1, m2);
    if (out_idx)
        out_idx[g] = make_int2(m1 == NCLT_KEY_INVALID ? -1 : (int)(m1 & NCLT_KEY_IDX_MASK),
                               m2 == NCLT_KEY_INVALID ? -1 : (int)(m2 & NCLT_KEY_IDX_MASK));
    if (out_dist)
        out_dist[g] = make_ushort2(m1 == NCLT_KEY_INVALID ? 65535 : (unsigned short)(m1 >> NCLT_KEY_SHIFT),
                                   m2 == NCLT_KEY_INVALID ? 65535 : (unsigned short)(m2 >> NCLT_KEY_SHIFT));
}

// ordered block compaction step: returns this thread's output slot (valid when flag) and
// advances `base` by the number of flagged threads. All TB threads must call.
__device__ __forceinline__ int compact_slot(bool flag, int& base, int* s_warp) {
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < TB / 32; ++w) {
        int v = s_warp[w];
        total += v;
        if (w < warp) before += v;
    }
    int slot = base + before + __popc(bal & ((1u << lane) - 1u));
    base += total;
    __syncthreads();
    return slot;
}

// K2 (a2): Lowe ratio on the per-keyframe top-2 + ordered compaction -> (a_row, b_row) pairs.
__global__ void __launch_bounds__(TB) k_ratio_compact(const uint2* __restrict__ keys, const int* __restrict__ a_count,
                                                      int a_stride_cnt, const int* __restrict__ cand, int C,
                                                      int a_rows_max, int num, int den, int2* out_pairs,
                                                      int* out_n) {
    __shared__ int s_warp[TB / 32];
    const int item = blockIdx.x;
    const int b = item / C;
    const int cnd = cand ? cand[item] : 0;
    int a_cnt = cnd < 0 ? 0 : (a_count ? a_count[b] : a_stride_cnt);
    int base = 0;
    for (int r0 = 0; r0 < a_cnt; r0 += TB) {
        int row = r0 + threadIdx.x;
        bool keep = false;
        uint32_t k1 = 0;
        if (row < a_cnt) {
            uint2 k = keys[(size_t)item * a_rows_max + row];
            k1 = k.x;
            if (k.y != NCLT_KEY_INVALID) {
                int d1 = (int)(k.x >> NCLT_KEY_SHIFT), d2 = (int)(k.y >> NCLT_KEY_SHIFT);
                keep = den * d1 < num * d2;
            }
        }
        int slot = compact_slot(keep, base, s_warp);
        if (keep) out_pairs[(size_t)item * a_rows_max + slot] = make_int2(row, (int)(k1 & NCLT_KEY_IDX_MASK));
    }
    if (threadIdx.x == 0) out_n[item] = base;
}

// K1' (a3): mutual nearest neighbours from the two directional top-1 results.
// fwd: rows = library (teach) descriptors, idx into the frame; bwd: rows = frame, idx into teach.
// Emits (teach_row, frame_row) in increasing teach_row = cv2 crossCheck output order.
__global__ void __launch_bounds__(TB) k_cross_combine(const uint2* __restrict__ fwd, const uint2* __restrict__ bwd,
                                                      SegView Lib, const int* __restrict__ cand, int C,
                                                      int fwd_rows_max, int bwd_rows_max, int2* out_pairs,
                                                      unsigned short* out_dist, int* out_n, int out_stride) {
    __shared__ int s_warp[TB / 32];
    const int item = blockIdx.x;
    const int cnd = cand ? cand[item] : item % C;
    int a_cnt = cnd < 0 ? 0 : (Lib.count ? Lib.count[cnd] : Lib.stride);
    int base = 0;
    for (int r0 = 0; r0 < a_cnt; r0 += TB) {
        int row = r0 + threadIdx.x;
        bool keep = false;
        int j = 0;
        unsigned short d = 0;
        if (row < a_cnt) {
            uint32_t k1 = fwd[(size_t)item * fwd_rows_max + row].x;
            if (k1 != NCLT_KEY_INVALID) {
                j = (int)(k1 & NCLT_KEY_IDX_MASK);
                d = (unsigned short)(k1 >> NCLT_KEY_SHIFT);
                uint32_t kb = bwd[(size_t)item * bwd_rows_max + j].x;
                keep = kb != NCLT_KEY_INVALID && (int)(kb & NCLT_KEY_IDX_MASK) == row;
            }
        }
        int slot = compact_slot(keep, base, s_warp);
        if (keep) {
            out_pairs[(size_t)item * out_stride + slot] = make_int2(row, j);
            if (out_dist) out_dist[(size_t)item * out_stride + slot] = d;
        }
    }
    if (threadIdx.x == 0) out_n[item] = base;
}

// Register-only POPC throughput probe: the roofline denominator for K1.
__global__ void __launch_bounds__(256) k_popc_peak(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u;
    uint32_t a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // 8 independent POPC chains; the feedback keeps the compiler from hoisting them
            s0 += __popc(a0 ^ s4); s1 += __popc(a1 ^ s5); s2 += __popc(a2 ^ s6); s3 += __popc(a3 ^ s7);
            s4 += __popc(a4 ^ s0); s5 += __popc(a5 ^ s1); s6 += __popc(a6 ^ s2); s7 += __popc(a7 ^ s3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}

}  // namespace

int launch_hamming_top2(nclt_ctx* c, const MatchLaunch& m, uint32_t idx_offset) {
    const int n_items = m.n_outer * m.C;
    if (n_items <= 0 || m.a_rows_max <= 0) return NCLT_OK;
    KParams p;
    p.A = m.A; p.B = m.B; p.cand = m.cand; p.C = m.C; p.swap = m.swap; p.a_rows_max = m.a_rows_max;
    p.nsplit = m.nsplit < 1 ? 1 : m.nsplit; p.b_seg_fixed = m.b_seg_fixed; p.idx_offset = idx_offset;
    p.out_keys = m.out_keys; p.out_idx = m.out_idx; p.out_dist = m.out_dist;
    // rows per thread: 4 when the A side is large, fewer for the small production shapes
    int R = m.a_rows_max > 2 * TB ? 4 : (m.a_rows_max > TB ? 2 : 1);
    int achunks = (m.a_rows_max + TB * R - 1) / (TB * R);
    dim3 grid(n_items, p.nsplit, achunks);
    nclt_prof_mark(c);
    if (R == 4) k_hamming_top2<4><<<grid, TB, 0, c->stream>>>(p);
    else if (R == 2) k_hamming_top2<2><<<grid, TB, 0, c->stream>>>(p);
    else k_hamming_top2<1><<<grid, TB, 0, c->stream>>>(p);
    nclt_prof_mark(c);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

// crossCheck against candidate lists: fwd[item][teach row] / bwd[item][frame row] top-1 keys from one pass (see the kernel)
int launch_hamming_cross(nclt_ctx* c, const SegView& frames, const SegView& lib, const int* cand, int n_outer, int C, int Nq,
                         int fwd_rows_max, uint2* fwd, uint2* bwd) {
    const int n_items = n_outer * C;
    if (n_items <= 0 || Nq <= 0) return NCLT_OK;
    KParams p{};
    p.A = frames; p.B = lib; p.cand = cand; p.C = C; p.swap = 0; p.a_rows_max = Nq; p.nsplit = 1; p.b_seg_fixed = -1;
    p.idx_offset = 0; p.out_keys = bwd;
    CU_TRY(c, cudaMemsetAsync(fwd, 0xFF, (size_t)n_items * fwd_rows_max * sizeof(uint2), c->stream));
    // rows per thread: 4 for 1000-row frames; 2 (and a second, usually empty, CTA per item) for ORB's 640-row buffers that
    // hold <= 500 keypoints - 40 registers instead of 64, six CTAs per SM
    const int R = Nq > 3 * TB ? 4 : (Nq > TB ? 2 : 1);
    dim3 grid(n_items, 1, (Nq + TB * R - 1) / (TB * R));
    nclt_prof_mark(c);
    if (R == 4) k_hamming_cross<4><<<grid, TB, 0, c->stream>>>(p, fwd, fwd_rows_max);
    else if (R == 2) k_hamming_cross<2><<<grid, TB, 0, c->stream>>>(p, fwd, fwd_rows_max);
    else k_hamming_cross<1><<<grid, TB, 0, c->stream>>>(p, fwd, fwd_rows_max);
    nclt_prof_mark(c);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

int launch_merge_top2(nclt_ctx* c, const uint2* parts, int n_items, int nparts, int rows, long long part_stride,
                      long long item_stride, uint2* out_keys, int2* out_idx, ushort2* out_dist) {
    long long total = (long long)n_items * rows;
    if (total <= 0) return NCLT_OK;
    int blocks = (int)((total + 255) / 256);
    k_merge_top2<<<blocks, 256, 0, c->stream>>>(parts, n_items, nparts, rows, part_stride, item_stride, out_keys,
                                                out_idx, out_dist);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

int launch_cross_combine(nclt_ctx* c, const uint2* fwd_keys, const uint2* bwd_keys, const SegView& Lib,
                         const int* cand, int n_outer, int C, int fwd_rows_max, int bwd_rows_max, int2* out_pairs,
                         unsigned short* out_dist, int* out_n, int out_stride) {
    int n_items = n_outer * C;
    if (n_items <= 0) return NCLT_OK;
    k_cross_combine<<<n_items, TB, 0, c->stream>>>(fwd_keys, bwd_keys, Lib, cand, C, fwd_rows_max, bwd_rows_max,
                                                   out_pairs, out_dist, out_n, out_stride);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

int launch_ratio_compact(nclt_ctx* c, const uint2* keys, const int* a_count, int a_stride_cnt, const int* cand,
                         int n_outer, int C, int a_rows_max, int num, int den, int2* out_pairs, int* out_n) {
    int n_items = n_outer * C;
    if (n_items <= 0) return NCLT_OK;
    k_ratio_compact<<<n_items, TB, 0, c->stream>>>(keys, a_count, a_stride_cnt, cand, C, a_rows_max, num, den,
                                                   out_pairs, out_n);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

// returns POPC32 operations per second measured with a register-only kernel
double run_popc_peak(nclt_ctx* c, int iters, float* ms_out) {
    const int blocks = c->sm_count * 8, threads = 256;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * threads * 4) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_popc_peak<<<blocks, threads, 0, c->stream>>>(d, iters / 8 + 1, 1u);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        k_popc_peak<<<blocks, threads, 0, c->stream>>>(d, iters, 12345u + rep);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    c->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (ms_out) *ms_out = best;
    double ops = (double)blocks * threads * (double)iters * 64.0;
    return ops / (best * 1e-3);
}
