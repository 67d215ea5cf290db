// Tensor-core Hamming top-2 for the "every frame against every keyframe" replay workloads
// (BASELINE.json configs 1, 2, 4; checkpoint_a_selftest.py:68-71 semantics per keyframe).
//
// The integer-pipe kernel (hamming.cu) sits at 99 % of the POPC roofline (2 comparisons/clk/SM).
// sm_100a has no native 1-bit MMA (SURVEY finding 6), so distances go through tcgen05 as
//     x . y = 256 - 2 * Hamming(a, b),   x, y in {+1,-1}^256 stored as fp8 e4m3,
// with fp16 accumulators in TMEM (exact: |sum| <= 256).  Measured building blocks (tools/tcbench.py):
// 32 comparisons/clk/SM of MMA at N = 256, packed 16-bit TMEM loads almost free, half2 top-2 tracking
// at 1.5 ALU ops per comparison.
//
// Only VALUES (best, second best) per (query, keyframe) come out of this kernel: the Lowe ratio test
// needs nothing else, and the train index of the (rare) survivors is recovered exactly, lowest index
// first, by an integer re-scan of that one keyframe (k_tc_ratio_recover).  Results are bit-identical
// to the integer path (tests/test_tc_match_gpu.py).
//
// Kernel k_tc_top2: persistent, warp specialised, one CTA per SM:
//   warps 0-7  epilogue: TMEM -> registers (tcgen05.ld .pack::16b), half2 running top-2 per row;
//              warp w owns TMEM lane quadrant w%4 and column half w/4
//   warp 8     producer: 1-D TMA bulk copies of pre-expanded operand tile images
//   warp 9     MMA issuer: 8 x tcgen05.mma (K = 32 B each) per 128 x N x 256 tile step
// A work item = MA (<= 3) resident 128-row query tiles x a contiguous range of library tiles.
#include "common.cuh"
#include "scratch.cuh"
#include "tc_common.cuh"
#include "tc_tiles.h"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <vector>

using namespace nclt_tc4;

namespace {

constexpr int MA = 3;                    // resident query tiles per work item
constexpr int A_TILE_BYTES = 128 * 256;  // 32 KB
constexpr int B_STAGE_BYTES = 256 * 256; // 64 KB
constexpr int NSTAGE = 2;
constexpr int TC_THREADS = 320;
constexpr uint32_t NEG_INF2 = 0xFC00FC00u;   // half2(-inf, -inf)
__device__ __forceinline__ __half2 neg_inf2() {
    uint32_t v = 0xFC00FC00u;
    return *reinterpret_cast<__half2*>(&v);
}

struct LibTile {
    uint32_t img_off256;   // byte offset / 256 into the library image buffer
    uint16_t n;            // rows in the tile image (multiple of 16, <= 256)
    uint16_t n_valid;      // real descriptor rows (<= n)
    int kf;                // keyframe id
    int last_of_kf;        // 1 if this is the keyframe's last tile
};

// ---- operand expansion -------------------------------------------------------------------
// rows: 32-byte descriptors; tile t covers rows [t*128, t*128+128) -> 32 KB image
__global__ void k_expand_queries(const uint32_t* __restrict__ desc, long long n_rows, uint8_t* img) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, chunk)
    long long total_rows = ((n_rows + 127) / 128) * 128;
    if (i >= total_rows * 16) return;
    long long row = i >> 4;
    int c = (int)(i & 15);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < n_rows) {
        uint32_t w = desc[row * 8 + (c >> 1)];
        v = tc::expand16((c & 1) ? (w >> 16) : (w & 0xFFFFu));
    }
    long long tile = row >> 7;
    int r = (int)(row & 127);
    *reinterpret_cast<uint4*>(img + tile * A_TILE_BYTES + tc::image_offset(128, r, c * 16)) = v;
}

__global__ void k_expand_library(const uint32_t* __restrict__ desc, const LibTile* __restrict__ tiles,
                                 const int* __restrict__ tile_row0, int n_tiles, uint8_t* img) {
    int t = blockIdx.x;
    if (t >= n_tiles) return;
    const LibTile lt = tiles[t];
    const long long row0 = tile_row0[t];
    uint8_t* dst = img + (size_t)lt.img_off256 * 256;
    for (int i = threadIdx.x; i < lt.n * 16; i += blockDim.x) {
        int r = i >> 4, c = i & 15;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < lt.n_valid) {
            uint32_t w = desc[(row0 + r) * 8 + (c >> 1)];
            v = tc::expand16((c & 1) ? (w >> 16) : (w & 0xFFFFu));
        }
        *reinterpret_cast<uint4*>(dst + tc::image_offset(lt.n, r, c * 16)) = v;
    }
}

// ---- main kernel ---------------------------------------------------------------------------
struct TcParams {
    const uint8_t* q_img;      // [n_mtiles][32 KB]
    const uint8_t* lib_img;
    const LibTile* tiles;
    int n_mtiles;
    int n_groups;              // ceil(n_mtiles / MA)
    int n_splits;
    const int* split_tile;     // [n_splits + 1] tile ranges (aligned to keyframes)
    long long rows_total;      // valid query rows (B * Nq)
    uint32_t* out;             // [n_kf][rows_pad] : d1 | d2 << 16
    long long rows_pad;        // n_mtiles * 128
    unsigned long long* clk;   // diagnostics: [0] max SM cycles, [1] max globaltimer ns of a CTA (nclt_ctx_tc_clock)
    int clk_slot;              // [16 + 2 * slot] / [17 + 2 * slot]: earliest CTA start / latest CTA end of this launch (globaltimer ns)
};

// Epilogue primitive. Exact running top-2 costs 3 half2 min/max per register and made the
// epilogue (ALU pipe) the bottleneck at 75 % tensor utilisation. Instead each thread keeps the
// plain maximum of two DISJOINT column subsets (even / odd registers); with the two half2 lanes
// and the two column-half warps that is 8 disjoint subsets per (row, keyframe).  Their largest
// value is the exact best; their second largest is a LOWER bound of the second-best accumulator,
// i.e. an UPPER bound d2_bound >= d2 of the second-best distance.  The ratio test can only pass
// if 5*d1 < 4*d2_bound; those rare rows are then re-evaluated exactly (k_tc_ratio_recover).
__device__ __forceinline__ void half2_max(uint32_t raw, __half2& M) {
    M = __hmax2(M, *reinterpret_cast<__half2*>(&raw));
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_top2(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                  // MA x 32 KB
    uint8_t* sB = smem + MA * A_TILE_BYTES;              // NSTAGE x 64 KB
    uint8_t* tail = sB + NSTAGE * B_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // 10 barriers
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;      // [2]
    uint64_t* b_empty = bars + 4;     // [2]
    uint64_t* acc_full = bars + 6;    // [2]
    uint64_t* acc_empty = bars + 8;   // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 10);
    uint32_t* xchg = s_tmem + 2;      // [4][32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(a_full, 1);
        tc::mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], 1);
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 8);
        }
        tc::mbar_fence_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(s_tmem, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const int n_items = p.n_groups * p.n_splits;
#ifdef NCLT_TC_TRACE
#define TR(st_, k_) { const uint32_t s__ = (st_) - 2000u; if (s__ < 128u && blockIdx.x == 0 && p.clk) p.clk[64 + s__ * 8 + (k_)] = clock64(); }
#else
#define TR(st_, k_)
#endif
    const long long clk0 = clock64();
    unsigned long long ns0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));

    if (warp == 8) {
        // =========================== producer ===========================
        if (lane == 0) {
            uint32_t it_cnt = 0, bt = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it_cnt) {
                const int split = item / p.n_groups, group = item % p.n_groups;
                const int m0 = group * MA;
                const int ma = min(MA, p.n_mtiles - m0);
                tc::mbar_wait(a_empty, (it_cnt & 1) ^ 1);
                tc::mbar_expect_tx(a_full, (uint32_t)ma * A_TILE_BYTES);
                for (int m = 0; m < ma; ++m)
                    tc::bulk_g2s(sA + m * A_TILE_BYTES, p.q_img + (size_t)(m0 + m) * A_TILE_BYTES, A_TILE_BYTES, a_full);
                for (int t = p.split_tile[split]; t < p.split_tile[split + 1]; ++t, ++bt) {
                    const int s = bt & 1;
                    const LibTile lt = p.tiles[t];
                    tc::mbar_wait(&b_empty[s], ((bt >> 1) & 1) ^ 1);
                    const uint32_t bytes = (uint32_t)lt.n * 256u;
                    tc::mbar_expect_tx(&b_full[s], bytes);
                    tc::bulk_g2s(sB + s * B_STAGE_BYTES, p.lib_img + (size_t)lt.img_off256 * 256, bytes, &b_full[s]);
                }
            }
        }
    } else if (warp == 9) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t it_cnt = 0, bt = 0, st = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it_cnt) {
                const int split = item / p.n_groups, group = item % p.n_groups;
                const int ma = min(MA, p.n_mtiles - group * MA);
                tc::mbar_wait(a_full, it_cnt & 1);
                for (int t = p.split_tile[split]; t < p.split_tile[split + 1]; ++t, ++bt) {
                    const int s = bt & 1;
                    const int n = p.tiles[t].n;
                    tc::mbar_wait(&b_full[s], (bt >> 1) & 1);
                    tc::tc_fence_after();
                    const uint32_t idesc = tc::idesc_f8(128, n, 0);
                    const uint32_t lboB = (uint32_t)n * 16u;
                    const uint32_t bbase = tc::smem_u32(sB + s * B_STAGE_BYTES);
                    for (int m = 0; m < ma; ++m, ++st) {
                        const int buf = st & 1;
                        tc::mbar_wait(&acc_empty[buf], ((st >> 1) & 1) ^ 1);
                        tc::tc_fence_after();
                        const uint32_t abase = tc::smem_u32(sA + m * A_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            uint64_t da = tc::smem_desc(abase + k * 4096u, 2048u, 128u);
                            uint64_t db = tc::smem_desc(bbase + k * 2u * lboB, lboB, 128u);
                            tc::mma_f8(tmem + buf * 256, da, db, idesc, k > 0 ? 1u : 0u);
                        }
                        tc::mma_commit(&acc_full[buf]);
                    }
                    tc::mma_commit(&b_empty[s]);       // stage reusable once these MMAs have read it
                }
                tc::mma_commit(a_empty);
            }
        }
    } else {
        // =========================== epilogue (warps 0-7) ===========================
        const int quad = warp & 3, half = warp >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t st = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int split = item / p.n_groups, group = item % p.n_groups;
            const int m0 = group * MA;
            const int ma = min(MA, p.n_mtiles - m0);
            __half2 M1[MA], M2[MA];
#pragma unroll
            for (int m = 0; m < MA; ++m) {
                M1[m] = neg_inf2();
                M2[m] = M1[m];
            }
            for (int t = p.split_tile[split]; t < p.split_tile[split + 1]; ++t) {
                const LibTile lt = p.tiles[t];
                const int c_lo = half * 128;
                const int nv = lt.n_valid;
#pragma unroll
                for (int m = 0; m < MA; ++m) {
                    if (m >= ma) break;
                    const int buf = st & 1;
                    tc::mbar_wait(&acc_full[buf], (st >> 1) & 1);
                    tc::tc_fence_after();
                    ++st;
                    if (c_lo < nv) {
                        uint32_t r0[32], r1[32];
                        const uint32_t taddr = tmem + buf * 256 + lane_base + (uint32_t)c_lo;
                        const bool second = c_lo + 64 < nv;
                        tc::tmem_ld32_pack16(taddr, r0);
                        if (second) tc::tmem_ld32_pack16(taddr + 64, r1);
                        tc::tmem_wait_ld();
                        if (c_lo + 64 <= nv) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) half2_max(r0[j], (j & 1) ? M2[m] : M1[m]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                int c = c_lo + 2 * j;
                                uint32_t v = r0[j];
                                if (c + 1 >= nv) v = (c >= nv) ? NEG_INF2 : ((v & 0xFFFFu) | 0xFC000000u);
                                half2_max(v, (j & 1) ? M2[m] : M1[m]);
                            }
                        }
                        if (second) {
                            if (c_lo + 128 <= nv) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) half2_max(r1[j], (j & 1) ? M2[m] : M1[m]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    int c = c_lo + 64 + 2 * j;
                                    uint32_t v = r1[j];
                                    if (c + 1 >= nv) v = (c >= nv) ? NEG_INF2 : ((v & 0xFFFFu) | 0xFC000000u);
                                    half2_max(v, (j & 1) ? M2[m] : M1[m]);
                                }
                            }
                        }
                    }
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
                }
                if (lt.last_of_kf) {
                    // ---- keyframe finished: merge even/odd lanes, the two column halves, write (d1,d2)
#pragma unroll
                    for (int m = 0; m < MA; ++m) {
                        if (m >= ma) break;
                        float2 a = __half22float2(M1[m]), b = __half22float2(M2[m]);
                        // top-2 of the four subset maxima {a.x, a.y, b.x, b.y}
                        float h1 = fmaxf(a.x, a.y), l1 = fminf(a.x, a.y);
                        float h2 = fmaxf(b.x, b.y), l2 = fminf(b.x, b.y);
                        float f1 = fmaxf(h1, h2);
                        float f2 = fmaxf(fminf(h1, h2), h1 >= h2 ? l1 : l2);
                        if (half == 1) {
                            __half2 pk = __floats2half2_rn(f1, f2);
                            xchg[quad * 32 + lane] = *reinterpret_cast<uint32_t*>(&pk);
                        }
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
                        if (half == 0) {
                            uint32_t raw = xchg[quad * 32 + lane];
                            float2 o = __half22float2(*reinterpret_cast<__half2*>(&raw));
                            float g1 = fmaxf(f1, o.x);
                            float g2 = fmaxf(fminf(f1, o.x), fmaxf(f2, o.y));
                            long long row = (long long)(m0 + m) * 128 + quad * 32 + lane;
                            uint32_t d1 = (uint32_t)((256.f - g1) * 0.5f);
                            uint32_t d2 = g2 < -300.f ? 0xFFFFu : (uint32_t)((256.f - g2) * 0.5f);
                            if (g1 < -300.f) d1 = 0xFFFFu;
                            p.out[(size_t)lt.kf * p.rows_pad + row] = d1 | (d2 << 16);   // (exact d1, upper bound of d2)
                        }
                        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
                        M1[m] = neg_inf2();
                        M2[m] = M1[m];
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0 && p.clk) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        atomicMax(p.clk, (unsigned long long)(clock64() - clk0));
        atomicMax(p.clk + 1, ns1 - ns0);
        atomicMin(p.clk + 16 + 2 * p.clk_slot, ns0);
        atomicMax(p.clk + 17 + 2 * p.clk_slot, ns1);
    }
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ==============================================================================================
// Block-scaled fp4 flavour (kind::mxf4, K = 64 per instruction): twice the MMA rate of fp8, +-1.0 as e2m1 nibbles,
// f32 accumulators (the only accumulator type of the block-scaled kinds; exact: |sum| <= 256).
//
// Round-2 design (k_tc4_top2).  The round-1 kernel read 32-bit accumulator cells (two batches of 120 registers per
// 240-column step, two TMEM round trips while the buffer is held) and spent ~800 clk per step against 480 clk of MMA.
// Three changes:
//   1. The bias comes out of the tensor core: a fifth K = 64 step per tile whose operands are constant rows
//      (tc_common.cuh, mx_bias_byte / MX_BIAS_SFA) adds 1.5 * 2^23 + 0x4000 to every cell, so the LOW 16 bits of the f32
//      cell are 0x4100 - 2 * Hamming - a positive fp16 bit pattern, monotone in -Hamming.  The epilogue reads with
//      .pack::16b (two cells per register): all 240 columns of a lane quadrant are 120 registers, ONE batch, and the
//      maxima are half2 (HMNMX2).  In isolation (tools/mxf4_probe.py, nclt_tc_bench_mxp) this epilogue disappears
//      behind the five MMAs: 600 clk per tile = the 5-MMA rate itself.
//   2. The bias row of the LIBRARY side lives in the tile image (2 extra 16-byte K chunks per row, 160 B per row):
//      padding rows carry a zero bias row, their cells stay 0 - below every valid pattern - so there is no masking
//      pass and no tcgen05.st in the loop.
//   3. Tiles are 240 consecutive library rows ACROSS keyframe boundaries, with every keyframe padded to whole 48-row
//      segments (five per tile): the tensor pipe takes ~118 clk per MMA for ANY N <= 240, so only full tiles run at its
//      rate (keyframe-aligned tiles: 1000 rows = 5 steps; running across: 4.2).  A keyframe can then only end at one
//      of five places in a tile, and the host precomputes where (`endmask`): the epilogue's straight-line maxima get
//      one `if` per segment in the tiles that have a boundary, nothing in the others.  (A first version padded to 16
//      rows and walked the boundaries in the kernel - a branch per 16-column register group, BSSY / BSYNC around each
//      because ptxas cannot see that the walk is warp-uniform: a boundary tile cost 3.6 x a plain one, +35 % on the
//      whole kernel.)  Tiles restart at "group" boundaries (the first keyframe boundary after >= 16 tiles' worth of
//      rows), the only places a work split may start - so the image does not depend on the batch size.
// The kernel either EMITS the Lowe-ratio candidates (row, keyframe) straight from the epilogue (ratio mode: no
// per-(row, keyframe) plane is ever written) or writes the (d1, d2 bound) plane (flat top-2 mode, config 5).
// TMEM: accumulator buffers at columns 0 and 240; scale factors 1.0 at [480, 496), 2^14 at [496, 512).
// ==============================================================================================
constexpr int MA4 = 4;
constexpr int A4_TILE_BYTES = 128 * 128;        // 16 KB
constexpr int B4_ROWS = 240;
static_assert(B4_ROWS == TILE_ROWS, "tc_tiles.h builds the table for this tile height");
constexpr int B4_ROW_BYTES = 160;               // 8 K chunks of the descriptor + 2 of the bias row
constexpr int B4_STAGE_BYTES = B4_ROWS * B4_ROW_BYTES;   // 38 400
constexpr int BX_ROW_BYTES = 192;               // crossCheck image: + 2 chunks of the column-index row (tc_common.cuh)
constexpr int BX_STAGE_BYTES = B4_ROWS * BX_ROW_BYTES;   // 46 080
constexpr int A4_BIAS_BYTES = 128 * 32;         // constant bias slab of the query side
#ifndef NCLT_CORESIDENT
#define NCLT_CORESIDENT 0
#endif
constexpr int NSTAGE4 = NCLT_CORESIDENT ? 2 : 3;
constexpr int TC4_THREADS = 352;   // 8 epilogue warps, the TMA producer, two MMA issuers
constexpr uint32_t SF_ONE_COL = 480, SF_BIAS_COL = 496;
constexpr uint32_t SFX_DATA_COL = 496, SFX_BIAS_COL = 504;     // crossCheck kernel: 2^7 at [496, 504), 2^15 at [504, 512)

// SEG4, GROUP_MIN_ROWS, LibTile4, Tiles4, build_tiles4: tc_tiles.h (host + device; the tile table is unit-tested on the CPU)
struct WorkEntry { int item; int q; };

struct Tc4Params {
    const uint8_t* q_img;      // [n_mtiles][16 KB]
    const uint8_t* lib_img;
    const LibTile4* tiles;
    int n_mtiles, n_groups, n_splits;
    const int* split_tile;     // [n_splits + 1] first tile of each split (a keyframe's first tile)
    const int* kf_count;
    int n_kf;
    long long rows_total;      // valid query rows (B * Nq)
    int Nq;
    const int* q_n;            // [B] valid rows per frame or null
    int mode;                  // 0: emit ratio candidates, 1: write the (d1 | d2bound << 16) plane
    int num, den;
    WorkEntry* work;
    int* work_count;           // [0] entries emitted
    int work_cap;
    int* overflow;
    uint32_t* out;             // mode 1: [n_kf][rows_pad]
    long long rows_pad;
    // crossCheck kernel (k_tc4<true>): nearest tile-side row WITH its index for every (query-side row, tile-side keyframe)
    int xpass;                 // 1: A = frame rows, B = library;  2: A = library rows, B = the frames (one "keyframe" each)
    uint32_t* x_out;           // uint2-strided key arrays of launch_cross_combine: bwd (pass 1) / fwd (pass 2); .x is written
    int x_stride;              // rows per item in x_out (Nq / Nmax)
    int x_n_kf;                // teach keyframes (item = frame * x_n_kf + teach keyframe)
    const int* b_pstart;       // [B-side keyframes + 1] image row where each B-side keyframe starts
    const int* a_row_kf;       // pass 2: teach keyframe of every library row
    const int* a_kf_start;     // pass 2: first library row of every teach keyframe
    unsigned long long* clk;
    int clk_slot;
    int* item_ctr;             // [1] next item of the launch (zeroed before it): the CTAs pull items instead of owning a fixed share
};

__global__ void k_expand_queries4(const uint32_t* __restrict__ desc, long long n_rows, uint8_t* img) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, 16-bit group)
    long long total_rows = ((n_rows + 127) / 128) * 128;
    if (i >= total_rows * 16) return;
    long long row = i >> 4;
    int c = (int)(i & 15);
    uint2 v = make_uint2(0, 0);
    if (row < n_rows) {
        uint32_t w = desc[row * 8 + (c >> 1)];
        v = tc::expand16_fp4((c & 1) ? (w >> 16) : (w & 0xFFFFu));
    }
    long long tile = row >> 7;
    int r = (int)(row & 127);
    *reinterpret_cast<uint2*>(img + tile * A4_TILE_BYTES + tc::image_offset4(128, r, c * 8)) = v;
}

// tile image: n rows x 160 bytes in the K-major no-swizzle layout (10 K chunks of 16 bytes); chunks 8-9 = bias row
// xc: crossCheck image (192-byte rows: bias row of the index-carrying encoding + the column-index row); kf_count may be
// null (every keyframe has kf_stride rows: the frames of a batch as a library) and kf_start null (keyframe k starts at
// k * kf_stride)
__global__ void k_expand_library4(const uint32_t* __restrict__ desc, const LibTile4* __restrict__ tiles, int n_tiles,
                                  const int* __restrict__ kf_pstart, const int* __restrict__ kf_start,
                                  const int* __restrict__ kf_count, int kf_stride, int n_kf, uint8_t* img, int xc) {
    int t = blockIdx.x;
    if (t >= n_tiles) return;
    const LibTile4 lt = tiles[t];
    uint8_t* dst = img + (size_t)lt.img_off256 * 256;
    const int gpr = xc ? 24 : 20;                  // 8-byte groups per row
    for (int i = threadIdx.x; i < lt.n * gpr; i += blockDim.x) {
        const int r = i / gpr, g = i % gpr;        // g: 8-byte group of the 160- / 192-byte row
        // image row -> (keyframe, row inside it): last keyframe whose image start is <= the row
        const int pr = lt.prow0 + r;
        int lo = 0, hi = n_kf;                     // invariant: kf_pstart[lo] <= pr < kf_pstart[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (kf_pstart[mid] <= pr) lo = mid; else hi = mid;
        }
        const int local = pr - kf_pstart[lo];
        uint2 v = make_uint2(0, 0);
        if (local < (kf_count ? kf_count[lo] : kf_stride)) {
            if (g < 16) {
                const size_t src = (kf_start ? (size_t)kf_start[lo] : (size_t)lo * kf_stride) + local;
                uint32_t w = desc[src * 8 + (g >> 1)];
                v = tc::expand16_fp4((g & 1) ? (w >> 16) : (w & 0xFFFFu));
            } else {
                uint32_t w[2] = {0, 0};
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const int kb = (g & 3) * 8 + b;      // byte of the 32-byte constant row
                    const uint8_t by = !xc ? tc::mx_bias_byte(true, kb) : g < 20 ? tc::mx_xbias_byte(true, kb) : tc::mx_index_byte(255 - r, kb);
                    w[b >> 2] |= (uint32_t)by << (8 * (b & 3));
                }
                v = make_uint2(w[0], w[1]);
            }
        }
        *reinterpret_cast<uint2*>(dst + tc::image_offset4(lt.n, r, g * 8)) = v;
    }
}

// Running maximum of 16-bit patterns (0 or 0x3F00..0x4100) kept in the HIGH half of a 32-bit accumulator.  A packed
// register holds two patterns (hi : lo); as IEEE f32 bit patterns such words are positive numbers ordered by their high
// half first, so  acc = max.f32(acc, r, r << 16)  leaves max(acc_hi, hi, lo) in the high half - one shift and one
// three-input FMNMX per register.  (The packed forms measured ~8 clk per instruction and sub-partition here:
// HMNMX2 / VHMNMX / VIMNMX.U16x2 made the maxima of a 240-column step cost 1000-2000 clk; FMNMX3 issues every 2 clk.)
__device__ __forceinline__ uint32_t fmax3u(uint32_t acc, uint32_t x, uint32_t y) {
    // plain fmaxf (ptxas fuses the pair into one FMNMX3 and is free to schedule it; an inline-asm max is not)
    return __float_as_uint(fmaxf(fmaxf(__uint_as_float(acc), __uint_as_float(x)), __uint_as_float(y)));
}
// Two groups of 8 packed registers (32 columns) -> the 8 running maxima.  Pass 1 takes the high halves as they are,
// then every register is shifted IN PLACE (no temporaries: the kernel sits at its register cap, and ptxas funnels
// temporaries through one register, serialising shift -> max pairs at ~13 clk each) and pass 2 takes the former low
// halves.  (The packed 16-bit min / max forms - HMNMX2, VHMNMX, VIMNMX[3].U16x2 - were measured at ~12 clk per
// instruction and sub-partition; FMNMX3 issues every 2 clk.)
template <int NA, int NB>
__device__ __forceinline__ void grp2(uint32_t (&x)[NA], int b0, uint32_t (&y)[NB], int b1, uint32_t (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j & 3] = fmax3u(acc[j & 3], x[b0 + j], y[b1 + j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[b0 + j] <<= 16; y[b1 + j] <<= 16; }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j & 3] = fmax3u(acc[j & 3], x[b0 + j], y[b1 + j]);
}
template <int NA>
__device__ __forceinline__ void grp1(uint32_t (&x)[NA], int b0, uint32_t (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j & 3] = fmax3u(acc[j & 3], x[b0 + j], x[b0 + j] << 16);
}
// One 48-column segment = 24 packed registers -> the 4 running maxima, same two-pass scheme.
__device__ __forceinline__ void seg24(uint32_t* x, uint32_t (&acc)[4]) {
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j & 3] = fmax3u(acc[j & 3], x[j], x[j + 12]);
#pragma unroll
    for (int j = 0; j < 24; ++j) x[j] <<= 16;
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j & 3] = fmax3u(acc[j & 3], x[j], x[j + 12]);
}

struct Acc4 { uint32_t v[4]; };            // running maxima (high halves) of 4 disjoint column subsets

struct Tc4Epilogue {
    const Tc4Params* p;         // the kernel's __grid_constant__ parameter block (constant bank, no local copy)
    long long row;          // this lane's query row
    bool row_ok;            // row < rows_total and the step is a real one
    int lane;

    // crossCheck kernel: the tile-side keyframe `kf` is finished for this lane's row.  best = largest index-carrying cell
    // (0: the keyframe has no valid row), row0 = image row of column 0 of the tile it was found in.
    __device__ __forceinline__ void xfinalize(int kf, uint32_t best, int row0) const {
        if (!row_ok || best == 0u) return;
        const uint32_t x = best - tc::MX_X_BASE;
        const uint32_t dist = 256u - (x >> 8);
        const int idx = row0 + 255 - (int)(x & 255u) - __ldg(p->b_pstart + kf);       // row inside the tile-side keyframe
        const uint32_t key = (dist << NCLT_KEY_SHIFT) | (uint32_t)idx;
        if (p->xpass == 1) {                 // frame rows against the library: bwd[(frame, keyframe)][frame row]
            const int b = (int)(row / p->Nq), j = (int)(row - (long long)b * p->Nq);
            if (p->q_n && j >= __ldg(p->q_n + b)) return;
            p->x_out[2 * (((size_t)b * p->x_n_kf + kf) * p->x_stride + j)] = key;
        } else {                             // library rows against the frames: fwd[(frame, keyframe)][teach row]
            const int ka = __ldg(p->a_row_kf + row);
            const int i = (int)row - __ldg(p->a_kf_start + ka);
            p->x_out[2 * (((size_t)kf * p->x_n_kf + ka) * p->x_stride + i)] = key;
        }
    }
    // Keyframe finished for this lane's row: top-2 of the 8 column-subset maxima -> (exact d1, upper bound of d2).
    // Inlined at its two call sites (one per query tile of the epilogue set): out-of-line calls
    // cost a stack frame and a local copy of the parameter block whose dependent loads sat on every step's path.
    __device__ __forceinline__ void finalize(int kf, const Acc4& a) const {
        // top-2 of the 4 subset maxima (high halves): second largest of a union = max(min of the maxima, max of the seconds)
        const uint32_t v0 = a.v[0] >> 16, v1 = a.v[1] >> 16, v2 = a.v[2] >> 16, v3 = a.v[3] >> 16;
        const uint32_t Ha = max(v0, v1), La = min(v0, v1), Hb = max(v2, v3), Lb = min(v2, v3);
        const uint32_t m1 = max(Ha, Hb), m2 = max(min(Ha, Hb), max(La, Lb));
        const uint32_t d1 = m1 ? (tc::MX_ZERO16 - m1) >> 1 : 0xFFFFu;
        const uint32_t d2 = m2 ? (tc::MX_ZERO16 - m2) >> 1 : 0xFFFFu;
        if (p->mode == 1) {
            if (row_ok) p->out[(size_t)kf * p->rows_pad + row] = d1 | (d2 << 16);
            return;
        }
        // ratio mode: the pair can only pass `den * d1 < num * d2` if it passes with the upper bound of d2
        bool cand = row_ok && d1 != 0xFFFFu && (d2 == 0xFFFFu || (uint32_t)p->den * d1 < (uint32_t)p->num * d2);
        if (!__any_sync(0xFFFFFFFFu, cand)) return;
        if (__ldg(p->kf_count + kf) < 2) return;            // knnMatch(k = 2) has no second neighbour: no `good` match
        int b = 0, q = 0;
        if (cand) {
            b = (int)(row / p->Nq);
            q = (int)(row - (long long)b * p->Nq);
            if (p->q_n && q >= __ldg(p->q_n + b)) cand = false;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
        if (!m) return;
        int base = 0;
        if (lane == 0) base = atomicAdd(p->work_count, __popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (cand) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < p->work_cap) p->work[pos] = WorkEntry{b * p->n_kf + kf, q};
            else atomicAdd(p->overflow, 1);
        }
    }
};

// XC = false: the matcher (ratio candidates / (d1, d2 bound) plane).  XC = true: the crossCheck flavour - index-carrying
// f32 cells (tc_common.cuh), six MMAs per step, the epilogue keeps the nearest tile-side row AND its index per
// (query-side row, tile-side keyframe) and writes launch_cross_combine's key arrays directly.
template <bool XC>
__global__ void __launch_bounds__(TC4_THREADS, 1) k_tc4_top2(const __grid_constant__ Tc4Params p) {
    constexpr int STAGE_BYTES = XC ? BX_STAGE_BYTES : B4_STAGE_BYTES;
    constexpr int ROW_BYTES = XC ? BX_ROW_BYTES : B4_ROW_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                   // MA4 x 16 KB
    uint8_t* sAb = sA + MA4 * A4_TILE_BYTES;              // 4 KB bias slab (query side, constant)
    uint8_t* sAi = sAb + A4_BIAS_BYTES;                   // XC: 4 KB index slab (all 1.0)
    uint8_t* sB = sAi + (XC ? A4_BIAS_BYTES : 0);         // NSTAGE4 x 38 400 (46 080)
    uint8_t* tail = sB + NSTAGE4 * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;      // [3]
    uint64_t* b_empty = bars + 5;     // [3]
    uint64_t* acc_full = bars + 8;    // [2]
    uint64_t* acc_empty = bars + 10;  // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 12);
    // Item queue of the CTA.  Items are PULLED from a launch-wide counter (a CTA that starts late - on an SM another
    // stream's kernels still occupy - simply takes fewer; with a fixed share per CTA it would finish late and the kernel
    // with it).  The producer lane fetches item k and publishes (id, k + 1) in slot k % 16; the issuers and the epilogue
    // warps, which walk the same sequence k = 0, 1, ..., wait for the tag.  The roles are never more than three items
    // apart (a_empty / acc_empty hand-overs), so 16 slots cannot be overrun.  id -1 = no more items.
    volatile int* q_id = reinterpret_cast<volatile int*>(s_tmem + 4);
    volatile uint32_t* q_seq = reinterpret_cast<volatile uint32_t*>(s_tmem + 4 + 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 16) q_seq[tid] = 0;
    if (tid == 0) {
        tc::mbar_init(a_full, 1);
        tc::mbar_init(a_empty, 2);                 // one tcgen05.commit per issuer
        for (int s = 0; s < NSTAGE4; ++s) {
            tc::mbar_init(&b_full[s], 1);
            tc::mbar_init(&b_empty[s], 2);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 4);
        }
        tc::mbar_fence_init();
    }
    if (XC) {
        tc::mx_fill_xslab(sAb, 128, 0, tid, TC4_THREADS);
        tc::mx_fill_xslab(sAi, 128, 1, tid, TC4_THREADS);
    } else {
        tc::mx_fill_bias_slab(sAb, 128, false, tid, TC4_THREADS);
    }
    tc::fence_proxy_async();
    if (warp == 0) {
        tc::tmem_alloc(s_tmem, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *s_tmem;
    if (warp < 4) {
        const uint32_t lb = (uint32_t)(warp * 32) << 16;
        tc::tmem_st16_const(tmem + lb + SF_ONE_COL, 0x7F7F7F7Fu);       // every scale factor of the real steps = 2^0
        if (XC) {
            tc::tmem_st8_const(tmem + lb + SFX_DATA_COL, tc::MX_X_SFA_DATA);    // query-side scale of the data steps = 2^7
            tc::tmem_st8_const(tmem + lb + SFX_BIAS_COL, tc::MX_X_SFA_BIAS);    // query-side scale of the bias step = 2^15
        } else {
            tc::tmem_st16_const(tmem + lb + SF_BIAS_COL, tc::MX_BIAS_SFA);  // query-side scale of the bias step = 2^14
        }
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const int n_items = p.n_groups * p.n_splits;
    auto fetch_item = [&](uint32_t k) -> int {          // producer lane only
        int id = atomicAdd(p.item_ctr, 1);
        if (id >= n_items) id = -1;
        q_id[k & 15] = id;
        __threadfence_block();
        q_seq[k & 15] = k + 1;
        return id;
    };
    auto take_item = [&](uint32_t k) -> int {           // issuers, epilogue warps (warp-uniform: every lane reads the slot)
        while (q_seq[k & 15] != k + 1) {}
        return q_id[k & 15];
    };
#ifdef NCLT_TC_TRACE
#define TR(st_, k_) { const uint32_t s__ = (st_) - 2000u; if (s__ < 128u && blockIdx.x == 0 && p.clk) p.clk[64 + s__ * 8 + (k_)] = clock64(); }
#else
#define TR(st_, k_)
#endif
    const long long clk0 = clock64();
    unsigned long long ns0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));

    if (warp == 8) {
        // =========================== producer ===========================
        if (lane == 0) {
            uint32_t it_cnt = 0, s = 0, ph = 0;
            for (;; ++it_cnt) {
                const int item = fetch_item(it_cnt);
                if (item < 0) break;
                const int split = item / p.n_groups, group = item % p.n_groups;
                const int m0 = group * MA4;
                const int ma = min(MA4, p.n_mtiles - m0);
                tc::mbar_wait_relaxed(a_empty, (it_cnt & 1) ^ 1);
                tc::mbar_expect_tx(a_full, (uint32_t)ma * A4_TILE_BYTES);
                for (int m = 0; m < ma; ++m)
                    tc::bulk_g2s(sA + m * A4_TILE_BYTES, p.q_img + (size_t)(m0 + m) * A4_TILE_BYTES, A4_TILE_BYTES, a_full);
                for (int t = p.split_tile[split]; t < p.split_tile[split + 1]; ++t) {
                    const LibTile4 lt = p.tiles[t];
                    tc::mbar_wait_relaxed(&b_empty[s], ph ^ 1);
                    const uint32_t bytes = (uint32_t)lt.n * ROW_BYTES;
                    tc::mbar_expect_tx(&b_full[s], bytes);
                    tc::bulk_g2s(sB + s * STAGE_BYTES, p.lib_img + (size_t)lt.img_off256 * 256, bytes, &b_full[s]);
                    if (++s == NSTAGE4) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 9) {
        // =========================== MMA issuers (warps 9 and 10) ===========================
        // Issuer i owns accumulator buffer i: the even steps (query tiles) of every library tile go to warp 9, the odd
        // ones to warp 10.  One warp per buffer because an issuer's own path per step - mbarrier try_wait, five
        // tcgen05.mma, tcgen05.commit - measured 650-700 clk while the tensor pipe is saturating shared memory (each
        // mbarrier operation ~250 clk then), more than the 600 clk the five MMAs execute in: a single issuer left the
        // pipe idle 40 % of the time.  The whole warp runs the loops (warp-uniform control flow and operands); one
        // elected lane issues the tcgen05 instructions.  Descriptors are base + offset.
        {
            const int iss = warp - 9;
            uint32_t it_cnt = 0, s = 0, ph = 0, st = 0;
            const uint64_t da_bias = tc::smem_desc(tc::smem_u32(sAb), 2048u, 128u);
            const uint64_t da0 = tc::smem_desc(tc::smem_u32(sA), 2048u, 128u);       // + (m * 16 KB + k * 4 KB) / 16
            const uint64_t db0 = tc::smem_desc(tc::smem_u32(sB), 0u, 128u);          // + stage / 16 + (n << 16) + 2 k n
            const uint32_t d = tmem + iss * B4_ROWS;
            for (;; ++it_cnt) {
                const int item = take_item(it_cnt);
                if (item < 0) break;
                const int split = item / p.n_groups, group = item % p.n_groups;
                const int ma = min(MA4, p.n_mtiles - group * MA4);
                // an odd number of query tiles gets one empty step per library tile: steps per tile stay even,
                // so issuer / epilogue set s always owns accumulator buffer s and sees every phase of its barriers
                const int ma_pad = (ma + 1) & ~1;
                tc::mbar_wait(a_full, it_cnt & 1);
                for (int t = p.split_tile[split]; t < p.split_tile[split + 1]; ++t) {
                    const uint32_t n = p.tiles[t].n;
                    tc::mbar_wait(&b_full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t idesc = tc::idesc_mxf4(128, (int)n);
                    // K-major, no swizzle: LBO = n * 16 bytes -> descriptor field n; K chunk pair k starts 2 k n * 16 bytes in
                    const uint64_t db_tile = db0 + (uint64_t)(s * (STAGE_BYTES >> 4)) + ((uint64_t)n << 16);
                    const uint64_t db_bias = db_tile + 8u * n;
                    for (int m = iss; m < ma_pad; m += 2) {
                        const uint32_t stm = st + (uint32_t)m;                   // global step number, stm & 1 == iss
                        tc::mbar_wait(&acc_empty[iss], ((stm >> 1) & 1) ^ 1);
                        tc::tc_fence_after();
                        TR(stm, 0);
                        if (tc::elect_one()) {
                            if (m < ma) {
                                const uint64_t da_m = da0 + (uint64_t)(m * (A4_TILE_BYTES >> 4));
                                if (XC) {
                                    // bias step first (accumulate = 0): 2^23 + 2^15; then 255 - column; then 128 x the +-1 products
                                    tc::mma_mxf4(d, da_bias, db_bias, idesc, 0u, tmem + SFX_BIAS_COL, tmem + SF_ONE_COL);
                                    tc::mma_mxf4(d, da_bias + (A4_BIAS_BYTES >> 4), db_tile + 10u * n, idesc, 1u, tmem + SF_ONE_COL,
                                                 tmem + SF_ONE_COL);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc::mma_mxf4(d, da_m + (uint64_t)(k * 256), db_tile + (uint64_t)(2u * k * n), idesc, 1u,
                                                     tmem + SFX_DATA_COL, tmem + SF_ONE_COL);
                                } else {
                                    // bias step first (accumulate = 0): 1.5 * 2^23 + 0x4000 in every cell of a non-padding row
                                    tc::mma_mxf4(d, da_bias, db_bias, idesc, 0u, tmem + SF_BIAS_COL, tmem + SF_ONE_COL);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc::mma_mxf4(d, da_m + (uint64_t)(k * 256), db_tile + (uint64_t)(2u * k * n), idesc, 1u,
                                                     tmem + SF_ONE_COL, tmem + SF_ONE_COL);
                                }
                            }
                            TR(stm, 1);
                            tc::mma_commit(&acc_full[iss]);
                        }
                        __syncwarp();
                        TR(stm, 2);
                    }
                    st += (uint32_t)ma_pad;
                    if (tc::elect_one()) tc::mma_commit(&b_empty[s]);      // this issuer's MMAs on the stage are done
                    __syncwarp();
                    if (++s == NSTAGE4) { s = 0; ph ^= 1; }
                }
                if (tc::elect_one()) tc::mma_commit(a_empty);
                __syncwarp();
            }
        }
    } else {
        // =========================== epilogue (warps 0-7) ===========================
        // Warps 0-3 take the even steps (query tiles) of every library tile, warps 4-7 the odd ones; warp w reads TMEM
        // lane quadrant w % 4, ALL columns, packed, in one batch, and hands the buffer back before the maxima.
        const int quad = warp & 3, set = warp >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        uint32_t st_base = 0;
#ifdef NCLT_TC_TIMING
        long long tt_acc[5] = {0, 0, 0, 0, 0}, tt_last = clock64();
#define TT(i) { long long now_ = clock64(); tt_acc[i] += now_ - tt_last; tt_last = now_; }
#else
#define TT(i)
#endif
        for (uint32_t it_k = 0;; ++it_k) {
            const int item = take_item(it_k);
            if (item < 0) break;
            const int split = item / p.n_groups, group = item % p.n_groups;
            const int m0 = group * MA4;
            const int ma = min(MA4, p.n_mtiles - m0);
            const int ma_pad = (ma + 1) & ~1;
            Acc4 acc[MA4 / 2];
#pragma unroll
            for (int mm = 0; mm < MA4 / 2; ++mm) acc[mm].v[0] = acc[mm].v[1] = acc[mm].v[2] = acc[mm].v[3] = 0;
            Tc4Epilogue ep{&p, 0, false, lane};
            const int t_end = p.split_tile[split + 1];
            for (int t = p.split_tile[split]; t < t_end; ++t) {
                // 16-byte table entries, L1 resident; read while the MMA is still running
                const uint32_t nm = __ldg(reinterpret_cast<const uint32_t*>(&p.tiles[t].n));     // n | endmask << 16
                const int n = (int)(nm & 0xFFFFu);
                const uint32_t endmask = nm >> 16;
                const int kf0 = __ldg(&p.tiles[t].kf0);
                const int prow0 = XC ? __ldg(&p.tiles[t].prow0) : 0;
#pragma unroll
                for (int mm = 0; mm < MA4 / 2; ++mm) {
                    const int m = set + 2 * mm;
                    if (m >= ma_pad) break;
                    const uint32_t st = st_base + (uint32_t)m;
                    const int buf = st & 1;                   // == set
                    TT(0);
                    tc::mbar_wait(&acc_full[buf], (st >> 1) & 1);
                    tc::tc_fence_after();
                    TT(1);
                    if (quad == 0) TR(st, 3);
                    if (m >= ma) {                            // the empty step of an odd group: just hand it back
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
                        break;
                    }
                    const uint32_t ta = tmem + buf * B4_ROWS + lane_base;
                    if constexpr (XC) {
                        // 32-bit cells do not pack: three batches of 96 / 96 / 48 columns (2 + 2 + 1 segments), the buffer
                        // goes back after the last batch that holds columns of this tile
                        Acc4& A = acc[mm];           // v[0]: best cell so far, v[1]: image row of column 0 of its tile
                        ep.row = (long long)(m0 + m) * 128 + quad * 32 + lane;
                        ep.row_ok = ep.row < p.rows_total;
                        int kf = kf0;
                        uint32_t r[96];
#pragma unroll
                        for (int bt = 0; bt < 3; ++bt) {
                            const int c0 = 96 * bt;
                            if (c0 < n) {
                                tc::tmem_ld32(ta + c0, r);
                                if (bt < 2) {
                                    tc::tmem_ld32(ta + c0 + 32, r + 32);
                                    tc::tmem_ld32(ta + c0 + 64, r + 64);
                                } else {
                                    tc::tmem_ld16(ta + c0 + 32, r + 32);
                                }
                                tc::tmem_wait_ld();
                                if (c0 + 96 >= n) {
                                    tc::tc_fence_before();
                                    __syncwarp();
                                    if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
                                    TR(st, 4 + quad);
                                }
#pragma unroll
                                for (int h = 0; h < (bt < 2 ? 2 : 1); ++h) {
                                    const int sg = 2 * bt + h;
                                    if (SEG4 * sg < n) {
                                        uint32_t m0_ = 0, m1_ = 0;
#pragma unroll
                                        for (int j = 0; j < 24; j += 2) {
                                            m0_ = fmax3u(m0_, r[48 * h + j], r[48 * h + 24 + j]);
                                            m1_ = fmax3u(m1_, r[48 * h + j + 1], r[48 * h + 24 + j + 1]);
                                        }
                                        const uint32_t mx = max(m0_, m1_);     // positive floats of one exponent: integer order
                                        // inside a segment the column bits break ties (lowest column); between segments
                                        // and tiles only the distance counts, strictly: the earlier rows keep a tie
                                        if ((mx >> 8) > (A.v[0] >> 8)) { A.v[0] = mx; A.v[1] = (uint32_t)prow0; }
                                        if ((endmask >> sg) & 1u) {
                                            ep.xfinalize(kf, A.v[0], (int)A.v[1]);
                                            A.v[0] = 0;
                                            ++kf;
                                        }
                                    }
                                }
                            }
                        }
                        continue;
                    }
                    // all 240 columns of the lane quadrant, packed: 120 registers, five loads back to back, no branches
                    // in between (columns beyond a short tile's n hold stale cells that are never looked at)
                    uint32_t r[120];
                    tc::tmem_ld32_pack16(ta, r);
                    tc::tmem_ld32_pack16(ta + 64, r + 32);
                    tc::tmem_ld32_pack16(ta + 128, r + 64);
                    tc::tmem_ld16_pack16(ta + 192, r + 96);
                    tc::tmem_ld8_pack16(ta + 224, r + 112);
                    tc::tmem_wait_ld();
#ifdef NCLT_TC_TIMING
                    if ((r[31] ^ r[63] ^ r[95] ^ r[111] ^ r[119]) == 0x12345u) s_tmem[1] = r[0];   // registers observably ready
#endif
                    TT(2);
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);     // every column is in registers
                    TT(3);
                    TR(st, 4 + quad);
                    Acc4& A = acc[mm];
                    if (endmask == 0) {
                        // no keyframe ends inside this tile (3 tiles in 4 with 1000-row keyframes; such a tile is full)
#pragma unroll
                        for (int sg = 0; sg < 5; ++sg) seg24(r + 24 * sg, A.v);
                    } else {
                        ep.row = (long long)(m0 + m) * 128 + quad * 32 + lane;
                        ep.row_ok = ep.row < p.rows_total;
                        int kf = kf0;
#pragma unroll
                        for (int sg = 0; sg < 5; ++sg) {
                            if (SEG4 * sg < n) {
                                seg24(r + 24 * sg, A.v);
                                if ((endmask >> sg) & 1u) {       // a keyframe ends with this segment
                                    ep.finalize(kf, A);
                                    A.v[0] = A.v[1] = A.v[2] = A.v[3] = 0;
                                    ++kf;
                                }
                            }
                        }
                    }
                    TT(4);
                }
                st_base += (uint32_t)ma_pad;
            }
        }
#ifdef NCLT_TC_TIMING
        // phase cycles of epilogue warps 0 and 4 of CTA 0: [other (loop, finalize), wait-full, loads, release, maxima]
        if (blockIdx.x == 0 && lane == 0 && quad == 0 && p.clk)
            for (int i = 0; i < 5; ++i) p.clk[2 + set * 5 + i] = (unsigned long long)tt_acc[i];
#endif
    }
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0 && p.clk) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        atomicMax(p.clk, (unsigned long long)(clock64() - clk0));
        atomicMax(p.clk + 1, ns1 - ns0);
        atomicMin(p.clk + 16 + 2 * p.clk_slot, ns0);
        atomicMax(p.clk + 17 + 2 * p.clk_slot, ns1);
    }
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- candidate filter + exact verification ----------------------------------------------------
// Rows whose (exact d1, upper bound of d2) could pass the Lowe ratio are candidates; each is then
// re-evaluated exactly on the integer pipe.  Three small kernels, no shared memory beyond a few ints,
// < 10 K registers per CTA (they can co-reside with a k_tc_top2 CTA of the next batch):
//   k_tc_candidates  one warp per (frame, keyframe): ordered candidate list written into the item's
//                    slice of out_pairs as (query, -1); a global work list gets one entry per candidate
//   k_tc_verify      one warp per work-list entry: exact best (lowest train row on ties) and exact
//                    second-best distance over the keyframe, exact ratio test -> train row or -1
//   k_tc_compact     one warp per item that had candidates: drops the -1 entries, keeps query order
struct WorkItem { int item; int slot; };

__global__ void __launch_bounds__(128) k_tc_candidates(const uint32_t* __restrict__ d12, long long rows_pad, int Nq,
                                                       const int* __restrict__ q_n, int n_kf, int n_items, int num, int den,
                                                       const int* __restrict__ kf_count, int2* out_pairs, int* out_n,
                                                       WorkItem* work, int* work_count, int work_cap) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const int b = item / n_kf, kf = item % n_kf;
    const int nq = q_n ? q_n[b] : Nq;
    const int nt = kf_count[kf];
    const uint32_t* src = d12 + (size_t)kf * rows_pad + (size_t)b * Nq;
    int2* dst = out_pairs + (size_t)item * Nq;
    int cnt = 0;
    auto is_cand = [&](uint32_t v) {
        const uint32_t d1 = v & 0xFFFFu, d2b = v >> 16;
        return d1 != 0xFFFFu && (d2b == 0xFFFFu || (uint32_t)den * d1 < (uint32_t)num * d2b);
    };
    if (nt >= 2 && (Nq & 3) == 0) {
        // 128 queries per iteration, one 16-byte load per lane (HBM-bound pass over 4 B per (query, keyframe));
        // candidates are rare (the planted keyframe's matches), so the ordered emission is the slow path
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
        for (int q0 = 0; q0 < nq; q0 += 128) {
            const int q = q0 + 4 * lane;
            uint4 v = make_uint4(0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu);
            if (q < nq) v = __ldg(src4 + (q >> 2));
            const bool f0 = q < nq && is_cand(v.x), f1 = q + 1 < nq && is_cand(v.y);
            const bool f2 = q + 2 < nq && is_cand(v.z), f3 = q + 3 < nq && is_cand(v.w);
            if (!__any_sync(0xFFFFFFFFu, f0 | f1 | f2 | f3)) continue;
            const unsigned b0 = __ballot_sync(0xFFFFFFFFu, f0), b1 = __ballot_sync(0xFFFFFFFFu, f1);
            const unsigned b2 = __ballot_sync(0xFFFFFFFFu, f2), b3 = __ballot_sync(0xFFFFFFFFu, f3);
            const unsigned lt = (1u << lane) - 1u;
            const int n = __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            int base = 0;
            if (lane == 0) base = atomicAdd(work_count, n);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            int r = __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);   // query order: lane-major
            const bool f[4] = {f0, f1, f2, f3};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (f[i]) {
                    dst[cnt + r] = make_int2(q + i, -1);
                    if (base + r < work_cap) work[base + r] = WorkItem{item, cnt + r};
                    ++r;
                }
            cnt += n;
        }
    } else if (nt >= 2) {
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const bool cand = q < nq && is_cand(src[q]);
            const unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
            if (m) {
                const int n = __popc(m);
                int base = 0;
                if (lane == 0) base = atomicAdd(work_count, n);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (cand) {
                    const int r = __popc(m & ((1u << lane) - 1u));
                    dst[cnt + r] = make_int2(q, -1);
                    if (base + r < work_cap) work[base + r] = WorkItem{item, cnt + r};
                }
                cnt += n;
            }
        }
    }
    if (lane == 0) out_n[item] = cnt;      // candidates for now; k_tc_compact turns it into verified pairs
}

__global__ void __launch_bounds__(128) k_tc_verify(const WorkItem* __restrict__ work, const int* __restrict__ work_count,
                                                   int work_cap, int Nq, int n_kf, int num, int den,
                                                   const uint4* __restrict__ q_desc, const uint4* __restrict__ lib_desc,
                                                   const int* __restrict__ kf_start, const int* __restrict__ kf_count,
                                                   int2* out_pairs) {
    const int lane = threadIdx.x & 31;
    const int n_work = min(*work_count, work_cap);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_work; w += warps) {
        const WorkItem wi = work[w];
        const int b = wi.item / n_kf, kf = wi.item % n_kf;
        int2* slot = out_pairs + (size_t)wi.item * Nq + wi.slot;
        const int qq = slot->x;
        const int nt = kf_count[kf];
        const uint4* trows = lib_desc + (size_t)kf_start[kf] * 2;
        const uint4* qa = q_desc + ((size_t)b * Nq + qq) * 2;
        const uint4 a0 = __ldg(qa), a1 = __ldg(qa + 1);
        uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;      // keys: dist << 16 | train row
        for (int j = lane; j < nt; j += 32) {
            const uint4 t0 = __ldg(trows + 2 * j), t1 = __ldg(trows + 2 * j + 1);
            uint32_t d = __popc(a0.x ^ t0.x) + __popc(a0.y ^ t0.y) + __popc(a0.z ^ t0.z) + __popc(a0.w ^ t0.w) +
                         __popc(a1.x ^ t1.x) + __popc(a1.y ^ t1.y) + __popc(a1.z ^ t1.z) + __popc(a1.w ^ t1.w);
            uint32_t key = (d << 16) | (uint32_t)j;
            uint32_t mx = max(m1, key);
            m1 = min(m1, key);
            m2 = min(m2, mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o);
            uint32_t lo = min(m1, o1), hi = max(m1, o1);
            m2 = min(min(m2, o2), hi);
            m1 = lo;
        }
        const bool pass = m2 != 0xFFFFFFFFu && (uint32_t)den * (m1 >> 16) < (uint32_t)num * (m2 >> 16);
        if (lane == 0) slot->y = pass ? (int)(m1 & 0xFFFFu) : -1;
    }
}

__global__ void __launch_bounds__(128) k_tc_compact(int n_items, int Nq, int2* out_pairs, int* out_n) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const int n = out_n[item];
    if (n == 0) return;
    int2* dst = out_pairs + (size_t)item * Nq;
    int kept = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        int2 e = make_int2(0, -1);
        if (i < n) e = dst[i];
        const bool keep = e.y >= 0;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
        __syncwarp();
        if (keep) dst[kept + __popc(m & ((1u << lane) - 1u))] = e;      // kept + rank <= i: never ahead of unread data
        __syncwarp();
        kept += __popc(m);
    }
    if (lane == 0) out_n[item] = kept;
}

// ---- fused fp4 ratio path: candidates come straight out of the matching kernel as (item, query row) entries ----------
//   k_tc4_verify  one warp per entry: exact best (lowest train row on ties) and exact second-best distance over the
//                 keyframe, exact ratio test; a passing entry sets bit `q` in its item's bitmap and stores the train row.
//                 Items get a bitmap slot on first use (atomicCAS on the item -> slot map).
//   k_tc4_emit    one warp per used slot: bitmap -> (queryIdx, trainIdx) pairs in increasing queryIdx, out_n[item]
struct Tc4Pool {
    int* item_slot;        // [n_items] -1 = none, -2 = being allocated, else slot
    int* slot_item;        // [slots]
    int* slot_count;       // [1]
    uint32_t* bitmap;      // [slots][nwords]
    int* tidx;             // [slots][Nq]
    int slots, nwords;
};

// measured per 512-frame replay step (207 k entries; an epilogue warp appends ~13 candidates of one keyframe together):
// strided over all warps 0.85 ms, runs of 64 0.67, 32 0.57, 16 0.54, 8 0.54; 256-thread CTAs with runs of 128 0.72
constexpr int VERIFY_THREADS = 128, VERIFY_RUN = 16;
#if NCLT_CORESIDENT
#define VERIFY_BOUNDS __launch_bounds__(VERIFY_THREADS, 16)
#else
#define VERIFY_BOUNDS __launch_bounds__(VERIFY_THREADS)
#endif
__global__ void VERIFY_BOUNDS k_tc4_verify(const WorkEntry* __restrict__ work, const int* __restrict__ work_count,
                                                    int work_cap, int Nq, int n_kf, int num, int den,
                                                    const uint4* __restrict__ q_desc, const uint4* __restrict__ lib_desc,
                                                    const int* __restrict__ kf_start, const int* __restrict__ kf_count,
                                                    Tc4Pool pool, int* overflow) {
    const int lane = threadIdx.x & 31;
    const int n_work = min(*work_count, work_cap);
    // An epilogue warp appends its candidates of one (32 query rows, keyframe) together, so neighbouring entries mostly
    // share the keyframe: a CTA walks runs of VERIFY_RUN consecutive entries, its warps side by side, and the keyframe's
    // rows (32 KB) are served by L1 instead of being streamed from L2 once per entry (the kernel was bound by L2 bandwidth:
    // 207 k entries x 32 KB = 6.6 GB per 512-frame step)
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int base = blockIdx.x * VERIFY_RUN; base < n_work; base += gridDim.x * VERIFY_RUN)
    for (int w = base + wib; w < min(base + VERIFY_RUN, n_work); w += wpb) {
        const WorkEntry we = work[w];
        const int b = we.item / n_kf, kf = we.item % n_kf;
        const int nt = kf_count[kf];
        const uint4* trows = lib_desc + (size_t)kf_start[kf] * 2;
        const uint4* qa = q_desc + ((size_t)b * Nq + we.q) * 2;
        const uint4 a0 = __ldg(qa), a1 = __ldg(qa + 1);
        uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;      // keys: dist << 16 | train row
        for (int j = lane; j < nt; j += 32) {
            const uint4 t0 = __ldg(trows + 2 * j), t1 = __ldg(trows + 2 * j + 1);
            uint32_t d = __popc(a0.x ^ t0.x) + __popc(a0.y ^ t0.y) + __popc(a0.z ^ t0.z) + __popc(a0.w ^ t0.w) +
                         __popc(a1.x ^ t1.x) + __popc(a1.y ^ t1.y) + __popc(a1.z ^ t1.z) + __popc(a1.w ^ t1.w);
            uint32_t key = (d << 16) | (uint32_t)j;
            uint32_t mx = max(m1, key);
            m1 = min(m1, key);
            m2 = min(m2, mx);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o);
            uint32_t lo = min(m1, o1), hi = max(m1, o1);
            m2 = min(min(m2, o2), hi);
            m1 = lo;
        }
        const bool pass = m2 != 0xFFFFFFFFu && (uint32_t)den * (m1 >> 16) < (uint32_t)num * (m2 >> 16);
        if (!pass) continue;
        if (lane == 0) {
            int slot = atomicCAS(&pool.item_slot[we.item], -1, -2);
            if (slot == -1) {                               // first passing entry of this item: take a slot
                slot = atomicAdd(pool.slot_count, 1);
                if (slot < pool.slots) pool.slot_item[slot] = we.item;
                else { atomicAdd(overflow, 1); slot = -3; }
                __threadfence();
                atomicExch(&pool.item_slot[we.item], slot);
            } else {
                while (slot == -2) slot = atomicAdd(&pool.item_slot[we.item], 0);      // the owner publishes within a few instructions
            }
            if (slot >= 0) {
                pool.tidx[(size_t)slot * Nq + we.q] = (int)(m1 & 0xFFFFu);
                atomicOr(&pool.bitmap[(size_t)slot * pool.nwords + (we.q >> 5)], 1u << (we.q & 31));
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_tc4_emit(Tc4Pool pool, int Nq, int2* out_pairs, int* out_n) {
    const int lane = threadIdx.x & 31;
    const int n_slots = min(*pool.slot_count, pool.slots);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slots; s += warps) {
        const int item = pool.slot_item[s];
        int2* dst = out_pairs + (size_t)item * Nq;
        const int* tidx = pool.tidx + (size_t)s * Nq;
        int base = 0;
        for (int w0 = 0; w0 < pool.nwords; w0 += 32) {
            const int w = w0 + lane;
            uint32_t bits = w < pool.nwords ? pool.bitmap[(size_t)s * pool.nwords + w] : 0u;
            int cnt = __popc(bits), incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            int pos = base + incl - cnt;
            while (bits) {
                const int bit = __ffs(bits) - 1;
                bits &= bits - 1;
                const int q = w * 32 + bit;
                dst[pos++] = make_int2(q, tidx[q]);
            }
            base += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) out_n[item] = base;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side: library image cache + launch
// ---------------------------------------------------------------------------------------
struct TcLibCache {
    int built_for_kf = -1, built_for_desc = -1;
    uint8_t* d_img = nullptr;
    LibTile* d_tiles = nullptr;
    int n_tiles = 0;
    std::vector<int> kf_first_tile;   // [n_kf + 1]
    // work split table cached on the device (keeps the steady-state call free of host->device copies,
    // so that a whole localisation step can be captured into a CUDA graph)
    int* d_split = nullptr;
    int split_groups = -1, split_n = 0;
    // fp4 flavour: tiles run across keyframe boundaries; a work split may only start where a tile GROUP starts
    LibTile4* d_tiles4 = nullptr;
    std::vector<int> grp_tile;           // [n_grp + 1] first tile of every group
    int* d_pstart = nullptr;             // [n_kf + 1] image row where every keyframe starts
    // crossCheck flavour only: the library rows as the query side of pass 2
    uint8_t* d_aimg = nullptr;           // A-tile image of all library rows (128 B per row)
    int* d_row_kf = nullptr;             // keyframe of every library row
};

static void tc_cache_free(TcLibCache* cch) {
    if (!cch) return;
    if (cch->d_img) cudaFree(cch->d_img);
    if (cch->d_tiles) cudaFree(cch->d_tiles);
    if (cch->d_tiles4) cudaFree(cch->d_tiles4);
    if (cch->d_split) cudaFree(cch->d_split);
    if (cch->d_pstart) cudaFree(cch->d_pstart);
    if (cch->d_aimg) cudaFree(cch->d_aimg);
    if (cch->d_row_kf) cudaFree(cch->d_row_kf);
    cch->d_pstart = cch->d_row_kf = nullptr;
    cch->d_aimg = nullptr;
    cch->d_img = nullptr;
    cch->d_tiles = nullptr;
    cch->d_tiles4 = nullptr;
    cch->d_split = nullptr;
    cch->split_groups = -1;
}

void nclt_tc_release(nclt_lib* L) {
    if (!L) return;
    for (void** slot : {&L->tc_cache, &L->tc4_cache, &L->tc4x_cache}) {
        if (*slot) {
            tc_cache_free(static_cast<TcLibCache*>(*slot));
            delete static_cast<TcLibCache*>(*slot);
            *slot = nullptr;
        }
    }
}

// fp4 = false: fp8 images (256 B per descriptor, keyframe-aligned tiles of <= 256 rows); fp4 = true: e2m1 images
// (128 B per descriptor + 32 B bias row, 240-row tiles across keyframe boundaries, keyframes padded to 48 rows)
static int tc_build_library(nclt_ctx* c, nclt_lib* L, bool fp4, bool xc = false) {
    void** slot = xc ? &L->tc4x_cache : fp4 ? &L->tc4_cache : &L->tc_cache;
    TcLibCache* cch = static_cast<TcLibCache*>(*slot);
    if (!cch) {
        cch = new TcLibCache();
        *slot = cch;
    }
    if (cch->built_for_kf == L->n_kf && cch->built_for_desc == L->n_desc) return NCLT_OK;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    tc_cache_free(cch);
    c->alloc_gen++;
    if (fp4) {
        Tiles4 tb;
        build_tiles4(L->h_count.data(), L->n_kf, xc ? BX_ROW_BYTES : B4_ROW_BYTES, tb);
        cch->grp_tile = tb.grp_tile;
        cch->n_tiles = (int)tb.tiles.size();
        tb.tiles.push_back(LibTile4{0, 48, 0, 0, 0});     // sentinel
        if (cch->n_tiles > 0) {
            CU_TRY(c, cudaMalloc(&cch->d_pstart, tb.pstart.size() * sizeof(int)));
            CU_TRY(c, cudaMemcpyAsync(cch->d_pstart, tb.pstart.data(), tb.pstart.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
            CU_TRY(c, cudaMalloc(&cch->d_img, tb.off256 * 256));
            CU_TRY(c, cudaMalloc(&cch->d_tiles4, tb.tiles.size() * sizeof(LibTile4)));
            CU_TRY(c, cudaMemcpyAsync(cch->d_tiles4, tb.tiles.data(), tb.tiles.size() * sizeof(LibTile4), cudaMemcpyHostToDevice, c->stream));
            k_expand_library4<<<cch->n_tiles, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(L->d_desc), cch->d_tiles4,
                                                                   cch->n_tiles, cch->d_pstart, L->d_start, L->d_count, 0, L->n_kf,
                                                                   cch->d_img, xc ? 1 : 0);
            c->launches++;
            CU_TRY(c, cudaGetLastError());
            if (xc) {
                // pass 2 of the crossCheck: the library rows are the QUERY side - A-tile image of all rows + row -> keyframe
                const long long rows_pad = ((long long)L->n_desc + 127) / 128 * 128;
                if (rows_pad > 0) {
                    CU_TRY(c, cudaMalloc(&cch->d_aimg, (size_t)rows_pad * 128));
                    k_expand_queries4<<<(unsigned)((rows_pad * 16 + 255) / 256), 256, 0, c->stream>>>(
                        reinterpret_cast<const uint32_t*>(L->d_desc), L->n_desc, cch->d_aimg);
                }
                std::vector<int> row_kf((size_t)std::max(L->n_desc, 1), 0);
                for (int k = 0; k < L->n_kf; ++k)
                    for (int r = 0; r < L->h_count[k]; ++r) row_kf[(size_t)L->h_start[k] + r] = k;
                CU_TRY(c, cudaMalloc(&cch->d_row_kf, row_kf.size() * sizeof(int)));
                CU_TRY(c, cudaMemcpyAsync(cch->d_row_kf, row_kf.data(), row_kf.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                c->launches++;
                CU_TRY(c, cudaStreamSynchronize(c->stream));      // row_kf is a local
            }
        }
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        cch->built_for_kf = L->n_kf;
        cch->built_for_desc = L->n_desc;
        return NCLT_OK;
    }
    std::vector<LibTile> tiles;
    std::vector<int> row0;
    cch->kf_first_tile.assign(L->n_kf + 1, 0);
    size_t off256 = 0;
    for (int k = 0; k < L->n_kf; ++k) {
        cch->kf_first_tile[k] = (int)tiles.size();
        int cnt = L->h_count[k], start = L->h_start[k];
        if (cnt == 0) {   // an empty keyframe still owns one (all-padding) tile so that it gets an output row
            LibTile t{(uint32_t)off256, 16, 0, k, 1};
            tiles.push_back(t);
            row0.push_back(start);
            off256 += 16;
            continue;
        }
        const int per = 256;
        for (int r = 0; r < cnt; r += per) {
            int nv = std::min(per, cnt - r);
            int n = (nv + 15) & ~15;
            LibTile t{(uint32_t)off256, (uint16_t)n, (uint16_t)nv, k, r + per >= cnt ? 1 : 0};
            tiles.push_back(t);
            row0.push_back(start + r);
            off256 += (size_t)n;     // n * 256 bytes / 256
        }
    }
    cch->kf_first_tile[L->n_kf] = (int)tiles.size();
    cch->n_tiles = (int)tiles.size();
    tiles.push_back(LibTile{0, 16, 0, -1, 0});     // sentinel: the epilogue prefetches entry t + 1
    if (cch->n_tiles == 0) { cch->built_for_kf = L->n_kf; cch->built_for_desc = L->n_desc; return NCLT_OK; }
    int* d_row0 = nullptr;
    CU_TRY(c, cudaMalloc(&cch->d_img, off256 * 256));
    CU_TRY(c, cudaMalloc(&cch->d_tiles, tiles.size() * sizeof(LibTile)));
    CU_TRY(c, cudaMalloc(&d_row0, row0.size() * sizeof(int)));
    CU_TRY(c, cudaMemcpyAsync(cch->d_tiles, tiles.data(), tiles.size() * sizeof(LibTile), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_row0, row0.data(), row0.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_expand_library<<<cch->n_tiles, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(L->d_desc), cch->d_tiles, d_row0,
                                                          cch->n_tiles, cch->d_img);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_row0);
    cch->built_for_kf = L->n_kf;
    cch->built_for_desc = L->n_desc;
    return NCLT_OK;
}

// ---- shared first stage: (exact d1, upper bound of d2) of every query row against every keyframe -------------
struct TcPlan {
    TcLibCache* cch;
    bool fp4;
    long long rows, rows_pad;
    int n_mtiles, n_groups, n_splits;
    size_t q_img_bytes, d12_bytes;
};

// library image + work split table for a batch of B x Nq query rows; no kernel launches, no scratch
// items = n_groups x n_splits are dealt round-robin to one persistent CTA per SM: the kernel lasts
// ceil(items / SMs) item-times plus ~10 us per item (query-tile reload, pipeline drain and refill), so among
// the split counts near the target (>= ~24 items per SM) pick the cheapest (e.g. 1000 groups: 4 splits -> 27.03 items
// per SM -> 28 rounds, 3.5 % of the last one idle; 5 splits -> 33.8 -> 34 rounds, 0.6 %: measured 22.5 -> 21.9 ms)
static int tc_choose_splits(int sm_count, int n_groups, int cap) {
    const int target = std::max(1, std::min(cap, (sm_count * 24 + n_groups - 1) / n_groups));
    int best = target;
    double best_cost = 1e30;
    for (int sp = std::max(1, target / 2); sp <= std::min(cap, 2 * target); ++sp) {
        const long long items = (long long)n_groups * sp;
        const long long rounds = (items + sm_count - 1) / sm_count;
        const double cost = (double)rounds * (1.0 / sp + 0.005);
        if (cost < best_cost - 1e-12) { best_cost = cost; best = sp; }
    }
    return best;
}
// split s starts at the legal start (a keyframe's first tile / a group's first tile) nearest to tile
// n_tiles * s / n_splits: equal tile counts, not equal keyframe counts, so ragged libraries stay balanced
static std::vector<int> tc_split_table(const std::vector<int>& starts, int n_tiles, int n_splits) {
    std::vector<int> split_tile(n_splits + 1);
    const int n_starts = (int)starts.size() - 1;
    for (int s = 0; s <= n_splits; ++s) {
        const int want = (int)((long long)n_tiles * s / n_splits);
        int k = (int)(std::lower_bound(starts.begin(), starts.end(), want) - starts.begin());      // first start >= want
        if (k > 0 && (k > n_starts || want - starts[k - 1] < starts[k] - want)) --k;
        split_tile[s] = starts[std::min(k, n_starts)];
    }
    split_tile[0] = 0;
    split_tile[n_splits] = n_tiles;
    for (int s = 1; s <= n_splits; ++s) split_tile[s] = std::max(split_tile[s], split_tile[s - 1]);
    return split_tile;
}

static int tc_plan(nclt_ctx* c, nclt_lib* L, int B, int Nq, bool fp4, TcPlan* pl, bool xc = false) {
    int rc;
    if ((rc = tc_build_library(c, L, fp4, xc))) return rc;
    TcLibCache* cch = static_cast<TcLibCache*>(xc ? L->tc4x_cache : fp4 ? L->tc4_cache : L->tc_cache);
    const int ma_tiles = fp4 ? MA4 : MA;
    const size_t a_tile_bytes = fp4 ? A4_TILE_BYTES : A_TILE_BYTES;
    const int n_kf = L->n_kf;
    pl->cch = cch;
    pl->fp4 = fp4;
    pl->rows = (long long)B * Nq;
    pl->n_mtiles = (int)((pl->rows + 127) / 128);
    pl->rows_pad = (long long)pl->n_mtiles * 128;
    pl->n_groups = (pl->n_mtiles + ma_tiles - 1) / ma_tiles;
    // a split starts at a keyframe's first tile (fp8) / at a tile group (fp4)
    pl->n_splits = tc_choose_splits(c->sm_count, pl->n_groups, fp4 ? std::max((int)cch->grp_tile.size() - 1, 1) : std::max(n_kf, 1));
    pl->q_img_bytes = (size_t)pl->n_mtiles * a_tile_bytes;
    pl->d12_bytes = (size_t)n_kf * pl->rows_pad * 4;
    if (n_kf == 0 || cch->n_tiles == 0) return NCLT_OK;
    if (cch->split_groups != pl->n_groups || cch->split_n != pl->n_splits) {
        const std::vector<int> split_tile = tc_split_table(fp4 ? cch->grp_tile : cch->kf_first_tile, cch->n_tiles, pl->n_splits);
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        c->alloc_gen++;
        if (cch->d_split) cudaFree(cch->d_split);
        cch->d_split = nullptr;
        CU_TRY(c, cudaMalloc(&cch->d_split, (pl->n_splits + 1) * sizeof(int)));
        CU_TRY(c, cudaMemcpy(cch->d_split, split_tile.data(), (pl->n_splits + 1) * sizeof(int), cudaMemcpyHostToDevice));
        cch->split_groups = pl->n_groups;
        cch->split_n = pl->n_splits;
    }
    return NCLT_OK;
}

static int tc_launch_persistent(nclt_ctx* c, const void* kernel, void* params, int grid, size_t smem, int threads) {
    CU_TRY(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NCLT_CORESIDENT || getenv("NCLT_TC_CARVEOUT")) {
        // the SM's shared-memory carve-out is fixed while a CTA is resident: ask for all of it, so that what this kernel
        // does not use (82 KB with two library-tile stages) is there for the other engine's tail kernels
        const char* env = getenv("NCLT_TC_CARVEOUT");
        CU_TRY(c, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, env ? atoi(env) : (int)cudaSharedmemCarveoutMaxShared));
    }
    // Two contexts that alternate batches (PipelinedLocalizer, bench.py) can make the matching kernel of one leave
    // `tail_sms` SMs to the short tail kernels of the other (nclt_ctx_set_tail_sms).  Measured with the final kernel
    // (512-frame steps, 0 / 4 / 8 / 12 / 20 SMs): 26.2-26.8k frames/s, no trend - the tail is ~400 SM-ms of
    // throughput-bound work per step, so the SMs it gets cost the matching kernel what they save.  Default 0.
    grid = std::max(1, std::min(grid, c->sm_count - c->tail_sms));
    if (const char* env = getenv("NCLT_TC_GRID")) {      // experiment knob
        int g = atoi(env);
        if (g > 0) grid = std::min(grid, g);
    }
    // Highest launch priority: when two engines alternate (PipelinedLocalizer) this kernel's CTAs must be placed
    // before the small tail CTAs of the previous batch (each of which would otherwise pin some of the shared memory
    // a matching CTA needs and delay it).
    static int prio_hi = 1 << 30;
    if (prio_hi == (1 << 30)) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        prio_hi = greatest;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = prio_hi;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* args[1] = {params};
    nclt_prof_mark(c);
    cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, args);
    nclt_prof_mark(c);
    c->launches++;
    if (e != cudaSuccess) return nclt_fail(c, NCLT_ERR_CUDA, "tensor matching kernel launch", e);
    return NCLT_OK;
}

static int tc_clock_slot(nclt_ctx* c, unsigned long long** clk, int* slot) {
    if (!c->d_tc_clk && c->prof) {      // diagnostics only in profile mode (allocation is not capturable)
        CU_TRY(c, cudaMalloc(&c->d_tc_clk, 512 + 8192));
        CU_TRY(c, cudaMemsetAsync(c->d_tc_clk, 0, 512 + 8192, c->stream));
        c->tc_clk_launch = 0;
    }
    *clk = c->prof ? c->d_tc_clk : nullptr;
    *slot = 0;
    if (*clk) {
        *slot = c->tc_clk_launch++ % 24;
        CU_TRY(c, cudaMemsetAsync(*clk, 0, 128, c->stream));
        CU_TRY(c, cudaMemsetAsync(*clk + 16 + 2 * *slot, 0xFF, 8, c->stream));     // start: atomicMin
        CU_TRY(c, cudaMemsetAsync(*clk + 17 + 2 * *slot, 0, 8, c->stream));        // end: atomicMax
    }
    return NCLT_OK;
}

// fp4: expands the queries and runs k_tc4_top2 in ratio-candidate mode (mode 0) or plane mode (mode 1)
static int tc4_run(nclt_ctx* c, const nclt_lib* L, const TcPlan& pl, const uint8_t* q, const int32_t* q_n, int Nq, uint8_t* q_img,
                   int mode, int num, int den, WorkEntry* work, int* work_count, int work_cap, uint32_t* d12) {
    TcLibCache* cch = pl.cch;
    {
        long long threads = pl.rows_pad * 16;
        k_expand_queries4<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(q), pl.rows, q_img);
        c->launches++;
    }
    Tc4Params p{};
    p.q_img = q_img; p.lib_img = cch->d_img; p.tiles = cch->d_tiles4; p.n_mtiles = pl.n_mtiles; p.n_groups = pl.n_groups;
    p.n_splits = pl.n_splits; p.split_tile = cch->d_split;
    p.kf_count = L->d_count; p.n_kf = L->n_kf;
    p.rows_total = pl.rows; p.Nq = Nq; p.q_n = q_n; p.mode = mode; p.num = num; p.den = den;
    p.work = work; p.work_count = work_count; p.work_cap = work_cap; p.overflow = c->d_overflow;
    p.out = d12; p.rows_pad = pl.rows_pad;
    int rc;
    if ((rc = tc_clock_slot(c, &p.clk, &p.clk_slot))) return rc;
    const size_t smem = (size_t)MA4 * A4_TILE_BYTES + A4_BIAS_BYTES + (size_t)NSTAGE4 * B4_STAGE_BYTES + 512;
    p.item_ctr = c->d_overflow + 1;
    CU_TRY(c, cudaMemsetAsync(p.item_ctr, 0, 4, c->stream));
    return tc_launch_persistent(c, (const void*)k_tc4_top2<false>, &p, std::min(c->sm_count, pl.n_groups * pl.n_splits), smem, TC4_THREADS);
}

// ---- crossCheck of every frame against every keyframe on the tensor cores ----------------------------------------
// cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(desc_t, desc_curr) (visual_landmark_matcher.py:211,327; exp 63's
// whole-library ranking, experiments/63_global_reloc/scripts/visual_landmark_matcher.py:314-345): pair (i, j) iff j is the
// nearest frame row of teach row i (lowest j on ties) and i the nearest teach row of j (lowest i on ties).  Two passes of
// k_tc4_top2<true> with the roles swapped - the index-carrying cells give argmin and distance exactly, no verification
// pass - fill the same fwd / bwd key arrays the integer path fills, and launch_cross_combine (hamming.cu) does the rest.
int tc4_match_cross_all(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, int Nmax,
                        int32_t* out_pairs, uint16_t* out_dist, int32_t* out_n) {
    TcPlan pl;
    int rc;
    if ((rc = tc_plan(c, L, B, Nq, true, &pl, true))) return rc;
    const int n_kf = L->n_kf;
    if (n_kf == 0) return NCLT_OK;
    TcLibCache* cch = pl.cch;
    const size_t items = (size_t)B * n_kf;
    // pass 2 geometry: the library rows are the query side, the frames the tile side (every frame = Nq rows; rows past
    // q_n[b] become zero image rows)
    Tiles4 ft;
    build_tiles4(nullptr, B, BX_ROW_BYTES, ft, Nq);
    const int f_tiles = (int)ft.tiles.size();
    const long long a_rows = L->n_desc;
    const int a_mtiles = (int)((a_rows + 127) / 128), a_groups = (a_mtiles + MA4 - 1) / MA4;
    const int f_splits = tc_choose_splits(c->sm_count, std::max(a_groups, 1), std::max((int)ft.grp_tile.size() - 1, 1));
    const std::vector<int> f_split = tc_split_table(ft.grp_tile, f_tiles, f_splits);
    ft.tiles.push_back(LibTile4{0, 48, 0, 0, 0});     // sentinel

    ScratchScope scope(c);
    const size_t fwd_bytes = pad256(items * Nmax * sizeof(uint2)), bwd_bytes = pad256(items * Nq * sizeof(uint2));
    const size_t f_img_bytes = pad256(ft.off256 * 256);
    size_t need = fwd_bytes + bwd_bytes + pad256(pl.q_img_bytes) + f_img_bytes + pad256(ft.tiles.size() * sizeof(LibTile4)) +
                  pad256(ft.pstart.size() * sizeof(int)) + pad256(f_split.size() * sizeof(int)) + 256;
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    uint2* fwd = cv.take<uint2>(items * Nmax);
    uint2* bwd = cv.take<uint2>(items * Nq);
    uint8_t* q_img = cv.take<uint8_t>(pl.q_img_bytes);
    uint8_t* f_img = cv.take<uint8_t>(ft.off256 * 256);
    LibTile4* f_tiles_d = cv.take<LibTile4>(ft.tiles.size());
    int* f_pstart_d = cv.take<int>(ft.pstart.size());
    int* f_split_d = cv.take<int>(f_split.size());
    // synchronous copies: the sources are locals (this entry point is not meant for graph capture)
    CU_TRY(c, cudaMemcpyAsync(f_tiles_d, ft.tiles.data(), ft.tiles.size() * sizeof(LibTile4), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(f_pstart_d, ft.pstart.data(), ft.pstart.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(f_split_d, f_split.data(), f_split.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaMemsetAsync(fwd, 0xFF, items * Nmax * sizeof(uint2), c->stream));       // NCLT_KEY_INVALID
    CU_TRY(c, cudaMemsetAsync(bwd, 0xFF, items * Nq * sizeof(uint2), c->stream));
    const size_t smem = (size_t)MA4 * A4_TILE_BYTES + 2 * A4_BIAS_BYTES + (size_t)NSTAGE4 * BX_STAGE_BYTES + 512;
    // ---- pass 1: frame rows against the library -> bwd[(frame, keyframe)][frame row] = nearest teach row
    if (cch->n_tiles > 0) {
        k_expand_queries4<<<(unsigned)((pl.rows_pad * 16 + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(q), pl.rows, q_img);
        c->launches++;
        Tc4Params p{};
        p.q_img = q_img; p.lib_img = cch->d_img; p.tiles = cch->d_tiles4; p.n_mtiles = pl.n_mtiles; p.n_groups = pl.n_groups;
        p.n_splits = pl.n_splits; p.split_tile = cch->d_split;
        p.kf_count = L->d_count; p.n_kf = n_kf; p.rows_total = pl.rows; p.Nq = Nq; p.q_n = q_n;
        p.xpass = 1; p.x_out = reinterpret_cast<uint32_t*>(bwd); p.x_stride = Nq; p.x_n_kf = n_kf; p.b_pstart = cch->d_pstart;
        if ((rc = tc_clock_slot(c, &p.clk, &p.clk_slot))) return rc;
        p.item_ctr = c->d_overflow + 1;
        CU_TRY(c, cudaMemsetAsync(p.item_ctr, 0, 4, c->stream));
        if ((rc = tc_launch_persistent(c, (const void*)k_tc4_top2<true>, &p, std::min(c->sm_count, pl.n_groups * pl.n_splits), smem, TC4_THREADS)))
            return rc;
    }
    // ---- pass 2: teach rows against the frames -> fwd[(frame, keyframe)][teach row] = nearest frame row
    if (a_rows > 0 && f_tiles > 0) {
        k_expand_library4<<<f_tiles, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(q), f_tiles_d, f_tiles, f_pstart_d, nullptr, q_n,
                                                         Nq, B, f_img, 1);
        c->launches++;
        Tc4Params p{};
        p.q_img = cch->d_aimg; p.lib_img = f_img; p.tiles = f_tiles_d; p.n_mtiles = a_mtiles; p.n_groups = a_groups;
        p.n_splits = f_splits; p.split_tile = f_split_d;
        p.kf_count = nullptr; p.n_kf = B; p.rows_total = a_rows; p.Nq = Nq; p.q_n = nullptr;
        p.xpass = 2; p.x_out = reinterpret_cast<uint32_t*>(fwd); p.x_stride = Nmax; p.x_n_kf = n_kf; p.b_pstart = f_pstart_d;
        p.a_row_kf = cch->d_row_kf; p.a_kf_start = L->d_start;
        if ((rc = tc_clock_slot(c, &p.clk, &p.clk_slot))) return rc;
        p.item_ctr = c->d_overflow + 1;
        CU_TRY(c, cudaMemsetAsync(p.item_ctr, 0, 4, c->stream));
        if ((rc = tc_launch_persistent(c, (const void*)k_tc4_top2<true>, &p, std::min(c->sm_count, a_groups * f_splits), smem, TC4_THREADS)))
            return rc;
    }
    CU_TRY(c, cudaGetLastError());
    return launch_cross_combine(c, fwd, bwd, SegView{L->d_desc, L->d_start, L->d_count, 0}, nullptr, B, n_kf, Nmax, Nq,
                                reinterpret_cast<int2*>(out_pairs), out_dist, out_n, Nmax);
}

// fp8: expands the queries into q_img and runs k_tc_top2: d12[kf * rows_pad + row] = d1 | d2bound << 16
static int tc_run_top2(nclt_ctx* c, const nclt_lib* L, const TcPlan& pl, const uint8_t* q, int Nq, uint8_t* q_img, uint32_t* d12) {
    if (pl.fp4) return tc4_run(c, L, pl, q, nullptr, Nq, q_img, 1, 0, 1, nullptr, nullptr, 0, d12);
    TcLibCache* cch = pl.cch;
    {
        long long threads = pl.rows_pad * 16;
        k_expand_queries<<<(unsigned)((threads + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(q), pl.rows, q_img);
        c->launches++;
    }
    TcParams p;
    p.q_img = q_img; p.lib_img = cch->d_img; p.tiles = cch->d_tiles; p.n_mtiles = pl.n_mtiles; p.n_groups = pl.n_groups;
    p.n_splits = pl.n_splits; p.split_tile = cch->d_split; p.rows_total = pl.rows; p.out = d12; p.rows_pad = pl.rows_pad;
    int rc;
    if ((rc = tc_clock_slot(c, &p.clk, &p.clk_slot))) return rc;
    const size_t smem = (size_t)MA * A_TILE_BYTES + (size_t)NSTAGE * B_STAGE_BYTES + 1024;
    return tc_launch_persistent(c, (const void*)k_tc_top2, &p, std::min(c->sm_count, pl.n_groups * pl.n_splits), smem, TC_THREADS);
}

// fp4: candidates straight from the matching kernel -> exact verification -> ordered pairs.  No (row, keyframe) plane.
static int tc4_match_ratio_all(nclt_ctx* c, nclt_lib* L, const TcPlan& pl, const uint8_t* q, const int32_t* q_n, int B, int Nq,
                               int num, int den, int32_t* out_pairs, int32_t* out_n) {
    int rc;
    const int n_kf = L->n_kf;
    const long long n_items = (long long)B * n_kf;
    // capacities: the work list takes every pair in small problems and 8 entries per query row otherwise (planted
    // matches are <= 1 per row; heavy-tie inputs produce more); item slots: 8 per frame.  What does not fit is COUNTED
    // in the context's overflow counter (nclt_ctx_overflow) - the caller re-runs such a batch on the integer engine.
    const long long all_pairs = pl.rows * (long long)n_kf;
    const int work_cap = (int)std::min<long long>(all_pairs, std::max<long long>(pl.rows * 8, 1LL << 20));
    const int slots = (int)std::min<long long>(n_items, std::max<long long>((long long)B * 8, 4096));
    const int nwords = (Nq + 31) / 32;
    ScratchScope scope(c);
    size_t need = pad256(pl.q_img_bytes) + pad256((size_t)work_cap * sizeof(WorkEntry)) + pad256(16) + pad256((size_t)n_items * 4) +
                  pad256((size_t)slots * 4) + pad256((size_t)slots * nwords * 4) + pad256((size_t)slots * Nq * 4) + 256;
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    uint8_t* q_img = cv.take<uint8_t>(pl.q_img_bytes);
    WorkEntry* work = cv.take<WorkEntry>((size_t)work_cap);
    int* counters = cv.take<int>(4);                       // [0] work entries, [1] slots used
    Tc4Pool pool;
    pool.item_slot = cv.take<int>((size_t)n_items);
    pool.slot_item = cv.take<int>((size_t)slots);
    pool.bitmap = cv.take<uint32_t>((size_t)slots * nwords);
    pool.tidx = cv.take<int>((size_t)slots * Nq);
    pool.slot_count = counters + 1;
    pool.slots = slots;
    pool.nwords = nwords;
    CU_TRY(c, cudaMemsetAsync(counters, 0, 16, c->stream));
    CU_TRY(c, cudaMemsetAsync(pool.item_slot, 0xFF, (size_t)n_items * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(pool.bitmap, 0, (size_t)slots * nwords * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(out_n, 0, (size_t)n_items * 4, c->stream));
    if ((rc = tc4_run(c, L, pl, q, q_n, Nq, q_img, 0, num, den, work, counters, work_cap, nullptr))) return rc;
    nclt_prof_mark_tag(c, 6);
    k_tc4_verify<<<c->sm_count * (2048 / VERIFY_THREADS), VERIFY_THREADS, 0, c->stream>>>(work, counters, work_cap, Nq, n_kf, num, den,
                                                         reinterpret_cast<const uint4*>(q), L->d_desc, L->d_start, L->d_count, pool,
                                                         c->d_overflow);
    nclt_prof_mark_tag(c, 6);
    k_tc4_emit<<<c->sm_count * 4, 128, 0, c->stream>>>(pool, Nq, reinterpret_cast<int2*>(out_pairs), out_n);
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    if (getenv("NCLT_DEBUG_WORK")) {       // diagnostics: candidates the epilogue emitted vs item slots used
        int h[2] = {0, 0};
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(h, counters, 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[nclt] ratio candidates: %d work entries (cap %d) for %lld query rows x %d keyframes, %d item slots\n", h[0],
                work_cap, pl.rows, n_kf, h[1]);
    }
    return NCLT_OK;
}

// all keyframes, every frame: knn2 + Lowe ratio via tensor cores. Same outputs as the integer path.
int tc_match_ratio_all(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, int num, int den,
                       int32_t* out_pairs, int32_t* out_n, bool fp4) {
    int rc;
    TcPlan pl;
    if ((rc = tc_plan(c, L, B, Nq, fp4, &pl))) return rc;
    const int n_kf = L->n_kf;
    if (n_kf == 0) return NCLT_OK;
    if (pl.cch->n_tiles == 0) {      // a library of empty keyframes: no matches anywhere
        CU_TRY(c, cudaMemsetAsync(out_n, 0, (size_t)B * n_kf * 4, c->stream));
        return NCLT_OK;
    }
    if (fp4) return tc4_match_ratio_all(c, L, pl, q, q_n, B, Nq, num, den, out_pairs, out_n);
    ScratchScope scope(c);
    // candidate work list: every (query, keyframe) pair may be a candidate in the worst case
    const long long all_pairs = pl.rows * (long long)n_kf;
    const int work_cap = (int)std::min<long long>(all_pairs, 1LL << 28);
    size_t need = pad256(pl.q_img_bytes) + pad256(pl.d12_bytes) + pad256((size_t)work_cap * sizeof(WorkItem)) + 256;
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    uint8_t* q_img = cv.take<uint8_t>(pl.q_img_bytes);
    uint32_t* d12 = cv.take<uint32_t>(pl.d12_bytes / 4);
    WorkItem* work = cv.take<WorkItem>((size_t)work_cap);
    int* work_count = cv.take<int>(1);
    if ((rc = tc_run_top2(c, L, pl, q, Nq, q_img, d12))) return rc;
    {
        const int n_items = B * n_kf;
        const unsigned blocks = (unsigned)(((long long)n_items * 32 + 127) / 128);
        CU_TRY(c, cudaMemsetAsync(work_count, 0, 4, c->stream));
        k_tc_candidates<<<blocks, 128, 0, c->stream>>>(d12, pl.rows_pad, Nq, q_n, n_kf, n_items, num, den, L->d_count,
                                                      reinterpret_cast<int2*>(out_pairs), out_n, work, work_count, work_cap);
        nclt_prof_mark_tag(c, 6);
        k_tc_verify<<<c->sm_count * 16, 128, 0, c->stream>>>(work, work_count, work_cap, Nq, n_kf, num, den,
                                                            reinterpret_cast<const uint4*>(q), L->d_desc, L->d_start,
                                                            L->d_count, reinterpret_cast<int2*>(out_pairs));
        nclt_prof_mark_tag(c, 6);
        k_tc_compact<<<blocks, 128, 0, c->stream>>>(n_items, Nq, reinterpret_cast<int2*>(out_pairs), out_n);
        c->launches += 3;
    }
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

// ---- flat global top-2 (BASELINE config 5) on the tensor cores -------------------------------------------------
// The global best row of a query lies in the keyframe kf_a with the smallest (d1, keyframe index) - global row indices
// grow with the keyframe index, so this is also the lowest global index among ties; the global second best is either
// the second best row of kf_a or the best row of kf_b, the smallest (d1, keyframe) among the OTHER keyframes.  Both
// keyframes are then re-scanned exactly on the integer pipe (2 x <= 1000 rows per query instead of millions).
__global__ void __launch_bounds__(256) k_flat_select(const uint32_t* __restrict__ d12, long long rows_pad, long long rows,
                                                     int n_kf, int2* __restrict__ sel /*[rows] (kf_a, kf_b), -1 = none*/) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;            // keys d1 << 16 | kf (n_kf < 65536)
    for (int kf = 0; kf < n_kf; ++kf) {
        const uint32_t d1 = d12[(size_t)kf * rows_pad + row] & 0xFFFFu;
        if (d1 == 0xFFFFu) continue;                         // keyframe without rows
        const uint32_t key = (d1 << 16) | (uint32_t)kf;
        const uint32_t mx = max(k1, key);
        k1 = min(k1, key);
        k2 = min(k2, mx);
    }
    sel[row] = make_int2(k1 == 0xFFFFFFFFu ? -1 : (int)(k1 & 0xFFFFu), k2 == 0xFFFFFFFFu ? -1 : (int)(k2 & 0xFFFFu));
}

__global__ void __launch_bounds__(128) k_flat_rescan(const int2* __restrict__ sel, long long rows, int Nq,
                                                     const int* __restrict__ q_n, const uint4* __restrict__ q_desc,
                                                     const uint4* __restrict__ lib_desc, const int* __restrict__ kf_start,
                                                     const int* __restrict__ kf_count, uint32_t idx_offset,
                                                     uint2* __restrict__ out_keys) {
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int b = (int)(row / Nq), qi = (int)(row % Nq);
    uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;            // keys dist << 23 | global row
    if (!q_n || qi < q_n[b]) {
        const int2 s = sel[row];
        const uint4 a0 = __ldg(q_desc + row * 2), a1 = __ldg(q_desc + row * 2 + 1);
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const int kf = which == 0 ? s.x : s.y;
            if (kf < 0) continue;
            const int start = kf_start[kf], nt = kf_count[kf];
            const uint4* trows = lib_desc + (size_t)start * 2;
            for (int j = lane; j < nt; j += 32) {
                const uint4 t0 = __ldg(trows + 2 * j), t1 = __ldg(trows + 2 * j + 1);
                const uint32_t d = __popc(a0.x ^ t0.x) + __popc(a0.y ^ t0.y) + __popc(a0.z ^ t0.z) + __popc(a0.w ^ t0.w) +
                                   __popc(a1.x ^ t1.x) + __popc(a1.y ^ t1.y) + __popc(a1.z ^ t1.z) + __popc(a1.w ^ t1.w);
                const uint32_t key = (d << NCLT_KEY_SHIFT) | (idx_offset + (uint32_t)(start + j));
                const uint32_t mx = max(m1, key);
                m1 = min(m1, key);
                m2 = min(m2, mx);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o);
            const uint32_t lo = min(m1, o1), hi = max(m1, o1);
            m2 = min(min(m2, o2), hi);
            m1 = lo;
        }
    }
    if (lane == 0) out_keys[row] = make_uint2(m1, m2);
}

int tc_match_flat2(nclt_ctx* c, nclt_lib* L, const uint8_t* q, const int32_t* q_n, int B, int Nq, uint32_t idx_offset,
                   uint32_t* out_keys, bool fp4) {
    int rc;
    TcPlan pl;
    if ((rc = tc_plan(c, L, B, Nq, fp4, &pl))) return rc;
    const int n_kf = L->n_kf;
    ScratchScope scope(c);
    size_t need = pad256(pl.q_img_bytes) + pad256(pl.d12_bytes) + pad256((size_t)pl.rows * sizeof(int2)) + 256;
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    uint8_t* q_img = cv.take<uint8_t>(pl.q_img_bytes);
    uint32_t* d12 = cv.take<uint32_t>(pl.d12_bytes / 4);
    int2* sel = cv.take<int2>((size_t)pl.rows);
    if (n_kf > 0 && pl.cch->n_tiles > 0) {
        if ((rc = tc_run_top2(c, L, pl, q, Nq, q_img, d12))) return rc;
    }
    k_flat_select<<<(unsigned)((pl.rows + 255) / 256), 256, 0, c->stream>>>(d12, pl.rows_pad, pl.rows, n_kf, sel);
    k_flat_rescan<<<(unsigned)((pl.rows * 32 + 127) / 128), 128, 0, c->stream>>>(
        sel, pl.rows, Nq, q_n, reinterpret_cast<const uint4*>(q), L->d_desc, L->d_start, L->d_count, idx_offset,
        reinterpret_cast<uint2*>(out_keys));
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

// SM clock the last profiled k_tc*_top2 launch actually ran at (cycles / wall time of its longest CTA)
#ifdef NCLT_TC_TRACE
// diagnostics build only: clock64 stamps of steps 2000..2127 of CTA 0 (tools/tc_clock.py prints them)
extern "C" int nclt_ctx_tc_trace(nclt_ctx* c, unsigned long long* out1024) {
    if (!c || !c->d_tc_clk) return NCLT_ERR_ARG;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaMemcpy(out1024, c->d_tc_clk + 64, 8192, cudaMemcpyDeviceToHost));
    return NCLT_OK;
}
#endif
extern "C" int nclt_ctx_tc_clock(nclt_ctx* c, double* mhz, double* kernel_ms, unsigned long long* raw64) {
    if (!c) return NCLT_ERR_ARG;
    unsigned long long h[64] = {0};
    if (c->d_tc_clk) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, cudaMemcpy(h, c->d_tc_clk, 512, cudaMemcpyDeviceToHost));
    }
    if (raw64) memcpy(raw64, h, 512);
    if (mhz) *mhz = h[1] ? (double)h[0] / (double)h[1] * 1e3 : 0.0;
    if (kernel_ms) *kernel_ms = (double)h[1] * 1e-6;
    return NCLT_OK;
}
