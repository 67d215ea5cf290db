// ORB feature extraction: `cv2.ORB_create(nfeatures=500).detectAndCompute(gray, None)` of
// scripts/common/visual_landmark_matcher.py:207,305-306 and visual_landmark_recorder.py:159,240-241, for batches of
// frames.  OpenCV is a third-party dependency of the reference (opencv-python, not vendored); the stages below follow
// its published ORB (features2d) and are pinned against cv2 4.13.0 bit for bit (keypoint order included):
//   gray      BGR -> gray, 15-bit fixed point (b*3735 + g*19235 + r*9798 + 2^14) >> 15
//   pyramid   8 levels, scale 1.2^l (float), each level resized from the previous one with the bit-exact bilinear
//             resize (INTER_LINEAR_EXACT: 8.8 fixed-point weights, horizontal then vertical, (v + 2^15) >> 16)
//   FAST      9-of-16 segment test, threshold 20, score = largest threshold keeping the pixel a corner, 3x3 strict NMS
//   Harris    7x7 block of 3x3 Sobel products, for the survivors of the first cut (>= 31 px from the border)
//   select    per level: retainBest(2n) by FAST score, then retainBest(n) by Harris.  The permutation OpenCV's
//             std::nth_element / std::partition calls leave behind defines the output order; orb_select.cuh
//             reproduces it on the device (one warp per level), the host std:: calls remain as fall-back
//   angle     intensity centroid over the radius-15 disc, cv::fastAtan2 polynomial
//   blur      7x7 sigma-2 Gaussian as OpenCV's separable float filter runs it on a pyramid sub-matrix: row pass
//             sequential FMAs, column pass symmetric pairs with FMAs, round half to even
//   rBRIEF    256 rotated pair tests (orb_pattern.h) on the blurred level
// Byte work bound by HBM / L2: every stage reads and writes u8 planes once; nothing is GEMM shaped.
#include "common.cuh"
#include "scratch.cuh"
#include "orb_pattern.h"
#include "orb_select.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int kLevels = 8;
constexpr int kEdge = 31;        // edgeThreshold
constexpr int kHalfPatch = 15;   // patchSize / 2
constexpr int kFastThr = 20;

struct OrbGeom {
    int w[kLevels], h[kLevels], pitch[kLevels];
    long long off[kLevels];     // byte offset of the level inside one frame's pyramid
    long long frame_bytes;
    float scale[kLevels];
    int umax[kHalfPatch + 2];
    float atan_p1, atan_p3, atan_p5, atan_p7;
    float gk[4];                // Gaussian taps: gk[0] centre .. gk[3] outermost
    float deg2rad;
};

__constant__ __align__(16) signed char c_pattern[256 * 4];

// One launch covers all levels without empty CTAs: blockIdx.x runs over the tiles of level 0, then level 1, ...
// The decode is a table in device memory, one int4 {level, bx, by, 0} per block, and the level's geometry another
// {w, h, pitch, byte offset}: two broadcast loads per thread.  (Computed from `first` / `bx` - a compare chain, a dynamic
// index into kernel parameters, which ptxas expands into predicated loads, and an integer division - it was 200 of the
// 975 instructions a FAST warp executes, ncu source view of round 2.)
struct BlockMap { int first[kLevels + 1]; int bx[kLevels]; const int4* tab; const int4* lev; };
__device__ __forceinline__ void block_of(const BlockMap& m, int b, int& l, int& bx, int& by) {
    const int4 e = __ldg(m.tab + b);
    l = e.x; bx = e.y; by = e.z;
}
__device__ __forceinline__ void level_of(const BlockMap& m, int l, int& w, int& h, int& pitch, long long& off) {
    const int4 e = __ldg(m.lev + l);
    w = e.x; h = e.y; pitch = e.z; off = e.w;
}

struct Cand { uint32_t tag, key; float score, harris; };          // tag = frame * 8 + level, key = y << 16 | x
struct Sel { int frame, level, x, y; float harris; int slot; };   // slot = row inside the frame's output

constexpr int kResizeRows = 8;      // rows per CTA of the copy / resize kernels (tiny CTAs are bound by the CTA launch rate)

// ---- level 0 ----
__global__ void __launch_bounds__(256) k_orb_level0(const uint8_t* __restrict__ src, int channels, int W, int H, int pitch,
                                                    long long frame_bytes, uint8_t* __restrict__ pyr) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.z;
    if (x >= W) return;
    for (int y = blockIdx.y * kResizeRows; y < min((blockIdx.y + 1) * kResizeRows, H); ++y) {
        const uint8_t* s = src + ((size_t)f * H + y) * (size_t)W * channels + (size_t)x * channels;
        int g;
        if (channels == 1) g = s[0];
        else g = (s[0] * 3735 + s[1] * 19235 + s[2] * 9798 + 16384) >> 15;
        pyr[(size_t)f * frame_bytes + (size_t)y * pitch + x] = (uint8_t)g;
    }
}

// ---- bit-exact bilinear resize of one level from the previous one ----
// tab: [sx(dw) | ax(dw) | sy(dh) | ay(dh)] int32, from the host in double arithmetic
__global__ void __launch_bounds__(256) k_orb_resize(const uint8_t* __restrict__ src, int sw, int sh, int spitch,
                                                    uint8_t* __restrict__ dst, int dw, int dh, int dpitch,
                                                    long long frame_bytes, const int* __restrict__ tab) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.z;
    if (x >= dw) return;
    const int sx = tab[x], ax = tab[dw + x];
    const int sx1 = min(sx + 1, sw - 1);
    const uint8_t* s = src + (size_t)f * frame_bytes;
    for (int y = blockIdx.y * kResizeRows; y < min((blockIdx.y + 1) * kResizeRows, dh); ++y) {     // short rows: several per CTA
        const int sy = tab[2 * dw + y], ay = tab[2 * dw + dh + y];
        const int sy1 = min(sy + 1, sh - 1);
        const uint8_t* r0 = s + (size_t)sy * spitch;
        const uint8_t* r1 = s + (size_t)sy1 * spitch;
        const unsigned h0 = r0[sx] * (256 - ax) + r0[sx1] * ax;
        const unsigned h1 = r1[sx] * (256 - ax) + r1[sx1] * ax;
        const unsigned v = h0 * (256 - ay) + h1 * ay;
        dst[(size_t)f * frame_bytes + (size_t)y * dpitch + x] = (uint8_t)((v + 32768u) >> 16);
    }
}

// ---- FAST 9-16 score map ----
__device__ __forceinline__ bool has9(unsigned m) {     // 9 contiguous set bits in a circular 16-bit mask
    unsigned r = m | (m << 16);
    unsigned a = r & (r >> 1);
    a &= a >> 2;
    a &= a >> 4;          // 8 contiguous
    a &= r >> 8;          // 9
    return (a & 0xFFFFu) != 0;
}

// 32 x 32 pixel tile per CTA (256 threads), staged in shared memory with its 3-pixel ring.
// Phase 1, four pixels per thread: the compass test - an arc of 9 contains one pixel of every opposite pair, so a
// pair inside the threshold band on both sides rules the corner out; survivors (a few per cent, but scattered over
// 40 % of the warps) go to a list in shared memory.  Phase 2: the threads walk that list densely - full segment test
// on 16-bit arc masks, score by doubling minima - so the expensive path runs in full warps.
constexpr int kFastTile = 32;
constexpr int kFastPitch = 48;      // 11 words per tile row: pixels x0 - 6 .. x0 + 37 (x0 - 6 is a multiple of 8)
__global__ void __launch_bounds__(256, 6) k_orb_fast(const uint8_t* __restrict__ pyr, OrbGeom g, BlockMap bm, int block0,
                                                  uint8_t* __restrict__ score) {
    __shared__ __align__(16) uint8_t tile[kFastTile + 6][kFastPitch];     // column c holds pixel x0 - 6 + c
    __shared__ unsigned short list[kFastTile * kFastTile];
    __shared__ int n_list;
    const int f = blockIdx.y;
    int l, bx, by;
    block_of(bm, blockIdx.x + block0, l, bx, by);
    int w, h, p;
    long long loff;
    level_of(bm, l, w, h, p, loff);
    const int x0 = bx * kFastTile + kEdge - 1, y0 = by * kFastTile + kEdge - 1;      // first pixel of the tile
    const int xe = w - kEdge, ye = h - kEdge;                                        // last pixel scored (inclusive)
    const size_t base = (size_t)f * g.frame_bytes + loff;
    if (threadIdx.x == 0) n_list = 0;
    const int tx = threadIdx.x & 31;
    // staging with 32-bit loads (rows are 16-byte aligned, x0 - 6 = 32 bx + 24); the clamps only touch pixels that are
    // never scored
    {   // 16 threads per tile row (11 words), 16 rows per pass: no index division
        const int k = threadIdx.x & 15;
        const int xw = min(x0 - 6 + 4 * k, p - 4);
#pragma unroll
        for (int r = threadIdx.x >> 4; r < kFastTile + 6; r += 16) {
            if (k < 11) {
                const uint8_t* row = pyr + base + (size_t)min(y0 + r - 3, h - 1) * p;
                *reinterpret_cast<uint32_t*>(&tile[r][4 * k]) = *reinterpret_cast<const uint32_t*>(row + xw);
            }
        }
    }
    __syncthreads();
    {
        const int x = x0 + tx;
        uint8_t* srow = score + base + (size_t)(y0 + (int)(threadIdx.x >> 5)) * p + x;     // this thread's pixel of row ty
#pragma unroll
        for (int r = 0; r < 4; ++r, srow += 8 * p) {
            const int ty = (threadIdx.x >> 5) + 8 * r;
            if (x > xe || y0 + ty > ye) continue;
            const uint8_t* c = &tile[ty + 3][tx + 6];
            const int v = c[0];
            const int e0 = v - c[3 * kFastPitch], e8 = v - c[-3 * kFastPitch], e4 = v - c[3], e12 = v - c[-3];
            const bool in0 = (e0 <= kFastThr && e0 >= -kFastThr) && (e8 <= kFastThr && e8 >= -kFastThr);
            const bool in4 = (e4 <= kFastThr && e4 >= -kFastThr) && (e12 <= kFastThr && e12 >= -kFastThr);
            if (in0 || in4) *srow = 0;
            else list[atomicAdd(&n_list, 1)] = (unsigned short)(ty * kFastTile + tx);
        }
    }
    __syncthreads();
    const int n = n_list;
    constexpr int P = kFastPitch;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int ty = list[i] / kFastTile, tx2 = list[i] % kFastTile;
        const uint8_t* c = &tile[ty + 3][tx2 + 6];
        const int v = c[0];
        int d[16];
        d[0] = v - c[3 * P];          d[1] = v - c[3 * P + 1];   d[2] = v - c[2 * P + 2];    d[3] = v - c[P + 3];
        d[4] = v - c[3];              d[5] = v - c[-P + 3];      d[6] = v - c[-2 * P + 2];   d[7] = v - c[-3 * P + 1];
        d[8] = v - c[-3 * P];         d[9] = v - c[-3 * P - 1];  d[10] = v - c[-2 * P - 2];  d[11] = v - c[-P - 3];
        d[12] = v - c[-3];            d[13] = v - c[P - 3];      d[14] = v - c[2 * P - 2];   d[15] = v - c[3 * P - 1];
        unsigned dark = 0, bright = 0;            // ring darker / brighter than the centre by more than the threshold
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            dark |= (d[k] > kFastThr ? 1u : 0u) << k;
            bright |= (d[k] < -kFastThr ? 1u : 0u) << k;
        }
        int s = 0;
        if (has9(dark) || has9(bright)) {
            // score = (max over the 16 arcs of 9 of the smallest difference of one sign) - 1.  Differences are biased
            // to non-negative values (a = 255 + d for "darker", b = 255 - d for "brighter") and the 9-wide circular
            // minima are built by doubling: windows of 2, 4, 8, then one more element.
            int a[16], b[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                a[k] = 255 + d[k];
                b[k] = 255 - d[k];
            }
            int a2[16], b2[16], a4[16], b4[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                a2[k] = min(a[k], a[(k + 1) & 15]);
                b2[k] = min(b[k], b[(k + 1) & 15]);
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                a4[k] = min(a2[k], a2[(k + 2) & 15]);
                b4[k] = min(b2[k], b2[(k + 2) & 15]);
            }
            int best = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int a9 = min(min(a4[k], a4[(k + 4) & 15]), a[(k + 8) & 15]);
                const int b9 = min(min(b4[k], b4[(k + 4) & 15]), b[(k + 8) & 15]);
                best = max(best, max(a9, b9));
            }
            s = best - 255 - 1;
        }
        score[base + (size_t)(y0 + ty) * p + x0 + tx2] = (uint8_t)s;
    }
}

// ---- NMS + border filter + Harris response, candidates appended to one list ----
__global__ void __launch_bounds__(256) k_orb_nms(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ score, OrbGeom g,
                                                 Cand* __restrict__ cand, unsigned* __restrict__ n_cand, unsigned cap,
                                                 float harris_scale4) {
    const int f = blockIdx.z / kLevels, l = blockIdx.z % kLevels;
    const int w = g.w[l], h = g.h[l], p = g.pitch[l];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31) + kEdge, y = blockIdx.y * 8 + (threadIdx.x >> 5) + kEdge;
    if (x >= w - kEdge || y >= h - kEdge) return;
    const size_t base = (size_t)f * g.frame_bytes + g.off[l];
    const uint8_t* sc = score + base + (size_t)y * p + x;
    const int s = sc[0];
    if (s == 0) return;
    if (!(s > sc[-1] && s > sc[1] && s > sc[-p - 1] && s > sc[-p] && s > sc[-p + 1] && s > sc[p - 1] && s > sc[p] &&
          s > sc[p + 1]))
        return;
    const uint8_t* c = pyr + base + (size_t)y * p + x;
    int a = 0, b = 0, cc = 0;
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx) {
            const uint8_t* q = c + dy * p + dx;
            const int Ix = (q[1] - q[-1]) * 2 + (q[-p + 1] - q[-p - 1]) + (q[p + 1] - q[p - 1]);
            const int Iy = (q[p] - q[-p]) * 2 + (q[p - 1] - q[-p - 1]) + (q[p + 1] - q[-p + 1]);
            a += Ix * Ix;
            b += Iy * Iy;
            cc += Ix * Iy;
        }
    const float fa = (float)a, fb = (float)b, fc = (float)cc;
    const float t = __fadd_rn(fa, fb);
    const float r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, t), t)),
                              harris_scale4);
    const unsigned slot = atomicAdd(n_cand, 1u);
    if (slot < cap) cand[slot] = Cand{(uint32_t)(f * kLevels + l), (uint32_t)(y << 16 | x), (float)s, r};
}

// ---- Gaussian blur, 32 x 64 output tile per CTA (threads = 32 columns x 8 row lanes, no index divisions) ----
constexpr int kBlurH = 64;
__global__ void __launch_bounds__(256) k_orb_blur(const uint8_t* __restrict__ pyr, OrbGeom g, BlockMap bm, uint8_t* __restrict__ blur) {
    const int f = blockIdx.y;
    int l, bx, by;
    block_of(bm, blockIdx.x, l, bx, by);
    int w, h, p;
    long long loff;
    level_of(bm, l, w, h, p, loff);
    const int x0 = bx * 32, y0 = by * kBlurH;
    const uint8_t* src = pyr + (size_t)f * g.frame_bytes + loff;
    uint8_t* dst = blur + (size_t)f * g.frame_bytes + loff;
    __shared__ float tile[kBlurH + 6][45];          // column c holds pixel x0 - 4 + c (44 staged, 11 words per row)
    __shared__ float rowp[kBlurH + 6][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // staging with 32-bit loads; clamped reads: only pixels >= 9 px inside the level are ever sampled, the clamps just
    // keep the reads inside the row
    {   // 16 threads per tile row (11 words), 16 rows per pass: no index division
        const int k = threadIdx.x & 15;
        const int xw = min(max(x0 - 4 + 4 * k, 0), p - 4);
        if (k < 11) {
#pragma unroll
            for (int r = threadIdx.x >> 4; r < kBlurH + 6; r += 16) {
                const uint8_t* row = src + (size_t)min(max(y0 + r - 3, 0), h - 1) * p;
                const uint32_t q = *reinterpret_cast<const uint32_t*>(row + xw);
                float* t = &tile[r][4 * k];
                t[0] = (float)(q & 0xFFu);
                t[1] = (float)((q >> 8) & 0xFFu);
                t[2] = (float)((q >> 16) & 0xFFu);
                t[3] = (float)(q >> 24);
            }
        }
    }
    __syncthreads();
    // row pass: four adjacent outputs per task from ten loaded values (each output keeps its own sequential FMA chain)
    for (int t = threadIdx.x; t < (kBlurH + 6) * 8; t += 256) {
        const int r = t >> 3, gx = (t & 7) * 4;
        float v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) v[j] = tile[r][gx + 1 + j];        // output x0 + gx + j taps x0 + gx + j - 3 ...
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = __fmul_rn(v[j], g.gk[3]);
            s = fmaf(v[j + 1], g.gk[2], s);
            s = fmaf(v[j + 2], g.gk[1], s);
            s = fmaf(v[j + 3], g.gk[0], s);
            s = fmaf(v[j + 4], g.gk[1], s);
            s = fmaf(v[j + 5], g.gk[2], s);
            s = fmaf(v[j + 6], g.gk[3], s);
            rowp[r][gx + j] = s;
        }
    }
    __syncthreads();
    // column pass: eight consecutive rows per thread from a 14-row window held in registers
    const int x = x0 + tx;
    if (x >= w) return;
    float c[14];
#pragma unroll
    for (int j = 0; j < 14; ++j) c[j] = rowp[ty * 8 + j][tx];
    uint8_t* out = dst + (size_t)(y0 + ty * 8) * p + x;
    const int nj = min(8, h - (y0 + ty * 8));
#pragma unroll
    for (int j = 0; j < 8; ++j, out += p) {
        if (j >= nj) break;
        float s = __fmul_rn(g.gk[0], c[j + 3]);
        s = fmaf(g.gk[1], __fadd_rn(c[j + 4], c[j + 2]), s);
        s = fmaf(g.gk[2], __fadd_rn(c[j + 5], c[j + 1]), s);
        s = fmaf(g.gk[3], __fadd_rn(c[j + 6], c[j]), s);
        const int v = __float2int_rn(s);
        *out = (uint8_t)min(max(v, 0), 255);
    }
}

// ---- orientation + rBRIEF, one warp per keypoint ----
__device__ __forceinline__ float fast_atan2_deg(float y, float x, const OrbGeom& g) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = 2.220446049250313e-16f;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(g.atan_p7, c2), g.atan_p5), c2), g.atan_p3), c2), g.atan_p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(g.atan_p7, c2), g.atan_p5), c2), g.atan_p3), c2), g.atan_p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256) k_orb_describe(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, OrbGeom g,
                                                      const Sel* __restrict__ sel, int n_sel, int cap,
                                                      const int* __restrict__ n_out,
                                                      float* __restrict__ out_kp, uint8_t* __restrict__ out_desc) {
    // n_out == nullptr: sel is a dense list of n_sel entries; else sel is [frame][cap] with n_out[frame] valid rows
    // the 256 test pairs, one 4-byte word each, transposed so that lane L's i-th test sits in bank L
    __shared__ uint32_t pat[8][32];
    {
        const int t = threadIdx.x, L = t >> 3, i = t & 7;        // test t = 8 * L + i  (byte L, bit i)
        pat[i][L] = reinterpret_cast<const uint32_t*>(c_pattern)[t];
    }
    __syncthreads();
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n_sel) return;
    if (n_out && (k % cap) >= n_out[k / cap]) return;
    const Sel s = sel[k];
    const int p = g.pitch[s.level];
    const size_t base = (size_t)s.frame * g.frame_bytes + g.off[s.level];
    const uint8_t* c = pyr + base + (size_t)s.y * p + s.x;
    // intensity centroid over the disc: lane = column u + 15, the warp walks the 31 rows (one coalesced read each)
    int m01 = 0, m10 = 0;
    {
        const int u = lane - kHalfPatch, au = u < 0 ? -u : u;
        const uint8_t* cr = c - kHalfPatch * p + u;
#pragma unroll          // static index into g.umax (a dynamic one is a chain of predicated parameter loads per row)
        for (int v = -kHalfPatch; v <= kHalfPatch; ++v, cr += p) {
            if (lane <= 2 * kHalfPatch && au <= g.umax[v < 0 ? -v : v]) {
                const int val = *cr;
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xFFFFFFFFu, m01, o);
        m10 += __shfl_xor_sync(0xFFFFFFFFu, m10, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10, g);
    const float sc = g.scale[s.level];
    const float px = s.level ? __fmul_rn((float)s.x, sc) : (float)s.x;
    const float py = s.level ? __fmul_rn((float)s.y, sc) : (float)s.y;
    // descriptor centre exactly as computeOrbDescriptors derives it from the scaled keypoint
    const float inv = __fdiv_rn(1.f, sc);
    const int cx = __float2int_rn(__fmul_rn(px, inv)), cy = __float2int_rn(__fmul_rn(py, inv));
    const float ar = __fmul_rn(angle, g.deg2rad);
    const float ca = (float)cos((double)ar), sa = (float)sin((double)ar);
    const uint8_t* bc = blur + base + (size_t)cy * p + cx;
    unsigned byte = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t q = pat[i][lane];
        const float x0 = (float)(signed char)(q & 0xFF), y0 = (float)(signed char)((q >> 8) & 0xFF);
        const float x1 = (float)(signed char)((q >> 16) & 0xFF), y1 = (float)(signed char)(q >> 24);
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sa)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sa), __fmul_rn(y0, ca)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sa)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sa), __fmul_rn(y1, ca)));
        const int t0 = bc[iy0 * p + ix0], t1 = bc[iy1 * p + ix1];
        byte |= (t0 < t1 ? 1u : 0u) << i;
    }
    const size_t row = (size_t)s.frame * cap + s.slot;
    out_desc[row * 32 + lane] = (uint8_t)byte;
    if (lane == 0) {
        float* o = out_kp + row * 6;
        o[0] = px;
        o[1] = py;
        o[2] = __fmul_rn(31.f, sc);
        o[3] = angle;
        o[4] = s.harris;
        o[5] = (float)s.level;
    }
}

// KeyPointsFilter::retainBest on the host with the real std:: algorithms (select mode 1 and the fall-back of the
// device selection): same calls, same comparators, hence the same permutation as OpenCV's vector<KeyPoint>.
void retain_best(std::vector<RespIdx>& v, int n_points) {
    if (n_points >= 0 && v.size() > (size_t)n_points) {
        if (n_points == 0) { v.clear(); return; }
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(),
                         [](const RespIdx& a, const RespIdx& b) { return a.r > b.r; });
        const float ambiguous = v[n_points - 1].r;
        auto e = std::partition(v.begin() + n_points, v.end(), [ambiguous](const RespIdx& a) { return a.r >= ambiguous; });
        v.resize(e - v.begin());
    }
}

// ---- device-side selection path ----
struct LevelTab { int cand_off[kLevels]; int n_level[kLevels]; int cand_per_frame; int row_off[kLevels]; int rows_total; int hit_stride; };

__device__ __forceinline__ bool nms_max(const uint8_t* sc, int p) {
    const int s = sc[0];
    if (s == 0) return false;
    // eight independent loads, one compare: max of the neighbours (a short-circuit chain would serialise the loads)
    const int n0 = sc[-1], n1 = sc[1], n2 = sc[-p - 1], n3 = sc[-p], n4 = sc[-p + 1], n5 = sc[p - 1], n6 = sc[p], n7 = sc[p + 1];
    return s > max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
}

// NMS survivors of every (frame, level), one warp per row: the row's survivors in ascending x (position, FAST score)
// go to the row's slot of `hx` / `hs` (hit_stride entries per row: a row holds at most one strict maximum per two
// pixels), their number to rowcnt.  FAST's row-major order is then "rows in order, hits in order" - k_orb_select1
// turns the per-row counts into offsets and gathers.  rowcnt: [F][rows_total], rows of level l start at lt.row_off[l].
constexpr int kNmsRowsPerCta = 32;
// bit k of the result: byte k of w is non-zero
__device__ __forceinline__ unsigned nz_nibble(uint32_t w) {
    const uint32_t t = (w | ((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
    return ((t >> 7) * 0x10204080u) >> 28;
}
__global__ void __launch_bounds__(256) k_orb_nms_rows(const uint8_t* __restrict__ score, OrbGeom g, LevelTab lt, BlockMap bm,
                                                      int* __restrict__ rowcnt, unsigned short* __restrict__ hx,
                                                      uint8_t* __restrict__ hs) {
    __shared__ unsigned short corner_list[8][512];     // per warp: the corners of one 512-pixel step of its row
    const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int l, bx, by;
    block_of(bm, blockIdx.x, l, bx, by);
    int w, h, p;
    long long loff;
    level_of(bm, l, w, h, p, loff);
    const int rows = h - 2 * kEdge, cols = w - 2 * kEdge;
    if (cols <= 0) {        // a level too narrow for any keypoint still owns row slots: their counts must read 0
        for (int r = by * kNmsRowsPerCta + (int)threadIdx.x; r < min((by + 1) * kNmsRowsPerCta, rows); r += 256)
            rowcnt[(size_t)f * lt.rows_total + lt.row_off[l] + r] = 0;
        return;
    }
    for (int r = by * kNmsRowsPerCta + warp; r < min((by + 1) * kNmsRowsPerCta, rows); r += 8) {     // short rows: several per warp
    const uint8_t* row = score + (size_t)f * g.frame_bytes + loff + (size_t)(r + kEdge) * p;
    const size_t slot = (size_t)f * lt.rows_total + lt.row_off[l] + r;
    unsigned short* ox = hx + slot * lt.hit_stride;
    uint8_t* os = hs + slot * lt.hit_stride;
    const int vi_max = (w - kEdge - 1) / 16;                 // 16 scores (one uint4) per lane and step
    const uint4* rowv = reinterpret_cast<const uint4*>(row);
    const int xlo = kEdge, xhi = w - kEdge;
    unsigned short* list = corner_list[warp];
    int pos = 0;
    for (int vi0 = kEdge / 16; vi0 <= vi_max; vi0 += 32) {
        const int vi = vi0 + lane;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (vi <= vi_max) q = rowv[vi];
        // bit k: pixel 16 vi + k is a corner (non-zero score) inside the border
        unsigned nz = 0;
        if (q.x | q.y | q.z | q.w) {
            nz = nz_nibble(q.x) | nz_nibble(q.y) << 4 | nz_nibble(q.z) << 8 | nz_nibble(q.w) << 12;
            const int xb = 16 * vi;
            if (xb < xlo) nz &= (xlo - xb < 16) ? ~((1u << (xlo - xb)) - 1u) : 0u;
            if (xb + 16 > xhi) nz &= (xhi > xb) ? ((1u << (xhi - xb)) - 1u) : 0u;
        }
        const int c = __popc(nz);
        if (__ballot_sync(0xFFFFFFFFu, c > 0) == 0) continue;
        int incl = c;                                   // lanes hold ascending x: exclusive scan of the per-lane counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += u;
        }
        // corners cluster (a lane may hold six while most hold none): they go to a list in ascending x first, and the
        // 3x3 test - eight neighbour loads per corner - then runs on the list with all lanes busy
        int o = incl - c;
        while (nz) {
            const int k = __ffs(nz) - 1;
            nz &= nz - 1;
            list[o++] = (unsigned short)(16 * vi + k);
        }
        const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        __syncwarp();
        for (int i0 = 0; i0 < total; i0 += 32) {
            const int i = i0 + lane;
            int x = 0;
            bool hit = false;
            if (i < total) {
                x = list[i];
                hit = nms_max(row + x, p);
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
            if (hit) {
                const int at = pos + __popc(m & ((1u << lane) - 1u));
                ox[at] = (unsigned short)x;
                os[at] = row[x];
            }
            pos += __popc(m);
        }
        __syncwarp();
    }
    if (lane == 0) rowcnt[slot] = pos;
    }
}

// The two retainBest passes, one warp per level (lane 0 runs the sequential algorithms of orb_select.cuh, all lanes
// do the copies), one CTA of 8 warps per frame, the working arrays in shared memory when the frame's candidates fit
// (kSelSmemEntries; white-noise frames do not and work in global memory):
//   k_orb_select1   retainBest(2 n_level) on the FAST scores                       -> work[], kept1[]
//   k_orb_harris    Harris response of the survivors only, one warp per keypoint   -> work[].r
//   k_orb_select2   retainBest(n_level) on the Harris responses, the frame's keypoint list in level order
// flags[0] |= 1: introselect ran out of its recursion budget (host fall-back), |= 2: more keypoints than out_cap.
constexpr int kSelSmemEntries = 13312;       // x (8 B element + 2 x 2 B stopper-list slots) = 156 KB
constexpr int kSelSmemBytes = kSelSmemEntries * 12;

__device__ __forceinline__ RespIdx* sel_buffer(RespIdx* smem, int* soff, const int* counts, int l, RespIdx* global_buf, bool& in_smem) {
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kLevels; ++k) {
            soff[k] = t;
            t += counts[k];
        }
        soff[kLevels] = t;
    }
    __syncthreads();
    in_smem = soff[kLevels] <= kSelSmemEntries;
    return in_smem ? smem + soff[l] : global_buf;
}

// retainBest by the whole warp (stopper lists, orb_select.cuh) when positions fit 16 bits, else by lane 0
__device__ __forceinline__ int retain_any(RespIdx* v, int n, int n_points, unsigned short* Ls, unsigned short* Rs) {
    if (n <= 65535) return orbsel::warp_retain_best(v, n, n_points, Ls, Rs);
    int m = 0;
    if ((threadIdx.x & 31) == 0) m = orbsel::retain_best(v, n, n_points);
    m = __shfl_sync(0xFFFFFFFFu, m, 0);
    __syncwarp();
    return m;
}

constexpr int kSel1Threads = 1024;
__global__ void __launch_bounds__(kSel1Threads) k_orb_select1(OrbGeom g, LevelTab lt, int* __restrict__ rowcnt, const unsigned short* __restrict__ hx,
                                                     const uint8_t* __restrict__ hs, uint32_t* __restrict__ key,
                                                     RespIdx* __restrict__ work, unsigned short* __restrict__ lists,
                                                     int* __restrict__ kept1, int* __restrict__ flags) {
    extern __shared__ RespIdx sel_smem[];
    __shared__ int soff[kLevels + 1];
    __shared__ int counts[kLevels];
    const int f = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool selector = warp < kLevels;            // warps 0..7 own a level each; the other 24 only help with the gather
    const int l = selector ? warp : 0;
    const size_t ob = (size_t)f * lt.cand_per_frame + lt.cand_off[l];
    int* rc_frame = rowcnt + (size_t)f * lt.rows_total;
    // per-row counts -> exclusive offsets, in place; n = survivors of the level
    int n = 0;
    if (selector) {
        const int rows = max(g.h[l] - 2 * kEdge, 0);
        int* rc = rc_frame + lt.row_off[l];
        for (int r0 = 0; r0 < rows; r0 += 32) {
            const int r = r0 + lane;
            const int c = r < rows ? rc[r] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += u;
            }
            if (r < rows) rc[r] = n + incl - c;
            n += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        if (lane == 0) counts[l] = n;
    }
    __syncthreads();
    bool in_smem;
    RespIdx* v = sel_buffer(sel_smem, soff, counts, l, work + ob, in_smem);
    // gather in row-major order: one thread per row of the frame (all levels), the whole CTA
    for (int R = threadIdx.x; R < lt.rows_total; R += kSel1Threads) {
        int lr = 0;
#pragma unroll
        for (int k = 1; k < kLevels; ++k) lr += (R >= lt.row_off[k]) ? 1 : 0;
        const int r = R - lt.row_off[lr];
        const int rows = max(g.h[lr] - 2 * kEdge, 0);
        const int o = rc_frame[R], e = (r + 1 < rows) ? rc_frame[R + 1] : counts[lr];
        const size_t obr = (size_t)f * lt.cand_per_frame + lt.cand_off[lr];
        RespIdx* vr = in_smem ? sel_smem + soff[lr] : work + obr;
        const size_t hb = ((size_t)f * lt.rows_total + R) * lt.hit_stride;
        for (int k = 0; k < e - o; ++k) {
            vr[o + k] = RespIdx{(float)hs[hb + k], o + k};
            key[obr + o + k] = (uint32_t)((r + kEdge) << 16 | hx[hb + k]);
        }
    }
    __syncthreads();
    if (!selector) return;
    unsigned short* Ls = in_smem ? reinterpret_cast<unsigned short*>(sel_smem + kSelSmemEntries) + 2 * soff[l] : lists + 2 * ob;
    unsigned short* Rs = Ls + n;
    int m = retain_any(v, n, 2 * lt.n_level[l], Ls, Rs);
    if (m < 0) {
        if (lane == 0) atomicOr(flags, 1);
        m = 0;
    }
    if (in_smem)
        for (int i = lane; i < m; i += 32) work[ob + i] = v[i];
    if (lane == 0) kept1[f * kLevels + l] = m;
}

// grid (kLevels * 8, F): 64 warps per (frame, level) stride over its survivors; lanes = the 49 cells of the 7x7 block
__global__ void __launch_bounds__(256) k_orb_harris(const uint8_t* __restrict__ pyr, OrbGeom g, LevelTab lt, float harris_scale4,
                                                    const uint32_t* __restrict__ key, const int* __restrict__ kept1,
                                                    RespIdx* __restrict__ work) {
    const int f = blockIdx.y, l = blockIdx.x >> 3, lane = threadIdx.x & 31;
    const int w0 = (blockIdx.x & 7) * 8 + (threadIdx.x >> 5);
    const int m = kept1[f * kLevels + l], p = g.pitch[l];
    const size_t ob = (size_t)f * lt.cand_per_frame + lt.cand_off[l];
    const uint8_t* img = pyr + (size_t)f * g.frame_bytes + g.off[l];
    for (int i = w0; i < m; i += 64) {
        const uint32_t k = key[ob + work[ob + i].i];
        const uint8_t* c = img + (size_t)(k >> 16) * p + (k & 0xFFFF);
        int a = 0, b = 0, cc = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int cell = lane + 32 * t;
            if (cell < 49) {
                const uint8_t* q = c + (cell / 7 - 3) * p + (cell % 7 - 3);
                const int Ix = (q[1] - q[-1]) * 2 + (q[-p + 1] - q[-p - 1]) + (q[p + 1] - q[p - 1]);
                const int Iy = (q[p] - q[-p]) * 2 + (q[p - 1] - q[-p - 1]) + (q[p + 1] - q[-p + 1]);
                a += Ix * Ix;
                b += Iy * Iy;
                cc += Ix * Iy;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
            b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
            cc += __shfl_xor_sync(0xFFFFFFFFu, cc, o);
        }
        if (lane == 0) {
            const float fa = (float)a, fb = (float)b, fc = (float)cc;
            const float t = __fadd_rn(fa, fb);
            work[ob + i].r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, t), t)),
                                       harris_scale4);
        }
    }
}

__global__ void __launch_bounds__(256) k_orb_select2(LevelTab lt, const uint32_t* __restrict__ key, const int* __restrict__ kept1,
                                                     RespIdx* __restrict__ work, unsigned short* __restrict__ lists,
                                                     Sel* __restrict__ sel, int out_cap, int* __restrict__ n_out,
                                                     int* __restrict__ flags) {
    extern __shared__ RespIdx sel_smem[];
    __shared__ int soff[kLevels + 1];
    __shared__ int kept[kLevels];
    const int f = blockIdx.x, l = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t ob = (size_t)f * lt.cand_per_frame + lt.cand_off[l];
    bool in_smem;
    RespIdx* v = sel_buffer(sel_smem, soff, kept1 + f * kLevels, l, work + ob, in_smem);
    const int n = kept1[f * kLevels + l];
    unsigned short* Ls = in_smem ? reinterpret_cast<unsigned short*>(sel_smem + kSelSmemEntries) + 2 * soff[l] : lists + 2 * ob;
    unsigned short* Rs = Ls + n;
    if (in_smem)
        for (int i = lane; i < n; i += 32) v[i] = work[ob + i];
    __syncwarp();
    int m = retain_any(v, n, lt.n_level[l], Ls, Rs);
    if (m < 0) {
        if (lane == 0) atomicOr(flags, 1);
        m = 0;
    }
    if (lane == 0) kept[l] = m;
    __syncthreads();
    int first = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kLevels; ++k) {
        if (k < l) first += kept[k];
        total += kept[k];
    }
    if (total > out_cap) {
        if (threadIdx.x == 0) {
            atomicOr(flags, 2);
            n_out[f] = 0;
        }
        return;
    }
    if (threadIdx.x == 0) n_out[f] = total;
    for (int i = lane; i < m; i += 32) {
        const RespIdx r = v[i];
        const uint32_t k = key[ob + r.i];
        sel[(size_t)f * out_cap + first + i] = Sel{f, l, (int)(k & 0xFFFF), (int)(k >> 16), r.r, first + i};
    }
}

}  // namespace

struct nclt_orb {
    int W = 0, H = 0, max_frames = 0, nfeatures = 500;
    OrbGeom g;
    int n_level[kLevels];
    unsigned cand_cap_per_frame = 0;
    uint8_t *d_pyr = nullptr, *d_blur = nullptr, *d_score = nullptr, *d_in = nullptr;
    Cand* d_cand = nullptr;
    unsigned* d_ncand = nullptr;
    int* d_tab[kLevels] = {};
    Sel* d_sel = nullptr;
    float* d_kp = nullptr;
    uint8_t* d_desc = nullptr;
    int out_cap = 0;
    float harris_scale4 = 0.f;
    std::vector<Cand> h_cand;
    std::vector<Sel> h_sel;
    // device-side selection (select_mode 0)
    int select_mode = 0;
    LevelTab lt;
    uint32_t* d_key = nullptr;
    unsigned short* d_hx = nullptr;      // per-row NMS survivors: x, FAST score
    uint8_t* d_hs = nullptr;
    int* d_kept1 = nullptr;
    RespIdx* d_work = nullptr;
    unsigned short* d_lists = nullptr;     // stopper lists of the warp partitions when a frame does not fit shared memory
    int *d_nout = nullptr, *d_flags = nullptr, *d_rowcnt = nullptr;
    BlockMap bm_fast, bm_blur, bm_rows;
    int4* d_tabs = nullptr;       // block / level tables the three maps point into
    cudaStream_t side = nullptr;       // the blur runs beside FAST / NMS / selection
    cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr, ev_rs = nullptr;
    cudaEvent_t ev_lvl[kLevels] = {};      // level l of the pyramid exists (recorded on the side stream)
    int* h_pinned = nullptr;      // [0] flags, [1..] n_out
    unsigned long long host_fallbacks = 0;
    // a submitted, not yet awaited call (nclt_orb_submit / nclt_orb_wait)
    bool pending = false, p_out_on_device = false;
    int p_F = 0;
    float* p_kp = nullptr;
    uint8_t* p_desc = nullptr;
    int32_t* p_n = nullptr;
};

extern "C" int nclt_orb_destroy(nclt_ctx* c, nclt_orb* o) {
    if (!o) return NCLT_OK;
    if (c) cudaSetDevice(c->device);
    cudaFree(o->d_pyr); cudaFree(o->d_blur); cudaFree(o->d_score); cudaFree(o->d_in); cudaFree(o->d_cand);
    cudaFree(o->d_ncand); cudaFree(o->d_sel); cudaFree(o->d_kp); cudaFree(o->d_desc);
    cudaFree(o->d_key); cudaFree(o->d_hx); cudaFree(o->d_hs); cudaFree(o->d_kept1); cudaFree(o->d_work); cudaFree(o->d_lists);
    cudaFree(o->d_nout); cudaFree(o->d_flags); cudaFree(o->d_rowcnt); cudaFree(o->d_tabs);
    if (o->side) cudaStreamDestroy(o->side);
    if (o->ev_pyr) cudaEventDestroy(o->ev_pyr);
    if (o->ev_blur) cudaEventDestroy(o->ev_blur);
    if (o->ev_rs) cudaEventDestroy(o->ev_rs);
    for (int l = 0; l < kLevels; ++l) if (o->ev_lvl[l]) cudaEventDestroy(o->ev_lvl[l]);
    if (o->h_pinned) cudaFreeHost(o->h_pinned);
    for (int l = 0; l < kLevels; ++l) cudaFree(o->d_tab[l]);
    delete o;
    return NCLT_OK;
}

extern "C" int nclt_orb_create(nclt_ctx* c, int W, int H, int max_frames, int out_cap, nclt_orb** out) {
    if (!c) return NCLT_ERR_ARG;
    if (!out || W < 200 || H < 200 || W > 8192 || H > 8192 || max_frames <= 0 || out_cap < 500)
        return nclt_fail(c, NCLT_ERR_ARG, "orb_create: bad arguments (W, H in [200, 8192], out_cap >= 500)");
    cudaSetDevice(c->device);
    nclt_orb* o = new nclt_orb();
    o->W = W; o->H = H; o->max_frames = max_frames; o->out_cap = out_cap;
    OrbGeom& g = o->g;
    // ORB_Impl: scaleFactor is stored as a double holding the float 1.2f; getScale() = (float)pow(scaleFactor, level)
    const double scale_factor = (double)1.2f;
    long long off = 0;
    unsigned cand_cap = 0;
    for (int l = 0; l < kLevels; ++l) {
        g.scale[l] = (float)std::pow(scale_factor, (double)l);
        g.w[l] = (int)lrintf((float)W / g.scale[l]);
        g.h[l] = (int)lrintf((float)H / g.scale[l]);
        g.pitch[l] = (g.w[l] + 15) & ~15;
        g.off[l] = off;
        off += (long long)g.pitch[l] * g.h[l];
        const int iw = std::max(g.w[l] - 2 * kEdge, 0), ih = std::max(g.h[l] - 2 * kEdge, 0);
        o->lt.cand_off[l] = (int)cand_cap;
        cand_cap += (unsigned)(((iw + 1) / 2) * ((ih + 1) / 2));        // strict 3x3 maxima: at most one per 2x2 cell
    }
    o->lt.cand_per_frame = (int)cand_cap;
    {
        int rows_total = 0, nf = 0, nb = 0, nr = 0;
        for (int l = 0; l < kLevels; ++l) {
            const int rows = std::max(g.h[l] - 2 * kEdge, 0), cols = std::max(g.w[l] - 2 * kEdge, 0);
            o->lt.row_off[l] = rows_total;
            rows_total += rows;
            o->bm_fast.first[l] = nf; o->bm_fast.bx[l] = std::max((cols + 2 + kFastTile - 1) / kFastTile, 1);
            nf += (rows > 0 && cols > 0) ? o->bm_fast.bx[l] * ((rows + 2 + kFastTile - 1) / kFastTile) : 0;
            o->bm_blur.first[l] = nb; o->bm_blur.bx[l] = (g.w[l] + 31) / 32;
            nb += o->bm_blur.bx[l] * ((g.h[l] + kBlurH - 1) / kBlurH);
            o->bm_rows.first[l] = nr; o->bm_rows.bx[l] = 1;
            nr += std::max((rows + kNmsRowsPerCta - 1) / kNmsRowsPerCta, 1);
        }
        o->lt.rows_total = std::max(rows_total, 1);
        o->lt.hit_stride = (std::max(g.w[0] - 2 * kEdge, 0) + 1) / 2 + 1;
        o->bm_fast.first[kLevels] = nf; o->bm_blur.first[kLevels] = nb; o->bm_rows.first[kLevels] = nr;
    }
    g.frame_bytes = (off + 255) & ~255LL;
    o->cand_cap_per_frame = cand_cap;
    {   // device tables of the three block maps + the level geometry (block_of / level_of)
        std::vector<int4> tabs;
        size_t at[3];
        BlockMap* maps[3] = {&o->bm_fast, &o->bm_blur, &o->bm_rows};
        for (int m = 0; m < 3; ++m) {
            at[m] = tabs.size();
            for (int l = 0; l < kLevels; ++l) {
                const int n = maps[m]->first[l + 1] - maps[m]->first[l], nbx = std::max(maps[m]->bx[l], 1);
                for (int t = 0; t < n; ++t) tabs.push_back(make_int4(l, t % nbx, t / nbx, 0));
            }
        }
        const size_t lev_at = tabs.size();
        for (int l = 0; l < kLevels; ++l) tabs.push_back(make_int4(g.w[l], g.h[l], g.pitch[l], (int)g.off[l]));
        if (cudaMalloc((void**)&o->d_tabs, tabs.size() * sizeof(int4)) != cudaSuccess ||
            cudaMemcpy(o->d_tabs, tabs.data(), tabs.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) {
            nclt_orb_destroy(c, o);
            return nclt_fail(c, NCLT_ERR_NOMEM, "orb_create: block tables");
        }
        for (int m = 0; m < 3; ++m) { maps[m]->tab = o->d_tabs + at[m]; maps[m]->lev = o->d_tabs + lev_at; }
    }
    // features per level (ORB_Impl::detectAndCompute)
    {
        const float factor = (float)(1.0 / scale_factor);
        float nd = o->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kLevels));
        int sum = 0;
        for (int l = 0; l < kLevels - 1; ++l) {
            o->n_level[l] = (int)lrintf(nd);
            sum += o->n_level[l];
            nd *= factor;
        }
        o->n_level[kLevels - 1] = std::max(o->nfeatures - sum, 0);
        for (int l = 0; l < kLevels; ++l) o->lt.n_level[l] = o->n_level[l];
    }
    // disc of the intensity centroid
    {
        const int hp = kHalfPatch;
        int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
        for (int v = 0; v <= vmax; ++v) g.umax[v] = (int)lrint(std::sqrt((double)hp * hp - v * v));
        for (int v = hp, v0 = 0; v >= vmin; --v) {
            while (g.umax[v0] == g.umax[v0 + 1]) ++v0;
            g.umax[v] = v0;
            ++v0;
        }
    }
    const float rad2deg = (float)(180 / 3.1415926535897932384626433832795);
    g.atan_p1 = 0.9997878412794807f * rad2deg;
    g.atan_p3 = -0.3258083974640975f * rad2deg;
    g.atan_p5 = 0.1555786518463281f * rad2deg;
    g.atan_p7 = -0.04432655554792128f * rad2deg;
    g.deg2rad = (float)(3.1415926535897932384626433832795 / 180.f);
    // getGaussianKernel(7, 2, CV_32F): bit patterns of the four distinct taps (centre first)
    const uint32_t gk_bits[4] = {0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu, 0x3d8fafb1u};
    for (int i = 0; i < 4; ++i) memcpy(&g.gk[i], &gk_bits[i], 4);
    {
        const float sc = 1.f / ((1 << 2) * 7 * 255.f);
        o->harris_scale4 = sc * sc * sc * sc;
    }
    const size_t fb = (size_t)g.frame_bytes * max_frames;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&o->d_pyr, fb); A((void**)&o->d_blur, fb); A((void**)&o->d_score, fb);
    A((void**)&o->d_in, (size_t)W * H * 3 * max_frames);
    A((void**)&o->d_cand, (size_t)cand_cap * max_frames * sizeof(Cand));
    A((void**)&o->d_ncand, 256);
    A((void**)&o->d_key, (size_t)cand_cap * max_frames * 4);
    A((void**)&o->d_kept1, (size_t)max_frames * kLevels * 4);
    A((void**)&o->d_work, (size_t)cand_cap * max_frames * sizeof(RespIdx));
    A((void**)&o->d_lists, (size_t)cand_cap * max_frames * 2 * sizeof(unsigned short));
    A((void**)&o->d_nout, (size_t)max_frames * 4);
    A((void**)&o->d_flags, 256);
    A((void**)&o->d_rowcnt, (size_t)o->lt.rows_total * max_frames * 4);
    A((void**)&o->d_hx, (size_t)o->lt.rows_total * max_frames * o->lt.hit_stride * sizeof(unsigned short));
    A((void**)&o->d_hs, (size_t)o->lt.rows_total * max_frames * o->lt.hit_stride);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&o->side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->ev_pyr, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->ev_blur, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->ev_rs, cudaEventDisableTiming);
    for (int l = 1; l < kLevels; ++l) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&o->ev_lvl[l], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&o->h_pinned, (size_t)(max_frames + 1) * 4);
    if (e == cudaSuccess) e = cudaMemset(o->d_flags, 0, 256);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_orb_select1, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelSmemBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_orb_select2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelSmemBytes);
    A((void**)&o->d_sel, (size_t)out_cap * max_frames * sizeof(Sel));
    A((void**)&o->d_kp, (size_t)out_cap * max_frames * 6 * sizeof(float));
    A((void**)&o->d_desc, (size_t)out_cap * max_frames * 32);
    // resize tables (resize.cpp, bit-exact linear): src = (d + 0.5) * (src / dst) - 0.5 in double, weight round(frac * 256)
    for (int l = 1; l < kLevels && e == cudaSuccess; ++l) {
        const int sw = g.w[l - 1], sh = g.h[l - 1], dw = g.w[l], dh = g.h[l];
        std::vector<int> tab(2 * dw + 2 * dh);
        auto fill = [](int src, int dst, int* s, int* a) {
            const double scale = (double)src / dst;
            for (int d = 0; d < dst; ++d) {
                double f = (d + 0.5) * scale - 0.5;
                int i = (int)std::floor(f);
                double fr = f - i;
                if (i < 0) { fr = 0; i = 0; }
                if (i >= src - 1) { fr = 0; i = src - 1; }
                s[d] = i;
                a[d] = (int)lrint(fr * 256);
            }
        };
        fill(sw, dw, tab.data(), tab.data() + dw);
        fill(sh, dh, tab.data() + 2 * dw, tab.data() + 2 * dw + dh);
        A((void**)&o->d_tab[l], tab.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(o->d_tab[l], tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_pattern, kOrbPattern, sizeof(kOrbPattern));
    if (e == cudaSuccess) e = cudaMemset(o->d_score, 0, fb);
    if (e != cudaSuccess) {
        nclt_orb_destroy(c, o);
        return nclt_fail(c, e == cudaErrorMemoryAllocation ? NCLT_ERR_NOMEM : NCLT_ERR_CUDA, "orb_create", e);
    }
    *out = o;
    return NCLT_OK;
}

extern "C" int nclt_orb_levels(const nclt_orb* o, int32_t* out_w, int32_t* out_h, int32_t* out_n, float* out_scale) {
    if (!o) return NCLT_ERR_ARG;
    for (int l = 0; l < kLevels; ++l) {
        if (out_w) out_w[l] = o->g.w[l];
        if (out_h) out_h[l] = o->g.h[l];
        if (out_n) out_n[l] = o->n_level[l];
        if (out_scale) out_scale[l] = o->g.scale[l];
    }
    return NCLT_OK;
}

// phase 1, common to both selection modes: pyramid, FAST score map, blurred pyramid
// The blur needs only the pyramid.  It is queued on the side stream behind `after` (an event on the main stream) so
// that it runs beside the selection kernels, which occupy one lane per level and leave the SMs idle.
static int orb_blur_beside(nclt_ctx* c, nclt_orb* o, int F, bool record = true) {
    if (record) CU_TRY(c, cudaEventRecord(o->ev_pyr, c->stream));
    CU_TRY(c, cudaStreamWaitEvent(o->side, o->ev_pyr, 0));
    k_orb_blur<<<dim3(o->bm_blur.first[kLevels], F), 256, 0, o->side>>>(o->d_pyr, o->g, o->bm_blur, o->d_blur);
    CU_TRY(c, cudaEventRecord(o->ev_blur, o->side));
    return NCLT_OK;
}

static int orb_front(nclt_ctx* c, nclt_orb* o, const uint8_t* d_img, int channels, int F, bool blur_later = false) {
    const OrbGeom& g = o->g;
    int rc_blur = 0;
    cudaStream_t st = c->stream;
    if (d_img)      // nullptr: level 0 is already in place
        k_orb_level0<<<dim3((o->W + 255) / 256, (o->H + kResizeRows - 1) / kResizeRows, F), 256, 0, st>>>(d_img, channels, o->W, o->H, g.pitch[0], g.frame_bytes, o->d_pyr);
    // the seven dependent resizes are small, latency-bound launches: they run on the side stream while FAST already
    // works on level 0 (45 % of all pixels); FAST of levels 1-7 follows when the chain is done
    CU_TRY(c, cudaEventRecord(o->ev_pyr, st));
    CU_TRY(c, cudaStreamWaitEvent(o->side, o->ev_pyr, 0));
    for (int l = 1; l < kLevels; ++l) {
        k_orb_resize<<<dim3((g.w[l] + 255) / 256, (g.h[l] + kResizeRows - 1) / kResizeRows, F), 256, 0, o->side>>>(o->d_pyr + g.off[l - 1], g.w[l - 1], g.h[l - 1],
                                                                                  g.pitch[l - 1], o->d_pyr + g.off[l], g.w[l], g.h[l],
                                                                                  g.pitch[l], g.frame_bytes, o->d_tab[l]);
        if (l == 1 || l == 2 || l == kLevels - 1) CU_TRY(c, cudaEventRecord(o->ev_lvl[l], o->side));
    }
    // FAST in four launches - level 0 at once, levels 1 and 2 as soon as each exists, levels 3-7 when the chain is
    // through (by then FAST has been busy for longer than the chain takes): the main stream no longer idles at the end of
    // level 0 waiting for level 7
    const int fast_from[5] = {0, 1, 2, 3, kLevels};
    for (int q = 0; q < 4; ++q) {
        const int b0 = o->bm_fast.first[fast_from[q]], b1 = o->bm_fast.first[fast_from[q + 1]];
        if (q > 0) CU_TRY(c, cudaStreamWaitEvent(st, o->ev_lvl[q < 3 ? q : kLevels - 1], 0));
        if (b1 > b0) k_orb_fast<<<dim3(b1 - b0, F), 256, 0, st>>>(o->d_pyr, g, o->bm_fast, b0, o->d_score);
    }
    // beside the light kernels that follow; the device-selection path queues it itself, behind the NMS pass
    if (!blur_later && (rc_blur = orb_blur_beside(c, o, F))) return rc_blur;
    c->launches += 13;
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

// select mode 1 (and the fall-back of mode 0): candidates to the host, std::sort into FAST's row-major order, the two
// retainBest passes per (frame, level) with the real std:: algorithms.  Fills d_sel (dense) and n_out.
static int orb_select_host(nclt_ctx* c, nclt_orb* o, int F, std::vector<int32_t>& n_out, int* n_sel_out) {
    const OrbGeom& g = o->g;
    cudaStream_t st = c->stream;
    CU_TRY(c, cudaMemsetAsync(o->d_ncand, 0, 4, st));
    const dim3 gf((g.w[0] - 2 * kEdge + 2 + 31) / 32, (g.h[0] - 2 * kEdge + 2 + 7) / 8, F * kLevels);
    const unsigned cap = o->cand_cap_per_frame * (unsigned)F;
    k_orb_nms<<<gf, 256, 0, st>>>(o->d_pyr, o->d_score, g, o->d_cand, o->d_ncand, cap, o->harris_scale4);
    c->launches += 1;
    unsigned n_cand = 0;
    CU_TRY(c, cudaMemcpyAsync(&n_cand, o->d_ncand, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(c, cudaStreamSynchronize(st));
    if (n_cand > cap) return nclt_fail(c, NCLT_ERR_STATE, "orb: candidate list overflow (cannot happen for strict 3x3 maxima)");
    o->h_cand.resize(n_cand);
    if (n_cand) {
        CU_TRY(c, cudaMemcpyAsync(o->h_cand.data(), o->d_cand, (size_t)n_cand * sizeof(Cand), cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaStreamSynchronize(st));
    }
    std::sort(o->h_cand.begin(), o->h_cand.end(),
              [](const Cand& a, const Cand& b) { return a.tag != b.tag ? a.tag < b.tag : a.key < b.key; });
    o->h_sel.clear();
    n_out.assign(F, 0);
    std::vector<RespIdx> v;
    size_t i = 0;
    while (i < o->h_cand.size()) {
        size_t j = i;
        const uint32_t tag = o->h_cand[i].tag;
        while (j < o->h_cand.size() && o->h_cand[j].tag == tag) ++j;
        const int f = (int)(tag / kLevels), l = (int)(tag % kLevels);
        v.resize(j - i);
        for (size_t k = i; k < j; ++k) v[k - i] = RespIdx{o->h_cand[k].score, (int)k};
        retain_best(v, 2 * o->n_level[l]);
        for (RespIdx& r : v) r.r = o->h_cand[r.i].harris;
        retain_best(v, o->n_level[l]);
        for (const RespIdx& r : v) {
            const Cand& cd = o->h_cand[r.i];
            if (n_out[f] >= o->out_cap)
                return nclt_fail(c, NCLT_ERR_STATE, "orb: more keypoints than out_cap (response ties); raise out_cap");
            o->h_sel.push_back(Sel{f, l, (int)(cd.key & 0xFFFF), (int)(cd.key >> 16), cd.harris, n_out[f]++});
        }
        i = j;
    }
    *n_sel_out = (int)o->h_sel.size();
    if (*n_sel_out)
        CU_TRY(c, cudaMemcpyAsync(o->d_sel, o->h_sel.data(), (size_t)*n_sel_out * sizeof(Sel), cudaMemcpyHostToDevice, st));
    return NCLT_OK;
}

static int orb_finish(nclt_ctx* c, nclt_orb* o);
static int orb_host_tail(nclt_ctx* c, nclt_orb* o, int F, float* out_kp, uint8_t* out_desc, int32_t* out_n, bool out_on_device);

// enqueue = true: return with everything queued (device selection only); orb_finish() completes the call
static int orb_run(nclt_ctx* c, nclt_orb* o, const uint8_t* img, bool img_on_device, int channels, int F, float* out_kp,
                   uint8_t* out_desc, int32_t* out_n, bool out_on_device, bool enqueue_only = false) {
    if (!c) return NCLT_ERR_ARG;
    if (!o || !img || (channels != 1 && channels != 3) || F <= 0 || F > o->max_frames || !out_kp || !out_desc || !out_n)
        return nclt_fail(c, NCLT_ERR_ARG, "orb_detect_and_compute: bad arguments");
    if (o->pending) return nclt_fail(c, NCLT_ERR_STATE, "orb: a submitted call is pending on this handle (nclt_orb_wait first)");
    cudaSetDevice(c->device);
    const OrbGeom& g = o->g;
    cudaStream_t st = c->stream;
    const uint8_t* d_img = img;
    if (img_on_device && channels == 1 && g.pitch[0] == o->W) {
        // gray device frames: one strided device copy into level 0 instead of the byte-per-thread kernel
        CU_TRY(c, cudaMemcpy2DAsync(o->d_pyr, (size_t)g.frame_bytes, img, (size_t)o->W * o->H, (size_t)o->W * o->H, F,
                                    cudaMemcpyDeviceToDevice, st));
        d_img = nullptr;
    }
    if (!img_on_device) {
        if (channels == 1 && g.pitch[0] == o->W) {
            // gray host frames land directly in level 0 of the pyramid (one strided copy, no staging, no copy kernel)
            CU_TRY(c, cudaMemcpy2DAsync(o->d_pyr, (size_t)g.frame_bytes, img, (size_t)o->W * o->H, (size_t)o->W * o->H, F,
                                        cudaMemcpyHostToDevice, st));
            d_img = nullptr;
        } else {
            CU_TRY(c, cudaMemcpyAsync(o->d_in, img, (size_t)o->W * o->H * channels * F, cudaMemcpyHostToDevice, st));
            d_img = o->d_in;
        }
    }
    int rc = orb_front(c, o, d_img, channels, F, o->select_mode != 1);
    if (rc) return rc;
    float* d_kp = out_on_device ? out_kp : o->d_kp;
    uint8_t* d_desc = out_on_device ? out_desc : o->d_desc;
    if (o->select_mode != 1) {
        // everything stays on the device; one read of (flags, n_out) at the end
        k_orb_nms_rows<<<dim3(o->bm_rows.first[kLevels], F), 256, 0, st>>>(o->d_score, g, o->lt, o->bm_rows, o->d_rowcnt, o->d_hx, o->d_hs);
        const size_t sel_smem = (size_t)kSelSmemBytes;
        // the blur (throughput-bound, needs only the pyramid) starts when the NMS pass is through: it runs beside the two
        // selection kernels, whose few warps leave the SMs idle, instead of sharing the SMs with the NMS pass.  select1 is
        // queued first so that its CTAs are placed before the blur's fill the machine.
        CU_TRY(c, cudaEventRecord(o->ev_pyr, st));
        k_orb_select1<<<F, kSel1Threads, sel_smem, st>>>(g, o->lt, o->d_rowcnt, o->d_hx, o->d_hs, o->d_key, o->d_work, o->d_lists, o->d_kept1,
                                                o->d_flags);
        if ((rc = orb_blur_beside(c, o, F, false))) return rc;
        k_orb_harris<<<dim3(kLevels * 8, F), 256, 0, st>>>(o->d_pyr, g, o->lt, o->harris_scale4, o->d_key, o->d_kept1, o->d_work);
        k_orb_select2<<<F, 256, sel_smem, st>>>(o->lt, o->d_key, o->d_kept1, o->d_work, o->d_lists, o->d_sel, o->out_cap, o->d_nout,
                                                o->d_flags);
        CU_TRY(c, cudaStreamWaitEvent(st, o->ev_blur, 0));
        k_orb_describe<<<(F * o->out_cap + 7) / 8, 256, 0, st>>>(o->d_pyr, o->d_blur, g, o->d_sel, F * o->out_cap, o->out_cap,
                                                                 o->d_nout, d_kp, d_desc);
        c->launches += 5;
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, cudaMemcpyAsync(o->h_pinned, o->d_flags, 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaMemcpyAsync(o->h_pinned + 1, o->d_nout, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
        if (!out_on_device) {
            CU_TRY(c, cudaMemcpyAsync(out_kp, d_kp, (size_t)F * o->out_cap * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
            CU_TRY(c, cudaMemcpyAsync(out_desc, d_desc, (size_t)F * o->out_cap * 32, cudaMemcpyDeviceToHost, st));
        } else {
            CU_TRY(c, cudaMemcpyAsync(out_n, o->d_nout, (size_t)F * 4, cudaMemcpyDeviceToDevice, st));
        }
        o->pending = true;
        o->p_F = F; o->p_kp = out_kp; o->p_desc = out_desc; o->p_n = out_n; o->p_out_on_device = out_on_device;
        return enqueue_only ? NCLT_OK : orb_finish(c, o);
    }
    return orb_host_tail(c, o, F, out_kp, out_desc, out_n, out_on_device);
}

// second half of a device-selection call: wait, look at the flags, hand over to the host selection if asked to
static int orb_finish(nclt_ctx* c, nclt_orb* o) {
    if (!o->pending) return NCLT_OK;
    o->pending = false;
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    const int F = o->p_F;
    CU_TRY(c, cudaStreamSynchronize(st));
    int flags = o->h_pinned[0];
    if (flags) CU_TRY(c, cudaMemsetAsync(o->d_flags, 0, 4, st));
    if (o->select_mode == 2) flags |= 1;      // diagnostic: exercise the hand-over to the host
    if (flags & 2) return nclt_fail(c, NCLT_ERR_STATE, "orb: more keypoints than out_cap (response ties); raise out_cap");
    if (flags & 1) {
        o->host_fallbacks++;         // introselect left its quick-select phase somewhere: redo the selection on the host
        return orb_host_tail(c, o, F, o->p_kp, o->p_desc, o->p_n, o->p_out_on_device);
    }
    if (!o->p_out_on_device) memcpy(o->p_n, o->h_pinned + 1, (size_t)F * 4);
    return NCLT_OK;
}

// host selection (mode 1, or the fall-back of the device selection) + description + read-back; synchronous
static int orb_host_tail(nclt_ctx* c, nclt_orb* o, int F, float* out_kp, uint8_t* out_desc, int32_t* out_n, bool out_on_device) {
    const OrbGeom& g = o->g;
    cudaStream_t st = c->stream;
    float* d_kp = out_on_device ? out_kp : o->d_kp;
    uint8_t* d_desc = out_on_device ? out_desc : o->d_desc;
    int rc;
    std::vector<int32_t> n_out;
    int n_sel = 0;
    if ((rc = orb_select_host(c, o, F, n_out, &n_sel))) return rc;
    CU_TRY(c, cudaStreamWaitEvent(st, o->ev_blur, 0));
    if (n_sel) {
        k_orb_describe<<<(n_sel + 7) / 8, 256, 0, st>>>(o->d_pyr, o->d_blur, g, o->d_sel, n_sel, o->out_cap, nullptr, d_kp, d_desc);
        c->launches += 1;
        CU_TRY(c, cudaGetLastError());
    }
    if (out_on_device) {
        CU_TRY(c, cudaMemcpyAsync(out_n, n_out.data(), (size_t)F * 4, cudaMemcpyHostToDevice, st));
    } else {
        memcpy(out_n, n_out.data(), (size_t)F * 4);
        CU_TRY(c, cudaMemcpyAsync(out_kp, d_kp, (size_t)F * o->out_cap * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaMemcpyAsync(out_desc, d_desc, (size_t)F * o->out_cap * 32, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(c, cudaStreamSynchronize(st));
    return NCLT_OK;
}

extern "C" int nclt_orb_submit(nclt_ctx* c, nclt_orb* o, const uint8_t* img, int channels, int F, float* out_kp, uint8_t* out_desc,
                               int32_t* out_n) {
    return orb_run(c, o, img, false, channels, F, out_kp, out_desc, out_n, false, true);
}
extern "C" int nclt_orb_submit_dev(nclt_ctx* c, nclt_orb* o, const uint8_t* img, int channels, int F, float* out_kp,
                                   uint8_t* out_desc, int32_t* out_n) {
    return orb_run(c, o, img, true, channels, F, out_kp, out_desc, out_n, true, true);
}
extern "C" int nclt_orb_wait(nclt_ctx* c, nclt_orb* o) {
    if (!c) return NCLT_ERR_ARG;
    if (!o) return nclt_fail(c, NCLT_ERR_ARG, "orb_wait: null handle");
    return orb_finish(c, o);
}

/* select mode: 0 = selection on the device (default), 1 = on the host with the std:: algorithms,
 * 2 = diagnostic: device selection, then behave as if it had asked for the host fall-back */
extern "C" int nclt_orb_set_select(nclt_ctx* c, nclt_orb* o, int mode) {
    if (!c) return NCLT_ERR_ARG;
    if (!o || mode < 0 || mode > 2) return nclt_fail(c, NCLT_ERR_ARG, "orb_set_select: mode 0 (device), 1 (host) or 2 (diagnostic)");
    o->select_mode = mode;
    return NCLT_OK;
}
extern "C" long long nclt_orb_host_fallbacks(const nclt_orb* o) { return o ? (long long)o->host_fallbacks : -1; }

extern "C" int nclt_orb_detect_and_compute(nclt_ctx* c, nclt_orb* o, const uint8_t* img, int channels, int F, float* out_kp,
                                           uint8_t* out_desc, int32_t* out_n) {
    return orb_run(c, o, img, false, channels, F, out_kp, out_desc, out_n, false);
}
extern "C" int nclt_orb_detect_and_compute_dev(nclt_ctx* c, nclt_orb* o, const uint8_t* img, int channels, int F,
                                               float* out_kp, uint8_t* out_desc, int32_t* out_n) {
    return orb_run(c, o, img, true, channels, F, out_kp, out_desc, out_n, true);
}

extern "C" int nclt_orb_debug_plane(nclt_ctx* c, nclt_orb* o, int what, int frame, int level, uint8_t* out) {
    if (!c) return NCLT_ERR_ARG;
    if (!o || !out || what < 0 || what > 2 || frame < 0 || frame >= o->max_frames || level < 0 || level >= kLevels)
        return nclt_fail(c, NCLT_ERR_ARG, "orb_debug_plane: bad arguments");
    cudaSetDevice(c->device);
    const uint8_t* base = (what == 0 ? o->d_pyr : what == 1 ? o->d_score : o->d_blur) + (size_t)frame * o->g.frame_bytes + o->g.off[level];
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaMemcpy2D(out, o->g.w[level], base, o->g.pitch[level], o->g.w[level], o->g.h[level], cudaMemcpyDeviceToHost));
    return NCLT_OK;
}
