// K6-K8: teach-time map builder - depth back-projection, ordered double compaction, Bresenham
// free-space ray tracing and saturating log-odds, order-exact.
//
// Replaces:
//   a10 relay depth_cb                      scripts/common/tf_wall_clock_relay.py:868-887
//   a11 mapper transform + filters + [::4]  scripts/common/teach_run_depth_mapper.py:133-153
//   a12 world_to_pix (truncation)           teach_run_depth_mapper.py:120-123
//   a13 _bresenham_mark                     teach_run_depth_mapper.py:172-195
//   a14 save (threshold, flipud)            teach_run_depth_mapper.py:208-216
//
// The reference updates a float32 log-odds grid sequentially, with clamps at +-5 that make the
// result depend on the order of updates.  Here the grid holds exact integers in units of 0.2
// (free -2, occupied +7, clamp +-25; thresholds: occupied <=> u >= 4, free <=> u <= -6), and
// every update x -> clamp(x + a) is an element of the monoid {x -> min(max(x + s, lo), hi)},
// closed under composition.  So:
//   stage A (k_occ_frame, one CTA per frame, all SMs in parallel): build the frame's rays, count
//     passes/hits per cell in a shared-memory window histogram (the privatised histogram), turn
//     every touched cell into ONE composed map.  Cells that only saw passes or only hits have a
//     closed form (their counts); cells that saw both get their exact interleaving from per-cell ray
//     bitmaps.  The frame's window (bounding box of its rays) is written out densely.
//   stage B (k_occ_apply, one CTA per 32 x 32 tile of the GRID, one thread per cell, all SMs in parallel):
//     every tile collects, in frame order, the frames of the batch whose window intersects it and each
//     thread applies their maps to ITS cell in that order: grid[cell] = map_f(... map_1(grid[cell])).
//     No sort, no atomics, and nothing sequential across the chip - the order lives inside a thread.
// The result equals the sequential reference update for update, ray by ray, frame by frame.
#include "common.cuh"
#include "scratch.cuh"

#include <algorithm>
#include <cfloat>
#include <climits>

namespace {

constexpr int OT = 512;             // threads per frame CTA
constexpr int MAX_RAYS = 8192;      // rays per frame handled by the fast path (depth mode: <= 4800)
constexpr int WIN_CAP = 40960;      // cells in the shared-memory window (bbox of one frame's rays)
constexpr int SUPER = 8192;         // stage A: samples staged per compaction pass
constexpr int PER_T = SUPER / OT;   // consecutive samples per thread in the ranking phase (16)
constexpr int ST_PLANE = SUPER + SUPER / 16 + 16;   // padded plane of the staging area (3 float planes + 1 byte plane < window)
constexpr int MIXED_CAP = 256;      // mixed cells resolved per bitmap pass (a frame with more takes several passes)
constexpr int TILE = 32;            // stage B: TILE x TILE grid cells per CTA, one thread per cell
constexpr int F_CHUNK = 2048;       // frames per (stage A, stage B) pair: bounds the scratch of one call
constexpr uint32_t OOB = 0xFFFFFFFFu;
constexpr uint32_t MIXED = 0x80000000u;

enum FrameStatus : int { FR_OK = 0, FR_EMPTY = 1, FR_NOHEIGHT = 2, FR_SENSOR_OOB = 3, FR_FALLBACK = 4 };

struct FrameHdr {    // 32 bytes
    int status;
    int n_rays;      // len(pts_map) after [::4] (incl. out-of-grid endpoints)
    int r0, c0;      // sensor cell
    int rmin, cmin;  // window origin (grid cell of window entry 0)
    int wr, wc;      // window rows / columns
};

struct OccGeom {
    int GH, GW;
    double ox, oy, res;
};

struct FrameIn {
    // depth mode
    const float* depth;       // [F][Hd][Wd] or null
    const uint16_t* depth16;  // [F][Hd][Wd] or null
    int Hd, Wd, step;
    int compact;              // 1: the buffer holds only the sampled rows (v = 0, step, 2*step, ...): [F][ceil(Hd/step)][Wd]
    float fx, fy, cx, cy;
    // point mode
    const float* pts;         // [F][Nmax][3] or null
    const int* pts_n;         // [F]
    int Nmax;
    const double* T;          // [F][16] row-major map <- camera_link
};

__device__ __forceinline__ int to_pix(double x, double origin, double res) { return (int)((x - origin) / res); }
// one coordinate of T @ [p;1] in OpenBLAS dgemm's order: an FMA chain over k (verified bit-exact)
__device__ __forceinline__ double xform(const double* t, double x, double y, double z) {
    return fma(t[3], 1.0, fma(t[2], z, fma(t[1], y, t[0] * x)));
}

// sample `s` of frame `f` -> camera_link point; returns relay validity
__device__ __forceinline__ bool sample_point(const FrameIn& in, int f, int s, int cols, float& X, float& Y, float& Z) {
    if (in.pts) {
        const float* p = in.pts + ((size_t)f * in.Nmax + s) * 3;
        X = p[0]; Y = p[1]; Z = p[2];
        return isfinite(X) && isfinite(Y) && isfinite(Z);      // _parse_pc2 finite filter
    }
    int v = (s / cols) * in.step, u = (s % cols) * in.step;
    size_t off = in.compact ? ((size_t)f * ((in.Hd + in.step - 1) / in.step) + v / in.step) * in.Wd + u
                            : ((size_t)f * in.Hd + v) * in.Wd + u;
    float z = in.depth ? in.depth[off] : (float)in.depth16[off] / 1000.0f;
    bool ok = (z > 0.3f) && (z < 10.0f) && isfinite(z);
    float px = ((float)u - in.cx) / in.fx * z;      // float32, this order, no FMA (-fmad=false)
    float py = ((float)v - in.cy) / in.fy * z;
    X = z; Y = -px; Z = -py;
    return ok;
}

// x -> min(max(x + s, lo), hi) in 19 bits
__device__ __forceinline__ uint32_t pack_map(int s, int lo, int hi) {
    s = s > 50 ? 50 : (s < -50 ? -50 : s);     // |s| >= 50 already saturates on [-25, 25]
    return ((uint32_t)(s + 64) << 12) | ((uint32_t)(lo + 32) << 6) | (uint32_t)(hi + 32);
}
// one window entry applied to a cell value: 0 = untouched, MIXED | id = composed map mix[id], else passes | hits << 16
__device__ __forceinline__ int apply_entry(int x, uint32_t v, const uint32_t* __restrict__ mix) {
    if (v == 0) return x;
    if (v & MIXED) {
        const uint32_t m = __ldg(mix + (v & 0x7FFFFFFFu));
        const int s = (int)((m >> 12) & 127) - 64, lo = (int)((m >> 6) & 63) - 32, hi = (int)(m & 63) - 32;
        return min(max(x + s, lo), hi);
    }
    const int nf = (int)(v & 0xFFFFu), no = (int)(v >> 16);
    return nf ? max(x - 2 * nf, -25) : min(x + 7 * no, 25);      // only passes / only hits: closed form
}

template <typename F>
__device__ __forceinline__ void bresenham(int r0, int c0, int r1, int c1, F&& visit) {
    int dr = abs(r1 - r0), dc = abs(c1 - c0);
    int sr = r0 < r1 ? 1 : -1, sc = c0 < c1 ? 1 : -1;
    int err = dr - dc, r = r0, c = c0;
    for (;;) {
        bool end = (r == r1) && (c == c1);
        visit(r, c, end);
        if (end) return;
        int e2 = 2 * err;
        if (e2 > -dc) { err -= dc; r += sr; }
        if (e2 < dr) { err += dr; c += sc; }
    }
}

// ---------------------------------------------------------------------------------------
// stage A
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OT) k_occ_frame(FrameIn in, OccGeom g, int F, FrameHdr* hdr,
                                                  uint32_t* wing /*[F][WIN_CAP] dense windows*/,
                                                  uint32_t* mixg /*[F][ray_cap] composed maps of the mixed cells*/,
                                                  uint32_t* rays_g /*[F][ray_cap]*/, int ray_cap,
                                                  uint32_t* bitmaps /*[gridDim.x][MIXED_CAP][2][nw_cap]*/, int nw_cap,
                                                  float* out_pts /*[F][cap][3] or null*/, int* out_pts_n, int pts_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* win = reinterpret_cast<uint32_t*>(smem_raw);                       // WIN_CAP
    uint32_t* rays = win + WIN_CAP;                                              // MAX_RAYS
    unsigned short* mixed_cell = reinterpret_cast<unsigned short*>(rays + MAX_RAYS);   // MAX_RAYS
    __shared__ int s_warp[OT / 32];
    __shared__ int s_bbox[4];
    __shared__ int s_cnt[4];      // n_mixed, n_delta, overflow flag
    __shared__ double sT[16];

    const int tid = threadIdx.x;
    for (int f = blockIdx.x; f < F; f += gridDim.x) {
        __syncthreads();
        if (tid < 16) sT[tid] = in.T[(size_t)f * 16 + tid];
        if (tid == 0) {
            s_bbox[0] = INT_MAX; s_bbox[1] = INT_MIN; s_bbox[2] = INT_MAX; s_bbox[3] = INT_MIN;
            s_cnt[0] = 0; s_cnt[1] = 0; s_cnt[2] = 0;
        }
        __syncthreads();
        const int cols = in.pts ? 1 : (in.Wd + in.step - 1) / in.step;
        const int rows = in.pts ? 1 : (in.Hd + in.step - 1) / in.step;
        const int n_samples = in.pts ? in.pts_n[f] : rows * cols;
        const int r0 = to_pix(sT[7], g.oy, g.res), c0 = to_pix(sT[3], g.ox, g.res);
        const bool sensor_ok = (0 <= r0 && r0 < g.GH && 0 <= c0 && c0 < g.GW);

        // ---- a10/a11: points, filters, ordered double compaction, every 4th ---------------
        // SUPER samples per pass.  Phase 1 (coalesced, all loads of a thread in flight at once): sample -> point +
        // relay validity, staged in shared memory (the window region, not needed yet).  Phase 2: every thread owns
        // PER_T CONSECUTIVE samples, counts its survivors of both filters, ONE block scan of the packed pair of
        // counts gives its ranks, and it emits its points / rays in order.  (The first version ranked 512 samples
        // at a time with two block scans each: 76 latency-exposed rounds per 640 x 480 frame.)
        int n_valid1 = 0, n_kept = 0;
        bool overflow = false;
        {
            float* stX = reinterpret_cast<float*>(win);
            float* stY = stX + ST_PLANE;
            float* stZ = stY + ST_PLANE;
            unsigned char* stV = reinterpret_cast<unsigned char*>(stZ + ST_PLANE);
            for (int s0 = 0; s0 < n_samples; s0 += SUPER) {
                const int n_here = min(SUPER, n_samples - s0);
#pragma unroll 4
                for (int j = tid; j < n_here; j += OT) {
                    float X, Y, Z;
                    const bool v1 = sample_point(in, f, s0 + j, cols, X, Y, Z);
                    const int a = j + (j >> 4);                 // one pad word per 16: phase 2 reads are conflict free
                    stX[a] = X; stY[a] = Y; stZ[a] = Z; stV[a] = v1 ? 1 : 0;
                }
                __syncthreads();
                const int j0 = tid * PER_T;
                unsigned m1 = 0, m2 = 0;
#pragma unroll
                for (int i = 0; i < PER_T; ++i) {
                    const int j = j0 + i, a = j + (j >> 4);
                    if (j < n_here && stV[a]) {
                        m1 |= 1u << i;
                        const double zm = xform(sT + 8, (double)stX[a], (double)stY[a], (double)stZ[a]);
                        if ((zm > 0.2) && (zm < 2.0)) m2 |= 1u << i;
                    }
                }
                // exclusive block scan of (count1 | count2 << 16): both totals <= SUPER < 65536
                const int mine = __popc(m1) | (__popc(m2) << 16);
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if ((tid & 31) >= o) incl += v;
                }
                if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
                __syncthreads();
                int before = 0, total = 0;
#pragma unroll
                for (int w = 0; w < OT / 32; ++w) {
                    const int v = s_warp[w];
                    total += v;
                    if (w < (tid >> 5)) before += v;
                }
                const int excl = before + incl - mine;
                int rank1 = n_valid1 + (excl & 0xFFFF), k = n_kept + (excl >> 16);
#pragma unroll
                for (int i = 0; i < PER_T; ++i) {
                    if (!((m1 >> i) & 1u)) continue;
                    const int j = j0 + i, a = j + (j >> 4);
                    const float X = stX[a], Y = stY[a], Z = stZ[a];
                    if (out_pts && rank1 < pts_cap) {
                        float* o = out_pts + ((size_t)f * pts_cap + rank1) * 3;
                        o[0] = X; o[1] = Y; o[2] = Z;
                    }
                    ++rank1;
                    if (!((m2 >> i) & 1u)) continue;
                    if ((k & 3) == 0) {
                        const int ray = k >> 2;
                        const double xm = xform(sT, (double)X, (double)Y, (double)Z);
                        const double ym = xform(sT + 4, (double)X, (double)Y, (double)Z);
                        const int c1 = to_pix(xm, g.ox, g.res), r1 = to_pix(ym, g.oy, g.res);
                        const bool inb = (0 <= r1 && r1 < g.GH && 0 <= c1 && c1 < g.GW);
                        const uint32_t packed = inb ? ((uint32_t)r1 << 16 | (uint32_t)c1) : OOB;
                        if (ray < MAX_RAYS) rays[ray] = packed;
                        if (ray < ray_cap) rays_g[(size_t)f * ray_cap + ray] = packed;
                        if (inb) {
                            atomicMin(&s_bbox[0], r1); atomicMax(&s_bbox[1], r1);
                            atomicMin(&s_bbox[2], c1); atomicMax(&s_bbox[3], c1);
                        }
                    }
                    ++k;
                }
                n_valid1 += total & 0xFFFF;
                n_kept += total >> 16;
                __syncthreads();                                // the stage is overwritten by the next pass / the window
            }
        }
        if (out_pts_n && tid == 0) out_pts_n[f] = n_valid1;
        const int n_rays = (n_kept + 3) >> 2;
        if (n_rays > MAX_RAYS) overflow = true;
        __syncthreads();

        int status = FR_OK;
        if (n_valid1 == 0) status = FR_EMPTY;
        else if (n_kept == 0) status = FR_NOHEIGHT;
        else if (!sensor_ok) status = FR_SENSOR_OOB;
        int wr = 0, wc = 0, rmin = 0, cmin = 0;
        if (status == FR_OK) {
            rmin = min(s_bbox[0], r0); int rmax = max(s_bbox[1], r0);
            cmin = min(s_bbox[2], c0); int cmax = max(s_bbox[3], c0);
            wr = rmax - rmin + 1; wc = cmax - cmin + 1;
            if (overflow || (long long)wr * wc > WIN_CAP || n_rays > ray_cap) status = FR_FALLBACK;
        }
        if (status != FR_OK) {
            if (tid == 0) {
                FrameHdr h{};
                h.status = status; h.n_rays = n_rays; h.r0 = r0; h.c0 = c0;
                hdr[f] = h;
            }
            continue;
        }
        const int ncell = wr * wc;
        for (int i = tid; i < ncell; i += OT) win[i] = 0;
        __syncthreads();

        // ---- a13 pass 1: privatised pass/hit histogram in shared memory --------------------
        for (int i = tid; i < n_rays; i += OT) {
            uint32_t pk = rays[i];
            if (pk == OOB) continue;
            int r1 = (int)(pk >> 16), c1 = (int)(pk & 0xFFFF);
            bresenham(r0, c0, r1, c1, [&](int r, int c, bool end) {
                atomicAdd(&win[(r - rmin) * wc + (c - cmin)], end ? 0x10000u : 1u);
            });
        }
        __syncthreads();

        // ---- classify touched cells: pure cells keep their counts (closed-form map), mixed cells get an id ---
        for (int i = tid; i < ncell; i += OT) {
            uint32_t v = win[i];
            if (v == 0) continue;
            if ((v & 0xFFFFu) && (v >> 16)) {
                int id = atomicAdd(&s_cnt[0], 1);
                win[i] = MIXED | (uint32_t)id;
                mixed_cell[id] = (unsigned short)i;
            }
        }
        __syncthreads();
        const int n_mixed = s_cnt[0];       // <= distinct endpoint cells <= n_rays <= ray_cap
        uint32_t* mixf = mixg + (size_t)f * ray_cap;
        // ---- mixed cells: exact interleaving of passes and hits by ray index, MIXED_CAP cells per pass ----
        const int nw = (n_rays + 31) >> 5;
        uint32_t* bm = bitmaps + (size_t)blockIdx.x * MIXED_CAP * 2 * nw_cap;   // [slot][2][nw_cap]
        for (int base = 0; base < n_mixed; base += MIXED_CAP) {
            const int cnt = min(MIXED_CAP, n_mixed - base);
            for (int i = tid; i < cnt * 2 * nw; i += OT) bm[(size_t)(i / nw) * nw_cap + (i % nw)] = 0;
            __syncthreads();
            for (int i = tid; i < n_rays; i += OT) {
                uint32_t pk = rays[i];
                if (pk == OOB) continue;
                int r1 = (int)(pk >> 16), c1 = (int)(pk & 0xFFFF);
                bresenham(r0, c0, r1, c1, [&](int r, int c, bool end) {
                    uint32_t v = win[(r - rmin) * wc + (c - cmin)];
                    if (v & MIXED) {
                        int slot = (int)(v & 0x7FFFFFFFu) - base;
                        if (slot >= 0 && slot < cnt)
                            atomicOr(&bm[((size_t)slot * 2 + (end ? 1 : 0)) * nw_cap + (i >> 5)], 1u << (i & 31));
                    }
                });
            }
            __syncthreads();
            for (int slot = tid; slot < cnt; slot += OT) {
                const uint32_t* pb = bm + (size_t)slot * 2 * nw_cap;
                const uint32_t* hb = pb + nw_cap;
                int sft = 0, lo = -1000000, hi = 1000000;
                for (int w = 0; w < nw; ++w) {
                    uint32_t p = pb[w], h = hb[w], m = p | h;
                    while (m) {
                        int b = __ffs(m) - 1;
                        m &= m - 1;
                        int a = ((h >> b) & 1u) ? 7 : -2;
                        sft += a;
                        lo = min(max(lo + a, -25), 25);
                        hi = min(max(hi + a, -25), 25);
                    }
                }
                mixf[base + slot] = pack_map(sft, lo, hi);
            }
            __syncthreads();
        }
        // ---- the frame's dense window -> global (stage B gathers from it, tile by tile) ---------------------
        uint32_t* wg = wing + (size_t)f * WIN_CAP;
        for (int i = tid; i < ncell; i += OT) wg[i] = win[i];
        if (tid == 0) {
            FrameHdr h{};
            h.status = FR_OK; h.n_rays = n_rays; h.r0 = r0; h.c0 = c0;
            h.rmin = rmin; h.cmin = cmin; h.wr = wr; h.wc = wc;
            hdr[f] = h;
        }
    }
}

// ---------------------------------------------------------------------------------------
// stage B: one CTA per TILE x TILE block of the grid, one thread per cell.  The frames whose window intersects
// the tile are collected IN ORDER (ordered block compaction of the header test) and every thread folds their
// entries for its own cell into its register copy of the cell: frames strictly in order per cell, cells and
// tiles in parallel on every SM.  Frames beyond stage A's capacity (FR_FALLBACK: window > WIN_CAP cells or more
// than MAX_RAYS rays) are walked ray by ray, in order, by the tile's first thread on a shared-memory copy of the
// tile - rays that cannot touch the tile are skipped; exact, parallel across tiles.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE * TILE) k_occ_apply(int* grid, OccGeom g, int F, const FrameHdr* __restrict__ hdr,
                                                           const uint32_t* __restrict__ wing,
                                                           const uint32_t* __restrict__ mixg,
                                                           const uint32_t* __restrict__ rays_g, int ray_cap, int tiles_c,
                                                           long long* counters) {
    constexpr int NT = TILE * TILE;
    __shared__ int s_list[NT];             // frames of this chunk that touch the tile, ascending
    __shared__ int4 s_win[NT];             // their (rmin, cmin, wr, wc); wr < 0 marks a fallback frame
    __shared__ int s_warp[NT / 32];
    __shared__ int s_x[NT];                // tile copy for the fallback walk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr0 = (blockIdx.x / tiles_c) * TILE, tc0 = (blockIdx.x % tiles_c) * TILE;
    const int r = tr0 + (tid >> 5), c = tc0 + (tid & 31);
    const bool inside = r < g.GH && c < g.GW;
    int x = inside ? grid[(size_t)r * g.GW + c] : 0;
    bool dirty = false;
    long long integ = 0, pts = 0, empty = 0;
    for (int f0 = 0; f0 < F; f0 += NT) {
        const int f = f0 + tid;
        bool hit = false;
        int4 w = make_int4(0, 0, 0, 0);
        if (f < F) {
            const FrameHdr h = hdr[f];
            if (h.status == FR_OK) {
                hit = h.rmin < tr0 + TILE && h.rmin + h.wr > tr0 && h.cmin < tc0 + TILE && h.cmin + h.wc > tc0;
                w = make_int4(h.rmin, h.cmin, h.wr, h.wc);
            } else if (h.status == FR_FALLBACK) {
                hit = true;
                w = make_int4(h.r0, h.c0, -1, h.n_rays);
            }
            if (blockIdx.x == 0) {
                if (h.status == FR_EMPTY) empty++;
                else if (h.status == FR_OK || h.status == FR_FALLBACK) { integ++; pts += h.n_rays; }
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = 0, n_list = 0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) {
            const int v = s_warp[k];
            n_list += v;
            if (k < warp) before += v;
        }
        if (hit) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            s_list[pos] = f;
            s_win[pos] = w;
        }
        __syncthreads();
        for (int k0 = 0; k0 < n_list; k0 += 4) {
            // up to four independent loads in flight, applied in frame order
            uint32_t v[4] = {0, 0, 0, 0};
            bool fb = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + u < n_list) {
                    const int4 ww = s_win[k0 + u];
                    if (ww.z < 0) fb = true;
                    else {
                        const int rr = r - ww.x, cc = c - ww.y;
                        if (rr >= 0 && rr < ww.z && cc >= 0 && cc < ww.w)
                            v[u] = __ldg(wing + (size_t)s_list[k0 + u] * WIN_CAP + rr * ww.w + cc);
                    }
                }
            }
            if (!fb) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (v[u]) { x = apply_entry(x, v[u], mixg + (size_t)s_list[k0 + u] * ray_cap); dirty = true; }
                continue;
            }
            for (int u = 0; u < 4 && k0 + u < n_list; ++u) {       // block-uniform: fb comes from shared memory
                const int4 ww = s_win[k0 + u];
                const int ff = s_list[k0 + u];
                if (ww.z >= 0) {
                    if (v[u]) { x = apply_entry(x, v[u], mixg + (size_t)ff * ray_cap); dirty = true; }
                    continue;
                }
                s_x[tid] = x;
                __syncthreads();
                if (tid == 0) {
                    const int fr0 = ww.x, fc0 = ww.y, n = min(ww.w, ray_cap);
                    for (int i = 0; i < n; ++i) {
                        const uint32_t pk = rays_g[(size_t)ff * ray_cap + i];
                        if (pk == OOB) continue;
                        const int r1 = (int)(pk >> 16), c1 = (int)(pk & 0xFFFF);
                        if (max(fr0, r1) < tr0 || min(fr0, r1) >= tr0 + TILE || max(fc0, c1) < tc0 || min(fc0, c1) >= tc0 + TILE)
                            continue;
                        bresenham(fr0, fc0, r1, c1, [&](int rr, int cc, bool end) {
                            const int lr = rr - tr0, lc = cc - tc0;
                            if (lr >= 0 && lr < TILE && lc >= 0 && lc < TILE) {
                                int* p = &s_x[lr * TILE + lc];
                                *p = end ? min(*p + 7, 25) : max(*p - 2, -25);
                            }
                        });
                    }
                }
                __syncthreads();
                x = s_x[tid];
                dirty = true;
                __syncthreads();
            }
        }
        __syncthreads();
    }
    if (inside && dirty) grid[(size_t)r * g.GW + c] = x;
    if (blockIdx.x == 0) {
        // frame counters (mapper:94-97): order independent
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            integ += __shfl_xor_sync(0xFFFFFFFFu, integ, o);
            pts += __shfl_xor_sync(0xFFFFFFFFu, pts, o);
            empty += __shfl_xor_sync(0xFFFFFFFFu, empty, o);
        }
        if (lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long*>(counters), (unsigned long long)integ);
            atomicAdd(reinterpret_cast<unsigned long long*>(counters + 1), (unsigned long long)pts);
            atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), (unsigned long long)empty);
        }
    }
}

// a14: threshold + flipud ; and the float32 view of the grid
__global__ void k_occ_render(const int* __restrict__ grid, int GH, int GW, unsigned char* img, float* logodds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GH * GW) return;
    int r = i / GW, c = i % GW;
    int u = grid[i];
    if (img) {
        unsigned char v = 205;
        if (u >= 4) v = 0;       // 0.2u > ln(0.65/0.35) = 0.619  <=>  u >= 4
        if (u <= -6) v = 254;    // 0.2u < ln(0.25/0.75) = -1.0986 <=> u <= -6
        img[(size_t)(GH - 1 - r) * GW + c] = v;
    }
    if (logodds) logodds[i] = (float)u * 0.2f;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// handle + launchers
// ---------------------------------------------------------------------------------------
struct nclt_occ {
    int device = 0;
    OccGeom g{};
    int* d_grid = nullptr;
    long long* d_counters = nullptr;
};

static const size_t OCC_SMEM = (size_t)WIN_CAP * 4 + (size_t)MAX_RAYS * 4 + (size_t)MAX_RAYS * 2;

extern "C" int nclt_occ_create(nclt_ctx* c, double origin_x, double origin_y, double res, int W, int H,
                               nclt_occ** out) {
    if (!c || !out || W <= 0 || H <= 0 || W > 65535 || H > 65535 || !(res > 0) || (long long)W * H > (1LL << 28))
        return nclt_fail(c, NCLT_ERR_ARG, "occ_create args");
    *out = nullptr;
    cudaSetDevice(c->device);
    nclt_occ* o = new nclt_occ();
    o->device = c->device;
    o->g.GH = H; o->g.GW = W; o->g.ox = origin_x; o->g.oy = origin_y; o->g.res = res;
    // the handle owns the grid and three counters only; every per-batch buffer (windows, ray lists, the mixed-cell
    // bitmaps) comes from the context scratch inside the integrate calls
    cudaError_t e = cudaMalloc(&o->d_grid, (size_t)W * H * 4);
    if (e == cudaSuccess) e = cudaMalloc(&o->d_counters, 3 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(o->d_grid, 0, (size_t)W * H * 4, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(o->d_counters, 0, 3 * sizeof(long long), c->stream);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_occ_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OCC_SMEM);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        if (o->d_grid) cudaFree(o->d_grid);
        if (o->d_counters) cudaFree(o->d_counters);
        delete o;
        return nclt_fail(c, e == cudaErrorMemoryAllocation ? NCLT_ERR_NOMEM : NCLT_ERR_CUDA, "occ_create", e);
    }
    *out = o;
    return NCLT_OK;
}

extern "C" int nclt_occ_destroy(nclt_ctx* c, nclt_occ* o) {
    if (!o) return NCLT_OK;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    cudaFree(o->d_grid);
    cudaFree(o->d_counters);
    delete o;
    return NCLT_OK;
}

extern "C" int nclt_occ_reset(nclt_ctx* c, nclt_occ* o) {
    if (!c || !o) return NCLT_ERR_ARG;
    cudaSetDevice(c->device);
    CU_TRY(c, cudaMemsetAsync(o->d_grid, 0, (size_t)o->g.GW * o->g.GH * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(o->d_counters, 0, 3 * sizeof(long long), c->stream));
    return NCLT_OK;
}

// shared driver: device pointers only.  Frames go through (stage A, stage B) pairs of at most F_CHUNK frames, so the
// scratch of one call is bounded (F_CHUNK x (164 KB window + ray list + mixed maps) + the per-CTA bitmaps).
static int occ_integrate_dev(nclt_ctx* c, nclt_occ* o, FrameIn in, int F, float* out_pts, int* out_pts_n,
                             int pts_cap) {
    if (F <= 0) return NCLT_OK;
    cudaSetDevice(c->device);
    int ray_cap;
    if (in.pts) ray_cap = in.Nmax / 4 + 1;
    else ray_cap = (((in.Hd + in.step - 1) / in.step) * ((in.Wd + in.step - 1) / in.step)) / 4 + 1;
    const int nw_cap = (std::min(ray_cap, MAX_RAYS) + 31) / 32;
    const int fc = std::min(F, F_CHUNK);
    const int gridA = std::min(fc, c->sm_count);
    ScratchScope scope(c);
    int rc;
    size_t need = pad256((size_t)fc * sizeof(FrameHdr)) + pad256((size_t)fc * WIN_CAP * 4) + 2 * pad256((size_t)fc * ray_cap * 4) +
                  pad256((size_t)gridA * MIXED_CAP * 2 * nw_cap * 4);
    if ((rc = nclt_scratch_reserve(c, need))) return rc;
    Carver cv(c);
    FrameHdr* hdr = cv.take<FrameHdr>(fc);
    uint32_t* wing = cv.take<uint32_t>((size_t)fc * WIN_CAP);
    uint32_t* mixg = cv.take<uint32_t>((size_t)fc * ray_cap);
    uint32_t* rays_g = cv.take<uint32_t>((size_t)fc * ray_cap);
    uint32_t* bitmaps = cv.take<uint32_t>((size_t)gridA * MIXED_CAP * 2 * nw_cap);
    const int tiles_c = (o->g.GW + TILE - 1) / TILE, tiles_r = (o->g.GH + TILE - 1) / TILE;
    for (int f0 = 0; f0 < F; f0 += F_CHUNK) {
        const int n = std::min(F_CHUNK, F - f0);
        FrameIn part = in;
        if (in.pts) {
            part.pts = in.pts + (size_t)f0 * in.Nmax * 3;
            part.pts_n = in.pts_n + f0;
        } else {
            const size_t rows = in.compact ? (size_t)((in.Hd + in.step - 1) / in.step) : (size_t)in.Hd;
            if (in.depth) part.depth = in.depth + (size_t)f0 * rows * in.Wd;
            if (in.depth16) part.depth16 = in.depth16 + (size_t)f0 * rows * in.Wd;
        }
        part.T = in.T + (size_t)f0 * 16;
        nclt_prof_mark_tag(c, 1);
        k_occ_frame<<<std::min(n, c->sm_count), OT, OCC_SMEM, c->stream>>>(
            part, o->g, n, hdr, wing, mixg, rays_g, ray_cap, bitmaps, nw_cap, out_pts ? out_pts + (size_t)f0 * pts_cap * 3 : nullptr,
            out_pts_n ? out_pts_n + f0 : nullptr, pts_cap);
        nclt_prof_mark_tag(c, 1);
        nclt_prof_mark_tag(c, 2);
        k_occ_apply<<<tiles_r * tiles_c, TILE * TILE, 0, c->stream>>>(o->d_grid, o->g, n, hdr, wing, mixg, rays_g, ray_cap, tiles_c,
                                                                      o->d_counters);
        nclt_prof_mark_tag(c, 2);
        c->launches += 2;
    }
    CU_TRY(c, cudaGetLastError());
    return NCLT_OK;
}

static int check_occ(nclt_ctx* c, nclt_occ* o) {
    if (!c || !o) return nclt_fail(c, NCLT_ERR_ARG, "occ: null handle");
    if (o->device != c->device) return nclt_fail(c, NCLT_ERR_ARG, "occ: grid lives on another device");
    return NCLT_OK;
}

static int occ_integrate_depth_impl(nclt_ctx* c, nclt_occ* o, const void* depth, int is_u16, int compact, int F, int Hd,
                                    int Wd, const double* T, double fx, double fy, double cx, double cy,
                                    float* out_pts, int32_t* out_pts_n, int pts_cap) {
    int rc = check_occ(c, o);
    if (rc) return rc;
    if (F < 0 || Hd <= 0 || Wd <= 0 || (F > 0 && (!depth || !T))) return nclt_fail(c, NCLT_ERR_ARG, "occ_integrate_depth args");
    FrameIn in{};
    if (is_u16) in.depth16 = static_cast<const uint16_t*>(depth);
    else in.depth = static_cast<const float*>(depth);
    in.Hd = Hd; in.Wd = Wd; in.step = 4;          // step = 4 (tf_wall_clock_relay.py:876)
    in.compact = compact;
    in.fx = (float)fx; in.fy = (float)fy; in.cx = (float)cx; in.cy = (float)cy;
    in.T = T;
    return occ_integrate_dev(c, o, in, F, out_pts, out_pts_n, pts_cap);
}

extern "C" int nclt_occ_integrate_depth_dev(nclt_ctx* c, nclt_occ* o, const void* depth, int is_u16, int F, int Hd,
                                            int Wd, const double* T, double fx, double fy, double cx, double cy,
                                            float* out_pts, int32_t* out_pts_n, int pts_cap) {
    return occ_integrate_depth_impl(c, o, depth, is_u16, 0, F, Hd, Wd, T, fx, fy, cx, cy, out_pts, out_pts_n, pts_cap);
}

// Host depth frames -> device staging.  The path only ever reads every 4th row (and every 4th pixel of it), so only
// those rows cross PCIe: one strided 2-D copy (source pitch = 4 rows) moves a quarter of the bytes into a compact
// [F][Hd/4][Wd] buffer.  Needs Hd % 4 == 0 (so that the sampled rows of consecutive frames are equally spaced);
// otherwise the frames are copied whole.  Returns the staging pointer and whether it is compact.
static int stage_depth(nclt_ctx* c, const void* depth, int is_u16, int F, int Hd, int Wd, char* stage, int* compact) {
    const size_t esz = is_u16 ? 2 : 4;
    cudaError_t e;
    if (Hd % 4 == 0) {
        *compact = 1;
        e = cudaMemcpy2DAsync(stage, (size_t)Wd * esz, depth, (size_t)4 * Wd * esz, (size_t)Wd * esz, (size_t)F * (Hd / 4),
                              cudaMemcpyHostToDevice, c->stream);
    } else {
        *compact = 0;
        e = cudaMemcpyAsync(stage, depth, (size_t)F * Hd * Wd * esz, cudaMemcpyHostToDevice, c->stream);
    }
    return e == cudaSuccess ? NCLT_OK : nclt_fail(c, NCLT_ERR_CUDA, "depth H2D", e);
}
static size_t staged_depth_bytes(int is_u16, int F, int Hd, int Wd) {
    const size_t esz = is_u16 ? 2 : 4;
    return (Hd % 4 == 0) ? (size_t)F * (Hd / 4) * Wd * esz : (size_t)F * Hd * Wd * esz;
}

extern "C" int nclt_occ_integrate_points_dev(nclt_ctx* c, nclt_occ* o, const float* pts, const int32_t* n, int F,
                                             int Nmax, const double* T) {
    int rc = check_occ(c, o);
    if (rc) return rc;
    if (F < 0 || Nmax <= 0 || (F > 0 && (!pts || !n || !T))) return nclt_fail(c, NCLT_ERR_ARG, "occ_integrate_points args");
    FrameIn in{};
    in.pts = pts; in.pts_n = n; in.Nmax = Nmax; in.step = 1; in.T = T;
    return occ_integrate_dev(c, o, in, F, nullptr, nullptr, 0);
}

extern "C" int nclt_occ_integrate_depth(nclt_ctx* c, nclt_occ* o, const void* depth, int is_u16, int F, int Hd,
                                        int Wd, const double* T, double fx, double fy, double cx, double cy) {
    int rc = check_occ(c, o);
    if (rc) return rc;
    if (F < 0 || Hd <= 0 || Wd <= 0 || (F > 0 && (!depth || !T))) return nclt_fail(c, NCLT_ERR_ARG, "occ_integrate_depth args");
    if (F == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    // staging from the context scratch (no cudaMalloc / cudaFree per call: both synchronise the device)
    const size_t db = staged_depth_bytes(is_u16, F, Hd, Wd);
    ScratchScope scope(c);
    if ((rc = nclt_scratch_reserve(c, pad256(db) + pad256((size_t)F * 128) + 512))) return rc;
    Carver cv(c);
    char* stage = cv.take<char>(db);
    double* dT = cv.take<double>((size_t)F * 16);
    int compact = 0;
    rc = stage_depth(c, depth, is_u16, F, Hd, Wd, stage, &compact);
    if (rc == NCLT_OK) {
        cudaError_t e = cudaMemcpyAsync(dT, T, (size_t)F * 128, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) rc = nclt_fail(c, NCLT_ERR_CUDA, "T H2D", e);
    }
    if (rc == NCLT_OK)
        rc = occ_integrate_depth_impl(c, o, stage, is_u16, compact, F, Hd, Wd, dT, fx, fy, cx, cy, nullptr, nullptr, 0);
    cudaStreamSynchronize(c->stream);
    return rc;
}

extern "C" int nclt_occ_integrate_points(nclt_ctx* c, nclt_occ* o, const float* pts, const int32_t* n, int F,
                                         int Nmax, const double* T) {
    int rc = check_occ(c, o);
    if (rc) return rc;
    if (F < 0 || Nmax <= 0 || (F > 0 && (!pts || !n || !T))) return nclt_fail(c, NCLT_ERR_ARG, "occ_integrate_points args");
    if (F == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    const size_t pb = (size_t)F * Nmax * 12;
    ScratchScope scope(c);
    if ((rc = nclt_scratch_reserve(c, pad256(pb) + pad256((size_t)F * 4) + pad256((size_t)F * 128) + 768))) return rc;
    Carver cv(c);
    float* dp = cv.take<float>((size_t)F * Nmax * 3);
    int* dn = cv.take<int>((size_t)F);
    double* dT = cv.take<double>((size_t)F * 16);
    cudaError_t e = cudaMemcpyAsync(dp, pts, pb, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dn, n, (size_t)F * 4, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dT, T, (size_t)F * 128, cudaMemcpyHostToDevice, c->stream);
    rc = e == cudaSuccess ? nclt_occ_integrate_points_dev(c, o, dp, dn, F, Nmax, dT)
                          : nclt_fail(c, NCLT_ERR_CUDA, "points H2D", e);
    cudaStreamSynchronize(c->stream);
    return rc;
}

// relay depth_cb alone: the PointCloud2 payload (the topic is also consumed by Nav2's obstacle layer)
extern "C" int nclt_depth_to_points(nclt_ctx* c, const void* depth, int is_u16, int F, int Hd, int Wd, double fx,
                                    double fy, double cx, double cy, float* out_pts, int32_t* out_n, int pts_cap) {
    if (!c) return NCLT_ERR_ARG;
    if (F < 0 || Hd <= 0 || Wd <= 0 || pts_cap <= 0 || (F > 0 && (!depth || !out_pts || !out_n)))
        return nclt_fail(c, NCLT_ERR_ARG, "depth_to_points args");
    if (F == 0) return NCLT_OK;
    cudaSetDevice(c->device);
    // a throw-away 1x1 grid: stage A alone produces the clouds
    nclt_occ* o = nullptr;
    int rc = nclt_occ_create(c, 0.0, 0.0, 1.0, 1, 1, &o);
    if (rc) return rc;
    const size_t db = staged_depth_bytes(is_u16, F, Hd, Wd);
    const size_t ob = (size_t)F * pts_cap * 12;
    {
        ScratchScope scope(c);
        rc = nclt_scratch_reserve(c, pad256(db) + pad256((size_t)F * 128) + pad256(ob) + pad256((size_t)F * 4) + 1024);
        if (rc == NCLT_OK) {
            Carver cv(c);
            char* stage = cv.take<char>(db);
            double* dT = cv.take<double>((size_t)F * 16);
            float* dp = cv.take<float>((size_t)F * pts_cap * 3);
            int* dn = cv.take<int>((size_t)F);
            int compact = 0;
            rc = stage_depth(c, depth, is_u16, F, Hd, Wd, stage, &compact);
            if (rc == NCLT_OK && cudaMemsetAsync(dT, 0, (size_t)F * 128, c->stream) != cudaSuccess)
                rc = nclt_fail(c, NCLT_ERR_CUDA, "memset T", cudaGetLastError());
            if (rc == NCLT_OK)
                rc = occ_integrate_depth_impl(c, o, stage, is_u16, compact, F, Hd, Wd, dT, fx, fy, cx, cy, dp, dn, pts_cap);
            if (rc == NCLT_OK) {
                cudaError_t e = cudaMemcpyAsync(out_pts, dp, ob, cudaMemcpyDeviceToHost, c->stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(out_n, dn, (size_t)F * 4, cudaMemcpyDeviceToHost, c->stream);
                if (e != cudaSuccess) rc = nclt_fail(c, NCLT_ERR_CUDA, "points D2H", e);
            }
            cudaStreamSynchronize(c->stream);
        }
    }
    nclt_occ_destroy(c, o);
    return rc;
}

extern "C" int nclt_occ_read(nclt_ctx* c, nclt_occ* o, float* out_logodds, uint8_t* out_pgm, int32_t* out_units,
                             int64_t* out_counters) {
    int rc = check_occ(c, o);
    if (rc) return rc;
    cudaSetDevice(c->device);
    const size_t cells = (size_t)o->g.GW * o->g.GH;
    ScratchScope scope(c);
    if ((rc = nclt_scratch_reserve(c, pad256(cells * 4) + pad256(cells)))) return rc;
    Carver cv(c);
    float* d_lo = cv.take<float>(cells);
    unsigned char* d_img = cv.take<unsigned char>(cells);
    k_occ_render<<<(unsigned)((cells + 255) / 256), 256, 0, c->stream>>>(o->d_grid, o->g.GH, o->g.GW,
                                                                         out_pgm ? d_img : nullptr,
                                                                         out_logodds ? d_lo : nullptr);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    if (out_logodds) CU_TRY(c, cudaMemcpyAsync(out_logodds, d_lo, cells * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_pgm) CU_TRY(c, cudaMemcpyAsync(out_pgm, d_img, cells, cudaMemcpyDeviceToHost, c->stream));
    if (out_units) CU_TRY(c, cudaMemcpyAsync(out_units, o->d_grid, cells * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_counters)
        CU_TRY(c, cudaMemcpyAsync(out_counters, o->d_counters, 3 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}
