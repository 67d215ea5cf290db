/*
 * TEST INFRASTRUCTURE (CPU oracle) - not product code.
 *
 * Plain-C restatement of the teach-time map builder:
 *   - relay depth_cb: depth image -> camera_link point cloud
 *       simulation/isaac/scripts/common/tf_wall_clock_relay.py:868-887 (== v55:1020-1039)
 *   - mapper cb: transform, height filter, [::4], Bresenham free-space + endpoint log-odds
 *       simulation/isaac/scripts/common/teach_run_depth_mapper.py:125-195
 *   - save: thresholds -> {0,205,254}, flipud        teach_run_depth_mapper.py:208-216
 * Pinned by tests/test_oracle_occupancy.py against golden vectors produced by the reference
 * modules themselves (imported unmodified under ROS stubs, oracle/make_golden_ref.py).
 *
 * Arithmetic notes that decide bit-exactness:
 *   - back-projection is float32, in the order ((u - cx) / fx) * z (NumPy NEP 50: float32 array
 *     with Python-float scalars stays float32), no FMA;
 *   - the map transform runs in OpenBLAS dgemm: per coordinate an FMA chain over k = 0..3,
 *     fma(T3, 1, fma(T2, z, fma(T1, y, T0*x))) - verified bit-exact against numpy here;
 *   - world_to_pix truncates toward zero (Python int()), so points up to one cell below the
 *     origin land in row/col 0;
 *   - grid updates are float32 adds of float32(-0.4) / float32(1.4) with clamps at -5 / +5.
 * Build: gcc -O2 -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* relay depth_cb, 32FC1 input. Returns the number of points written (row-major order). */
int orc_depth_to_points(const float* depth, int H, int W, int step, double fx, double fy, double cx, double cy,
                        float* pts) {
    int n = 0;
    const float fxf = (float)fx, fyf = (float)fy, cxf = (float)cx, cyf = (float)cy;
    for (int v = 0; v < H; v += step)
        for (int u = 0; u < W; u += step) {
            float z = depth[(size_t)v * W + u];
            if (!((z > 0.3f) && (z < 10.0f) && isfinite(z))) continue;
            float px = ((float)u - cxf) / fxf * z;
            float py = ((float)v - cyf) / fyf * z;
            pts[3 * n] = z;
            pts[3 * n + 1] = -px;
            pts[3 * n + 2] = -py;
            n++;
        }
    return n;
}

/* 16UC1 input: depth = u16.astype(float32) / 1000.0 */
int orc_depth16_to_points(const uint16_t* depth, int H, int W, int step, double fx, double fy, double cx,
                          double cy, float* pts) {
    int n = 0;
    const float fxf = (float)fx, fyf = (float)fy, cxf = (float)cx, cyf = (float)cy;
    for (int v = 0; v < H; v += step)
        for (int u = 0; u < W; u += step) {
            float z = (float)depth[(size_t)v * W + u] / 1000.0f;
            if (!((z > 0.3f) && (z < 10.0f))) continue;
            float px = ((float)u - cxf) / fxf * z;
            float py = ((float)v - cyf) / fyf * z;
            pts[3 * n] = z;
            pts[3 * n + 1] = -px;
            pts[3 * n + 2] = -py;
            n++;
        }
    return n;
}

static int to_pix(double x, double origin, double res) { return (int)((x - origin) / res); }

/* Bresenham exactly as teach_run_depth_mapper.py:172-195; `upd(r, c, is_endpoint)` per cell */
#define BRESENHAM(r0, c0, r1, c1, FREE_STMT, OCC_STMT)        \
    do {                                                      \
        int dr = abs((r1) - (r0)), dc = abs((c1) - (c0));     \
        int sr = (r0) < (r1) ? 1 : -1, sc = (c0) < (c1) ? 1 : -1; \
        int err = dr - dc, r = (r0), c = (c0);                \
        for (;;) {                                            \
            if (!(r == (r1) && c == (c1))) { FREE_STMT; }     \
            else { OCC_STMT; break; }                         \
            int e2 = 2 * err;                                 \
            if (e2 > -dc) { err -= dc; r += sr; }             \
            if (e2 < dr) { err += dr; c += sc; }              \
        }                                                     \
    } while (0)

#include <stdlib.h>

/* the camera_link -> map transform of one coordinate, in dgemm's FMA order */
static double xform(const double* t, double x, double y, double z) {
    return fma(t[3], 1.0, fma(t[2], z, fma(t[1], y, t[0] * x)));
}

/* the (r1,c1) endpoints the mapper will ray-trace for one cloud, in order; returns their number
 * (= len(pts_map) after [::4], including out-of-grid points, flagged inb=0), or -1 if the frame
 * is dropped because the sensor cell is outside the grid, -2 if the cloud is empty, -3 if no
 * point survives the height filter. */
int orc_mapper_rays(const double* T, const float* pts, int n, int GH, int GW, double ox, double oy, double res,
                    int* r0c0, int* rays /* [n/4+1][2] */, unsigned char* inb) {
    if (n == 0) return -2;
    int kept = 0, m = 0;
    for (int i = 0; i < n; i++) {
        double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        double zm = xform(T + 8, x, y, z);
        if (!((zm > 0.2) && (zm < 2.0))) continue;
        if ((kept++ & 3) != 0) continue;
        double xm = xform(T, x, y, z), ym = xform(T + 4, x, y, z);
        int c1 = to_pix(xm, ox, res), r1 = to_pix(ym, oy, res);
        rays[2 * m] = r1;
        rays[2 * m + 1] = c1;
        inb[m] = (0 <= r1 && r1 < GH && 0 <= c1 && c1 < GW);
        m++;
    }
    if (kept == 0) return -3;
    int c0 = to_pix(T[3], ox, res), r0 = to_pix(T[7], oy, res);
    r0c0[0] = r0;
    r0c0[1] = c0;
    if (!(0 <= r0 && r0 < GH && 0 <= c0 && c0 < GW)) return -1;
    return m;
}

/* mapper cb on the float32 log-odds grid (reference semantics). counters: [frames_integrated,
 * total_points_integrated, frames_skipped_empty]. Returns rays traced. */
int orc_mapper_integrate(float* grid, int GH, int GW, double ox, double oy, double res, const double* T,
                         const float* pts, int n, long long* counters, int* rays_buf, unsigned char* inb_buf) {
    const float L_FREE = (float)-0.4, L_OCC = (float)1.4;
    int r0c0[2];
    int m = orc_mapper_rays(T, pts, n, GH, GW, ox, oy, res, r0c0, rays_buf, inb_buf);
    if (m == -2) { counters[2]++; return 0; }
    if (m < 0) return 0;
    int traced = 0;
    for (int i = 0; i < m; i++) {
        if (!inb_buf[i]) continue;
        int r1 = rays_buf[2 * i], c1 = rays_buf[2 * i + 1];
        BRESENHAM(r0c0[0], r0c0[1], r1, c1,
                  { float v = grid[(size_t)r * GW + c] + L_FREE; grid[(size_t)r * GW + c] = (v > -5.0f) ? v : -5.0f; },
                  { float v = grid[(size_t)r * GW + c] + L_OCC; grid[(size_t)r * GW + c] = (v < 5.0f) ? v : 5.0f; });
        traced++;
    }
    counters[0]++;
    counters[1] += m;
    return traced;
}

/* the exact-integer model (units of 0.2: free -2, occ +7, clamp +-25) used by the CUDA path */
int orc_mapper_integrate_int(int* grid, int GH, int GW, double ox, double oy, double res, const double* T,
                             const float* pts, int n, long long* counters, int* rays_buf, unsigned char* inb_buf) {
    int r0c0[2];
    int m = orc_mapper_rays(T, pts, n, GH, GW, ox, oy, res, r0c0, rays_buf, inb_buf);
    if (m == -2) { counters[2]++; return 0; }
    if (m < 0) return 0;
    int traced = 0;
    for (int i = 0; i < m; i++) {
        if (!inb_buf[i]) continue;
        int r1 = rays_buf[2 * i], c1 = rays_buf[2 * i + 1];
        BRESENHAM(r0c0[0], r0c0[1], r1, c1,
                  { int v = grid[(size_t)r * GW + c] - 2; grid[(size_t)r * GW + c] = v > -25 ? v : -25; },
                  { int v = grid[(size_t)r * GW + c] + 7; grid[(size_t)r * GW + c] = v < 25 ? v : 25; });
        traced++;
    }
    counters[0]++;
    counters[1] += m;
    return traced;
}

/* save(): 205 unknown, grid > log(0.65/0.35) -> 0, grid < log(0.25/0.75) -> 254, flipud */
void orc_mapper_render(const float* grid, int GH, int GW, unsigned char* img) {
    const double occ_th = log(0.65 / (1 - 0.65)), free_th = log(0.25 / (1 - 0.25));
    for (int r = 0; r < GH; r++)
        for (int c = 0; c < GW; c++) {
            double g = grid[(size_t)r * GW + c];
            unsigned char v = 205;
            if (g > occ_th) v = 0;
            if (g < free_th) v = 254;
            img[(size_t)(GH - 1 - r) * GW + c] = v;
        }
}

void orc_mapper_render_int(const int* grid, int GH, int GW, unsigned char* img) {
    for (int r = 0; r < GH; r++)
        for (int c = 0; c < GW; c++) {
            int g = grid[(size_t)r * GW + c];
            unsigned char v = 205;
            if (g >= 4) v = 0;      /* 0.8 > 0.619 >= 0.6 */
            if (g <= -6) v = 254;   /* -1.2 < -1.0986 <= -1.0 */
            img[(size_t)(GH - 1 - r) * GW + c] = v;
        }
}
