// Hit-count occupancy grid (datasets/rover/scripts/occupancy_astar.py:142-187 `build_occupancy`):
// classified world points -> floor / obstacle hit counts per X-Z cell (the reference's np.add.at) -> unknown / free /
// occupied.  Integer counts are commutative, so the scatter is plain atomics - no ordering problem as in the log-odds
// mapper.  Grid bounds come from the data (min / max of the classified points -+ 0.5 m), hence three small kernels:
// bounds reduction, scatter, classification.  HBM bound: 25 B per point in, 9 B per cell out.
#include "common.cuh"
#include "scratch.cuh"

#include <cmath>
#include <vector>

namespace {

struct Bounds { double xmin, xmax, zmin, zmax; long long n; };

__global__ void __launch_bounds__(256) k_hit_bounds(const double* __restrict__ pts, const int8_t* __restrict__ lab, long long N,
                                                    Bounds* __restrict__ part) {
    double xmin = INFINITY, xmax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
    long long n = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        if (lab[i] < 0) continue;
        const double x = pts[3 * i], z = pts[3 * i + 2];
        xmin = fmin(xmin, x); xmax = fmax(xmax, x);
        zmin = fmin(zmin, z); zmax = fmax(zmax, z);
        ++n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(0xFFFFFFFFu, xmin, o));
        xmax = fmax(xmax, __shfl_xor_sync(0xFFFFFFFFu, xmax, o));
        zmin = fmin(zmin, __shfl_xor_sync(0xFFFFFFFFu, zmin, o));
        zmax = fmax(zmax, __shfl_xor_sync(0xFFFFFFFFu, zmax, o));
        n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    }
    __shared__ Bounds s[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s[warp] = Bounds{xmin, xmax, zmin, zmax, n};
    __syncthreads();
    if (threadIdx.x == 0) {
        Bounds b = s[0];
        for (int w = 1; w < 8; ++w) {
            b.xmin = fmin(b.xmin, s[w].xmin); b.xmax = fmax(b.xmax, s[w].xmax);
            b.zmin = fmin(b.zmin, s[w].zmin); b.zmax = fmax(b.zmax, s[w].zmax);
            b.n += s[w].n;
        }
        part[blockIdx.x] = b;
    }
}

// cell of a coordinate exactly like `np.clip(((p - p_min) / res).astype(int), 0, n - 1)`
__device__ __forceinline__ int cell_of(double p, double pmin, double res, int n) {
    const double q = __ddiv_rn(__dsub_rn(p, pmin), res);
    long long c = (long long)q;                     // truncation toward zero
    if (!(q == q)) c = 0;
    return (int)(c < 0 ? 0 : (c > n - 1 ? n - 1 : c));
}

__global__ void __launch_bounds__(256) k_hit_scatter(const double* __restrict__ pts, const int8_t* __restrict__ lab, long long N,
                                                     double xmin, double zmin, double res, int nx, int nz,
                                                     int* __restrict__ floor_grid, int* __restrict__ obs_grid) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int l = lab[i];
        if (l != 0 && l != 1) continue;
        const int xi = cell_of(pts[3 * i], xmin, res, nx), zi = cell_of(pts[3 * i + 2], zmin, res, nz);
        atomicAdd((l == 0 ? floor_grid : obs_grid) + (size_t)zi * nx + xi, 1);
    }
}

__global__ void __launch_bounds__(256) k_hit_classify(const int* __restrict__ floor_grid, const int* __restrict__ obs_grid,
                                                      long long cells, int min_total, int min_obstacle, int8_t* __restrict__ occ) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const int f = floor_grid[i], o = obs_grid[i];
    occ[i] = (f + o >= min_total) ? (o >= min_obstacle ? 1 : 0) : -1;       // occupancy_astar.py:177-181
}

}  // namespace

extern "C" int nclt_hitcount_occupancy(nclt_ctx* c, const double* points, const int8_t* labels, long long N, double grid_res,
                                       int min_total, int min_obstacle, long long cell_cap, double* out_origin,
                                       int32_t* out_dims, int8_t* out_occ, int32_t* out_floor, int32_t* out_obs) {
    if (!c) return NCLT_ERR_ARG;
    if (!points || !labels || N <= 0 || !(grid_res > 0) || !out_origin || !out_dims || !out_occ || cell_cap <= 0)
        return nclt_fail(c, NCLT_ERR_ARG, "hitcount: bad arguments");
    cudaSetDevice(c->device);
    ScratchScope scope(c);
    const int nblk = (int)std::min<long long>((N + 255) / 256, (long long)c->sm_count * 16);
    int rc;
    if ((rc = nclt_scratch_reserve(c, pad256((size_t)N * 24) + pad256((size_t)N) + pad256((size_t)nblk * sizeof(Bounds)) +
                                          2 * pad256((size_t)cell_cap * 4) + pad256((size_t)cell_cap) + 2048)))
        return rc;
    Carver cv(c);
    double* d_pts = cv.take<double>((size_t)N * 3);
    int8_t* d_lab = cv.take<int8_t>((size_t)N);
    Bounds* d_part = cv.take<Bounds>((size_t)nblk);
    int* d_floor = cv.take<int>((size_t)cell_cap);
    int* d_obs = cv.take<int>((size_t)cell_cap);
    int8_t* d_occ = cv.take<int8_t>((size_t)cell_cap);
    CU_TRY(c, cudaMemcpyAsync(d_pts, points, (size_t)N * 24, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_lab, labels, (size_t)N, cudaMemcpyHostToDevice, c->stream));
    k_hit_bounds<<<nblk, 256, 0, c->stream>>>(d_pts, d_lab, N, d_part);
    c->launches++;
    std::vector<Bounds> part(nblk);
    CU_TRY(c, cudaMemcpyAsync(part.data(), d_part, (size_t)nblk * sizeof(Bounds), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    Bounds b{INFINITY, -INFINITY, INFINITY, -INFINITY, 0};
    for (const Bounds& p : part) {
        b.xmin = std::fmin(b.xmin, p.xmin); b.xmax = std::fmax(b.xmax, p.xmax);
        b.zmin = std::fmin(b.zmin, p.zmin); b.zmax = std::fmax(b.zmax, p.zmax);
        b.n += p.n;
    }
    if (b.n == 0) return nclt_fail(c, NCLT_ERR_ARG, "hitcount: no classified points (labels >= 0)");
    // occupancy_astar.py:155-161, same double expressions
    const double x_min = b.xmin - 0.5, x_max = b.xmax + 0.5, z_min = b.zmin - 0.5, z_max = b.zmax + 0.5;
    const long long nx = (long long)((x_max - x_min) / grid_res) + 1, nz = (long long)((z_max - z_min) / grid_res) + 1;
    out_origin[0] = x_min;
    out_origin[1] = z_min;
    out_dims[0] = (int32_t)nx;
    out_dims[1] = (int32_t)nz;
    if (nx <= 0 || nz <= 0 || nx > (1 << 30) || nz > (1 << 30) || nx * nz > cell_cap)
        return nclt_fail(c, NCLT_ERR_ARG, "hitcount: grid larger than cell_cap (out_dims holds the size needed)");
    const long long cells = nx * nz;
    CU_TRY(c, cudaMemsetAsync(d_floor, 0, (size_t)cells * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(d_obs, 0, (size_t)cells * 4, c->stream));
    k_hit_scatter<<<nblk, 256, 0, c->stream>>>(d_pts, d_lab, N, x_min, z_min, grid_res, (int)nx, (int)nz, d_floor, d_obs);
    k_hit_classify<<<(unsigned)((cells + 255) / 256), 256, 0, c->stream>>>(d_floor, d_obs, cells, min_total, min_obstacle, d_occ);
    c->launches += 2;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out_occ, d_occ, (size_t)cells, cudaMemcpyDeviceToHost, c->stream));
    if (out_floor) CU_TRY(c, cudaMemcpyAsync(out_floor, d_floor, (size_t)cells * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_obs) CU_TRY(c, cudaMemcpyAsync(out_obs, d_obs, (size_t)cells * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}
