// libnclt_b200_diag.so - micro-benchmarks and single-tile probes of the tensor-core building blocks (NOT part of the
// product library libnclt_b200.so; links against it for the context / scratch helpers).  include/nclt_b200_diag.h.
// Diagnostic: one tcgen05 tile (M = 128 queries x N train rows x K = 256 bits) end to end -
// validates the operand image layout, the shared-memory / instruction descriptors, the TMEM
// accumulator layout (f16 and f32) and the packed 16-bit TMEM load against a NumPy popcount
// (tests/test_tc_gpu.py), and times the building blocks of the tensor-core Hamming path
// (MMA issue rate, TMEM read rate, half2 min/max rate) for DESIGN.md.
#include "../common.cuh"
#include "../scratch.cuh"
#include "../tc_common.cuh"

#include <cuda_fp16.h>

namespace {

__global__ void __launch_bounds__(128) k_tc_probe(const uint32_t* __restrict__ a_bits /*[128][8]*/,
                                                  const uint32_t* __restrict__ b_bits /*[N][8]*/, int N, int c_fmt,
                                                  int ld_mode, uint32_t* out /*[128][N]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                 // 128 x 256 B
    uint8_t* sB = smem + 32768;         // N x 256 B
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // bits -> fp8 image (generic proxy writes)
    for (int i = tid; i < 128 * 16; i += 128) {
        int r = i >> 4, c = i & 15;
        uint32_t w = a_bits[r * 8 + (c >> 1)];
        uint32_t bits16 = (c & 1) ? (w >> 16) : (w & 0xFFFFu);
        *reinterpret_cast<uint4*>(sA + tc::image_offset(128, r, c * 16)) = tc::expand16(bits16);
    }
    for (int i = tid; i < N * 16; i += 128) {
        int r = i >> 4, c = i & 15;
        uint32_t w = b_bits[r * 8 + (c >> 1)];
        uint32_t bits16 = (c & 1) ? (w >> 16) : (w & 0xFFFFu);
        *reinterpret_cast<uint4*>(sB + tc::image_offset(N, r, c * 16)) = tc::expand16(bits16);
    }
    tc::fence_proxy_async();
    if (tid == 0) {
        tc::mbar_init(&s_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(&s_tmem, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;

    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f8(128, N, c_fmt);
        const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
        for (int k = 0; k < 8; ++k) {
            uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
            uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
            tc::mma_f8(tmem, da, db, idesc, k > 0 ? 1u : 0u);
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::tc_fence_after();

    // dump: lane quadrant of this warp, all columns (raw 32-bit cells, or packed pairs)
    const uint32_t row = warp * 32 + lane;
    const int ncols = ld_mode == 1 ? N / 2 : N;      // packed: one register per 2 columns
    for (int c0 = 0; c0 < N; c0 += (ld_mode == 1 ? 64 : 32)) {
        uint32_t r[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        if (ld_mode == 1) tc::tmem_ld32_pack16(taddr, r);
        else tc::tmem_ld32(taddr, r);
        tc::tmem_wait_ld();
        int o0 = ld_mode == 1 ? c0 / 2 : c0;
        for (int j = 0; j < 32; ++j)
            if (o0 + j < ncols) out[(size_t)row * ncols + o0 + j] = r[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- micro-benchmarks -------------------------------------------------------------------------
// mode 0: back-to-back MMAs (M=128, N, K=32) on resident smem tiles, `iters` x 8 k-steps
// mode 1: + every warp reads the accumulator tile back with packed 16-bit loads (no math)
// mode 2: + half2 top-2 max tracking on what was read
__global__ void __launch_bounds__(192) k_tc_bench(int N, int iters, int mode, float* sink, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 32768;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (32768 + N * 256) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x38383838u, 0xB838B838u, 0x3838B8B8u, 0xB8B83838u);
    tc::fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 128); }
        tc::mbar_fence_init();
    }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    long long t0 = clock64();
    if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_f8(128, N, 0);
            const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
            for (int it = 0; it < iters; ++it) {
                int buf = it & 1;
                if (mode > 0 && it >= 2) tc::mbar_wait(&s_empty[buf], ((it >> 1) - 1) & 1);
                tc::tc_fence_after();
                for (int k = 0; k < 8; ++k) {
                    uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
                    uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
                    tc::mma_f8(tmem + buf * 256, da, db, idesc, k > 0 ? 1u : 0u);
                }
                tc::mma_commit(&s_full[buf]);
            }
        }
    } else if (warp < 4) {
        __half2 m1 = __float2half2_rn(-1000.f), m2 = m1;
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
            int buf = it & 1;
            tc::mbar_wait(&s_full[buf], (it >> 1) & 1);
            tc::tc_fence_after();
            if (mode > 0) {
                for (int c0 = 0; c0 < N; c0 += 64) {
                    uint32_t r[32];
                    tc::tmem_ld32_pack16(tmem + buf * 256 + ((uint32_t)(warp * 32) << 16) + c0, r);
                    tc::tmem_wait_ld();
                    if (mode == 2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            __half2 v = *reinterpret_cast<__half2*>(&r[j]);
                            __half2 lo = __hmin2(m1, v);
                            m1 = __hmax2(m1, v);
                            m2 = __hmax2(m2, lo);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc ^= r[j];
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&s_empty[buf]);
            }
        }
        float2 f1 = __half22float2(m1), f2 = __half22float2(m2);
        if (f1.x + f1.y + f2.x + f2.y + (float)acc == 12345.f) sink[tid] = 1.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int nclt_tc_probe(nclt_ctx* c, const uint8_t* a_desc, const uint8_t* b_desc, int N, int c_fmt, int ld_mode,
                             uint32_t* out) {
    if (!c || !a_desc || !b_desc || !out || N < 16 || N > 256 || (N % 16)) return nclt_fail(c, NCLT_ERR_ARG, "tc_probe args");
    cudaSetDevice(c->device);
    ScratchScope scope(c);
    int rc;
    size_t out_elems = (size_t)128 * (ld_mode == 1 ? N / 2 : N);
    if ((rc = nclt_scratch_reserve(c, pad256(128 * 32) + pad256((size_t)N * 32) + pad256(out_elems * 4)))) return rc;
    Carver cv(c);
    uint32_t* da = cv.take<uint32_t>(128 * 8);
    uint32_t* db = cv.take<uint32_t>((size_t)N * 8);
    uint32_t* dout = cv.take<uint32_t>(out_elems);
    CU_TRY(c, cudaMemcpyAsync(da, a_desc, 128 * 32, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(db, b_desc, (size_t)N * 32, cudaMemcpyHostToDevice, c->stream));
    size_t smem = 32768 + (size_t)N * 256;
    CU_TRY(c, cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_probe<<<1, 128, smem, c->stream>>>(da, db, N, c_fmt, ld_mode, dout);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, dout, out_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}

// returns comparisons (pairs) per second over the whole GPU for the chosen mode
extern "C" double nclt_tc_bench(nclt_ctx* c, int N, int iters, int mode, double* cycles_per_tile) {
    if (!c || N < 16 || N > 256 || (N % 16) || iters < 4) return -1.0;
    cudaSetDevice(c->device);
    float* sink = nullptr;
    long long* cyc = nullptr;
    int blocks = c->sm_count;
    if (cudaMalloc(&sink, 192 * 4) != cudaSuccess || cudaMalloc(&cyc, blocks * 8) != cudaSuccess) return -1.0;
    size_t smem = 32768 + (size_t)N * 256;
    cudaFuncSetAttribute(k_tc_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_tc_bench<<<blocks, 192, smem, c->stream>>>(N, 8, mode, sink, cyc);
    cudaEventRecord(e0, c->stream);
    k_tc_bench<<<blocks, 192, smem, c->stream>>>(N, iters, mode, sink, cyc);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (cycles_per_tile) *cycles_per_tile = (double)h / iters;
    c->launches += 2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) { nclt_fail(c, NCLT_ERR_CUDA, "tc_bench", e); return -1.0; }
    return (double)blocks * iters * 128.0 * N / (ms * 1e-3);
}

// =========================================================================================
// kind::mxf4 probe (helpers in tc_common.cuh): one 128 x N x 256 tile, optionally with the accumulators
// pre-loaded with MX_MAGIC; the caller checks accumulator == 256 - 2 * Hamming (tools/mxf4_probe.py,
// tests/test_tc_gpu.py).
// =========================================================================================
namespace {

__global__ void __launch_bounds__(128) k_tc_probe_mxf4(const uint32_t* __restrict__ a_bits, const uint32_t* __restrict__ b_bits,
                                                       int N, int magic, uint32_t* out /*[128][N] f32 bits*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;              // 128 x 128 B
    uint8_t* sB = smem + 16384;      // N x 128 B
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * 16; i += 128) {
        int r = i >> 4, c = i & 15;           // c = 16-bit group -> 8 bytes at byte offset c*8
        uint32_t w = a_bits[r * 8 + (c >> 1)];
        *reinterpret_cast<uint2*>(sA + tc::image_offset4(128, r, c * 8)) = tc::expand16_fp4((c & 1) ? (w >> 16) : (w & 0xFFFFu));
    }
    for (int i = tid; i < N * 16; i += 128) {
        int r = i >> 4, c = i & 15;
        uint32_t w = b_bits[r * 8 + (c >> 1)];
        *reinterpret_cast<uint2*>(sB + tc::image_offset4(N, r, c * 8)) = tc::expand16_fp4((c & 1) ? (w >> 16) : (w & 0xFFFFu));
    }
    uint8_t* sAb = sB + (size_t)N * 128;      // bias slabs (magic == 2): 128 x 32 B, N x 32 B
    uint8_t* sBb = sAb + 4096;
    if (magic == 2) {
        tc::mx_fill_bias_slab(sAb, 128, false, tid, 128);
        tc::mx_fill_bias_slab(sBb, N, true, tid, 128);
    }
    tc::fence_proxy_async();
    if (tid == 0) { tc::mbar_init(&s_bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    tc::tmem_st32_const(tmem + lane_base + 480u, 0x7F7F7F7Fu);    // scale factors = 1.0 everywhere
    if (magic == 2) tc::tmem_st16_const(tmem + lane_base + 496u, tc::MX_BIAS_SFA);   // [480,496) = 1.0 (A and B), [496,512) = 2^14
    if (magic == 1)
        for (int c0 = 0; c0 < N; c0 += 16) tc::tmem_st16_const(tmem + lane_base + (uint32_t)c0, tc::MX_MAGIC);
    tc::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_mxf4(128, N);
        const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
        if (magic == 2)     // the bias step: constant operands, A scale 2^14
            tc::mma_mxf4(tmem, tc::smem_desc(tc::smem_u32(sAb), lboA, 128), tc::smem_desc(tc::smem_u32(sBb), lboB, 128), idesc, 0u,
                         tmem + 496u, tmem + 480u);
        for (int k = 0; k < 4; ++k) {
            uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
            uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
            tc::mma_mxf4(tmem, da, db, idesc, (k > 0 || magic) ? 1u : 0u, tmem + 480u, magic == 2 ? tmem + 480u : tmem + 496u);
        }
        tc::mma_commit(&s_bar);
    }
    tc::mbar_wait(&s_bar, 0);
    tc::tc_fence_after();
    const uint32_t row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem + lane_base + (uint32_t)c0, r);
        tc::tmem_wait_ld();
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) out[(size_t)row * N + c0 + j] = r[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Rates of the mxf4 path on resident tiles (N = 240):
//   mode 0: MMAs only            mode 1: + f32 loads and fmaxf
//   mode 2: + the real epilogue: pre-loaded accumulators, packed 16-bit loads, half2 max, re-arm stores
// NWQ epilogue warps per TMEM lane quadrant (2: 120 columns each, 3: 80 columns each).
template <int NWQ>
__global__ void __launch_bounds__(32 * (4 * NWQ + 2)) k_tc_bench_mxf4(int N, int iters, int mode, float* sink, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + N * 128) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x2A2A2A2Au, 0xA2A2A2A2u, 0x22AA22AAu, 0xAA22AA22u);
    uint8_t* sAb = sB + (size_t)N * 128;
    uint8_t* sBb = sAb + 4096;
    const bool bias = mode == 8 || mode == 9;
    if (bias) {
        tc::mx_fill_bias_slab(sAb, 128, false, tid, blockDim.x);
        tc::mx_fill_bias_slab(sBb, N, true, tid, blockDim.x);
        mode = mode == 8 ? 4 : 0;
    }
    tc::fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 4 * NWQ); }
        tc::mbar_fence_init();
    }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 4) {
        const uint32_t lb = (uint32_t)(warp * 32) << 16;
        tc::tmem_st32_const(tmem + lb + 480u, 0x7F7F7F7Fu);
        if (bias) tc::tmem_st16_const(tmem + lb + 496u, tc::MX_BIAS_SFA);
        if (mode == 2)
            for (int c0 = 0; c0 < 480; c0 += 16) tc::tmem_st16_const(tmem + lb + (uint32_t)c0, tc::MX_MAGIC);
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = clock64();
    if (warp == 4 * NWQ + 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_mxf4(128, N);
            const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
            for (int it = 0; it < iters; ++it) {
                int buf = it & 1;
                if (mode > 0 && it >= 2) tc::mbar_wait(&s_empty[buf], ((it >> 1) - 1) & 1);
                tc::tc_fence_after();
                if (bias)
                    tc::mma_mxf4(tmem + buf * 240, tc::smem_desc(tc::smem_u32(sAb), lboA, 128),
                                 tc::smem_desc(tc::smem_u32(sBb), lboB, 128), idesc, 0u, tmem + 496u, tmem + 480u);
                for (int k = 0; k < 4; ++k) {
                    uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
                    uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
                    tc::mma_mxf4(tmem + buf * 240, da, db, idesc, (k > 0 || mode == 2 || bias) ? 1u : 0u, tmem + 480u,
                                 bias ? tmem + 480u : tmem + 496u);
                }
                tc::mma_commit(&s_full[buf]);
            }
        }
    } else if (warp < 4 * NWQ) {
        const int quad = warp & 3, part = warp >> 2;
        const uint32_t lb = (uint32_t)(quad * 32) << 16;
        float m0 = -1e30f, m1 = -1e30f;
        __half2 h0 = __float2half2_rn(0.f), h1 = h0;
        for (int it = 0; it < iters; ++it) {
            int buf = it & 1;
            tc::mbar_wait(&s_full[buf], (it >> 1) & 1);
            tc::tc_fence_after();
            if (mode == 1) {
                const int c_lo = part * (N / NWQ);
                for (int c0 = 0; c0 + 32 <= N / NWQ; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(tmem + buf * 240 + lb + c_lo + c0, r);
                    tc::tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        m0 = fmaxf(m0, __uint_as_float(r[j]));
                        m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
                    }
                }
            } else if (mode == 2) {
                if (NWQ == 2) {          // 120 columns: 64 + 32 + 16 + 8
                    const uint32_t ta = tmem + buf * 240 + lb + part * 120;
                    uint32_t r0[32], r1[16], r2[8], r3[4];
                    tc::tmem_ld32_pack16(ta, r0);
                    tc::tmem_ld16_pack16(ta + 64, r1);
                    tc::tmem_ld8_pack16(ta + 96, r2);
                    tc::tmem_ld4_pack16(ta + 112, r3);
                    tc::tmem_wait_ld();
                    tc::tmem_st64_const(ta, tc::MX_MAGIC);
                    tc::tmem_st32_const(ta + 64, tc::MX_MAGIC);
                    tc::tmem_st16_const(ta + 96, tc::MX_MAGIC);
                    tc::tmem_st8_const(ta + 112, tc::MX_MAGIC);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r0[j])); }
#pragma unroll
                    for (int j = 0; j < 16; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r1[j])); }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r2[j])); }
#pragma unroll
                    for (int j = 0; j < 4; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r3[j])); }
                    tc::tmem_wait_st();
                } else {                 // 80 columns: 64 + 16
                    const uint32_t ta = tmem + buf * 240 + lb + part * 80;
                    uint32_t r0[32], r2[8];
                    tc::tmem_ld32_pack16(ta, r0);
                    tc::tmem_ld8_pack16(ta + 64, r2);
                    tc::tmem_wait_ld();
                    tc::tmem_st64_const(ta, tc::MX_MAGIC);
                    tc::tmem_st16_const(ta + 64, tc::MX_MAGIC);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r0[j])); }
#pragma unroll
                    for (int j = 0; j < 8; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r2[j])); }
                    tc::tmem_wait_st();
                }
            }
            else if (mode == 4) {    // packed loads + half2 max only (no re-arm stores)
                const uint32_t ta = tmem + buf * 240 + lb + part * 120;
                uint32_t r0[32], r1[16], r2[8], r3[4];
                tc::tmem_ld32_pack16(ta, r0);
                tc::tmem_ld16_pack16(ta + 64, r1);
                tc::tmem_ld8_pack16(ta + 96, r2);
                tc::tmem_ld4_pack16(ta + 112, r3);
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r0[j])); }
#pragma unroll
                for (int j = 0; j < 16; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r1[j])); }
#pragma unroll
                for (int j = 0; j < 8; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r2[j])); }
#pragma unroll
                for (int j = 0; j < 4; ++j) { __half2& h = (j & 1) ? h1 : h0; h = __hmax2(h, *reinterpret_cast<__half2*>(&r3[j])); }
            } else if (mode == 5) {  // f32 cells, three-input max, all 120 columns, two batches in flight
                const uint32_t ta = tmem + buf * 240 + lb + part * 120;
                uint32_t a[32], b[32], c2[32], d[16], e[8];
                tc::tmem_ld32(ta, a);
                tc::tmem_ld32(ta + 32, b);
                tc::tmem_wait_ld();
                tc::tmem_ld32(ta + 64, c2);
                tc::tmem_ld16(ta + 96, d);
                tc::tmem_ld8(ta + 112, e);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    m0 = tc::fmax3(m0, __uint_as_float(a[j]), __uint_as_float(a[j + 1]));
                    m1 = tc::fmax3(m1, __uint_as_float(b[j]), __uint_as_float(b[j + 1]));
                }
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 2) m0 = tc::fmax3(m0, __uint_as_float(c2[j]), __uint_as_float(c2[j + 1]));
#pragma unroll
                for (int j = 0; j < 16; j += 2) m1 = tc::fmax3(m1, __uint_as_float(d[j]), __uint_as_float(d[j + 1]));
#pragma unroll
                for (int j = 0; j < 8; j += 2) m1 = tc::fmax3(m1, __uint_as_float(e[j]), __uint_as_float(e[j + 1]));
            }
            else if (mode == 6) {    // f32 cells, NWQ = 3 (80 columns) or 4 (60 columns)
                const uint32_t ta = tmem + buf * 240 + lb + part * (240 / NWQ);
                uint32_t a[32], b[32], d[16], e[8], f[4];
                tc::tmem_ld32(ta, a);
                if (NWQ == 3) {
                    tc::tmem_ld32(ta + 32, b);
                    tc::tmem_ld16(ta + 64, d);
                } else {
                    tc::tmem_ld16(ta + 32, d);
                    tc::tmem_ld8(ta + 48, e);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(f[0]), "=r"(f[1]), "=r"(f[2]), "=r"(f[3]) : "r"(ta + 56) : "memory");
                }
                tc::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 2) m0 = tc::fmax3(m0, __uint_as_float(a[j]), __uint_as_float(a[j + 1]));
#pragma unroll
                for (int j = 0; j < 16; j += 2) m1 = tc::fmax3(m1, __uint_as_float(d[j]), __uint_as_float(d[j + 1]));
                if (NWQ == 3) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) m1 = tc::fmax3(m1, __uint_as_float(b[j]), __uint_as_float(b[j + 1]));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j += 2) m0 = tc::fmax3(m0, __uint_as_float(e[j]), __uint_as_float(e[j + 1]));
                    m1 = tc::fmax3(m1, __uint_as_float(f[0]), __uint_as_float(f[1]));
                    m1 = tc::fmax3(m1, __uint_as_float(f[2]), __uint_as_float(f[3]));
                }
            }
            if (mode > 0) {
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&s_empty[buf]);
            }
        }
        float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        if (m0 + m1 + f0.x + f0.y + f1.x + f1.y == 12345.f) sink[tid] = 1.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

extern "C" int nclt_tc_probe_mxf4(nclt_ctx* c, const uint8_t* a_desc, const uint8_t* b_desc, int N, int magic, uint32_t* out) {
    if (!c || !a_desc || !b_desc || !out || N < 16 || N > 240 || (N % 16)) return nclt_fail(c, NCLT_ERR_ARG, "tc_probe_mxf4 args");
    cudaSetDevice(c->device);
    ScratchScope scope(c);
    int rc;
    size_t out_elems = (size_t)128 * N;
    if ((rc = nclt_scratch_reserve(c, pad256(128 * 32) + pad256((size_t)N * 32) + pad256(out_elems * 4)))) return rc;
    Carver cv(c);
    uint32_t* da = cv.take<uint32_t>(128 * 8);
    uint32_t* db = cv.take<uint32_t>((size_t)N * 8);
    uint32_t* dout = cv.take<uint32_t>(out_elems);
    CU_TRY(c, cudaMemcpyAsync(da, a_desc, 128 * 32, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(db, b_desc, (size_t)N * 32, cudaMemcpyHostToDevice, c->stream));
    size_t smem = 16384 + (size_t)N * 128 + 4096 + (size_t)N * 32;
    CU_TRY(c, cudaFuncSetAttribute(k_tc_probe_mxf4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_probe_mxf4<<<1, 128, smem, c->stream>>>(da, db, N, magic, dout);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, dout, out_elems * 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return NCLT_OK;
}

// mode 0/1/2 as above; mode 3 = mode 2 with three epilogue warps per lane quadrant
extern "C" double nclt_tc_bench_mxf4(nclt_ctx* c, int N, int iters, int mode, double* cycles_per_tile) {
    if (!c || N < 16 || N > 240 || (N % 16) || iters < 4 || (mode >= 2 && N != 240)) return -1.0;
    cudaSetDevice(c->device);
    float* sink = nullptr;
    long long* cyc = nullptr;
    int blocks = c->sm_count;
    if (cudaMalloc(&sink, 512 * 4) != cudaSuccess || cudaMalloc(&cyc, blocks * 8) != cudaSuccess) return -1.0;
    size_t smem = 16384 + (size_t)N * 128 + 4096 + (size_t)N * 32;
    cudaFuncSetAttribute(k_tc_bench_mxf4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tc_bench_mxf4<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tc_bench_mxf4<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto go = [&](int n_it) {
        if (mode == 3) k_tc_bench_mxf4<3><<<blocks, 448, smem, c->stream>>>(N, n_it, 2, sink, cyc);
        else if (mode == 6) k_tc_bench_mxf4<3><<<blocks, 448, smem, c->stream>>>(N, n_it, 6, sink, cyc);
        else if (mode == 7) k_tc_bench_mxf4<4><<<blocks, 576, smem, c->stream>>>(N, n_it, 6, sink, cyc);
        else k_tc_bench_mxf4<2><<<blocks, 320, smem, c->stream>>>(N, n_it, mode, sink, cyc);
    };
    go(8);
    cudaEventRecord(e0, c->stream);
    go(iters);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (cycles_per_tile) *cycles_per_tile = (double)h / iters;
    c->launches += 2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) { nclt_fail(c, NCLT_ERR_CUDA, "tc_bench_mxf4", e); return -1.0; }
    return (double)blocks * iters * 128.0 * N / (ms * 1e-3);
}

// ---- the packed fp4 design in isolation: five MMAs per 128 x 240 tile (bias step first), two epilogue warp sets that
// take alternate tiles (set s owns TMEM buffer s), every warp reads ALL 240 columns of its lane quadrant with packed
// 16-bit loads in ONE batch (120 registers, a single tcgen05.wait::ld), hands the buffer back, then runs the half2 maxima.
namespace {
__global__ void __launch_bounds__(320, 1) k_tc_bench_mxp(int iters, int variant, float* sink, long long* cycles) {
    constexpr int N = 240;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    uint8_t* sAb = sB + (size_t)N * 128;
    uint8_t* sBb = sAb + 4096;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + N * 128) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x2A2A2A2Au, 0xA2A2A2A2u, 0x22AA22AAu, 0xAA22AA22u);
    tc::mx_fill_bias_slab(sAb, 128, false, tid, blockDim.x);
    tc::mx_fill_bias_slab(sBb, N, true, tid, blockDim.x);
    tc::fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 4); }
        tc::mbar_fence_init();
    }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 4) {
        const uint32_t lb = (uint32_t)(warp * 32) << 16;
        tc::tmem_st16_const(tmem + lb + 480u, 0x7F7F7F7Fu);
        tc::tmem_st16_const(tmem + lb + 496u, tc::MX_BIAS_SFA);
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = clock64();
    if (warp == 9) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_mxf4(128, N);
            const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
            for (int it = 0; it < iters; ++it) {
                const int buf = it & 1;
                if (it >= 2) tc::mbar_wait(&s_empty[buf], ((it >> 1) - 1) & 1);
                tc::tc_fence_after();
                tc::mma_mxf4(tmem + buf * 240, tc::smem_desc(tc::smem_u32(sAb), lboA, 128), tc::smem_desc(tc::smem_u32(sBb), lboB, 128),
                             idesc, 0u, tmem + 496u, tmem + 480u);
                for (int k = 0; k < 4; ++k) {
                    uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
                    uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
                    tc::mma_mxf4(tmem + buf * 240, da, db, idesc, 1u, tmem + 480u, tmem + 480u);
                }
                tc::mma_commit(&s_full[buf]);
            }
        }
    } else if (warp < 8) {
        const int quad = warp & 3, set = warp >> 2;
        const uint32_t lb = (uint32_t)(quad * 32) << 16;
        __half2 h[4];
        for (int i = 0; i < 4; ++i) h[i] = __float2half2_rn(0.f);
        for (int it = set; it < iters; it += 2) {
            const int buf = set;
            tc::mbar_wait(&s_full[buf], (it >> 1) & 1);
            tc::tc_fence_after();
            const uint32_t ta = tmem + buf * 240 + lb;
            uint32_t a[32], b[32], c2[32], d[16], e[8];
            tc::tmem_ld32_pack16(ta, a);
            tc::tmem_ld32_pack16(ta + 64, b);
            if (variant == 1) tc::tmem_wait_ld();          // variant 1: two batches (what 32-bit cells force)
            tc::tmem_ld32_pack16(ta + 128, c2);
            tc::tmem_ld16_pack16(ta + 192, d);
            tc::tmem_ld8_pack16(ta + 224, e);
            tc::tmem_wait_ld();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_empty[buf]);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                h[j & 3] = __hmax2(h[j & 3], *reinterpret_cast<__half2*>(&a[j]));
                h[j & 3] = __hmax2(h[j & 3], *reinterpret_cast<__half2*>(&b[j]));
                h[j & 3] = __hmax2(h[j & 3], *reinterpret_cast<__half2*>(&c2[j]));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j & 3] = __hmax2(h[j & 3], *reinterpret_cast<__half2*>(&d[j]));
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j & 3] = __hmax2(h[j & 3], *reinterpret_cast<__half2*>(&e[j]));
        }
        float acc = 0;
        for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); acc += f.x + f.y; }
        if (acc == 12345.f) sink[tid] = 1.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" double nclt_tc_bench_mxp(nclt_ctx* c, int iters, int variant, double* cycles_per_tile) {
    if (!c || iters < 4) return -1.0;
    cudaSetDevice(c->device);
    float* sink = nullptr;
    long long* cyc = nullptr;
    const int blocks = c->sm_count;
    if (cudaMalloc(&sink, 512 * 4) != cudaSuccess || cudaMalloc(&cyc, blocks * 8) != cudaSuccess) return -1.0;
    const size_t smem = 16384 + (size_t)240 * 128 + 4096 + (size_t)240 * 32;
    cudaFuncSetAttribute(k_tc_bench_mxp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_tc_bench_mxp<<<blocks, 320, smem, c->stream>>>(8, variant, sink, cyc);
    cudaEventRecord(e0, c->stream);
    k_tc_bench_mxp<<<blocks, 320, smem, c->stream>>>(iters, variant, sink, cyc);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (cycles_per_tile) *cycles_per_tile = (double)h / iters;
    c->launches += 2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) { nclt_fail(c, NCLT_ERR_CUDA, "tc_bench_mxp", e); return -1.0; }
    return (double)blocks * iters * 128.0 * 240 / (ms * 1e-3);
}

// ---- TMEM read-port microbenchmark: NW warps loop tcgen05.ld (32 columns each, `batch` loads per wait) --------
namespace {
__global__ void k_tmem_bw(int iters, int batch, int with_max, float* sink, long long* cycles) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lb = (uint32_t)((warp & 3) * 32) << 16;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t a[32], b[32], c[32], d[32];
        const uint32_t ta = tmem + lb + (uint32_t)((it * 32 * batch + (warp >> 2) * 128) & 255);
        tc::tmem_ld32(ta, a);
        if (batch >= 2) tc::tmem_ld32(ta + 32, b);
        if (batch >= 3) tc::tmem_ld32(ta + 64, c);
        if (batch >= 4) tc::tmem_ld32(ta + 96, d);
        tc::tmem_wait_ld();
        if (with_max) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                m0 = tc::fmax3(m0, __uint_as_float(a[j]), __uint_as_float(a[j + 1]));
                if (batch >= 2) m1 = tc::fmax3(m1, __uint_as_float(b[j]), __uint_as_float(b[j + 1]));
                if (batch >= 3) m2 = tc::fmax3(m2, __uint_as_float(c[j]), __uint_as_float(c[j + 1]));
                if (batch >= 4) m3 = tc::fmax3(m3, __uint_as_float(d[j]), __uint_as_float(d[j + 1]));
            }
        } else {
            acc ^= a[0] ^ a[31];
            if (batch >= 2) acc ^= b[0] ^ b[31];
            if (batch >= 3) acc ^= c[0] ^ c[31];
            if (batch >= 4) acc ^= d[0] ^ d[31];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (m0 + m1 + m2 + m3 + (float)acc == 12345.f) sink[tid] = 1.f;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace

// bytes per clock per SM that `warps` warps pull out of TMEM with 32-column loads, `batch` loads per wait
extern "C" double nclt_tmem_bw(nclt_ctx* c, int warps, int batch, int with_max) {
    if (!c || warps < 1 || warps > 32 || batch < 1 || batch > 4) return -1.0;
    cudaSetDevice(c->device);
    float* sink = nullptr;
    long long* cyc = nullptr;
    if (cudaMalloc(&sink, 1024 * 4) != cudaSuccess || cudaMalloc(&cyc, c->sm_count * 8) != cudaSuccess) return -1.0;
    const int iters = 4000;
    k_tmem_bw<<<c->sm_count, warps * 32, 0, c->stream>>>(64, batch, with_max, sink, cyc);
    k_tmem_bw<<<c->sm_count, warps * 32, 0, c->stream>>>(iters, batch, with_max, sink, cyc);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(sink);
    cudaFree(cyc);
    c->launches += 2;
    if (e != cudaSuccess || h <= 0) return -1.0;
    return (double)warps * iters * batch * 32.0 * 128.0 / (double)h;
}

// ---- mxf4 epilogue design probe: 16 epilogue warps in two sets that take alternate tiles (set s owns accumulator
// buffer s), accumulators pre-armed with MX_MAGIC, packed 16-bit loads, re-arm stores, three-input u16x2 maxima ----
namespace {
template <int variant>
__global__ void __launch_bounds__(576) k_tc_bench_mx16(int iters, float* sink, long long* cycles) {
    constexpr int N = 240;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + N * 128) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x2A2A2A2Au, 0xA2A2A2A2u, 0x22AA22AAu, 0xAA22AA22u);
    tc::fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_empty[s], 8); }
        tc::mbar_fence_init();
    }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 4) {
        const uint32_t lb = (uint32_t)(warp * 32) << 16;
        tc::tmem_st32_const(tmem + lb + 480u, 0x7F7F7F7Fu);
        for (int c0 = 0; c0 < 480; c0 += 16) tc::tmem_st16_const(tmem + lb + (uint32_t)c0, tc::MX_MAGIC);
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = clock64();
    if (warp == 17) {
        if (lane == 0) {
            const uint32_t idesc = tc::idesc_mxf4(128, N);
            const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
            for (int it = 0; it < iters; ++it) {
                int buf = it & 1;
                if (it >= 2) tc::mbar_wait(&s_empty[buf], ((it >> 1) - 1) & 1);
                tc::tc_fence_after();
                for (int k = 0; k < 4; ++k) {
                    uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
                    uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
                    tc::mma_mxf4(tmem + buf * 240, da, db, idesc, 1u, tmem + 480u, tmem + 496u);
                }
                tc::mma_commit(&s_full[buf]);
            }
        }
    } else if (warp < 16) {
        const int quad = warp & 3, part = (warp >> 2) & 1, set = warp >> 3;
        const uint32_t ta = tmem + set * 240 + ((uint32_t)(quad * 32) << 16) + part * 120;
        uint32_t h0 = 0, h1 = 0;
        for (int it = set; it < iters; it += 2) {
            tc::mbar_wait(&s_full[set], (it >> 1) & 1);
            tc::tc_fence_after();
            uint32_t r0[32], r1[16], r2[8], r3[4];
            tc::tmem_ld32_pack16(ta, r0);
            tc::tmem_ld16_pack16(ta + 64, r1);
            tc::tmem_ld8_pack16(ta + 96, r2);
            tc::tmem_ld4_pack16(ta + 112, r3);
            tc::tmem_wait_ld();
            // STTM takes a block of N consecutive registers: small stores keep the constant block small
#pragma unroll
            for (int c0 = 0; c0 < 120; c0 += 8) tc::tmem_st8_const(ta + c0, tc::MX_MAGIC);
            tc::tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&s_empty[set]);
            if (variant == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) { uint32_t& h = (j & 2) ? h1 : h0; h = __vimax3_u16x2(h, r0[j], r0[j + 1]); }
#pragma unroll
                for (int j = 0; j < 16; j += 2) { uint32_t& h = (j & 2) ? h1 : h0; h = __vimax3_u16x2(h, r1[j], r1[j + 1]); }
#pragma unroll
                for (int j = 0; j < 8; j += 2) { uint32_t& h = (j & 2) ? h1 : h0; h = __vimax3_u16x2(h, r2[j], r2[j + 1]); }
                h0 = __vimax3_u16x2(h0, r3[0], r3[1]);
                h1 = __vimax3_u16x2(h1, r3[2], r3[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) { uint32_t& h = (j & 1) ? h1 : h0; h = __vmaxu2(h, r0[j]); }
#pragma unroll
                for (int j = 0; j < 16; ++j) { uint32_t& h = (j & 1) ? h1 : h0; h = __vmaxu2(h, r1[j]); }
#pragma unroll
                for (int j = 0; j < 8; ++j) { uint32_t& h = (j & 1) ? h1 : h0; h = __vmaxu2(h, r2[j]); }
#pragma unroll
                for (int j = 0; j < 4; ++j) { uint32_t& h = (j & 1) ? h1 : h0; h = __vmaxu2(h, r3[j]); }
            }
        }
        if (h0 + h1 == 12345u) sink[tid] = 1.f;
    }
    tc::tc_fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" double nclt_tc_bench_mx16(nclt_ctx* c, int iters, int variant, double* cycles_per_tile) {
    if (!c || iters < 4) return -1.0;
    cudaSetDevice(c->device);
    float* sink = nullptr;
    long long* cyc = nullptr;
    int blocks = c->sm_count;
    if (cudaMalloc(&sink, 1024 * 4) != cudaSuccess || cudaMalloc(&cyc, blocks * 8) != cudaSuccess) return -1.0;
    size_t smem = 16384 + (size_t)240 * 128;
    cudaFuncSetAttribute(k_tc_bench_mx16<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_tc_bench_mx16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (variant == 0) k_tc_bench_mx16<0><<<blocks, 576, smem, c->stream>>>(8, sink, cyc);
    else k_tc_bench_mx16<1><<<blocks, 576, smem, c->stream>>>(8, sink, cyc);
    cudaEventRecord(e0, c->stream);
    if (variant == 0) k_tc_bench_mx16<0><<<blocks, 576, smem, c->stream>>>(iters, sink, cyc);
    else k_tc_bench_mx16<1><<<blocks, 576, smem, c->stream>>>(iters, sink, cyc);
    cudaEventRecord(e1, c->stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (cycles_per_tile) *cycles_per_tile = (double)h / iters;
    c->launches += 2;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) { nclt_fail(c, NCLT_ERR_CUDA, "tc_bench_mx16", e); return -1.0; }
    return (double)blocks * iters * 128.0 * 240.0 / (ms * 1e-3);
}

// ---- do tcgen05.mma streams issued by TWO different warps pipeline like one stream? -------------------------
// variant 0: one thread issues every tile (buffers alternate); variant 1: warp 0 issues the even tiles into buffer 0,
// warp 1 the odd tiles into buffer 1; each issuer waits for its own previous tile (commit -> mbarrier) first.
namespace {
__global__ void __launch_bounds__(128) k_tc_bench_two_issuers(int iters, int variant, long long* cycles) {
    constexpr int N = 240;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_done[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (16384 + N * 128) / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x2A2A2A2Au, 0xA2A2A2A2u, 0x22AA22AAu, 0xAA22AA22u);
    tc::fence_proxy_async();
    if (tid == 0) {
        tc::mbar_init(&s_done[0], 1);
        tc::mbar_init(&s_done[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) { tc::tmem_alloc(&s_tmem, 512); tc::tmem_relinquish(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 4) {
        tc::tmem_st32_const(tmem + ((uint32_t)(warp * 32) << 16) + 480u, 0x7F7F7F7Fu);
        tc::tmem_wait_st();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    long long t0 = clock64();
    const uint32_t idesc = tc::idesc_mxf4(128, N);
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
    auto issue = [&](int buf) {
        for (int k = 0; k < 4; ++k) {
            uint64_t da = tc::smem_desc(tc::smem_u32(sA) + k * 2 * lboA, lboA, 128);
            uint64_t db = tc::smem_desc(tc::smem_u32(sB) + k * 2 * lboB, lboB, 128);
            tc::mma_mxf4(tmem + buf * 240, da, db, idesc, k > 0 ? 1u : 0u, tmem + 480u, tmem + 496u);
        }
        tc::mma_commit(&s_done[buf]);
    };
    if (variant == 0) {
        if (tid == 0)
            for (int it = 0; it < iters; ++it) {
                const int buf = it & 1;
                if (it >= 2) tc::mbar_wait(&s_done[buf], ((it >> 1) - 1) & 1);
                tc::tc_fence_after();
                issue(buf);
            }
    } else if (warp < 2 && lane == 0) {
        for (int it = warp; it < iters; it += 2) {
            if (it >= 2) tc::mbar_wait(&s_done[warp], ((it >> 1) - 1) & 1);
            tc::tc_fence_after();
            issue(warp);
        }
    }
    if (tid == 0) {                          // wait for the last two tiles
        tc::mbar_wait(&s_done[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
        tc::mbar_wait(&s_done[(iters - 2) & 1], ((iters - 2) >> 1) & 1);
    }
    tc::tc_fence_before();
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace

extern "C" double nclt_tc_bench_two_issuers(nclt_ctx* c, int iters, int variant) {
    if (!c || iters < 4 || (iters & 1)) return -1.0;
    cudaSetDevice(c->device);
    long long* cyc = nullptr;
    if (cudaMalloc(&cyc, c->sm_count * 8) != cudaSuccess) return -1.0;
    size_t smem = 16384 + (size_t)240 * 128;
    cudaFuncSetAttribute(k_tc_bench_two_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_tc_bench_two_issuers<<<c->sm_count, 128, smem, c->stream>>>(8, variant, cyc);
    k_tc_bench_two_issuers<<<c->sm_count, 128, smem, c->stream>>>(iters, variant, cyc);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(cyc);
    c->launches += 2;
    if (e != cudaSuccess) { nclt_fail(c, NCLT_ERR_CUDA, "tc_bench_two_issuers", e); return -1.0; }
    return (double)h / iters;
}

// ---- CPU-callable hooks for the host logic of the fp4 engines (no GPU needed; tests/test_tc_tiles_cpu.py) -------------
#include "../tc_tiles.h"
// tile table of a library (counts[n_kf], or every keyframe `stride` rows when counts == NULL): fills the first `cap`
// entries of the per-tile arrays, pstart[n_kf + 1] and grp_tile[<= grp_cap]; returns the number of tiles
extern "C" int nclt_diag_tiles4(const int* counts, int n_kf, int stride, int row_bytes, int cap, unsigned* img_off256, int* n,
                                int* endmask, int* kf0, int* prow0, int* pstart, int* grp_tile, int grp_cap, int* n_grp) {
    nclt_tc4::Tiles4 t;
    nclt_tc4::build_tiles4(counts, n_kf, row_bytes, t, stride);
    for (int i = 0; i < (int)t.tiles.size() && i < cap; ++i) {
        img_off256[i] = t.tiles[i].img_off256; n[i] = t.tiles[i].n; endmask[i] = t.tiles[i].endmask;
        kf0[i] = t.tiles[i].kf0; prow0[i] = t.tiles[i].prow0;
    }
    for (int k = 0; k <= n_kf; ++k) pstart[k] = t.pstart[k];
    for (int g = 0; g < (int)t.grp_tile.size() && g < grp_cap; ++g) grp_tile[g] = t.grp_tile[g];
    if (n_grp) *n_grp = (int)t.grp_tile.size();
    return (int)t.tiles.size();
}
// the 32-byte (64 e2m1 nibbles) constant rows of the encodings: which = 0 / 1 matcher bias row (query / library side),
// 2 / 3 crossCheck bias row (query / library side), 4 crossCheck index row of value `arg`
extern "C" int nclt_diag_mx_row(int which, int arg, unsigned char* out32) {
    for (int b = 0; b < 32; ++b)
        out32[b] = which == 0 ? tc::mx_bias_byte(false, b) : which == 1 ? tc::mx_bias_byte(true, b) : which == 2 ? tc::mx_xbias_byte(false, b)
                 : which == 3 ? tc::mx_xbias_byte(true, b) : tc::mx_index_byte(arg, b);
    return 0;
}
