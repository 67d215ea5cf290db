// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers for the tensor-core Hamming path (sm_100a).
//
// Operand encoding: a 256-bit ORB descriptor becomes 256 fp8 (e4m3) values, bit 0 -> +1.0 (0x38),
// bit 1 -> -1.0 (0xB8), so that  x . y = 256 - 2 * Hamming(a, b)  - exact in fp16/fp32
// accumulators (|sum| <= 256).  Tiles live in shared memory in the canonical K-major,
// no-swizzle UMMA layout: core matrices of 8 rows x 16 bytes,
//     offset(row r, k) = (k/16) * LBO + (r/8) * 128 + (r%8) * 16 + k%16,   LBO = rows * 16,
// i.e. [16 k-chunks][rows/8][8][16 B]; global memory keeps the operands in exactly this image
// form so that one 1-D TMA bulk copy per tile fills shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tc {

constexpr uint32_t FP8_POS1 = 0x38u;   // +1.0 e4m3
constexpr uint32_t FP8_NEG1 = 0xB8u;   // -1.0 e4m3

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// for waiters with slack (the TMA producer): back off between polls instead of hammering the barrier unit
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}
// 1-D TMA bulk copy global -> shared with byte-count completion on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (elect.sync): control flow around it stays warp-uniform, so ptxas keeps the operands of
// the tcgen05 instructions in uniform registers instead of moving them there one by one (R2UR) from a divergent lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n.reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMEM ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp as alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------
// K-major, no swizzle: LBO = byte distance between the two 16-byte K chunks of one K=32 MMA,
// SBO = byte distance between 8-row groups. version = 1 (Blackwell), layout_type = 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// kind::f8f6f4, A = B = e4m3 (format 0), both K-major; c_fmt 0 = f16, 1 = f32 accumulators
__host__ __device__ constexpr uint32_t idesc_f8(int M, int N, int c_fmt) {
    return ((uint32_t)c_fmt << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread -> arrive on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM loads: lane = TMEM data path (row), N consecutive 32-bit columns per thread ---------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// .pack::16b: two 16-bit accumulators of adjacent columns per destination register
__device__ __forceinline__ void tmem_ld32_pack16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16_pack16(uint32_t taddr, uint32_t* r) {   // 32 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8_pack16(uint32_t taddr, uint32_t* r) {    // 16 columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4_pack16(uint32_t taddr, uint32_t* r) {    // 8 columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.pack::16b.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {   // 16 columns, 32-bit cells
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {    // 8 columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// registers -> TMEM, one constant into 8 / 16 / 32 / 64 consecutive columns of this warp's 32 lanes (no wait)
#define NCLT_R8 "%1,%1,%1,%1,%1,%1,%1,%1"
#define NCLT_R16 NCLT_R8 "," NCLT_R8
#define NCLT_R32 NCLT_R16 "," NCLT_R16
#define NCLT_R64 NCLT_R32 "," NCLT_R32
__device__ __forceinline__ void tmem_st8_const(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {" NCLT_R8 "};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st16_const(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {" NCLT_R16 "};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st32_const(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {" NCLT_R32 "};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st64_const(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {" NCLT_R64 "};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- kind::mxf4 (block-scaled fp4, K = 64 per instruction, f32 accumulators) ------------------------------
// +-1.0 as e2m1 nibbles (0x2 / 0xA); every scale factor is UE8M0 0x7F = 1.0 (a TMEM region filled with
// 0x7F7F7F7F, so the scale-factor layout does not matter).
// 16 descriptor bits -> 16 nibbles = 8 bytes; bit t -> nibble t (low nibble first; A and B agree, which is all a dot product needs)
__device__ __forceinline__ uint2 expand16_fp4(uint32_t bits16) {
    uint32_t w[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t b = (bits16 >> (8 * i)) & 0xFFu, s = 0;
#pragma unroll
        for (int t = 0; t < 8; ++t) s |= ((b >> t) & 1u) << (4 * t + 3);
        w[i] = 0x22222222u | s;
    }
    return make_uint2(w[0], w[1]);
}
// byte offset of (row r, byte kb of its 128-byte fp4 row) inside a K-major no-swizzle tile image of `rows` rows
__host__ __device__ __forceinline__ uint32_t image_offset4(int rows, int r, int kb) {
    return (uint32_t)(kb >> 4) * (uint32_t)rows * 16u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kb & 15);
}
// block-scaled instruction descriptor: a/b format E2M1 (1) at bits 7 / 10, N>>3 at 17, scale format UE8M0 at 23,
// M>>4 at 24, scale-factor ids 0
__host__ __device__ constexpr uint32_t idesc_mxf4(int M, int N) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_mxf4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, uint32_t sfa,
                                         uint32_t sfb) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
}
// Accumulators are pre-loaded with 1.5 * 2^23 + 0x4000 (f32 bits 0x4B404000): the exact integer sum
// 256 - 2H then sits in the low mantissa bits, so the LOW 16 bits of the f32 cell are 0x4100 - 2H, a positive
// normal fp16 bit pattern that is monotone in -H: .pack::16b loads + half2 max work exactly as on fp16 accumulators.
constexpr uint32_t MX_MAGIC = 0x4B404000u;
constexpr uint32_t MX_ZERO16 = 0x4100u;      // low 16 bits at H = 0
// The same bias produced BY THE TENSOR CORE: one extra K = 64 step whose operands are constant rows
//   u = 21 x 6.0, 3.0, 1.0, 0...   (A side)      v = 21 x 6.0, 4.0, 1.0, 0...   (B side)
// so u . v = 21 * 36 + 12 + 1 = 769, with scale factors 2^14 (A, UE8M0 0x8D) x 1.0 (B): 769 * 2^14 = 1.5 * 2^23 + 0x4000.
// All scale-factor bytes of a region are equal, so the scale-factor layout in TMEM is irrelevant; every partial sum is a
// multiple of 2^14 below 2^24, i.e. exact whatever the internal alignment of the adder tree.  Issued FIRST (accumulate = 0);
// the four real steps then add +-1 products to an accumulator of 2^23-magnitude (ulp 1): exact.
constexpr uint32_t MX_BIAS_SFA = 0x8D8D8D8Du;     // UE8M0 2^14 in every byte
__host__ __device__ __forceinline__ uint8_t mx_bias_byte(bool b_side, int byte /*0..31 of the 64-nibble row*/) {
    return byte < 10 ? 0x77 : byte == 10 ? (b_side ? 0x67 : 0x57) : byte == 11 ? 0x02 : 0x00;
}
// ---- index-carrying cells (crossCheck on the tensor core) ---------------------------------------------------------
// cell = 2^23 + 256 * (256 - H) + (255 - column): the row maximum is the nearest library row of the tile AND its column,
// lowest column on ties (cv2's tie rule).  Six K = 64 steps per tile: a bias step with u . v = 257 at A-scale 2^15
// (= 2^23 + 2^15), an index step whose library-side row for column c sums to 255 - c (A side all 1.0, scales 1.0), and
// the four data steps at A-scale 2^7 (128 * (256 - 2 H) = 256 * (256 - H) - 2^15).  Every partial sum is an integer
// multiple of its scale below 2^24: exact.  Decode: x = bits - 0x4B000000; H = 256 - (x >> 8); column = 255 - (x & 255).
constexpr uint32_t MX_X_SFA_DATA = 0x86868686u;     // UE8M0 2^7 in every byte
constexpr uint32_t MX_X_SFA_BIAS = 0x8E8E8E8Eu;     // UE8M0 2^15
constexpr uint32_t MX_X_BASE = 0x4B000000u;         // f32 bits of 2^23
// u = 7 x 6.0, 4.0, 1.0    v = 7 x 6.0, 1.0, 1.0     u . v = 252 + 4 + 1 = 257
__host__ __device__ __forceinline__ uint8_t mx_xbias_byte(bool b_side, int byte /*0..31*/) {
    return byte < 3 ? 0x77 : byte == 3 ? (b_side ? 0x27 : 0x67) : byte == 4 ? 0x02 : 0x00;
}
// library-side index row of column c: e2m1 values summing to n = 255 - c (sixes, then the remainder as 4 / 3 / 2 / 1 [+ 1])
__host__ __device__ __forceinline__ uint8_t mx_index_byte(int n, int byte /*0..31*/) {
    const int n6 = n / 6, rem = n - 6 * n6;
    uint8_t out = 0;
    for (int h = 0; h < 2; ++h) {
        const int e = 2 * byte + h - n6;          // position after the sixes
        uint8_t code = 0;
        if (e < 0) code = 7;                      // 6.0
        else if (e == 0) code = rem == 0 ? 0 : rem == 1 ? 2 : rem == 2 ? 4 : rem == 3 ? 5 : 6;   // 1, 2, 3, 4 (5 = 4 + 1)
        else if (e == 1) code = rem == 5 ? 2 : 0;
        out |= (uint8_t)(code << (4 * h));
    }
    return out;
}
// query-side constant slab of the crossCheck kernel: rows of K = 64: which = 0 bias (u above), 1 index (all 1.0)
__device__ __forceinline__ void mx_fill_xslab(uint8_t* slab, int rows, int which, int tid, int nthreads) {
    for (int i = tid; i < rows * 2; i += nthreads) {       // (row, 16-byte k chunk)
        const int r = i >> 1, kc = i & 1;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)(which ? 0x22 : mx_xbias_byte(false, kc * 16 + j * 4 + b)) << (8 * b);
            w[j] = v;
        }
        *reinterpret_cast<uint4*>(slab + (uint32_t)kc * (uint32_t)rows * 16u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// fills a K = 64 (32 bytes per row) K-major no-swizzle bias slab of `rows` rows; all threads of the CTA
__device__ __forceinline__ void mx_fill_bias_slab(uint8_t* slab, int rows, bool b_side, int tid, int nthreads) {
    for (int i = tid; i < rows * 2; i += nthreads) {       // (row, 16-byte k chunk)
        const int r = i >> 1, kc = i & 1;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) v |= (uint32_t)mx_bias_byte(b_side, kc * 16 + j * 4 + b) << (8 * b);
            w[j] = v;
        }
        *reinterpret_cast<uint4*>(slab + (uint32_t)kc * (uint32_t)rows * 16u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u) =
            make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// byte offset of element (row r, k-byte k) inside a tile image of `rows` rows
__host__ __device__ __forceinline__ uint32_t image_offset(int rows, int r, int k) {
    return (uint32_t)(k >> 4) * (uint32_t)rows * 16u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 15);
}

// 16 descriptor bits -> 16 fp8 bytes (+-1.0), as one uint4
__device__ __forceinline__ uint4 expand16(uint32_t bits16) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t b = (bits16 >> (4 * i)) & 0xFu;
        // spread 4 bits into the sign bits of 4 bytes: bit j -> byte j bit 7
        uint32_t s = ((b & 1u) << 7) | ((b & 2u) << 14) | ((b & 4u) << 21) | ((b & 8u) << 28);
        w[i] = 0x38383838u | s;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

}  // namespace tc
