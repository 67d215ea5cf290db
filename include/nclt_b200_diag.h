/* nclt_b200_diag.h - diagnostic / micro-benchmark entry points, exported by libnclt_b200_diag.so
 * (nclt-slam-project_b200/csrc/diag/, built next to the product library and linked against it).
 *
 * NOT part of the drop-in boundary (include/nclt_b200.h) and NOT in the product library: these exist so that the
 * building blocks of the tensor-core matcher can be validated and measured in isolation (tests/test_tc_gpu.py,
 * tools/tcbench.py, tools/mxf4_probe.py, tools/tmem_bw.py, tools/two_issuers.py, bench.py's roofline peak).
 * Two read-outs of product-kernel state stay in libnclt_b200.so and are marked below.
 * Same conventions: int functions return 0 or a negative NCLT_ERR_*; double functions return < 0 on error. */
#ifndef NCLT_B200_DIAG_H
#define NCLT_B200_DIAG_H
#include "nclt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One 128 x N x 256-bit tile through tcgen05.mma kind::f8f6f4 (+-1 as e4m3): a_desc u8[128,32], b_desc u8[N,32]
 * (HOST), N multiple of 16 <= 256; c_fmt 0 = f16 / 1 = f32 accumulators; ld_mode 1 = .pack::16b TMEM loads.
 * out: raw 32-bit TMEM cells u32[128,N] (or [128,N/2] packed).  Expected value: 256 - 2 * Hamming. */
int nclt_tc_probe(nclt_ctx* ctx, const uint8_t* a_desc, const uint8_t* b_desc, int N, int c_fmt, int ld_mode,
                  uint32_t* out);
/* The same through kind::mxf4.block_scale (+-1 as e2m1, all scale factors 1.0, f32 accumulators), N <= 240;
 * magic = 1 pre-loads the accumulators with 1.5 * 2^23 + 0x4000 (cell bits = 0x4B404000 + 256 - 2 * Hamming);
 * magic = 2 produces the same bias with one extra MMA step on constant operands (tc_common.cuh, MX_BIAS_SFA). */
int nclt_tc_probe_mxf4(nclt_ctx* ctx, const uint8_t* a_desc, const uint8_t* b_desc, int N, int magic, uint32_t* out);
/* Rates on resident tiles, one CTA per SM; return comparisons/s, *cycles_per_tile = SM clocks per 128 x N tile.
 * nclt_tc_bench (fp8): mode 0 MMA only, 1 + packed TMEM read-back, 2 + exact half2 top-2.
 * nclt_tc_bench_mxf4: mode 0 MMA only, 1 + f32 loads and fmaxf, 2/3 pre-armed + packed loads + re-arm stores
 * (2 / 3 epilogue warps per lane quadrant), 4 packed loads + half2 max, 5-7 f32 loads + three-input max
 * (2 / 3 / 4 warps per quadrant), 8 = five MMAs per tile (bias step first) + packed loads + half2 max, 9 = the five
 * MMAs alone.  nclt_tc_bench_mx16: 16 epilogue warps in two sets, variant 0 = VIMNMX3.u16x2. */
double nclt_tc_bench(nclt_ctx* ctx, int N, int iters, int mode, double* cycles_per_tile);
double nclt_tc_bench_mxf4(nclt_ctx* ctx, int N, int iters, int mode, double* cycles_per_tile);
double nclt_tc_bench_mx16(nclt_ctx* ctx, int iters, int variant, double* cycles_per_tile);
/* the packed design of k_tc4_top2 in isolation (round 2): five MMAs per tile (bias step first), two epilogue sets on
 * alternate tiles, all 240 columns of a lane quadrant in ONE batch of packed loads (variant 0) or in two (variant 1) */
double nclt_tc_bench_mxp(nclt_ctx* ctx, int iters, int variant, double* cycles_per_tile);
/* TMEM read-out: bytes per clock per SM that `warps` warps obtain with `batch` 32-column tcgen05.ld per wait. */
double nclt_tmem_bw(nclt_ctx* ctx, int warps, int batch, int with_max);
/* SM clocks per 128 x 240 x 256 mxf4 tile when one thread issues every tile (variant 0) or two warps alternate. */
double nclt_tc_bench_two_issuers(nclt_ctx* ctx, int iters, int variant);
/* [in libnclt_b200.so] With nclt_ctx_profile(ctx, 1): effective SM clock (clock64 / globaltimer of the longest CTA) and duration of the
 * last k_tc*_top2 launch; raw64 (optional, 64 x u64): [0] cycles, [1] ns, [2..15] phase counters of a
 * -DNCLT_TC_TIMING build, [16 + 2i] / [17 + 2i] globaltimer ns of the earliest CTA start / latest CTA end of the
 * i-th most recent launches (ring of 24, slot = launch number mod 24). */
int nclt_ctx_tc_clock(nclt_ctx* ctx, double* mhz, double* kernel_ms, unsigned long long* raw64);

/* [in libnclt_b200.so] Intermediate planes of the last nclt_orb_detect_and_compute call (stage-by-stage parity tests): what 0 = pyramid
 * level, 1 = FAST score map (score, 0 = no corner; only defined >= 30 px from the border), 2 = blurred level.
 * out: HOST u8[h,w] of that level (nclt_orb_levels gives w, h). */
int nclt_orb_debug_plane(nclt_ctx* ctx, nclt_orb* orb, int what, int frame, int level, uint8_t* out);

/* CPU-callable hooks for the host logic of the fp4 engines (no GPU needed; tests/test_tc_tiles_cpu.py): the tile table
 * of a library image (csrc/tc_tiles.h) and the constant operand rows of the bias / index encodings (csrc/tc_common.cuh) */
int nclt_diag_tiles4(const int* counts, int n_kf, int stride, int row_bytes, int cap, unsigned* img_off256, int* n,
                     int* endmask, int* kf0, int* prow0, int* pstart, int* grp_tile, int grp_cap, int* n_grp);
int nclt_diag_mx_row(int which, int arg, unsigned char* out32);

#ifdef __cplusplus
}
#endif
#endif
