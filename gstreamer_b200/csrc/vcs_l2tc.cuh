// gstreamer_b200/csrc/vcs_l2tc.cuh — the exact-2:1 / 8-tap kernel with both FIR passes on the 5th-generation tensor
// cores (product code, sm_100a only: tcgen05.mma.kind::i8 with TMEM accumulators).
//
// Why: the measured ablation of vcs_lanczos2_kernel (profiles/r02_l2_ablation_*.txt) puts 1.15 us/frame in the horizontal
// FIR's dot products alone and as much again in the shuffles, shifts and packs around it, against an HBM floor of 3.2 us:
// the SIMT kernel is bound by integer issue, not by memory.  Both FIRs are banded u8 x s8 matrix products with exact
// 32-bit accumulation, so they can leave the issue slots altogether:
//
//   H pass   D_h[j][l] = sum_x band[j][x] * img[l][x]       A = tap band of the strip's 128 output columns (s8, K-major)
//                                                           B = the staged input lines of one channel (u8, K = pixels)
//   V pass   D_v[c][r] = sum_l hs[c][l]   * vband[r][l]     A = h-scaled samples, one row per output column (u8, K = lines)
//                                                           B = tap band of the tile's output rows (s8)
//
// With the band as the A operand the accumulator of an output COLUMN lives in one TMEM lane, so the thread that owns the
// lane reads consecutive LINES of its column, packs four of them per word and writes the V pass's K-major operand with
// 16-byte stores; the V accumulators come back one output column per lane again, which makes the final stores coalesced
// rows.  Rounding is one more K step against a block of ones ((acc + 32) >> 6, or taps times 4 and + 128 so that the
// saturated result is byte 1 of a 16-bit saturating pack); garbage in staged samples outside the frame meets zero taps
// (the reference folds its edge taps inward, see pack of the bands below), so no clamped copies are needed.
// What stays SIMT is what is not linear: de-interleaving and the two rounding chroma up-sampling steps
// (video-chroma.c:687-699, video-orc.orc:2705-2735), saturation, and the AYUV -> ARGB mulhi matrix.
//
// Shared-memory operands use the un-swizzled canonical K-major layout (8 rows x 16 bytes = one contiguous 128-byte core
// matrix; LBO = distance of the next 16-byte K chunk, SBO = distance of the next 8 rows), which plain 16-byte stores and
// cp.async produce without any address swizzling.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#include "common.h"
#include "vcs_device.h"
#include "vcs_plan.h"
#include "vcs_lanczos2.cuh"

namespace b200 {

constexpr int TC_THREADS = 256;
constexpr int TC_TW = 128;                                 // output columns per tile = rows of the band (UMMA M)
constexpr int TC_TH = 28;                                  // output rows per tile
constexpr int TC_N = 64;                                   // staged input lines (2 * TH + 6 = 62 used) = UMMA N of the H pass
constexpr int TC_CHUNKS = 18;                              // 16-byte chunks per staged line: 288 input bytes (2 * 128 + 6, 16-aligned start)
constexpr int TC_X_LEAD = 16;                              // the staged window starts 16 pixels left of the strip's first tap centre - 3
constexpr int TC_IMG_LBO = 8 * 128 + 16;                   // bytes between chunk columns of a staged plane (+16: bank spread)
constexpr int TC_IMG_BYTES = TC_CHUNKS * TC_IMG_LBO;       // 18 720
constexpr int TC_BAND_CHUNKS = TC_CHUNKS + 2;              // + the rounding K step
constexpr int TC_BAND_LBO = 16 * 128;                      // 128 rows
constexpr int TC_BAND_BYTES = TC_BAND_CHUNKS * TC_BAND_LBO;   // 40 960
constexpr int TC_ONES_BYTES = 2 * TC_BAND_LBO;             // [2 chunks][128 rows][16]: every byte 1
constexpr int TC_VROWS = 32;                               // UMMA N of the V pass (28 used)
constexpr int TC_VB_CHUNKS = TC_N / 16 + 2;
constexpr int TC_VB_LBO = (TC_VROWS / 8) * 128;            // 512
constexpr int TC_VB_BYTES = TC_VB_CHUNKS * TC_VB_LBO;      // 3 072
constexpr int TC_HS_LBO = 16 * 128;
constexpr int TC_HS_BYTES = (TC_N / 16) * TC_HS_LBO;       // per channel 8 192
// shared-memory map
constexpr int TC_OFF_BAND = 0;
constexpr int TC_OFF_ONES = TC_OFF_BAND + TC_BAND_BYTES;
constexpr int TC_OFF_VB = TC_OFF_ONES + TC_ONES_BYTES;
constexpr int TC_OFF_IMG = TC_OFF_VB + TC_VB_BYTES;         // 3 planes; the h-scaled operand re-uses this space after the H pass
constexpr int TC_OFF_BAR = TC_OFF_IMG + 3 * TC_IMG_BYTES;
constexpr int TC_SMEM = TC_OFF_BAR + 64;
constexpr int TC_TMEM_COLS = 256;                           // 3 x 64 accumulator columns (H), re-used as 3 x 32 (V)
static_assert (3 * TC_HS_BYTES <= 3 * TC_IMG_BYTES, "h-scaled operand must fit the staged planes");
static_assert (TC_OFF_IMG % 16 == 0 && TC_IMG_BYTES % 16 == 0, "operand alignment");

struct L2tcDev {
  const uint8_t *band;          // [strips][TC_BAND_BYTES]   tap band of each 128-column strip, canonical layout, + rounding step
  const uint8_t *vband;         // [row tiles][TC_VB_BYTES]
  const uint8_t *hx4, *vx4;     // per strip / row tile: taps are stored times 4 (rounding 128, result in byte 1)
  int strips, row_tiles;
  unsigned magic_rt;            // 2^32 / row_tiles + 1: t / row_tiles == umulhi (t, magic) for the tile counts that occur
};

// ---- PTX wrappers (sm_100a) -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }
__device__ __forceinline__ uint64_t tc_desc (uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
  // matrix descriptor, un-swizzled K-major: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout 0 << 61
  return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo >> 4) & 0x3fffu) << 16) |
      ((uint64_t) ((sbo >> 4) & 0x3fffu) << 32) | ((uint64_t) 1 << 46);
}
// instruction descriptor for kind::i8: D = s32, A / B formats (0 = u8, 1 = s8), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t tc_idesc (int a_signed, int b_signed, int n, int m)
{
  return (2u << 4) | ((uint32_t) a_signed << 7) | ((uint32_t) b_signed << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24);
}
__device__ __forceinline__ void tc_mma (uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
#ifndef B200_CUDA_EMU
  asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r" (d_tmem), "l" (adesc), "l" (bdesc), "r" (idesc), "r" (accumulate), "r" (0u) : "memory");
#endif
}
__device__ __forceinline__ void tc_commit (uint32_t bar)
{
#ifndef B200_CUDA_EMU
  asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r" (bar) : "memory");
#endif
}
__device__ __forceinline__ void tc_fence_before () {
#ifndef B200_CUDA_EMU
  asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_fence_after () {
#ifndef B200_CUDA_EMU
  asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_fence_async_smem () {
#ifndef B200_CUDA_EMU
  asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_bar_wait (uint32_t bar, uint32_t parity)
{
#ifndef B200_CUDA_EMU
  uint32_t ok;
  do {
    asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r" (ok) : "r" (bar), "r" (parity) : "memory");
  } while (!ok);
#endif
}
__device__ __forceinline__ void tc_cp16 (uint32_t saddr, const void *g)
{
#ifndef B200_CUDA_EMU
  asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r" (saddr), "l" (g) : "memory");
#endif
}
__device__ __forceinline__ void tc_cp_wait ()
{
#ifndef B200_CUDA_EMU
  asm volatile ("cp.async.wait_all;" ::: "memory");
#endif
}
#define TC_LD16(v, taddr)                                                                                          \
  asm volatile ("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"  \
      : "=r" (v[0]), "=r" (v[1]), "=r" (v[2]), "=r" (v[3]), "=r" (v[4]), "=r" (v[5]), "=r" (v[6]), "=r" (v[7]),       \
        "=r" (v[8]), "=r" (v[9]), "=r" (v[10]), "=r" (v[11]), "=r" (v[12]), "=r" (v[13]), "=r" (v[14]), "=r" (v[15])   \
      : "r" (taddr) : "memory")
__device__ __forceinline__ void tc_ld_wait () {
#ifndef B200_CUDA_EMU
  asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
}

// four consecutive FIR sums -> four rounded, saturated bytes (lowest byte = first sum); the rounding constant is already
// in the accumulator (the extra K step)
template <bool X4>
__device__ __forceinline__ unsigned tc_pack4 (int a0, int a1, int a2, int a3)
{
  if (X4) return __byte_perm (pack_sat_u16x2 (a1, a0), pack_sat_u16x2 (a3, a2), 0x7531);
  return pack_sat2 (a1 >> 6, a0 >> 6, pack_sat2 (a3 >> 6, a2 >> 6, 0u));
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
// One persistent CTA per SM, warp-specialised (the canonical sm_100 pipeline, with SIMT stages where a GEMM has TMA):
//
//   warps 0-7   PRODUCERS   stage tile k+1: Y by cp.async, prepared chroma by SIMT, tap bands -> planes[k&1]
//   warps 8-15  CONSUMERS   tile k: H epilogue (TMEM -> bytes -> V operand), V epilogue (TMEM -> matrix -> stores)
//   warp 16     MMA ISSUER  one lane: H pass of tile k+1 (30 MMAs), then V pass of tile k (9 MMAs)
//
// mbarriers:  full[2] (producers -> issuer: planes staged), empty[2] (commit of the H pass: planes free again),
// dh_full[2] / dh_empty[2] (H accumulators of the two TMEM buffers), hs_full (consumers -> issuer: V operand written),
// dv_full (commit of the V pass).  TMEM: 2 x 192 columns for the H pass, 96 for the V pass.
constexpr int TC_PROD_WARPS = 8, TC_CONS_WARPS = 8;
constexpr int TC_THREADS2 = 32 * (TC_PROD_WARPS + TC_CONS_WARPS + 1);
constexpr int TC_VB_RING = 4;
constexpr int TC2_OFF_BAND = 0;
constexpr int TC2_OFF_ONES = TC2_OFF_BAND + TC_BAND_BYTES;
constexpr int TC2_OFF_VB = TC2_OFF_ONES + TC_ONES_BYTES;
constexpr int TC2_OFF_IMG = TC2_OFF_VB + TC_VB_RING * TC_VB_BYTES;      // 2 buffers x 3 planes
constexpr int TC2_OFF_HS = TC2_OFF_IMG + 2 * 3 * TC_IMG_BYTES;
constexpr int TC2_OFF_BAR = TC2_OFF_HS + 3 * TC_HS_BYTES;
constexpr int TC2_SMEM = TC2_OFF_BAR + 128;
constexpr int TC2_TMEM_COLS = 512;
constexpr int TC2_DV_COL = 2 * 3 * TC_N;                                // 384: the V accumulators
enum { TCB_FULL0 = 0, TCB_FULL1, TCB_EMPTY0, TCB_EMPTY1, TCB_DHF0, TCB_DHF1, TCB_DHE0, TCB_DHE1, TCB_HSF, TCB_DVF, TCB_COUNT };

__device__ __forceinline__ void tc_bar_wait_guard (uint32_t bar, uint32_t parity)
{
#ifndef B200_CUDA_EMU
  uint32_t ok, spins = 0;
  do {
    asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r" (ok) : "r" (bar), "r" (parity) : "memory");
    if (!ok) {
      __nanosleep (40);                                           // a waiting warp should not compete for issue slots
      if (++spins > (1u << 24)) __trap ();                        // a lost arrival must fail loudly, not hang the device
    }
  } while (!ok);
#endif
}
__device__ __forceinline__ bool tc_bar_test (uint32_t bar, uint32_t parity)
{
#ifndef B200_CUDA_EMU
  uint32_t ok;
  asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r" (ok) : "r" (bar), "r" (parity) : "memory");
  return ok != 0;
#else
  return true;
#endif
}
__device__ __forceinline__ void tc_bar_arrive (uint32_t bar)
{
#ifndef B200_CUDA_EMU
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r" (bar) : "memory");
#endif
}

template <bool X4, bool FULL>
__device__ __forceinline__ void tc_v_rows (const VcsDev & P, const int (&a)[3][16], uint8_t *dst, int rows)
{
#pragma unroll
  for (int i = 0; i < 14; i++) {
    if (FULL || i < rows) {
      unsigned yuv = X4 ? __byte_perm (pack_sat_u16x2 (a[1][i], a[0][i]), pack_sat_u16x2 (0, a[2][i]), 0x7531)
          : pack_sat2 (a[1][i] >> 6, a[0][i] >> 6, pack_sat2 (0, a[2][i] >> 6, 0u));
      yuv ^= 0x00808080u;
      const int wy = prmt_s (yuv, 0x8800u), wu = prmt_s (yuv, 0x9911u), wv = prmt_s (yuv, 0xaa22u);
      const int ty = ((wy * P.p1) >> 16) + 128;
      const int r = ty + ((wv * P.p2) >> 16);
      const int b = ty + ((wu * P.p3) >> 16);
      const int gg = ty + ((wu * P.p4) >> 16) + ((wv * P.p5) >> 16);
      const unsigned argb = pack_sat2 (r, 255, pack_sat2 (b, gg, 0u));
      *(unsigned *) dst = __byte_perm (argb, 0, P.sel);
    }
    dst += P.stride_out;
  }
}

template <int DBG>
__global__ void __launch_bounds__ (TC_THREADS2, 1)
vcs_l2tc_kernel (const VcsDev P, const L2tcDev L, const VcsBatch frames, int n_frames, unsigned *dbg)
{
#ifndef B200_CUDA_EMU
  extern __shared__ __align__ (128) uint8_t sm[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync (0xffffffffu, tid >> 5, 0);
  const uint32_t s_base = tc_smem_u32 (sm);
  const uint32_t s_band = s_base + TC2_OFF_BAND, s_ones = s_base + TC2_OFF_ONES, s_vb = s_base + TC2_OFF_VB,
      s_img = s_base + TC2_OFF_IMG, s_hs = s_base + TC2_OFF_HS, s_bar = s_base + TC2_OFF_BAR, s_tptr = s_bar + 8 * TCB_COUNT;
  auto bar = [&] (int i) { return s_bar + 8u * (uint32_t) i; };

  if (tid == 0) {
    const int counts[TCB_COUNT] = {TC_PROD_WARPS, TC_PROD_WARPS, 1, 1, 1, 1, TC_CONS_WARPS, TC_CONS_WARPS, TC_CONS_WARPS, 1};
    for (int i = 0; i < TCB_COUNT; i++)
      asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r" (bar (i)), "r" (counts[i]) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_PROD_WARPS + TC_CONS_WARPS) {
    asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r" (s_tptr), "r" (TC2_TMEM_COLS) : "memory");
    asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < TC_ONES_BYTES / 16; i += TC_THREADS2)
    *(uint4 *) (sm + TC2_OFF_ONES + 16 * i) = make_uint4 (0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
  tc_fence_async_smem ();
  tc_fence_before ();
  __syncthreads ();
  tc_fence_after ();
  const uint32_t tmem = *(volatile uint32_t *) (sm + TC2_OFF_BAR + 8 * TCB_COUNT);

  const int per_strip = n_frames * L.row_tiles;
  const int n_tiles = L.strips * per_strip;
  const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
  const unsigned magic_ps = 0xffffffffu / (unsigned) per_strip + 1u;    // tiles < 2^16 * ... : exact for t < 2^32 / per_strip
  auto tile_of = [&] (int k, int & strip, int & f, int & rt) {
    const unsigned t = blockIdx.x + (unsigned) k * gridDim.x;
    strip = (int) __umulhi (t, magic_ps);
    if ((unsigned) (strip + 1) * (unsigned) per_strip <= t) strip++;          // the reciprocals round down at most one short
    const unsigned rem = t - (unsigned) strip * (unsigned) per_strip;
    f = (int) __umulhi (rem, L.magic_rt);
    if ((unsigned) (f + 1) * (unsigned) L.row_tiles <= rem) f++;
    rt = (int) (rem - (unsigned) f * (unsigned) L.row_tiles);
  };

  if (warp < TC_PROD_WARPS) {
    // ================================================================ PRODUCERS
    const int crows = P.ih >> 1;
    const unsigned selU = P.u_index ? 0x7531u : 0x6420u, selV = P.u_index ? 0x6420u : 0x7531u;
    const unsigned nselU = P.u_index ? 0x5321u : 0x4321u, nselV = P.u_index ? 0x4321u : 0x5321u;
    int prev_strip = -1;
    for (int k = 0; k < my_tiles; k++) {
      int strip, f, rt;
      tile_of (k, strip, f, rt);
      const int buf = k & 1, use = k >> 1;
      if (use > 0) tc_bar_wait_guard (bar (TCB_EMPTY0 + buf), (use - 1) & 1);          // H pass of tile k-2 has read these planes
      if (strip != prev_strip) {
        // the band is single-buffered: the H pass of tile k-1 must have finished with the old one
        if (k > 0) tc_bar_wait_guard (bar (TCB_EMPTY0 + (buf ^ 1)), ((k - 1) >> 1) & 1);
        const uint8_t *src = L.band + (size_t) strip * TC_BAND_BYTES;
        for (int i = tid; i < TC_BAND_BYTES / 16; i += 32 * TC_PROD_WARPS) tc_cp16 (s_band + 16 * i, src + 16 * i);
        prev_strip = strip;
      }
      const uint8_t *__restrict__ in = frames.in[f];
      const uint8_t *__restrict__ plane_y = in + P.off_y;
      const uint8_t *__restrict__ plane_c = in + P.off_c;
      const int x0 = strip * TC_TW, oy0 = rt * TC_TH;
      const int X0 = 2 * x0 - TC_X_LEAD, R0 = 2 * oy0 - 3;
      const uint32_t img = s_img + buf * 3 * TC_IMG_BYTES;
      uint8_t *img_u = sm + TC2_OFF_IMG + buf * 3 * TC_IMG_BYTES + TC_IMG_BYTES, *img_v = img_u + TC_IMG_BYTES;
      if (tid < TC_VB_BYTES / 16)
        tc_cp16 (s_vb + (k % TC_VB_RING) * TC_VB_BYTES + 16 * tid, L.vband + (size_t) rt * TC_VB_BYTES + 16 * tid);
      // Y: a warp takes lines w, w + 8, ...; lanes 0..17 one 16-byte chunk each
      {
        const int x = X0 + 16 * lane;
        const bool xin = lane < TC_CHUNKS && x >= 0 && x + 16 <= P.iw;
        const uint8_t *src = plane_y + (ptrdiff_t) (R0 + warp) * P.stride_y + x;
        const ptrdiff_t step = (ptrdiff_t) TC_PROD_WARPS * P.stride_y;
        for (int li = warp; li < 2 * TC_TH + 6; li += TC_PROD_WARPS, src += step) {
          const int y = R0 + li;
          if (xin && y >= 0 && y < P.ih) tc_cp16 (img + lane * TC_IMG_LBO + (li >> 3) * 128 + (li & 7) * 16, src);
        }
      }
      // chroma: item = (line group g, 8-pixel slot); units of 32 items, the odd unit rotates over the warps.  All raw
      // chroma of a warp's (at most 3) units is requested before the first byte-SIMD instruction needs any of it.
      {
        constexpr int SLOTS = 2 * TC_CHUNKS - 2, ITEMS = (TC_N / 4) * SLOTS, UNITS = (ITEMS + 31) / 32;
        constexpr int MAXU = (UNITS + TC_PROD_WARPS - 1) / TC_PROD_WARPS;
        const int u0 = (warp + k) % TC_PROD_WARPS;
        // tiles that touch no frame border: no row clamps, no right-edge replication, no column tests
        const bool interior = R0 >= 1 && R0 + 4 * (TC_N / 4) + 2 <= P.ih && X0 >= 0 && X0 + 16 * TC_CHUNKS + 8 <= P.iw;
        uint2 raw[MAXU][3];
        unsigned nxt[MAXU][3];
        int offs[MAXU];
        bool redge[MAXU];
#pragma unroll
        for (int j = 0; j < MAXU; j++) {
          const int item = 32 * (u0 + j * TC_PROD_WARPS) + lane;
          offs[j] = -1; redge[j] = false;
          if (u0 + j * TC_PROD_WARPS < UNITS && item < ITEMS) {
            const int g = item / SLOTS, slot = item - g * SLOTS + 1;
            const int x = X0 + 8 * slot;
            if (interior) {
              offs[j] = (slot >> 1) * TC_IMG_LBO + (g >> 1) * 128 + (g & 1) * 64 + (slot & 1) * 8;
              const uint8_t *row = plane_c + (ptrdiff_t) (oy0 - 2 + 2 * g) * P.stride_c + x;    // (R0 + 4g - 1) >> 1 == oy0 - 2 + 2g
#pragma unroll
              for (int kk = 0; kk < 3; kk++, row += P.stride_c) {
                raw[j][kk] = __ldg ((const uint2 *) row);
                nxt[j][kk] = (unsigned) __ldg ((const unsigned short *) (row + 8));
              }
            } else if (x >= 0 && x < P.iw) {
              redge[j] = x + 8 >= P.iw;
              const int m2 = (R0 + 4 * g - 1) >> 1;
              offs[j] = (slot >> 1) * TC_IMG_LBO + (g >> 1) * 128 + (g & 1) * 64 + (slot & 1) * 8;   // line 4g + r: + 16 r
#pragma unroll
              for (int kk = 0; kk < 3; kk++) {
                const int cr = min (max (m2 + kk, 0), crows - 1);
                const uint8_t *row = plane_c + (size_t) cr * P.stride_c + x;
                raw[j][kk] = __ldg ((const uint2 *) row);
                nxt[j][kk] = redge[j] ? 0u : (unsigned) __ldg ((const unsigned short *) (row + 8));
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < MAXU; j++) {
          if (offs[j] < 0) continue;
          unsigned ulo[3], uhi[3], vlo[3], vhi[3];
#pragma unroll
          for (int kk = 0; kk < 3; kk++) {
            const uint2 c = raw[j][kk];
            const unsigned ue = __byte_perm (c.x, c.y, selU), ve = __byte_perm (c.x, c.y, selV);
            const unsigned un = redge[j] ? __byte_perm (ue, ue, 0x3321) : __byte_perm (ue, nxt[j][kk], nselU);
            const unsigned vn = redge[j] ? __byte_perm (ve, ve, 0x3321) : __byte_perm (ve, nxt[j][kk], nselV);
            const unsigned uo = avg_ceil4 (ue, un), vo = avg_ceil4 (ve, vn);
            ulo[kk] = __byte_perm (ue, uo, 0x5140); uhi[kk] = __byte_perm (ue, uo, 0x7362);
            vlo[kk] = __byte_perm (ve, vo, 0x5140); vhi[kk] = __byte_perm (ve, vo, 0x7362);
          }
          uint2 U[4], V[4];
          {
            unsigned q;
            q = avg_floor4 (ulo[0], ulo[1]); U[0].x = avg_ceil4 (ulo[0], q); U[1].x = avg_ceil4 (ulo[1], q);
            q = avg_floor4 (uhi[0], uhi[1]); U[0].y = avg_ceil4 (uhi[0], q); U[1].y = avg_ceil4 (uhi[1], q);
            q = avg_floor4 (ulo[1], ulo[2]); U[2].x = avg_ceil4 (ulo[1], q); U[3].x = avg_ceil4 (ulo[2], q);
            q = avg_floor4 (uhi[1], uhi[2]); U[2].y = avg_ceil4 (uhi[1], q); U[3].y = avg_ceil4 (uhi[2], q);
            q = avg_floor4 (vlo[0], vlo[1]); V[0].x = avg_ceil4 (vlo[0], q); V[1].x = avg_ceil4 (vlo[1], q);
            q = avg_floor4 (vhi[0], vhi[1]); V[0].y = avg_ceil4 (vhi[0], q); V[1].y = avg_ceil4 (vhi[1], q);
            q = avg_floor4 (vlo[1], vlo[2]); V[2].x = avg_ceil4 (vlo[1], q); V[3].x = avg_ceil4 (vlo[2], q);
            q = avg_floor4 (vhi[1], vhi[2]); V[2].y = avg_ceil4 (vhi[1], q); V[3].y = avg_ceil4 (vhi[2], q);
          }
#pragma unroll
          for (int r = 0; r < 4; r++) {
            *(uint2 *) (img_u + offs[j] + 16 * r) = U[r];
            *(uint2 *) (img_v + offs[j] + 16 * r) = V[r];
          }
        }
      }
      tc_cp_wait ();
      tc_fence_async_smem ();
      __syncwarp ();
      if (lane == 0) tc_bar_arrive (bar (TCB_FULL0 + buf));
    }
  } else if (warp < TC_PROD_WARPS + TC_CONS_WARPS) {
    // ================================================================ CONSUMERS
    const int cw = warp - TC_PROD_WARPS, qd = cw & 3, hf = cw >> 2;
    const int col = 32 * qd + lane;
    const uint32_t lane_base = tmem + ((uint32_t) (32 * qd) << 16);
    for (int k = 0; k < my_tiles; k++) {
      int strip, f, rt;
      tile_of (k, strip, f, rt);
      const int acc = k & 1, use = k >> 1;
      // ---------------- H epilogue: 32 lines of one column per thread and channel
      tc_bar_wait_guard (bar (TCB_DHF0 + acc), use & 1);
      tc_fence_after ();
      const bool hx4 = __ldg (L.hx4 + strip) != 0;
#pragma unroll 1
      for (int ch = 0; ch < 3; ch++) {
        int v[32];
        const uint32_t ta = lane_base + acc * 3 * TC_N + ch * TC_N + 32 * hf;
        TC_LD16 (v, ta);
        TC_LD16 ((v + 16), (ta + 16));
        tc_ld_wait ();
        if (DBG == 1 && dbg && blockIdx.x == 0 && k == 0) {
#pragma unroll
          for (int i = 0; i < 32; i++) dbg[(ch * 128 + col) * 64 + 32 * hf + i] = (unsigned) v[i];
        }
        uint4 w0, w1;
        if (hx4) {
          w0.x = tc_pack4<true> (v[0], v[1], v[2], v[3]); w0.y = tc_pack4<true> (v[4], v[5], v[6], v[7]);
          w0.z = tc_pack4<true> (v[8], v[9], v[10], v[11]); w0.w = tc_pack4<true> (v[12], v[13], v[14], v[15]);
          w1.x = tc_pack4<true> (v[16], v[17], v[18], v[19]); w1.y = tc_pack4<true> (v[20], v[21], v[22], v[23]);
          w1.z = tc_pack4<true> (v[24], v[25], v[26], v[27]); w1.w = tc_pack4<true> (v[28], v[29], v[30], v[31]);
        } else {
          w0.x = tc_pack4<false> (v[0], v[1], v[2], v[3]); w0.y = tc_pack4<false> (v[4], v[5], v[6], v[7]);
          w0.z = tc_pack4<false> (v[8], v[9], v[10], v[11]); w0.w = tc_pack4<false> (v[12], v[13], v[14], v[15]);
          w1.x = tc_pack4<false> (v[16], v[17], v[18], v[19]); w1.y = tc_pack4<false> (v[20], v[21], v[22], v[23]);
          w1.z = tc_pack4<false> (v[24], v[25], v[26], v[27]); w1.w = tc_pack4<false> (v[28], v[29], v[30], v[31]);
        }
        uint8_t *hs = sm + TC2_OFF_HS + ch * TC_HS_BYTES + (2 * hf) * TC_HS_LBO + col * 16;
        *(uint4 *) hs = w0;
        *(uint4 *) (hs + TC_HS_LBO) = w1;
      }
      tc_fence_before ();
      tc_fence_async_smem ();
      __syncwarp ();
      if (lane == 0) { tc_bar_arrive (bar (TCB_DHE0 + acc)); tc_bar_arrive (bar (TCB_HSF)); }
      // ---------------- V epilogue: 14 rows of one column per thread
      tc_bar_wait_guard (bar (TCB_DVF), k & 1);
      tc_fence_after ();
      {
        const int x0 = strip * TC_TW, oy0 = rt * TC_TH, ox = x0 + col, r0 = 14 * hf;
        int a[3][16];
        const uint32_t ta = lane_base + TC2_DV_COL + r0;
        TC_LD16 (a[0], ta);
        TC_LD16 (a[1], (ta + TC_VROWS));
        TC_LD16 (a[2], (ta + 2 * TC_VROWS));
        tc_ld_wait ();
        tc_fence_before ();                                        // the next V pass may overwrite the accumulators once hs_full comes
        const int rows = min (min (TC_TH, P.oh - oy0) - r0, 14);
        if (ox < P.ow && rows > 0) {
          uint8_t *dst = frames.out[f] + P.off_out + (size_t) (oy0 + r0) * P.stride_out + (size_t) ox * 4u;
          const bool vx4 = __ldg (L.vx4 + rt) != 0;              // warp-uniform: branches, not predicates
          if (rows == 14) { if (vx4) tc_v_rows<true, true> (P, a, dst, 14); else tc_v_rows<false, true> (P, a, dst, 14); }
          else { if (vx4) tc_v_rows<true, false> (P, a, dst, rows); else tc_v_rows<false, false> (P, a, dst, rows); }
        }
      }
    }
  } else if (lane == 0) {
    // ================================================================ MMA ISSUER (one lane)
    constexpr uint32_t IDESC_H = tc_idesc (1, 0, TC_N, 128), IDESC_V = tc_idesc (0, 1, TC_VROWS, 128);
    // Both passes are issued from one lane in whatever order their inputs become ready: the H pass of tile k+1 must not
    // wait behind the V pass of tile k (that one waits for the consumers), nor the other way round (the H pass waits for
    // the producers).  The H pass runs at most one tile ahead (two TMEM / plane buffers).
    // descriptors differ only in their start-address field (bits 0..13, units of 16 bytes): base descriptor + offset
    const uint64_t d_band = tc_desc (s_band, TC_BAND_LBO, 128), d_ones = tc_desc (s_ones, TC_BAND_LBO, 128);
    const uint64_t d_img = tc_desc (s_img, TC_IMG_LBO, 128), d_hs = tc_desc (s_hs, TC_HS_LBO, 128);
    const uint64_t d_vb = tc_desc (s_vb, TC_VB_LBO, 128);
    auto issue_h_ch = [&] (int k, int ch) {                       // one channel of the H pass: 9 K steps + the rounding step
      const int buf = k & 1;
      const uint64_t img = d_img + (uint64_t) (((buf * 3 + ch) * TC_IMG_BYTES) >> 4);
      const uint32_t acc = tmem + buf * 3 * TC_N + ch * TC_N;
#pragma unroll
      for (int s = 0; s < TC_CHUNKS / 2; s++)
        tc_mma (acc, d_band + (uint64_t) ((2 * s * TC_BAND_LBO) >> 4), img + (uint64_t) ((2 * s * TC_IMG_LBO) >> 4), IDESC_H, s > 0);
      tc_mma (acc, d_band + (uint64_t) ((TC_CHUNKS * TC_BAND_LBO) >> 4), d_ones, IDESC_H, 1);
      if (ch == 2) {
        tc_commit (bar (TCB_EMPTY0 + buf));
        tc_commit (bar (TCB_DHF0 + buf));
      }
    };
    auto issue_v = [&] (int k) {
      tc_fence_after ();
      const uint64_t vb = d_vb + (uint64_t) (((k % TC_VB_RING) * TC_VB_BYTES) >> 4);
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const uint64_t hs = d_hs + (uint64_t) ((ch * TC_HS_BYTES) >> 4);
        const uint32_t acc = tmem + TC2_DV_COL + ch * TC_VROWS;
#pragma unroll
        for (int s = 0; s < TC_N / 32; s++)
          tc_mma (acc, hs + (uint64_t) ((2 * s * TC_HS_LBO) >> 4), vb + (uint64_t) ((2 * s * TC_VB_LBO) >> 4), IDESC_V, s > 0);
        tc_mma (acc, d_ones, vb + (uint64_t) ((2 * (TC_N / 32) * TC_VB_LBO) >> 4), IDESC_V, 1);
      }
      tc_commit (bar (TCB_DVF));
    };
    // The tensor pipe runs its MMAs in issue order, so the short V pass (consumers wait for it) is looked at between the
    // three channel-sized bites of the long H pass instead of after all of it.
    int nh = 0, hch = 0, nv = 0;                                  // hch: channels of H pass nh already issued
    uint32_t idle = 0;
    while (nv < my_tiles) {
      bool progressed = false;
      if (nv < nh && tc_bar_test (bar (TCB_HSF), nv & 1)) { issue_v (nv); nv++; progressed = true; }
      if (nh < my_tiles && nh <= nv + 1) {
        const int buf = nh & 1, use = nh >> 1;
        if (hch > 0 || (tc_bar_test (bar (TCB_FULL0 + buf), use & 1) && (use == 0 || tc_bar_test (bar (TCB_DHE0 + buf), (use - 1) & 1)))) {
          if (hch == 0) tc_fence_after ();
          issue_h_ch (nh, hch);
          if (++hch == 3) { hch = 0; nh++; }
          progressed = true;
        }
      }
      if (progressed) idle = 0;
      else if (++idle > (1u << 27)) __trap ();
    }
  }

  tc_fence_before ();
  __syncthreads ();
  if (warp == TC_PROD_WARPS + TC_CONS_WARPS)
    asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r" (tmem), "r" (TC2_TMEM_COLS) : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------------- host side
struct L2tcTables {
  std::vector<uint8_t> band, vband, hx4, vx4;
  int strips = 0, row_tiles = 0;
  bool ok = false;
};

// canonical K-major byte position of element (row, k) in an operand whose 16-byte K chunks are `lbo` apart
inline size_t tc_pos (int row, int k, int lbo) { return (size_t) (k >> 4) * lbo + (size_t) row * 16 + (k & 15); }

inline L2tcTables build_l2tc_tables (const VcsPlan & p, const Lanczos2Tables & l2)
{
  L2tcTables t;
  if (!l2.ok || !l2.alpha_opaque || p.planar) return t;           // same eligibility as the SIMT 2:1 kernel, opaque alpha, semi-planar only
  if ((p.in.width & 15) || (p.in.stride[0] & 15) || (p.in.stride[1] & 15) || (p.in.offset[0] & 15) || (p.in.offset[1] & 15)) return t;
  const int ow = p.out.width, oh = p.out.height, iw = p.in.width, ih = p.in.height;
  t.strips = (ow + TC_TW - 1) / TC_TW; t.row_tiles = (oh + TC_TH - 1) / TC_TH;
  t.band.assign ((size_t) t.strips * TC_BAND_BYTES, 0); t.hx4.assign (t.strips, 1);
  t.vband.assign ((size_t) t.row_tiles * TC_VB_BYTES, 0); t.vx4.assign (t.row_tiles, 1);
  for (int s = 0; s < t.strips; s++) {
    const int x0 = s * TC_TW, X0 = 2 * x0 - TC_X_LEAD;
    for (int j = 0; j < TC_TW && x0 + j < ow; j++)
      for (int k = 0; k < 8; k++) { const int v = 4 * p.h.coef[(size_t) (x0 + j) * 8 + k]; if (v < -128 || v > 127) t.hx4[s] = 0; }
    const int scale = t.hx4[s] ? 4 : 1;
    uint8_t *b = t.band.data () + (size_t) s * TC_BAND_BYTES;
    for (int j = 0; j < TC_TW; j++) {
      b[tc_pos (j, 16 * TC_CHUNKS, TC_BAND_LBO)] = (uint8_t) (t.hx4[s] ? 127 : 32);     // rounding step against the ones block ...
      if (t.hx4[s]) b[tc_pos (j, 16 * TC_CHUNKS + 1, TC_BAND_LBO)] = 1;                // ... 127 + 1 = 128 (s8 holds no 128)
      if (x0 + j >= ow) continue;
      for (int k = 0; k < 8; k++) {
        const int tap = p.h.coef[(size_t) (x0 + j) * 8 + k];
        if (!tap) continue;
        const int xr = (int) p.h.offset[x0 + j] + k - X0;
        if (xr < 0 || xr >= 16 * TC_CHUNKS || (int) p.h.offset[x0 + j] + k >= iw) return t;
        b[tc_pos (j, xr, TC_BAND_LBO)] = (uint8_t) (int8_t) (scale * tap);
      }
    }
  }
  for (int rt = 0; rt < t.row_tiles; rt++) {
    const int oy0 = rt * TC_TH, R0 = 2 * oy0 - 3;
    for (int r = 0; r < TC_TH && oy0 + r < oh; r++)
      for (int k = 0; k < 8; k++) { const int v = 4 * p.v.coef[(size_t) (oy0 + r) * 8 + k]; if (v < -128 || v > 127) t.vx4[rt] = 0; }
    const int scale = t.vx4[rt] ? 4 : 1;
    uint8_t *b = t.vband.data () + (size_t) rt * TC_VB_BYTES;
    for (int r = 0; r < TC_VROWS; r++) {
      b[tc_pos (r, TC_N, TC_VB_LBO)] = (uint8_t) (t.vx4[rt] ? 127 : 32);
      if (t.vx4[rt]) b[tc_pos (r, TC_N + 1, TC_VB_LBO)] = 1;
      if (r >= TC_TH || oy0 + r >= oh) continue;
      for (int k = 0; k < 8; k++) {
        const int tap = p.v.coef[(size_t) (oy0 + r) * 8 + k];
        if (!tap) continue;
        const int li = (int) p.v.offset[oy0 + r] + k - R0;
        if (li < 0 || li >= 2 * TC_TH + 6 || (int) p.v.offset[oy0 + r] + k >= ih) return t;
        b[tc_pos (r, li, TC_VB_LBO)] = (uint8_t) (int8_t) (scale * tap);
      }
    }
  }
  t.ok = true;
  return t;
}

struct L2tcState {
  uint8_t *d_band = nullptr, *d_vband = nullptr, *d_hx4 = nullptr, *d_vx4 = nullptr;
  L2tcDev dev;
  bool ready = false;
};

inline int prepare_l2tc (const L2tcTables & t, L2tcState * st)
{
  int rc;
  if (!t.ok) return B200_OK;
  if ((rc = upload (&st->d_band, t.band.data (), t.band.size ())) != B200_OK) return rc;
  if ((rc = upload (&st->d_vband, t.vband.data (), t.vband.size ())) != B200_OK) return rc;
  if ((rc = upload (&st->d_hx4, t.hx4.data (), t.hx4.size ())) != B200_OK) return rc;
  if ((rc = upload (&st->d_vx4, t.vx4.data (), t.vx4.size ())) != B200_OK) return rc;
  st->dev.band = st->d_band; st->dev.vband = st->d_vband; st->dev.hx4 = st->d_hx4; st->dev.vx4 = st->d_vx4;
  st->dev.strips = t.strips; st->dev.row_tiles = t.row_tiles;
  st->dev.magic_rt = 0xffffffffu / (unsigned) t.row_tiles + 1u;
  st->ready = true;
  return B200_OK;
}

inline int launch_l2tc (const VcsDev & d, const L2tcState & st, const VcsBatch & batch, int n, cudaStream_t stream,
    unsigned *dbg = nullptr, int dbg_mode = 0)
{
  int rc;
  int dev = 0;
  B200_CUDA_TRY (cudaGetDevice (&dev));
  const int tiles = st.dev.strips * st.dev.row_tiles * n;
  const int grid = min (tiles, sm_count (dev));
  if (dbg_mode == 1) {
    if ((rc = allow_max_dyn_smem (vcs_l2tc_kernel<1>)) != B200_OK) return rc;
    vcs_l2tc_kernel<1> <<<grid, TC_THREADS2, TC2_SMEM, stream>>> (d, st.dev, batch, n, dbg);
  } else if (dbg_mode == 2) {
    if ((rc = allow_max_dyn_smem (vcs_l2tc_kernel<2>)) != B200_OK) return rc;
    vcs_l2tc_kernel<2> <<<grid, TC_THREADS2, TC2_SMEM, stream>>> (d, st.dev, batch, n, dbg);
  } else {
    static bool attr_done[16] = {false};
    if (!attr_done[dev & 15]) {
      if ((rc = allow_max_dyn_smem (vcs_l2tc_kernel<0>)) != B200_OK) return rc;
      attr_done[dev & 15] = true;
    }
    vcs_l2tc_kernel<0> <<<grid, TC_THREADS2, TC2_SMEM, stream>>> (d, st.dev, batch, n, nullptr);
  }
  B200_CUDA_TRY (cudaGetLastError ());
  return B200_OK;
}

}  // namespace b200
