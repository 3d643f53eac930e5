// gstreamer_b200/csrc/vcs_planes.cuh — YUV -> same YUV family (NV12->NV12, NV21->NV21, I420/YV12 ->
// I420/YV12): the reference's plane-scaling fast path (product code, sm_100a).
//
//   convert_scale_planes            gst-libs/gst/video/video-converter.c:7757-7769
//   setup_scale                     :7958-8245  (which kernel each plane gets; built on the host, vcs_plan.cpp)
//   convert_plane_{h,v,hv}_halve    :7399-7680  -> video_orc_planar_chroma_444_422 / _422_420 / _444_420
//   convert_plane_{h,v,hv}_double   :7348-7615  -> replication
//   gst_video_scaler_2d             gst-libs/gst/video/video-scaler.c:1451-1640 with video_scale_h_near_u8/_u16,
//                                   video_scale_h_2tap_1u8 (ldreslinb), video_scale_h_ntap_u8, video_scale_v_*_u8
//
// Correctness-first kernel: one thread per output byte evaluates its separable filter directly from the source plane
// (the first pass is re-evaluated per tap of the second instead of being staged), every intermediate rounded to
// 8 bits exactly where the reference rounds.  One launch covers all planes of all frames of a batch.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "vcs_device.h"
#include "vcs_kernels.cuh"
#include "vcs_plan.h"

namespace b200 {

struct PlaneAxisDev {
  const uint32_t *offset;
  const int16_t *coef;
  int mode, n_taps;              // PassMode; taps per output
};

struct PlaneDev {
  unsigned long long src_off, dst_off;
  int sstride, dstride, iw, ih, ow, oh, ne, mode, h_first;
  PlaneAxisDev h, v;
};

struct PlanesParams {
  PlaneDev pl[3];
  int n_planes;
};

// one pixel component of one line through the horizontal scaler
__device__ __forceinline__ int plane_h (const PlaneDev & Q, const uint8_t *line, int x, int c)
{
  const PlaneAxisDev & A = Q.h;
  if (A.mode == PASS_COPY) return line[(int) A.offset[x] * Q.ne + c];
  if (A.mode == PASS_2TAP) {                                       // ldreslinb; the second tap is unused (f == 0) at the edge
    const int i0 = (int) A.offset[x], f = A.coef[x];
    const int i1 = min (i0 + 1, Q.iw - 1);
    return lerp_h_u8 (line[i0 * Q.ne + c], line[i1 * Q.ne + c], f);
  }
  const int16_t *t = A.coef + (size_t) x * A.n_taps;
  const uint8_t *s = line + (int) A.offset[x] * Q.ne + c;
  int acc = 0;
  for (int k = 0; k < A.n_taps; k++) acc += (int) s[k * Q.ne] * (int) t[k];
  return fir_round_u8 (acc);
}

__global__ void __launch_bounds__ (256)
vcs_planes_kernel (const PlanesParams P, const VcsBatch frames)
{
  const int plane = blockIdx.z % P.n_planes, frame = blockIdx.z / P.n_planes;
  const PlaneDev & Q = P.pl[plane];
  const int xb = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (xb >= Q.ow * Q.ne || y >= Q.oh) return;
  const uint8_t *__restrict__ src = frames.in[frame] + Q.src_off;
  uint8_t *__restrict__ dst = frames.out[frame] + Q.dst_off;
  const int x = xb / Q.ne, c = xb - x * Q.ne;
  int v;
  switch (Q.mode) {
    case PM_COPY:
      v = src[(size_t) y * Q.sstride + xb];
      break;
    case PM_HALVE_V:                                               // avgub of the two lines
      v = (src[(size_t) (2 * y) * Q.sstride + xb] + src[(size_t) (2 * y + 1) * Q.sstride + xb] + 1) >> 1;
      break;
    case PM_HALVE_H:
      v = (src[(size_t) y * Q.sstride + 2 * xb] + src[(size_t) y * Q.sstride + 2 * xb + 1] + 1) >> 1;
      break;
    case PM_HALVE_HV: {                                            // vertical averages first, then the pair
      const uint8_t *a = src + (size_t) (2 * y) * Q.sstride + 2 * xb, *b = a + Q.sstride;
      const int t1 = (a[0] + b[0] + 1) >> 1, t2 = (a[1] + b[1] + 1) >> 1;
      v = (t1 + t2 + 1) >> 1;
      break;
    }
    case PM_DOUBLE:
      v = src[(size_t) (Q.ih == Q.oh ? y : y >> 1) * Q.sstride + (Q.iw == Q.ow ? xb : xb >> 1)];
      break;
    default: {
      const PlaneAxisDev & V = Q.v;
      const int r0 = (int) V.offset[y];
      if (Q.h_first) {
        // h-scaled lines r0 .. r0 + n_taps - 1 (rounded to bytes), then the vertical function
        if (V.mode == PASS_COPY) v = plane_h (Q, src + (size_t) r0 * Q.sstride, x, c);
        else if (V.mode == PASS_2TAP)
          v = lerp_v_u8 (plane_h (Q, src + (size_t) r0 * Q.sstride, x, c),
              plane_h (Q, src + (size_t) (r0 + 1) * Q.sstride, x, c), V.coef[y]);
        else {
          const int16_t *t = V.coef + (size_t) y * V.n_taps;
          int acc = 0;
          for (int k = 0; k < V.n_taps; k++) acc += plane_h (Q, src + (size_t) (r0 + k) * Q.sstride, x, c) * (int) t[k];
          v = fir_round_u8 (acc);
        }
      } else {
        // vertical first: the v-scaled bytes of the columns the horizontal window reads, then the horizontal function
        const PlaneAxisDev & H = Q.h;
        auto vcol = [&] (int col) -> int {
          const uint8_t *s = src + (size_t) r0 * Q.sstride + col * Q.ne + c;
          if (V.mode == PASS_COPY) return s[0];
          if (V.mode == PASS_2TAP) return lerp_v_u8 (s[0], s[Q.sstride], V.coef[y]);
          const int16_t *t = V.coef + (size_t) y * V.n_taps;
          int acc = 0;
          for (int k = 0; k < V.n_taps; k++) acc += (int) s[(size_t) k * Q.sstride] * (int) t[k];
          return fir_round_u8 (acc);
        };
        const int i0 = (int) H.offset[x];
        if (H.mode == PASS_COPY) v = vcol (i0);
        else if (H.mode == PASS_2TAP) v = lerp_h_u8 (vcol (i0), vcol (min (i0 + 1, Q.iw - 1)), H.coef[x]);
        else {
          const int16_t *t = H.coef + (size_t) x * H.n_taps;
          int acc = 0;
          for (int k = 0; k < H.n_taps; k++) acc += vcol (i0 + k) * (int) t[k];
          v = fir_round_u8 (acc);
        }
      }
    }
  }
  dst[(size_t) y * Q.dstride + xb] = (uint8_t) v;
}

struct PlanesState {
  uint32_t *d_off[3][2] = {{nullptr}};
  int16_t *d_coef[3][2] = {{nullptr}};
  PlanesParams params;
  bool ready = false;
};

inline int prepare_planes (const VcsPlan & p, PlanesState * st)
{
  memset (&st->params, 0, sizeof (st->params));
  st->params.n_planes = p.n_planes;
  for (int i = 0; i < p.n_planes; i++) {
    const PlanePlan & q = p.planes[i];
    PlaneDev & d = st->params.pl[i];
    d.src_off = p.in.offset[q.src_plane]; d.dst_off = p.out.offset[i];
    d.sstride = p.in.stride[q.src_plane]; d.dstride = p.out.stride[i];
    d.iw = q.iw; d.ih = q.ih; d.ow = q.ow; d.oh = q.oh; d.ne = q.ne; d.mode = q.mode; d.h_first = q.h_first ? 1 : 0;
    if (q.mode != PM_SCALE) continue;
    const AxisPlan *ax[2] = {&q.h, &q.v};
    PlaneAxisDev *dv[2] = {&d.h, &d.v};
    for (int a = 0; a < 2; a++) {
      int s;
      if ((s = upload (&st->d_off[i][a], ax[a]->offset.data (), ax[a]->offset.size ())) != B200_OK) return s;
      if ((s = upload (&st->d_coef[i][a], ax[a]->coef.data (), ax[a]->coef.size ())) != B200_OK) return s;
      dv[a]->offset = st->d_off[i][a]; dv[a]->coef = st->d_coef[i][a];
      dv[a]->mode = ax[a]->mode; dv[a]->n_taps = ax[a]->n_taps;
    }
  }
  st->ready = true;
  return B200_OK;
}

inline void free_planes (PlanesState * st)
{
  for (int i = 0; i < 3; i++)
    for (int a = 0; a < 2; a++) { cudaFree (st->d_off[i][a]); cudaFree (st->d_coef[i][a]); }
}

inline int launch_planes (const PlanesState & st, const VcsBatch & batch, int n, cudaStream_t stream)
{
  int wmax = 0, hmax = 0;
  for (int i = 0; i < st.params.n_planes; i++) {
    wmax = max (wmax, st.params.pl[i].ow * st.params.pl[i].ne);
    hmax = max (hmax, st.params.pl[i].oh);
  }
  const dim3 grid ((wmax + 63) / 64, (hmax + 3) / 4, st.params.n_planes * n);
  vcs_planes_kernel <<<grid, 256, 0, stream>>> (st.params, batch);
  B200_CUDA_TRY (cudaGetLastError ());
  return B200_OK;
}

}  // namespace b200
