/* oracle/oracle_video.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Restatement of the videoconvertscale hot path for 4:2:0 semi-planar YUV -> packed
 * 8-bit RGB with scaling, as whole-frame passes in the exact stage order and
 * integer arithmetic of the reference's pull chain:
 *
 *   unpack (nearest chroma)          video-format.c:1593-1641 (NV12), :1679+ (NV21)
 *   chroma upsample h then v         video-converter.c:2991-3021, video-chroma.c:309-327, :687-699
 *   [downscale h/v]  -> matrix -> [upscale h/v]     video-converter.c:1685-1718, :2509-2539
 *   h scale                          video-scaler.c:596-618 (2-tap), :621-760 (n-tap), :462-580 (nearest)
 *   v scale                          video-scaler.c:846-879 (2-tap), :923-1072 (4/n-tap), :828-844 (nearest)
 *   AYUV->ARGB matrix                video-converter.c:1209-1216; video-orc.orc:1634-1688
 *   pack                             video-format.c:1445-1452 (+ :1473-1520 siblings)
 *
 * Defined for n-threads=1 (the element default, gstvideoconvertscale.c:144).
 * Pinned byte-for-byte against oracle/_ref (the reference's own sources) by
 * tests/test_oracle_vs_ref.py.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CLAMPI(x,lo,hi) ((x) > (hi) ? (hi) : ((x) < (lo) ? (lo) : (x)))
#define ROUND_UP_4(n) (((n) + 3) & ~3)
#define ROUND_UP_2(n) (((n) + 1) & ~1)

/* ======================================================================= taps */

/* video-resampler.c:142-202 */
static double
sinc_ (double x)
{
  if (x == 0)
    return 1;
  return sin (M_PI * x) / (M_PI * x);
}

static double
envelope_ (double x)
{
  if (x <= -1 || x >= 1)
    return 0;
  return sinc_ (x);
}

typedef struct
{
  int method;
  double fx, ex, b, c, sharpen;
} TapParams;

static double
get_tap (const TapParams * p, int l, int xi, double x)
{
  int xl = xi + l;
  switch (p->method) {
    case ORC_RS_NEAREST:       /* video-resampler.c:156-160 */
      return 1.0;
    case ORC_RS_LINEAR:{       /* :162-176 */
      double a = fabs (x - xl) * p->fx;
      return a < 1.0 ? 1.0 - a : 0.0;
    }
    case ORC_RS_CUBIC:{        /* :178-201 */
      double a = fabs (x - xl) * p->fx, a2 = a * a, a3 = a2 * a, b = p->b, c = p->c;
      if (a <= 1.0)
        return ((12.0 - 9.0 * b - 6.0 * c) * a3 + (-18.0 + 12.0 * b + 6.0 * c) * a2 +
            (6.0 - 2.0 * b)) / 6.0;
      else if (a <= 2.0)
        return ((-b - 6.0 * c) * a3 + (6.0 * b + 30.0 * c) * a2 + (-12.0 * b - 48.0 * c) * a +
            (8.0 * b + 24.0 * c)) / 6.0;
      return 0.0;
    }
    case ORC_RS_SINC:          /* :203-208 */
      return sinc_ ((x - xl) * p->fx);
    case ORC_RS_LANCZOS:       /* :210-216 */
    default:
      return (sinc_ ((x - xl) * p->fx) - p->sharpen) * envelope_ ((x - xl) * p->ex);
  }
}

int
oracle_resampler_taps (const OracleResamplerOpts * o, int in_size, int out_size,
    uint32_t * offset, double *taps)
{
  /* gst_video_resampler_init, video-resampler.c:343-429 */
  TapParams p;
  int max_taps = o->max_taps_opt > 0 ? o->max_taps_opt : 128;
  int n_taps = o->n_taps_req;
  double envelope = 2.0, scale_factor, dx, corr;
  int j, l, tap_offs;

  if (in_size <= 0 || out_size <= 0)
    return -1;
  p.method = o->method;
  p.sharpen = o->sharpen;
  p.b = o->cubic_b;
  p.c = o->cubic_c;
  scale_factor = in_size / (double) out_size;
  if (scale_factor > 1.0)
    p.fx = (1.0 / scale_factor) * o->sharpness;
  else
    p.fx = 1.0 * o->sharpness;

  if (n_taps > max_taps)
    n_taps = max_taps;
  switch (o->method) {
    case ORC_RS_NEAREST:
      envelope = o->envelope;
      if (n_taps == 0)
        n_taps = 1;
      break;
    case ORC_RS_LINEAR:
      envelope = 1.0;
      break;
    case ORC_RS_CUBIC:
      envelope = 2.0;
      break;
    default:
      envelope = o->envelope;
      break;
  }
  if (n_taps == 0) {
    dx = ceil (2.0 * envelope / p.fx);
    n_taps = (int) CLAMPI (dx, 0, max_taps);
  }
  p.fx = 2.0 * envelope / n_taps;
  p.ex = 2.0 / n_taps;
  if (n_taps > in_size)
    n_taps = in_size;

  /* resampler_calculate_taps, video-resampler.c:204-288 */
  tap_offs = (n_taps - 1) / 2;
  corr = (n_taps == 1 ? 0.0 : 0.5);
  for (j = 0; j < out_size; j++) {
    double ox, x, weight = 0, *t = taps + (size_t) j * n_taps;
    int xi;
    ox = (0.5 + (double) j - 0.0) / out_size;
    x = ox * (double) in_size - corr;
    x = CLAMPI (x, 0, in_size - 1);
    xi = (int) floor (x - tap_offs);
    offset[j] = (uint32_t) xi;
    for (l = 0; l < n_taps; l++) {
      t[l] = get_tap (&p, l, xi, x);
      weight += t[l];
    }
    for (l = 0; l < n_taps; l++)
      t[l] /= weight;
    if (xi < 0) {
      int sh = -xi;
      for (l = 0; l < sh; l++)
        t[sh] += t[l];
      for (l = 0; l < n_taps - sh; l++)
        t[l] = t[sh + l];
      for (; l < n_taps; l++)
        t[l] = 0;
      offset[j] += sh;
    }
    if (xi > in_size - n_taps) {
      int sh = xi - (in_size - n_taps);
      for (l = 0; l < sh; l++)
        t[n_taps - sh - 1] += t[n_taps - sh + l];
      for (l = 0; l < n_taps - sh; l++)
        t[n_taps - 1 - l] = t[n_taps - 1 - sh - l];
      for (l = 0; l < sh; l++)
        t[l] = 0;
      offset[j] -= sh;
    }
  }
  return n_taps;
}

int
oracle_quantize_taps (const double *src, int16_t * dst, int n, int precision)
{
  /* resampler_convert_coeff, video-scaler.c:338-388: bisection on the rounding
   * bias until the integer taps sum to 1<<precision (may fail -> keeps last try) */
  double multiplier = (double) (1 << precision);
  double l_offset = 0.0, h_offset = 1.0, offset = 0.5;
  int i, j, exact = 0;
  for (i = 0; i < 64; i++) {
    int sum = 0;
    for (j = 0; j < n; j++) {
      int16_t tap = (int16_t) floor (offset + src[j] * multiplier);
      dst[j] = tap;
      sum += tap;
    }
    if (sum == (1 << precision)) {
      exact = 1;
      break;
    }
    if (l_offset == h_offset)
      break;
    if (sum < (1 << precision)) {
      if (offset > l_offset)
        l_offset = offset;
      offset += (h_offset - l_offset) / 2;
    } else {
      if (offset < h_offset)
        h_offset = offset;
      offset -= (h_offset - l_offset) / 2;
    }
  }
  return exact;
}

/* ===================================================================== layout */

static int
fmt_is_rgb (int f)
{
  return f >= ORC_FMT_RGBx && f <= ORC_FMT_ABGR;
}

/* 4:2:2 / 4:4:4 inputs (packed YUY2 / UYVY / YVYU, planar Y42B / Y444): where the samples of a line live.  Chroma is
 * sub-sampled horizontally by hshift and not at all vertically. */
static int
fmt_is_422_444 (int f)
{
  return f == ORC_FMT_YUY2 || f == ORC_FMT_UYVY || f == ORC_FMT_YVYU || f == ORC_FMT_Y42B || f == ORC_FMT_Y444;
}

int
oracle_vcs_default_desc (OracleVcsDesc * d, int in_format, int in_w, int in_h,
    int out_format, int out_w, int out_h, int method, int max_taps_opt)
{
  memset (d, 0, sizeof (*d));
  if (in_format != ORC_FMT_NV12 && in_format != ORC_FMT_NV21 && in_format != ORC_FMT_I420 &&
      in_format != ORC_FMT_YV12 && !fmt_is_rgb (in_format) && !fmt_is_422_444 (in_format))
    return -1;
  d->in_format = in_format;
  d->in_width = in_w;
  d->in_height = in_h;
  d->in_stride[0] = ROUND_UP_4 (in_w);
  d->in_offset[0] = 0;
  if (fmt_is_rgb (in_format)) {
    d->in_stride[0] = in_w * 4;         /* video-info.c:890-894 */
  } else if (in_format == ORC_FMT_YUY2 || in_format == ORC_FMT_UYVY || in_format == ORC_FMT_YVYU) {
    d->in_stride[0] = ROUND_UP_4 (in_w * 2);    /* :882-889 */
  } else if (in_format == ORC_FMT_Y42B) {       /* :1020-1029 */
    d->in_stride[1] = d->in_stride[2] = ((in_w + 7) & ~7) / 2;
    d->in_offset[1] = (size_t) d->in_stride[0] * in_h;
    d->in_offset[2] = d->in_offset[1] + (size_t) d->in_stride[1] * in_h;
  } else if (in_format == ORC_FMT_Y444) {       /* :1030-1041 */
    d->in_stride[1] = d->in_stride[2] = d->in_stride[0];
    d->in_offset[1] = (size_t) d->in_stride[0] * in_h;
    d->in_offset[2] = d->in_offset[1] * 2;
  } else if (in_format == ORC_FMT_I420 || in_format == ORC_FMT_YV12) {
    /* video-info.c:997-1009 (YV12: same planes, 1 and 2 swapped in the format description) */
    d->in_stride[1] = d->in_stride[2] = ROUND_UP_4 (ROUND_UP_2 (in_w) / 2);
    d->in_offset[1] = (size_t) d->in_stride[0] * ROUND_UP_2 (in_h);
    d->in_offset[2] = d->in_offset[1] + (size_t) d->in_stride[1] * (ROUND_UP_2 (in_h) / 2);
  } else {
    /* video-info.c:1053-1063 */
    d->in_stride[1] = d->in_stride[0];
    d->in_offset[1] = (size_t) d->in_stride[0] * ROUND_UP_2 (in_h);
  }
  /* video-info.c:165-185, :211-225 */
  d->in_matrix = in_h > 576 ? ORC_CM_BT709 : ORC_CM_BT601;
  d->in_range = ORC_RANGE_16_235;
  d->in_chroma_site = in_h > 576 ? ORC_SITE_H_COSITED : ORC_SITE_NONE;
  if (fmt_is_rgb (in_format)) {
    d->in_matrix = ORC_CM_RGB;
    d->in_range = ORC_RANGE_0_255;
    d->in_chroma_site = 0;
  }
  d->out_format = out_format;
  d->out_width = out_w;
  d->out_height = out_h;
  d->out_stride[0] = out_w * 4;         /* video-info.c:890-894 */
  if (out_format == ORC_FMT_I420 || out_format == ORC_FMT_YV12) {
    d->out_stride[0] = ROUND_UP_4 (out_w);
    d->out_stride[1] = d->out_stride[2] = ROUND_UP_4 (ROUND_UP_2 (out_w) / 2);
    d->out_offset[1] = (size_t) d->out_stride[0] * ROUND_UP_2 (out_h);
    d->out_offset[2] = d->out_offset[1] + (size_t) d->out_stride[1] * (ROUND_UP_2 (out_h) / 2);
  } else if (out_format == ORC_FMT_NV12 || out_format == ORC_FMT_NV21) {
    d->out_stride[0] = d->out_stride[1] = ROUND_UP_4 (out_w);
    d->out_offset[1] = (size_t) d->out_stride[0] * ROUND_UP_2 (out_h);
  }
  d->rs.method = method;
  d->rs.max_taps_opt = max_taps_opt;
  d->rs.envelope = 2.0;
  d->rs.sharpness = 1.0;
  d->rs.sharpen = 0.0;
  d->rs.cubic_b = 1.0 / 3.0;
  d->rs.cubic_c = 1.0 / 3.0;
  return 0;
}

size_t
oracle_vcs_in_size (const OracleVcsDesc * d)
{
  if (fmt_is_rgb (d->in_format) || d->in_format == ORC_FMT_YUY2 || d->in_format == ORC_FMT_UYVY || d->in_format == ORC_FMT_YVYU)
    return d->in_offset[0] + (size_t) d->in_stride[0] * d->in_height;
  if (d->in_format == ORC_FMT_Y42B || d->in_format == ORC_FMT_Y444)
    return d->in_offset[2] + (size_t) d->in_stride[2] * d->in_height;
  if (d->in_format == ORC_FMT_I420 || d->in_format == ORC_FMT_YV12)
    return d->in_offset[2] + (size_t) d->in_stride[2] * (ROUND_UP_2 (d->in_height) / 2);
  return d->in_offset[1] + (size_t) d->in_stride[1] * (ROUND_UP_2 (d->in_height) / 2);
}

size_t
oracle_vcs_out_size (const OracleVcsDesc * d)
{
  if (d->out_format == ORC_FMT_I420 || d->out_format == ORC_FMT_YV12)
    return d->out_offset[2] + (size_t) d->out_stride[2] * (ROUND_UP_2 (d->out_height) / 2);
  if (d->out_format == ORC_FMT_NV12 || d->out_format == ORC_FMT_NV21)
    return d->out_offset[1] + (size_t) d->out_stride[1] * (ROUND_UP_2 (d->out_height) / 2);
  return d->out_offset[0] + (size_t) d->out_stride[0] * d->out_height;
}

/* ===================================================================== matrix */

typedef struct
{
  double dm[4][4];
} Mat;

static void
mat_identity (Mat * m)
{
  int i, j;
  for (i = 0; i < 4; i++)
    for (j = 0; j < 4; j++)
      m->dm[i][j] = (i == j);
}

/* dst = a * b (video-converter.c:925-941) */
static void
mat_mul (Mat * dst, const Mat * a, const Mat * b)
{
  Mat t;
  int i, j, k;
  for (i = 0; i < 4; i++)
    for (j = 0; j < 4; j++) {
      double x = 0;
      for (k = 0; k < 4; k++)
        x += a->dm[i][k] * b->dm[k][j];
      t.dm[i][j] = x;
    }
  *dst = t;
}

static void
mat_offset (Mat * m, double a1, double a2, double a3)
{
  Mat a;
  mat_identity (&a);
  a.dm[0][3] = a1;
  a.dm[1][3] = a2;
  a.dm[2][3] = a3;
  mat_mul (m, &a, m);
}

static void
mat_scale (Mat * m, double a1, double a2, double a3)
{
  Mat a;
  mat_identity (&a);
  a.dm[0][0] = a1;
  a.dm[1][1] = a2;
  a.dm[2][2] = a3;
  mat_mul (m, &a, m);
}

int
oracle_vcs_matrix (const OracleVcsDesc * d, int p[5], int im[4][4])
{
  /* chain_convert (video-converter.c:1720-1868) for 8-bit YUV in, 8-bit RGB out,
   * gamma/primaries modes NONE: identity -> compute_matrix_to_RGB (:1373-1404)
   * -> compute_matrix_to_YUV (:1406-1442, RGB out: range scaling only)
   * -> prepare_matrix (:1324-1370): x256, rint. */
  Mat m;
  double Kr, Kb, Kg;
  int offset[3], scale[3], i, j;
  mat_identity (&m);
  /* gst_video_color_range_offsets for AYUV 8 bit (video-color.c:204-252) */
  if (d->in_range == ORC_RANGE_16_235) {
    offset[0] = 16; scale[0] = 219;
    offset[1] = offset[2] = 128; scale[1] = scale[2] = 224;
  } else {
    offset[0] = 0; scale[0] = 255;
    offset[1] = offset[2] = 128; scale[1] = scale[2] = 255;
  }
  mat_offset (&m, -offset[0], -offset[1], -offset[2]);
  mat_scale (&m, 1 / ((float) scale[0]), 1 / ((float) scale[1]), 1 / ((float) scale[2]));
  switch (d->in_matrix) {      /* video-color.c:423-459 */
    case ORC_CM_FCC: Kr = 0.30; Kb = 0.11; break;
    case ORC_CM_BT709: Kr = 0.2126; Kb = 0.0722; break;
    case ORC_CM_BT601: Kr = 0.2990; Kb = 0.1140; break;
    case ORC_CM_SMPTE240M: Kr = 0.212; Kb = 0.087; break;
    case ORC_CM_BT2020: Kr = 0.2627; Kb = 0.0593; break;
    default: return -1;
  }
  Kg = 1.0 - Kr - Kb;
  {                             /* color_matrix_YCbCr_to_RGB, :1027-1040 */
    Mat k = { {{1., 0., 2 * (1 - Kr), 0.},
        {1., -2 * Kb * (1 - Kb) / Kg, -2 * Kr * (1 - Kr) / Kg, 0.},
        {1., 2 * (1 - Kb), 0., 0.},
        {0., 0., 0., 1.}} };
    mat_mul (&m, &k, &m);
  }
  /* out: ARGB 0-255 full range */
  mat_scale (&m, (float) 255, (float) 255, (float) 255);
  mat_offset (&m, 0, 0, 0);
  mat_scale (&m, 256.0f, 256.0f, 256.0f);
  for (i = 0; i < 4; i++)
    for (j = 0; j < 4; j++)
      im[i][j] = (int) rint (m.dm[i][j]);
  /* is_ayuv_to_rgb_matrix (:1218-1228) must hold for the fast path */
  if (im[0][0] != im[1][0] || im[1][0] != im[2][0] || im[0][1] != 0 || im[2][2] != 0)
    return -2;
  p[0] = im[0][0];
  p[1] = im[0][2];
  p[2] = im[2][1];
  p[3] = im[1][1];
  p[4] = im[1][2];
  return 0;
}

static int
kr_kb (int matrix, double *Kr, double *Kb)
{
  switch (matrix) {             /* video-color.c:423-459 */
    case ORC_CM_FCC: *Kr = 0.30; *Kb = 0.11; return 0;
    case ORC_CM_BT709: *Kr = 0.2126; *Kb = 0.0722; return 0;
    case ORC_CM_BT601: *Kr = 0.2990; *Kb = 0.1140; return 0;
    case ORC_CM_SMPTE240M: *Kr = 0.212; *Kb = 0.087; return 0;
    case ORC_CM_BT2020: *Kr = 0.2627; *Kb = 0.0593; return 0;
    default: return -1;
  }
}

/* colorimetry of a 4:2:0 output fed by packed RGB: explicit, else the caps defaults of the output size */
static void
rgb_in_out_colorimetry (const OracleVcsDesc * d, int *matrix, int *range, int *site)
{
  *matrix = d->out_matrix ? d->out_matrix : (d->out_height > 576 ? ORC_CM_BT709 : ORC_CM_BT601);
  *range = d->out_range ? d->out_range : ORC_RANGE_16_235;
  *site = d->out_chroma_site ? d->out_chroma_site : (d->out_height > 576 ? ORC_SITE_H_COSITED : ORC_SITE_NONE);
}

int
oracle_vcs_matrix_rgb2yuv (const OracleVcsDesc * d, int im[4][4])
{
  /* chain_convert (video-converter.c:1720-1868), 8-bit ARGB in, 8-bit AYUV out, no gamma / primaries remap:
   * identity -> compute_matrix_to_RGB (:1373-1404; RGB input: only the range normalisation of the unpack format,
   * gst_video_color_range_offsets video-color.c:204-252) -> compute_matrix_to_YUV (:1406-1442): color_matrix_RGB_to_YCbCr
   * (:1037-1066) with the OUTPUT matrix, output range scale + offset -> prepare_matrix (:1324-1370): x256, rint */
  Mat m;
  double Kr, Kb, Kg, x;
  int i, j, matrix, range, site, t;
  static const uint8_t corner[8][3] = { {0, 0, 0}, {0, 0, 255}, {0, 255, 0}, {0, 255, 255}, {255, 0, 0}, {255, 0, 255},
    {255, 255, 0}, {255, 255, 255} };
  rgb_in_out_colorimetry (d, &matrix, &range, &site);
  mat_identity (&m);
  if (d->in_range == ORC_RANGE_16_235) {
    mat_offset (&m, -16, -16, -16);
    mat_scale (&m, 1 / ((float) 219), 1 / ((float) 219), 1 / ((float) 219));
  } else {
    mat_offset (&m, 0, 0, 0);
    mat_scale (&m, 1 / ((float) 255), 1 / ((float) 255), 1 / ((float) 255));
  }
  if (kr_kb (matrix, &Kr, &Kb))
    return -1;
  Kg = 1.0 - Kr - Kb;
  {
    Mat k;
    mat_identity (&k);
    k.dm[0][0] = Kr; k.dm[0][1] = Kg; k.dm[0][2] = Kb;
    x = 1 / (2 * (1 - Kb));
    k.dm[1][0] = -x * Kr; k.dm[1][1] = -x * Kg; k.dm[1][2] = x * (1 - Kb);
    x = 1 / (2 * (1 - Kr));
    k.dm[2][0] = x * (1 - Kr); k.dm[2][1] = -x * Kg; k.dm[2][2] = -x * Kb;
    mat_mul (&m, &k, &m);
  }
  if (range == ORC_RANGE_16_235) {
    mat_scale (&m, (float) 219, (float) 224, (float) 224);
    mat_offset (&m, 16, 128, 128);
  } else {
    mat_scale (&m, (float) 255, (float) 255, (float) 255);
    mat_offset (&m, 0, 128, 128);
  }
  mat_scale (&m, 256.0f, 256.0f, 256.0f);
  for (i = 0; i < 4; i++)
    for (j = 0; j < 4; j++)
      im[i][j] = (int) rint (m.dm[i][j]);
  /* is_no_clip_matrix (:1262-1300) -> video_converter_matrix8_table; a clipping matrix would run video_orc_matrix8,
   * whose SIMD program and C backup disagree: not restated */
  for (t = 0; t < 8; t++)
    for (i = 0; i < 3; i++) {
      int v = (im[i][0] * corner[t][0] + im[i][1] * corner[t][1] + im[i][2] * corner[t][2] + im[i][3]) >> 8;
      if (v < 0 || v > 255)
        return -2;
    }
  return 0;
}

/* video_converter_matrix8_table (:1178-1200) with the tables of videoconvert_convert_init_tables (:1110-1134): the
 * three 16-bit fields of the packed 64-bit sums never borrow from each other for a no-clip matrix, so every component
 * is (row . (r,g,b) + offset) >> 8 */
static void
matrix_line_rgb2yuv (uint8_t * px, int n, int im[4][4])
{
  int i, c;
  for (i = 0; i < n; i++, px += 4) {
    const int r = px[1], g = px[2], b = px[3];
    for (c = 0; c < 3; c++)
      px[1 + c] = (uint8_t) ((im[c][0] * r + im[c][1] * g + im[c][2] * b + im[c][3]) >> 8);
  }
}

/* video_orc_convert_AYUV_ARGB, video-orc.orc:1634-1688, in place on one line */
static inline int16_t
splatbw (uint8_t b)
{
  return (int16_t) (uint16_t) ((b << 8) | b);
}

static inline int
sat_s8 (int v)
{
  return CLAMPI (v, -128, 127);
}

static void
matrix_line (uint8_t * px, int n, const int p[5])
{
  int i;
  int16_t p1 = (int16_t) p[0], p2 = (int16_t) p[1], p3 = (int16_t) p[2], p4 = (int16_t) p[3],
      p5 = (int16_t) p[4];
  for (i = 0; i < n; i++, px += 4) {
    uint8_t a = (uint8_t) (px[0] - 128), y = (uint8_t) (px[1] - 128);
    uint8_t u = (uint8_t) (px[2] - 128), v = (uint8_t) (px[3] - 128);
    int16_t wy = (int16_t) ((splatbw (y) * p1) >> 16);
    int16_t r = (int16_t) (wy + (int16_t) ((splatbw (v) * p2) >> 16));
    int16_t b = (int16_t) (wy + (int16_t) ((splatbw (u) * p3) >> 16));
    int16_t g = (int16_t) (wy + (int16_t) ((splatbw (u) * p4) >> 16));
    g = (int16_t) (g + (int16_t) ((splatbw (v) * p5) >> 16));
    px[0] = (uint8_t) (a + 128);
    px[1] = (uint8_t) (sat_s8 (r) + 128);
    px[2] = (uint8_t) (sat_s8 (g) + 128);
    px[3] = (uint8_t) (sat_s8 (b) + 128);
  }
}

/* ===================================================================== stages */

/* unpack one line to AYUV, chroma replicated (video-format.c:1595-1641) */
static void
unpack_line (const OracleVcsDesc * d, const uint8_t * in, int y, uint8_t * dst)
{
  const uint8_t *sy = in + d->in_offset[0] + (size_t) d->in_stride[0] * y;
  if (fmt_is_422_444 (d->in_format)) {
    /* unpack_YUY2 / _UYVY / _YVYU (video-format.c:155-197, :232-274, :...), unpack_Y42B (:1009-1050), unpack_Y444
     * (:1091-1104): the chroma sample of a pixel pair feeds both pixels (none shared for 4:4:4), every line has its own */
    int i;
    for (i = 0; i < d->in_width; i++) {
      int yy, u, v;
      switch (d->in_format) {
        case ORC_FMT_YUY2: yy = sy[2 * i]; u = sy[4 * (i >> 1) + 1]; v = sy[4 * (i >> 1) + 3]; break;
        case ORC_FMT_YVYU: yy = sy[2 * i]; v = sy[4 * (i >> 1) + 1]; u = sy[4 * (i >> 1) + 3]; break;
        case ORC_FMT_UYVY: yy = sy[2 * i + 1]; u = sy[4 * (i >> 1)]; v = sy[4 * (i >> 1) + 2]; break;
        case ORC_FMT_Y42B:
          yy = sy[i];
          u = in[d->in_offset[1] + (size_t) d->in_stride[1] * y + (i >> 1)];
          v = in[d->in_offset[2] + (size_t) d->in_stride[2] * y + (i >> 1)];
          break;
        default:
          yy = sy[i];
          u = in[d->in_offset[1] + (size_t) d->in_stride[1] * y + i];
          v = in[d->in_offset[2] + (size_t) d->in_stride[2] * y + i];
          break;
      }
      dst[i * 4 + 0] = 0xff; dst[i * 4 + 1] = (uint8_t) yy; dst[i * 4 + 2] = (uint8_t) u; dst[i * 4 + 3] = (uint8_t) v;
    }
    return;
  }
  if (d->in_format == ORC_FMT_I420 || d->in_format == ORC_FMT_YV12) {
    /* unpack_I420 -> video_orc_unpack_I420 (video-format.c:100-115, video-orc.orc:63-79): loadupdb = each
     * chroma sample feeds two pixels; YV12 keeps U in plane 2 */
    const int pu = d->in_format == ORC_FMT_YV12 ? 2 : 1, pv = 3 - pu;
    const uint8_t *su = in + d->in_offset[pu] + (size_t) d->in_stride[pu] * (y >> 1);
    const uint8_t *sv = in + d->in_offset[pv] + (size_t) d->in_stride[pv] * (y >> 1);
    int i;
    for (i = 0; i < d->in_width; i++) {
      dst[i * 4 + 0] = 0xff;
      dst[i * 4 + 1] = sy[i];
      dst[i * 4 + 2] = su[i >> 1];
      dst[i * 4 + 3] = sv[i >> 1];
    }
    return;
  }
  const uint8_t *suv = in + d->in_offset[1] + (size_t) d->in_stride[1] * (y >> 1);
  int x, ui = d->in_format == ORC_FMT_NV21 ? 1 : 0;
  for (x = 0; x < d->in_width; x++) {
    dst[x * 4 + 0] = 0xff;
    dst[x * 4 + 1] = sy[x];
    dst[x * 4 + 2] = suv[(x & ~1) + ui];
    dst[x * 4 + 3] = suv[(x & ~1) + (ui ^ 1)];
  }
}

/* packed RGB line -> ARGB: unpack_copy4 / unpack_BGRA / unpack_ABGR / unpack_RGBA (video-format.c:536-547, :1433-1443,
 * :1457-1470, :1489-1502); the x formats share them, their padding byte travels as alpha */
static void
unpack_line_rgb (const OracleVcsDesc * d, const uint8_t * in, int y, uint8_t * dst)
{
  const uint8_t *s = in + d->in_offset[0] + (size_t) d->in_stride[0] * y;
  int i;
  for (i = 0; i < d->in_width; i++, s += 4, dst += 4) {
    switch (d->in_format) {
      case ORC_FMT_BGRA: case ORC_FMT_BGRx: dst[0] = s[3]; dst[1] = s[2]; dst[2] = s[1]; dst[3] = s[0]; break;
      case ORC_FMT_RGBA: case ORC_FMT_RGBx: dst[0] = s[3]; dst[1] = s[0]; dst[2] = s[1]; dst[3] = s[2]; break;
      case ORC_FMT_ABGR: case ORC_FMT_xBGR: dst[0] = s[0]; dst[1] = s[3]; dst[2] = s[2]; dst[3] = s[1]; break;
      default: dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2]; dst[3] = s[3]; break;
    }
  }
}

/* video_chroma_up_h2_cs_u8 (video-chroma.c:687-699) / video_chroma_up_h2_u8 (:277-296) */
static void
chroma_h_line (uint8_t * p, int width, int cosited)
{
  int i;
  if (cosited) {
    for (i = 1; i < width - 1; i += 2) {
      p[2 + 4 * i] = (uint8_t) ((p[2 + 4 * (i - 1)] + p[2 + 4 * (i + 1)] + 1) >> 1);
      p[3 + 4 * i] = (uint8_t) ((p[3 + 4 * (i - 1)] + p[3 + 4 * (i + 1)] + 1) >> 1);
    }
  } else {
    int tr0, tr1 = p[2], tb0, tb1 = p[3];
    for (i = 1; i < width - 1; i += 2) {
      tr0 = tr1, tr1 = p[2 + 4 * (i + 1)];
      tb0 = tb1, tb1 = p[3 + 4 * (i + 1)];
      p[2 + 4 * i] = (uint8_t) ((3 * tr0 + tr1 + 2) >> 2);
      p[3 + 4 * i] = (uint8_t) ((3 * tb0 + tb1 + 2) >> 2);
      p[2 + 4 * (i + 1)] = (uint8_t) ((tr0 + 3 * tr1 + 2) >> 2);
      p[3 + 4 * (i + 1)] = (uint8_t) ((tb0 + 3 * tb1 + 2) >> 2);
    }
  }
}

typedef struct
{
  int n_taps;
  uint32_t *offset;
  double *taps;
  int16_t *taps_s16;            /* out_size * n_taps, lazily quantised */
  int in_size, out_size, inc;
} Scaler;

static int
scaler_init (Scaler * s, const OracleResamplerOpts * o, int in_size, int out_size)
{
  memset (s, 0, sizeof (*s));
  s->offset = malloc (sizeof (uint32_t) * out_size);
  s->taps = malloc (sizeof (double) * out_size * ORACLE_MAX_TAPS);
  s->n_taps = oracle_resampler_taps (o, in_size, out_size, s->offset, s->taps);
  s->in_size = in_size;
  s->out_size = out_size;
  /* video-scaler.c:254-257 */
  s->inc = out_size == 1 ? 0 : ((in_size - 1) << 16) / (out_size - 1) - 1;
  return s->n_taps > 0 ? 0 : -1;
}

static void
scaler_quantize (Scaler * s, int precision)
{
  int i;
  s->taps_s16 = malloc (sizeof (int16_t) * s->out_size * s->n_taps);
  for (i = 0; i < s->out_size; i++)
    oracle_quantize_taps (s->taps + (size_t) i * s->n_taps, s->taps_s16 + (size_t) i * s->n_taps,
        s->n_taps, precision);
}

static void
scaler_clear (Scaler * s)
{
  free (s->offset);
  free (s->taps);
  free (s->taps_s16);
}

static inline uint8_t
scale_round_u8 (int acc)
{
  /* addw 32; shrsw 6; convsuswb  (video-orc.orc:2474-2481): 16-bit wrap-around */
  int16_t w = (int16_t) (acc + 32);
  int v = w >> 6;
  return (uint8_t) CLAMPI (v, 0, 255);
}

/* horizontal pass over a packed 4x8-bit image: src (sw x h) -> dst (dw x h) */
static void
hscale_image (Scaler * s, const uint8_t * src, int sw, uint8_t * dst, int dw, int h)
{
  int x, y, c, k;
  if (s->n_taps > 2 && !s->taps_s16)
    scaler_quantize (s, 6);     /* SCALE_U8_LQ, video-scaler.c:632-637 */
  for (y = 0; y < h; y++) {
    const uint8_t *sl = src + (size_t) y * sw * 4;
    uint8_t *dl = dst + (size_t) y * dw * 4;
    for (x = 0; x < dw; x++) {
      if (s->n_taps == 1) {     /* video_scale_h_near_u32, :558-590 */
        memcpy (dl + 4 * x, sl + 4 * s->offset[x], 4);
      } else if (s->n_taps == 2) {      /* video_scale_h_2tap_4u8 -> ldreslinl, :609-618 */
        int tmp = x * s->inc, i0 = tmp >> 16, f = (tmp >> 8) & 0xff;
        int i1 = (i0 + 1 < sw) ? i0 + 1 : i0;   /* f==0 whenever i0+1 leaves the line */
        for (c = 0; c < 4; c++)
          dl[4 * x + c] = (uint8_t) ((sl[4 * i0 + c] * (256 - f) + sl[4 * i1 + c] * f) >> 8);
      } else {                  /* video_scale_h_ntap_u8, :621-760 */
        const int16_t *t = s->taps_s16 + (size_t) x * s->n_taps;
        for (c = 0; c < 4; c++) {
          int acc = 0;
          for (k = 0; k < s->n_taps; k++)
            acc += (int16_t) (sl[4 * (s->offset[x] + k) + c] * t[k]);
          dl[4 * x + c] = scale_round_u8 (acc);
        }
      }
    }
  }
}

/* vertical pass: src (w x sh) -> dst (w x dh) */
static void
vscale_image (Scaler * s, const uint8_t * src, int sh, uint8_t * dst, int dh, int w)
{
  int x, y, k, n = w * 4;
  (void) sh;
  if (s->n_taps >= 2 && !s->taps_s16)
    scaler_quantize (s, s->n_taps == 2 ? 8 : 6);       /* video-scaler.c:857, :937, :998 */
  for (y = 0; y < dh; y++) {
    uint8_t *dl = dst + (size_t) y * n;
    const uint8_t *s0 = src + (size_t) s->offset[y] * n;
    if (s->n_taps == 1) {       /* video_scale_v_near_u8, :828-835 */
      memcpy (dl, s0, n);
    } else if (s->n_taps == 2) {        /* video_orc_resample_v_2tap_u8_lq, video-orc.orc:2212-2228 */
      int16_t p1 = s->taps_s16[(size_t) y * 2 + 1];
      const uint8_t *s1 = s0 + n;
      for (x = 0; x < n; x++) {
        int16_t w2 = (int16_t) (s1[x] - s0[x]);
        w2 = (int16_t) (w2 * p1);
        w2 = (int16_t) (w2 + 128);
        dl[x] = (uint8_t) (((uint16_t) w2 >> 8) + s0[x]);
      }
    } else {                    /* video_scale_v_4tap_u8 / _ntap_u8, :923-1072 */
      const int16_t *t = s->taps_s16 + (size_t) y * s->n_taps;
      for (x = 0; x < n; x++) {
        int acc = 0;
        for (k = 0; k < s->n_taps; k++)
          acc += (int16_t) (s0[(size_t) k * n + x] * t[k]);
        dl[x] = scale_round_u8 (acc);
      }
    }
  }
}

static void
pack_line (int fmt, const uint8_t * argb, uint8_t * d, int n)
{
  /* video-format.c pack_BGRA :1445, pack_ABGR :1473, pack_RGBA :1505, pack_copy4 (ARGB),
   * RGBx/BGRx/xRGB/xBGR share the 4-byte shuffles of their alpha siblings */
  int i;
  for (i = 0; i < n; i++, argb += 4, d += 4) {
    uint8_t a = argb[0], r = argb[1], g = argb[2], b = argb[3];
    switch (fmt) {
      case ORC_FMT_BGRA: case ORC_FMT_BGRx: d[0] = b; d[1] = g; d[2] = r; d[3] = a; break;
      case ORC_FMT_RGBA: case ORC_FMT_RGBx: d[0] = r; d[1] = g; d[2] = b; d[3] = a; break;
      case ORC_FMT_ABGR: case ORC_FMT_xBGR: d[0] = a; d[1] = b; d[2] = g; d[3] = r; break;
      default: d[0] = a; d[1] = r; d[2] = g; d[3] = b; break;
    }
  }
}

/* chroma down-sampling, horizontal, in place on one AYUV line: video_chroma_down_h2_u8 -> video_orc_chroma_down_h2_u8
 * (video-chroma.c:398-409, video-orc.orc:2657-2671) averages each pixel pair into its first pixel; the co-sited
 * variant video_chroma_down_h2_cs_u8 (:742-764) filters 3-1 / 1-2-1 / 1-3 around every even pixel.  Only the even
 * pixels are read by the 4:2:0 pack functions. */
static void
chroma_down_h_line (uint8_t * p, int width, int cosited)
{
  int i, c;
  if (!cosited) {
    for (i = 0; i + 1 < width; i += 2)
      for (c = 2; c < 4; c++)
        p[c + 4 * i] = (uint8_t) ((p[c + 4 * i] + p[c + 4 * (i + 1)] + 1) >> 1);
    return;
  }
  if (width < 2)
    return;
  for (c = 2; c < 4; c++)
    p[c] = (uint8_t) ((3 * p[c] + p[c + 4] + 2) >> 2);
  for (i = 2; i < width - 2; i += 2)
    for (c = 2; c < 4; c++)
      p[c + 4 * i] = (uint8_t) ((p[c + 4 * (i - 1)] + 2 * p[c + 4 * i] + p[c + 4 * (i + 1)] + 2) >> 2);
  if (i < width)
    for (c = 2; c < 4; c++)
      p[c + 4 * i] = (uint8_t) ((p[c + 4 * (i - 1)] + 3 * p[c + 4 * i] + 2) >> 2);
}

/* pack one AYUV line into a 4:2:0 frame: pack_planar_420 (video-format.c:117-148), pack_NV12 (:1642-1672),
 * pack_NV21 (:1818-1848): Y of every line; chroma only from even lines (IS_CHROMA_LINE_420), taken from the even
 * pixel of each pair, plus the last pixel of an odd width */
static void
pack_line_420 (const OracleVcsDesc * d, const uint8_t * ayuv, uint8_t * out, int y)
{
  int w = d->out_width, i;
  uint8_t *dy = out + d->out_offset[0] + (size_t) d->out_stride[0] * y;
  for (i = 0; i < w; i++)
    dy[i] = ayuv[4 * i + 1];
  if (y & 1)
    return;
  if (d->out_format == ORC_FMT_I420 || d->out_format == ORC_FMT_YV12) {
    const int pu = d->out_format == ORC_FMT_YV12 ? 2 : 1, pv = 3 - pu;
    uint8_t *du = out + d->out_offset[pu] + (size_t) d->out_stride[pu] * (y >> 1);
    uint8_t *dv = out + d->out_offset[pv] + (size_t) d->out_stride[pv] * (y >> 1);
    for (i = 0; i < w; i += 2) {
      du[i >> 1] = ayuv[4 * i + 2];
      dv[i >> 1] = ayuv[4 * i + 3];
    }
  } else {
    const int ui = d->out_format == ORC_FMT_NV21 ? 1 : 0;
    uint8_t *duv = out + d->out_offset[1] + (size_t) d->out_stride[1] * (y >> 1);
    for (i = 0; i < w; i += 2) {
      duv[i + ui] = ayuv[4 * i + 2];
      duv[i + (ui ^ 1)] = ayuv[4 * i + 3];
    }
  }
}

/* Which input lines does the chain pull through the chroma upsampler, and how is each
 * paired?  The upsample cache hands out lines in PAIRS (n_lines=2, offset=-1,
 * video-chroma.c:997) starting at whichever line is requested first after a gap
 * (gst_line_cache_get_lines skip-ahead, video-converter.c:571-617; do_upsample_lines
 * :2991-3021).  Requests are monotonic, so: a requested line y that directly follows a
 * requested "first of pair" y-1 is the pair's second line; otherwise y opens a new
 * pair (y, y+1) - except y==0, which is the second line of the clamped pair (-1,0).
 * mode: 0 = keep own chroma row, 1 = first of pair, 2 = second of pair. */
static void
chroma_plan (const OracleVcsDesc * d, const Scaler * vs, uint8_t * mode)
{
  int ih = d->in_height, y, r, k;
  uint8_t *req = calloc (ih, 1);
  int prev_first = -2;
  if (vs) {
    for (r = 0; r < vs->out_size; r++)
      for (k = 0; k < vs->n_taps; k++)
        req[vs->offset[r] + k] = 1;
  } else
    memset (req, 1, ih);
  for (y = 0; y < ih; y++) {
    mode[y] = 0;
    if (!req[y])
      continue;
    if (y == 0)
      mode[y] = 0;              /* pair (-1,0): both lines carry chroma row 0 */
    else if (prev_first == y - 1)
      mode[y] = 2;
    else {
      mode[y] = 1;
      prev_first = y;
    }
  }
  free (req);
}

/* ============================================================== YUV -> same YUV family: plane scaling
 * Fast path convert_scale_planes (video-converter.c:7757-7769, table rows NV12->NV12, I420->I420, I420<->YV12 with
 * keeps_size = FALSE): every plane is scaled on its own by setup_scale's choice (:8092-8245) —
 * luma with the element's method, chroma with the chroma resampler (LINEAR unless the method is NEAREST),
 * exact 2:1 / 1:2 steps of single-byte planes by the planar_chroma averaging / doubling kernels, everything
 * else by gst_video_scaler_2d (video-scaler.c:1451-1640). */

/* one line through the horizontal scaler, n_elems bytes per pixel (1 or 2): video_scale_h_near_u8/_u16,
 * video_scale_h_2tap_1u8 (ldreslinb stepping), video_scale_h_ntap_u8 */
static void
hscale_line_n (Scaler * s, const uint8_t * sl, uint8_t * dl, int dw, int ne)
{
  int x, c, k;
  if (s->n_taps >= 2 && !(s->n_taps == 2 && (ne == 1 || ne == 4)) && !s->taps_s16)
    scaler_quantize (s, 6);
  for (x = 0; x < dw; x++) {
    if (s->n_taps == 1) {
      memcpy (dl + ne * x, sl + ne * s->offset[x], ne);
    } else if (s->n_taps == 2 && (ne == 1 || ne == 4)) {
      /* video_scale_h_2tap_1u8 (mono planes) / video_scale_h_2tap_4u8 (4-byte pixels): edge-aligned 16.16 stepping;
       * 2-byte pixels (the interleaved UV plane) take the n-tap routine (video-scaler.c:1286-1292) */
      int tmp = x * s->inc, i0 = tmp >> 16, f = (tmp >> 8) & 0xff;
      int i1 = (i0 + 1 < s->in_size) ? i0 + 1 : i0;
      for (c = 0; c < ne; c++)
        dl[ne * x + c] = (uint8_t) ((sl[ne * i0 + c] * (256 - f) + sl[ne * i1 + c] * f) >> 8);
    } else {
      const int16_t *t = s->taps_s16 + (size_t) x * s->n_taps;
      for (c = 0; c < ne; c++) {
        int acc = 0;
        for (k = 0; k < s->n_taps; k++)
          acc += (int16_t) (sl[ne * (s->offset[x] + k) + c] * t[k]);
        dl[ne * x + c] = scale_round_u8 (acc);
      }
    }
  }
}

/* one output line of the vertical scaler from `lines[]` (n bytes each): v_near / v_2tap / v_4tap / v_ntap */
static void
vscale_line_n (Scaler * s, const uint8_t ** lines, uint8_t * dl, int y, int n)
{
  int x, k;
  if (s->n_taps >= 2 && !s->taps_s16)
    scaler_quantize (s, s->n_taps == 2 ? 8 : 6);
  if (s->n_taps == 1) {
    memcpy (dl, lines[0], n);
  } else if (s->n_taps == 2) {
    int16_t p1 = s->taps_s16[(size_t) y * 2 + 1];
    for (x = 0; x < n; x++) {
      int16_t w2 = (int16_t) (lines[1][x] - lines[0][x]);
      w2 = (int16_t) (w2 * p1);
      w2 = (int16_t) (w2 + 128);
      dl[x] = (uint8_t) (((uint16_t) w2 >> 8) + lines[0][x]);
    }
  } else {
    const int16_t *t = s->taps_s16 + (size_t) y * s->n_taps;
    for (x = 0; x < n; x++) {
      int acc = 0;
      for (k = 0; k < s->n_taps; k++)
        acc += (int16_t) (lines[k][x] * t[k]);
      dl[x] = scale_round_u8 (acc);
    }
  }
}

/* gst_video_scaler_2d for a whole plane */
static void
scale_plane_2d (Scaler * hs, Scaler * vs, const uint8_t * src, int sstride, int iw, uint8_t * dst, int dstride,
    int ow, int oh, int ne)
{
  int y, j;
  const uint8_t *lines[ORACLE_MAX_TAPS];
  if (!vs) {
    for (y = 0; y < oh; y++) {
      if (hs)
        hscale_line_n (hs, src + (size_t) y * sstride, dst + (size_t) y * dstride, ow, ne);
      else
        memcpy (dst + (size_t) y * dstride, src + (size_t) y * sstride, (size_t) ow * ne);
    }
    return;
  }
  if (!hs) {
    for (y = 0; y < oh; y++) {
      for (j = 0; j < vs->n_taps; j++)
        lines[j] = src + (size_t) (vs->offset[y] + j) * sstride;
      vscale_line_n (vs, lines, dst + (size_t) y * dstride, y, ow * ne);
    }
    return;
  }
  if ((long) ow * vs->offset[oh - 1] <= (long) ow * oh) {
    /* horizontal first: every needed input line is h-scaled (the reference caches them in a ring; same values) */
    uint8_t *tmp = malloc ((size_t) vs->in_size * ow * ne);
    uint8_t *have = calloc (vs->in_size, 1);
    for (y = 0; y < oh; y++) {
      for (j = 0; j < vs->n_taps; j++) {
        int l = vs->offset[y] + j;
        if (!have[l]) {
          hscale_line_n (hs, src + (size_t) l * sstride, tmp + (size_t) l * ow * ne, ow, ne);
          have[l] = 1;
        }
        lines[j] = tmp + (size_t) l * ow * ne;
      }
      vscale_line_n (vs, lines, dst + (size_t) y * dstride, y, ow * ne);
    }
    free (tmp);
    free (have);
  } else {
    /* vertical first into a temp line of input width, then horizontal */
    uint8_t *tmp = malloc ((size_t) iw * ne + 16);
    for (y = 0; y < oh; y++) {
      for (j = 0; j < vs->n_taps; j++)
        lines[j] = src + (size_t) (vs->offset[y] + j) * sstride;
      vscale_line_n (vs, lines, tmp, y, iw * ne);
      hscale_line_n (hs, tmp, dst + (size_t) y * dstride, ow, ne);
    }
    free (tmp);
  }
}

#define AVGUB(a, b) ((uint8_t) (((a) + (b) + 1) >> 1))

static int
convert_planes (const OracleVcsDesc * d, const uint8_t * in, uint8_t * out)
{
  int semi = d->out_format == ORC_FMT_NV12 || d->out_format == ORC_FMT_NV21;
  int n_planes = semi ? 2 : 3, i;
  if (fmt_is_rgb (d->out_format)) {
    /* table rows ARGB->ARGB ... BGRx->BGRx (video-converter.c:8879-8896): one plane of 4-byte pixels through
     * gst_video_scaler_2d with the element's method; every byte, the alpha / padding one included, is a channel */
    OracleResamplerOpts rs = d->rs;
    Scaler hs, vs;
    int need_h = d->in_width != d->out_width, need_v = d->in_height != d->out_height, y;
    if (!need_h && !need_v) {
      for (y = 0; y < d->out_height; y++)
        memcpy (out + d->out_offset[0] + (size_t) y * d->out_stride[0], in + d->in_offset[0] + (size_t) y * d->in_stride[0],
            (size_t) d->out_width * 4);
      return 0;
    }
    if (need_h && scaler_init (&hs, &rs, d->in_width, d->out_width))
      return -1;
    if (need_v && scaler_init (&vs, &rs, d->in_height, d->out_height))
      return -1;
    scale_plane_2d (need_h ? &hs : NULL, need_v ? &vs : NULL, in + d->in_offset[0], d->in_stride[0], d->in_width,
        out + d->out_offset[0], d->out_stride[0], d->out_width, d->out_height, 4);
    if (need_h)
      scaler_clear (&hs);
    if (need_v)
      scaler_clear (&vs);
    return 0;
  }
  int iw = d->in_width, ih = d->in_height, ow = d->out_width, oh = d->out_height;
  for (i = 0; i < n_planes; i++) {
    /* plane i of the output holds component i (Y,U,V) for I420 and (Y,V,U) for YV12; the source plane is the
     * one holding the same component (fsplane) */
    int sp = i;
    int pw, ph, qw, qh, ne = (semi && i == 1) ? 2 : 1, x, y;
    /* chroma plane size of the INPUT: 4:2:0 halves both directions, Y42B the width only, Y444 neither */
    const int in_wsub = d->in_format == ORC_FMT_Y444 ? 0 : 1, in_hsub = (d->in_format == ORC_FMT_Y444 || d->in_format == ORC_FMT_Y42B) ? 0 : 1;
    const uint8_t *s;
    uint8_t *dp;
    int ss, ds, method;
    OracleResamplerOpts rs = d->rs;
    Scaler hs, vs;
    int need_h = 0, need_v = 0;
    if (!semi && i > 0 && (d->in_format == ORC_FMT_YV12) != (d->out_format == ORC_FMT_YV12))
      sp = 3 - i;               /* exactly one side is YV12: U and V planes swap */
    pw = i ? (iw + in_wsub) >> in_wsub : iw; ph = i ? (ih + in_hsub) >> in_hsub : ih;
    qw = i ? (ow + 1) / 2 : ow; qh = i ? (oh + 1) / 2 : oh;
    s = in + d->in_offset[sp]; ss = d->in_stride[sp];
    dp = out + d->out_offset[i]; ds = d->out_stride[i];
    /* resample_method = (i == 0 ? method : cr_method), cr_method = LINEAR unless method == NEAREST (:7981-7986) */
    method = rs.method;
    if (i > 0 && method != ORC_RS_NEAREST)
      method = ORC_RS_LINEAR;
    rs.method = method;
    if (pw == qw && ph == qh) {
      for (y = 0; y < qh; y++)
        memcpy (dp + (size_t) y * ds, s + (size_t) y * ss, (size_t) qw * ne);
      continue;
    }
    if (ne == 1 && method == ORC_RS_LINEAR && ((pw == qw && ph == 2 * qh) || (ph == qh && pw == 2 * qw) ||
            (pw == 2 * qw && ph == 2 * qh))) {
      /* convert_plane_v_halve / _h_halve / _hv_halve: video_orc_planar_chroma_422_420 / _444_422 / _444_420 */
      for (y = 0; y < qh; y++)
        for (x = 0; x < qw; x++) {
          if (pw == qw)
            dp[(size_t) y * ds + x] = AVGUB (s[(size_t) (2 * y) * ss + x], s[(size_t) (2 * y + 1) * ss + x]);
          else if (ph == qh)
            dp[(size_t) y * ds + x] = AVGUB (s[(size_t) y * ss + 2 * x], s[(size_t) y * ss + 2 * x + 1]);
          else {
            uint8_t t1 = AVGUB (s[(size_t) (2 * y) * ss + 2 * x], s[(size_t) (2 * y + 1) * ss + 2 * x]);
            uint8_t t2 = AVGUB (s[(size_t) (2 * y) * ss + 2 * x + 1], s[(size_t) (2 * y + 1) * ss + 2 * x + 1]);
            dp[(size_t) y * ds + x] = AVGUB (t1, t2);
          }
        }
      continue;
    }
    if (ne == 1 && method == ORC_RS_NEAREST && ((pw == qw && 2 * ph == qh) || (ph == qh && 2 * pw == qw) ||
            (2 * pw == qw && 2 * ph == qh))) {
      /* convert_plane_v_double / _h_double / _hv_double: plain replication */
      for (y = 0; y < qh; y++)
        for (x = 0; x < qw; x++)
          dp[(size_t) y * ds + x] = s[(size_t) (ph == qh ? y : y / 2) * ss + (pw == qw ? x : x / 2)];
      continue;
    }
    need_h = pw != qw;
    need_v = ph != qh;
    if (need_h && scaler_init (&hs, &rs, pw, qw))
      return -1;
    if (need_v && scaler_init (&vs, &rs, ph, qh))
      return -1;
    scale_plane_2d (need_h ? &hs : NULL, need_v ? &vs : NULL, s, ss, pw, dp, ds, qw, qh, ne);
    if (need_h)
      scaler_clear (&hs);
    if (need_v)
      scaler_clear (&vs);
  }
  return 0;
}

/* table rows YUY2 / UYVY -> I420 / YV12 (video-converter.c:8493-8507: unchanged size, same matrix; range and chroma-site are
 * not looked at): convert_YUY2_I420 / convert_UYVY_I420 (:3954-4028, :4913-4986) -> video_orc_convert_YUY2_I420 /
 * _UYVY_I420 (video-orc.orc:716-744, :781-809): per line pair the luma of both lines is copied - (width + 1) / 2 PAIRS, so an
 * odd width writes one byte of row padding - and each chroma sample is avgub of the pair's two lines, no horizontal filter.
 * A last odd line goes through unpack + pack_planar_420 (video-format.c:117-148): its own chroma, even pixels. */
static int
convert_yuy2_i420 (const OracleVcsDesc * d, const uint8_t * in, uint8_t * out)
{
  const int w = d->in_width, h = d->in_height, h2 = h & ~1, np = (w + 1) / 2;
  const int yo = d->in_format == ORC_FMT_UYVY ? 1 : 0, uo = d->in_format == ORC_FMT_UYVY ? 0 : 1, vo = uo + 2;
  const int pu = d->out_format == ORC_FMT_YV12 ? 2 : 1, pv = 3 - pu;
  int y, i;
  for (y = 0; y < h2; y += 2) {
    const uint8_t *s1 = in + d->in_offset[0] + (size_t) d->in_stride[0] * y, *s2 = s1 + d->in_stride[0];
    uint8_t *y1 = out + d->out_offset[0] + (size_t) d->out_stride[0] * y, *y2 = y1 + d->out_stride[0];
    uint8_t *du = out + d->out_offset[pu] + (size_t) d->out_stride[pu] * (y >> 1);
    uint8_t *dv = out + d->out_offset[pv] + (size_t) d->out_stride[pv] * (y >> 1);
    for (i = 0; i < np; i++) {
      y1[2 * i] = s1[4 * i + yo]; y1[2 * i + 1] = s1[4 * i + yo + 2];
      y2[2 * i] = s2[4 * i + yo]; y2[2 * i + 1] = s2[4 * i + yo + 2];
      du[i] = AVGUB (s1[4 * i + uo], s2[4 * i + uo]);
      dv[i] = AVGUB (s1[4 * i + vo], s2[4 * i + vo]);
    }
  }
  if (h2 != h) {
    uint8_t *line = malloc ((size_t) w * 4);
    unpack_line (d, in, h - 1, line);
    pack_line_420 (d, line, out, h - 1);
    free (line);
  }
  return 0;
}

int
oracle_vcs_convert (const OracleVcsDesc * d, const uint8_t * in, uint8_t * out)
{
  int iw = d->in_width, ih = d->in_height, ow = d->out_width, oh = d->out_height;
  int p[5], im[4][4], y, cw, ch;
  uint8_t *cur, *tmp, *mode;
  Scaler hs, vs;
  int have_h = iw != ow, have_v = ih != oh, pass;
  int yuv_out = 0, out_site = 0, rgb_in = 0, rgb_out = 0;
  long s0, s3;

  if (fmt_is_rgb (d->in_format)) {
    /* packed RGB -> 4:2:0 (the encoder-feeding direction): no table row, generic chain: unpack to ARGB, the scalers
     * that shrink, the RGB -> YUV matrix (chain_convert :1720-1868 -> video_converter_matrix8_table), the scalers
     * that grow, chroma down-sampling (RGB has no sub-sampling: only the down side exists, :2850-2895), 4:2:0 pack */
    int m_, r_;
    if (d->out_format == d->in_format)
      return convert_planes (d, in, out);
    if (fmt_is_rgb (d->out_format)) {
      /* another byte order: the chain with no matrix stage (both sides name the RGB matrix) and, at the default
       * alpha value 1.0, no alpha stage either (convert_get_alpha_mode :2263-2294): unpack, scalers, pack */
      rgb_in = rgb_out = 1;
    } else {
      if (!(d->out_format == ORC_FMT_I420 || d->out_format == ORC_FMT_YV12 || d->out_format == ORC_FMT_NV12 ||
              d->out_format == ORC_FMT_NV21))
        return -1;
      if (oracle_vcs_matrix_rgb2yuv (d, im) != 0)
        return -1;
      rgb_in = yuv_out = 1;
      rgb_in_out_colorimetry (d, &m_, &r_, &out_site);
    }
  } else {
    int in_planar = d->in_format == ORC_FMT_I420 || d->in_format == ORC_FMT_YV12;
    int out_planar = d->out_format == ORC_FMT_I420 || d->out_format == ORC_FMT_YV12;
    if ((in_planar && out_planar) || ((d->out_format == ORC_FMT_NV12 || d->out_format == ORC_FMT_NV21) &&
            d->in_format == d->out_format))
      return convert_planes (d, in, out);
    if ((d->in_format == ORC_FMT_Y42B || d->in_format == ORC_FMT_Y444) && out_planar) {
      /* planar 4:2:2 / 4:4:4 -> planar 4:2:0: plane-scaling table rows like I420 -> I420 (video-converter.c:8607-8628,
       * same colour matrix required, :8989): every output plane is scaled from the plane holding the same component */
      if (d->out_matrix && d->out_matrix != d->in_matrix)
        return -1;
      return convert_planes (d, in, out);
    }
    if (out_planar || d->out_format == ORC_FMT_NV12 || d->out_format == ORC_FMT_NV21) {
      /* the other 4:2:0 pairs (NV12 <-> I420, NV12 <-> NV21 ...) have no table row: generic chain, with
       * chain_downsample (video-converter.c:2018-2032) and the 4:2:0 pack functions at its end.  chain_convert
       * (:1720-1868) adds no matrix stage when both sides name the same colour matrix (the range is not
       * compared); a differing matrix selects video_orc_matrix8, whose SIMD program and C backup disagree
       * (video-orc.orc:2079-2133 vs video-converter.c:1136-1176): not restated */
      if (d->out_matrix && d->out_matrix != d->in_matrix)
        return -1;
      yuv_out = 1;
      out_site = d->out_chroma_site ? d->out_chroma_site : d->in_chroma_site;
      if (fmt_is_422_444 (d->in_format)) {
        /* 4:2:2 / 4:4:4 -> 4:2:0 through the chain (packed inputs; planar ones to the semi-planar outputs - their planar
         * outputs were plane-scaling rows above): the sub-sampling changes, so the element's fixation does not carry
         * the input's chroma-site over (gstvideoconvertscale.c:1411-1424): the output keeps the default of its size */
        if (!d->out_chroma_site)
          out_site = oh > 576 ? ORC_SITE_H_COSITED : ORC_SITE_NONE;
        if ((d->in_format == ORC_FMT_YUY2 || d->in_format == ORC_FMT_UYVY) && out_planar && iw == ow && ih == oh)
          return convert_yuy2_i420 (d, in, out);
      }
    }
  }
  if (yuv_out && !rgb_in && !fmt_is_422_444 (d->in_format) && iw == ow && ih == oh && out_site == d->in_chroma_site && !d->force_resample) {
    /* video_converter_compute_resample (:2850-2895): same sub-sampling, site and size -> no chroma resampler
     * on either side; unpack replicates every chroma sample and pack reads it back */
    uint8_t *line = malloc ((size_t) iw * 4);
    for (y = 0; y < ih; y++) {
      unpack_line (d, in, y, line);
      pack_line_420 (d, line, out, y);
    }
    free (line);
    return 0;
  }
  if (!yuv_out && !rgb_out && oracle_vcs_matrix (d, p, im) != 0)
    return -1;
  if (rgb_in) {
    if (have_h && scaler_init (&hs, &d->rs, iw, ow))
      return -1;
    if (have_v && scaler_init (&vs, &d->rs, ih, oh))
      return -1;
    cur = malloc ((size_t) iw * ih * 4);
    for (y = 0; y < ih; y++)
      unpack_line_rgb (d, in, y, cur + (size_t) y * iw * 4);
    goto scale_passes;
  }
  if (!yuv_out && !rgb_in && (d->in_format == ORC_FMT_I420 || d->in_format == ORC_FMT_YV12) && iw == ow && ih == oh) {
    /* fast path (video-converter.c:8766-8800 table rows, keeps_size): convert_I420_BGRA / _ARGB /
     * _pack_ARGB (:6772-6988) -> video_orc_convert_I420_BGRA (video-orc.orc:1859-1911): the chroma
     * sample of row y>>1 is used as is for both of its pixels (no up-sampling filter), same mulhi
     * matrix as AYUV->ARGB, alpha = 255 */
    uint8_t *line = malloc ((size_t) iw * 4);
    for (y = 0; y < ih; y++) {
      unpack_line (d, in, y, line);
      matrix_line (line, iw, p);
      pack_line (d->out_format, line, out + d->out_offset[0] + (size_t) d->out_stride[0] * y, ow);
    }
    free (line);
    return 0;
  }
  /* chain_hscale / chain_vscale always build in_width->out_width and
   * in_height->out_height scalers (video-converter.c:1625-1683) */
  if (have_h && scaler_init (&hs, &d->rs, iw, ow))
    return -1;
  if (have_v && scaler_init (&vs, &d->rs, ih, oh))
    return -1;

  /* unpack + chroma upsample, whole frame: h filter on every line, then the v filter
   * on the pairs the pull chain forms.  V_COSITED selects the "IMPLEMENT ME"
   * resampler (video-chroma.c:1000): h filter only, n_lines=1. */
  cur = malloc ((size_t) iw * ih * 4);
  mode = malloc (ih);
  chroma_plan (d, have_v ? &vs : NULL, mode);
  for (y = 0; y < ih; y++) {
    unpack_line (d, in, y, cur + (size_t) y * iw * 4);
    if (d->in_format != ORC_FMT_Y444)   /* 4:4:4: gst_video_chroma_resample_new (.., 0, 0) is NULL (video-chroma.c:1054-1055) */
      chroma_h_line (cur + (size_t) y * iw * 4, iw, (d->in_chroma_site & ORC_SITE_H_COSITED) != 0);
  }
  /* 4:2:2: v_factor 0 selects video_chroma_none (video-chroma.c:989-994): horizontal filter only, line by line */
  if (!(d->in_chroma_site & ORC_SITE_V_COSITED) && !fmt_is_422_444 (d->in_format)) {
    /* both lines of a pair are filtered from the ORIGINAL (h-filtered) values, so
     * work from a copy of the chroma of the partner line */
    uint8_t *orig = malloc ((size_t) iw * ih * 4);
    memcpy (orig, cur, (size_t) iw * ih * 4);
    for (y = 0; y < ih; y++) {
      int x, c, yo = -1;
      uint8_t *l = cur + (size_t) y * iw * 4;
      if (mode[y] == 1)
        yo = y + 1 < ih ? y + 1 : ih - 1;       /* do_unpack_lines clamp, :2973 */
      else if (mode[y] == 2)
        yo = y - 1;
      if (yo < 0)
        continue;
      for (x = 0; x < iw; x++)
        for (c = 2; c < 4; c++) {
          int own = orig[((size_t) y * iw + x) * 4 + c], oth = orig[((size_t) yo * iw + x) * 4 + c];
          l[4 * x + c] = (uint8_t) ((3 * own + oth + 2) >> 2);  /* FILT_3_1 / FILT_1_3 */
        }
    }
    free (orig);
  }
  free (mode);

scale_passes:
  cw = iw;
  ch = ih;
  s0 = (long) iw * ih;
  s3 = (long) ow * oh;
  /* pass 0 = chain_scale(force=FALSE) before the matrix, pass 1 = chain_scale(force=TRUE)
   * after it (video-converter.c:2517-2531, :1685-1718) */
  for (pass = 0; pass < 2; pass++) {
    if (pass == 1 && !yuv_out && !rgb_out) {
      for (y = 0; y < ch; y++)
        matrix_line (cur + (size_t) y * cw * 4, cw, p);
    }
    if (pass == 1 && rgb_in && yuv_out) {
      for (y = 0; y < ch; y++)
        matrix_line_rgb2yuv (cur + (size_t) y * cw * 4, cw, im);
    }
    if (pass == 0 && !(s3 <= s0))
      continue;
    if (cw == ow && ch == oh)
      continue;
    {
      long s1 = (long) ow * ch, s2 = (long) cw * oh;
      int h_first = s1 <= s2, step;
      for (step = 0; step < 2; step++) {
        int do_h = (step == 0) == (h_first != 0);
        if (do_h && cw != ow) {
          tmp = malloc ((size_t) ow * ch * 4);
          hscale_image (&hs, cur, cw, tmp, ow, ch);
          free (cur);
          cur = tmp;
          cw = ow;
        } else if (!do_h && ch != oh) {
          tmp = malloc ((size_t) cw * oh * 4);
          vscale_image (&vs, cur, ch, tmp, oh, cw);
          free (cur);
          cur = tmp;
          ch = oh;
        }
      }
    }
  }
  if (yuv_out) {
    /* do_downsample_lines (:3194-3222): the pack stage pulls line after line; every even line 2k makes the
     * down-sampler fetch the pair (2k, 2k+1) and filter it IN PLACE: vertical average of the chroma into line 2k
     * (video_chroma_down_v2_u8, video-chroma.c:434-442; the V_COSITED resampler is the unimplemented one that only
     * runs the horizontal filter, :774-785), then the horizontal filter on line 2k.  For an odd height the pair's
     * second line is line `oh`: the vertical scaler clamps it back to oh-1 (:3070-3080), but without a vertical
     * scaler the request reaches do_unpack_lines' clamp (:2973) through a fresh up-sampler pair (oh, oh+1), i.e.
     * the last source line with its chroma row NOT vertically filtered. */
    uint8_t *extra = NULL;
    if ((oh & 1) && !have_v && !(out_site & ORC_SITE_V_COSITED) && !rgb_in) {   /* RGB: the clamped line IS the last line */
      uint8_t *one = malloc ((size_t) iw * 4);
      unpack_line (d, in, ih - 1, one);
      if (d->in_format != ORC_FMT_Y444)     /* 4:4:4 has no up-sampler: the clamped request is the last line as unpacked */
        chroma_h_line (one, iw, (d->in_chroma_site & ORC_SITE_H_COSITED) != 0);
      if (have_h) {
        extra = malloc ((size_t) ow * 4);
        hscale_image (&hs, one, iw, extra, ow, 1);
        free (one);
      } else
        extra = one;
    }
    for (y = 0; y < oh; y += 2) {
      uint8_t *l0 = cur + (size_t) y * ow * 4;
      if (!(out_site & ORC_SITE_V_COSITED)) {
        const uint8_t *l1 = y + 1 < oh ? l0 + (size_t) ow * 4 : (extra ? extra : l0);
        int x, c;
        for (x = 0; x < ow; x++)
          for (c = 2; c < 4; c++)
            l0[4 * x + c] = (uint8_t) ((l0[4 * x + c] + l1[4 * x + c] + 1) >> 1);
      }
      chroma_down_h_line (l0, ow, (out_site & ORC_SITE_H_COSITED) != 0);
    }
    free (extra);
    for (y = 0; y < oh; y++)
      pack_line_420 (d, cur + (size_t) y * ow * 4, out, y);
  } else
    for (y = 0; y < oh; y++)
      pack_line (d->out_format, cur + (size_t) y * ow * 4,
          out + d->out_offset[0] + (size_t) d->out_stride[0] * y, ow);
  free (cur);
  if (have_h)
    scaler_clear (&hs);
  if (have_v)
    scaler_clear (&vs);
  return 0;
}


/* ===================================================================== borders */

static unsigned long long
gcd_u64 (unsigned long long a, unsigned long long b)
{
  while (b) {
    unsigned long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}

void
oracle_vcs_borders (int in_w, int in_h, int out_w, int out_h, int dest[4])
{
  /* gst_video_convert_scale_set_info (gstvideoconvertscale.c:920-952) with par 1/1 on both sides:
   * from_dar = in_w/in_h and to_dar = out_w/out_h reduced (gst_util_fraction_multiply); different ->
   * to_h = out_w * d / n (gst_util_uint64_scale_int: floor); fits -> bars above/below, else to_w = out_h * n / d */
  unsigned long long g1 = gcd_u64 (in_w, in_h), g2 = gcd_u64 (out_w, out_h);
  int n = (int) (in_w / g1), dd = (int) (in_h / g1);
  int borders_w = 0, borders_h = 0;
  if (n != (int) (out_w / g2) || dd != (int) (out_h / g2)) {
    int to_h = (int) (((unsigned long long) out_w * dd) / n);
    if (to_h <= out_h)
      borders_h = out_h - to_h;
    else
      borders_w = out_w - (int) (((unsigned long long) out_h * n) / dd);
  }
  dest[0] = borders_w / 2;
  dest[1] = borders_h / 2;
  dest[2] = out_w - borders_w;
  dest[3] = out_h - borders_h;
}

int
oracle_vcs_convert_dest (const OracleVcsDesc * d, int dest_x, int dest_y, int dest_w, int dest_h, uint32_t border_argb,
    const uint8_t * in, uint8_t * out)
{
  OracleVcsDesc q = *d;
  const int W = d->out_width, H = d->out_height;
  const int yuv = !fmt_is_rgb (d->out_format);
  const int planar = d->out_format == ORC_FMT_I420 || d->out_format == ORC_FMT_YV12;
  int x, y, c, st;
  uint8_t argb[4] = { (uint8_t) (border_argb >> 24), (uint8_t) (border_argb >> 16), (uint8_t) (border_argb >> 8),
    (uint8_t) border_argb };
  uint8_t px[4];
  if (yuv) {                    /* out_x / out_y are rounded down to the chroma grid (video-converter.c:2335-2336) */
    dest_x &= ~1;
    dest_y &= ~1;
  }
  /* :2338-2362 */
  if (dest_w > W - dest_x)
    dest_w = W - dest_x;
  dest_w = CLAMPI (dest_w, 0, W);
  if (dest_h > H - dest_y)
    dest_h = H - dest_y;
  dest_h = CLAMPI (dest_h, 0, H);
  if (dest_w < 1 || dest_h < 1)
    return -1;                  /* an empty rectangle: no chain is built (:2509-2510) and the rows of the rectangle are
                                 * never touched - degenerate, not restated */
  /* the chain runs on the rectangle: same layout, shifted plane origins; colorimetry defaults stay those of the
   * whole frame */
  q.out_width = dest_w;
  q.out_height = dest_h;
  /* video_converter_compute_resample (:2850-2895) compares the input with the whole output frame */
  q.force_resample = d->in_width != W || d->in_height != H;
  if (!yuv) {
    q.out_offset[0] += (size_t) dest_y * d->out_stride[0] + (size_t) dest_x * 4;
  } else {
    q.out_offset[0] += (size_t) dest_y * d->out_stride[0] + dest_x;
    if (planar) {
      q.out_offset[1] += (size_t) (dest_y / 2) * d->out_stride[1] + dest_x / 2;
      q.out_offset[2] += (size_t) (dest_y / 2) * d->out_stride[2] + dest_x / 2;
    } else
      q.out_offset[1] += (size_t) (dest_y / 2) * d->out_stride[1] + dest_x;
    if (fmt_is_rgb (d->in_format)) {
      int m_, r_, s_;
      rgb_in_out_colorimetry (d, &m_, &r_, &s_);
      q.out_matrix = m_;
      q.out_range = r_;
      q.out_chroma_site = s_;
    }
  }
  st = oracle_vcs_convert (&q, in, out);
  if (st != 0)
    return st;
  if (dest_w == W && dest_h == H)
    return 0;
  /* border pixel: setup_borderline (:2189-2258) */
  if (!yuv) {
    switch (d->out_format) {
      case ORC_FMT_BGRA: case ORC_FMT_BGRx: px[0] = argb[3]; px[1] = argb[2]; px[2] = argb[1]; px[3] = argb[0]; break;
      case ORC_FMT_RGBA: case ORC_FMT_RGBx: px[0] = argb[1]; px[1] = argb[2]; px[2] = argb[3]; px[3] = argb[0]; break;
      case ORC_FMT_ABGR: case ORC_FMT_xBGR: px[0] = argb[0]; px[1] = argb[3]; px[2] = argb[2]; px[3] = argb[1]; break;
      default: memcpy (px, argb, 4); break;
    }
    for (y = 0; y < H; y++)
      for (x = 0; x < W; x++)
        if (y < dest_y || y >= dest_y + dest_h || x < dest_x || x >= dest_x + dest_w)
          memcpy (out + d->out_offset[0] + (size_t) y * d->out_stride[0] + 4 * (size_t) x, px, 4);
  } else {
    /* identity -> compute_matrix_to_YUV (force) with the OUTPUT colorimetry -> rint (no x256 here: the matrix maps
     * [0,1] RGB... it is applied to 8-bit r,g,b and shifted by 8, offsets added by hand: :2207-2226) */
    Mat m;
    double Kr, Kb, Kg, xk;
    int im[3][3], matrix, range, site, yv, uv, vv, cx0, cy0, cw, chh, CW, CH;
    if (fmt_is_rgb (d->in_format))
      rgb_in_out_colorimetry (d, &matrix, &range, &site);
    else {
      matrix = d->out_matrix ? d->out_matrix : d->in_matrix;
      range = d->out_range ? d->out_range : d->in_range;
    }
    mat_identity (&m);
    if (kr_kb (matrix, &Kr, &Kb))
      return -1;
    Kg = 1.0 - Kr - Kb;
    {
      Mat k;
      mat_identity (&k);
      k.dm[0][0] = Kr; k.dm[0][1] = Kg; k.dm[0][2] = Kb;
      xk = 1 / (2 * (1 - Kb));
      k.dm[1][0] = -xk * Kr; k.dm[1][1] = -xk * Kg; k.dm[1][2] = xk * (1 - Kb);
      xk = 1 / (2 * (1 - Kr));
      k.dm[2][0] = xk * (1 - Kr); k.dm[2][1] = -xk * Kg; k.dm[2][2] = -xk * Kb;
      mat_mul (&m, &k, &m);
    }
    if (range == ORC_RANGE_16_235) {
      mat_scale (&m, (float) 219, (float) 224, (float) 224);
      mat_offset (&m, 16, 128, 128);
    } else {
      mat_scale (&m, (float) 255, (float) 255, (float) 255);
      mat_offset (&m, 0, 128, 128);
    }
    for (y = 0; y < 3; y++)
      for (x = 0; x < 3; x++)
        im[y][x] = (int) rint (m.dm[y][x]);
    yv = 16 + ((argb[1] * im[0][0] + argb[2] * im[0][1] + argb[3] * im[0][2]) >> 8);
    uv = 128 + ((argb[1] * im[1][0] + argb[2] * im[1][1] + argb[3] * im[1][2]) >> 8);
    vv = 128 + ((argb[1] * im[2][0] + argb[2] * im[2][1] + argb[3] * im[2][2]) >> 8);
    yv = CLAMPI (yv, 0, 255);
    uv = CLAMPI (uv, 0, 255);
    vv = CLAMPI (vv, 0, 255);
    for (y = 0; y < H; y++)
      for (x = 0; x < W; x++)
        if (y < dest_y || y >= dest_y + dest_h || x < dest_x || x >= dest_x + dest_w)
          out[d->out_offset[0] + (size_t) y * d->out_stride[0] + x] = (uint8_t) yv;
    /* chroma planes: GST_VIDEO_FORMAT_INFO_SCALE_WIDTH / _HEIGHT round up (convert_fill_border :7209-7221) */
    cx0 = (dest_x + 1) / 2; cy0 = (dest_y + 1) / 2; cw = (dest_w + 1) / 2; chh = (dest_h + 1) / 2;
    CW = (W + 1) / 2; CH = (H + 1) / 2;
    for (y = 0; y < CH; y++)
      for (x = 0; x < CW; x++) {
        if (!(y < cy0 || y >= cy0 + chh || x < cx0 || x >= cx0 + cw))
          continue;
        if (planar) {
          const int pu = d->out_format == ORC_FMT_YV12 ? 2 : 1, pv = 3 - pu;
          out[d->out_offset[pu] + (size_t) y * d->out_stride[pu] + x] = (uint8_t) uv;
          out[d->out_offset[pv] + (size_t) y * d->out_stride[pv] + x] = (uint8_t) vv;
        } else {
          const int ui = d->out_format == ORC_FMT_NV21 ? 1 : 0;
          out[d->out_offset[1] + (size_t) y * d->out_stride[1] + 2 * x + ui] = (uint8_t) uv;
          out[d->out_offset[1] + (size_t) y * d->out_stride[1] + 2 * x + (ui ^ 1)] = (uint8_t) vv;
        }
      }
    (void) c;
  }
  return 0;
}
