/* oracle/refshim/refdrv_video.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points (ctypes-friendly) over the *unmodified* reference
 * converter compiled in place from /root/reference:
 *   gst_video_info_set_format  (gst-libs/gst/video/video-info.c:292)
 *   gst_video_converter_new    (gst-libs/gst/video/video-converter.c:2577)
 *   gst_video_converter_frame  (gst-libs/gst/video/video-converter.c:2782)
 * with the option bag the videoconvertscale element builds in set_info()
 * (gst/videoconvertscale/gstvideoconvertscale.c:991-1087).
 *
 * Used to (1) validate oracle/ (the restatement), (2) generate the golden
 * fixtures under tests/golden/, (3) time the reference CPU path
 * (bench.py --impl reference, cpu_baseline kind "reference").
 */
#include <gst/video/video.h>

/* mirrors GstVideoScaleMethod of the element (gstvideoconvertscale.h) */
enum
{
  REF_SCALE_NEAREST = 0, REF_SCALE_BILINEAR, REF_SCALE_4TAP, REF_SCALE_LANCZOS,
  REF_SCALE_BILINEAR2, REF_SCALE_SINC, REF_SCALE_HERMITE, REF_SCALE_SPLINE,
  REF_SCALE_CATROM, REF_SCALE_MITCHELL
};

typedef struct
{
  GstVideoConverter *convert;
  GstVideoInfo in_info, out_info;
} RefVcs;

static void
apply_defaults_like_caps (GstVideoInfo * info, int matrix, int range, int chroma_site)
{
  /* gst_video_info_from_caps() fills chroma-site (video-info.c:211-225) when the
   * caps carry none; gst_video_info_set_format() alone leaves it UNKNOWN. */
  if (GST_VIDEO_FORMAT_INFO_IS_YUV (info->finfo)) {
    if (chroma_site < 0)
      info->chroma_site = info->height > 576 ?
          GST_VIDEO_CHROMA_SITE_H_COSITED : GST_VIDEO_CHROMA_SITE_NONE;
    else
      info->chroma_site = (GstVideoChromaSite) chroma_site;
  }
  if (matrix >= 0)
    info->colorimetry.matrix = (GstVideoColorMatrix) matrix;
  if (range >= 0)
    info->colorimetry.range = (GstVideoColorRange) range;
}

/* query the default system-memory layout: strides/offsets/size (video-info.c fill_planes) */
int
ref_video_info (int format, int width, int height, int *n_planes, int stride[4],
    size_t offset[4], size_t * size)
{
  GstVideoInfo info;
  int i;
  if (!gst_video_info_set_format (&info, (GstVideoFormat) format, width, height))
    return -1;
  *n_planes = GST_VIDEO_INFO_N_PLANES (&info);
  for (i = 0; i < 4; i++) {
    stride[i] = info.stride[i];
    offset[i] = info.offset[i];
  }
  *size = info.size;
  return 0;
}

int
ref_video_format_from_string (const char *s)
{
  return (int) gst_video_format_from_string (s);
}

/* the element's DAR-preserving borders (gstvideoconvertscale.c:926-952 -> GST_VIDEO_CONVERTER_OPT_DEST_*, :1068-1072):
 * the destination rectangle of the NEXT ref_vcs_new() call; w < 0 = the whole output frame */
static int next_dest[4] = { 0, 0, -1, -1 };
static unsigned next_border = 0xff000000u;      /* DEFAULT_OPT_BORDER_ARGB, video-converter.c:778 */

void
ref_vcs_next_border_argb (unsigned argb)
{
  next_border = argb;
}

void
ref_vcs_next_dest (int x, int y, int w, int h)
{
  next_dest[0] = x;
  next_dest[1] = y;
  next_dest[2] = w;
  next_dest[3] = h;
}

/* in_cm / in_range / in_site: -1 = caps default.  strides NULL = default layout. */
RefVcs *
ref_vcs_new (int in_format, int in_w, int in_h, const int *in_stride,
    const size_t *in_offset, int in_cm, int in_range, int in_site,
    int out_format, int out_w, int out_h, const int *out_stride,
    const size_t *out_offset, int out_cm, int out_range, int out_site,
    int method, int n_threads, int dither, double envelope, double sharpness,
    double sharpen)
{
  RefVcs *r = g_new0 (RefVcs, 1);
  GstStructure *options;
  int i;

  if (!gst_video_info_set_format (&r->in_info, (GstVideoFormat) in_format, in_w, in_h) ||
      !gst_video_info_set_format (&r->out_info, (GstVideoFormat) out_format, out_w, out_h)) {
    g_free (r);
    return NULL;
  }
  apply_defaults_like_caps (&r->in_info, in_cm, in_range, in_site);
  apply_defaults_like_caps (&r->out_info, out_cm, out_range, out_site);
  for (i = 0; i < 4; i++) {
    if (in_stride) {
      r->in_info.stride[i] = in_stride[i];
      r->in_info.offset[i] = in_offset[i];
    }
    if (out_stride) {
      r->out_info.stride[i] = out_stride[i];
      r->out_info.offset[i] = out_offset[i];
    }
  }

  options = gst_structure_new_static_str_empty ("videoconvertscale");
  switch (method) {
    case REF_SCALE_NEAREST:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_NEAREST, NULL);
      break;
    case REF_SCALE_BILINEAR:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_LINEAR,
          GST_VIDEO_RESAMPLER_OPT_MAX_TAPS, G_TYPE_INT, 2, NULL);
      break;
    case REF_SCALE_4TAP:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_SINC,
          GST_VIDEO_RESAMPLER_OPT_MAX_TAPS, G_TYPE_INT, 4, NULL);
      break;
    case REF_SCALE_LANCZOS:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_LANCZOS, NULL);
      break;
    case REF_SCALE_BILINEAR2:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_LINEAR, NULL);
      break;
    case REF_SCALE_SINC:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_SINC, NULL);
      break;
    case REF_SCALE_HERMITE:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_CUBIC,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_B, G_TYPE_DOUBLE, (gdouble) 0.0,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_C, G_TYPE_DOUBLE, (gdouble) 0.0, NULL);
      break;
    case REF_SCALE_SPLINE:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_CUBIC,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_B, G_TYPE_DOUBLE, (gdouble) 1.0,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_C, G_TYPE_DOUBLE, (gdouble) 0.0, NULL);
      break;
    case REF_SCALE_CATROM:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_CUBIC,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_B, G_TYPE_DOUBLE, (gdouble) 0.0,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_C, G_TYPE_DOUBLE, (gdouble) 0.5, NULL);
      break;
    case REF_SCALE_MITCHELL:
    default:
      gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_RESAMPLER_METHOD,
          GST_TYPE_VIDEO_RESAMPLER_METHOD, GST_VIDEO_RESAMPLER_METHOD_CUBIC,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_B, G_TYPE_DOUBLE, (gdouble) 1.0 / 3.0,
          GST_VIDEO_RESAMPLER_OPT_CUBIC_C, G_TYPE_DOUBLE, (gdouble) 1.0 / 3.0, NULL);
      break;
  }
  /* same keys/values as the element, borders = 0 (no DAR padding) */
  gst_structure_set_static_str (options,
      GST_VIDEO_RESAMPLER_OPT_ENVELOPE, G_TYPE_DOUBLE, envelope,
      GST_VIDEO_RESAMPLER_OPT_SHARPNESS, G_TYPE_DOUBLE, sharpness,
      GST_VIDEO_RESAMPLER_OPT_SHARPEN, G_TYPE_DOUBLE, sharpen,
      GST_VIDEO_CONVERTER_OPT_DEST_X, G_TYPE_INT, next_dest[2] < 0 ? 0 : next_dest[0],
      GST_VIDEO_CONVERTER_OPT_DEST_Y, G_TYPE_INT, next_dest[2] < 0 ? 0 : next_dest[1],
      GST_VIDEO_CONVERTER_OPT_DEST_WIDTH, G_TYPE_INT, next_dest[2] < 0 ? out_w : next_dest[2],
      GST_VIDEO_CONVERTER_OPT_DEST_HEIGHT, G_TYPE_INT, next_dest[2] < 0 ? out_h : next_dest[3],
      GST_VIDEO_CONVERTER_OPT_DITHER_METHOD, GST_TYPE_VIDEO_DITHER_METHOD, dither,
      GST_VIDEO_CONVERTER_OPT_DITHER_QUANTIZATION, G_TYPE_UINT, 1u,
      GST_VIDEO_CONVERTER_OPT_CHROMA_RESAMPLER_METHOD, GST_TYPE_VIDEO_RESAMPLER_METHOD,
      GST_VIDEO_RESAMPLER_METHOD_LINEAR,
      GST_VIDEO_CONVERTER_OPT_ALPHA_MODE, GST_TYPE_VIDEO_ALPHA_MODE, GST_VIDEO_ALPHA_MODE_COPY,
      GST_VIDEO_CONVERTER_OPT_ALPHA_VALUE, G_TYPE_DOUBLE, 1.0,
      GST_VIDEO_CONVERTER_OPT_CHROMA_MODE, GST_TYPE_VIDEO_CHROMA_MODE, GST_VIDEO_CHROMA_MODE_FULL,
      GST_VIDEO_CONVERTER_OPT_MATRIX_MODE, GST_TYPE_VIDEO_MATRIX_MODE, GST_VIDEO_MATRIX_MODE_FULL,
      GST_VIDEO_CONVERTER_OPT_GAMMA_MODE, GST_TYPE_VIDEO_GAMMA_MODE, GST_VIDEO_GAMMA_MODE_NONE,
      GST_VIDEO_CONVERTER_OPT_PRIMARIES_MODE, GST_TYPE_VIDEO_PRIMARIES_MODE,
      GST_VIDEO_PRIMARIES_MODE_NONE,
      GST_VIDEO_CONVERTER_OPT_THREADS, G_TYPE_UINT, (guint) n_threads, NULL);
  if (next_border != 0xff000000u)
    gst_structure_set_static_str (options, GST_VIDEO_CONVERTER_OPT_BORDER_ARGB, G_TYPE_UINT, (guint) next_border, NULL);
  next_border = 0xff000000u;

  next_dest[2] = next_dest[3] = -1;
  if (method < 0) {
    /* GstVideoAggregatorConvertPad without a converter-config: gst_video_converter_new (..., NULL)
     * (gstvideoaggregator.c:508-513) — every option at its default */
    gst_structure_free (options);
    options = NULL;
  }
  r->convert = gst_video_converter_new (&r->in_info, &r->out_info, options);
  if (!r->convert) {
    g_free (r);
    return NULL;
  }
  return r;
}

void
ref_vcs_convert (RefVcs * r, const guint8 * in, guint8 * out)
{
  GstVideoFrame src, dst;
  guint i;
  memset (&src, 0, sizeof (src));
  memset (&dst, 0, sizeof (dst));
  src.info = r->in_info;
  dst.info = r->out_info;
  for (i = 0; i < GST_VIDEO_INFO_N_PLANES (&r->in_info); i++)
    src.data[i] = (gpointer) (in + r->in_info.offset[i]);
  for (i = 0; i < GST_VIDEO_INFO_N_PLANES (&r->out_info); i++)
    dst.data[i] = out + r->out_info.offset[i];
  gst_video_converter_frame (r->convert, &src, &dst);
}

void
ref_vcs_free (RefVcs * r)
{
  if (!r)
    return;
  gst_video_converter_free (r->convert);
  g_free (r);
}

/* --- direct access to the reference scaler tables, for oracle tap-table checks --------
 * gst_video_scaler_new / get_coeff: gst-libs/gst/video/video-scaler.c:209, :313 */
int
ref_scaler_taps (int resampler_method, int max_taps_opt, int in_size, int out_size,
    double envelope, double sharpness, double sharpen, int *n_taps, unsigned *offsets,
    double *taps /* out_size * 128 */ )
{
  GstStructure *opts = gst_structure_new_static_str_empty ("o");
  GstVideoScaler *sc;
  guint i, nt = 0, off;
  gst_structure_set_static_str (opts,
      GST_VIDEO_RESAMPLER_OPT_ENVELOPE, G_TYPE_DOUBLE, envelope,
      GST_VIDEO_RESAMPLER_OPT_SHARPNESS, G_TYPE_DOUBLE, sharpness,
      GST_VIDEO_RESAMPLER_OPT_SHARPEN, G_TYPE_DOUBLE, sharpen, NULL);
  if (max_taps_opt > 0)
    gst_structure_set_static_str (opts, GST_VIDEO_RESAMPLER_OPT_MAX_TAPS, G_TYPE_INT,
        max_taps_opt, NULL);
  sc = gst_video_scaler_new ((GstVideoResamplerMethod) resampler_method,
      GST_VIDEO_SCALER_FLAG_NONE, 0, in_size, out_size, opts);
  for (i = 0; i < (guint) out_size; i++) {
    const gdouble *t = gst_video_scaler_get_coeff (sc, i, &off, &nt);
    offsets[i] = off;
    memcpy (taps + (size_t) i * nt, t, sizeof (double) * nt);
  }
  *n_taps = (int) nt;
  gst_video_scaler_free (sc);
  gst_structure_free (opts);
  return 0;
}
