/* oracle/refshim/refdrv_comp.c — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's own blend kernels (gst/compositor/blend.c + the ORC C backups in
 * compositororc-dist.c, compiled in place) in the call sequence of
 * blend_pads()/_draw_background() (gst/compositor/compositor.c:1619-1697): background
 * fill over [0,height), then every pad in z-order through the `blend` family (opaque
 * backgrounds) or the `overlay` family (transparent background).
 */
#include <gst/video/video.h>
#include "blend.h"

typedef struct
{
  const guint8 *data; int width, height, stride;
  int xpos, ypos; double alpha; int op;
} RefPad;

static void
fill_frame (GstVideoFrame * f, int format, guint8 * data, int width, int height, int stride)
{
  memset (f, 0, sizeof (*f));
  gst_video_info_set_format (&f->info, (GstVideoFormat) format, width, height);
  f->info.stride[0] = stride;
  f->data[0] = data;
}

int
ref_compositor (int out_format, guint8 * dst, int width, int height, int stride, int background,
    const RefPad * pads, int n_pads)
{
  static gsize inited = 0;
  GstVideoFrame out;
  BlendFunction blend, overlay, composite;
  FillCheckerFunction fill_checker;
  FillColorFunction fill_color;
  int i;
  if (!inited) {
    gst_compositor_init_blend ();
    inited = 1;
  }
  switch (out_format) {         /* compositor.c:844-867 */
    case GST_VIDEO_FORMAT_ARGB:
      blend = gst_compositor_blend_argb; overlay = gst_compositor_overlay_argb;
      fill_checker = gst_compositor_fill_checker_argb; fill_color = gst_compositor_fill_color_argb;
      break;
    case GST_VIDEO_FORMAT_BGRA:
      blend = gst_compositor_blend_bgra; overlay = gst_compositor_overlay_bgra;
      fill_checker = gst_compositor_fill_checker_bgra; fill_color = gst_compositor_fill_color_bgra;
      break;
    case GST_VIDEO_FORMAT_ABGR:
      blend = gst_compositor_blend_abgr; overlay = gst_compositor_overlay_abgr;
      fill_checker = gst_compositor_fill_checker_abgr; fill_color = gst_compositor_fill_color_abgr;
      break;
    case GST_VIDEO_FORMAT_RGBA:
      blend = gst_compositor_blend_rgba; overlay = gst_compositor_overlay_rgba;
      fill_checker = gst_compositor_fill_checker_rgba; fill_color = gst_compositor_fill_color_rgba;
      break;
    default:
      return -1;
  }
  fill_frame (&out, out_format, dst, width, height, stride);
  composite = blend;
  switch (background) {         /* compositor.c:1619-1675; RGB 0-255: black 0,0,0 white 255,255,255 (:1131-1149) */
    case 0: fill_checker (&out, 0, height); break;
    case 1: fill_color (&out, 0, height, 0, 0, 0); break;
    case 2: fill_color (&out, 0, height, 255, 255, 255); break;
    default:
      for (i = 0; i < height; i++)
        memset (dst + (gsize) i * stride, 0, (gsize) width * 4);
      composite = overlay;
      break;
  }
  for (i = 0; i < n_pads; i++) {
    GstVideoFrame src;
    fill_frame (&src, out_format, (guint8 *) pads[i].data, pads[i].width, pads[i].height,
        pads[i].stride);
    /* COMPOSITOR_BLEND_MODE_* share the operator's numbering (blend.h, compositor.c:1800-1813) */
    composite (&src, pads[i].xpos, pads[i].ypos, pads[i].alpha, &out, 0, height,
        (GstCompositorBlendMode) pads[i].op);
  }
  return 0;
}


/* 4:2:0 formats (I420, YV12, NV12, NV21): every pad frame has the output's format, default plane
 * layouts (gst_video_info_set_format); range_16_235 selects the black/white levels the way
 * gst_video_color_range_offsets does (compositor.c:1131-1149). */
int
ref_compositor_yuv (int format, guint8 * dst, int width, int height, int background, int range_16_235,
    const RefPad * pads, int n_pads)
{
  static gsize inited = 0;
  GstVideoFrame out;
  BlendFunction blend;
  FillCheckerFunction fill_checker;
  FillColorFunction fill_color;
  int i, p;
  if (!inited) {
    gst_compositor_init_blend ();
    inited = 1;
  }
  switch (format) {             /* compositor.c:982-1005 */
    case GST_VIDEO_FORMAT_I420:
      blend = gst_compositor_blend_i420; fill_checker = gst_compositor_fill_checker_i420; fill_color = gst_compositor_fill_color_i420;
      break;
    case GST_VIDEO_FORMAT_YV12:
      blend = gst_compositor_blend_yv12; fill_checker = gst_compositor_fill_checker_yv12; fill_color = gst_compositor_fill_color_yv12;
      break;
    case GST_VIDEO_FORMAT_NV12:
      blend = gst_compositor_blend_nv12; fill_checker = gst_compositor_fill_checker_nv12; fill_color = gst_compositor_fill_color_nv12;
      break;
    case GST_VIDEO_FORMAT_NV21:
      blend = gst_compositor_blend_nv21; fill_checker = gst_compositor_fill_checker_nv21; fill_color = gst_compositor_fill_color_nv21;
      break;
    case GST_VIDEO_FORMAT_Y444:
      blend = gst_compositor_blend_y444; fill_checker = gst_compositor_fill_checker_y444; fill_color = gst_compositor_fill_color_y444;
      break;
    case GST_VIDEO_FORMAT_Y42B:
      blend = gst_compositor_blend_y42b; fill_checker = gst_compositor_fill_checker_y42b; fill_color = gst_compositor_fill_color_y42b;
      break;
#define HIGH(fmt, name) case GST_VIDEO_FORMAT_##fmt: \
      blend = gst_compositor_blend_##name; fill_checker = gst_compositor_fill_checker_##name; fill_color = gst_compositor_fill_color_##name; break;
    HIGH (I420_10LE, i420_10le) HIGH (I420_12LE, i420_12le) HIGH (I422_10LE, i422_10le) HIGH (I422_12LE, i422_12le)
    HIGH (Y444_10LE, y444_10le) HIGH (Y444_12LE, y444_12le) HIGH (Y444_16LE, y444_16le)
#undef HIGH
    default:
      return -1;
  }
  memset (&out, 0, sizeof (out));
  gst_video_info_set_format (&out.info, (GstVideoFormat) format, width, height);
  for (p = 0; p < (int) GST_VIDEO_INFO_N_PLANES (&out.info); p++)
    out.data[p] = dst + out.info.offset[p];
  switch (background) {
    case 0: fill_checker (&out, 0, height); break;
    case 1: case 2: {
      /* compositor.c:1131-1149: black / white from the reference's own range offsets at the format's depth */
      gint offset[GST_VIDEO_MAX_COMPONENTS], scale[GST_VIDEO_MAX_COMPONENTS];
      gst_video_color_range_offsets (range_16_235 ? GST_VIDEO_COLOR_RANGE_16_235 : GST_VIDEO_COLOR_RANGE_0_255, out.info.finfo,
          offset, scale);
      fill_color (&out, 0, height, background == 1 ? offset[0] : scale[0] + offset[0], offset[1], offset[2]);
      break;
    }
    default:
      /* the element's own loop is a static function (_draw_background, compositor.c:1640-1670): the visible bytes of
       * every plane row are zeroed, the stride padding is left alone; overlay == blend for these formats */
      for (p = 0; p < (int) GST_VIDEO_INFO_N_PLANES (&out.info); p++) {
        gint comp[GST_VIDEO_MAX_COMPONENTS], row;
        gst_video_format_info_component (out.info.finfo, p, comp);
        for (row = 0; row < GST_VIDEO_FRAME_COMP_HEIGHT (&out, comp[0]); row++)
          memset ((guint8 *) out.data[p] + (gsize) row * GST_VIDEO_FRAME_PLANE_STRIDE (&out, p), 0,
              GST_VIDEO_FRAME_COMP_WIDTH (&out, comp[0]) * GST_VIDEO_FRAME_COMP_PSTRIDE (&out, comp[0]));
      }
      break;
  }
  for (i = 0; i < n_pads; i++) {
    GstVideoFrame src;
    memset (&src, 0, sizeof (src));
    gst_video_info_set_format (&src.info, (GstVideoFormat) format, pads[i].width, pads[i].height);
    for (p = 0; p < (int) GST_VIDEO_INFO_N_PLANES (&src.info); p++)
      src.data[p] = (guint8 *) pads[i].data + src.info.offset[p];
    blend (&src, pads[i].xpos, pads[i].ypos, pads[i].alpha, &out, 0, height, (GstCompositorBlendMode) pads[i].op);
  }
  return 0;
}
