/* gst/gstb200elements.h — GTypes of the three elements and shared helpers. */
#ifndef GST_B200_ELEMENTS_H
#define GST_B200_ELEMENTS_H

#include <gst/gst.h>
#include <gst/video/video.h>
#include <b200dsp.h>

G_BEGIN_DECLS

#define GST_TYPE_CUDA_VIDEO_CONVERT_SCALE (gst_cuda_video_convert_scale_get_type ())
GType gst_cuda_video_convert_scale_get_type (void);
#define GST_TYPE_B200_CUDA_COMPOSITOR (gst_b200_cuda_compositor_get_type ())
GType gst_b200_cuda_compositor_get_type (void);
#define GST_TYPE_CUDA_AUDIO_RESAMPLE (gst_cuda_audio_resample_get_type ())
GType gst_cuda_audio_resample_get_type (void);

/* GstVideoInfo -> b200_video_info: every enum value is shared with GStreamer, so this is a field
 * copy.  Plane offsets are taken relative to plane 0 of the mapped frame. */
static inline void
gst_b200_video_info_from_gst (b200_video_info * d, const GstVideoInfo * s)
{
  guint i;
  memset (d, 0, sizeof (*d));
  d->format = GST_VIDEO_INFO_FORMAT (s);
  d->width = GST_VIDEO_INFO_WIDTH (s);
  d->height = GST_VIDEO_INFO_HEIGHT (s);
  for (i = 0; i < GST_VIDEO_INFO_N_PLANES (s) && i < B200_VIDEO_MAX_PLANES; i++) {
    d->stride[i] = GST_VIDEO_INFO_PLANE_STRIDE (s, i);
    d->offset[i] = GST_VIDEO_INFO_PLANE_OFFSET (s, i) - GST_VIDEO_INFO_PLANE_OFFSET (s, 0);
  }
  d->color_matrix = s->colorimetry.matrix;
  d->color_range = s->colorimetry.range;
  d->chroma_site = s->chroma_site;
}

/* b200_status -> what the reference would do at the same point (SURVEY §8b error convention) */
#define GST_B200_FLOW_FROM_STATUS(elem, st, what) G_STMT_START {                              \
  if ((st) != B200_OK) {                                                                      \
    GST_ELEMENT_ERROR (elem, LIBRARY, FAILED, ("%s failed: %s", what, b200_strerror (st)),    \
        ("%s", b200_last_cuda_error ()));                                                     \
    return GST_FLOW_ERROR;                                                                    \
  }                                                                                           \
} G_STMT_END

G_END_DECLS
#endif
