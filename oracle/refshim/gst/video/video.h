/* refshim: <gst/video/video.h> reduced to the headers the arithmetic sources
 * need (the real umbrella header pulls in every GObject-based class of
 * libgstvideo).  The included headers are the reference's own, found through
 * -I$(REF)/gst-plugins-base/gst-libs. */
#ifndef __GST_VIDEO_H__
#define __GST_VIDEO_H__
#include <gst/gst.h>
#include <gst/video/video-prelude.h>
typedef struct _GstVideoAlignment GstVideoAlignment;
typedef struct _GstVideoRectangle GstVideoRectangle;
#include <gst/video/video-format.h>
#include <gst/video/video-color.h>
#include <gst/video/video-dither.h>
#include <gst/video/video-info.h>
#include <gst/video/video-frame.h>
#include <gst/video/video-enumtypes.h>
#include <gst/video/video-converter.h>
#include <gst/video/video-scaler.h>
#include <gst/video/video-multiview.h>
G_BEGIN_DECLS
struct _GstVideoAlignment { guint padding_top, padding_bottom, padding_left, padding_right;
  guint stride_align[GST_VIDEO_MAX_PLANES]; };
struct _GstVideoRectangle { gint x, y, w, h; };
/* meta plumbing referenced (never executed) by gst_video_converter_transform_metas() */
typedef struct { GstVideoInfo *in_info; GstVideoInfo *out_info; } GstVideoMetaTransform;
typedef struct { const GstVideoInfo *in_info; GstVideoRectangle in_rectangle;
  const GstVideoInfo *out_info; GstVideoRectangle out_rectangle; gfloat matrix[3][3]; } GstVideoMetaTransformMatrix;
GQuark gst_video_meta_transform_matrix_get_quark (void);
GQuark gst_video_meta_transform_scale_get_quark (void);
void gst_video_meta_transform_matrix_init (GstVideoMetaTransformMatrix * trans,
    const GstVideoInfo * in_info, const GstVideoRectangle * in_rectangle,
    const GstVideoInfo * out_info, const GstVideoRectangle * out_rectangle);
#define GST_META_TAG_VIDEO_STR "video"
#define GST_META_TAG_VIDEO_ORIENTATION_STR "orientation"
#define GST_META_TAG_VIDEO_SIZE_STR "size"
#define GST_META_TAG_VIDEO_COLORSPACE_STR "colorspace"
G_END_DECLS
#endif
