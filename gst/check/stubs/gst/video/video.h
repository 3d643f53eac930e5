/* declarations only: GstVideoInfo / GstVideoFrame accessors as in gst-plugins-base/gst-libs/gst/video/video-info.h,
 * video-frame.h, video-format.h, video-converter.h (enums the element's properties use) */
#ifndef B200_STUB_VIDEO_H
#define B200_STUB_VIDEO_H
#include <gst/gst.h>
#define GST_VIDEO_MAX_PLANES 4
typedef enum { GST_VIDEO_FORMAT_UNKNOWN = 0 } GstVideoFormat;
typedef struct _GstVideoFormatInfo { GstVideoFormat format; guint flags; } GstVideoFormatInfo;
typedef struct { gint range, matrix, transfer, primaries; } GstVideoColorimetry;
typedef struct _GstVideoInfo {
  const GstVideoFormatInfo *finfo; gint interlace_mode; guint flags; gint width, height; gsize size; gint views;
  gint chroma_site; GstVideoColorimetry colorimetry; gint par_n, par_d, fps_n, fps_d;
  gsize offset[GST_VIDEO_MAX_PLANES]; gint stride[GST_VIDEO_MAX_PLANES];
} GstVideoInfo;
typedef struct _GstVideoFrame { GstVideoInfo info; guint flags; GstBuffer *buffer; gpointer meta; gint id;
  gpointer data[GST_VIDEO_MAX_PLANES]; GstMapInfo map[GST_VIDEO_MAX_PLANES]; } GstVideoFrame;
#define GST_VIDEO_INFO_FORMAT(i) ((i)->finfo->format)
#define GST_VIDEO_INFO_WIDTH(i) ((i)->width)
#define GST_VIDEO_INFO_HEIGHT(i) ((i)->height)
#define GST_VIDEO_INFO_SIZE(i) ((i)->size)
#define GST_VIDEO_INFO_N_PLANES(i) 3u
#define GST_VIDEO_INFO_PLANE_STRIDE(i, p) ((i)->stride[p])
#define GST_VIDEO_INFO_PLANE_OFFSET(i, p) ((i)->offset[p])
#define GST_VIDEO_INFO_IS_YUV(i) (((i)->finfo->flags & 1) != 0)
#define GST_VIDEO_FORMAT_INFO_IS_YUV(f) (((f)->flags & 1) != 0)
#define GST_VIDEO_FRAME_PLANE_DATA(f, p) ((f)->data[p])
#define GST_VIDEO_FRAME_PLANE_STRIDE(f, p) ((f)->info.stride[p])
#define GST_VIDEO_FRAME_WIDTH(f) ((f)->info.width)
#define GST_VIDEO_FRAME_HEIGHT(f) ((f)->info.height)
gboolean gst_video_info_from_caps (GstVideoInfo * info, const GstCaps * caps);
gboolean gst_video_frame_map (GstVideoFrame * frame, const GstVideoInfo * info, GstBuffer * buffer, GstMapFlags flags);
void gst_video_frame_unmap (GstVideoFrame * frame);
GstVideoFormat gst_video_format_from_string (const gchar * format);
const GstVideoFormatInfo *gst_video_format_get_info (GstVideoFormat format);
GstBufferPool *gst_video_buffer_pool_new (void);
#define GST_BUFFER_POOL_OPTION_VIDEO_META "GstBufferPoolOptionVideoMeta"
GType gst_video_meta_api_get_type (void);
#define GST_VIDEO_META_API_TYPE (gst_video_meta_api_get_type ())
typedef struct { gint x, y, w, h; } GstVideoRectangle;
void gst_video_center_rect (const GstVideoRectangle * src, const GstVideoRectangle * dst, GstVideoRectangle * result, gboolean scaling);
/* property enums (video-converter.h:177-262, video-resampler.h:49, video-dither.h:42) and their GTypes */
enum { GST_VIDEO_RESAMPLER_METHOD_NEAREST, GST_VIDEO_RESAMPLER_METHOD_LINEAR, GST_VIDEO_RESAMPLER_METHOD_CUBIC };
enum { GST_VIDEO_ALPHA_MODE_COPY, GST_VIDEO_ALPHA_MODE_SET, GST_VIDEO_ALPHA_MODE_MULT };
enum { GST_VIDEO_CHROMA_MODE_FULL, GST_VIDEO_CHROMA_MODE_UPSAMPLE_ONLY };
enum { GST_VIDEO_MATRIX_MODE_FULL, GST_VIDEO_MATRIX_MODE_INPUT_ONLY };
enum { GST_VIDEO_GAMMA_MODE_NONE, GST_VIDEO_GAMMA_MODE_REMAP };
enum { GST_VIDEO_PRIMARIES_MODE_NONE, GST_VIDEO_PRIMARIES_MODE_MERGE_ONLY };
enum { GST_VIDEO_DITHER_NONE, GST_VIDEO_DITHER_VERTERR, GST_VIDEO_DITHER_FLOYD_STEINBERG, GST_VIDEO_DITHER_SIERRA_LITE, GST_VIDEO_DITHER_BAYER };
GType gst_video_resampler_method_get_type (void); GType gst_video_alpha_mode_get_type (void);
GType gst_video_chroma_mode_get_type (void); GType gst_video_matrix_mode_get_type (void);
GType gst_video_gamma_mode_get_type (void); GType gst_video_primaries_mode_get_type (void);
GType gst_video_dither_method_get_type (void);
#define GST_TYPE_VIDEO_RESAMPLER_METHOD (gst_video_resampler_method_get_type ())
#define GST_TYPE_VIDEO_ALPHA_MODE (gst_video_alpha_mode_get_type ())
#define GST_TYPE_VIDEO_CHROMA_MODE (gst_video_chroma_mode_get_type ())
#define GST_TYPE_VIDEO_MATRIX_MODE (gst_video_matrix_mode_get_type ())
#define GST_TYPE_VIDEO_GAMMA_MODE (gst_video_gamma_mode_get_type ())
#define GST_TYPE_VIDEO_PRIMARIES_MODE (gst_video_primaries_mode_get_type ())
#define GST_TYPE_VIDEO_DITHER_METHOD (gst_video_dither_method_get_type ())
#endif
