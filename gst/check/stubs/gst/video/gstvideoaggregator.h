/* declarations only: gst-plugins-base/gst-libs/gst/video/gstvideoaggregator.h:71-330 and gstreamer/libs/gst/base/gstaggregator.h */
#ifndef B200_STUB_VIDEOAGG_H
#define B200_STUB_VIDEOAGG_H
#include <gst/video/video.h>
typedef struct _GstAggregator { GstElement parent; GstPad *srcpad; } GstAggregator;
typedef struct _GstAggregatorClass {
  GstElementClass parent_class;
  gboolean (*start) (GstAggregator * aggregator);
  gboolean (*stop) (GstAggregator * aggregator);
  gboolean (*src_query) (GstAggregator * aggregator, GstQuery * query);
  gboolean (*sink_query) (GstAggregator * aggregator, gpointer aggregator_pad, GstQuery * query);
  gboolean (*decide_allocation) (GstAggregator * self, GstQuery * query);
  gboolean (*propose_allocation) (GstAggregator * self, gpointer pad, GstQuery * decide_query, GstQuery * query);
} GstAggregatorClass;
#define GST_AGGREGATOR_CLASS(k) ((GstAggregatorClass *) (k))
#define GST_TYPE_AGGREGATOR_PAD ((GType) 0x202)
typedef struct _GstVideoAggregator { GstAggregator aggregator; GstVideoInfo info; } GstVideoAggregator;
typedef struct _GstVideoAggregatorPad { GstObject parent; GstVideoInfo info; } GstVideoAggregatorPad;
typedef struct _GstVideoAggregatorPadClass {
  GObjectClass parent_class;
  void (*update_conversion_info) (GstVideoAggregatorPad * pad);
  gboolean (*prepare_frame) (GstVideoAggregatorPad * pad, GstVideoAggregator * vagg, GstBuffer * buffer, GstVideoFrame * prepared_frame);
  void (*clean_frame) (GstVideoAggregatorPad * pad, GstVideoAggregator * vagg, GstVideoFrame * prepared_frame);
} GstVideoAggregatorPadClass;
typedef struct _GstVideoAggregatorClass {
  GstAggregatorClass parent_class;
  GstCaps *(*update_caps) (GstVideoAggregator * vagg, GstCaps * caps);
  GstFlowReturn (*aggregate_frames) (GstVideoAggregator * vagg, GstBuffer * outbuf);
  GstFlowReturn (*create_output_buffer) (GstVideoAggregator * vagg, GstBuffer ** outbuffer);
  void (*find_best_format) (GstVideoAggregator * vagg, GstCaps * downstream_caps, GstVideoInfo * best_info, gboolean * at_least_one_alpha);
} GstVideoAggregatorClass;
#define GST_TYPE_VIDEO_AGGREGATOR ((GType) 0x200)
#define GST_TYPE_VIDEO_AGGREGATOR_PAD ((GType) 0x201)
#define GST_VIDEO_AGGREGATOR_CLASS(k) ((GstVideoAggregatorClass *) (k))
#define GST_VIDEO_AGGREGATOR_PAD(p) ((GstVideoAggregatorPad *) (p))
#define GST_VIDEO_AGGREGATOR_PAD_CLASS(k) ((GstVideoAggregatorPadClass *) (k))
GstVideoFrame *gst_video_aggregator_pad_get_prepared_frame (GstVideoAggregatorPad * pad);
gboolean gst_video_aggregator_pad_has_current_buffer (GstVideoAggregatorPad * pad);
GstBuffer *gst_video_aggregator_pad_get_current_buffer (GstVideoAggregatorPad * pad);
#endif
