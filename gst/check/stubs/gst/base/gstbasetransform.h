/* declarations only: GstBaseTransformClass vmethods as in gstreamer/libs/gst/base/gstbasetransform.h:220-300 */
#ifndef B200_STUB_BASETRANSFORM_H
#define B200_STUB_BASETRANSFORM_H
#include <gst/gst.h>
typedef struct _GstBaseTransform { GstElement element; GstPad *sinkpad, *srcpad; } GstBaseTransform;
typedef struct _GstBaseTransformClass {
  GstElementClass parent_class;
  gboolean passthrough_on_same_caps, transform_ip_on_passthrough;
  GstCaps *(*transform_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, GstCaps * filter);
  GstCaps *(*fixate_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, GstCaps * othercaps);
  gboolean (*accept_caps) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps);
  gboolean (*set_caps) (GstBaseTransform * trans, GstCaps * incaps, GstCaps * outcaps);
  gboolean (*query) (GstBaseTransform * trans, GstPadDirection direction, GstQuery * query);
  gboolean (*decide_allocation) (GstBaseTransform * trans, GstQuery * query);
  gboolean (*filter_meta) (GstBaseTransform * trans, GstQuery * query, GType api, const GstStructure * params);
  gboolean (*propose_allocation) (GstBaseTransform * trans, GstQuery * decide_query, GstQuery * query);
  gboolean (*transform_size) (GstBaseTransform * trans, GstPadDirection direction, GstCaps * caps, gsize size, GstCaps * othercaps, gsize * othersize);
  gboolean (*get_unit_size) (GstBaseTransform * trans, GstCaps * caps, gsize * size);
  gboolean (*start) (GstBaseTransform * trans);
  gboolean (*stop) (GstBaseTransform * trans);
  gboolean (*sink_event) (GstBaseTransform * trans, GstEvent * event);
  gboolean (*src_event) (GstBaseTransform * trans, GstEvent * event);
  GstFlowReturn (*prepare_output_buffer) (GstBaseTransform * trans, GstBuffer * input, GstBuffer ** outbuf);
  gboolean (*copy_metadata) (GstBaseTransform * trans, GstBuffer * input, GstBuffer * outbuf);
  gboolean (*transform_meta) (GstBaseTransform * trans, GstBuffer * outbuf, gpointer meta, GstBuffer * inbuf);
  void (*before_transform) (GstBaseTransform * trans, GstBuffer * buffer);
  GstFlowReturn (*transform) (GstBaseTransform * trans, GstBuffer * inbuf, GstBuffer * outbuf);
  GstFlowReturn (*transform_ip) (GstBaseTransform * trans, GstBuffer * buf);
} GstBaseTransformClass;
#define GST_TYPE_BASE_TRANSFORM ((GType) 0x100)
#define GST_BASE_TRANSFORM(o) ((GstBaseTransform *) (o))
#define GST_BASE_TRANSFORM_CLASS(k) ((GstBaseTransformClass *) (k))
#define GST_BASE_TRANSFORM_SRC_PAD(t) (((GstBaseTransform *) (t))->srcpad)
#define GST_BASE_TRANSFORM_FLOW_DROPPED GST_FLOW_CUSTOM_SUCCESS
void gst_base_transform_set_passthrough (GstBaseTransform * trans, gboolean passthrough);
gboolean gst_base_transform_is_passthrough (GstBaseTransform * trans);
#endif
