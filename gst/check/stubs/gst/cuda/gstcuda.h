/* declarations only: gst-plugins-bad/gst-libs/gst/cuda/gstcudacontext.h, gstcudamemory.h, gstcudastream.h,
 * gstcudabufferpool.h, gstcudautils.h, cuda-gst.h */
#ifndef B200_STUB_GSTCUDA_H
#define B200_STUB_GSTCUDA_H
#include <gst/gst.h>
typedef struct _GstCudaContext GstCudaContext; typedef struct _GstCudaStream GstCudaStream; typedef struct _GstCudaMemory GstCudaMemory;
typedef void *CUstream; typedef int CUresult;
#define GST_CAPS_FEATURE_MEMORY_CUDA_MEMORY "memory:CUDAMemory"
#define GST_MAP_CUDA (GST_MAP_FLAG_LAST << 1)
#define GST_CUDA_MEMORY_CAST(m) ((GstCudaMemory *) (m))
#define GST_CUDA_MEMORY_TRANSFER_NEED_SYNC 0
gboolean gst_is_cuda_memory (GstMemory * mem);
GstCudaStream *gst_cuda_memory_get_stream (GstCudaMemory * mem);
void gst_cuda_memory_sync (GstCudaMemory * mem);
gboolean gst_cuda_context_push (GstCudaContext * ctx);
gboolean gst_cuda_context_pop (gpointer * cuda_ctx);
GstCudaStream *gst_cuda_stream_new (GstCudaContext * context);
CUstream gst_cuda_stream_get_handle (GstCudaStream * stream);
void gst_clear_cuda_stream (GstCudaStream ** stream);
gboolean gst_cuda_ensure_element_context (GstElement * element, gint device_id, GstCudaContext ** cuda_ctx);
gboolean gst_cuda_handle_set_context (GstElement * element, GstContext * context, gint device_id, GstCudaContext ** cuda_ctx);
gboolean gst_cuda_handle_context_query (GstElement * element, GstQuery * query, GstCudaContext * cuda_ctx);
GstBufferPool *gst_cuda_buffer_pool_new (GstCudaContext * context);
gboolean b200_stub_is_cuda_pool (gpointer p);
#define GST_IS_CUDA_BUFFER_POOL(p) b200_stub_is_cuda_pool (p)
CUresult CuStreamSynchronize (CUstream stream);
#endif
