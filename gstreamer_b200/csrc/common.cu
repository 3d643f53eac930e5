// gstreamer_b200/csrc/common.cu — status strings, device helpers, pinned host memory.
#include "common.h"

#include <stdio.h>
#include <string.h>

namespace b200 {

static thread_local char g_last_cuda[512] = "";

int cuda_fail (cudaError_t e, const char *what, const char *file, int line)
{
  snprintf (g_last_cuda, sizeof (g_last_cuda), "%s:%d: %s -> %s (%s)", file, line, what,
      cudaGetErrorName (e), cudaGetErrorString (e));
  // clear the sticky-less error state so later calls report their own failures
  cudaGetLastError ();
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
    return B200_ERR_NO_DEVICE;
  if (e == cudaErrorMemoryAllocation)
    return B200_ERR_NOMEM;
  return B200_ERR_CUDA;
}

int sm_count (int device)
{
  static int cache[64];
  if (device < 0 || device >= 64) return 148;
  if (cache[device] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0)
      n = 148;
    cache[device] = n;
  }
  return cache[device];
}

}  // namespace b200

extern "C" {

const char *b200_strerror (int status)
{
  switch (status) {
    case B200_OK: return "ok";
    case B200_ERR_INVALID_ARG: return "invalid argument";
    case B200_ERR_UNSUPPORTED: return "unsupported format or mode";
    case B200_ERR_NO_DEVICE: return "no CUDA device (libb200dsp has no CPU fallback)";
    case B200_ERR_CUDA: return "CUDA call failed";
    case B200_ERR_NOMEM: return "out of memory";
    case B200_ERR_STATE: return "invalid handle state";
    default: return "unknown status";
  }
}

const char *b200_last_cuda_error (void) { return b200::g_last_cuda; }

int b200_version (void) { return B200DSP_VERSION_MAJOR * 100 + B200DSP_VERSION_MINOR; }

int b200_device_count (void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount (&n);
  if (e != cudaSuccess)
    return b200::cuda_fail (e, "cudaGetDeviceCount", __FILE__, __LINE__);
  return n;
}

int b200_host_alloc (size_t size, void **ptr)
{
  if (!ptr || size == 0) return B200_ERR_INVALID_ARG;
  B200_CUDA_TRY (cudaHostAlloc (ptr, size, cudaHostAllocPortable));
  return B200_OK;
}

int b200_host_free (void *ptr)
{
  if (!ptr) return B200_OK;
  B200_CUDA_TRY (cudaFreeHost (ptr));
  return B200_OK;
}

}  // extern "C"
