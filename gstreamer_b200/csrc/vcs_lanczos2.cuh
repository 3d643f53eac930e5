// gstreamer_b200/csrc/vcs_lanczos2.cuh — specialised 2:1 lanczos NV12->RGB kernel (product).
// Placeholder until the specialised path lands; the generic kernel covers every plan.
#pragma once
#include "common.h"
#include "vcs_device.h"
namespace b200 {
inline int prepare_lanczos2 (const VcsDev &, int) { return B200_ERR_UNSUPPORTED; }
inline int launch_lanczos2 (const VcsDev &, const VcsBatch &, int, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
}
