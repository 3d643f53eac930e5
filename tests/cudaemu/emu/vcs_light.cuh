// tests/cudaemu/emu/vcs_light.cuh — TEST INFRASTRUCTURE: the light kernel (PTX, shuffles) is not emulated.
#pragma once
#include "common.h"
#include "vcs_device.h"
#include "vcs_plan.h"
namespace b200 {
inline void (*light_kernel_for (const VcsPlan &)) () { return nullptr; }
inline int launch_light (const VcsDev &, const VcsPlan &, const VcsBatch &, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
}  // namespace b200
