// tests/cudaemu/emu/vcs_ntap.cuh — TEST INFRASTRUCTURE: the n-tap kernel (dp4a PTX) is not emulated.
#pragma once
#include "common.h"
#include "vcs_device.h"
#include "vcs_plan.h"
namespace b200 {
struct NtapState { int *d_h = nullptr, *d_v = nullptr; bool ready = false; };
inline int prepare_ntap (const VcsPlan &, NtapState *) { return B200_OK; }
inline int launch_ntap (const VcsDev &, const VcsPlan &, const NtapState &, const VcsBatch &, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
}  // namespace b200
