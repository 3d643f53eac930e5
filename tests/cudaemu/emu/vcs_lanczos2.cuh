// tests/cudaemu/emu/vcs_lanczos2.cuh — TEST INFRASTRUCTURE: stand-in for the PTX / warp-shuffle kernel header.
// Keeps the packed-byte helpers other kernels borrow and reports the specialised kernel as not eligible.
#pragma once
#include <vector>
#include "common.h"
#include "vcs_device.h"
#include "vcs_plan.h"

namespace b200 {

inline unsigned avg_floor4 (unsigned a, unsigned b) { return (a & b) + (((a ^ b) & 0xfefefefeu) >> 1); }
inline unsigned avg_ceil4 (unsigned a, unsigned b) { return (a | b) - (((a ^ b) & 0xfefefefeu) >> 1); }

struct Lanczos2Tables { bool ok = false; bool alpha_opaque = false; };
struct Lanczos2State { int4 *d_htab = nullptr, *d_vtab = nullptr; int variant = 0; };
inline Lanczos2Tables build_lanczos2_tables (const VcsPlan &) { return Lanczos2Tables (); }
inline int prepare_lanczos2 (const Lanczos2Tables &, const VcsDev &, Lanczos2State *) { return B200_ERR_UNSUPPORTED; }
inline int launch_lanczos2 (const VcsDev &, const Lanczos2State &, const VcsBatch &, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }

}  // namespace b200
