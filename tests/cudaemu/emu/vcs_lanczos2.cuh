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
inline int prmt_s (unsigned a, unsigned sel) { return (int) __byte_perm (a, 0, sel); }       // prmt.b32 incl. sign replication
inline int sra6 (int acc) { return acc >> 6; }
// cvt.pack.sat.u8.s32.b32 d, a, b, c: d = { c[15:0], sat_u8 (a), sat_u8 (b) }  (b in the lowest byte)
inline unsigned pack_sat2 (int a, int b, unsigned c)
{
  const unsigned sa = (unsigned) min (max (a, 0), 255), sb = (unsigned) min (max (b, 0), 255);
  return (c << 16) | (sa << 8) | sb;
}
constexpr int L2_THREADS = 256;

struct Lanczos2Tables { bool ok = false; bool alpha_opaque = false; };
struct Lanczos2State { int4 *d_htab = nullptr, *d_vtab = nullptr; int variant = 0; };
inline Lanczos2Tables build_lanczos2_tables (const VcsPlan &) { return Lanczos2Tables (); }
inline int prepare_lanczos2 (const Lanczos2Tables &, const VcsDev &, Lanczos2State *) { return B200_ERR_UNSUPPORTED; }
inline int launch_lanczos2 (const VcsDev &, const Lanczos2State &, const VcsBatch &, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }

}  // namespace b200
