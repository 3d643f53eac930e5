// tools/l2lab.cu — development lab for the headline kernel (NOT product code, never loaded by the library).
//
// Times instantiations of vcs_lanczos2_kernel on the bench shape (3840x2160 NV12 -> 1920x1080 BGRA, lanczos, 32 frames per
// launch out of a 64-frame ring >> L2) and checks every parity variant byte for byte against the product's default
// instantiation.  The ABL != 0 instantiations are NON-PARITY ablations: they answer "what does this stage cost" in
// microseconds instead of by counting instructions.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/l2lab.cu -o tools/l2lab \
//        -Lgstreamer_b200 -lb200dsp -Xlinker -rpath -Xlinker '$ORIGIN/../gstreamer_b200'
//   tools/l2lab [name-filter]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../gstreamer_b200/csrc/common.h"
#include "../gstreamer_b200/csrc/vcs_plan.h"
#include "../gstreamer_b200/csrc/vcs_device.h"
#include "../gstreamer_b200/csrc/vcs_kernels.cuh"
#include "../gstreamer_b200/csrc/vcs_lanczos2.cuh"
#include "../gstreamer_b200/csrc/vcs_lanczos2_v2.cuh"
#include "../gstreamer_b200/csrc/vcs_l2tc.cuh"
#ifdef L2LAB_EXTRA
#include L2LAB_EXTRA
#endif

using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf ("CUDA error %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); exit (1); } } while (0)

__global__ void fill_random (uint8_t * p, size_t n, unsigned seed)
{
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const size_t step = (size_t) gridDim.x * blockDim.x;
  for (; i < n / 4; i += step) {
    unsigned x = (unsigned) i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    ((unsigned *) p)[i] = x;
  }
}

struct Lab {
  VcsPlan plan;
  VcsDev dev;
  Lanczos2Tables tab;
  Lanczos2State st;
  int16_t *d_hsum = nullptr, *d_vsum = nullptr;
  static const int RING = 64, PER = 32;
  uint8_t *in[RING], *out[RING];
  size_t in_bytes, out_bytes;
  std::vector<uint8_t> ref;      // output of frame 0 from the reference instantiation
  L2tcTables tct;
  L2tcState tc;
  Lanczos2V2Tables v2t;
  Lanczos2V2State v2;
};

void launch_tc (Lab & L, const VcsBatch & b, cudaStream_t s)
{
  if (launch_l2tc (L.dev, L.tc, b, Lab::PER, s) != B200_OK) { printf ("launch_l2tc failed: %s\n", b200_last_cuda_error ()); exit (1); }
}

// H pass of the tensor kernel in isolation: raw accumulators of tile 0 (frame 0, strip 0, row tile 0) against a plain
// CPU evaluation of the same banded product for the luma plane
int tc_debug (Lab & L)
{
  unsigned *d_dbg; const size_t n = 3 * 128 * 64;
  CK (cudaMalloc ((void **) &d_dbg, n * 4)); CK (cudaMemset (d_dbg, 0xee, n * 4));
  VcsBatch b; b.in[0] = L.in[0]; b.out[0] = L.out[0];
  if (launch_l2tc (L.dev, L.tc, b, 1, 0, d_dbg, 1) != B200_OK) { printf ("launch failed: %s\n", b200_last_cuda_error ()); return 1; }
  cudaError_t e = cudaDeviceSynchronize ();
  if (e != cudaSuccess) { printf ("kernel error: %s\n", cudaGetErrorString (e)); return 1; }
  std::vector<unsigned> dbg (n); std::vector<uint8_t> frame (L.in_bytes);
  CK (cudaMemcpy (dbg.data (), d_dbg, n * 4, cudaMemcpyDeviceToHost));
  CK (cudaMemcpy (frame.data (), L.in[0], L.in_bytes, cudaMemcpyDeviceToHost));
  const VcsPlan & p = L.plan;
  const int scale = L.tct.hx4[0] ? 4 : 1, rnd = L.tct.hx4[0] ? 128 : 32;
  long bad = 0, checked = 0;
  for (int j = 0; j < 128; j++)
    for (int li = 3; li < 62; li++) {                  // lines 0..2 are above the frame in row tile 0
      const int y = li - 3;
      int acc = rnd;
      for (int k = 0; k < 8; k++)
        acc += scale * p.h.coef[(size_t) j * 8 + k] * frame[p.in.offset[0] + (size_t) y * p.in.stride[0] + p.h.offset[j] + k];
      const int got = (int) dbg[(0 * 128 + j) * 64 + li];
      checked++;
      if (got != acc) { if (bad < 8) printf ("  Y mismatch col %d line %d: got %d want %d\n", j, li, got, acc); bad++; }
    }
  printf ("tc_debug: H pass luma accumulators of tile 0: %ld of %ld differ (x4 %d)\n", bad, checked, (int) L.tct.hx4[0]);
  printf ("  sample U accumulators col 5: %d %d %d %d\n", (int) dbg[(128 + 5) * 64 + 10], (int) dbg[(128 + 5) * 64 + 11], (int) dbg[(128 + 5) * 64 + 12], (int) dbg[(128 + 5) * 64 + 13]);
  cudaFree (d_dbg);
  return bad != 0;
}

typedef void (*launch_fn) (Lab & L, const VcsBatch & b, cudaStream_t s);

template <int SEL, int PF, int MINB = 4, int TH = 60, int NT = 256>
void launch_v2 (Lab & L, const VcsBatch & b, cudaStream_t s)
{
  if (launch_lanczos2_v2_sel<SEL, PF, MINB, TH, NT> (L.dev, L.st, L.v2, b, Lab::PER, s) != B200_OK) { printf ("launch_v2 failed: %s\n", b200_last_cuda_error ()); exit (1); }
}


template <int MINB, int TH, int NWC, bool X4, int ABL>
void launch_l2 (Lab & L, const VcsBatch & b, cudaStream_t s)
{
  auto kern = vcs_lanczos2_kernel<true, MINB, TH, NWC, X4, ABL>;
  static bool done = false;
  if (!done) { CK (cudaFuncSetAttribute (kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L2Shape<TH, NWC>::SMEM)); done = true; }
  dim3 grid ((L.dev.ow + L2Shape<TH, NWC>::TW - 1) / L2Shape<TH, NWC>::TW, (L.dev.oh + TH - 1) / TH, Lab::PER);
  kern <<<grid, L2_THREADS, L2Shape<TH, NWC>::SMEM, s>>> (L.dev, L.st.dev, b);
}

struct Variant { const char *name; launch_fn fn; bool parity; };

int main (int argc, char **argv)
{
  const char *filter = argc > 1 ? argv[1] : "";
  const int reps = argc > 2 ? atoi (argv[2]) : 20;
  Lab L;
  b200_video_info ii, oi;
  b200_video_info_set_format (&ii, B200_VIDEO_FORMAT_NV12, 3840, 2160);
  b200_video_info_set_format (&oi, B200_VIDEO_FORMAT_BGRA, 1920, 1080);
  b200_vcs_config cfg;
  b200_vcs_config_init (&cfg);
  cfg.method = B200_SCALE_LANCZOS;
  if (build_vcs_plan (&ii, &oi, &cfg, &L.plan) != B200_OK) { printf ("plan failed\n"); return 1; }
  L.tab = build_lanczos2_tables (L.plan);
  if (!L.tab.ok || !L.tab.x4_ok) { printf ("tables not eligible\n"); return 1; }
  const VcsPlan & p = L.plan;
  VcsDev & d = L.dev;
  memset (&d, 0, sizeof (d));
  d.iw = p.in.width; d.ih = p.in.height; d.ow = p.out.width; d.oh = p.out.height;
  d.stride_y = p.in.stride[0]; d.stride_c = p.in.stride[1]; d.stride_out = p.out.stride[0];
  d.off_y = p.in.offset[0]; d.off_c = p.in.offset[1]; d.off_out = p.out.offset[0];
  d.u_index = p.u_index; d.h_cosited = p.h_cosited; d.v_pairs = p.v_pairs;
  d.p1 = p.p[0]; d.p2 = p.p[1]; d.p3 = p.p[2]; d.p4 = p.p[3]; d.p5 = p.p[4];
  d.sel = p.byte_sel[0] | (p.byte_sel[1] << 4) | (p.byte_sel[2] << 8) | (p.byte_sel[3] << 12);
  upload (&L.d_hsum, p.h.sum.data (), p.h.sum.size ());
  upload (&L.d_vsum, p.v.sum.data (), p.v.sum.size ());
  d.h.sum = L.d_hsum; d.v.sum = L.d_vsum;
  if (prepare_lanczos2 (L.tab, d, &L.st) != B200_OK) { printf ("prepare failed\n"); return 1; }
  L.in_bytes = b200_video_info_size (&ii); L.out_bytes = b200_video_info_size (&oi);
  for (int i = 0; i < Lab::RING; i++) {
    CK (cudaMalloc ((void **) &L.in[i], L.in_bytes));
    CK (cudaMalloc ((void **) &L.out[i], L.out_bytes));
    fill_random <<<592, 256>>> (L.in[i], L.in_bytes, 1234567u * (i + 1));
  }
  CK (cudaDeviceSynchronize ());
  L.tct = build_l2tc_tables (L.plan, L.tab);
  if (!L.tct.ok) { printf ("tensor tables not eligible\n"); return 1; }
  if (prepare_l2tc (L.tct, &L.tc) != B200_OK) { printf ("prepare_l2tc failed\n"); return 1; }
  if (!strcmp (filter, "tcdbg")) return tc_debug (L);
  L.v2t = build_lanczos2_v2_tables (L.plan, L.tab);
  if (!L.v2t.ok || prepare_lanczos2_v2 (L.v2t, &L.v2) != B200_OK) { printf ("v2 tables not eligible\n"); return 1; }

  std::vector<Variant> vs = {
    {"ref_x4_default", launch_l2<4, 60, 1, true, 0>, true},
    {"plain_tables", launch_l2<4, 60, 1, false, 0>, true},
    {"v2_bgra", launch_v2<0x0123, 0>, true},
    {"v2_bgra_prefetch_all", launch_v2<0x0123, 1>, true},
    {"v2_bgra_prefetch_chroma", launch_v2<0x0123, 2>, true},
    {"v2_runtime_sel", launch_v2<-1, 0>, true},
    {"tcgen05_both_passes", launch_tc, true},
    {"abl1_no_chroma_prep", launch_l2<4, 60, 1, true, 1>, false},
    {"abl2_no_hfir", launch_l2<4, 60, 1, true, 2>, false},
    {"abl3_no_chroma_no_hfir", launch_l2<4, 60, 1, true, 3>, false},
    {"abl4_h_phase_only", launch_l2<4, 60, 1, true, 4>, false},
    {"abl8_v_phase_only", launch_l2<4, 60, 1, true, 8>, false},
    {"abl16_no_matrix", launch_l2<4, 60, 1, true, 16>, false},
    {"abl32_no_vfir", launch_l2<4, 60, 1, true, 32>, false},
    {"abl48_v_loads_stores_only", launch_l2<4, 60, 1, true, 48>, false},
    {"abl7_h_loads_stores_only", launch_l2<4, 60, 1, true, 7>, false},
    {"abl63_skeleton", launch_l2<4, 60, 1, true, 51>, false},
#ifdef L2LAB_EXTRA_VARIANTS
    L2LAB_EXTRA_VARIANTS
#endif
  };
  cudaStream_t s;
  CK (cudaStreamCreate (&s));
  cudaEvent_t e0, e1;
  CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
  std::vector<uint8_t> got (L.out_bytes);
  for (const Variant & v : vs) {
    if (filter[0] && !strstr (v.name, filter) && strcmp (v.name, "ref_x4_default")) continue;
    VcsBatch b[2];
    for (int k = 0; k < 2; k++)
      for (int i = 0; i < Lab::PER; i++) { b[k].in[i] = L.in[k * Lab::PER + i]; b[k].out[i] = L.out[k * Lab::PER + i]; }
    for (int i = 0; i < Lab::RING; i++) CK (cudaMemsetAsync (L.out[i], 0x5a, L.out_bytes, s));
    for (int w = 0; w < 3; w++) v.fn (L, b[w & 1], s);
    CK (cudaStreamSynchronize (s));
    CK (cudaGetLastError ());
    CK (cudaEventRecord (e0, s));
    for (int r = 0; r < reps; r++) v.fn (L, b[r & 1], s);
    CK (cudaEventRecord (e1, s));
    CK (cudaStreamSynchronize (s));
    CK (cudaGetLastError ());
    float ms;
    CK (cudaEventElapsedTime (&ms, e0, e1));
    const double us_frame = ms * 1e3 / reps / Lab::PER;
    const double gbs = 20736000.0 / (us_frame * 1e-6) / 1e9;
    std::string verdict = "non-parity";
    if (v.parity) {
      verdict = "bit-exact vs ref";
      for (int f : {0, 31, 33}) {
        CK (cudaMemcpy (got.data (), L.out[f], L.out_bytes, cudaMemcpyDeviceToHost));
        if (!strcmp (v.name, "ref_x4_default")) { if (f == 0) L.ref = got; }
        else if (f == 0 && memcmp (got.data (), L.ref.data (), L.out_bytes)) {
          size_t k = 0; while (got[k] == L.ref[k]) k++;
          verdict = "MISMATCH at byte " + std::to_string (k);
        }
      }
    }
    printf ("%-34s %8.3f us/frame  %7.1f GB/s  frac %.3f  %s\n", v.name, us_frame, gbs, gbs / 6571.6, verdict.c_str ());
    fflush (stdout);
  }
  return 0;
}
