// oracle/refcuda/refcuda_bench.cu — TEST / MEASUREMENT INFRASTRUCTURE (never linked into the product).
//
// Drives the REFERENCE's own CUDA converter kernel on the bench shape, as the `gpu_baseline` of bench.py: the kernel
// source is compiled where it lies under /root/reference (included by path below, never copied into this repository):
//   gst-plugins-bad/gst-libs/gst/cuda/kernel/gstcudaconverter.cu:1358  GstCudaConverterMain, with
//   -DSAMPLER=SampleNV12 -DOUTPUT=OutputBGRA as gstcudaconverter.cpp:1549-1551 selects them for NV12 -> BGRA.
// Host side restated from the reference: textures per plane as gst_cuda_memory_get_texture builds them
// (gstcudamemory.cpp:880-907: pitch2D, normalized coordinates, linear filter, clamp, values read as normalized floats),
// the constant buffer of gst_cuda_converter_setup / _update_transform (gstcudaconverter.cpp:1129-1203), one launch per
// frame with 16 x 16 blocks over the output (gstcudaconverter.cpp:49-51, :2078-2080).
// NOT the same arithmetic as videoconvertscale: the texture unit interpolates bilinearly in 9-bit fixed point and the
// matrix runs in float - this is the reference's GPU path (cudaconvertscale), timed beside ours, not a parity oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define SAMPLER SampleNV12
#define OUTPUT OutputBGRA
#include REF_KERNEL_PATH

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString (e_), __FILE__, __LINE__); exit (1); } } while (0)

static cudaTextureObject_t make_tex (const uint8_t * ptr, int w, int h, size_t pitch, int channels)
{
  cudaResourceDesc rd; memset (&rd, 0, sizeof (rd));
  rd.resType = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr = (void *) ptr;
  rd.res.pitch2D.desc = channels == 1 ? cudaCreateChannelDesc<unsigned char> () : cudaCreateChannelDesc<uchar2> ();
  rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td; memset (&td, 0, sizeof (td));
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 1;
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  cudaTextureObject_t t = 0;
  CK (cudaCreateTextureObject (&t, &rd, &td, nullptr));
  return t;
}

__global__ void fill_random (uint8_t * p, size_t n, unsigned seed)
{
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n / 4; i += (size_t) gridDim.x * blockDim.x) {
    unsigned x = (unsigned) i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    ((unsigned *) p)[i] = x;
  }
}

int main (int argc, char **argv)
{
  const int iw = 3840, ih = 2160, ow = 1920, oh = 1080, RING = 64, PER = 32;
  const int reps = argc > 1 ? atoi (argv[1]) : 20;
  const size_t in_bytes = (size_t) iw * ih * 3 / 2, out_bytes = (size_t) ow * oh * 4;
  std::vector<uint8_t *> in (RING), out (RING);
  std::vector<cudaTextureObject_t> ty (RING), tc (RING);
  for (int i = 0; i < RING; i++) {
    CK (cudaMalloc ((void **) &in[i], in_bytes)); CK (cudaMalloc ((void **) &out[i], out_bytes));
    fill_random <<<592, 256>>> (in[i], in_bytes, 977u * (i + 1));
    ty[i] = make_tex (in[i], iw, ih, iw, 1);
    tc[i] = make_tex (in[i] + (size_t) iw * ih, iw / 2, ih / 2, iw, 2);
  }
  // constant buffer: BT.709 16-235 YCbCr -> full-range RGB in normalized floats, identity transform, no border, no blend
  ConstBuffer cb; memset (&cb, 0, sizeof (cb));
  const double Kr = 0.2126, Kb = 0.0722, Kg = 1.0 - Kr - Kb, sy = 255.0 / 219.0, sc = 255.0 / 224.0;
  const double m[3][3] = {{sy, 0, 2 * (1 - Kr) * sc}, {sy, -2 * Kb * (1 - Kb) / Kg * sc, -2 * Kr * (1 - Kr) / Kg * sc}, {sy, 2 * (1 - Kb) * sc, 0}};
  for (int i = 0; i < 3; i++) {
    cb.matrix.CoeffX[i] = (float) m[0][i]; cb.matrix.CoeffY[i] = (float) m[1][i]; cb.matrix.CoeffZ[i] = (float) m[2][i];
    cb.matrix.Min[i] = 0.f; cb.matrix.Max[i] = 1.f;
  }
  for (int r = 0; r < 3; r++) {
    const double off = -(m[r][0] * 16.0 / 255.0 + m[r][1] * 128.0 / 255.0 + m[r][2] * 128.0 / 255.0);
    cb.matrix.Offset[r] = (float) off;
  }
  cb.width = ow; cb.height = oh; cb.left = 0; cb.top = 0; cb.right = ow; cb.bottom = oh; cb.view_width = ow; cb.view_height = oh;
  cb.border_w = 1.f; cb.fill_border = 0; cb.alpha = 1.f; cb.do_blend = 0; cb.do_convert = 1;
  cb.transform_u[0] = 1.f; cb.transform_v[1] = 1.f;
  CK (cudaDeviceSynchronize ());
  cudaStream_t s; CK (cudaStreamCreate (&s));
  cudaEvent_t e0, e1; CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
  const dim3 block (16, 16), grid ((ow + 15) / 16, (oh + 15) / 16);
  auto step = [&] (int k) {
    for (int f = 0; f < PER; f++) {
      const int i = (k & 1) * PER + f;
      GstCudaConverterMain <<<grid, block, 0, s>>> (ty[i], tc[i], 0, 0, out[i], nullptr, nullptr, nullptr, ow * 4, 0, cb, 0, 0);
    }
  };
  for (int w = 0; w < 3; w++) step (w);
  CK (cudaStreamSynchronize (s)); CK (cudaGetLastError ());
  CK (cudaEventRecord (e0, s));
  for (int r = 0; r < reps; r++) step (r);
  CK (cudaEventRecord (e1, s));
  CK (cudaStreamSynchronize (s)); CK (cudaGetLastError ());
  float ms; CK (cudaEventElapsedTime (&ms, e0, e1));
  const double us = ms * 1e3 / reps / PER;
  std::vector<uint8_t> px (16);
  CK (cudaMemcpy (px.data (), out[0], 16, cudaMemcpyDeviceToHost));
  printf ("{\"kernel\": \"GstCudaConverterMain<SampleNV12, OutputBGRA> (reference gstcudaconverter.cu:1358, texture-bilinear, float matrix)\", "
      "\"us_per_frame\": %.3f, \"mpix_per_s_in\": %.1f, \"achieved_gbs\": %.1f, \"launches_timed\": %d, \"first_pixel_bgra\": [%d, %d, %d, %d]}\n",
      us, (double) iw * ih / us, 20736000.0 / us / 1e3, reps * PER, px[0], px[1], px[2], px[3]);
  return 0;
}
