// tools/ubench.cu — instruction-throughput microbenchmarks on sm_100a (development aid, not product).
// Each kernel runs ITER x 16 independent instances of one operation per thread with all SMs full;
// reports warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define DEF(name, BODY)                                                            \
  __global__ void k_##name (unsigned *out, unsigned seed)                          \
  {                                                                                \
    unsigned a[16];                                                                \
    _Pragma ("unroll") for (int i = 0; i < 16; i++) a[i] = seed + threadIdx.x * 16 + i; \
    unsigned b = seed * 3 + threadIdx.x, c = seed ^ 0x55aa;                        \
    for (int it = 0; it < ITER; it++) {                                            \
      _Pragma ("unroll") for (int i = 0; i < 16; i++) { BODY; }                    \
    }                                                                              \
    unsigned s = 0;                                                                \
    _Pragma ("unroll") for (int i = 0; i < 16; i++) s += a[i];                     \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b + c;                        \
  }

DEF (lop3, a[i] = (a[i] & b) ^ c)
DEF (shf, asm volatile ("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b)))
DEF (shr, a[i] = ((int) a[i] >> 6) + 1)
DEF (prmt, asm volatile ("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a[i]) : "r"(b)))
DEF (iadd3, a[i] = a[i] + b + c)
DEF (i2ip, asm volatile ("cvt.pack.sat.u8.s32.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c)))
DEF (vimnmx, a[i] = min ((int) a[i], (int) b) + 0)
DEF (idp4a, asm volatile ("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)))
DEF (imad, a[i] = a[i] * b + c)
DEF (imadhi, a[i] = __mulhi ((int) a[i], (int) b))
DEF (shfl, a[i] = __shfl_down_sync (0xffffffffu, a[i], 1))
DEF (fadd, a[i] = __float_as_uint (__fadd_rn (__uint_as_float (a[i]), __uint_as_float (b))))
DEF (fmul, a[i] = __float_as_uint (__fmul_rn (__uint_as_float (a[i]), __uint_as_float (b))))
DEF (mix_alu_fma, a[i] = (i & 1) ? ((a[i] & b) ^ c) : (a[i] * b + c))
DEF (mix_lop_dp, if (i & 1) a[i] = (a[i] & b) ^ c; else asm volatile ("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c)))
DEF (mix_fmul_fadd, a[i] = (i & 1) ? __float_as_uint (__fadd_rn (__uint_as_float (a[i]), __uint_as_float (b))) : __float_as_uint (__fmul_rn (__uint_as_float (a[i]), __uint_as_float (b))))

template <typename K>
void run (const char *name, K kern, unsigned *out, int sms)
{
  cudaEvent_t e0, e1;
  cudaEventCreate (&e0); cudaEventCreate (&e1);
  const int blocks = sms * 8, threads = 256;
  kern <<<blocks, threads>>> (out, 1);
  cudaDeviceSynchronize ();
  cudaEventRecord (e0);
  kern <<<blocks, threads>>> (out, 2);
  cudaEventRecord (e1);
  cudaDeviceSynchronize ();
  float ms; cudaEventElapsedTime (&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute (&clk_khz, cudaDevAttrClockRate, 0);
  const double winst = (double) blocks * (threads / 32) * ITER * 16;
  const double clocks = ms * 1e-3 * clk_khz * 1e3;
  printf ("%-14s %8.3f ms  %6.2f warp-inst/clk/SM (at nominal %d MHz)\n", name, ms, winst / clocks / sms, clk_khz / 1000);
}

int main ()
{
  int sms; cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned *out; cudaMalloc (&out, sms * 8 * 256 * 4);
#define R(n) run (#n, k_##n, out, sms)
  R (lop3); R (shf); R (shr); R (prmt); R (iadd3); R (i2ip); R (vimnmx); R (idp4a); R (imad); R (imadhi);
  R (shfl); R (fadd); R (fmul); R (mix_alu_fma); R (mix_lop_dp); R (mix_fmul_fadd);
  return 0;
}
