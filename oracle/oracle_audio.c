/* oracle/oracle_audio.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Restatement of the audioresample hot path for F32 interleaved audio with the element's
 * defaults (kaiser window, filter-mode auto, cubic filter interpolation):
 *   option plumbing        gst/audioresample/gstaudioresample.c:374-396 (make_options),
 *                          gst-libs/gst/audio/audio-resampler.c:1277-1328 (options_set_quality)
 *   rate reduction         audio-resampler.c:1502-1560 (gst_audio_resampler_update)
 *   filter design          :927-965 (calculate_kaiser_params), :1062-1208 (resampler_calculate_taps),
 *                          :205-215 (get_kaiser_tap), :287-323 (make_taps), :259-268 (float taps)
 *   phase cache            :503-561 (get_taps_gfloat_full), :360-373 (make_coeff_gfloat_cubic),
 *                          audio-resampler-x86-sse.c:139-167 (interpolate_gfloat_cubic_sse)
 *   inner products         audio-resampler-x86-sse.c:27-46 (full), :82-120 (cubic, interpolated mode)
 *   framing                :1648-1678 (get_out_frames), :1750-1805 (resample), macros.h:62-100
 * The SSE lane order is reproduced in scalar C: the real x86 build of the reference takes the SSE
 * functions (audio-resampler-x86.h:29-70), and that is what oracle/_ref is compiled as.
 * Build with -ffp-contract=off (oracle/Makefile): no FMA may be formed.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "besi0_coeffs.inc"

#define ROUND_UP_8(n) (((n) + 7) & ~7)

/* I0(x): piecewise polynomials, evaluated exactly like dbesi0.c:113-145 */
static double
bessel_i0 (double x)
{
  double w = fabs (x), t, y;
  int k, i;
  if (w < 8.5) {
    t = w * w * 0.0625;
    k = 13 * ((int) t);
    y = ORACLE_I0_LOW[k];
    for (i = 1; i < 13; i++)
      y = y * t + ORACLE_I0_LOW[k + i];
  } else if (w < 12.5) {
    k = (int) w;
    t = w - k;
    k = 14 * (k - 8);
    y = ORACLE_I0_MID[k];
    for (i = 1; i < 14; i++)
      y = y * t + ORACLE_I0_MID[k + i];
  } else {
    t = 60 / w;
    k = 9 * ((int) t);
    y = ORACLE_I0_HIGH[k];
    for (i = 1; i < 9; i++)
      y = y * t + ORACLE_I0_HIGH[k + i];
    y = y * sqrt (t) * exp (w);
  }
  return y;
}

struct OracleArs
{
  int channels, in_rate, out_rate;      /* rates after gcd reduction */
  int samp_inc, samp_frac, samp_index, samp_phase, skip;
  int n_taps, oversample, n_phases, full;
  double cutoff, beta;
  float *table;                 /* (oversample + 4) rows of n_taps, the oversampled prototype */
  float *cache;                 /* n_phases rows of n_taps (FULL mode) */
  unsigned char *have;
  float **sbuf;                 /* per channel history + input */
  size_t samples_len, samples_avail;
};

static const struct { double cutoff, down, atten, trbw; } kaiser_q[11] = {
  {0.860, 0.96511, 60, 0.7}, {0.880, 0.96591, 65, 0.29}, {0.910, 0.96923, 70, 0.145},
  {0.920, 0.97600, 80, 0.105}, {0.940, 0.97979, 85, 0.087}, {0.940, 0.98085, 95, 0.077},
  {0.945, 0.99471, 100, 0.068}, {0.950, 1.0, 105, 0.055}, {0.960, 1.0, 110, 0.045},
  {0.968, 1.0, 115, 0.039}, {0.975, 1.0, 120, 0.0305}
};
static const int oversample_q[11] = { 4, 4, 4, 8, 8, 16, 16, 16, 16, 32, 32 };

static int
gcd_ (int a, int b)
{
  while (b) { int t = a; a = b; b = t % b; }
  return a < 0 ? -a : a;
}

/* one prototype row: n_taps kaiser-windowed sinc values at offset x, normalised, as float */
static void
make_row (const OracleArs * r, float *res, double x)
{
  int i, n = r->n_taps;
  double weight = 0.0, *tmp = malloc (sizeof (double) * n);
  for (i = 0; i < n; i++) {
    double xx = x + i, y = M_PI * xx, s, w;
    s = (y == 0.0 ? r->cutoff : sin (y * r->cutoff) / y);
    w = 2.0 * xx / n;
    tmp[i] = s * bessel_i0 (r->beta * sqrt (fmax (1 - w * w, 0)));
    weight += tmp[i];
  }
  for (i = 0; i < n; i++)
    res[i] = (float) (tmp[i] / weight);
  free (tmp);
}

OracleArs *
oracle_ars_new (int in_rate, int out_rate, int channels, int quality)
{
  OracleArs *r;
  double Fc, A, tr_bw, B, dw;
  int g, n, oversample, i;
  if (in_rate <= 0 || out_rate <= 0 || channels <= 0 || quality < 0 || quality > 10)
    return NULL;
  r = calloc (1, sizeof (*r));
  r->channels = channels;
  /* update(): first call runs with options == NULL: max_error 0.1, samp_phase 0 -> plain gcd */
  g = gcd_ (in_rate, out_rate);
  r->in_rate = in_rate / g;
  r->out_rate = out_rate / g;
  r->samp_inc = r->in_rate / r->out_rate;
  r->samp_frac = r->in_rate % r->out_rate;

  /* options_set_quality(kaiser) then calculate_kaiser_params: the option values override the
   * DEFAULT_QUALITY row */
  Fc = kaiser_q[quality].cutoff;
  if (out_rate < in_rate)
    Fc *= kaiser_q[quality].down;
  A = kaiser_q[quality].atten;
  tr_bw = kaiser_q[quality].trbw;
  if (A > 50)
    B = 0.1102 * (A - 8.7);
  else if (A >= 21)
    B = 0.5842 * pow (A - 21, 0.4) + 0.07886 * (A - 21);
  else
    B = 0.0;
  dw = 2 * M_PI * tr_bw;
  n = (int) ((A - 8.0) / (2.285 * dw));
  r->beta = B;
  r->n_taps = n + 1;
  r->cutoff = Fc;
  if (r->out_rate < r->in_rate) {
    r->cutoff = r->cutoff * r->out_rate / r->in_rate;
    r->n_taps = (int) (((unsigned long long) r->n_taps * r->in_rate) / r->out_rate);
  }
  r->n_taps = ROUND_UP_8 (r->n_taps);
  /* cubic filter interpolation: oversampling from the quality, halved while the decimation
   * ratio allows it (audio-resampler.c:1119-1140) */
  {
    int mult = 2;
    oversample = oversample_q[quality];
    while (oversample > 1) {
      if (mult * r->out_rate >= r->in_rate)
        break;
      mult *= 2;
      oversample >>= 1;
    }
  }
  r->oversample = oversample;
  /* filter-mode auto; the element stores the threshold as UINT, the resampler reads it as INT,
   * so the default 1048576 always applies; VARIABLE_RATE is set by the
   * element so the first clause never selects FULL */
  if (4 * r->n_taps * r->out_rate < 1048576)
    r->full = 1;
  r->n_phases = r->out_rate;
  r->table = calloc ((size_t) (oversample + 4) * r->n_taps, sizeof (float));
  for (i = 0; i < oversample + 4; i++)
    make_row (r, r->table + (size_t) i * r->n_taps, -(r->n_taps / 2) + i / (double) oversample);
  if (r->full) {
    r->cache = calloc ((size_t) r->n_phases * r->n_taps, sizeof (float));
    r->have = calloc (r->n_phases, 1);
  }
  r->sbuf = calloc (channels, sizeof (float *));
  oracle_ars_reset (r);
  return r;
}

void
oracle_ars_free (OracleArs * r)
{
  int c;
  if (!r)
    return;
  for (c = 0; c < r->channels; c++)
    free (r->sbuf[c]);
  free (r->sbuf);
  free (r->table);
  free (r->cache);
  free (r->have);
  free (r);
}

void
oracle_ars_reset (OracleArs * r)
{
  int c;
  for (c = 0; c < r->channels; c++)
    if (r->sbuf[c])
      memset (r->sbuf[c], 0, sizeof (float) * (r->n_taps / 2));
  r->samp_index = 0;
  r->samples_avail = r->n_taps / 2 - 1;
  /* note: samp_phase and skip are left alone, exactly like gst_audio_resampler_reset() */
}

static void
cubic_coeff (int num, int denom, float ic[4])
{
  /* make_coeff_gfloat_cubic, audio-resampler.c:360-373 */
  float x = (float) num / denom, x2 = x * x, x3 = x2 * x;
  ic[0] = 0.16667f * (x3 - x);
  ic[1] = x + 0.5f * (x2 - x3);
  ic[3] = -0.33333f * x + 0.5f * x2 - 0.16667f * x3;
  ic[2] = (float) 1.0 - ic[0] - ic[1] - ic[3];
}

/* taps of one phase in FULL mode, built lazily like get_taps_gfloat_full() */
static const float *
phase_taps (OracleArs * r, int phase)
{
  float *res = r->cache + (size_t) phase * r->n_taps;
  if (!r->have[phase]) {
    int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->n_phases;
    int frac = pos % r->n_phases, i, n = r->n_taps;
    const float *c0 = r->table + (size_t) offset * n, *c1 = c0 + n, *c2 = c1 + n, *c3 = c2 + n;
    float ic[4];
    cubic_coeff (frac, r->n_phases, ic);
    for (i = 0; i < n; i++) {
      /* interpolate_gfloat_cubic_sse: (c0*f0 + c1*f1) + (c2*f2 + c3*f3) */
      float t0 = c0[i] * ic[0], t1 = c1[i] * ic[1], t2 = c2[i] * ic[2], t3 = c3[i] * ic[3];
      t0 = t0 + t1;
      t2 = t2 + t3;
      res[i] = t0 + t2;
    }
    r->have[phase] = 1;
  }
  return res;
}

int
oracle_ars_phase_taps (OracleArs * r, int phase, float *taps)
{
  if (!r->full || phase < 0 || phase >= r->n_phases)
    return -1;
  memcpy (taps, phase_taps (r, phase), sizeof (float) * r->n_taps);
  return r->n_taps;
}

int
oracle_ars_info (OracleArs * r, int *n_taps, int *n_phases, int *in_step, int *out_step,
    int *filter_mode, int *oversample)
{
  *n_taps = r->n_taps;
  *n_phases = r->full ? r->n_phases : 0;  /* only the FULL mode sets n_phases (audio-resampler.c:1173-1178) */
  *in_step = r->in_rate;
  *out_step = r->out_rate;
  *filter_mode = r->full ? 1 : 0;       /* GST_AUDIO_RESAMPLER_FILTER_MODE_{INTERPOLATED=0,FULL=1} */
  *oversample = r->oversample;
  return 0;
}

size_t
oracle_ars_get_out_frames (OracleArs * r, size_t in_frames)
{
  size_t need = r->n_taps + r->samp_index + r->skip, avail = r->samples_avail + in_frames, out;
  if (avail < need)
    return 0;
  out = (avail - need) * r->out_rate;
  if (out < (size_t) r->samp_phase)
    return 0;
  return ((out - r->samp_phase) / r->in_rate) + 1;
}

size_t
oracle_ars_get_in_frames (OracleArs * r, size_t out_frames)
{
  size_t in_frames = (r->samp_phase + out_frames * r->samp_frac) / r->out_rate;
  return in_frames + out_frames * r->samp_inc;
}

size_t
oracle_ars_max_latency (OracleArs * r)
{
  return r->n_taps / 2;
}

/* inner_product_gfloat_full_1_sse: four lanes, mul then add, (l0+l2)+(l1+l3) */
static float
dot_full (const float *a, const float *b, int len)
{
  float s[4] = { 0, 0, 0, 0 };
  int i, l;
  for (i = 0; i < len; i += 4)
    for (l = 0; l < 4; l++) {
      float p = a[i + l] * b[i + l];
      s[l] = s[l] + p;
    }
  return (s[0] + s[2]) + (s[1] + s[3]);
}

/* inner_product_gfloat_cubic_1_sse (interpolated filter mode) */
static float
dot_cubic (const float *a, const float *c0, int stride, int len, const float ic[4])
{
  float s[4][4];
  int i, l, k;
  memset (s, 0, sizeof (s));
  for (i = 0; i < len; i += 4)
    for (k = 0; k < 4; k++)
      for (l = 0; l < 4; l++) {
        float p = a[i + l] * c0[(size_t) k * stride + i + l];
        s[k][l] = s[k][l] + p;
      }
  for (l = 0; l < 4; l++) {
    float t0 = s[0][l] * ic[0], t1 = s[1][l] * ic[1], t2 = s[2][l] * ic[2], t3 = s[3][l] * ic[3];
    t0 = t0 + t1;
    t2 = t2 + t3;
    s[0][l] = t0 + t2;
  }
  return (s[0][0] + s[0][2]) + (s[0][1] + s[0][3]);
}

size_t
oracle_ars_process (OracleArs * r, const float *in, size_t in_frames, float *out,
    size_t out_capacity)
{
  size_t out_frames = oracle_ars_get_out_frames (r, in_frames), avail, need, di, consumed;
  int c, ch = r->channels, samp_index = 0, samp_phase = 0;
  if (out_frames > out_capacity)
    out_frames = out_capacity;
  /* gst_audio_resampler_resample(), audio-resampler.c:1750-1805 */
  if ((size_t) r->skip >= in_frames) {
    r->skip -= (int) in_frames;
    return out_frames;
  }
  r->samp_index += r->skip;
  avail = r->samples_avail;
  if (r->samples_len < in_frames + avail) {
    for (c = 0; c < ch; c++) {
      float *n = calloc (in_frames + avail, sizeof (float));
      if (r->sbuf[c])
        memcpy (n, r->sbuf[c], avail * sizeof (float));
      free (r->sbuf[c]);
      r->sbuf[c] = n;
    }
    r->samples_len = in_frames + avail;
  }
  for (c = 0; c < ch; c++) {    /* deinterleave_gfloat, :879-897 */
    float *s = r->sbuf[c] + avail;
    size_t i;
    for (i = 0; i < in_frames; i++)
      s[i] = in ? in[i * ch + c] : 0.0f;
  }
  r->samples_avail = avail += in_frames;
  need = r->n_taps + r->samp_index;
  if (avail < need || out_frames == 0)
    return out_frames;
  for (c = 0; c < ch; c++) {    /* MAKE_RESAMPLE_FUNC, audio-resampler-macros.h:62-100 */
    float *ip = r->sbuf[c];
    samp_index = r->samp_index;
    samp_phase = r->samp_phase;
    for (di = 0; di < out_frames; di++) {
      const float *ipp = ip + samp_index;
      if (r->full) {
        out[di * ch + c] = dot_full (ipp, phase_taps (r, samp_phase), r->n_taps);
      } else {                  /* get_taps_gfloat_cubic, :567-600 */
        int pos = samp_phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
        float ic[4];
        cubic_coeff (pos % r->out_rate, r->out_rate, ic);
        out[di * ch + c] = dot_cubic (ipp, r->table + (size_t) offset * r->n_taps, r->n_taps,
            r->n_taps, ic);
      }
      samp_index += r->samp_inc;
      samp_phase += r->samp_frac;
      if (samp_phase >= r->out_rate) {
        samp_phase -= r->out_rate;
        samp_index += 1;
      }
    }
    if (avail > (size_t) samp_index)
      memmove (ip, ip + samp_index, (avail - samp_index) * sizeof (float));
  }
  consumed = samp_index - r->samp_index;
  r->samp_index = 0;
  r->samp_phase = samp_phase;
  if (consumed > 0) {
    if (avail > consumed) {
      r->samples_avail = avail - consumed;
    } else {
      r->samples_avail = 0;
      r->skip = (int) (consumed - avail);
    }
  }
  return out_frames;
}
