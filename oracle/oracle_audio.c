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
  int copy;                     /* nearest method, or equal rates (setup_functions, audio-resampler.c:1019-1020): *o = *a */
  int linear, isize;            /* sinc-filter-interpolation=linear: two table rows per phase, 11x the oversampling */
  int method, interp_none;      /* ORACLE_ARS_METHOD_*; sinc-filter-interpolation=none (FULL mode: exact taps per phase) */
  int quality, opt_filter_mode, opt_interpolation;      /* the option bag, kept for oracle_ars_update () */
  int raw_down;                 /* out_rate < in_rate on the UN-reduced rates (options_set_quality sees those) */
  double cutoff, beta;
  int fmt, bps;                 /* ORACLE_AFMT_*, bytes per sample */
  float *table;                 /* (oversample + 4) rows of n_taps, the oversampled prototype (typed by fmt) */
  float *cache;                 /* n_phases rows of n_taps (FULL mode) (typed by fmt) */
  unsigned char *have;
  float **sbuf;                 /* per channel history + input (typed by fmt) */
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

/* convert_taps_gint16_c / _gint32_c (audio-resampler.c:217-258): round with an adjustable bias found by
 * bisection so that the integer taps sum to (1 << precision) - 1 */
static void
convert_taps_int (const double *tmp, void *taps, double weight, int n, int precision, int is16)
{
  long long one = (1LL << precision) - 1;
  double multiplier = (double) one, offset = 0.5, l_offset = 0.0, h_offset = 1.0;
  int i, j;
  for (i = 0; i < 32; i++) {
    long long sum = 0;
    for (j = 0; j < n; j++)
      sum += (long long) floor (offset + tmp[j] * multiplier / weight);
    if (sum == one)
      break;
    if (l_offset == h_offset)
      break;
    if (sum < one) {
      if (offset > l_offset)
        l_offset = offset;
      offset += (h_offset - l_offset) / 2;
    } else {
      if (offset < h_offset)
        h_offset = offset;
      offset -= (h_offset - l_offset) / 2;
    }
  }
  for (j = 0; j < n; j++) {
    double v = floor (offset + tmp[j] * multiplier / weight);
    if (is16)
      ((int16_t *) taps)[j] = (int16_t) v;
    else
      ((int32_t *) taps)[j] = (int32_t) v;
  }
}

/* one prototype row: n_taps kaiser-windowed sinc values at offset x, normalised, in the sample format
 * (make_taps + convert_taps_<type>, audio-resampler.c:287-323, :217-268) */
static void
make_row (const OracleArs * r, void *res, double x)
{
  int i, n = r->n_taps;
  double weight = 0.0, *tmp = malloc (sizeof (double) * n);
  for (i = 0; i < n; i++) {
    double xx = x + i, y = M_PI * xx, s, w;
    s = (y == 0.0 ? r->cutoff : sin (y * r->cutoff) / y);
    if (r->method == ORACLE_ARS_METHOD_LINEAR) {        /* get_linear_tap, audio-resampler.c:163-168 */
      tmp[i] = ((n + 1) & ~1) / 2 - fabs (xx);
    } else if (r->method == ORACLE_ARS_METHOD_CUBIC) {  /* get_cubic_tap :170-190 with the default b = 1, c = 0 (:97-98) */
      const double b = 1.0, c = 0.0;
      double a = fabs (xx * 4.0) / n, a2 = a * a, a3 = a2 * a;
      if (a <= 1.0)
        tmp[i] = ((12.0 - 9.0 * b - 6.0 * c) * a3 + (-18.0 + 12.0 * b + 6.0 * c) * a2 + (6.0 - 2.0 * b)) / 6.0;
      else if (a <= 2.0)
        tmp[i] = ((-b - 6.0 * c) * a3 + (6.0 * b + 30.0 * c) * a2 + (-12.0 * b - 48.0 * c) * a + (8.0 * b + 24.0 * c)) / 6.0;
      else
        tmp[i] = 0.0;
    } else if (r->method == ORACLE_ARS_METHOD_BLACKMAN_NUTTALL) {      /* get_blackman_nuttall_tap, audio-resampler.c:192-203 */
      w = 2.0 * y / n + M_PI;
      tmp[i] = s * (0.3635819 - 0.4891775 * cos (w) + 0.1365995 * cos (2 * w) - 0.0106411 * cos (3 * w));
    } else {                    /* get_kaiser_tap, :205-215 */
      w = 2.0 * xx / n;
      tmp[i] = s * bessel_i0 (r->beta * sqrt (fmax (1 - w * w, 0)));
    }
    weight += tmp[i];
  }
  switch (r->fmt) {
    case ORACLE_AFMT_S16: convert_taps_int (tmp, res, weight, n, 15, 1); break;
    case ORACLE_AFMT_S32: convert_taps_int (tmp, res, weight, n, 31, 0); break;
    case ORACLE_AFMT_F64:
      for (i = 0; i < n; i++)
        ((double *) res)[i] = tmp[i] / weight;
      break;
    default:
      for (i = 0; i < n; i++)
        ((float *) res)[i] = (float) (tmp[i] / weight);
  }
  free (tmp);
}

/* resampler_calculate_taps () with the element's option bag (audio-resampler.c:1064-1215), on the handle's current
 * (reduced) rates; called from the constructor and from oracle_ars_update () */
static void
ars_configure (OracleArs * r)
{
  static const struct { int n_taps; double cutoff; } blackman_q[11] = { {8, 0.5}, {16, 0.6}, {24, 0.72}, {32, 0.8},
    {48, 0.85}, {64, 0.90}, {80, 0.92}, {96, 0.933}, {128, 0.950}, {148, 0.955}, {160, 0.960} };
  double Fc, A, tr_bw, B, dw;
  int n, oversample, i;
  const int quality = r->quality, method = r->method;
  int filter_mode = r->opt_filter_mode, interpolation = r->opt_interpolation;
  free (r->table); free (r->cache); free (r->have);
  r->table = r->cache = NULL; r->have = NULL;
  r->copy = method == ORACLE_ARS_METHOD_NEAREST || r->in_rate == r->out_rate;
  /* options_set_quality(kaiser) then calculate_kaiser_params: the option values override the
   * DEFAULT_QUALITY row */
  Fc = kaiser_q[quality].cutoff;
  if (r->raw_down)
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
  if (method == ORACLE_ARS_METHOD_BLACKMAN_NUTTALL) {   /* options_set_quality :1299-1305 -> calculate_taps :1084-1090 */
    r->n_taps = blackman_q[quality].n_taps;
    r->cutoff = blackman_q[quality].cutoff;
  }
  /* the methods without a sinc table (resampler_calculate_taps :1070-1117): 2 taps (nearest: never scaled), 2 (linear),
   * 4 (cubic; options_set_quality :1286-1297); no rounding up to 8 - the SIMD inner products run into the 16 zero taps
   * every table row ends with (TAPS_OVERREAD :34, :979), which changes no sum; FULL mode, no table interpolation */
  if (method <= ORACLE_ARS_METHOD_CUBIC) {
    r->n_taps = method == ORACLE_ARS_METHOD_CUBIC ? 4 : 2;
    filter_mode = ORACLE_ARS_MODE_FULL;
    interpolation = ORACLE_ARS_INTERP_NONE;
  }
  if (r->out_rate < r->in_rate && method != ORACLE_ARS_METHOD_NEAREST) {
    r->cutoff = r->cutoff * r->out_rate / r->in_rate;
    r->n_taps = (int) (((unsigned long long) r->n_taps * r->in_rate) / r->out_rate);
  }
  if (method > ORACLE_ARS_METHOD_CUBIC)
    r->n_taps = ROUND_UP_8 (r->n_taps);
  /* cubic filter interpolation: oversampling from the quality, halved while the decimation
   * ratio allows it (audio-resampler.c:1119-1140) */
  if (interpolation != ORACLE_ARS_INTERP_NONE) {
    int mult = 2;
    oversample = oversample_q[quality];
    while (oversample > 1) {
      if (mult * r->out_rate >= r->in_rate)
        break;
      mult *= 2;
      oversample >>= 1;
    }
    if (interpolation == ORACLE_ARS_INTERP_LINEAR)
      oversample *= 11;         /* :1131-1137 */
  } else
    oversample = 1;             /* :1141-1143 */
  r->oversample = oversample;
  r->linear = interpolation == ORACLE_ARS_INTERP_LINEAR;
  r->isize = r->linear ? 2 : 4; /* :1186-1197 */
  /* filter-mode auto; the element stores the threshold as UINT, the resampler reads it as INT,
   * so the default 1048576 always applies; VARIABLE_RATE is set by the
   * element so the first clause never selects FULL */
  if (filter_mode == ORACLE_ARS_MODE_AUTO)
    r->full = r->bps * r->n_taps * r->out_rate < 1048576;       /* bps * n_taps * out_rate, :1153 */
  else
    r->full = filter_mode == ORACLE_ARS_MODE_FULL;
  /* an interpolated table without an interpolation falls back to the default cubic one - with the oversampling of 1
   * computed above (:1167-1170) */
  r->interp_none = r->full && interpolation == ORACLE_ARS_INTERP_NONE;
  r->n_phases = r->out_rate;
  r->table = calloc ((size_t) (oversample + r->isize) * r->n_taps, r->bps);
  for (i = 0; i < oversample + r->isize && method > ORACLE_ARS_METHOD_CUBIC; i++)     /* no table without a table interpolation (:1181-1201) */
    make_row (r, (char *) r->table + (size_t) i * r->n_taps * r->bps, -(r->n_taps / 2) + i / (double) oversample);
  if (r->full) {
    r->cache = calloc ((size_t) r->n_phases * r->n_taps, r->bps);
    r->have = calloc (r->n_phases, 1);
  }
}

OracleArs *
oracle_ars_new (int in_rate, int out_rate, int channels, int quality)
{
  return oracle_ars_new_fmt (in_rate, out_rate, channels, quality, ORACLE_AFMT_F32);
}

OracleArs *
oracle_ars_new_fmt (int in_rate, int out_rate, int channels, int quality, int fmt)
{
  return oracle_ars_new_opts (in_rate, out_rate, channels, quality, fmt, ORACLE_ARS_METHOD_KAISER, ORACLE_ARS_MODE_AUTO,
      ORACLE_ARS_INTERP_CUBIC);
}

OracleArs *
oracle_ars_new_opts (int in_rate, int out_rate, int channels, int quality, int fmt, int method, int filter_mode,
    int interpolation)
{
  OracleArs *r;
  int g;
  if (in_rate <= 0 || out_rate <= 0 || channels <= 0 || quality < 0 || quality > 10)
    return NULL;
  if (fmt < 0 || fmt > ORACLE_AFMT_F64)
    return NULL;
  if (method < ORACLE_ARS_METHOD_NEAREST || method > ORACLE_ARS_METHOD_KAISER ||
      filter_mode < ORACLE_ARS_MODE_INTERPOLATED || filter_mode > ORACLE_ARS_MODE_AUTO ||
      interpolation < ORACLE_ARS_INTERP_NONE || interpolation > ORACLE_ARS_INTERP_CUBIC)
    return NULL;
  r = calloc (1, sizeof (*r));
  r->method = method;
  r->channels = channels;
  r->fmt = fmt;
  r->bps = fmt == ORACLE_AFMT_S16 ? 2 : (fmt == ORACLE_AFMT_F64 ? 8 : 4);
  /* update(): first call runs with options == NULL: max_error 0.1, samp_phase 0 -> plain gcd */
  g = gcd_ (in_rate, out_rate);
  r->in_rate = in_rate / g;
  r->out_rate = out_rate / g;
  r->samp_inc = r->in_rate / r->out_rate;
  r->samp_frac = r->in_rate % r->out_rate;
  r->quality = quality; r->opt_filter_mode = filter_mode; r->opt_interpolation = interpolation;
  r->raw_down = out_rate < in_rate;
  ars_configure (r);
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
      memset (r->sbuf[c], 0, (size_t) r->bps * (r->n_taps / 2));
  r->samp_index = 0;
  r->samples_avail = r->n_taps / 2 - 1;
  /* note: samp_phase and skip are left alone, exactly like gst_audio_resampler_reset() */
}

/* gst_audio_resampler_update () as the element drives it (gst_audio_resample_update_state -> gst_audio_converter_update_config,
 * gstaudioresample.c:398-437, audio-converter.c:349-375): new rates AND a fresh option bag, so the filter is re-designed;
 * the stream position survives: the phase is rescaled, the common divisor of the rates is reduced only as far as the phase
 * error stays below 0.1 (DEFAULT_OPT_MAX_PHASE_ERROR), and the history moves by half the change of the tap count
 * (audio-resampler.c:1503-1615).  in_rate / out_rate <= 0 keep the old value - of the REDUCED rates, like the reference. */
int
oracle_ars_update (OracleArs * r, int in_rate, int out_rate)
{
  int g, samp_phase, old_n_taps = r->n_taps, c;
  if (in_rate <= 0)
    in_rate = r->in_rate;
  if (out_rate <= 0)
    out_rate = r->out_rate;
  samp_phase = (int) (((unsigned long long) r->samp_phase * (unsigned long long) out_rate) / (unsigned long long) r->out_rate);
  g = gcd_ (in_rate, out_rate);
  while (g > 1) {
    double ph1 = (double) samp_phase / out_rate;
    double ph2 = (double) (samp_phase / g) / (out_rate / g);
    int factor = 2;
    if (fabs (ph1 - ph2) < 0.1)
      break;
    while (g % factor != 0)
      factor++;
    g /= factor;
  }
  r->raw_down = out_rate < in_rate;
  r->samp_phase = samp_phase / g;
  r->in_rate = in_rate / g;
  r->out_rate = out_rate / g;
  r->samp_inc = r->in_rate / r->out_rate;
  r->samp_frac = r->in_rate % r->out_rate;
  ars_configure (r);
  if (old_n_taps > 0 && old_n_taps != r->n_taps) {
    const int diff = (r->n_taps - old_n_taps) / 2, bps = r->bps;
    size_t need = (size_t) r->n_taps;                              /* get_sample_bufs (resampler, n_taps) */
    long long bytes = (long long) r->samples_avail * bps, soff = (long long) r->samp_index * bps, doff = soff;
    if (need < r->samp_index + r->samples_avail + (size_t) (diff > 0 ? diff : 0))
      need = r->samp_index + r->samples_avail + (size_t) (diff > 0 ? diff : 0);   /* the reference would write past its buffer */
    if (r->samples_len < need) {
      for (c = 0; c < r->channels; c++) {
        char *n = calloc (need, bps);
        if (r->sbuf[c])
          memcpy (n, r->sbuf[c], r->samples_avail * bps);
        free (r->sbuf[c]);
        r->sbuf[c] = (float *) n;
      }
      r->samples_len = need;
    }
    if (diff < 0) {
      soff += (long long) -diff * bps;
      bytes -= (long long) -diff * bps;
    } else
      doff += (long long) diff * bps;
    if (bytes > 0)
      for (c = 0; c < r->channels; c++)
        memmove ((char *) r->sbuf[c] + doff, (char *) r->sbuf[c] + soff, (size_t) bytes);
    r->samples_avail += diff;
  }
  return 0;
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

/* make_coeff_gfloat_linear, audio-resampler.c:333-340 */
static void
linear_coeff (int num, int denom, float ic[4])
{
  float x = (float) num / denom;
  ic[0] = ic[2] = x;
  ic[1] = ic[3] = (float) 1.0 - x;
}

/* taps of one phase in FULL mode, built lazily like get_taps_gfloat_full() */
static const float *
phase_taps (OracleArs * r, int phase)
{
  float *res = r->cache + (size_t) phase * r->n_taps;
  if (!r->have[phase] && r->interp_none) {
    /* GST_AUDIO_RESAMPLER_FILTER_INTERPOLATION_NONE (:517-525): the taps of this phase computed directly */
    make_row (r, res, 1.0 - r->n_taps / 2 - (double) phase / r->n_phases);
    r->have[phase] = 1;
  }
  if (!r->have[phase]) {
    int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->n_phases;
    int frac = pos % r->n_phases, i, n = r->n_taps;
    const float *c0 = r->table + (size_t) offset * n, *c1 = c0 + n, *c2 = c1 + n, *c3 = c2 + n;
    float ic[4];
    if (r->linear) {            /* interpolate_gfloat_linear_sse (audio-resampler-x86-sse.c:113-137): c0*f0 + c1*f1 */
      linear_coeff (frac, r->n_phases, ic);
      for (i = 0; i < n; i++) {
        float t0 = c0[i] * ic[0], t1 = c1[i] * ic[1];
        res[i] = t0 + t1;
      }
      r->have[phase] = 1;
      return res;
    }
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
  *n_phases = r->full && r->method != ORACLE_ARS_METHOD_NEAREST ? r->n_phases : 0;  /* only the FULL mode sets n_phases (audio-resampler.c:1173-1178) */
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
    for (l = 0; l < 4 && i + l < len; l++) {    /* a tap count off the lane width ends in the row's zero taps */
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

/* inner_product_gfloat_linear_1_sse (audio-resampler-x86-sse.c:48-74): four lane sums per row, blended per lane as
 * (s0 - s1) * ic[0] + s1 */
static float
dot_linear (const float *a, const float *c0, int stride, int len, const float ic[4])
{
  float s[2][4];
  int i, l, k;
  memset (s, 0, sizeof (s));
  for (i = 0; i < len; i += 4)
    for (k = 0; k < 2; k++)
      for (l = 0; l < 4; l++) {
        float p = a[i + l] * c0[(size_t) k * stride + i + l];
        s[k][l] = s[k][l] + p;
      }
  for (l = 0; l < 4; l++) {
    float d = s[0][l] - s[1][l];
    d = d * ic[0];
    s[0][l] = d + s[1][l];
  }
  return (s[0][0] + s[0][2]) + (s[0][1] + s[0][3]);
}

/* ---- S16 / S32 / F64 (audio-resampler.c macros + audio-resampler-x86-sse2.c / -sse41.c, the
 * implementations an x86 build selects, audio-resampler-x86.h:29-70) --------------------------- */

static int64_t
sat64 (int64_t v, int64_t lo, int64_t hi)
{
  return v < lo ? lo : (v > hi ? hi : v);
}

/* make_coeff_gint16_cubic / _gint32_cubic (audio-resampler.c:350-372) */
static void
cubic_coeff_int (int num, int denom, int prec, int64_t ic[4])
{
  int64_t one = ((int64_t) 1 << prec) - 1;
  int64_t x = ((int64_t) num << prec) / denom, x2, x3;
  if (prec == 15) {             /* type2 = gint32: products wrap in 32 bits */
    int32_t x32 = (int32_t) x, x2_32 = (int32_t) ((int32_t) (x32 * x32) >> 15), x3_32 = (int32_t) ((int32_t) (x2_32 * x32) >> 15);
    int16_t c0 = (int16_t) ((((x3_32 - x32) << 15) / 6) >> 15);
    int16_t c1 = (int16_t) (x32 + ((x2_32 - x3_32) >> 1));
    int16_t c3 = (int16_t) (-(((x32 << 15) / 3) >> 15) + (x2_32 >> 1) - (((x3_32 << 15) / 6) >> 15));
    int16_t c2 = (int16_t) ((int32_t) one - c0 - c1 - c3);
    ic[0] = c0; ic[1] = c1; ic[2] = c2; ic[3] = c3;
    return;
  }
  x2 = (x * x) >> prec;
  x3 = (x2 * x) >> prec;
  ic[0] = (int32_t) ((((x3 - x) << prec) / 6) >> prec);
  ic[1] = (int32_t) (x + ((x2 - x3) >> 1));
  ic[3] = (int32_t) (-(((x << prec) / 3) >> prec) + (x2 >> 1) - (((x3 << prec) / 6) >> prec));
  ic[2] = (int32_t) (one - ic[0] - ic[1] - ic[3]);
}

/* make_coeff_gint16_linear / _gint32_linear (audio-resampler.c:325-332) */
static void
linear_coeff_int (int num, int denom, int prec, int64_t ic[4])
{
  int64_t x = ((int64_t) num << prec) / denom;
  ic[0] = ic[2] = x;
  ic[1] = ic[3] = (((int64_t) 1 << prec) - 1) - x;
}

/* make_coeff_gdouble_linear (:333-340) */
static void
linear_coeff_f64 (int num, int denom, double ic[4])
{
  double x = (double) num / denom;
  ic[0] = ic[2] = x;
  ic[1] = ic[3] = (double) 1.0 - x;
}

/* make_coeff_gdouble_cubic: the literals are float constants promoted to double */
static void
cubic_coeff_f64 (int num, int denom, double ic[4])
{
  double x = (double) num / denom, x2 = x * x, x3 = x2 * x;
  ic[0] = 0.16667f * (x3 - x);
  ic[1] = x + 0.5f * (x2 - x3);
  ic[3] = -0.33333f * x + 0.5f * x2 - 0.16667f * x3;
  ic[2] = (double) 1.0 - ic[0] - ic[1] - ic[3];
}

/* FULL-mode taps of one phase in the handle's format (get_taps_<type>_full + interpolate_<type>_cubic) */
static const void *
phase_taps_any (OracleArs * r, int phase)
{
  size_t n = r->n_taps;
  char *res = (char *) r->cache + (size_t) phase * n * r->bps;
  if (!r->have[phase] && r->interp_none) {
    make_row (r, res, 1.0 - r->n_taps / 2 - (double) phase / r->n_phases);
    r->have[phase] = 1;
  }
  if (!r->have[phase]) {
    int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->n_phases;
    int frac = pos % r->n_phases;
    size_t i;
    const char *c0 = (const char *) r->table + (size_t) offset * n * r->bps;
    if (r->linear && r->fmt == ORACLE_AFMT_S16) {       /* interpolate_gint16_linear_sse2 (audio-resampler-x86-sse2.c:267-299) */
      const int16_t *a = (const int16_t *) c0, *b = a + n;
      int64_t ic[4];
      linear_coeff_int (frac, r->n_phases, 15, ic);
      for (i = 0; i < n; i++) {
        int32_t t = (int32_t) ((uint32_t) (a[i] * (int32_t) ic[0]) + (uint32_t) (b[i] * (int32_t) ic[1]) + (1u << 14));
        ((int16_t *) res)[i] = (int16_t) sat64 (t >> 15, -32768, 32767);
      }
    } else if (r->linear && r->fmt == ORACLE_AFMT_S32) {        /* interpolate_gint32_linear_c (audio-resampler.c:375-390) */
      const int32_t *a = (const int32_t *) c0, *b = a + n;
      int64_t ic[4];
      linear_coeff_int (frac, r->n_phases, 31, ic);
      for (i = 0; i < n; i++) {
        /* (c0 - c1) * x + (c1 << 31), in 64 bits; the store truncates */
        uint64_t t = (uint64_t) (((int64_t) a[i] - (int64_t) b[i]) * ic[0]) + ((uint64_t) (int64_t) b[i] << 31);
        ((int32_t *) res)[i] = (int32_t) (uint32_t) ((int64_t) (t + ((uint64_t) 1 << 30)) >> 31);
      }
    } else if (r->linear) {     /* interpolate_gdouble_linear_sse2 (audio-resampler-x86-sse2.c:344-366) */
      const double *a = (const double *) c0, *b = a + n;
      double ic[4];
      linear_coeff_f64 (frac, r->n_phases, ic);
      for (i = 0; i < n; i++) {
        double t0 = a[i] * ic[0], t1 = b[i] * ic[1];
        ((double *) res)[i] = t0 + t1;
      }
    } else if (r->fmt == ORACLE_AFMT_S16) {    /* interpolate_gint16_cubic_sse2: 32-bit sums, +2^14, >>15, packs */
      const int16_t *a = (const int16_t *) c0, *b = a + n, *c = b + n, *d = c + n;
      int64_t ic[4];
      cubic_coeff_int (frac, r->n_phases, 15, ic);
      for (i = 0; i < n; i++) {
        int32_t t = (int32_t) ((uint32_t) (a[i] * (int32_t) ic[0]) + (uint32_t) (b[i] * (int32_t) ic[1]) +
            (uint32_t) (c[i] * (int32_t) ic[2]) + (uint32_t) (d[i] * (int32_t) ic[3]) + (1u << 14));
        ((int16_t *) res)[i] = (int16_t) sat64 (t >> 15, -32768, 32767);
      }
    } else if (r->fmt == ORACLE_AFMT_S32) {     /* interpolate_gint32_cubic_c */
      const int32_t *a = (const int32_t *) c0, *b = a + n, *c = b + n, *d = c + n;
      int64_t ic[4];
      cubic_coeff_int (frac, r->n_phases, 31, ic);
      for (i = 0; i < n; i++) {
        int64_t t = (int64_t) a[i] * ic[0] + (int64_t) b[i] * ic[1] + (int64_t) c[i] * ic[2] + (int64_t) d[i] * ic[3];
        t = (t + ((int64_t) 1 << 30)) >> 31;
        ((int32_t *) res)[i] = (int32_t) sat64 (t, -((int64_t) 1 << 31), ((int64_t) 1 << 31) - 1);
      }
    } else {                    /* interpolate_gdouble_cubic_sse2: (c0*f0 + c1*f1) + (c2*f2 + c3*f3) */
      const double *a = (const double *) c0, *b = a + n, *c = b + n, *d = c + n;
      double ic[4];
      cubic_coeff_f64 (frac, r->n_phases, ic);
      for (i = 0; i < n; i++) {
        double t0 = a[i] * ic[0], t1 = b[i] * ic[1], t2 = c[i] * ic[2], t3 = d[i] * ic[3];
        t0 = t0 + t1;
        t2 = t2 + t3;
        ((double *) res)[i] = t0 + t2;
      }
    }
    r->have[phase] = 1;
  }
  return res;
}

/* one output sample at window `a`, phase `phase`, written to *o (type by format) */
static void
resample_one_any (OracleArs * r, const void *a, int phase, void *o)
{
  int n = r->n_taps, i, k;
  if (r->copy) {                /* inner_product_<type>_nearest_1_c (audio-resampler.c:602-612): *o = *a */
    memcpy (o, a, r->bps);
    return;
  }
  if (r->fmt == ORACLE_AFMT_S16) {
    const int16_t *x = a;
    if (r->full) {              /* inner_product_gint16_full_1_sse2 */
      const int16_t *t = phase_taps_any (r, phase);
      uint32_t sum = 0;
      for (i = 0; i < n; i++)
        sum += (uint32_t) (x[i] * (int32_t) t[i]);
      *(int16_t *) o = (int16_t) sat64 ((int32_t) (sum + (1u << 14)) >> 15, -32768, 32767);
    } else if (r->linear) {     /* get_taps_gint16_linear + inner_product_gint16_linear_1_sse2 (audio-resampler-x86-sse2.c:57-108) */
      int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
      const int16_t *c = (const int16_t *) r->table + (size_t) offset * n;
      int64_t ic[4];
      uint32_t s[2][4], acc = 0;
      int l;
      linear_coeff_int (pos % r->out_rate, r->out_rate, 15, ic);
      memset (s, 0, sizeof (s));
      /* pmaddwd: 32-bit lane l sums the tap pairs (2l, 2l+1) of every group of eight */
      for (k = 0; k < 2; k++)
        for (i = 0; i < n; i++)
          s[k][(i >> 1) & 3] += (uint32_t) (x[i] * (int32_t) c[(size_t) k * n + i]);
      /* every LANE: srai 15, then pmaddwd of its low 16 bits with the coefficient, before the lanes meet */
      for (k = 0; k < 2; k++)
        for (l = 0; l < 4; l++)
          acc += (uint32_t) ((int32_t) (int16_t) ((int32_t) s[k][l] >> 15) * (int32_t) (int16_t) ic[k]);
      *(int16_t *) o = (int16_t) sat64 ((int32_t) (acc + (1u << 14)) >> 15, -32768, 32767);
    } else {                    /* get_taps_gint16_cubic + inner_product_gint16_cubic_1_sse2 */
      int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
      const int16_t *c = (const int16_t *) r->table + (size_t) offset * n;
      int64_t ic[4];
      uint32_t s[4] = { 0, 0, 0, 0 }, acc = 0;
      cubic_coeff_int (pos % r->out_rate, r->out_rate, 15, ic);
      for (k = 0; k < 4; k++)
        for (i = 0; i < n; i++)
          s[k] += (uint32_t) (x[i] * (int32_t) c[(size_t) k * n + i]);
      for (k = 0; k < 4; k++)   /* srai 15, then pmaddwd with the low 16 bits */
        acc += (uint32_t) ((int32_t) (int16_t) ((int32_t) s[k] >> 15) * (int32_t) ic[k]);
      *(int16_t *) o = (int16_t) sat64 ((int32_t) (acc + (1u << 14)) >> 15, -32768, 32767);
    }
  } else if (r->fmt == ORACLE_AFMT_S32) {
    const int32_t *x = a;
    const int64_t lo = -((int64_t) 1 << 31), hi = ((int64_t) 1 << 31) - 1;
    if (r->full) {              /* inner_product_gint32_full_1_sse41 */
      const int32_t *t = phase_taps_any (r, phase);
      uint64_t sum = 0;
      for (i = 0; i < n; i++)
        sum += (uint64_t) ((int64_t) x[i] * t[i]);
      *(int32_t *) o = (int32_t) sat64 (((int64_t) sum + (1 << 30)) >> 31, lo, hi);
    } else {                    /* inner_product_gint32_cubic_1_sse41 / _linear_1_sse41 (audio-resampler-x86-sse41.c:70-112): the
                                 * same lane arithmetic over four or two table rows */
      int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
      const int32_t *c = (const int32_t *) r->table + (size_t) offset * n;
      int64_t ic[4];
      uint64_t s[4][2], acc = 0;
      int l;
      if (r->linear)
        linear_coeff_int (pos % r->out_rate, r->out_rate, 31, ic);
      else
        cubic_coeff_int (pos % r->out_rate, r->out_rate, 31, ic);
      memset (s, 0, sizeof (s));
      /* two 64-bit lanes per row: pmuldq of (a0,a1 | a2,a3) pairs puts even taps in lane 0, odd taps in lane 1 */
      for (k = 0; k < r->isize; k++)
        for (i = 0; i < n; i++)
          s[k][i & 1] += (uint64_t) ((int64_t) x[i] * c[(size_t) k * n + i]);
      /* each LANE is shifted (srli 31) and multiplied (pmuldq: low 32 bits, signed) before the lanes meet */
      for (k = 0; k < r->isize; k++)
        for (l = 0; l < 2; l++)
          acc += (uint64_t) ((int64_t) (int32_t) (uint32_t) (s[k][l] >> 31) * (int64_t) (int32_t) ic[k]);
      *(int32_t *) o = (int32_t) sat64 (((int64_t) acc + (1 << 30)) >> 31, lo, hi);
    }
  } else {
    const double *x = a;
    if (r->full) {              /* inner_product_gdouble_full_1_sse2: two lanes by tap parity */
      const double *t = phase_taps_any (r, phase);
      double s0 = 0, s1 = 0;
      for (i = 0; i < n; i += 2) {
        double p0 = x[i] * t[i], p1 = i + 1 < n ? x[i + 1] * t[i + 1] : 0.0;
        s0 = s0 + p0;
        s1 = s1 + p1;
      }
      *(double *) o = s0 + s1;
    } else if (r->linear) {     /* inner_product_gdouble_linear_1_sse2 (audio-resampler-x86-sse2.c:195-220) */
      int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
      const double *c = (const double *) r->table + (size_t) offset * n;
      double ic[4], s[2][2], l[2];
      linear_coeff_f64 (pos % r->out_rate, r->out_rate, ic);
      memset (s, 0, sizeof (s));
      for (i = 0; i < n; i += 2)
        for (k = 0; k < 2; k++) {
          double p0 = x[i] * c[(size_t) k * n + i], p1 = x[i + 1] * c[(size_t) k * n + i + 1];
          s[k][0] = s[k][0] + p0;
          s[k][1] = s[k][1] + p1;
        }
      for (i = 0; i < 2; i++) {
        double d = s[0][i] - s[1][i];
        d = d * ic[0];
        l[i] = d + s[1][i];
      }
      *(double *) o = l[0] + l[1];
    } else {                    /* inner_product_gdouble_cubic_1_sse2 */
      int pos = phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
      const double *c = (const double *) r->table + (size_t) offset * n;
      double ic[4], s[4][2], l[2];
      cubic_coeff_f64 (pos % r->out_rate, r->out_rate, ic);
      memset (s, 0, sizeof (s));
      for (i = 0; i < n; i += 2)
        for (k = 0; k < 4; k++) {
          double p0 = x[i] * c[(size_t) k * n + i], p1 = x[i + 1] * c[(size_t) k * n + i + 1];
          s[k][0] = s[k][0] + p0;
          s[k][1] = s[k][1] + p1;
        }
      for (i = 0; i < 2; i++) {
        double t0 = s[0][i] * ic[0], t1 = s[1][i] * ic[1], t2 = s[2][i] * ic[2], t3 = s[3][i] * ic[3];
        t0 = t0 + t1;
        t2 = t2 + t3;
        l[i] = t0 + t2;
      }
      *(double *) o = l[0] + l[1];
    }
  }
}

size_t
oracle_ars_process_any (OracleArs * r, const void *in, size_t in_frames, void *out, size_t out_capacity)
{
  size_t out_frames = oracle_ars_get_out_frames (r, in_frames), avail, need, di, consumed;
  int c, ch = r->channels, samp_index = 0, samp_phase = 0, bps = r->bps;
  if (r->fmt == ORACLE_AFMT_F32)
    return oracle_ars_process (r, in, in_frames, out, out_capacity);
  if (out_frames > out_capacity)
    out_frames = out_capacity;
  if ((size_t) r->skip >= in_frames) {
    r->skip -= (int) in_frames;
    return out_frames;
  }
  r->samp_index += r->skip;
  avail = r->samples_avail;
  if (r->samples_len < in_frames + avail) {
    for (c = 0; c < ch; c++) {
      char *n = calloc (in_frames + avail, bps);
      if (r->sbuf[c])
        memcpy (n, r->sbuf[c], avail * bps);
      free (r->sbuf[c]);
      r->sbuf[c] = (float *) n;
    }
    r->samples_len = in_frames + avail;
  }
  for (c = 0; c < ch; c++) {    /* deinterleave_<type> */
    char *s = (char *) r->sbuf[c] + avail * bps;
    size_t i;
    for (i = 0; i < in_frames; i++) {
      if (in)
        memcpy (s + i * bps, (const char *) in + (i * ch + c) * bps, bps);
      else
        memset (s + i * bps, 0, bps);
    }
  }
  r->samples_avail = avail += in_frames;
  need = r->n_taps + r->samp_index;
  if (avail < need || out_frames == 0)
    return out_frames;
  for (c = 0; c < ch; c++) {
    char *ip = (char *) r->sbuf[c];
    samp_index = r->samp_index;
    samp_phase = r->samp_phase;
    for (di = 0; di < out_frames; di++) {
      resample_one_any (r, ip + (size_t) samp_index * bps, samp_phase, (char *) out + (di * ch + c) * bps);
      samp_index += r->samp_inc;
      samp_phase += r->samp_frac;
      if (samp_phase >= r->out_rate) {
        samp_phase -= r->out_rate;
        samp_index += 1;
      }
    }
    if (avail > (size_t) samp_index)
      memmove (ip, ip + (size_t) samp_index * bps, (avail - samp_index) * bps);
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
      if (r->copy) {            /* inner_product_gfloat_nearest_1_c */
        out[di * ch + c] = ipp[0];
      } else if (r->full) {
        out[di * ch + c] = dot_full (ipp, phase_taps (r, samp_phase), r->n_taps);
      } else {                  /* get_taps_gfloat_cubic, :567-600 */
        int pos = samp_phase * r->oversample, offset = (r->oversample - 1) - pos / r->out_rate;
        float ic[4];
        if (r->linear) {        /* get_taps_gfloat_linear */
          linear_coeff (pos % r->out_rate, r->out_rate, ic);
          out[di * ch + c] = dot_linear (ipp, r->table + (size_t) offset * r->n_taps, r->n_taps, r->n_taps, ic);
        } else {
        cubic_coeff (pos % r->out_rate, r->out_rate, ic);
        out[di * ch + c] = dot_cubic (ipp, r->table + (size_t) offset * r->n_taps, r->n_taps,
            r->n_taps, ic);
        }
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
