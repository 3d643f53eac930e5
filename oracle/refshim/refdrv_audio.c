/* oracle/refshim/refdrv_audio.c — TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points over the *unmodified* reference resampler compiled in place
 * (gst-libs/gst/audio/audio-resampler.c + its SSE inner products), configured the way the
 * audioresample element configures it: make_options() (gst/audioresample/gstaudioresample.c:374-396)
 * -> GstAudioConverter chain_resample() (gst-libs/gst/audio/audio-converter.c:903-940, flag
 * VARIABLE_RATE from gstaudioresample.c:422) -> gst_audio_resampler_new().
 * F32 interleaved in and out, kaiser method, filter-mode auto, cubic interpolation.
 */
#include <gst/audio/audio.h>
#include "audio-resampler-private.h"

#define OPT_METHOD "GstAudioConverter.resampler-method"

GstAudioResampler *ref_ars_new_fmt (int in_rate, int out_rate, int channels, int quality, int format);

GstAudioResampler *
ref_ars_new (int in_rate, int out_rate, int channels, int quality)
{
  return ref_ars_new_fmt (in_rate, out_rate, channels, quality, GST_AUDIO_FORMAT_F32);
}

/* format: GstAudioFormat of the samples the resampler works on — S16, S32, F32 or F64 in native
 * endianness, the formats audioresample hands over unconverted (audio-converter.c:700-727) */
GstAudioResampler *ref_ars_new_opts (int in_rate, int out_rate, int channels, int quality, int format, int method,
    int filter_mode, int interpolation);

GstAudioResampler *
ref_ars_new_fmt (int in_rate, int out_rate, int channels, int quality, int format)
{
  return ref_ars_new_opts (in_rate, out_rate, channels, quality, format, GST_AUDIO_RESAMPLER_METHOD_KAISER,
      GST_AUDIO_RESAMPLER_FILTER_MODE_AUTO, GST_AUDIO_RESAMPLER_FILTER_INTERPOLATION_CUBIC);
}

/* the element's properties resample-method / sinc-filter-mode / sinc-filter-interpolation as make_options passes them
 * (gstaudioresample.c:374-396); enum values are the reference's (audio-resampler.h:112-160) */
GstAudioResampler *
ref_ars_new_opts (int in_rate, int out_rate, int channels, int quality, int format, int method, int filter_mode,
    int interpolation)
{
  GstStructure *options = gst_structure_new_static_str_empty ("resampler-options");
  GstAudioResampler *r;
  gst_audio_resampler_options_set_quality ((GstAudioResamplerMethod) method, quality, in_rate, out_rate, options);
  gst_structure_set_static_str (options,
      OPT_METHOD, GST_TYPE_AUDIO_RESAMPLER_METHOD, (GstAudioResamplerMethod) method,
      GST_AUDIO_RESAMPLER_OPT_FILTER_MODE, GST_TYPE_AUDIO_RESAMPLER_FILTER_MODE,
      (GstAudioResamplerFilterMode) filter_mode,
      GST_AUDIO_RESAMPLER_OPT_FILTER_MODE_THRESHOLD, G_TYPE_UINT, (guint) 1048576,
      GST_AUDIO_RESAMPLER_OPT_FILTER_INTERPOLATION, GST_TYPE_AUDIO_RESAMPLER_FILTER_INTERPOLATION,
      (GstAudioResamplerFilterInterpolation) interpolation, NULL);
  r = gst_audio_resampler_new ((GstAudioResamplerMethod) method,
      GST_AUDIO_RESAMPLER_FLAG_VARIABLE_RATE, (GstAudioFormat) format, channels, in_rate, out_rate,
      options);
  gst_structure_free (options);
  return r;
}

/* the element's rate change: gst_audio_resample_update_state () builds a fresh option bag with make_options () for the
 * new rates and hands both to gst_audio_converter_update_config () -> gst_audio_resampler_update ()
 * (gstaudioresample.c:398-437, audio-converter.c:349-375) */
int
ref_ars_update (GstAudioResampler * r, int in_rate, int out_rate, int quality, int method, int filter_mode, int interpolation)
{
  GstStructure *options = gst_structure_new_static_str_empty ("resampler-options");
  gboolean ok;
  gst_audio_resampler_options_set_quality ((GstAudioResamplerMethod) method, quality, in_rate, out_rate, options);
  gst_structure_set_static_str (options,
      OPT_METHOD, GST_TYPE_AUDIO_RESAMPLER_METHOD, (GstAudioResamplerMethod) method,
      GST_AUDIO_RESAMPLER_OPT_FILTER_MODE, GST_TYPE_AUDIO_RESAMPLER_FILTER_MODE,
      (GstAudioResamplerFilterMode) filter_mode,
      GST_AUDIO_RESAMPLER_OPT_FILTER_MODE_THRESHOLD, G_TYPE_UINT, (guint) 1048576,
      GST_AUDIO_RESAMPLER_OPT_FILTER_INTERPOLATION, GST_TYPE_AUDIO_RESAMPLER_FILTER_INTERPOLATION,
      (GstAudioResamplerFilterInterpolation) interpolation, NULL);
  ok = gst_audio_resampler_update (r, in_rate, out_rate, options);
  gst_structure_free (options);
  return ok ? 0 : -1;
}

void ref_ars_free (GstAudioResampler * r) { gst_audio_resampler_free (r); }
void ref_ars_reset (GstAudioResampler * r) { gst_audio_resampler_reset (r); }
size_t ref_ars_get_out_frames (GstAudioResampler * r, size_t n) { return gst_audio_resampler_get_out_frames (r, n); }
size_t ref_ars_get_in_frames (GstAudioResampler * r, size_t n) { return gst_audio_resampler_get_in_frames (r, n); }
size_t ref_ars_max_latency (GstAudioResampler * r) { return gst_audio_resampler_get_max_latency (r); }

/* like gst_audio_resample_process() (gstaudioresample.c:743-883): compute out length, resample */
size_t
ref_ars_process (GstAudioResampler * r, const void *in, size_t in_frames, void *out,
    size_t out_capacity)
{
  size_t n = gst_audio_resampler_get_out_frames (r, in_frames);
  gpointer ins[1] = { (gpointer) in }, outs[1] = { out };
  if (n > out_capacity)
    n = out_capacity;
  gst_audio_resampler_resample (r, in ? ins : NULL, in_frames, outs, n);
  return n;
}

int
ref_ars_info (GstAudioResampler * r, int *n_taps, int *n_phases, int *in_step, int *out_step,
    int *filter_mode, int *oversample)
{
  *n_taps = r->n_taps;
  *n_phases = r->n_phases;
  *in_step = r->in_rate;
  *out_step = r->out_rate;
  *filter_mode = r->filter_mode;
  *oversample = r->oversample;
  return 0;
}

/* taps of one phase exactly as the resample loop would cache them (FULL mode) */
gpointer get_taps_gfloat_full (GstAudioResampler * resampler, gint * samp_index, gint * samp_phase,
    gfloat icoeff[4]);
int
ref_ars_phase_taps (GstAudioResampler * r, int phase, float *taps)
{
  gint si = 0, sp = phase;
  gfloat ic[4];
  gfloat *t;
  if (r->filter_mode != GST_AUDIO_RESAMPLER_FILTER_MODE_FULL || phase < 0 || phase >= r->n_phases)
    return -1;
  t = get_taps_gfloat_full (r, &si, &sp, ic);
  memcpy (taps, t, sizeof (float) * r->n_taps);
  return r->n_taps;
}
