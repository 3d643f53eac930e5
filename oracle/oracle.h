/* oracle/oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's raw-frame DSP hot path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker for the CUDA path.  The product library (gstreamer_b200/csrc) never
 * links or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every entry point here is checked byte-for-byte (and
 * float bit-for-bit) against the reference's own sources compiled in place
 * (oracle/_ref/libgstref.so, see oracle/Makefile) in tests/test_oracle_vs_ref.py,
 * and against fixtures generated from that library under tests/golden/.
 *
 * Enum values are GStreamer's own (gst-libs/gst/video/video-format.h:195+,
 * video-color.h:40-83, video-chroma.h:43-52, video-resampler.h:45-49).
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GstVideoFormat subset */
enum { ORC_FMT_I420 = 2, ORC_FMT_YV12 = 3, ORC_FMT_YUY2 = 4, ORC_FMT_UYVY = 5, ORC_FMT_Y42B = 18, ORC_FMT_YVYU = 19,
  ORC_FMT_Y444 = 20, ORC_FMT_RGBx = 7, ORC_FMT_BGRx = 8, ORC_FMT_xRGB = 9,
  ORC_FMT_xBGR = 10, ORC_FMT_RGBA = 11, ORC_FMT_BGRA = 12, ORC_FMT_ARGB = 13, ORC_FMT_ABGR = 14,
  ORC_FMT_NV12 = 23, ORC_FMT_NV21 = 24,
  /* compositor outputs only (planar high bit depth, little endian; GstVideoFormat values) */
  ORC_FMT_I420_10LE = 43, ORC_FMT_I422_10LE = 45, ORC_FMT_Y444_10LE = 47, ORC_FMT_I420_12LE = 73, ORC_FMT_I422_12LE = 75,
  ORC_FMT_Y444_12LE = 77, ORC_FMT_Y444_16LE = 88 };
/* GstVideoResamplerMethod */
enum { ORC_RS_NEAREST = 0, ORC_RS_LINEAR = 1, ORC_RS_CUBIC = 2, ORC_RS_SINC = 3, ORC_RS_LANCZOS = 4 };
/* GstVideoColorMatrix / Range / ChromaSite */
enum { ORC_CM_RGB = 1, ORC_CM_FCC = 2, ORC_CM_BT709 = 3, ORC_CM_BT601 = 4, ORC_CM_SMPTE240M = 5, ORC_CM_BT2020 = 6 };
enum { ORC_RANGE_0_255 = 1, ORC_RANGE_16_235 = 2 };
enum { ORC_SITE_NONE = 1, ORC_SITE_H_COSITED = 2, ORC_SITE_V_COSITED = 4, ORC_SITE_ALT_LINE = 8 };

#define ORACLE_MAX_TAPS 128

/* ---------------- video: resampler tap tables ----------------------------- */
typedef struct {
  int method;          /* ORC_RS_* */
  int max_taps_opt;    /* GST_VIDEO_RESAMPLER_OPT_MAX_TAPS, 0 = unset (128) */
  int n_taps_req;      /* GST_VIDEO_CONVERTER_OPT_RESAMPLER_TAPS, 0 = auto */
  double envelope, sharpness, sharpen, cubic_b, cubic_c;
} OracleResamplerOpts;

/* video-resampler.c:343-429 + :204-288.  taps must hold out_size*ORACLE_MAX_TAPS doubles.
 * returns n_taps (>0) or -1. */
int oracle_resampler_taps (const OracleResamplerOpts * o, int in_size, int out_size,
    uint32_t * offset, double *taps);
/* video-scaler.c:338-388.  returns 1 when the integer taps sum exactly to 1<<precision */
int oracle_quantize_taps (const double *src, int16_t * dst, int n, int precision);

/* ---------------- video: convert + scale ----------------------------------- */
typedef struct {
  int in_format, in_width, in_height;
  int in_stride[4];
  size_t in_offset[4];
  int in_matrix, in_range, in_chroma_site;      /* ORC_CM_*, ORC_RANGE_*, ORC_SITE_* flags */
  int out_format, out_width, out_height;
  int out_stride[4];
  size_t out_offset[4];
  OracleResamplerOpts rs;
  /* YUV output only; 0 = carried over from the input the way the element's caps fixation does when the
   * input caps hold them (gstvideoconvertscale.c:1335-1427) */
  int out_matrix, out_chroma_site;
  /* packed RGB input -> 4:2:0 output: 0 = caps default of the output size (the fixation forwards only primaries and
   * transfer across a YUV/RGB change, gstvideoconvertscale.c:1394-1408); out_range 0 = 16-235 */
  int out_range;
  /* set by oracle_vcs_convert_dest: the chroma resamplers of a 4:2:0 -> other 4:2:0 chain exist because the input
   * size differs from the size of the whole OUTPUT FRAME, even if it equals the destination rectangle's */
  int force_resample;
} OracleVcsDesc;

/* fills the default system-memory layout (video-info.c fill_planes :1053-1063, :890-894)
 * and the caps-default colorimetry / chroma-site (video-info.c:165-225) */
int oracle_vcs_default_desc (OracleVcsDesc * d, int in_format, int in_w, int in_h,
    int out_format, int out_w, int out_h, int method, int max_taps_opt);
size_t oracle_vcs_in_size (const OracleVcsDesc * d);
size_t oracle_vcs_out_size (const OracleVcsDesc * d);
/* packed RGB in, YUV out: the x256 integer matrix of chain_convert (rows Y,U,V; columns R,G,B,offset); 0 when the
 * reference would take its table path (is_no_clip_matrix, video-converter.c:1262-1300), -2 otherwise */
int oracle_vcs_matrix_rgb2yuv (const OracleVcsDesc * d, int im[4][4]);
/* the fast AYUV->ARGB matrix parameters p1..p5 (video-converter.c:1209-1216, :1324-1442) */
int oracle_vcs_matrix (const OracleVcsDesc * d, int p[5], int im[4][4]);
/* whole-frame conversion; 0 on success */
int oracle_vcs_convert (const OracleVcsDesc * d, const uint8_t * in, uint8_t * out);

/* the element's add-borders geometry (gstvideoconvertscale.c:926-952): borders that keep the display aspect ratio
 * when both pixel aspect ratios are 1/1.  Returns dest x, y, width, height inside the out_w x out_h frame. */
void oracle_vcs_borders (int in_w, int in_h, int out_w, int out_h, int dest[4]);
/* whole-frame conversion into the destination rectangle of the output frame, the rest filled with the border colour
 * (GST_VIDEO_CONVERTER_OPT_DEST_*, fill-border, border-argb default 0xff000000: video-converter.c:2333-2366,
 * setup_borderline :2189-2258, convert_fill_border :7190-7300).  d describes the WHOLE output frame. */
int oracle_vcs_convert_dest (const OracleVcsDesc * d, int dest_x, int dest_y, int dest_w, int dest_h,
    uint32_t border_argb, const uint8_t * in, uint8_t * out);

/* ---------------- compositor ------------------------------------------------- */
enum { ORC_BG_CHECKER = 0, ORC_BG_BLACK = 1, ORC_BG_WHITE = 2, ORC_BG_TRANSPARENT = 3 };
enum { ORC_OP_SOURCE = 0, ORC_OP_OVER = 1, ORC_OP_ADD = 2 };
typedef struct {
  const uint8_t *data; int width, height, stride;
  int xpos, ypos; double alpha; int op;
} OraclePad;
/* blend.c:42-159, :178-237; compositor.c:1619-1697.  format: RGBA/BGRA/ARGB/ABGR family byte order
 * is given by out_format (alpha byte index 3 for RGBA/BGRA, 0 for ARGB/ABGR). */
int oracle_compositor (int out_format, uint8_t * dst, int width, int height, int stride,
    int background, const OraclePad * pads, int n_pads);

/* 4:2:0 output and pads (I420, YV12, NV12, NV21; default plane layouts): blend.c PLANAR_YUV_BLEND / NV_YUV_BLEND */
size_t oracle_compositor_yuv_size (int format, int width, int height);
int oracle_compositor_yuv (int format, uint8_t * dst, int width, int height, int background, int range_16_235,
    const OraclePad * pads, int n_pads);

/* ---------------- audio resampler -------------------------------------------- */
typedef struct OracleArs OracleArs;
/* gstaudioresample.c:374-396 (element option plumbing) + audio-resampler.c:1344-1424.
 * quality 0..10, F32 interleaved, kaiser method, filter-mode auto, cubic interpolation. */
OracleArs *oracle_ars_new (int in_rate, int out_rate, int channels, int quality);
/* sample formats the element hands to the resampler unconverted (audio-converter.c:700-727):
 * F32 (default), S16, S32, F64, native endianness */
enum { ORACLE_AFMT_F32 = 0, ORACLE_AFMT_S16 = 1, ORACLE_AFMT_S32 = 2, ORACLE_AFMT_F64 = 3 };
OracleArs *oracle_ars_new_fmt (int in_rate, int out_rate, int channels, int quality, int fmt);
/* the element's resample-method / sinc-filter-mode / sinc-filter-interpolation properties (gstaudioresample.c:160-186,
 * make_options :374-396); values are the reference's enums (audio-resampler.h:104-106, :138-140, :169-178).  Restated:
 * every method, every filter mode, every table interpolation. */
enum { ORACLE_ARS_METHOD_NEAREST = 0, ORACLE_ARS_METHOD_LINEAR = 1, ORACLE_ARS_METHOD_CUBIC = 2,
  ORACLE_ARS_METHOD_BLACKMAN_NUTTALL = 3, ORACLE_ARS_METHOD_KAISER = 4 };
enum { ORACLE_ARS_MODE_INTERPOLATED = 0, ORACLE_ARS_MODE_FULL = 1, ORACLE_ARS_MODE_AUTO = 2 };
enum { ORACLE_ARS_INTERP_NONE = 0, ORACLE_ARS_INTERP_LINEAR = 1, ORACLE_ARS_INTERP_CUBIC = 2 };
OracleArs *oracle_ars_new_opts (int in_rate, int out_rate, int channels, int quality, int fmt, int method,
    int filter_mode, int interpolation);
/* like oracle_ars_process for any format: in/out are interleaved samples of the handle's format */
size_t oracle_ars_process_any (OracleArs * r, const void *in, size_t in_frames, void *out, size_t out_capacity);
void oracle_ars_free (OracleArs * r);
void oracle_ars_reset (OracleArs * r);
/* rate change on a live stream (gst_audio_resampler_update with the element's fresh option bag, audio-resampler.c:1503) */
int oracle_ars_update (OracleArs * r, int in_rate, int out_rate);
size_t oracle_ars_get_out_frames (OracleArs * r, size_t in_frames);
size_t oracle_ars_get_in_frames (OracleArs * r, size_t out_frames);
size_t oracle_ars_max_latency (OracleArs * r);
/* in == NULL pushes silence (drain). returns frames written */
size_t oracle_ars_process (OracleArs * r, const float *in, size_t in_frames, float *out,
    size_t out_frames);
/* introspection for the product's host-side plan tests */
int oracle_ars_info (OracleArs * r, int *n_taps, int *n_phases, int *in_step, int *out_step,
    int *filter_mode, int *oversample);
/* exact float taps of one phase as the reference would cache them (n_taps floats) */
int oracle_ars_phase_taps (OracleArs * r, int phase, float *taps);

#ifdef __cplusplus
}
#endif
#endif
