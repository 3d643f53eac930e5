/* oracle/oracle_comp.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Restatement of the compositor's per-frame pixel work for packed 4x8-bit formats with alpha:
 *   background            compositor.c:1619-1675 (_draw_background); blend.c:178-237
 *   pad loop, z-order     compositor.c:1678-1697 (blend_pads)
 *   clip + dispatch       blend.c:42-97 (BLEND_A32), :99-159 (_blend_loop / _overlay_loop)
 *   pixel arithmetic      compositororc.orc:163-264 (blend_argb/bgra), :196-223 (source),
 *                         :343-486 (overlay), :488-560 (overlay addition); div255w, divluw
 * Pinned byte-for-byte against the reference's own blend.c + compositororc-dist.c
 * (oracle/_ref, refdrv_comp.c) by tests/test_oracle_vs_ref.py.
 */
#include "oracle.h"

#include <string.h>

static inline unsigned
div255w (unsigned x)
{
  return ((x & 0xffff) * 0x8081u) >> 23;
}

static inline unsigned
divluw (unsigned num, unsigned den)
{
  unsigned q;
  den &= 0xff;
  if (den == 0)
    return 255;
  q = (num & 0xffff) / den;
  return q > 255 ? 255 : q;
}

static void
px_source (uint8_t * d, const uint8_t * s, unsigned s_alpha, int ai)
{
  /* compositor_orc_source_*: copy colour, alpha = div255 (As * alpha) */
  unsigned a = div255w (s[ai] * s_alpha);
  memcpy (d, s, 4);
  d[ai] = (uint8_t) a;
}

static void
px_blend (uint8_t * d, const uint8_t * s, unsigned s_alpha, int ai)
{
  /* compositor_orc_blend_*: every byte (alpha byte included) is blended, then alpha := 0xff */
  unsigned a = div255w (s[ai] * s_alpha), c;
  for (c = 0; c < 4; c++)
    d[c] = (uint8_t) div255w (((s[c] * a) & 0xffff) + ((d[c] * (255 - a)) & 0xffff));
  d[ai] = 0xff;
}

static void
px_overlay (uint8_t * d, const uint8_t * s, unsigned s_alpha, int ai, int addition)
{
  /* compositor_orc_overlay_* / _addition */
  unsigned as = div255w (s[ai] * s_alpha);
  unsigned ad = div255w (d[ai] * (255 - as));
  unsigned asum = (ad + as) & 0xffff, c;
  unsigned dst_alpha = d[ai];
  for (c = 0; c < 4; c++) {
    unsigned v = (((s[c] * as) & 0xffff) + ((d[c] * ad) & 0xffff)) & 0xffff;
    d[c] = (uint8_t) divluw (v, asum);
  }
  d[ai] = (uint8_t) (addition ? (dst_alpha + as) : asum);
}

int
oracle_compositor (int out_format, uint8_t * dst, int width, int height, int stride,
    int background, const OraclePad * pads, int n_pads)
{
  int ai, x, y, i;
  switch (out_format) {
    case ORC_FMT_BGRA: case ORC_FMT_RGBA: ai = 3; break;
    case ORC_FMT_ARGB: case ORC_FMT_ABGR: ai = 0; break;
    default: return -1;
  }
  /* background */
  for (y = 0; y < height; y++) {
    uint8_t *row = dst + (size_t) y * stride;
    for (x = 0; x < width; x++) {
      uint8_t *p = row + 4 * x;
      switch (background) {
        case ORC_BG_CHECKER:{
          static const int tab[] = { 80, 160, 80, 160 };
          int v = tab[((y & 0x8) >> 3) + ((x & 0x8) >> 3)];
          p[0] = p[1] = p[2] = p[3] = (uint8_t) v;
          p[ai] = 0xff;
          break;
        }
        case ORC_BG_BLACK: p[0] = p[1] = p[2] = p[3] = 0; p[ai] = 0xff; break;
        case ORC_BG_WHITE: p[0] = p[1] = p[2] = p[3] = 0xff; break;
        default: p[0] = p[1] = p[2] = p[3] = 0; break;
      }
    }
  }
  /* pads in z-order */
  for (i = 0; i < n_pads; i++) {
    const OraclePad *pad = &pads[i];
    int s_alpha = (int) (pad->alpha * 255), xpos = pad->xpos, ypos = pad->ypos;
    int sw = pad->width, sh = pad->height;
    const uint8_t *src = pad->data;
    s_alpha = s_alpha < 0 ? 0 : (s_alpha > 255 ? 255 : s_alpha);
    if (s_alpha == 0)
      continue;
    if (xpos < 0) { src += -xpos * 4; sw -= -xpos; xpos = 0; }
    if (ypos < 0) { src += (size_t) (-ypos) * pad->stride; sh -= -ypos; ypos = 0; }
    if (xpos + sw > width) sw = width - xpos;
    if (ypos + sh > height) sh = height - ypos;
    if (sw <= 0 || sh <= 0)
      continue;
    for (y = 0; y < sh; y++) {
      const uint8_t *s = src + (size_t) y * pad->stride;
      uint8_t *d = dst + (size_t) (ypos + y) * stride + 4 * xpos;
      for (x = 0; x < sw; x++, s += 4, d += 4) {
        if (pad->op == ORC_OP_SOURCE) {
          if (s_alpha == 255) memcpy (d, s, 4);
          else px_source (d, s, s_alpha, ai);
        } else if (background != ORC_BG_TRANSPARENT) {
          px_blend (d, s, s_alpha, ai);
        } else {
          px_overlay (d, s, s_alpha, ai, pad->op == ORC_OP_ADD);
        }
      }
    }
  }
  return 0;
}


/* ======================================================================= planar / semi-planar YUV formats
 * blend.c PLANAR_YUV_BLEND (:246-401: I420 / YV12 / Y444 / Y42B at 8 bits, :613-646 the little-endian 10 / 12 / 16-bit
 * I420 / I422 / Y444 families with compositor_orc_blend_u10 / _u12 / _u16) and NV_YUV_BLEND (:1386-1500, NV12 / NV21),
 * their fill_checker / fill_color (:403-500, :502-587 PLANAR_YUV_HIGH_*, :1502-1600), compositor_orc_blend_u8 .. _u16
 * (compositororc.orc:20-100), black / white levels compositor.c:1131-1149 (gst_video_color_range_offsets,
 * video-color.c:204-252).  Default plane layouts (video-info.c:997-1063, :1142-1188) for the output and every pad. */

typedef struct
{
  int n_planes;                 /* 3 planar, 2 semi-planar */
  int stride[3];
  size_t offset[3];
  int pu, pv;                   /* plane of U and of V (planar); semi-planar: U byte index in the pair */
  int ws, hs;                   /* chroma sub-sampling shifts (w_sub / h_sub of components 1, 2) */
  int xr, yr;                   /* x_round / y_round of the format's blend function: positions round UP to a multiple */
  int es, nbits;                /* bytes per sample (pstride of a planar component), significant bits */
  size_t size;
} YuvLayout;

#define SCALE_SUB(v, sub) (-((-(v)) >> (sub)))     /* GST_VIDEO_SUB_SCALE */
#define ROUND_UP_N(v, n) (((v) + (n) - 1) & ~((n) - 1))

static int
yuv_layout (int format, int w, int h, YuvLayout * l)
{
  int hh = (h + 1) & ~1;
  memset (l, 0, sizeof (*l));
  l->es = 1; l->nbits = 8; l->ws = l->hs = 1; l->xr = l->yr = 2; l->n_planes = 3; l->pu = 1; l->pv = 2;
  switch (format) {
    case ORC_FMT_I420: case ORC_FMT_YV12:
      l->stride[0] = (w + 3) & ~3;
      l->stride[1] = l->stride[2] = ((((w + 1) & ~1) / 2) + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->offset[2] = l->offset[1] + (size_t) l->stride[1] * (hh / 2);
      l->size = l->offset[2] + (size_t) l->stride[2] * (hh / 2);
      l->pu = format == ORC_FMT_YV12 ? 2 : 1;
      l->pv = 3 - l->pu;
      return 0;
    case ORC_FMT_NV12: case ORC_FMT_NV21:
      l->n_planes = 2;
      l->stride[0] = (w + 3) & ~3;
      l->stride[1] = l->stride[0];
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->size = l->offset[1] + (size_t) l->stride[1] * (hh / 2);
      l->pu = format == ORC_FMT_NV21 ? 1 : 0;
      return 0;
    case ORC_FMT_Y42B:            /* video-info.c:1020-1029 */
      l->hs = 0; l->yr = 1;
      l->stride[0] = (w + 3) & ~3;
      l->stride[1] = l->stride[2] = ((w + 7) & ~7) / 2;
      l->offset[1] = (size_t) l->stride[0] * h;
      l->offset[2] = l->offset[1] + (size_t) l->stride[1] * h;
      l->size = l->offset[2] + (size_t) l->stride[2] * h;
      return 0;
    case ORC_FMT_Y444:            /* :1030-1041 */
      l->ws = l->hs = 0; l->xr = l->yr = 1;
      l->stride[0] = l->stride[1] = l->stride[2] = (w + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * h;
      l->offset[2] = l->offset[1] * 2;
      l->size = (size_t) l->stride[0] * h * 3;
      return 0;
    case ORC_FMT_I420_10LE: case ORC_FMT_I420_12LE:      /* :1142-1156 */
      l->es = 2; l->nbits = format == ORC_FMT_I420_10LE ? 10 : 12;
      l->stride[0] = (w * 2 + 3) & ~3;
      l->stride[1] = l->stride[2] = (w + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->offset[2] = l->offset[1] + (size_t) l->stride[1] * (hh / 2);
      l->size = l->offset[2] + (size_t) l->stride[2] * (hh / 2);
      return 0;
    case ORC_FMT_I422_10LE: case ORC_FMT_I422_12LE:      /* :1157-1169 */
      l->es = 2; l->nbits = format == ORC_FMT_I422_10LE ? 10 : 12; l->hs = 0; l->yr = 1;
      l->stride[0] = (w * 2 + 3) & ~3;
      l->stride[1] = l->stride[2] = (w + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->offset[2] = l->offset[1] + (size_t) l->stride[1] * hh;
      l->size = l->offset[2] + (size_t) l->stride[2] * hh;
      return 0;
    case ORC_FMT_Y444_10LE: case ORC_FMT_Y444_12LE: case ORC_FMT_Y444_16LE:      /* :1170-1188 */
      l->es = 2; l->nbits = format == ORC_FMT_Y444_10LE ? 10 : (format == ORC_FMT_Y444_12LE ? 12 : 16);
      l->ws = l->hs = 0; l->xr = l->yr = 1;
      l->stride[0] = l->stride[1] = l->stride[2] = (w * 2 + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * h;
      l->offset[2] = l->offset[1] * 2;
      l->size = (size_t) l->stride[0] * h * 3;
      return 0;
    default:
      return -1;
  }
}

size_t
oracle_compositor_yuv_size (int format, int width, int height)
{
  YuvLayout l;
  return yuv_layout (format, width, height, &l) ? 0 : l.size;
}

/* compositor_orc_blend_u8 over a w x h byte rectangle, or the alpha == 1 / SOURCE row copy */
static void
blend_plane_u8 (const uint8_t * src, uint8_t * dest, int sstride, int dstride, int w, int h, double alpha, int op)
{
  int x, y, b_alpha;
  if (op == ORC_OP_SOURCE)
    alpha = 1.0;
  if (alpha == 0.0 || w <= 0 || h <= 0)
    return;
  if (alpha == 1.0) {
    for (y = 0; y < h; y++)
      memcpy (dest + (size_t) y * dstride, src + (size_t) y * sstride, w);
    return;
  }
  b_alpha = (int) (alpha * 255);
  b_alpha = b_alpha < 0 ? 0 : (b_alpha > 255 ? 255 : b_alpha);
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      uint8_t *d = dest + (size_t) y * dstride + x;
      uint16_t t1 = *d, t2 = src[(size_t) y * sstride + x];
      t2 = (uint16_t) (t2 - t1);
      t2 = (uint16_t) (t2 * (uint16_t) b_alpha);        /* mullw */
      t1 = (uint16_t) (t1 << 8);
      t2 = (uint16_t) (t1 + t2);
      t2 = (uint16_t) (t2 >> 8);
      *d = (uint8_t) ((int16_t) t2 < 0 ? 0 : ((int16_t) t2 > 255 ? 255 : t2));       /* convsuswb */
    }
}

/* compositor_orc_blend_u10 / _u12 / _u16 (compositororc.orc:38-100) over w x h little-endian 16-bit samples: 32-bit wrapping
 * arithmetic, logical shift, signed-to-unsigned-word saturation; alpha = CLAMP ((int) (alpha * range), 0, range) */
static void
blend_plane_u16 (const uint8_t * src, uint8_t * dest, int sstride, int dstride, int w, int h, double alpha, int op, int nbits)
{
  int x, y, range = (1 << nbits) - 1, b_alpha;
  if (op == ORC_OP_SOURCE)
    alpha = 1.0;
  if (alpha == 0.0 || w <= 0 || h <= 0)
    return;
  if (alpha == 1.0) {
    for (y = 0; y < h; y++)
      memcpy (dest + (size_t) y * dstride, src + (size_t) y * sstride, (size_t) w * 2);
    return;
  }
  b_alpha = (int) (alpha * range);
  b_alpha = b_alpha < 0 ? 0 : (b_alpha > range ? range : b_alpha);
  for (y = 0; y < h; y++)
    for (x = 0; x < w; x++) {
      uint8_t *d = dest + (size_t) y * dstride + 2 * x;
      const uint8_t *s = src + (size_t) y * sstride + 2 * x;
      uint32_t t1 = (uint32_t) d[0] | ((uint32_t) d[1] << 8), t2 = (uint32_t) s[0] | ((uint32_t) s[1] << 8);
      int32_t r;
      t2 = t2 - t1;                                     /* subl */
      t2 = t2 * (uint32_t) b_alpha;                     /* mulll: low 32 bits */
      t1 = t1 << nbits;                                 /* shll */
      t2 = t1 + t2;                                     /* addl */
      t2 = t2 >> nbits;                                 /* shrul */
      r = (int32_t) t2;                                 /* convsuslw */
      r = r < 0 ? 0 : (r > 65535 ? 65535 : r);
      d[0] = (uint8_t) r; d[1] = (uint8_t) (r >> 8);
    }
}

static void
blend_plane (const YuvLayout * L, const uint8_t * src, uint8_t * dest, int sstride, int dstride, int w, int h, double alpha, int op)
{
  if (L->es == 1) blend_plane_u8 (src, dest, sstride, dstride, w, h, alpha, op);
  else blend_plane_u16 (src, dest, sstride, dstride, w, h, alpha, op, L->nbits);
}

int
oracle_compositor_yuv (int format, uint8_t * dst, int width, int height, int background, int range_16_235,
    const OraclePad * pads, int n_pads)
{
  YuvLayout L;
  int x, y, i, p;
  if (yuv_layout (format, width, height, &L))
    return -1;
  /* ---- background: fill_checker_* / fill_color_* / memset 0 */
  {
    int cw = SCALE_SUB (width, L.ws), ch = SCALE_SUB (height, L.hs);
    int sh = L.nbits - 8;
    /* gst_video_color_range_offsets: 16-235 -> offset 1 << (depth - 4), scale 219 << (depth - 8); 0-255 -> 0, (1 << depth) - 1 */
    int colY = background == ORC_BG_BLACK ? (range_16_235 ? 16 << sh : 0) : (range_16_235 ? 235 << sh : (1 << L.nbits) - 1);
    int colC = 1 << (L.nbits - 1);
    for (y = 0; y < height; y++)
      for (x = 0; x < width; x++) {
        static const int tab[] = { 80, 160, 80, 160 };
        uint8_t *d = dst + (size_t) y * L.stride[0] + (size_t) x * L.es;
        int v;
        if (background == ORC_BG_CHECKER)
          v = tab[((y & 0x8) >> 3) + ((x & 0x8) >> 3)] << sh;
        else if (background == ORC_BG_TRANSPARENT)
          v = 0;
        else
          v = colY;
        d[0] = (uint8_t) v;
        if (L.es == 2) d[1] = (uint8_t) (v >> 8);
      }
    for (p = 1; p < L.n_planes; p++)
      for (y = 0; y < ch; y++) {
        uint8_t *row = dst + L.offset[p] + (size_t) y * L.stride[p];
        int n = L.n_planes == 2 ? 2 * cw : cw, v = background == ORC_BG_TRANSPARENT ? 0 : colC;
        for (x = 0; x < n; x++) {
          row[(size_t) x * L.es] = (uint8_t) v;
          if (L.es == 2) row[(size_t) x * 2 + 1] = (uint8_t) (v >> 8);
        }
      }
  }
  /* ---- pads in z-order */
  for (i = 0; i < n_pads; i++) {
    const OraclePad *pad = &pads[i];
    YuvLayout S;
    int xpos = pad->xpos, ypos = pad->ypos, xoffset = 0, yoffset = 0;
    int bw = pad->width, bh = pad->height, cxpos, cypos, cxoff, cyoff, cw, ch;
    const uint8_t *src = pad->data;
    if (yuv_layout (format, pad->width, pad->height, &S))
      return -1;
    xpos = ROUND_UP_N (xpos, L.xr);     /* x_round / y_round (also for negative values: arithmetic on two's complement) */
    ypos = ROUND_UP_N (ypos, L.yr);
    if (xpos < 0) { xoffset = -xpos; bw -= -xpos; xpos = 0; }
    if (ypos < 0) { yoffset = -ypos; bh -= -ypos; ypos = 0; }
    if (xoffset >= pad->width || yoffset >= pad->height)
      continue;
    if (xpos + bw > width) bw = width - xpos;
    if (ypos + bh > height) bh = height - ypos;
    if (bw <= 0 || bh <= 0)
      continue;
    /* Y */
    blend_plane (&L, src + S.offset[0] + (size_t) xoffset * L.es + (size_t) yoffset * S.stride[0],
        dst + L.offset[0] + (size_t) xpos * L.es + (size_t) ypos * L.stride[0], S.stride[0], L.stride[0], bw, bh, pad->alpha, pad->op);
    /* chroma: widths / positions by GST_VIDEO_FORMAT_INFO_SCALE_WIDTH (round up), rows by a plain shift (blend.c:371-376) */
    cw = SCALE_SUB (bw, L.ws); ch = SCALE_SUB (bh, L.hs);
    cxpos = xpos == 0 ? 0 : SCALE_SUB (xpos, L.ws); cypos = ypos == 0 ? 0 : ypos >> L.hs;
    cxoff = xoffset == 0 ? 0 : SCALE_SUB (xoffset, L.ws); cyoff = yoffset == 0 ? 0 : yoffset >> L.hs;
    if (L.n_planes == 3) {
      for (p = 1; p < 3; p++)   /* component order U then V; same arithmetic on both planes */
        blend_plane (&L, src + S.offset[p] + (size_t) cxoff * L.es + (size_t) cyoff * S.stride[p],
            dst + L.offset[p] + (size_t) cxpos * L.es + (size_t) cypos * L.stride[p], S.stride[p], L.stride[p], cw, ch, pad->alpha, pad->op);
    } else {
      blend_plane_u8 (src + S.offset[1] + 2 * cxoff + (size_t) cyoff * S.stride[1],
          dst + L.offset[1] + 2 * cxpos + (size_t) cypos * L.stride[1], S.stride[1], L.stride[1], 2 * cw, ch, pad->alpha, pad->op);
    }
  }
  return 0;
}
