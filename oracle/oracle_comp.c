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


/* ======================================================================= 4:2:0 formats
 * blend.c PLANAR_YUV_BLEND (:246-401, I420 / YV12) and NV_YUV_BLEND (:1386-1500, NV12 / NV21),
 * their fill_checker / fill_color (:403-500, :1502-1600), compositor_orc_blend_u8
 * (compositororc.orc:20-36), black / white levels compositor.c:1131-1149.  Default plane layouts
 * (video-info.c:997-1009, :1053-1063) for the output and every pad. */

typedef struct
{
  int n_planes;                 /* 3 planar, 2 semi-planar */
  int stride[3];
  size_t offset[3];
  int pu, pv;                   /* plane of U and of V (planar); semi-planar: U byte index in the pair */
  size_t size;
} Yuv420Layout;

static int
yuv420_layout (int format, int w, int h, Yuv420Layout * l)
{
  int hh = (h + 1) & ~1;
  memset (l, 0, sizeof (*l));
  l->stride[0] = (w + 3) & ~3;
  switch (format) {
    case ORC_FMT_I420: case ORC_FMT_YV12:
      l->n_planes = 3;
      l->stride[1] = l->stride[2] = ((((w + 1) & ~1) / 2) + 3) & ~3;
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->offset[2] = l->offset[1] + (size_t) l->stride[1] * (hh / 2);
      l->size = l->offset[2] + (size_t) l->stride[2] * (hh / 2);
      l->pu = format == ORC_FMT_YV12 ? 2 : 1;
      l->pv = 3 - l->pu;
      return 0;
    case ORC_FMT_NV12: case ORC_FMT_NV21:
      l->n_planes = 2;
      l->stride[1] = l->stride[0];
      l->offset[1] = (size_t) l->stride[0] * hh;
      l->size = l->offset[1] + (size_t) l->stride[1] * (hh / 2);
      l->pu = format == ORC_FMT_NV21 ? 1 : 0;
      return 0;
    default:
      return -1;
  }
}

size_t
oracle_compositor_yuv_size (int format, int width, int height)
{
  Yuv420Layout l;
  return yuv420_layout (format, width, height, &l) ? 0 : l.size;
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

#define SCALE2(v) (-((-(v)) >> 1))      /* GST_VIDEO_FORMAT_INFO_SCALE_WIDTH / _HEIGHT for a 2x sub-sampled component */

int
oracle_compositor_yuv (int format, uint8_t * dst, int width, int height, int background, int range_16_235,
    const OraclePad * pads, int n_pads)
{
  Yuv420Layout L;
  int x, y, i, p;
  if (yuv420_layout (format, width, height, &L))
    return -1;
  /* ---- background: fill_checker_* / fill_color_* / memset 0 */
  {
    int cw = SCALE2 (width), ch = SCALE2 (height);
    int colY = background == ORC_BG_BLACK ? (range_16_235 ? 16 : 0) : (range_16_235 ? 235 : 255);
    for (y = 0; y < height; y++)
      for (x = 0; x < width; x++) {
        static const int tab[] = { 80, 160, 80, 160 };
        uint8_t *d = dst + (size_t) y * L.stride[0] + x;
        if (background == ORC_BG_CHECKER)
          *d = (uint8_t) tab[((y & 0x8) >> 3) + ((x & 0x8) >> 3)];
        else if (background == ORC_BG_TRANSPARENT)
          *d = 0;
        else
          *d = (uint8_t) colY;
      }
    for (p = 1; p < L.n_planes; p++)
      for (y = 0; y < ch; y++)
        memset (dst + L.offset[p] + (size_t) y * L.stride[p], background == ORC_BG_TRANSPARENT ? 0 : 0x80,
            L.n_planes == 2 ? 2 * cw : cw);
  }
  /* ---- pads in z-order */
  for (i = 0; i < n_pads; i++) {
    const OraclePad *pad = &pads[i];
    Yuv420Layout S;
    int xpos = pad->xpos, ypos = pad->ypos, xoffset = 0, yoffset = 0;
    int bw = pad->width, bh = pad->height, cxpos, cypos, cxoff, cyoff, cw, ch;
    const uint8_t *src = pad->data;
    if (yuv420_layout (format, pad->width, pad->height, &S))
      return -1;
    xpos = (xpos + 1) & ~1;     /* GST_ROUND_UP_2 (also for negative values: arithmetic on two's complement) */
    ypos = (ypos + 1) & ~1;
    if (xpos < 0) { xoffset = -xpos; bw -= -xpos; xpos = 0; }
    if (ypos < 0) { yoffset = -ypos; bh -= -ypos; ypos = 0; }
    if (xoffset >= pad->width || yoffset >= pad->height)
      continue;
    if (xpos + bw > width) bw = width - xpos;
    if (ypos + bh > height) bh = height - ypos;
    if (bw <= 0 || bh <= 0)
      continue;
    /* Y */
    blend_plane_u8 (src + S.offset[0] + xoffset + (size_t) yoffset * S.stride[0],
        dst + L.offset[0] + xpos + (size_t) ypos * L.stride[0], S.stride[0], L.stride[0], bw, bh, pad->alpha, pad->op);
    /* chroma */
    cw = SCALE2 (bw); ch = SCALE2 (bh);
    cxpos = xpos == 0 ? 0 : SCALE2 (xpos); cypos = ypos == 0 ? 0 : ypos >> 1;
    cxoff = xoffset == 0 ? 0 : SCALE2 (xoffset); cyoff = yoffset == 0 ? 0 : yoffset >> 1;
    if (L.n_planes == 3) {
      for (p = 1; p < 3; p++)   /* component order U then V; same arithmetic on both planes */
        blend_plane_u8 (src + S.offset[p] + cxoff + (size_t) cyoff * S.stride[p],
            dst + L.offset[p] + cxpos + (size_t) cypos * L.stride[p], S.stride[p], L.stride[p], cw, ch, pad->alpha, pad->op);
    } else {
      blend_plane_u8 (src + S.offset[1] + 2 * cxoff + (size_t) cyoff * S.stride[1],
          dst + L.offset[1] + 2 * cxpos + (size_t) cypos * L.stride[1], S.stride[1], L.stride[1], 2 * cw, ch, pad->alpha, pad->op);
    }
  }
  return 0;
}
