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
