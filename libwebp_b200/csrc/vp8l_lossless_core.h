// vp8l_lossless_core.h -- whole-picture VP8L (lossless WebP) on the device: the last stage.
//
// The entropy-coded stream of a lossless picture is the same syntax as the VP8L payload of an ALPH chunk behind a
// 5-byte header, so passes A and B are vp8l_alpha_core.h's (alph_parse_header with lossless = 1, alph_decode_pixels:
// the 32-bit DecodeImageData shape, vp8l_dec.c:1138-1293). What differs is what becomes of the ARGB words:
//   VP8LDecodeImage / ProcessRows / ApplyInverseTransforms / EmitRows          src/dec/vp8l_dec.c:700-808, 1706-1775
//   VP8LInverseTransform, VP8LColorIndexInverseTransform, MapARGB              src/dsp/lossless.c:341-442
//   VP8LConvertFromBGRA (+ WebPApplyAlphaMultiply / 4444 for the premultiplied modes)   src/dsp/lossless.c:562-637
// The reference converts 16 rows at a time through a cache; the inverse transforms only ever look up and to the left, so
// running them over the whole picture gives the same words. One thread block per picture: transforms in place, then
// one thread per pixel of the output window (options.use_cropping: the window is NOT snapped to even coordinates for
// lossless pictures, webp_dec.c:816-820; options.flip: rows written bottom-up).
// Rows below the window are never decoded (DecodeImageData stops at io->crop_bottom), so whatever the transforms make of
// them is never looked at.
#ifndef LIBWEBP_B200_VP8L_LOSSLESS_CORE_H_
#define LIBWEBP_B200_VP8L_LOSSLESS_CORE_H_

#include "vp8_dev.h"
#include "vp8_pixel_core.h"
#include "vp8l_alpha_core.h"

// ARGB of picture pixel (sx, sy) after the inverse transforms: the palette look-up / unbundling of a COLOR_INDEXING
// transform happens here (VP8LColorIndexInverseTransform, lossless.c:341-385: the index travels in green).
AL_FN uint32_t vp8l_fetch(const AlphaHdr* hd, const uint32_t* px, int sx, int sy) {
  const int xs = hd->px_stride;
  if (hd->ntrans > 0 && hd->ttype[0] == AL_T_COLOR_INDEXING) {
    const int bits = hd->tbits[0], bpp = 8 >> bits;
    const uint32_t packed = (px[(size_t)sy * xs + (sx >> bits)] >> 8) & 0xff;
    return hd->palette[(packed >> ((sx & ((1 << bits) - 1)) * bpp)) & ((1u << bpp) - 1u)];
  }
  return px[(size_t)sy * xs + sx];
}

// VP8RGBToY / VP8RGBToU / VP8RGBToV (src/dsp/yuv.h:186-204), YUV_FIX = 16.
AL_FN int vp8l_clip_uv(int uv, int rounding) {
  uv = (uv + rounding + (128 << 18)) >> 18;
  return ((uv & ~0xff) == 0) ? uv : (uv < 0) ? 0 : 255;
}
// One row's contribution to chroma sample ux: two pixels summed and doubled, the last pixel of an odd width times four
// (WebPConvertARGBToUV_C, src/dsp/yuv.c:129-170).
AL_FN void vp8l_row_uv(const AlphaHdr* hd, const uint32_t* px, const ImgDesc& im, int ux, int y, int* u, int* v) {
  const int x0 = 2 * ux;
  const uint32_t v0 = vp8l_fetch(hd, px, x0 + im.crop_x, y + im.crop_y);
  int r, g, b;
  if (x0 + 1 < im.out_w) {
    const uint32_t v1 = vp8l_fetch(hd, px, x0 + 1 + im.crop_x, y + im.crop_y);
    r = (int)(((v0 >> 15) & 0x1fe) + ((v1 >> 15) & 0x1fe)); g = (int)(((v0 >> 7) & 0x1fe) + ((v1 >> 7) & 0x1fe));
    b = (int)(((v0 << 1) & 0x1fe) + ((v1 << 1) & 0x1fe));
  } else {
    r = (int)((v0 >> 14) & 0x3fc); g = (int)((v0 >> 6) & 0x3fc); b = (int)((v0 << 2) & 0x3fc);
  }
  *u = vp8l_clip_uv(-9719 * r - 19081 * g + 28800 * b, 1 << 17);
  *v = vp8l_clip_uv(28800 * r - 24116 * g - 4684 * b, 1 << 17);
}

// MODE_YUV / MODE_YUVA from a lossless picture (EmitRowsYUVA / ConvertToYUVA, vp8l_dec.c:660-686): luma per pixel; chroma
// per row from horizontal pairs, even rows stored, odd rows averaged into them with rounding; alpha copied.
// Layout as the VP8 output stage writes it: y | u | v [| a], strides w, (w+1)/2, (w+1)/2, w; flip reverses every plane.
AL_FN void vp8l_emit_yuva(const AlphaHdr* hd, const ImgDesc& im, const uint32_t* px, uint8_t* out, int tid, int nt) {
  const int w = im.out_w, h = im.out_h, uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
  const int flip = (im.flags & VP8B_FLAG_FLIP) != 0;
  uint8_t* yo = out;
  uint8_t* uo = out + (size_t)im.out_stride * h;
  uint8_t* vo = uo + (size_t)uvw * uvh;
  uint8_t* ao = vo + (size_t)uvw * uvh;
  const size_t total = (size_t)w * (size_t)h;
  for (size_t i = (size_t)tid; i < total; i += (size_t)nt) {
    const int x = (int)(i % (size_t)w), y = (int)(i / (size_t)w);
    const uint32_t p = vp8l_fetch(hd, px, x + im.crop_x, y + im.crop_y);
    const int luma = 16839 * (int)((p >> 16) & 0xff) + 33059 * (int)((p >> 8) & 0xff) + 6420 * (int)(p & 0xff);
    const int yd = flip ? h - 1 - y : y;
    yo[(size_t)yd * im.out_stride + x] = (uint8_t)((luma + (1 << 15) + (16 << 16)) >> 16);
    if (im.csp == 12) ao[(size_t)yd * w + x] = (uint8_t)(p >> 24);
  }
  const size_t uv_total = (size_t)uvw * (size_t)uvh;
  for (size_t i = (size_t)tid; i < uv_total; i += (size_t)nt) {
    const int ux = (int)(i % (size_t)uvw), uy = (int)(i / (size_t)uvw);
    int u, v;
    vp8l_row_uv(hd, px, im, ux, 2 * uy, &u, &v);
    if (2 * uy + 1 < h) {
      int u1, v1;
      vp8l_row_uv(hd, px, im, ux, 2 * uy + 1, &u1, &v1);
      u = (u + u1 + 1) >> 1; v = (v + v1 + 1) >> 1;
    }
    const int yd = flip ? uvh - 1 - uy : uy;
    uo[(size_t)yd * uvw + ux] = (uint8_t)u;
    vo[(size_t)yd * uvw + ux] = (uint8_t)v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// options.use_scaling on a lossless picture (AllocateAndInitRescaler, EmitRescaledRowsRGBA / EmitRescaledRowsYUVA, Export /
// ExportYUVA, vp8l_dec.c:560-645, 688-737): the window's ARGB words are multiplied by their alpha (WebPMultARGBRows,
// alpha_processing.c:140-158), the four byte channels go through four rescalers of the same geometry (num_channels = 4),
// every exported row is divided by its rescaled alpha again (WebPMultARGBRow inverse) and converted like an unscaled row.
// `pm` = the premultiplied window as interleaved B, G, R, A bytes (4 * out_w * out_h bytes of scratch).
AL_FN uint32_t vp8l_premultiply(uint32_t argb) {
  if (argb >= 0xff000000u) return argb;
  if (argb <= 0x00ffffffu) return 0;
  const uint32_t a = argb >> 24;
  return (argb & 0xff000000u) | mult_by_alpha(argb & 0xff, a, 0) | (mult_by_alpha((argb >> 8) & 0xff, a, 0) << 8) |
         (mult_by_alpha((argb >> 16) & 0xff, a, 0) << 16);
}
// The inverse on an exported pixel: 32-bit wrap-around products, and quotients above 255 spill into the next channel,
// both exactly as WebPMultARGBRow_C leaves them.
AL_FN uint32_t vp8l_unmultiply(uint32_t argb) {
  if (argb >= 0xff000000u) return argb;
  if (argb <= 0x00ffffffu) return 0;
  const uint32_t scale = (255u << 24) / (argb >> 24);
  uint32_t out = argb & 0xff000000u;
  out |= (((argb & 0xff) * scale + (1u << 23)) >> 24);
  out |= ((((argb >> 8) & 0xff) * scale + (1u << 23)) >> 24) << 8;
  out |= ((((argb >> 16) & 0xff) * scale + (1u << 23)) >> 24) << 16;
  return out;
}
struct Vp8lScaledColumn { Rescaler ch[4]; };
AL_FN void vp8l_column_init(Vp8lScaledColumn& c, const ImgDesc& im, const uint8_t* pm) {
  for (int k = 0; k < 4; ++k) {
    rescaler_init(c.ch[k], pm + k, 4 * im.out_w, im.out_w, im.out_h, im.dst_w, im.dst_h);
    c.ch[k].px_step = 4;
  }
}
AL_FN uint32_t vp8l_column_next(Vp8lScaledColumn& c, int x) {   // the next exported pixel of output column x
  const uint32_t b = (uint32_t)rescaler_next(c.ch[0], x) & 0xff, g = (uint32_t)rescaler_next(c.ch[1], x) & 0xff;
  const uint32_t r = (uint32_t)rescaler_next(c.ch[2], x) & 0xff, a = (uint32_t)rescaler_next(c.ch[3], x) & 0xff;
  return vp8l_unmultiply((a << 24) | (r << 16) | (g << 8) | b);
}
AL_FN void vp8l_uv_of(uint32_t v0, uint32_t v1, int pair, int* u, int* v) {   // WebPConvertARGBToUV_C on one or two pixels
  int r, g, b;
  if (pair) {
    r = (int)(((v0 >> 15) & 0x1fe) + ((v1 >> 15) & 0x1fe)); g = (int)(((v0 >> 7) & 0x1fe) + ((v1 >> 7) & 0x1fe));
    b = (int)(((v0 << 1) & 0x1fe) + ((v1 << 1) & 0x1fe));
  } else {
    r = (int)((v0 >> 14) & 0x3fc); g = (int)((v0 >> 6) & 0x3fc); b = (int)((v0 << 2) & 0x3fc);
  }
  *u = vp8l_clip_uv(-9719 * r - 19081 * g + 28800 * b, 1 << 17);
  *v = vp8l_clip_uv(28800 * r - 24116 * g - 4684 * b, 1 << 17);
}
AL_FN void vp8l_store_pixel(int csp, uint32_t argb, uint8_t* orow, int x) {   // VP8LConvertFromBGRA, one pixel
  const int a = (int)(argb >> 24), r = (int)((argb >> 16) & 0xff), g = (int)((argb >> 8) & 0xff), b = (int)(argb & 0xff);
  if (csp == 0 || csp == 2) {
    uint8_t* o = orow + 3 * x;
    if (csp == 0) { o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b; } else { o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r; }
  } else if (csp == 5 || csp == 6 || csp == 10) {
    const uint32_t p2 = pack_pixel2(csp, r, g, b, a);
    orow[2 * x] = (uint8_t)p2; orow[2 * x + 1] = (uint8_t)(p2 >> 8);
  } else {
    const uint32_t p4 = pack_pixel4(csp, r, g, b, a);
    uint8_t* o = orow + 4 * x;
    o[0] = (uint8_t)p4; o[1] = (uint8_t)(p4 >> 8); o[2] = (uint8_t)(p4 >> 16); o[3] = (uint8_t)(p4 >> 24);
  }
}
AL_FN void vp8l_emit_scaled(const AlphaHdr* hd, const ImgDesc& im, const uint32_t* px, uint8_t* pm, uint8_t* out, int tid, int nt) {
  const int sw = im.out_w, sh = im.out_h, dw = im.dst_w, dh = im.dst_h;
  const int flip = (im.flags & VP8B_FLAG_FLIP) != 0;
  const size_t total = (size_t)sw * (size_t)sh;
  for (size_t i = (size_t)tid; i < total; i += (size_t)nt) {
    const uint32_t p = vp8l_premultiply(vp8l_fetch(hd, px, (int)(i % (size_t)sw) + im.crop_x, (int)(i / (size_t)sw) + im.crop_y));
    pm[4 * i] = (uint8_t)p; pm[4 * i + 1] = (uint8_t)(p >> 8); pm[4 * i + 2] = (uint8_t)(p >> 16); pm[4 * i + 3] = (uint8_t)(p >> 24);
  }
  AL_BLOCK_SYNC();
  if (im.csp != 11 && im.csp != 12) {
    for (int x = tid; x < dw; x += nt) {
      Vp8lScaledColumn c;
      vp8l_column_init(c, im, pm);
      for (int k = 0; k < dh; ++k) vp8l_store_pixel(im.csp, vp8l_column_next(c, x), out + (size_t)(flip ? dh - 1 - k : k) * im.out_stride, x);
    }
    return;
  }
  // MODE_YUV / MODE_YUVA: one work item per chroma column = two luma columns (ConvertToYUVA on every exported row)
  const int uvw = (dw + 1) >> 1, uvh = (dh + 1) >> 1;
  uint8_t* uo = out + (size_t)im.out_stride * dh;
  uint8_t* vo = uo + (size_t)uvw * uvh;
  uint8_t* ao = vo + (size_t)uvw * uvh;
  for (int ux = tid; ux < uvw; ux += nt) {
    const int x0 = 2 * ux, pair = x0 + 1 < dw;
    Vp8lScaledColumn c0, c1;
    vp8l_column_init(c0, im, pm);
    if (pair) vp8l_column_init(c1, im, pm);
    int u_even = 0, v_even = 0;
    for (int k = 0; k < dh; ++k) {
      const uint32_t p0 = vp8l_column_next(c0, x0), p1 = pair ? vp8l_column_next(c1, x0 + 1) : 0;
      const int kd = flip ? dh - 1 - k : k;
      for (int j = 0; j <= pair; ++j) {
        const uint32_t p = j ? p1 : p0;
        const int luma = 16839 * (int)((p >> 16) & 0xff) + 33059 * (int)((p >> 8) & 0xff) + 6420 * (int)(p & 0xff);
        out[(size_t)kd * im.out_stride + x0 + j] = (uint8_t)((luma + (1 << 15) + (16 << 16)) >> 16);
        if (im.csp == 12) ao[(size_t)kd * dw + x0 + j] = (uint8_t)(p >> 24);
      }
      int u, v;
      vp8l_uv_of(p0, p1, pair, &u, &v);
      if (k & 1) { u = (u_even + u + 1) >> 1; v = (v_even + v + 1) >> 1; } else { u_even = u; v_even = v; }
      const int uyd = flip ? uvh - 1 - (k >> 1) : (k >> 1);
      uo[(size_t)uyd * uvw + ux] = (uint8_t)u;
      vo[(size_t)uyd * uvw + ux] = (uint8_t)v;
    }
  }
}

// px = xsize x height coded words (pass B), out = the picture's slot in the output arena; pm = scratch for options.use_scaling.
AL_FN void vp8l_finish_picture(const AlphaHdr* hd, const ImgDesc& im, uint32_t* px, const uint32_t* tdata, uint8_t* out, uint8_t* pm,
                               int tid, int nt) {
  al_inverse_transforms(hd, px, tdata, im.height, tid, nt);
  if (im.dst_w != 0) { vp8l_emit_scaled(hd, im, px, pm, out, tid, nt); return; }
  const int w = im.out_w, h = im.out_h, csp = im.csp;
  if (csp == 11 || csp == 12) { vp8l_emit_yuva(hd, im, px, out, tid, nt); return; }
  const int obpp = (csp == 0 || csp == 2) ? 3 : (csp == 5 || csp == 6 || csp == 10) ? 2 : 4;
  const size_t total = (size_t)w * (size_t)h;
  for (size_t i = (size_t)tid; i < total; i += (size_t)nt) {
    const int x = (int)(i % (size_t)w), y = (int)(i / (size_t)w);
    const int sx = x + im.crop_x, sy = y + im.crop_y;
    const uint32_t argb = vp8l_fetch(hd, px, sx, sy);
    const int a = (int)(argb >> 24), r = (int)((argb >> 16) & 0xff), g = (int)((argb >> 8) & 0xff), b = (int)(argb & 0xff);
    uint8_t* o = out + (size_t)((im.flags & VP8B_FLAG_FLIP) ? h - 1 - y : y) * im.out_stride + (size_t)x * obpp;
    if (obpp == 4) {
      const uint32_t p4 = pack_pixel4(csp, r, g, b, a);
      if ((((uintptr_t)o) & 3) == 0) *(uint32_t*)o = p4;
      else { o[0] = (uint8_t)p4; o[1] = (uint8_t)(p4 >> 8); o[2] = (uint8_t)(p4 >> 16); o[3] = (uint8_t)(p4 >> 24); }
    } else if (obpp == 2) {
      const uint32_t p2 = pack_pixel2(csp, r, g, b, a);
      o[0] = (uint8_t)p2; o[1] = (uint8_t)(p2 >> 8);
    } else if (csp == 0) {
      o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
    } else {
      o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
    }
  }
}

#endif  // LIBWEBP_B200_VP8L_LOSSLESS_CORE_H_
