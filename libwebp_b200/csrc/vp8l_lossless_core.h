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
// them is never looked at. MODE_YUV / MODE_YUVA output and options.use_scaling are refused by the host planner.
#ifndef LIBWEBP_B200_VP8L_LOSSLESS_CORE_H_
#define LIBWEBP_B200_VP8L_LOSSLESS_CORE_H_

#include "vp8_dev.h"
#include "vp8_pixel_core.h"
#include "vp8l_alpha_core.h"

// px = xsize x height coded words (pass B), out = the picture's slot in the output arena.
AL_FN void vp8l_finish_picture(const AlphaHdr* hd, const ImgDesc& im, uint32_t* px, const uint32_t* tdata, uint8_t* out, int tid, int nt) {
  al_inverse_transforms(hd, px, tdata, im.height, tid, nt);
  const int has_palette = hd->ntrans > 0 && hd->ttype[0] == AL_T_COLOR_INDEXING;
  const int bits = has_palette ? hd->tbits[0] : 0, bpp = 8 >> bits;
  const int xs = hd->xsize;
  const int w = im.out_w, h = im.out_h, csp = im.csp;
  const int obpp = (csp == 0 || csp == 2) ? 3 : (csp == 5 || csp == 6 || csp == 10) ? 2 : 4;
  const size_t total = (size_t)w * (size_t)h;
  for (size_t i = (size_t)tid; i < total; i += (size_t)nt) {
    const int x = (int)(i % (size_t)w), y = (int)(i / (size_t)w);
    const int sx = x + im.crop_x, sy = y + im.crop_y;
    uint32_t argb;
    if (has_palette) {   // VP8LColorIndexInverseTransform, lossless.c:341-385: the index travels in green
      const uint32_t packed = (px[(size_t)sy * xs + (sx >> bits)] >> 8) & 0xff;
      argb = hd->palette[(packed >> ((sx & ((1 << bits) - 1)) * bpp)) & ((1u << bpp) - 1u)];
    } else {
      argb = px[(size_t)sy * xs + sx];
    }
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
