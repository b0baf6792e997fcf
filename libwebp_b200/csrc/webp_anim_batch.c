/* webp_anim_batch.c -- every frame of an animated WebP file decoded as ONE batch (include/webp/decode_batch.h:
 * WebPAnimBatchGetInfo, WebPAnimDecodeBatch). Host C; the pixels come from WebPDecodeBatch().
 *
 * The reference reconstructs an animation frame by frame: WebPAnimDecoderGetNext (src/demux/anim_decode.c:325-440) takes
 * frame k from the demuxer, calls WebPDecode on its payload straight into the frame's rectangle of the canvas, then
 * blends the pixels the frame left (partly) transparent against the previous canvas and disposes. Each WebPDecode there
 * is a GPU round trip here. The frames of a file are independent as bitstreams (each ANMF payload is a complete
 * [ALPH] VP8 / VP8L image), so this entry point decodes them all in one WebPDecodeBatch into frame-sized buffers and
 * then replays the reference's canvas logic -- key-frame rule, blending (IsKeyFrame :197, BlendPixelNonPremult :231,
 * BlendPixelPremult :283, FindBlendRangeAtRow :302), disposal -- on the host, in frame order. The container walk follows
 * src/demux/demux.c (ParseVP8X / ParseAnimationFrame / StoreFrame: :200-560) for well-formed files; anything it does not
 * recognise is refused (VP8_STATUS_BITSTREAM_ERROR), never guessed at. */
#include <stdlib.h>
#include <string.h>

#include "vp8_container.h"
#include "webp/decode_batch.h"

static uint32_t rd24(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
static uint32_t rd32(const uint8_t* p) { return rd24(p) | ((uint32_t)p[3] << 24); }

typedef struct AnimFrame {
  const uint8_t* payload;   /* [ALPH chunk] + VP8 / VP8L chunk, as WebPDecode takes them */
  size_t size;
  int x, y, w, h, duration, dispose_bg, no_blend, has_alpha;
} AnimFrame;

/* Walks RIFF/VP8X/ANIM/ANMF. frames may be NULL (count only). Returns the frame count, or -1. */
static int walk(const uint8_t* data, size_t size, WebPAnimBatchInfo* info, AnimFrame* frames, int max_frames) {
  size_t pos = 12, end;
  int n = 0, seen_anim = 0;
  uint32_t flags;
  if (data == NULL || size < 30 || memcmp(data, "RIFF", 4) != 0 || memcmp(data + 8, "WEBP", 4) != 0) return -1;
  end = (size_t)rd32(data + 4) + 8;
  if (end > size) return -1;             /* the whole file has to be there */
  if (memcmp(data + 12, "VP8X", 4) != 0 || rd32(data + 16) != 10) return -1;
  flags = data[20];
  if (!(flags & 0x02)) return -1;        /* no animation flag: not an animated file */
  info->canvas_width = 1 + (int)rd24(data + 24);
  info->canvas_height = 1 + (int)rd24(data + 27);
  info->bgcolor = 0; info->loop_count = 0;
  pos = 30;
  while (pos + 8 <= end) {
    const uint8_t* ck = data + pos;
    const size_t csz = rd32(ck + 4), padded = csz + (csz & 1);
    if (csz > end - pos - 8) return -1;
    if (memcmp(ck, "ANIM", 4) == 0) {
      if (csz < 6) return -1;
      /* stored blue, green, red, alpha (container spec) = the uint32 the demuxer hands out (demux.c:ParseVP8XChunks) */
      info->bgcolor = rd32(ck + 8);
      info->loop_count = (int)(ck[12] | (ck[13] << 8));
      seen_anim = 1;
    } else if (memcmp(ck, "ANMF", 4) == 0) {
      AnimFrame f;
      size_t sub = 16, img_end = 0;
      int found_image = 0;
      if (!seen_anim || csz < 16 + 8) return -1;
      memset(&f, 0, sizeof(f));
      f.x = 2 * (int)rd24(ck + 8); f.y = 2 * (int)rd24(ck + 11);
      f.w = 1 + (int)rd24(ck + 14); f.h = 1 + (int)rd24(ck + 17);
      f.duration = (int)rd24(ck + 20);
      f.dispose_bg = ck[23] & 1; f.no_blend = (ck[23] >> 1) & 1;
      if (f.x + f.w > info->canvas_width || f.y + f.h > info->canvas_height) return -1;
      /* sub-chunks: an optional ALPH, then the image; whatever follows the image inside the frame is ignored (StoreFrame) */
      f.payload = ck + 8 + 16;
      while (sub + 8 <= csz && !found_image) {
        const uint8_t* s = ck + 8 + sub;
        const size_t ssz = rd32(s + 4), spad = ssz + (ssz & 1);
        if (ssz > csz - sub - 8) return -1;
        if (memcmp(s, "VP8 ", 4) == 0 || memcmp(s, "VP8L", 4) == 0) { found_image = 1; img_end = sub + 8 + ssz; }
        else if (memcmp(s, "ALPH", 4) != 0) return -1;   /* anything else ends the demuxer's scan without an image (StoreFrame) */
        sub += 8 + spad;
      }
      if (!found_image) return -1;
      f.size = (size_t)(ck + 8 + img_end - f.payload);
      {
        WebPBitstreamFeatures ft;
        if (vp8b_get_features(f.payload, f.size, &ft) != VP8_STATUS_OK) return -1;
        if (ft.width != f.w || ft.height != f.h) return -1;     /* demux.c:CheckFrameBounds / StoreFrame */
        f.has_alpha = ft.has_alpha;
      }
      if (frames != NULL && n < max_frames) frames[n] = f;
      ++n;
    }
    pos += 8 + padded;
  }
  if (n == 0) return -1;
  info->frame_count = n;
  return n;
}

int WebPAnimBatchGetInfo(const uint8_t* data, size_t data_size, WebPAnimBatchInfo* info) {
  WebPAnimBatchInfo tmp;
  if (info == NULL) return 0;
  if (walk(data, data_size, &tmp, NULL, 0) < 0) return 0;
  *info = tmp;
  return 1;
}

/* ---- the reference's blending, channel positions as bytes in memory (anim_decode.c:214-300; little-endian host) */
static uint8_t blend_channel(uint32_t src, uint8_t src_a, uint32_t dst, uint8_t dst_a, uint32_t scale, int shift) {
  const uint8_t s = (uint8_t)((src >> shift) & 0xff), d = (uint8_t)((dst >> shift) & 0xff);
  const uint32_t unscaled = (uint32_t)s * src_a + (uint32_t)d * dst_a;
  return (uint8_t)((unscaled * scale) >> 24);
}

static uint32_t blend_nonpremult(uint32_t src, uint32_t dst) {
  const uint8_t src_a = (uint8_t)(src >> 24);
  if (src_a == 0) return dst;
  {
    const uint8_t dst_a = (uint8_t)(dst >> 24);
    const uint8_t dst_factor_a = (uint8_t)((dst_a * (256 - src_a)) >> 8);
    const uint8_t blend_a = (uint8_t)(src_a + dst_factor_a);
    const uint32_t scale = (1UL << 24) / blend_a;
    const uint8_t c0 = blend_channel(src, src_a, dst, dst_factor_a, scale, 0);
    const uint8_t c1 = blend_channel(src, src_a, dst, dst_factor_a, scale, 8);
    const uint8_t c2 = blend_channel(src, src_a, dst, dst_factor_a, scale, 16);
    return (uint32_t)c0 | ((uint32_t)c1 << 8) | ((uint32_t)c2 << 16) | ((uint32_t)blend_a << 24);
  }
}

static uint32_t channelwise_multiply(uint32_t pix, uint32_t scale) {
  const uint32_t mask = 0x00FF00FF;
  const uint32_t rb = ((pix & mask) * scale) >> 8;
  const uint32_t ag = ((pix >> 8) & mask) * scale;
  return (rb & mask) | (ag & ~mask);
}

static void blend_row(uint32_t* src, const uint32_t* dst, int n, int premult) {
  int i;
  for (i = 0; i < n; ++i) {
    const uint8_t a = (uint8_t)(src[i] >> 24);
    if (a != 0xff) src[i] = premult ? src[i] + channelwise_multiply(dst[i], 256 - a) : blend_nonpremult(src[i], dst[i]);
  }
}

VP8StatusCode WebPAnimDecodeBatch(const uint8_t* data, size_t data_size, WEBP_CSP_MODE mode, uint8_t* canvases,
                                  size_t canvases_size, int* timestamps, const WebPBatchOptions* options) {
  WebPAnimBatchInfo info;
  AnimFrame* fr = NULL;
  WebPBatchItem* items = NULL;
  WebPDecoderConfig* cfgs = NULL;
  uint8_t* pixels = NULL;
  uint8_t* disposed = NULL;     /* the previous canvas after its disposal (dec->prev_frame_disposed_) */
  size_t* offs = NULL;
  VP8StatusCode st = VP8_STATUS_OUT_OF_MEMORY;
  int n, k, premult, prev_key = 0;
  size_t total = 0, canvas_bytes;
  /* WebPAnimDecoderNewInternal, anim_decode.c:96-104: the four modes a canvas can have */
  if (mode != MODE_RGBA && mode != MODE_BGRA && mode != MODE_rgbA && mode != MODE_bgrA) return VP8_STATUS_INVALID_PARAM;
  premult = (mode == MODE_rgbA || mode == MODE_bgrA);
  n = walk(data, data_size, &info, NULL, 0);
  if (n < 0) return VP8_STATUS_BITSTREAM_ERROR;
  canvas_bytes = (size_t)4 * info.canvas_width * info.canvas_height;
  if (canvases == NULL || canvases_size / canvas_bytes < (size_t)n) return VP8_STATUS_INVALID_PARAM;
  fr = (AnimFrame*)calloc((size_t)n, sizeof(*fr));
  items = (WebPBatchItem*)calloc((size_t)n, sizeof(*items));
  cfgs = (WebPDecoderConfig*)calloc((size_t)n, sizeof(*cfgs));
  offs = (size_t*)calloc((size_t)n + 1, sizeof(*offs));
  disposed = (uint8_t*)calloc(1, canvas_bytes);
  if (fr == NULL || items == NULL || cfgs == NULL || offs == NULL || disposed == NULL) goto End;
  if (walk(data, data_size, &info, fr, n) != n) { st = VP8_STATUS_BITSTREAM_ERROR; goto End; }
  for (k = 0; k < n; ++k) { offs[k] = total; total += (size_t)4 * fr[k].w * fr[k].h; }
  pixels = (uint8_t*)WebPBatchHostAlloc(total ? total : 4);    /* page-locked: one download */
  if (pixels == NULL) goto End;
  /* ---- every frame an item of one batch */
  for (k = 0; k < n; ++k) {
    if (!WebPInitDecoderConfig(&cfgs[k])) { st = VP8_STATUS_INVALID_PARAM; goto End; }
    cfgs[k].output.colorspace = mode;
    cfgs[k].output.is_external_memory = 1;
    cfgs[k].output.u.RGBA.rgba = pixels + offs[k];
    cfgs[k].output.u.RGBA.stride = 4 * fr[k].w;
    cfgs[k].output.u.RGBA.size = (size_t)4 * fr[k].w * fr[k].h;
    items[k].data = fr[k].payload; items[k].data_size = fr[k].size; items[k].config = &cfgs[k];
  }
  st = WebPDecodeBatch(items, n, options);
  if (st != VP8_STATUS_OK) goto End;      /* the reference stops at the first frame that fails (anim_decode.c:376) */
  /* ---- canvases, in order */
  for (k = 0; k < n; ++k) {
    const AnimFrame* f = &fr[k];
    const AnimFrame* p = k > 0 ? &fr[k - 1] : NULL;
    uint8_t* canvas = canvases + (size_t)k * canvas_bytes;
    const int cw = info.canvas_width, ch = info.canvas_height;
    int is_key, y;
    if (k == 0) is_key = 1;
    else if ((!f->has_alpha || f->no_blend) && f->w == cw && f->h == ch) is_key = 1;
    else is_key = p->dispose_bg && ((p->w == cw && p->h == ch) || prev_key);
    if (is_key) memset(canvas, 0, canvas_bytes); else memcpy(canvas, disposed, canvas_bytes);
    for (y = 0; y < f->h; ++y) {
      memcpy(canvas + ((size_t)(f->y + y) * cw + f->x) * 4, pixels + offs[k] + (size_t)y * f->w * 4, (size_t)f->w * 4);
    }
    if (k > 0 && !f->no_blend && !is_key) {
      if (!p->dispose_bg) {
        for (y = 0; y < f->h; ++y) {
          const size_t o = (size_t)(f->y + y) * cw + f->x;
          blend_row((uint32_t*)canvas + o, (const uint32_t*)disposed + o, f->w, premult);
        }
      } else {
        for (y = 0; y < f->h; ++y) {   /* FindBlendRangeAtRow: the parts of this frame's row outside the previous rectangle */
          const int cy = f->y + y, src_max_x = f->x + f->w, dst_max_x = p->x + p->w, dst_max_y = p->y + p->h;
          int left1 = -1, width1 = 0, left2 = -1, width2 = 0;
          if (cy < p->y || cy >= dst_max_y || f->x >= dst_max_x || src_max_x <= p->x) { left1 = f->x; width1 = f->w; }
          else {
            if (f->x < p->x) { left1 = f->x; width1 = p->x - f->x; }
            if (src_max_x > dst_max_x) { left2 = dst_max_x; width2 = src_max_x - dst_max_x; }
          }
          if (width1 > 0) { const size_t o = (size_t)cy * cw + left1; blend_row((uint32_t*)canvas + o, (const uint32_t*)disposed + o, width1, premult); }
          if (width2 > 0) { const size_t o = (size_t)cy * cw + left2; blend_row((uint32_t*)canvas + o, (const uint32_t*)disposed + o, width2, premult); }
        }
      }
    }
    if (timestamps != NULL) timestamps[k] = (k > 0 ? timestamps[k - 1] : 0) + f->duration;
    prev_key = is_key;
    memcpy(disposed, canvas, canvas_bytes);
    if (f->dispose_bg) {
      for (y = 0; y < f->h; ++y) memset(disposed + ((size_t)(f->y + y) * cw + f->x) * 4, 0, (size_t)f->w * 4);
    }
  }
  st = VP8_STATUS_OK;
End:
  WebPBatchHostFree(pixels);
  free(fr); free(items); free(cfgs); free(offs); free(disposed);
  return st;
}
