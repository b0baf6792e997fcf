/* vp8_container.c -- host side, plain C: RIFF/WebP container walk and the byte-aligned part of the VP8
 * frame header. This is all the host ever reads of a file: the bool-coded headers, modes and tokens are
 * parsed on the device (vp8_parse.cu).
 *
 * Behaviour follows the reference's ParseHeadersInternal (src/dec/webp_dec.c:277-412: ParseRIFF :54,
 * ParseVP8X :93, ParseOptionalChunks :146, ParseVP8Header :222), VP8GetInfo (src/dec/vp8_dec.c:107-147) and
 * the first checks of VP8GetHeaders (src/dec/vp8_dec.c:286-345), including their status codes. */
#include "vp8_container.h"

#include <string.h>

#define TAG(p, s) (memcmp((p), (s), 4) == 0)
#define RIFF_HDR 12u
#define CHUNK_HDR 8u
#define VP8X_PAYLOAD 10u
#define MAX_PAYLOAD (~0u - CHUNK_HDR - 1u)

static uint32_t rd24(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16); }
static uint32_t rd32(const uint8_t* p) { return rd24(p) | ((uint32_t)p[3] << 24); }

/* A cursor over the not-yet-consumed input. */
typedef struct { const uint8_t* p; size_t left; } Cur;
static void skip(Cur* c, size_t n) { c->p += n; c->left -= n; }

int vp8b_parse_container(const uint8_t* data, size_t size, int have_all_data, Vp8Container* out) {
  Cur c;
  size_t riff = 0;
  uint32_t vp8x_flags = 0;
  int canvas_w = 0, canvas_h = 0;
  memset(out, 0, sizeof(*out));
  if (data == NULL || size < RIFF_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
  c.p = data; c.left = size;

  /* "RIFF" <size> "WEBP" -- optional: a bare VP8/VP8L payload is accepted too. */
  if (TAG(c.p, "RIFF")) {
    const uint32_t declared = rd32(c.p + 4);
    if (!TAG(c.p + 8, "WEBP")) return VP8_STATUS_BITSTREAM_ERROR;
    if (declared < 4 + CHUNK_HDR || declared > MAX_PAYLOAD) return VP8_STATUS_BITSTREAM_ERROR;
    if (have_all_data && declared > size - CHUNK_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
    riff = declared;
    skip(&c, RIFF_HDR);
  }

  /* "VP8X": canvas size + feature flags. */
  if (c.left < CHUNK_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
  if (TAG(c.p, "VP8X")) {
    if (rd32(c.p + 4) != VP8X_PAYLOAD) return VP8_STATUS_BITSTREAM_ERROR;
    if (c.left < CHUNK_HDR + VP8X_PAYLOAD) return VP8_STATUS_NOT_ENOUGH_DATA;
    vp8x_flags = rd32(c.p + 8);
    canvas_w = 1 + (int)rd24(c.p + 12);
    canvas_h = 1 + (int)rd24(c.p + 15);
    if ((uint64_t)canvas_w * (uint64_t)canvas_h >= (1ull << 32)) return VP8_STATUS_BITSTREAM_ERROR;
    skip(&c, CHUNK_HDR + VP8X_PAYLOAD);
    out->found_vp8x = 1;
    if (riff == 0) return VP8_STATUS_BITSTREAM_ERROR;   /* VP8X is only legal inside RIFF */
  }
  out->has_alpha = (vp8x_flags >> 4) & 1;
  out->has_animation = (vp8x_flags >> 1) & 1;
  out->width = canvas_w;
  out->height = canvas_h;
  if (out->found_vp8x && out->has_animation && !have_all_data) return VP8_STATUS_OK;   /* features only */

  /* Optional chunks in front of the image data; remember ALPH. */
  if (c.left < 4) return VP8_STATUS_NOT_ENOUGH_DATA;
  if ((riff > 0 && out->found_vp8x) || (riff == 0 && !out->found_vp8x && TAG(c.p, "ALPH"))) {
    uint32_t walked = 4 + CHUNK_HDR + VP8X_PAYLOAD;
    for (;;) {
      uint32_t payload, padded;
      if (c.left < CHUNK_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
      payload = rd32(c.p + 4);
      if (payload > MAX_PAYLOAD) return VP8_STATUS_BITSTREAM_ERROR;
      padded = (CHUNK_HDR + payload + 1) & ~1u;
      walked += padded;
      if (riff > 0 && walked > riff) return VP8_STATUS_BITSTREAM_ERROR;
      if (TAG(c.p, "VP8 ") || TAG(c.p, "VP8L")) break;
      if (c.left < padded) return VP8_STATUS_NOT_ENOUGH_DATA;
      if (TAG(c.p, "ALPH")) { out->alpha_offset = (size_t)(c.p - data) + CHUNK_HDR; out->alpha_size = payload; out->has_alph_chunk = 1; }
      skip(&c, padded);
    }
  }

  /* "VP8 " / "VP8L" chunk header, or a bare bitstream. */
  if (c.left < CHUNK_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
  if (TAG(c.p, "VP8 ") || TAG(c.p, "VP8L")) {
    const uint32_t payload = rd32(c.p + 4);
    out->is_lossless = TAG(c.p, "VP8L");
    if (riff >= 4 + CHUNK_HDR && payload > riff - (4 + CHUNK_HDR)) return VP8_STATUS_BITSTREAM_ERROR;
    if (have_all_data && payload > c.left - CHUNK_HDR) return VP8_STATUS_NOT_ENOUGH_DATA;
    out->chunk_size = payload;
    skip(&c, CHUNK_HDR);
  } else {
    out->is_lossless = (c.left >= 5 && c.p[0] == 0x2f && (c.p[4] >> 5) == 0);
    out->chunk_size = c.left;
  }
  if (out->chunk_size > MAX_PAYLOAD) return VP8_STATUS_BITSTREAM_ERROR;
  if (!out->has_animation) out->format = out->is_lossless ? 2 : 1;

  if (!out->is_lossless) {
    /* 3-byte frame tag, start code 9d 01 2a, 14-bit width and height (RFC 6386 9.1) */
    uint32_t tag;
    int w, h;
    if (c.left < 10) return VP8_STATUS_NOT_ENOUGH_DATA;
    if (!(c.p[3] == 0x9d && c.p[4] == 0x01 && c.p[5] == 0x2a)) return VP8_STATUS_BITSTREAM_ERROR;
    tag = rd24(c.p);
    w = (int)(rd24(c.p + 6) & 0x3fff);
    h = (int)((((uint32_t)c.p[9] << 8) | c.p[8]) & 0x3fff);
    if ((tag & 1) != 0) return VP8_STATUS_BITSTREAM_ERROR;          /* not a key frame */
    if (((tag >> 1) & 7) > 3) return VP8_STATUS_BITSTREAM_ERROR;    /* unknown profile */
    if (((tag >> 4) & 1) == 0) return VP8_STATUS_BITSTREAM_ERROR;   /* invisible frame */
    if ((tag >> 5) >= out->chunk_size) return VP8_STATUS_BITSTREAM_ERROR;
    if (w == 0 || h == 0) return VP8_STATUS_BITSTREAM_ERROR;
    out->part0_size = tag >> 5;
    out->width = w; out->height = h;
  } else {
    uint32_t bits;
    if (c.left < 5) return VP8_STATUS_NOT_ENOUGH_DATA;
    if (c.p[0] != 0x2f || (c.p[4] >> 5) != 0) return VP8_STATUS_BITSTREAM_ERROR;
    bits = rd32(c.p + 1);
    out->width = (int)(bits & 0x3fff) + 1;
    out->height = (int)((bits >> 14) & 0x3fff) + 1;
    out->has_alpha |= (int)((bits >> 28) & 1);
  }
  if (out->found_vp8x && (canvas_w != out->width || canvas_h != out->height)) return VP8_STATUS_BITSTREAM_ERROR;
  out->has_alpha |= out->has_alph_chunk;
  out->frame_offset = (size_t)(c.p - data);
  out->frame_size = c.left;
  out->complete = 1;
  return VP8_STATUS_OK;
}

VP8StatusCode vp8b_get_features(const uint8_t* data, size_t size, WebPBitstreamFeatures* f) {
  Vp8Container c;
  const int st = vp8b_parse_container(data, size, 0, &c);
  memset(f, 0, sizeof(*f));
  if (st == VP8_STATUS_OK || (st == VP8_STATUS_NOT_ENOUGH_DATA && c.found_vp8x)) {
    f->width = c.width; f->height = c.height;
    /* an ALPH chunk seen before the data ran out counts too (webp_dec.c:397-404) */
    f->has_alpha = c.has_alpha | c.has_alph_chunk; f->has_animation = c.has_animation; f->format = c.format;
    return VP8_STATUS_OK;
  }
  return (VP8StatusCode)st;
}

/* ---------------------------------------------------------------------------------------------------------
 * Partition-count pre-scan. The partition count sits behind the bool-coded segment and filter headers; the
 * host needs it only to size the token-parse thread blocks (one warp per partition), so it walks those few
 * dozen symbols with a minimal byte-wise boolean decoder. The device parses the same header again and its
 * result is authoritative (vp8_parse_core.h). */
typedef struct { const uint8_t* p; const uint8_t* end; uint32_t value, range; int shifted; } MiniBool;

static uint32_t mb_next(MiniBool* d) { return (d->p < d->end) ? *d->p++ : 0u; }

static int mb_get(MiniBool* d, int prob) {   /* RFC 6386 section 7.3 */
  const uint32_t split = 1 + (((d->range - 1) * (uint32_t)prob) >> 8);
  int bit = 0;
  if (d->value >= (split << 8)) { d->range -= split; d->value -= split << 8; bit = 1; } else { d->range = split; }
  while (d->range < 128) {
    d->value <<= 1; d->range <<= 1;
    if (++d->shifted == 8) { d->shifted = 0; d->value |= mb_next(d); }
  }
  return bit;
}
static uint32_t mb_lit(MiniBool* d, int n) { uint32_t v = 0; while (n-- > 0) v = (v << 1) | (uint32_t)mb_get(d, 128); return v; }
static void mb_skip_flagged(MiniBool* d, int n) { if (mb_get(d, 128)) { mb_lit(d, n); } }

int vp8b_prescan_partitions(const uint8_t* part0, size_t part0_size) {
  MiniBool d;
  int i;
  d.p = part0; d.end = part0 + part0_size; d.range = 255; d.shifted = 0;
  d.value = mb_next(&d) << 8;
  d.value |= mb_next(&d);
  mb_lit(&d, 2);                                  /* colorspace, clamp type */
  if (mb_get(&d, 128)) {                          /* segmentation enabled */
    const int update_map = mb_get(&d, 128);
    if (mb_get(&d, 128)) {                        /* update segment feature data */
      mb_get(&d, 128);                            /* absolute / delta */
      for (i = 0; i < 4; ++i) mb_skip_flagged(&d, 7 + 1);
      for (i = 0; i < 4; ++i) mb_skip_flagged(&d, 6 + 1);
    }
    if (update_map) for (i = 0; i < 3; ++i) mb_skip_flagged(&d, 8);
  }
  mb_lit(&d, 1 + 6 + 3);                          /* filter type, level, sharpness */
  if (mb_get(&d, 128)) {                          /* loop-filter deltas enabled */
    if (mb_get(&d, 128)) {
      for (i = 0; i < 8; ++i) mb_skip_flagged(&d, 6 + 1);
    }
  }
  return 1 << mb_lit(&d, 2);
}

/* ParsePartitions (vp8_dec.c:188-222): the size table (three bytes per partition but the last) follows the first partition,
   sizes are clipped to what is left. Byte 0xFF at the head of a partition is what no encoder writes and what makes the
   reference's reader leave its range (vp8_parse_core.h:RefBits). `rest` = bytes of the frame after its 10-byte header. */
int vp8b_partition_starts_with_ff(const uint8_t* part0, size_t part0_size, size_t rest, int num_parts) {
  const uint8_t* sz;
  const uint8_t* start;
  size_t left;
  int p;
  if (part0_size > 0 && part0_size <= rest && part0[0] == 0xFF) return 1;
  if (part0_size > rest) return 0;
  sz = part0 + part0_size;
  left = rest - part0_size;
  if (left < 3u * (size_t)(num_parts - 1)) return 0;
  start = sz + 3 * (num_parts - 1);
  left -= 3u * (size_t)(num_parts - 1);
  for (p = 0; p < num_parts; ++p) {
    size_t psize = left;
    if (p < num_parts - 1) {
      psize = (size_t)sz[0] | ((size_t)sz[1] << 8) | ((size_t)sz[2] << 16);
      if (psize > left) psize = left;
      sz += 3;
    }
    if (psize > 0 && start[0] == 0xFF) return 1;
    start += psize;
    left -= psize;
  }
  return 0;
}

/* Has the LAST token partition begun inside the bytes at hand? That is when VP8GetHeaders stops answering "suspended" for a
   stream that is still arriving (ParsePartitions, vp8_dec.c:188-222: `part_start < buf_end`), i.e. when the reference's incremental
   decoder gets its output area. `rest` = bytes at hand after the frame's 10-byte header. */
int vp8b_last_partition_begun(const uint8_t* part0, size_t part0_size, size_t rest, int num_parts) {
  const uint8_t* sz;
  size_t left, start;
  int p;
  if (part0_size > rest) return 0;
  sz = part0 + part0_size;
  left = rest - part0_size;
  if (left < 3u * (size_t)(num_parts - 1)) return 0;
  left -= 3u * (size_t)(num_parts - 1);
  start = 0;
  for (p = 0; p < num_parts - 1; ++p) {
    size_t psize = (size_t)sz[0] | ((size_t)sz[1] << 8) | ((size_t)sz[2] << 16);
    if (psize > left) psize = left;
    start += psize; left -= psize;
    sz += 3;
  }
  (void)start;
  return left > 0;
}
