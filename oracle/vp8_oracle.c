// vp8_oracle.c -- TEST INFRASTRUCTURE ONLY. See vp8_oracle.h for scope, citations and how it is pinned.
// Whole-frame, sequential, no SIMD, no threads: parse everything, reconstruct in raster order, filter in
// raster order, convert. Written for readability; every stage keeps its intermediate so vp8o_dump can
// hand it to the kernel tests.
#include "vp8_oracle.h"

#include <stdlib.h>
#include <string.h>

#include "vp8_tables.h"

// =====================================================================================================
// Boolean decoder. State kept as in the reference: range stored minus one, in [127,254] once normalised
// (src/utils/bit_reader_utils.h:93-103), a 64-bit value filled seven bytes at a time while eight or more are left and one
// byte at a time after that (BITS = 56: bit_reader_inl_utils.h:58-103, bit_reader_utils.c:88-101); `eof` is raised exactly
// when a decode needs a byte that does not exist. On every stream an encoder writes the refill width does not show in the
// decoded bits; it does when a partition starts with byte 0xFF, which puts the value outside its range from the first bit on
// (whatever has overflowed above the low eight bits is dropped when the next seven bytes come in), so the width is kept.
typedef struct {
  const uint8_t* p;
  const uint8_t* end;
  const uint8_t* wide_end;   // seven bytes at once while p is below this
  uint32_t range;
  uint64_t value;
  int bits;   // number of not-yet-consumed bits in `value`, minus 8
  int eof;
  uint64_t decodes;   // boolean decodes so far (vp8o_count_decodes: the length of this stream's dependent chain)
} BoolDec;

static void bd_refill(BoolDec* d) {
  if (d->p < d->wide_end) {
    uint64_t w = 0;
    int k;
    for (k = 0; k < 7; ++k) w = (w << 8) | d->p[k];
    d->p += 7;
    d->value = (d->value << 56) | w;
    d->bits += 56;
  } else if (d->p < d->end) {
    d->value = (d->value << 8) | *d->p++;
    d->bits += 8;
  } else if (!d->eof) {
    d->value <<= 8;
    d->bits += 8;
    d->eof = 1;
  } else {
    d->bits = 0;
  }
}

static void bd_init(BoolDec* d, const uint8_t* start, size_t size) {
  d->p = start;
  d->end = start + size;
  d->wide_end = size >= 8 ? start + size - 7 : start;
  d->range = 255 - 1;
  d->value = 0;
  d->bits = -8;
  d->eof = 0;
  d->decodes = 0;
  bd_refill(d);
}

static int ilog2(uint32_t v) {   // floor(log2(v)), v >= 1
  int n = 0;
  while (v >>= 1) ++n;
  return n;
}

static int bd_bit(BoolDec* d, int prob) {   // bit_reader_inl_utils.h:107-136
  uint32_t range = d->range;
  uint32_t split, top;
  int bit, shift;
  if (d->bits < 0) bd_refill(d);
  ++d->decodes;
  split = (range * (uint32_t)prob) >> 8;
  top = (uint32_t)(d->value >> d->bits);
  if (top > split) {
    range -= split;
    d->value -= (uint64_t)(split + 1) << d->bits;
    bit = 1;
  } else {
    range = split + 1;
    bit = 0;
  }
  shift = 7 ^ ilog2(range);
  range <<= shift;
  d->bits -= shift;
  d->range = range - 1;
  return bit;
}

// Sign of a coefficient (VP8GetSigned, bit_reader_inl_utils.h:139-158): an even split decided by the sign bit of a 32-bit
// difference; the same as bd_bit(d, 0x80) while the value is inside its range. Returns 1 for negative.
static int bd_sign(BoolDec* d) {
  uint32_t split, top, mask;
  if (d->bits < 0) bd_refill(d);
  ++d->decodes;
  split = d->range >> 1;
  top = (uint32_t)(d->value >> d->bits);
  mask = (uint32_t)((int32_t)(split - top) >> 31);
  d->value -= (uint64_t)((split + 1) & mask) << d->bits;
  d->bits -= 1;
  d->range = (d->range + mask) | 1u;
  return (int)(mask & 1u);
}

static uint32_t bd_value(BoolDec* d, int nbits) {   // bit_reader_utils.c:106
  uint32_t v = 0;
  while (nbits-- > 0) v |= (uint32_t)bd_bit(d, 0x80) << nbits;
  return v;
}

static int bd_signed_value(BoolDec* d, int nbits) {   // bit_reader_utils.c:114
  const int v = (int)bd_value(d, nbits);
  return bd_bit(d, 0x80) ? -v : v;
}

// =====================================================================================================
// Container walk (src/dec/webp_dec.c:54-412).
typedef struct {
  const uint8_t* vp8;       // start of the VP8 frame (frame tag)
  size_t vp8_size;          // bytes available from there to the end of the input
  size_t chunk_size;        // declared payload size
  const uint8_t* alpha;
  size_t alpha_size;
  int is_lossless, has_alpha, has_animation, width, height, found_vp8x, format;
} Container;

static uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t le24(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16); }

#define MAX_CHUNK_PAYLOAD (~0U - 8 - 1)

// `all_data` = the caller promises the whole file is there (decode) vs a feature probe.
static int parse_container(const uint8_t* data, size_t size, int all_data, Container* c) {
  size_t riff_size = 0;
  int found_riff, canvas_w = 0, canvas_h = 0;
  uint32_t flags = 0;
  memset(c, 0, sizeof(*c));
  if (data == NULL || size < 12) return VP8O_NOT_ENOUGH_DATA;
  if (!memcmp(data, "RIFF", 4)) {
    uint32_t sz;
    if (memcmp(data + 8, "WEBP", 4)) return VP8O_BITSTREAM_ERROR;
    sz = le32(data + 4);
    if (sz < 12 || sz > MAX_CHUNK_PAYLOAD) return VP8O_BITSTREAM_ERROR;
    if (all_data && sz > size - 8) return VP8O_NOT_ENOUGH_DATA;
    riff_size = sz;
    data += 12; size -= 12;
  }
  found_riff = riff_size > 0;
  if (size < 8) return VP8O_NOT_ENOUGH_DATA;
  if (!memcmp(data, "VP8X", 4)) {
    if (le32(data + 4) != 10) return VP8O_BITSTREAM_ERROR;
    if (size < 18) return VP8O_NOT_ENOUGH_DATA;
    flags = le32(data + 8);
    canvas_w = 1 + (int)le24(data + 12);
    canvas_h = 1 + (int)le24(data + 15);
    if ((uint64_t)canvas_w * (uint64_t)canvas_h >= (1ull << 32)) return VP8O_BITSTREAM_ERROR;
    data += 18; size -= 18;
    c->found_vp8x = 1;
  }
  if (!found_riff && c->found_vp8x) return VP8O_BITSTREAM_ERROR;
  c->has_alpha = !!(flags & 0x10);
  c->has_animation = !!(flags & 0x02);
  c->width = canvas_w; c->height = canvas_h;
  if (c->found_vp8x && c->has_animation && !all_data) return VP8O_OK;   // features come from VP8X alone
  if (size < 4) return VP8O_NOT_ENOUGH_DATA;
  if ((found_riff && c->found_vp8x) || (!found_riff && !c->found_vp8x && !memcmp(data, "ALPH", 4))) {
    uint32_t total = 4 + 8 + 10;
    for (;;) {
      uint32_t csz, disk;
      if (size < 8) return VP8O_NOT_ENOUGH_DATA;
      csz = le32(data + 4);
      if (csz > MAX_CHUNK_PAYLOAD) return VP8O_BITSTREAM_ERROR;
      disk = (8 + csz + 1) & ~1u;
      total += disk;
      if (riff_size > 0 && total > riff_size) return VP8O_BITSTREAM_ERROR;
      if (!memcmp(data, "VP8 ", 4) || !memcmp(data, "VP8L", 4)) break;
      if (size < disk) return VP8O_NOT_ENOUGH_DATA;
      if (!memcmp(data, "ALPH", 4)) { c->alpha = data + 8; c->alpha_size = csz; }
      data += disk; size -= disk;
    }
  }
  if (size < 8) return VP8O_NOT_ENOUGH_DATA;
  {
    const int is_vp8 = !memcmp(data, "VP8 ", 4), is_vp8l = !memcmp(data, "VP8L", 4);
    if (is_vp8 || is_vp8l) {
      const uint32_t sz = le32(data + 4);
      if (riff_size >= 12 && sz > riff_size - 12) return VP8O_BITSTREAM_ERROR;
      if (all_data && sz > size - 8) return VP8O_NOT_ENOUGH_DATA;
      c->chunk_size = sz;
      data += 8; size -= 8;
      c->is_lossless = is_vp8l;
    } else {
      c->is_lossless = (size >= 5 && data[0] == 0x2f && (data[4] >> 5) == 0);   // VP8LCheckSignature
      c->chunk_size = size;
    }
  }
  if (c->chunk_size > MAX_CHUNK_PAYLOAD) return VP8O_BITSTREAM_ERROR;
  if (!c->has_animation) c->format = c->is_lossless ? 2 : 1;
  if (!c->is_lossless) {
    uint32_t bits;
    int w, h;
    if (size < 10) return VP8O_NOT_ENOUGH_DATA;
    // VP8GetInfo, src/dec/vp8_dec.c:107-147
    if (!(data[3] == 0x9d && data[4] == 0x01 && data[5] == 0x2a)) return VP8O_BITSTREAM_ERROR;
    bits = le24(data);
    w = ((data[7] << 8) | data[6]) & 0x3fff;
    h = ((data[9] << 8) | data[8]) & 0x3fff;
    if ((bits & 1) || ((bits >> 1) & 7) > 3 || !((bits >> 4) & 1) || (bits >> 5) >= c->chunk_size || w == 0 || h == 0) {
      return VP8O_BITSTREAM_ERROR;
    }
    if (c->found_vp8x && (canvas_w != w || canvas_h != h)) return VP8O_BITSTREAM_ERROR;
    c->width = w; c->height = h;
  } else {
    int w, h;
    if (size < 5) return VP8O_NOT_ENOUGH_DATA;
    if (data[0] != 0x2f || (data[4] >> 5) != 0) return VP8O_BITSTREAM_ERROR;
    {
      const uint32_t b = le32(data + 1);
      w = (int)(b & 0x3fff) + 1; h = (int)((b >> 14) & 0x3fff) + 1;
      c->has_alpha |= (int)((b >> 28) & 1);
    }
    if (c->found_vp8x && (canvas_w != w || canvas_h != h)) return VP8O_BITSTREAM_ERROR;
    c->width = w; c->height = h;
  }
  c->has_alpha |= (c->alpha != NULL);
  c->vp8 = data;
  c->vp8_size = size;
  return VP8O_OK;
}

int vp8o_features(const uint8_t* data, size_t size, int* f) {
  Container c;
  const int st = parse_container(data, size, 0, &c);
  f[0] = f[1] = f[2] = f[3] = f[4] = 0;
  // ParseHeadersInternal: OK, or NOT_ENOUGH_DATA once a VP8X header was seen, still reports the features.
  if (st == VP8O_OK || (st == VP8O_NOT_ENOUGH_DATA && c.found_vp8x)) {
    f[0] = c.width; f[1] = c.height; f[2] = c.has_alpha; f[3] = c.has_animation;
    f[4] = c.format;
    return VP8O_OK;
  }
  return st;
}

// =====================================================================================================
// Frame state.
typedef struct {
  uint8_t imodes[16];
  uint8_t is_i4x4, uvmode, skip, segment;
} MbModes;

typedef struct { uint8_t limit, ilevel, inner, hev; } FInfo;

typedef struct {
  int width, height, mb_w, mb_h;
  int use_segment, update_map, absolute_delta;
  int seg_quant[4], seg_filter[4];
  uint8_t seg_prob[3];
  int simple, level, sharpness, use_lf_delta, ref_lf_delta[4], mode_lf_delta[4];
  int filter_type;
  int num_parts;
  BoolDec br;            // partition 0
  BoolDec parts[8];
  int dq[4][6];
  uint8_t prob[4][8][3][11];
  int use_skip, skip_p;
  FInfo fstr[4][2];
  MbModes* modes;
  int16_t* coeffs;
  uint32_t* nz;          // 2 per MB
  FInfo* finfo;
  int ys, uvs;           // plane strides
  uint8_t *y, *u, *v;    // reconstruction (filtered in place)
  uint8_t* unfiltered;   // optional copy taken before filtering (dump only)
} Frame;

static int clipi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

// VP8GetHeaders (src/dec/vp8_dec.c:263-395) on the VP8 payload.
static int parse_frame_header(Frame* f, const uint8_t* buf, size_t size) {
  uint32_t tag, part0;
  BoolDec* br = &f->br;
  int s, i, t, b, c, p;
  if (size < 4) return VP8O_NOT_ENOUGH_DATA;
  tag = le24(buf);
  if (((tag >> 1) & 7) > 3) return VP8O_BITSTREAM_ERROR;
  if (!((tag >> 4) & 1)) return VP8O_UNSUPPORTED_FEATURE;
  part0 = tag >> 5;
  buf += 3; size -= 3;
  if (tag & 1) return VP8O_UNSUPPORTED_FEATURE;   // not a key frame (parse_container already rejects it)
  if (size < 7) return VP8O_NOT_ENOUGH_DATA;
  if (!(buf[0] == 0x9d && buf[1] == 0x01 && buf[2] == 0x2a)) return VP8O_BITSTREAM_ERROR;
  f->width = ((buf[4] << 8) | buf[3]) & 0x3fff;
  f->height = ((buf[6] << 8) | buf[5]) & 0x3fff;
  buf += 7; size -= 7;
  f->mb_w = (f->width + 15) >> 4;
  f->mb_h = (f->height + 15) >> 4;
  if (part0 > size) return VP8O_NOT_ENOUGH_DATA;
  bd_init(br, buf, part0);
  buf += part0; size -= part0;
  bd_value(br, 1);   // colorspace
  bd_value(br, 1);   // clamp type
  // segment header (vp8_dec.c:162-196)
  f->absolute_delta = 1;
  f->seg_prob[0] = f->seg_prob[1] = f->seg_prob[2] = 255;
  f->use_segment = (int)bd_value(br, 1);
  if (f->use_segment) {
    f->update_map = (int)bd_value(br, 1);
    if (bd_value(br, 1)) {
      f->absolute_delta = (int)bd_value(br, 1);
      for (s = 0; s < 4; ++s) f->seg_quant[s] = bd_value(br, 1) ? bd_signed_value(br, 7) : 0;
      for (s = 0; s < 4; ++s) f->seg_filter[s] = bd_value(br, 1) ? bd_signed_value(br, 6) : 0;
    }
    if (f->update_map) {
      for (s = 0; s < 3; ++s) f->seg_prob[s] = bd_value(br, 1) ? (uint8_t)bd_value(br, 8) : 255u;
    }
  }
  if (br->eof) return VP8O_BITSTREAM_ERROR;
  // filter header (vp8_dec.c:237-260)
  f->simple = (int)bd_value(br, 1);
  f->level = (int)bd_value(br, 6);
  f->sharpness = (int)bd_value(br, 3);
  f->use_lf_delta = (int)bd_value(br, 1);
  if (f->use_lf_delta && bd_value(br, 1)) {
    for (i = 0; i < 4; ++i) if (bd_value(br, 1)) f->ref_lf_delta[i] = bd_signed_value(br, 6);
    for (i = 0; i < 4; ++i) if (bd_value(br, 1)) f->mode_lf_delta[i] = bd_signed_value(br, 6);
  }
  f->filter_type = (f->level == 0) ? 0 : f->simple ? 1 : 2;
  if (br->eof) return VP8O_BITSTREAM_ERROR;
  // token partitions (vp8_dec.c:203-234)
  {
    const uint8_t* sz = buf;
    const uint8_t* start;
    size_t left = size, last;
    f->num_parts = 1 << bd_value(br, 2);
    last = (size_t)f->num_parts - 1;
    if (size < 3 * last) return VP8O_NOT_ENOUGH_DATA;
    start = buf + last * 3;
    left -= last * 3;
    for (p = 0; p < (int)last; ++p) {
      size_t psize = le24(sz);
      if (psize > left) psize = left;
      bd_init(&f->parts[p], start, psize);
      start += psize; left -= psize; sz += 3;
    }
    bd_init(&f->parts[last], start, left);
    if (!(start < buf + size)) return VP8O_NOT_ENOUGH_DATA;
  }
  // quantisers (quant_dec.c:62-112)
  {
    const int base = (int)bd_value(br, 7);
    const int dy1dc = bd_value(br, 1) ? bd_signed_value(br, 4) : 0;
    const int dy2dc = bd_value(br, 1) ? bd_signed_value(br, 4) : 0;
    const int dy2ac = bd_value(br, 1) ? bd_signed_value(br, 4) : 0;
    const int duvdc = bd_value(br, 1) ? bd_signed_value(br, 4) : 0;
    const int duvac = bd_value(br, 1) ? bd_signed_value(br, 4) : 0;
    for (s = 0; s < 4; ++s) {
      int q;
      if (f->use_segment) {
        q = f->seg_quant[s] + (f->absolute_delta ? 0 : base);
      } else if (s > 0) {
        memcpy(f->dq[s], f->dq[0], sizeof(f->dq[0]));
        continue;
      } else {
        q = base;
      }
      f->dq[s][0] = kVp8DcQ[clipi(q + dy1dc, 0, 127)];
      f->dq[s][1] = kVp8AcQ[clipi(q, 0, 127)];
      f->dq[s][2] = kVp8DcQ[clipi(q + dy2dc, 0, 127)] * 2;
      f->dq[s][3] = (kVp8AcQ[clipi(q + dy2ac, 0, 127)] * 101581) >> 16;
      if (f->dq[s][3] < 8) f->dq[s][3] = 8;
      f->dq[s][4] = kVp8DcQ[clipi(q + duvdc, 0, 117)];
      f->dq[s][5] = kVp8AcQ[clipi(q + duvac, 0, 127)];
    }
  }
  bd_value(br, 1);   // update_proba, ignored on key frames
  // coefficient probabilities (tree_dec.c:515-538)
  i = 0;
  for (t = 0; t < 4; ++t) for (b = 0; b < 8; ++b) for (c = 0; c < 3; ++c) for (p = 0; p < 11; ++p, ++i) {
    f->prob[t][b][c][p] = bd_bit(br, kVp8CoeffUpdateProba[i]) ? (uint8_t)bd_value(br, 8) : kVp8CoeffProba0[i];
  }
  f->use_skip = (int)bd_value(br, 1);
  if (f->use_skip) f->skip_p = (int)bd_value(br, 8);
  return VP8O_OK;
}

// PrecomputeFilterStrengths (src/dec/frame_dec.c:265-313).
static void filter_strengths(Frame* f) {
  int s, i4;
  for (s = 0; s < 4; ++s) {
    int base = f->level;
    if (f->use_segment) base = f->seg_filter[s] + (f->absolute_delta ? 0 : f->level);
    for (i4 = 0; i4 <= 1; ++i4) {
      FInfo* const info = &f->fstr[s][i4];
      int level = base;
      if (f->use_lf_delta) {
        level += f->ref_lf_delta[0];
        if (i4) level += f->mode_lf_delta[0];
      }
      level = clipi(level, 0, 63);
      memset(info, 0, sizeof(*info));
      if (level > 0) {
        int ilevel = level;
        if (f->sharpness > 0) {
          ilevel >>= (f->sharpness > 4) ? 2 : 1;
          if (ilevel > 9 - f->sharpness) ilevel = 9 - f->sharpness;
        }
        if (ilevel < 1) ilevel = 1;
        info->ilevel = (uint8_t)ilevel;
        info->limit = (uint8_t)(2 * level + ilevel);
        info->hev = (level >= 40) ? 2 : (level >= 15) ? 1 : 0;
      }
      info->inner = (uint8_t)i4;
    }
  }
}

// =====================================================================================================
// Intra modes (src/dec/tree_dec.c:290-367). Mode numbering: common_dec.h:18-40.
enum { DC_PRED = 0, TM_PRED = 1, V_PRED = 2, H_PRED = 3,
       B_DC = 0, B_TM, B_VE, B_HE, B_RD, B_VR, B_LD, B_VL, B_HD, B_HU };

// Sub-block mode tree as {child-if-0, child-if-1}; a value <= 0 is a leaf holding -mode.
static const int8_t kBModeTree[9][2] = {
  { -B_DC, 1 }, { -B_TM, 2 }, { -B_VE, 3 }, { 4, 6 }, { -B_HE, 5 }, { -B_RD, -B_VR }, { -B_LD, 7 },
  { -B_VL, 8 }, { -B_HD, -B_HU }
};

static int parse_modes(Frame* f) {
  BoolDec* br = &f->br;
  uint8_t* top = (uint8_t*)calloc((size_t)f->mb_w * 4, 1);   // B_DC == 0
  uint8_t left[4];
  int mx, my, x, y;
  if (top == NULL) return VP8O_OUT_OF_MEMORY;
  for (my = 0; my < f->mb_h; ++my) {
    memset(left, B_DC, 4);
    for (mx = 0; mx < f->mb_w; ++mx) {
      MbModes* m = &f->modes[my * f->mb_w + mx];
      uint8_t* t = top + 4 * mx;
      m->segment = 0;
      if (f->update_map) {
        m->segment = !bd_bit(br, f->seg_prob[0]) ? (uint8_t)bd_bit(br, f->seg_prob[1])
                                                 : (uint8_t)(bd_bit(br, f->seg_prob[2]) + 2);
      }
      m->skip = f->use_skip ? (uint8_t)bd_bit(br, f->skip_p) : 0;
      m->is_i4x4 = !bd_bit(br, 145);
      if (!m->is_i4x4) {
        const int ymode = bd_bit(br, 156) ? (bd_bit(br, 128) ? TM_PRED : H_PRED)
                                          : (bd_bit(br, 163) ? V_PRED : DC_PRED);
        m->imodes[0] = (uint8_t)ymode;
        memset(t, ymode, 4);
        memset(left, ymode, 4);
      } else {
        for (y = 0; y < 4; ++y) {
          int ymode = left[y];
          for (x = 0; x < 4; ++x) {
            const uint8_t* pr = &kVp8BModeProba[(t[x] * 10 + ymode) * 9];
            int node = 0;
            do { node = kBModeTree[node][bd_bit(br, pr[node])]; } while (node > 0);
            ymode = -node;
            t[x] = (uint8_t)ymode;
            m->imodes[y * 4 + x] = (uint8_t)ymode;
          }
          left[y] = (uint8_t)ymode;
        }
      }
      m->uvmode = !bd_bit(br, 142) ? DC_PRED : !bd_bit(br, 114) ? V_PRED : bd_bit(br, 183) ? TM_PRED : H_PRED;
    }
    if (br->eof) { free(top); return VP8O_NOT_ENOUGH_DATA; }
  }
  free(top);
  return VP8O_OK;
}

// =====================================================================================================
// Coefficient tokens (src/dec/vp8_dec.c:400-635).
static const uint8_t kZigzag[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
static const uint8_t kBand[17] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0 };
static const uint8_t kCatProbs[4][12] = {
  { 173, 148, 140, 0 }, { 176, 155, 140, 135, 0 }, { 180, 157, 141, 134, 130, 0 },
  { 254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0 }
};

static int large_value(BoolDec* d, const uint8_t* p) {
  int v;
  if (!bd_bit(d, p[3])) {
    v = !bd_bit(d, p[4]) ? 2 : 3 + bd_bit(d, p[5]);
  } else if (!bd_bit(d, p[6])) {
    if (!bd_bit(d, p[7])) {
      v = 5 + bd_bit(d, 159);
    } else {
      v = 7 + 2 * bd_bit(d, 165);
      v += bd_bit(d, 145);
    }
  } else {
    const int b1 = bd_bit(d, p[8]);
    const int b0 = bd_bit(d, p[9 + b1]);
    const int cat = 2 * b1 + b0;
    const uint8_t* tab;
    v = 0;
    for (tab = kCatProbs[cat]; *tab; ++tab) v += v + bd_bit(d, *tab);
    v += 3 + (8 << cat);
  }
  return v;
}

// Returns (index of the last non-zero coefficient) + 1. `probs` = f->prob[type].
static int block_coeffs(BoolDec* d, uint8_t (*probs)[3][11], int ctx, const int* dq, int n, int16_t* out) {
  const uint8_t* p = probs[kBand[n]][ctx];
  for (; n < 16; ++n) {
    if (!bd_bit(d, p[0])) return n;
    while (!bd_bit(d, p[1])) {
      p = probs[kBand[++n]][0];
      if (n == 16) return 16;
    }
    {
      int v;
      if (!bd_bit(d, p[2])) {
        v = 1;
        p = probs[kBand[n + 1]][1];
      } else {
        v = large_value(d, p);
        p = probs[kBand[n + 1]][2];
      }
      out[kZigzag[n]] = (int16_t)((bd_sign(d) ? -v : v) * dq[n > 0]);
    }
  }
  return 16;
}

static void inverse_wht(const int16_t* in, int16_t* out) {   // src/dsp/dec.c:137-162
  int tmp[16], i;
  for (i = 0; i < 4; ++i) {
    const int a0 = in[0 + i] + in[12 + i], a1 = in[4 + i] + in[8 + i];
    const int a2 = in[4 + i] - in[8 + i], a3 = in[0 + i] - in[12 + i];
    tmp[0 + i] = a0 + a1; tmp[8 + i] = a0 - a1; tmp[4 + i] = a3 + a2; tmp[12 + i] = a3 - a2;
  }
  for (i = 0; i < 4; ++i) {
    const int dc = tmp[0 + i * 4] + 3;
    const int a0 = dc + tmp[3 + i * 4], a1 = tmp[1 + i * 4] + tmp[2 + i * 4];
    const int a2 = tmp[1 + i * 4] - tmp[2 + i * 4], a3 = dc - tmp[3 + i * 4];
    out[0] = (int16_t)((a0 + a1) >> 3); out[16] = (int16_t)((a3 + a2) >> 3);
    out[32] = (int16_t)((a0 - a1) >> 3); out[48] = (int16_t)((a3 - a2) >> 3);
    out += 64;
  }
}

static uint32_t nz_code(uint32_t acc, int nz, int dc_nz) {
  return (acc << 2) | (uint32_t)((nz > 3) ? 3 : (nz > 1) ? 2 : dc_nz);
}

typedef struct { uint8_t nz, nz_dc; } NzCtx;

static int parse_tokens(Frame* f) {
  NzCtx* top = (NzCtx*)calloc((size_t)f->mb_w, sizeof(NzCtx));
  int mx, my, x, y, ch;
  if (top == NULL) return VP8O_OUT_OF_MEMORY;
  for (my = 0; my < f->mb_h; ++my) {
    BoolDec* d = &f->parts[my & (f->num_parts - 1)];
    NzCtx left = { 0, 0 };
    for (mx = 0; mx < f->mb_w; ++mx) {
      const int idx = my * f->mb_w + mx;
      const MbModes* m = &f->modes[idx];
      NzCtx* t = &top[mx];
      int16_t* dst = f->coeffs + (size_t)idx * 384;
      uint32_t nzy = 0, nzuv = 0;
      int skip = f->use_skip ? m->skip : 0;
      if (!skip) {
        const int* q = f->dq[m->segment];
        uint8_t tnz, lnz;
        uint32_t out_t, out_l;
        int first, type;
        if (!m->is_i4x4) {
          int16_t dc[16] = { 0 };
          const int nz = block_coeffs(d, f->prob[1], t->nz_dc + left.nz_dc, q + 2, 0, dc);
          t->nz_dc = left.nz_dc = (uint8_t)(nz > 0);
          if (nz > 1) {
            inverse_wht(dc, dst);
          } else {
            const int dc0 = (dc[0] + 3) >> 3;
            for (x = 0; x < 16; ++x) dst[x * 16] = (int16_t)dc0;
          }
          first = 1; type = 0;
        } else {
          first = 0; type = 3;
        }
        tnz = t->nz & 0x0f; lnz = left.nz & 0x0f;
        for (y = 0; y < 4; ++y) {
          int l = lnz & 1;
          uint32_t acc = 0;
          for (x = 0; x < 4; ++x) {
            const int nz = block_coeffs(d, f->prob[type], l + (tnz & 1), q + 0, first, dst);
            l = (nz > first);
            tnz = (uint8_t)((tnz >> 1) | (l << 7));
            acc = nz_code(acc, nz, dst[0] != 0);
            dst += 16;
          }
          tnz >>= 4;
          lnz = (uint8_t)((lnz >> 1) | (l << 7));
          nzy = (nzy << 8) | acc;
        }
        out_t = tnz; out_l = lnz >> 4;
        for (ch = 0; ch < 4; ch += 2) {
          uint32_t acc = 0;
          tnz = (uint8_t)(t->nz >> (4 + ch)); lnz = (uint8_t)(left.nz >> (4 + ch));
          for (y = 0; y < 2; ++y) {
            int l = lnz & 1;
            for (x = 0; x < 2; ++x) {
              const int nz = block_coeffs(d, f->prob[2], l + (tnz & 1), q + 4, 0, dst);
              l = (nz > 0);
              tnz = (uint8_t)((tnz >> 1) | (l << 3));
              acc = nz_code(acc, nz, dst[0] != 0);
              dst += 16;
            }
            tnz >>= 2;
            lnz = (uint8_t)((lnz >> 1) | (l << 5));
          }
          nzuv |= acc << (4 * ch);
          out_t |= (uint32_t)(tnz << 4) << ch;
          out_l |= (uint32_t)(lnz & 0xf0) << ch;
        }
        t->nz = (uint8_t)out_t; left.nz = (uint8_t)out_l;
        skip = !(nzy | nzuv);
      } else {
        t->nz = left.nz = 0;
        if (!m->is_i4x4) t->nz_dc = left.nz_dc = 0;
      }
      f->nz[idx * 2 + 0] = nzy; f->nz[idx * 2 + 1] = nzuv;
      if (f->filter_type > 0) {
        f->finfo[idx] = f->fstr[m->segment][m->is_i4x4];
        f->finfo[idx].inner |= (uint8_t)!skip;
      }
      if (d->eof) { free(top); return VP8O_NOT_ENOUGH_DATA; }
    }
  }
  free(top);
  return VP8O_OK;
}

// =====================================================================================================
// Reconstruction (src/dec/frame_dec.c:71-196, src/dsp/dec.c). Works in a 32-byte-stride scratch block laid
// out like the reference's yuv_b_ so every predictor reads its neighbours at the same relative offsets.
#define BPS 32
static uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
#define MUL1(a) ((((a) * 20091) >> 16) + (a))
#define MUL2(a) (((a) * 35468) >> 16)

static void idct_add(const int16_t* in, uint8_t* dst) {   // TransformOne_C, dsp/dec.c:44-81
  int tmp[16], i;
  for (i = 0; i < 4; ++i) {
    const int a = in[i] + in[8 + i], b = in[i] - in[8 + i];
    const int c = MUL2(in[4 + i]) - MUL1(in[12 + i]), d = MUL1(in[4 + i]) + MUL2(in[12 + i]);
    tmp[4 * i + 0] = a + d; tmp[4 * i + 1] = b + c; tmp[4 * i + 2] = b - c; tmp[4 * i + 3] = a - d;
  }
  for (i = 0; i < 4; ++i) {
    const int dc = tmp[i] + 4;
    const int a = dc + tmp[8 + i], b = dc - tmp[8 + i];
    const int c = MUL2(tmp[4 + i]) - MUL1(tmp[12 + i]), d = MUL1(tmp[4 + i]) + MUL2(tmp[12 + i]);
    uint8_t* r = dst + i * BPS;
    r[0] = clip8(r[0] + ((a + d) >> 3)); r[1] = clip8(r[1] + ((b + c) >> 3));
    r[2] = clip8(r[2] + ((b - c) >> 3)); r[3] = clip8(r[3] + ((a - d) >> 3));
  }
}

static void idct_ac3_add(const int16_t* in, uint8_t* dst) {   // TransformAC3_C, dsp/dec.c:84-95
  const int a = in[0] + 4, c4 = MUL2(in[4]), d4 = MUL1(in[4]), c1 = MUL2(in[1]), d1 = MUL1(in[1]);
  const int dcs[4] = { a + d4, a + c4, a - c4, a - d4 };
  int y;
  for (y = 0; y < 4; ++y) {
    uint8_t* r = dst + y * BPS;
    r[0] = clip8(r[0] + ((dcs[y] + d1) >> 3)); r[1] = clip8(r[1] + ((dcs[y] + c1) >> 3));
    r[2] = clip8(r[2] + ((dcs[y] - c1) >> 3)); r[3] = clip8(r[3] + ((dcs[y] - d1) >> 3));
  }
}

static void idct_dc_add(const int16_t* in, uint8_t* dst) {   // TransformDC_C, dsp/dec.c:113-121
  const int dc = (in[0] + 4) >> 3;
  int x, y;
  for (y = 0; y < 4; ++y) for (x = 0; x < 4; ++x) dst[x + y * BPS] = clip8(dst[x + y * BPS] + dc);
}

static void do_transform(uint32_t bits, const int16_t* src, uint8_t* dst) {   // frame_dec.c:43-58
  switch (bits >> 30) {
    case 3: idct_add(src, dst); break;
    case 2: idct_ac3_add(src, dst); break;
    case 1: idct_dc_add(src, dst); break;
    default: break;
  }
}

static void do_uv_transform(uint32_t bits, const int16_t* src, uint8_t* dst) {   // frame_dec.c:60-69
  int k;
  if (!(bits & 0xff)) return;
  for (k = 0; k < 4; ++k) {
    uint8_t* d = dst + (k & 1) * 4 + (k >> 1) * 4 * BPS;
    if (bits & 0xaa) idct_add(src + k * 16, d);
    else if (src[k * 16]) idct_dc_add(src + k * 16, d);
  }
}

static void fill(uint8_t* dst, int v, int size) {
  int y;
  for (y = 0; y < size; ++y) memset(dst + y * BPS, v, (size_t)size);
}

static void pred_tm(uint8_t* dst, int size) {   // TrueMotion, dsp/dec.c:173-186
  const uint8_t* top = dst - BPS;
  int x, y;
  for (y = 0; y < size; ++y) {
    for (x = 0; x < size; ++x) dst[x + y * BPS] = clip8(top[x] + dst[y * BPS - 1] - top[-1]);
  }
}

// mode: 0 DC, 1 TM, 2 V, 3 H, 4 DC-no-top, 5 DC-no-left, 6 DC-nothing. size 16 (luma) or 8 (chroma).
static void pred_block(uint8_t* dst, int mode, int size) {
  const int sh = (size == 16) ? 4 : 3;
  int x, y, dc;
  switch (mode) {
    case 0:
      dc = size;
      for (x = 0; x < size; ++x) dc += dst[x - BPS] + dst[x * BPS - 1];
      fill(dst, dc >> (sh + 1), size);
      break;
    case 1: pred_tm(dst, size); break;
    case 2: for (y = 0; y < size; ++y) memcpy(dst + y * BPS, dst - BPS, (size_t)size); break;
    case 3: for (y = 0; y < size; ++y) memset(dst + y * BPS, dst[y * BPS - 1], (size_t)size); break;
    case 4:
      dc = size >> 1;
      for (y = 0; y < size; ++y) dc += dst[y * BPS - 1];
      fill(dst, dc >> sh, size);
      break;
    case 5:
      dc = size >> 1;
      for (x = 0; x < size; ++x) dc += dst[x - BPS];
      fill(dst, dc >> sh, size);
      break;
    default: fill(dst, 0x80, size); break;
  }
}

#define AVG3(a, b, c) ((uint8_t)(((a) + 2 * (b) + (c) + 2) >> 2))
#define AVG2(a, b) ((uint8_t)(((a) + (b) + 1) >> 1))
#define DST(x, y) dst[(x) + (y) * BPS]

static void pred4(uint8_t* dst, int mode) {   // dsp/dec.c:256-410
  const uint8_t* top = dst - BPS;
  const int X = top[-1], A = top[0], B = top[1], C = top[2], D = top[3];
  const int E = top[4], F = top[5], G = top[6], H = top[7];
  const int I = dst[-1], J = dst[-1 + BPS], K = dst[-1 + 2 * BPS], L = dst[-1 + 3 * BPS];
  int x, y;
  switch (mode) {
    case B_DC: {
      int dc = 4;
      for (x = 0; x < 4; ++x) dc += top[x] + dst[-1 + x * BPS];
      fill(dst, dc >> 3, 4);
      break;
    }
    case B_TM: pred_tm(dst, 4); break;
    case B_VE: {
      const uint8_t v[4] = { AVG3(X, A, B), AVG3(A, B, C), AVG3(B, C, D), AVG3(C, D, E) };
      for (y = 0; y < 4; ++y) memcpy(dst + y * BPS, v, 4);
      break;
    }
    case B_HE:
      memset(dst + 0 * BPS, AVG3(X, I, J), 4); memset(dst + 1 * BPS, AVG3(I, J, K), 4);
      memset(dst + 2 * BPS, AVG3(J, K, L), 4); memset(dst + 3 * BPS, AVG3(K, L, L), 4);
      break;
    case B_RD:
      DST(0, 3) = AVG3(J, K, L);
      DST(1, 3) = DST(0, 2) = AVG3(I, J, K);
      DST(2, 3) = DST(1, 2) = DST(0, 1) = AVG3(X, I, J);
      DST(3, 3) = DST(2, 2) = DST(1, 1) = DST(0, 0) = AVG3(A, X, I);
      DST(3, 2) = DST(2, 1) = DST(1, 0) = AVG3(B, A, X);
      DST(3, 1) = DST(2, 0) = AVG3(C, B, A);
      DST(3, 0) = AVG3(D, C, B);
      break;
    case B_VR:
      DST(0, 0) = DST(1, 2) = AVG2(X, A); DST(1, 0) = DST(2, 2) = AVG2(A, B);
      DST(2, 0) = DST(3, 2) = AVG2(B, C); DST(3, 0) = AVG2(C, D);
      DST(0, 3) = AVG3(K, J, I); DST(0, 2) = AVG3(J, I, X);
      DST(0, 1) = DST(1, 3) = AVG3(I, X, A); DST(1, 1) = DST(2, 3) = AVG3(X, A, B);
      DST(2, 1) = DST(3, 3) = AVG3(A, B, C); DST(3, 1) = AVG3(B, C, D);
      break;
    case B_LD:
      DST(0, 0) = AVG3(A, B, C);
      DST(1, 0) = DST(0, 1) = AVG3(B, C, D);
      DST(2, 0) = DST(1, 1) = DST(0, 2) = AVG3(C, D, E);
      DST(3, 0) = DST(2, 1) = DST(1, 2) = DST(0, 3) = AVG3(D, E, F);
      DST(3, 1) = DST(2, 2) = DST(1, 3) = AVG3(E, F, G);
      DST(3, 2) = DST(2, 3) = AVG3(F, G, H);
      DST(3, 3) = AVG3(G, H, H);
      break;
    case B_VL:
      DST(0, 0) = AVG2(A, B); DST(1, 0) = DST(0, 2) = AVG2(B, C);
      DST(2, 0) = DST(1, 2) = AVG2(C, D); DST(3, 0) = DST(2, 2) = AVG2(D, E);
      DST(0, 1) = AVG3(A, B, C); DST(1, 1) = DST(0, 3) = AVG3(B, C, D);
      DST(2, 1) = DST(1, 3) = AVG3(C, D, E); DST(3, 1) = DST(2, 3) = AVG3(D, E, F);
      DST(3, 2) = AVG3(E, F, G); DST(3, 3) = AVG3(F, G, H);
      break;
    case B_HD:
      DST(0, 0) = DST(2, 1) = AVG2(I, X); DST(0, 1) = DST(2, 2) = AVG2(J, I);
      DST(0, 2) = DST(2, 3) = AVG2(K, J); DST(0, 3) = AVG2(L, K);
      DST(3, 0) = AVG3(A, B, C); DST(2, 0) = AVG3(X, A, B);
      DST(1, 0) = DST(3, 1) = AVG3(I, X, A); DST(1, 1) = DST(3, 2) = AVG3(J, I, X);
      DST(1, 2) = DST(3, 3) = AVG3(K, J, I); DST(1, 3) = AVG3(L, K, J);
      break;
    default:   // B_HU
      DST(0, 0) = AVG2(I, J); DST(2, 0) = DST(0, 1) = AVG2(J, K); DST(2, 1) = DST(0, 2) = AVG2(K, L);
      DST(1, 0) = AVG3(I, J, K); DST(3, 0) = DST(1, 1) = AVG3(J, K, L); DST(3, 1) = DST(1, 2) = AVG3(K, L, L);
      DST(3, 2) = DST(2, 2) = DST(0, 3) = DST(1, 3) = DST(2, 3) = DST(3, 3) = (uint8_t)L;
      break;
  }
}

static int check_mode(int mx, int my, int mode) {   // frame_dec.c:28-37
  if (mode == DC_PRED) {
    if (mx == 0) return (my == 0) ? 6 : 5;
    return (my == 0) ? 4 : 0;
  }
  return mode;
}

static void reconstruct(Frame* f) {
  // scratch: row -1 and column -1..-4 hold the neighbours. Offsets follow src/dec/vp8i_dec.h:37-62.
  uint8_t buf[BPS * 17 + BPS * 9 + 64];
  uint8_t* const yb = buf + BPS + 8;
  uint8_t* const ub = buf + BPS * 18 + 8;
  uint8_t* const vb = ub + 16;
  int mx, my, j, n;
  memset(buf, 0, sizeof(buf));
  for (my = 0; my < f->mb_h; ++my) {
    for (j = 0; j < 16; ++j) yb[j * BPS - 1] = 129;
    for (j = 0; j < 8; ++j) ub[j * BPS - 1] = vb[j * BPS - 1] = 129;
    if (my > 0) {
      yb[-1 - BPS] = ub[-1 - BPS] = vb[-1 - BPS] = 129;
    } else {
      memset(yb - BPS - 1, 127, 16 + 4 + 1);
      memset(ub - BPS - 1, 127, 8 + 1);
      memset(vb - BPS - 1, 127, 8 + 1);
    }
    for (mx = 0; mx < f->mb_w; ++mx) {
      const int idx = my * f->mb_w + mx;
      const MbModes* m = &f->modes[idx];
      const int16_t* coeffs = f->coeffs + (size_t)idx * 384;
      uint32_t bits = f->nz[idx * 2];
      const uint32_t bits_uv = f->nz[idx * 2 + 1];
      uint8_t* const yo = f->y + (size_t)my * 16 * f->ys + mx * 16;
      uint8_t* const uo = f->u + (size_t)my * 8 * f->uvs + mx * 8;
      uint8_t* const vo = f->v + (size_t)my * 8 * f->uvs + mx * 8;
      if (mx > 0) {   // left neighbours (incl. the top-left corner) = right edge of the previous macroblock
        for (j = -1; j < 16; ++j) memcpy(&yb[j * BPS - 4], &yb[j * BPS + 12], 4);
        for (j = -1; j < 8; ++j) { memcpy(&ub[j * BPS - 4], &ub[j * BPS + 4], 4); memcpy(&vb[j * BPS - 4], &vb[j * BPS + 4], 4); }
      }
      if (my > 0) {   // top neighbours = last unfiltered row of the macroblock above
        memcpy(yb - BPS, yo - f->ys, 16);
        memcpy(ub - BPS, uo - f->uvs, 8);
        memcpy(vb - BPS, vo - f->uvs, 8);
      }
      if (m->is_i4x4) {
        uint8_t* tr = yb - BPS + 16;
        if (my > 0) {
          if (mx >= f->mb_w - 1) memset(tr, yb[-BPS + 15], 4);
          else memcpy(tr, yo - f->ys + 16, 4);
        }
        memcpy(tr + 4 * BPS, tr, 4); memcpy(tr + 8 * BPS, tr, 4); memcpy(tr + 12 * BPS, tr, 4);
        for (n = 0; n < 16; ++n, bits <<= 2) {
          uint8_t* dst = yb + (n & 3) * 4 + (n >> 2) * 4 * BPS;
          pred4(dst, m->imodes[n]);
          do_transform(bits, coeffs + n * 16, dst);
        }
      } else {
        pred_block(yb, check_mode(mx, my, m->imodes[0]), 16);
        if (bits != 0) {
          for (n = 0; n < 16; ++n, bits <<= 2) do_transform(bits, coeffs + n * 16, yb + (n & 3) * 4 + (n >> 2) * 4 * BPS);
        }
      }
      {
        const int mode = check_mode(mx, my, m->uvmode);
        pred_block(ub, mode, 8);
        pred_block(vb, mode, 8);
        do_uv_transform(bits_uv >> 0, coeffs + 16 * 16, ub);
        do_uv_transform(bits_uv >> 8, coeffs + 20 * 16, vb);
      }
      for (j = 0; j < 16; ++j) memcpy(yo + (size_t)j * f->ys, yb + j * BPS, 16);
      for (j = 0; j < 8; ++j) { memcpy(uo + (size_t)j * f->uvs, ub + j * BPS, 8); memcpy(vo + (size_t)j * f->uvs, vb + j * BPS, 8); }
    }
  }
}

// =====================================================================================================
// Loop filter (src/dec/frame_dec.c:203-260, src/dsp/dec.c:484-693). Clip tables written as clamps
// (src/dsp/dec_clip_tables.c:340-367).
static int sclip1(int v) { return clipi(v, -128, 127); }
static int sclip2(int v) { return clipi(v, -16, 15); }
static int iabs(int v) { return v < 0 ? -v : v; }

static void filter2(uint8_t* p, int step) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  const int a = 3 * (q0 - p0) + sclip1(p1 - q1);
  const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3);
  p[-step] = clip8(p0 + a2);
  p[0] = clip8(q0 - a1);
}

static void filter4(uint8_t* p, int step) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  const int a = 3 * (q0 - p0);
  const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3), a3 = (a1 + 1) >> 1;
  p[-2 * step] = clip8(p1 + a3); p[-step] = clip8(p0 + a2);
  p[0] = clip8(q0 - a1); p[step] = clip8(q1 - a3);
}

static void filter6(uint8_t* p, int step) {
  const int p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step], q2 = p[2 * step];
  const int a = sclip1(3 * (q0 - p0) + sclip1(p1 - q1));
  const int a1 = (27 * a + 63) >> 7, a2 = (18 * a + 63) >> 7, a3 = (9 * a + 63) >> 7;
  p[-3 * step] = clip8(p2 + a3); p[-2 * step] = clip8(p1 + a2); p[-step] = clip8(p0 + a1);
  p[0] = clip8(q0 - a1); p[step] = clip8(q1 - a2); p[2 * step] = clip8(q2 - a3);
}

static int hev(const uint8_t* p, int step, int t) {
  return iabs(p[-2 * step] - p[-step]) > t || iabs(p[step] - p[0]) > t;
}

static int needs_filter(const uint8_t* p, int step, int t) {
  return 4 * iabs(p[-step] - p[0]) + iabs(p[-2 * step] - p[step]) <= t;
}

static int needs_filter2(const uint8_t* p, int step, int t, int it) {
  const int p3 = p[-4 * step], p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step];
  const int q0 = p[0], q1 = p[step], q2 = p[2 * step], q3 = p[3 * step];
  if (4 * iabs(p0 - q0) + iabs(p1 - q1) > t) return 0;
  return iabs(p3 - p2) <= it && iabs(p2 - p1) <= it && iabs(p1 - p0) <= it &&
         iabs(q3 - q2) <= it && iabs(q2 - q1) <= it && iabs(q1 - q0) <= it;
}

// `across` = distance between the pixels straddling the edge, `along` = distance between edge positions.
static void simple_edge(uint8_t* p, int across, int along, int thresh) {
  int i;
  for (i = 0; i < 16; ++i) if (needs_filter(p + i * along, across, 2 * thresh + 1)) filter2(p + i * along, across);
}

static void normal_edge(uint8_t* p, int across, int along, int size, int thresh, int ithresh, int hev_t, int mb_edge) {
  while (size-- > 0) {
    if (needs_filter2(p, across, 2 * thresh + 1, ithresh)) {
      if (hev(p, across, hev_t)) filter2(p, across);
      else if (mb_edge) filter6(p, across);
      else filter4(p, across);
    }
    p += along;
  }
}

static void loop_filter(Frame* f) {
  int mx, my, k;
  for (my = 0; my < f->mb_h; ++my) {
    for (mx = 0; mx < f->mb_w; ++mx) {
      const FInfo* fi = &f->finfo[my * f->mb_w + mx];
      const int limit = fi->limit, il = fi->ilevel, ht = fi->hev;
      uint8_t* const y = f->y + (size_t)my * 16 * f->ys + mx * 16;
      const int ys = f->ys, uvs = f->uvs;
      if (limit == 0) continue;
      if (f->filter_type == 1) {
        if (mx > 0) simple_edge(y, 1, ys, limit + 4);
        if (fi->inner) for (k = 1; k < 4; ++k) simple_edge(y + 4 * k, 1, ys, limit);
        if (my > 0) simple_edge(y, ys, 1, limit + 4);
        if (fi->inner) for (k = 1; k < 4; ++k) simple_edge(y + 4 * k * ys, ys, 1, limit);
      } else {
        uint8_t* const u = f->u + (size_t)my * 8 * uvs + mx * 8;
        uint8_t* const v = f->v + (size_t)my * 8 * uvs + mx * 8;
        if (mx > 0) {
          normal_edge(y, 1, ys, 16, limit + 4, il, ht, 1);
          normal_edge(u, 1, uvs, 8, limit + 4, il, ht, 1);
          normal_edge(v, 1, uvs, 8, limit + 4, il, ht, 1);
        }
        if (fi->inner) {
          for (k = 1; k < 4; ++k) normal_edge(y + 4 * k, 1, ys, 16, limit, il, ht, 0);
          normal_edge(u + 4, 1, uvs, 8, limit, il, ht, 0);
          normal_edge(v + 4, 1, uvs, 8, limit, il, ht, 0);
        }
        if (my > 0) {
          normal_edge(y, ys, 1, 16, limit + 4, il, ht, 1);
          normal_edge(u, uvs, 1, 8, limit + 4, il, ht, 1);
          normal_edge(v, uvs, 1, 8, limit + 4, il, ht, 1);
        }
        if (fi->inner) {
          for (k = 1; k < 4; ++k) normal_edge(y + 4 * k * ys, ys, 1, 16, limit, il, ht, 0);
          normal_edge(u + 4 * uvs, uvs, 1, 8, limit, il, ht, 0);
          normal_edge(v + 4 * uvs, uvs, 1, 8, limit, il, ht, 0);
        }
      }
    }
  }
}

// =====================================================================================================
// Output (src/dec/io_dec.c:25-109, src/dsp/upsampling.c:37-93, src/dsp/yuv.h:59-144, src/dsp/yuv.c:22-63).
static int yuv_clip(int v) { return ((v & ~16383) == 0) ? (v >> 6) : (v < 0) ? 0 : 255; }
static int mult_hi(int v, int c) { return (v * c) >> 8; }

static void put_pixel(int y, int u, int v, int csp, uint8_t* dst) {
  const int r = yuv_clip(mult_hi(y, 19077) + mult_hi(v, 26149) - 14234);
  const int g = yuv_clip(mult_hi(y, 19077) - mult_hi(u, 6419) - mult_hi(v, 13320) + 8708);
  const int b = yuv_clip(mult_hi(y, 19077) + mult_hi(u, 33050) - 17685);
  switch (csp) {
    case VP8O_RGB: dst[0] = (uint8_t)r; dst[1] = (uint8_t)g; dst[2] = (uint8_t)b; break;
    case VP8O_BGR: dst[0] = (uint8_t)b; dst[1] = (uint8_t)g; dst[2] = (uint8_t)r; break;
    case VP8O_RGBA: case VP8O_rgbA: dst[0] = (uint8_t)r; dst[1] = (uint8_t)g; dst[2] = (uint8_t)b; dst[3] = 0xff; break;
    case VP8O_BGRA: case VP8O_bgrA: dst[0] = (uint8_t)b; dst[1] = (uint8_t)g; dst[2] = (uint8_t)r; dst[3] = 0xff; break;
    default: dst[0] = 0xff; dst[1] = (uint8_t)r; dst[2] = (uint8_t)g; dst[3] = (uint8_t)b; break;   // ARGB / Argb
  }
}

static int csp_bpp(int csp) { return (csp == VP8O_RGB || csp == VP8O_BGR) ? 3 : 4; }

// One call of the reference's line-pair upsampler; bottom_y/bottom_dst may be NULL.
static void upsample_pair(const uint8_t* top_y, const uint8_t* bot_y, const uint8_t* top_u, const uint8_t* top_v,
                          const uint8_t* cur_u, const uint8_t* cur_v, uint8_t* top_dst, uint8_t* bot_dst,
                          int len, int csp) {
  const int bpp = csp_bpp(csp);
  const int last_pair = (len - 1) >> 1;
  uint32_t tl = top_u[0] | ((uint32_t)top_v[0] << 16);
  uint32_t l = cur_u[0] | ((uint32_t)cur_v[0] << 16);
  int x;
  {
    const uint32_t uv0 = (3 * tl + l + 0x00020002u) >> 2;
    put_pixel(top_y[0], uv0 & 0xff, (uv0 >> 16), csp, top_dst);
  }
  if (bot_y != NULL) {
    const uint32_t uv0 = (3 * l + tl + 0x00020002u) >> 2;
    put_pixel(bot_y[0], uv0 & 0xff, (uv0 >> 16), csp, bot_dst);
  }
  for (x = 1; x <= last_pair; ++x) {
    const uint32_t t = top_u[x] | ((uint32_t)top_v[x] << 16);
    const uint32_t uv = cur_u[x] | ((uint32_t)cur_v[x] << 16);
    const uint32_t avg = tl + t + l + uv + 0x00080008u;
    const uint32_t d12 = (avg + 2 * (t + l)) >> 3;
    const uint32_t d03 = (avg + 2 * (tl + uv)) >> 3;
    {
      const uint32_t uv0 = (d12 + tl) >> 1, uv1 = (d03 + t) >> 1;
      put_pixel(top_y[2 * x - 1], uv0 & 0xff, (uv0 >> 16) & 0xff, csp, top_dst + (2 * x - 1) * bpp);
      put_pixel(top_y[2 * x], uv1 & 0xff, (uv1 >> 16) & 0xff, csp, top_dst + (2 * x) * bpp);
    }
    if (bot_y != NULL) {
      const uint32_t uv0 = (d03 + l) >> 1, uv1 = (d12 + uv) >> 1;
      put_pixel(bot_y[2 * x - 1], uv0 & 0xff, (uv0 >> 16) & 0xff, csp, bot_dst + (2 * x - 1) * bpp);
      put_pixel(bot_y[2 * x], uv1 & 0xff, (uv1 >> 16) & 0xff, csp, bot_dst + (2 * x) * bpp);
    }
    tl = t; l = uv;
  }
  if (!(len & 1)) {
    {
      const uint32_t uv0 = (3 * tl + l + 0x00020002u) >> 2;
      put_pixel(top_y[len - 1], uv0 & 0xff, (uv0 >> 16), csp, top_dst + (len - 1) * bpp);
    }
    if (bot_y != NULL) {
      const uint32_t uv0 = (3 * l + tl + 0x00020002u) >> 2;
      put_pixel(bot_y[len - 1], uv0 & 0xff, (uv0 >> 16), csp, bot_dst + (len - 1) * bpp);
    }
  }
}

static void emit_rgb(const Frame* f, int csp, int fancy, uint8_t* out, int stride) {
  const int w = f->width, h = f->height;
  int y;
  if (!fancy) {
    const int bpp = csp_bpp(csp);
    int x;
    for (y = 0; y < h; ++y) {
      const uint8_t* yr = f->y + (size_t)y * f->ys;
      const uint8_t* ur = f->u + (size_t)(y >> 1) * f->uvs;
      const uint8_t* vr = f->v + (size_t)(y >> 1) * f->uvs;
      for (x = 0; x < w; ++x) put_pixel(yr[x], ur[x >> 1], vr[x >> 1], csp, out + (size_t)y * stride + x * bpp);
    }
    return;
  }
  // EmitFancyRGB over the whole picture: row 0 alone, then pairs (2k-1, 2k), then the last row if h is even.
  upsample_pair(f->y, NULL, f->u, f->v, f->u, f->v, out, NULL, w, csp);
  for (y = 1; y + 1 < h; y += 2) {
    const int k = (y + 1) >> 1;
    upsample_pair(f->y + (size_t)y * f->ys, f->y + (size_t)(y + 1) * f->ys,
                  f->u + (size_t)(k - 1) * f->uvs, f->v + (size_t)(k - 1) * f->uvs,
                  f->u + (size_t)k * f->uvs, f->v + (size_t)k * f->uvs,
                  out + (size_t)y * stride, out + (size_t)(y + 1) * stride, w, csp);
  }
  if (!(h & 1)) {
    const int k = (h >> 1) - 1;
    upsample_pair(f->y + (size_t)(h - 1) * f->ys, NULL, f->u + (size_t)k * f->uvs, f->v + (size_t)k * f->uvs,
                  f->u + (size_t)k * f->uvs, f->v + (size_t)k * f->uvs, out + (size_t)(h - 1) * stride, NULL, w, csp);
  }
}

// =====================================================================================================
static void frame_free(Frame* f) {
  free(f->modes); free(f->coeffs); free(f->nz); free(f->finfo); free(f->y); free(f->unfiltered);
}

// Runs the pipeline up to (and including) the loop filter.
static int decode_frame(const uint8_t* data, size_t size, int bypass_filter, int keep_unfiltered, Frame* f, Container* c) {
  int st = parse_container(data, size, 1, c);
  size_t nmb, ysz, uvsz;
  memset(f, 0, sizeof(*f));
  if (st != VP8O_OK) return st;
  if (c->has_animation) return VP8O_UNSUPPORTED_FEATURE;
  if (c->is_lossless) return VP8O_UNSUPPORTED_FEATURE;   // out of scope for the lossy path
  if (c->alpha != NULL) return VP8O_UNSUPPORTED_FEATURE; // ALPH plane: next row of SURVEY.md 8(f)
  st = parse_frame_header(f, c->vp8, c->vp8_size);
  if (st != VP8O_OK) return st;
  if (bypass_filter) f->filter_type = 0;   // VP8EnterCritical, frame_dec.c:557-560
  filter_strengths(f);
  nmb = (size_t)f->mb_w * f->mb_h;
  f->ys = 16 * f->mb_w; f->uvs = 8 * f->mb_w;
  ysz = (size_t)f->ys * 16 * f->mb_h; uvsz = (size_t)f->uvs * 8 * f->mb_h;
  f->modes = (MbModes*)calloc(nmb, sizeof(MbModes));
  f->coeffs = (int16_t*)calloc(nmb * 384, sizeof(int16_t));
  f->nz = (uint32_t*)calloc(nmb * 2, sizeof(uint32_t));
  f->finfo = (FInfo*)calloc(nmb, sizeof(FInfo));
  f->y = (uint8_t*)calloc(ysz + 2 * uvsz, 1);
  if (!f->modes || !f->coeffs || !f->nz || !f->finfo || !f->y) return VP8O_OUT_OF_MEMORY;
  f->u = f->y + ysz; f->v = f->u + uvsz;
  // The reference interleaves mode rows and token rows (vp8_dec.c:646-660); they read different partitions,
  // so parsing all modes first gives the same result. Error precedence is kept row-wise below.
  {
    const int st_modes = parse_modes(f);
    const int st_tok = (st_modes == VP8O_OK) ? parse_tokens(f) : st_modes;
    if (st_tok != VP8O_OK) return st_tok;
  }
  reconstruct(f);
  if (keep_unfiltered) {
    f->unfiltered = (uint8_t*)malloc(ysz + 2 * uvsz);
    if (!f->unfiltered) return VP8O_OUT_OF_MEMORY;
    memcpy(f->unfiltered, f->y, ysz + 2 * uvsz);
  }
  if (f->filter_type > 0) loop_filter(f);
  return VP8O_OK;
}

int vp8o_decode(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size, int stride) {
  Frame f;
  Container c;
  int st, feat[5];
  st = vp8o_features(data, size, feat);   // WebPDecode: the probe's NOT_ENOUGH_DATA becomes BITSTREAM_ERROR
  if (st != VP8O_OK) return st == VP8O_NOT_ENOUGH_DATA ? VP8O_BITSTREAM_ERROR : st;
  st = decode_frame(data, size, flags & VP8O_FLAG_BYPASS_FILTER, 0, &f, &c);
  if (st == VP8O_OK) {
    const int w = f.width, h = f.height;
    if (csp == VP8O_YUV) {
      const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
      int y;
      if (out_size < (size_t)w * h + 2 * (size_t)uvw * uvh) { frame_free(&f); return VP8O_INVALID_PARAM; }
      for (y = 0; y < h; ++y) memcpy(out + (size_t)y * w, f.y + (size_t)y * f.ys, (size_t)w);
      for (y = 0; y < uvh; ++y) {
        memcpy(out + (size_t)w * h + (size_t)y * uvw, f.u + (size_t)y * f.uvs, (size_t)uvw);
        memcpy(out + (size_t)w * h + (size_t)uvw * uvh + (size_t)y * uvw, f.v + (size_t)y * f.uvs, (size_t)uvw);
      }
    } else {
      const int bpp = csp_bpp(csp);
      if (stride < w * bpp || out_size < (size_t)stride * (h - 1) + (size_t)w * bpp) { frame_free(&f); return VP8O_INVALID_PARAM; }
      emit_rgb(&f, csp, !(flags & VP8O_FLAG_NO_FANCY), out, stride);
    }
  }
  frame_free(&f);
  return st;
}

int vp8o_dump(const uint8_t* data, size_t size, Vp8oDump* d) {
  Frame f;
  Container c;
  size_t nmb, i, plane;
  memset(d, 0, sizeof(*d));
  d->status = decode_frame(data, size, 0, 1, &f, &c);
  if (d->status != VP8O_OK) { frame_free(&f); return d->status; }
  nmb = (size_t)f.mb_w * f.mb_h;
  plane = (size_t)f.ys * 16 * f.mb_h + 2 * (size_t)f.uvs * 8 * f.mb_h;
  d->width = f.width; d->height = f.height; d->mb_w = f.mb_w; d->mb_h = f.mb_h;
  d->filter_type = f.filter_type; d->num_parts = f.num_parts;
  memcpy(d->dq, f.dq, sizeof(d->dq));
  d->modes = (uint8_t*)malloc(nmb * 20);
  d->finfo = (uint8_t*)malloc(nmb * 4);
  for (i = 0; i < nmb; ++i) {
    memcpy(d->modes + i * 20, f.modes[i].imodes, 16);
    d->modes[i * 20 + 16] = f.modes[i].is_i4x4; d->modes[i * 20 + 17] = f.modes[i].uvmode;
    d->modes[i * 20 + 18] = f.modes[i].skip; d->modes[i * 20 + 19] = f.modes[i].segment;
    d->finfo[i * 4 + 0] = f.finfo[i].limit; d->finfo[i * 4 + 1] = f.finfo[i].ilevel;
    d->finfo[i * 4 + 2] = f.finfo[i].inner; d->finfo[i * 4 + 3] = f.finfo[i].hev;
  }
  d->coeffs = f.coeffs; f.coeffs = NULL;
  d->nz = f.nz; f.nz = NULL;
  d->y_stride = f.ys; d->uv_stride = f.uvs;
  d->unfiltered = f.unfiltered; f.unfiltered = NULL;
  d->filtered = (uint8_t*)malloc(plane);
  memcpy(d->filtered, f.y, plane);
  frame_free(&f);
  return VP8O_OK;
}

// Boolean decodes per entropy-coded stream of the frame: out[0] = first partition (headers + intra modes),
// out[1..8] = token partitions, out[9] = number of token partitions. bench.py turns these into "cycles per decode".
int vp8o_count_decodes(const uint8_t* data, size_t size, uint64_t out[10]) {
  Frame f;
  Container c;
  int p;
  const int st = decode_frame(data, size, 0, 1, &f, &c);
  memset(out, 0, 10 * sizeof(out[0]));
  if (st == VP8O_OK) {
    out[0] = f.br.decodes;
    for (p = 0; p < f.num_parts && p < 8; ++p) out[1 + p] = f.parts[p].decodes;
    out[9] = (uint64_t)f.num_parts;
  }
  frame_free(&f);
  return st;
}

void vp8o_dump_free(Vp8oDump* d) {
  free(d->modes); free(d->coeffs); free(d->nz); free(d->finfo); free(d->unfiltered); free(d->filtered);
  memset(d, 0, sizeof(*d));
}
