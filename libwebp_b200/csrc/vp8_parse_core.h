// vp8_parse_core.h -- bitstream side of the batched VP8 decoder: boolean decoder, frame header, intra modes,
// coefficient tokens. One lane walks one entropy-coded stream; everything it needs lives in registers and
// in the shared-memory blocks handed in by the kernels of vp8_parse.cu.
//
// Replaces, for a whole batch at once:
//   VP8GetHeaders / ParseSegmentHeader / ParseFilterHeader / ParsePartitions   src/dec/vp8_dec.c:162-395
//   VP8ParseQuant src/dec/quant_dec.c:62-112, VP8ParseProba src/dec/tree_dec.c:515-538
//   PrecomputeFilterStrengths src/dec/frame_dec.c:265-313
//   VP8ParseIntraModeRow / ParseIntraMode src/dec/tree_dec.c:290-367
//   VP8DecodeMB / ParseResiduals / GetCoeffs / GetLargeValue src/dec/vp8_dec.c:400-635
//   VP8BitReader src/utils/bit_reader_inl_utils.h:107-157 (state kept as range-1, like the reference)
//
// The same source is compiled by nvcc (product) and by g++ with -DVP8_EMU (tests/emu: a host build used only
// to unit-test this logic where no GPU exists; it is never part of the shipped library).
#ifndef LIBWEBP_B200_VP8_PARSE_CORE_H_
#define LIBWEBP_B200_VP8_PARSE_CORE_H_

#include "vp8_dev.h"

#if defined(__CUDACC__) && !defined(VP8_EMU)
#define VP8_FN __device__ __forceinline__
#define VP8_MFN __device__ __forceinline__   // member functions
#define VP8_TABLE static __constant__ const
#define VP8_CLZ(x) __clz((int)(x))
#define VP8_BSWAP(x) __byte_perm((x), 0, 0x0123)
#else
#define VP8_FN static inline
#define VP8_MFN inline
#define VP8_TABLE static const
#ifndef VP8_EMU_VECTORS
#define VP8_EMU_VECTORS
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
#endif
#define VP8_CLZ(x) __builtin_clz((unsigned)(x))
#define VP8_BSWAP(x) __builtin_bswap32(x)
#endif

#include "vp8_tables.cuh"

VP8_TABLE uint8_t kZigzagPos[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
// coefficient index -> probability band, as a byte offset (band * 33) into one type's [8][3][11] table
VP8_TABLE uint16_t kBandOff[17] = { 0, 33, 66, 99, 198, 132, 165, 198, 198, 198, 198, 198, 198, 198, 198, 231, 0 };
VP8_TABLE uint8_t kCatProb[4][12] = {
  { 173, 148, 140, 0 }, { 176, 155, 140, 135, 0 }, { 180, 157, 141, 134, 130, 0 },
  { 254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0 }
};

// ---------------------------------------------------------------------------------------------------------
// Boolean decoder (VP8GetBit, bit_reader_inl_utils.h:107-136) shaped for few instructions per decode.
//   V:vlo   64-bit left-aligned window on the stream, `nbits` of it valid; the decode only looks at V's top byte
//   R24     (range - 1) << 24, range - 1 in [127, 254] as in the reference, so that
//             split            = umulhi(R24, prob)              one IMAD.HI
//             (split + 1)<<24  = split * 2^24 + 2^24            one IMAD
//             bit              = V >= (split + 1) << 24         one compare, no shift of V
//           and the new true range, still << 24, normalises with one count-leading-zeros.
//   nxt     the next aligned word of the stream, fetched one refill ahead of its use (a warp never waits on HBM)
// bd_fill() tops the window up to > 32 valid bits; every decode consumes at most 7, so callers place one
// bd_fill() per four decodes instead of testing before each (see parse_block).
// Words wholly past the stream read as zero; any decode that starts in them is reported by bd_eof(), which
// reproduces the reference's eof_ flag (bit_reader_utils.c:88-101) from the bit position alone.
#if defined(__CUDACC__) && !defined(VP8_EMU)
#define VP8_UNROLL4 _Pragma("unroll 1")
#define VP8_UMULHI(a, b) __umulhi((a), (b))
#define VP8_LDG(p) __ldg(p)
VP8_FN uint32_t vp8_shr_clamp(uint32_t w, int n) { return __funnelshift_rc(w, 0u, (uint32_t)n); }   // 0 when n >= 32
VP8_FN uint32_t vp8_shl_clamp(uint32_t w, int n) { return __funnelshift_lc(0u, w, (uint32_t)n); }
VP8_FN uint32_t vp8_shl_pair(uint32_t hi, uint32_t lo, int n) { return __funnelshift_l(lo, hi, (uint32_t)n); }
VP8_FN int vp8_shiftamt(uint32_t x) { int r; asm("bfind.shiftamt.u32 %0, %1;" : "=r"(r) : "r"(x)); return r; }   // clz, x != 0
#else
#define VP8_UNROLL4
#define VP8_UMULHI(a, b) ((uint32_t)(((uint64_t)(a) * (uint64_t)(b)) >> 32))
#define VP8_LDG(p) (*(p))
VP8_FN uint32_t vp8_shr_clamp(uint32_t w, int n) { return n >= 32 ? 0u : (w >> n); }
VP8_FN uint32_t vp8_shl_clamp(uint32_t w, int n) { return n >= 32 ? 0u : (w << n); }
VP8_FN uint32_t vp8_shl_pair(uint32_t hi, uint32_t lo, int n) { return n == 0 ? hi : ((hi << n) | (lo >> (32 - n))); }
VP8_FN int vp8_shiftamt(uint32_t x) { return __builtin_clz(x); }
#endif

struct BoolDec {
  const uint32_t* wp;    // word after `nxt`
  const uint32_t* wend;  // first word wholly past the stream: from there on zeros are shifted in
  const uint32_t* wbase; // stream bits moved into the window so far = 32 * (wp - wbase) - bias8
  uint32_t V, vlo, nxt;  // nxt holds the raw (little-endian) word
  int nbits;             // valid bits in V:vlo
  uint32_t R24;
  int last_shift;        // renormalisation shift of the most recent decode
  int bias8;
  int64_t limit;         // 8*size - 8: a decode starting beyond this bit position reads past the end
};

VP8_FN uint32_t bd_fetch(BoolDec& d) {
  const uint32_t* p = d.wp++;
  return (p < d.wend) ? VP8_LDG(p) : 0u;
}

VP8_FN void bd_init(BoolDec& d, const uint8_t* start, uint32_t size) {
  const uintptr_t a = (uintptr_t)start;
  const int off = (int)(a & 3);
  d.wp = (const uint32_t*)(a - off);
  d.wend = (const uint32_t*)((a + size + 3) & ~(uintptr_t)3);
  d.wbase = d.wp + 1;
  d.V = VP8_BSWAP(bd_fetch(d)) << (8 * off);
  d.vlo = 0;
  d.nbits = 32 - 8 * off;
  d.nxt = bd_fetch(d);
  d.bias8 = 8 * off;
  d.R24 = 254u << 24;
  d.last_shift = 0;
  d.limit = 8 * (int64_t)size - 8;
}

// True when the reference's reader would have raised eof_: some decode started with fewer than 8 real bits
// left. Positions are monotonic, so testing the most recent decode is enough (and size==0 trips at once).
VP8_FN int bd_eof(const BoolDec& d) {
  const int64_t loaded = 32 * (int64_t)(d.wp - d.wbase) - d.bias8;
  return (loaded - d.nbits - d.last_shift) > d.limit;
}

// Window -> more than 32 valid bits. vlo is empty whenever nbits <= 32.
VP8_FN void bd_fill(BoolDec& d) {
  if (d.nbits <= 32) {
    const uint32_t w = VP8_BSWAP(d.nxt);
    d.nxt = bd_fetch(d);
    d.V |= vp8_shr_clamp(w, d.nbits);
    d.vlo = vp8_shl_clamp(w, 32 - d.nbits);
    d.nbits += 32;
  }
}

// bd_fill() for a warp whose lanes all wait when one waits (vp8_tokens_lockstep.h): the word fetched here is not
// looked at before the NEXT fill -- not even to replace it by zero past the end of the stream, which bd_fetch()
// does with a select that has to wait for the load. The load address is clamped instead and the zero is chosen
// when the word is consumed.
VP8_FN void bd_fill_lookahead(BoolDec& d) {
  if (d.nbits <= 32) {
    const uint32_t* p = d.wp;   // d.nxt came from p - 1
    const uint32_t w = (p - 1 < d.wend) ? VP8_BSWAP(d.nxt) : 0u;
    d.nxt = VP8_LDG(p < d.wend ? p : d.wend);   // wend itself lies inside the arena's tail padding
    d.wp = p + 1;
    d.V |= vp8_shr_clamp(w, d.nbits);
    d.vlo = vp8_shl_clamp(w, 32 - d.nbits);
    d.nbits += 32;
  }
}

// One decode given (split + 1) << 24; needs >= 8 valid bits.
VP8_FN int bd_decode(BoolDec& d, uint32_t s1) {
  const int bit = d.V >= s1;
  uint32_t r24 = s1;                       // true range << 24
  if (bit) { r24 = d.R24 + (1u << 24) - s1; d.V -= s1; }
  const int shift = vp8_shiftamt(r24);
  d.R24 = (r24 << shift) - (1u << 24);
  d.V = vp8_shl_pair(d.V, d.vlo, shift);
  d.vlo <<= shift;
  d.nbits -= shift;
  d.last_shift = shift;
  return bit;
}

VP8_FN int bd_bit_nofill(BoolDec& d, uint32_t prob) {
  return bd_decode(d, (VP8_UMULHI(d.R24, prob) << 24) + (1u << 24));
}
// prob = 128 (VP8GetSigned / VP8GetValue bits): split = range >> 1
VP8_FN int bd_half_nofill(BoolDec& d) { return bd_decode(d, ((d.R24 >> 25) << 24) + (1u << 24)); }

VP8_FN int bd_bit(BoolDec& d, uint32_t prob) {
  bd_fill(d);
  return bd_bit_nofill(d, prob);
}

VP8_FN uint32_t bd_value(BoolDec& d, int n) {
  uint32_t v = 0;
  while (n-- > 0) v |= (uint32_t)bd_bit(d, 0x80) << n;
  return v;
}

VP8_FN int bd_signed(BoolDec& d, int n) {
  const int v = (int)bd_value(d, n);
  return bd_bit(d, 0x80) ? -v : v;
}

// ---------------------------------------------------------------------------------------------------------
// The reference's reader taken literally, for the one case where the window above is not equivalent: a partition whose
// first byte is 0xFF. That byte breaks the decoder's invariant value < range from the first bit on (no encoder writes it),
// and what the reference then decodes depends on the mechanics of its reader (bit_reader_utils.c:26-45,105-119,
// bit_reader_inl_utils.h:58-136 with BITS = 56): seven bytes enter a 64-bit value at a time while eight or more are left (so
// whatever overflowed above the low eight bits is dropped at that moment, not earlier), single bytes after that, one zero byte
// past the end raises eof_; the compare is done on the value cut to 32 bits, and the sign of a coefficient comes from the
// sign bit of a 32-bit difference. Slow (64-bit arithmetic, a byte at a time): only the images the host flags
// (VP8B_FLAG_LITERAL_READER) are parsed with it, by k_parse_literal. Same function names as BoolDec, so the parsers below
// are written once for both.
struct RefBits {
  const uint8_t* buf;
  const uint8_t* end;
  const uint8_t* wide_end;   // seven bytes at once while buf is below this (eight readable bytes left)
  uint64_t value;
  uint32_t range;            // range - 1, as the reference keeps it
  int bits;                  // position of the bit after the eight in use; negative: load before the next decode
  int eof;
};

VP8_FN void rb_load(RefBits& d) {
  if (d.buf < d.wide_end) {
    uint64_t w = 0;
    for (int k = 0; k < 7; ++k) w = (w << 8) | d.buf[k];
    d.buf += 7;
    d.value = w | (d.value << 56);
    d.bits += 56;
  } else if (d.buf < d.end) {
    d.value = (uint64_t)(*d.buf++) | (d.value << 8);
    d.bits += 8;
  } else if (!d.eof) {
    d.value <<= 8;
    d.bits += 8;
    d.eof = 1;
  } else {
    d.bits = 0;
  }
}

VP8_FN void bd_init(RefBits& d, const uint8_t* start, uint32_t size) {
  d.buf = start; d.end = start + size;
  d.wide_end = size >= 8 ? start + size - 7 : start;
  d.value = 0; d.range = 254; d.bits = -8; d.eof = 0;
  rb_load(d);
}
VP8_FN int bd_eof(const RefBits& d) { return d.eof; }
VP8_FN void bd_fill(RefBits&) {}
VP8_FN int bd_bit(RefBits& d, uint32_t prob) {
  uint32_t range = d.range;
  if (d.bits < 0) rb_load(d);
  const int pos = d.bits;
  const uint32_t split = (range * prob) >> 8;
  const uint32_t v32 = (uint32_t)(d.value >> pos);   // cut to 32 bits: part of the behaviour
  const int bit = v32 > split;
  if (bit) { range -= split; d.value -= (uint64_t)(split + 1) << pos; } else { range = split + 1; }
  const int shift = VP8_CLZ(range) - 24;             // range in [1, 255] here
  d.bits -= shift;
  d.range = (range << shift) - 1;
  return bit;
}
VP8_FN int bd_bit_nofill(RefBits& d, uint32_t prob) { return bd_bit(d, prob); }
// the sign of a coefficient (VP8GetSigned): not the same as bd_bit(d, 128) once the value has left its range
VP8_FN int bd_half_nofill(RefBits& d) {
  if (d.bits < 0) rb_load(d);
  const int pos = d.bits;
  const uint32_t split = d.range >> 1;
  const uint32_t v32 = (uint32_t)(d.value >> pos);
  const uint32_t mask = (uint32_t)((int32_t)(split - v32) >> 31);   // all ones: negative
  d.bits -= 1;
  d.range = (d.range + mask) | 1u;
  d.value -= (uint64_t)((split + 1) & mask) << pos;
  return (int)(mask & 1u);
}
VP8_FN uint32_t bd_value(RefBits& d, int n) {
  uint32_t v = 0;
  while (n-- > 0) v |= (uint32_t)bd_bit(d, 0x80) << n;
  return v;
}
VP8_FN int bd_signed(RefBits& d, int n) {
  const int v = (int)bd_value(d, n);
  return bd_bit(d, 0x80) ? -v : v;
}

VP8_FN int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

// ---------------------------------------------------------------------------------------------------------
// Frame header (partition 0 prefix). `frame` points at the 3-byte frame tag. Fills *h; returns the status.
// The host already validated tag, signature, dimensions and part0_size <= available (vp8_container.c).
template <class BD>
VP8_FN int parse_frame_header(BD& br, const uint8_t* frame, const ImgDesc& im, FrameHdr* h) {
  const uint8_t* buf = frame + 10 + im.part0_size;         // first byte after partition 0
  const uint32_t size = im.vp8_size - 10 - im.part0_size;  // bytes left for the token partitions
  int seg_quant[4] = { 0, 0, 0, 0 }, seg_filter[4] = { 0, 0, 0, 0 };
  int ref_lf0 = 0, mode_lf0 = 0;
  int absolute_delta = 1;
  bd_init(br, frame + 10, im.part0_size);
  bd_bit(br, 0x80);   // colorspace
  bd_bit(br, 0x80);   // clamp type
  // segment header (reference: vp8_dec.c:162-196)
  h->seg_prob[0] = h->seg_prob[1] = h->seg_prob[2] = 255;
  h->update_map = 0;
  const int use_segment = bd_bit(br, 0x80);
  if (use_segment) {
    h->update_map = (uint8_t)bd_bit(br, 0x80);
    if (bd_bit(br, 0x80)) {
      absolute_delta = bd_bit(br, 0x80);
      for (int s = 0; s < 4; ++s) seg_quant[s] = bd_bit(br, 0x80) ? bd_signed(br, 7) : 0;
      for (int s = 0; s < 4; ++s) seg_filter[s] = bd_bit(br, 0x80) ? bd_signed(br, 6) : 0;
    }
    if (h->update_map) {
      for (int s = 0; s < 3; ++s) h->seg_prob[s] = bd_bit(br, 0x80) ? (uint8_t)bd_value(br, 8) : 255u;
    }
  }
  if (bd_eof(br)) return VP8B_BITSTREAM_ERROR;
  // filter header (vp8_dec.c:237-260)
  const int simple = bd_bit(br, 0x80);
  const int level = (int)bd_value(br, 6);
  const int sharpness = (int)bd_value(br, 3);
  const int use_lf_delta = bd_bit(br, 0x80);
  if (use_lf_delta && bd_bit(br, 0x80)) {
    for (int i = 0; i < 4; ++i) if (bd_bit(br, 0x80)) { const int v = bd_signed(br, 6); if (i == 0) ref_lf0 = v; }
    for (int i = 0; i < 4; ++i) if (bd_bit(br, 0x80)) { const int v = bd_signed(br, 6); if (i == 0) mode_lf0 = v; }
  }
  int filter_type = (level == 0) ? 0 : simple ? 1 : 2;
  if (bd_eof(br)) return VP8B_BITSTREAM_ERROR;
  // token partitions (vp8_dec.c:203-234)
  {
    const int nparts = 1 << bd_value(br, 2);
    const uint32_t last = (uint32_t)nparts - 1;
    if (size < 3 * last) return VP8B_NOT_ENOUGH_DATA;
    uint32_t off = (uint32_t)(buf - frame) + 3 * last;
    uint32_t left = size - 3 * last;
    for (uint32_t p = 0; p < last; ++p) {
      uint32_t psize = buf[3 * p] | (buf[3 * p + 1] << 8) | (buf[3 * p + 2] << 16);
      if (psize > left) psize = left;
      h->part_off[p] = off; h->part_size[p] = psize;
      off += psize; left -= psize;
    }
    h->part_off[last] = off; h->part_size[last] = left;
    h->num_parts = (uint8_t)nparts;
    if (!(off < im.vp8_size)) return VP8B_NOT_ENOUGH_DATA;
  }
  // quantisers (quant_dec.c:62-112)
  {
    const int base = (int)bd_value(br, 7);
    const int dy1dc = bd_bit(br, 0x80) ? bd_signed(br, 4) : 0;
    const int dy2dc = bd_bit(br, 0x80) ? bd_signed(br, 4) : 0;
    const int dy2ac = bd_bit(br, 0x80) ? bd_signed(br, 4) : 0;
    const int duvdc = bd_bit(br, 0x80) ? bd_signed(br, 4) : 0;
    const int duvac = bd_bit(br, 0x80) ? bd_signed(br, 4) : 0;
    for (int s = 0; s < 4; ++s) {
      int q = base;
      if (use_segment) q = seg_quant[s] + (absolute_delta ? 0 : base);
      else if (s > 0) { for (int k = 0; k < 6; ++k) h->dq[s][k] = h->dq[0][k]; h->dither[s] = h->dither[0]; continue; }
      h->dither[s] = 0;
      if (im.dither_f > 0 && q + duvac < 12) {   // VP8InitDithering, frame_dec.c:328-349 (uv_quant_ = q + dquv_ac)
        const int idx = (q + duvac < 0) ? 0 : q + duvac;
        const int amp = (idx < 3) ? 8 - idx : (idx < 5) ? 4 : (idx < 8) ? 2 : 1;   // kQuantToDitherAmp
        h->dither[s] = (uint8_t)((im.dither_f * amp) >> 3);
      }
      int y2ac = (kVp8AcQ[clampi(q + dy2ac, 0, 127)] * 101581) >> 16;
      if (y2ac < 8) y2ac = 8;
      h->dq[s][0] = kVp8DcQ[clampi(q + dy1dc, 0, 127)];
      h->dq[s][1] = (int16_t)kVp8AcQ[clampi(q, 0, 127)];
      h->dq[s][2] = (int16_t)(kVp8DcQ[clampi(q + dy2dc, 0, 127)] * 2);
      h->dq[s][3] = (int16_t)y2ac;
      h->dq[s][4] = kVp8DcQ[clampi(q + duvdc, 0, 117)];
      h->dq[s][5] = (int16_t)kVp8AcQ[clampi(q + duvac, 0, 127)];
    }
  }
  bd_bit(br, 0x80);   // update_proba: ignored on key frames
  // coefficient probabilities (tree_dec.c:515-538)
  for (int i = 0; i < 1056; ++i) {
    h->prob[i] = bd_bit(br, kVp8CoeffUpdateProba[i]) ? (uint8_t)bd_value(br, 8) : kVp8CoeffProba0[i];
  }
  h->use_skip = (uint8_t)bd_bit(br, 0x80);
  h->skip_p = h->use_skip ? (uint8_t)bd_value(br, 8) : 0;
  // loop-filter strengths per segment and block type (frame_dec.c:265-313)
  if (im.flags & VP8B_FLAG_BYPASS_FILTER) filter_type = 0;   // VP8EnterCritical, frame_dec.c:557-560
  h->filter_type = (uint8_t)filter_type;
  {   // rows below the crop window (plus what the loop filter reads past it) are never decoded, nor checked for eof
    const int extra = (filter_type == 2) ? 8 : (filter_type == 1) ? 2 : 0;   // kFilterExtraRows, frame_dec.c:201
    const int rows = ((int)im.crop_y + (int)im.out_h + 15 + extra) >> 4;
    h->rows = rows < (int)im.mb_h ? rows : (int)im.mb_h;
  }
  for (int s = 0; s < 4; ++s) {
    int base_level = level;
    if (use_segment) base_level = seg_filter[s] + (absolute_delta ? 0 : level);
    for (int i4 = 0; i4 <= 1; ++i4) {
      int lvl = base_level;
      if (use_lf_delta) lvl += ref_lf0 + (i4 ? mode_lf0 : 0);
      lvl = clampi(lvl, 0, 63);
      int limit = 0, ilevel = 0, hev = 0;
      if (lvl > 0) {
        ilevel = lvl;
        if (sharpness > 0) {
          ilevel >>= (sharpness > 4) ? 2 : 1;
          if (ilevel > 9 - sharpness) ilevel = 9 - sharpness;
        }
        if (ilevel < 1) ilevel = 1;
        limit = 2 * lvl + ilevel;
        hev = (lvl >= 40) ? 2 : (lvl >= 15) ? 1 : 0;
      }
      h->fstr[s][i4][0] = (uint8_t)limit; h->fstr[s][i4][1] = (uint8_t)ilevel;
      h->fstr[s][i4][2] = (uint8_t)i4;    h->fstr[s][i4][3] = (uint8_t)hev;
    }
  }
  return VP8B_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Intra modes of the whole frame (partition 0, strictly serial). `top` = mb_w words of scratch holding the
// four bottom sub-block modes of the macroblock row above (one byte each); `bprob` = kVp8BModeProba
// ([top mode][left mode][9]) staged in shared memory by the kernel (a pointer to the table itself elsewhere).
// Writes MbInfo x,y,w. Returns VP8B_OK or VP8B_NOT_ENOUGH_DATA (checked once per macroblock row, like
// vp8_dec.c:651-654). Sub-block mode tree of tree_dec.c:28-36 written out as straight-line decisions.
template <class BD>
VP8_FN uint32_t parse_bmode(BD& d, const uint8_t* p) {
  bd_fill(d);
  if (!bd_bit_nofill(d, p[0])) return M_DC;
  if (!bd_bit_nofill(d, p[1])) return M_TM;
  if (!bd_bit_nofill(d, p[2])) return M_VE;
  if (!bd_bit_nofill(d, p[3])) {
    bd_fill(d);
    if (!bd_bit_nofill(d, p[4])) return M_HE;
    return bd_bit_nofill(d, p[5]) ? M_VR : M_RD;
  }
  bd_fill(d);
  if (!bd_bit_nofill(d, p[6])) return M_LD;
  if (!bd_bit_nofill(d, p[7])) return M_VL;
  return bd_bit_nofill(d, p[8]) ? M_HU : M_HD;
}

template <class BD>
VP8_FN int parse_intra_modes(BD& br, const ImgDesc& im, const FrameHdr* h, uint32_t* top, const uint8_t* bprob,
                             uint32_t* mbinfo /* 4 words per MB */, int* fail_row = 0) {
  const int mb_w = im.mb_w, mb_h = h->rows;
  const int update_map = h->update_map, use_skip = h->use_skip, skip_p = h->skip_p;
  const uint32_t sp0 = h->seg_prob[0], sp1 = h->seg_prob[1], sp2 = h->seg_prob[2];
  for (int mx = 0; mx < mb_w; ++mx) top[mx] = 0;   // M_DC == 0
  for (int my = 0; my < mb_h; ++my) {
    uint32_t left = 0;   // four left modes, byte y
    for (int mx = 0; mx < mb_w; ++mx) {
      uint32_t* out = mbinfo + 4 * ((size_t)my * mb_w + mx);
      uint32_t w = 0, m0 = 0, m1 = 0;
      bd_fill(br);   // segment (<= 2) + skip + block size: at most 4 decodes
      if (update_map) {
        const uint32_t seg = !bd_bit_nofill(br, sp0) ? (uint32_t)bd_bit_nofill(br, sp1) : (uint32_t)bd_bit_nofill(br, sp2) + 2u;
        w |= seg << MBW_SEG_SHIFT;
      }
      if (use_skip && bd_bit_nofill(br, skip_p)) w |= MBW_SKIP;
      if (bd_bit_nofill(br, 145)) {   // i16
        bd_fill(br);
        const uint32_t ymode = bd_bit_nofill(br, 156) ? (bd_bit_nofill(br, 128) ? M_TM : M_HE)
                                                      : (bd_bit_nofill(br, 163) ? M_VE : M_DC);
        m0 = ymode;
        top[mx] = ymode * 0x01010101u;
        left = ymode * 0x01010101u;
      } else {
        // The modes above and to the left rotate through their words instead of being indexed: the one the next sub-block
        // needs is always the low byte, a new mode enters at the top (after four, a row's modes are back in column
        // order: the row above of the next row; after four rows `left` holds the right-most column in row order), and
        // the sixteen MbInfo nibbles shift down through m1:m0 the same way (the first one ends up lowest).
        uint32_t t = top[mx];
        w |= MBW_I4X4;
        VP8_UNROLL4
        for (int y = 0; y < 4; ++y) {
          uint32_t ymode = left & 0xff;
          VP8_UNROLL4
          for (int x = 0; x < 4; ++x) {
            ymode = parse_bmode(br, bprob + (t & 0xff) * 90 + ymode * 9);
            t = (t >> 8) | (ymode << 24);
            m0 = (m0 >> 4) | (m1 << 28);
            m1 = (m1 >> 4) | (ymode << 28);
          }
          left = (left >> 8) | (ymode << 24);
        }
        top[mx] = t;
      }
      bd_fill(br);
      const uint32_t uvmode = !bd_bit_nofill(br, 142) ? M_DC : !bd_bit_nofill(br, 114) ? M_VE : bd_bit_nofill(br, 183) ? M_TM : M_HE;
      w |= uvmode << MBW_UVMODE_SHIFT;
      uint4 o4; o4.x = m0; o4.y = m1; o4.z = 0; o4.w = w;
      *(uint4*)out = o4;
    }
    if (bd_eof(br)) { if (fail_row) *fail_row = my; return VP8B_NOT_ENOUGH_DATA; }
  }
  return VP8B_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Coefficient tokens. Window discipline: bd_fill() leaves > 32 valid bits and a decode uses at most 7, so up
// to four bd_*_nofill() calls may follow one bd_fill().
template <class BD>
VP8_FN int large_value(BD& d, const uint8_t* p) {   // GetLargeValue, vp8_dec.c:411-440; sign decoded by the caller
  int v;
  bd_fill(d);
  if (!bd_bit_nofill(d, p[3])) {
    v = !bd_bit_nofill(d, p[4]) ? 2 : 3 + bd_bit_nofill(d, p[5]);
  } else if (!bd_bit_nofill(d, p[6])) {
    bd_fill(d);
    if (!bd_bit_nofill(d, p[7])) {
      v = 5 + bd_bit_nofill(d, 159);
    } else {
      v = 7 + 2 * bd_bit_nofill(d, 165);
      v += bd_bit_nofill(d, 145);
    }
  } else {
    bd_fill(d);
    const int b1 = bd_bit_nofill(d, p[8]);
    const int b0 = bd_bit_nofill(d, p[9 + b1]);
    const int cat = 2 * b1 + b0;
    v = 0;
    for (const uint8_t* tab = kCatProb[cat]; *tab; ++tab) v += v + bd_bit(d, *tab);
    v += 3 + (8 << cat);
    bd_fill(d);
  }
  return v;
}

// Probabilities laid out by coefficient position instead of band, so that walking a block is pure pointer
// arithmetic: [type 4][position 17][ctx 3][11] = 4 x 561 bytes (position 16 is only ever addressed, never read).
#define VP8B_POSPROB_TYPE 561
#define VP8B_POSPROB_BYTES (4 * VP8B_POSPROB_TYPE)
VP8_TABLE uint8_t kBandOfPos[17] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0 };
// byte k of the expanded table (any thread subset may fill its slice)
VP8_FN uint8_t posprob_byte(const uint8_t* prob /* [4][8][3][11] */, int k) {
  const int t = k / VP8B_POSPROB_TYPE, r = k % VP8B_POSPROB_TYPE;
  return prob[t * 264 + kBandOfPos[r / 33] * 33 + r % 33];
}

// One 4x4 block. `tp` = this block type's [17 positions][3 ctx][11] table (shared memory), `out` = 16 int16 in
// HBM (pre-zeroed): the decoded LEVELS in parse (zigzag) order; dequantisation and the zigzag scatter happen
// where the coefficients are consumed (recon_macroblock). Returns nz = index of the last decoded coefficient
// + 1 (GetCoeffs, vp8_dec.c:443-469).
template <class BD>
VP8_FN int parse_block(BD& d, const uint8_t* tp, int ctx, int n, int16_t* out) {
  const uint8_t* pz = tp + n * 33;       // position n, ctx 0
  const uint8_t* p = pz + ctx * 11;
  for (;;) {
    bd_fill(d);
    if (!bd_bit_nofill(d, p[0])) break;
    while (!bd_bit_nofill(d, p[1])) {
      pz += 33;
      p = pz;
      if (++n == 16) return 16;
      bd_fill(d);
    }
    int v;
    if (!bd_bit_nofill(d, p[2])) {
      v = 1;
      p = pz + (33 + 11);
    } else {
      v = large_value(d, p);
      p = pz + (33 + 22);
    }
    if (bd_half_nofill(d)) v = -v;
    out[n] = (int16_t)v;
    pz += 33;
    if (++n == 16) break;
  }
  return n;
}

// Progress hand-off between the token partitions of one image: partition p publishes how many macroblocks
// it has finished (counted over all its rows); the partition owning the next row waits on it. No-ops for a
// single partition. Defined by the including translation unit.
#ifndef VP8_WAIT_PROGRESS
#define VP8_WAIT_PROGRESS(ptr, need) ((void)0)
#define VP8_PUBLISH_PROGRESS(ptr, val) ((void)0)
#endif

// Blocks of a macroblock in parse order (seq 0 = Y2, 1..16 luma, 17..24 chroma), one word each:
//  [3:0] bit of the top context  [7:4] bit of the left context  [12:8] shift of the 2-bit nz code
//  13 chroma  14 luma  [16:15] block type when the macroblock is i16 (luma of i4x4 macroblocks: type 3)
//  [18:17] dequantiser pair  [23:19] block index inside the macroblock's coefficients
#define SQ_(tb, lb, sh, fl, type, qp, blk) ((uint32_t)(tb) | ((uint32_t)(lb) << 4) | ((uint32_t)(sh) << 8) | (fl) | ((uint32_t)(type) << 15) | ((uint32_t)(qp) << 17) | ((uint32_t)(blk) << 19))
#define SQ_CHROMA (1u << 13)
#define SQ_LUMA (1u << 14)
#define SQ_Y(b) SQ_((b) & 3, (b) >> 2, 30 - 2 * (b), SQ_LUMA, 0, 0, (b))
#define SQ_C(c) SQ_(4 + ((c) & 1) + 2 * ((c) >> 2), 4 + (((c) >> 1) & 1) + 2 * ((c) >> 2), 8 * ((c) >> 2) + 6 - 2 * ((c) & 3), SQ_CHROMA, 2, 2, 16 + (c))
VP8_TABLE uint32_t kBlockSeq[25] = {
  SQ_(8, 8, 0, 0, 1, 1, 24),
  SQ_Y(0), SQ_Y(1), SQ_Y(2), SQ_Y(3), SQ_Y(4), SQ_Y(5), SQ_Y(6), SQ_Y(7),
  SQ_Y(8), SQ_Y(9), SQ_Y(10), SQ_Y(11), SQ_Y(12), SQ_Y(13), SQ_Y(14), SQ_Y(15),
  SQ_C(0), SQ_C(1), SQ_C(2), SQ_C(3), SQ_C(4), SQ_C(5), SQ_C(6), SQ_C(7)
};

// Non-zero context of one macroblock column / row: bits 0-3 luma, 4-5 U, 6-7 V, bit 8 = Y2 (nz_dc).
// State of one token partition of an image; it parses macroblock rows part, part+P, ... (vp8_dec.c:649-650).
template <class BD>
struct TokenPartT {
  BD d;
  int done;     // macroblocks finished by this partition (published to the partition owning the next row)
  int status;   // VP8B_OK or VP8B_NOT_ENOUGH_DATA
};
typedef TokenPartT<BoolDec> TokenPart;

template <class BD>
VP8_FN void token_part_init(TokenPartT<BD>& tp, const uint8_t* frame, const FrameHdr* h, int part) {
  bd_init(tp.d, frame + h->part_off[part], h->part_size[part]);
  tp.done = 0;
  tp.status = VP8B_OK;
}

// One macroblock row `my` of partition `part` (= my % P).
//   probs     : VP8B_POSPROB_BYTES, this frame's coefficient probabilities by position (shared memory)
//   topctx    : (P+1) rows x mb_w uint16 ring of per-column contexts (shared by the image's partitions)
//   progress  : P counters (shared); see above
// Writes coefficients and MbInfo z / w.
// SINK: what happens to a macroblock's levels once they are parsed. The default leaves them in the dense plane (800 bytes per
// macroblock at `coeffs`); a sink with kPerMb = 1 gets every macroblock in the same 400-level buffer `coeffs` (handed over
// all zero, to be left all zero) -- k_parse_literal turns it into the token stream there.
struct TokDensePlane {
  static const int kPerMb = 0;
  VP8_MFN void mb(int, int, int16_t*) {}
};
template <class BD, class SINK = TokDensePlane>
VP8_FN void parse_token_row(TokenPartT<BD>& tp, const ImgDesc& im, const FrameHdr* h, int part, int my, const uint8_t* probs,
                            uint16_t* topctx, volatile int* progress, uint32_t* mbinfo, int16_t* coeffs, SINK sink = SINK()) {
  const int P = h->num_parts, mb_w = im.mb_w;
  const int use_skip = h->use_skip;
  BD& d = tp.d;
  const uint16_t* trow = topctx + (size_t)((my + P) % (P + 1)) * mb_w;   // written by row my-1
  uint16_t* orow = topctx + (size_t)(my % (P + 1)) * mb_w;
  const int prev = (part + P - 1) % P;                // partition that owns row my-1
  const int prev_rows = (my - 1 - prev) / P;          // rows it finished before row my-1 (meaningful if my>0)
  uint32_t lctx = 0;
  uint32_t* info = mbinfo + 4 * ((size_t)my * mb_w);
  int16_t* dst = SINK::kPerMb ? coeffs : coeffs + (size_t)my * mb_w * VP8B_COEFFS_PER_MB;
  uint32_t w_next = info[3];
  for (int mx = 0; mx < mb_w; ++mx, info += 4, dst += SINK::kPerMb ? 0 : VP8B_COEFFS_PER_MB) {
    uint32_t w = w_next;
    if (mx + 1 < mb_w) w_next = info[7];   // next macroblock's flags: fetched a whole macroblock ahead of their use
    uint32_t tctx = 0;
    if (my > 0) {
      if (P > 1) VP8_WAIT_PROGRESS(&progress[prev], prev_rows * mb_w + mx + 1);
      tctx = trow[mx];
    }
    uint32_t nzy = 0, nzuv = 0;
    const uint32_t is_i4 = (w >> 16) & 1u;   // MBW_I4X4
    if (!(use_skip && (w & MBW_SKIP))) {
      // Contexts as shift registers, like ParseResiduals (vp8_dec.c:517-609): the bit of the next block sits in bit 0.
      const uint8_t* yprobs = probs + 3 * VP8B_POSPROB_TYPE;   // luma with DC (i4x4 macroblocks)
      int first = 0;
      if (!is_i4) {   // Y2 block: the 16 luma DCs; the luma blocks then start at coefficient 1 with the type-0 tables
        const int ctx = (int)(((tctx >> 8) & 1u) + ((lctx >> 8) & 1u));
        const int nz = parse_block(d, probs + 1 * VP8B_POSPROB_TYPE, ctx, 0, dst + 24 * 16);
        const uint32_t f = (nz > 0) ? 0x100u : 0u;
        tctx = (tctx & 0xffu) | f;
        lctx = (lctx & 0xffu) | f;
        if (nz > 0) w |= MBW_HAS_Y2;
        first = 1;
        yprobs = probs;
      }
      int16_t* out = dst;
      uint32_t tnz = tctx & 0x0f, lnz = lctx & 0x0f;
      for (int y = 0; y < 4; ++y) {
        uint32_t l = lnz & 1;
        for (int x = 0; x < 4; ++x) {
          const int nz = parse_block(d, yprobs, (int)(l + (tnz & 1)), first, out);
          l = (nz > first);
          tnz = (tnz >> 1) | (l << 7);
          // nz code; a lone DC level counts as "DC only" here and is re-examined after dequantisation (recon_macroblock)
          nzy = (nzy << 2) | (uint32_t)((nz > 3) ? 3 : (nz > 1) ? 2 : (int)l);
          out += 16;
        }
        tnz >>= 4;
        lnz = (lnz >> 1) | (l << 7);
      }
      uint32_t out_t = tnz, out_l = lnz >> 4;
      for (int ch = 0; ch < 4; ch += 2) {   // U then V: 2x2 blocks each
        uint32_t acc = 0;
        tnz = (tctx >> (4 + ch)) & 0x0f;
        lnz = (lctx >> (4 + ch)) & 0x0f;
        for (int y = 0; y < 2; ++y) {
          uint32_t l = lnz & 1;
          for (int x = 0; x < 2; ++x) {
            const int nz = parse_block(d, probs + 2 * VP8B_POSPROB_TYPE, (int)(l + (tnz & 1)), 0, out);
            l = (nz > 0);
            tnz = ((tnz >> 1) | (l << 3)) & 0xff;
            acc = (acc << 2) | (uint32_t)((nz > 3) ? 3 : (nz > 1) ? 2 : (int)l);
            out += 16;
          }
          tnz >>= 2;
          lnz = ((lnz >> 1) | (l << 5)) & 0xff;
        }
        nzuv |= acc << (4 * ch);
        out_t |= ((tnz << 4) << ch) & 0xff;
        out_l |= ((lnz & 0xf0) << ch) & 0xff;
      }
      tctx = (tctx & 0x100u) | (out_t & 0xff);
      lctx = (lctx & 0x100u) | (out_l & 0xff);
    } else {
      tctx &= is_i4 ? 0x100u : 0u;
      lctx &= is_i4 ? 0x100u : 0u;
    }
    info[2] = nzy;
    info[3] = (w & 0xffff0000u) | nzuv;
    sink.mb(mx, my, dst);
    orow[mx] = (uint16_t)tctx;
    ++tp.done;
    if (P > 1) VP8_PUBLISH_PROGRESS(&progress[part], tp.done);
    if (bd_eof(d)) tp.status = VP8B_NOT_ENOUGH_DATA;   // keep going: later partitions wait on our progress
  }
}

#endif  // LIBWEBP_B200_VP8_PARSE_CORE_H_
