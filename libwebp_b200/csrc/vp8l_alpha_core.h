// vp8l_alpha_core.h -- ALPH chunk decoding on the device: the VP8L subset that carries the alpha plane of a lossy
// WebP, the row unfilters, and nothing else of the lossless codec.
//
// Replaces, for a whole batch at once:
//   ALPHInit / ALPHDecode / VP8DecompressAlphaRows                      src/dec/alpha_dec.c:48-233
//   VP8LDecodeAlphaHeader / DecodeImageStream / ReadTransform /
//   ReadHuffmanCodes(+Helper) / ReadHuffmanCode / ReadHuffmanCodeLengths /
//   DecodeAlphaData / DecodeImageData / ExtractPalettedAlphaRows       src/dec/vp8l_dec.c:159-1700
//   VP8LBitReader                                                       src/utils/bit_reader_utils.c:141-222, .h:134-189
//   BuildHuffmanTable                                                   src/utils/huffman_utils.c:43-209
//   VP8LColorIndexInverseTransformAlpha                                 src/dsp/lossless.c:341-385
//   WebPUnfilters (horizontal / vertical / gradient)                    src/dsp/filters.c:121-234
//
// One thread walks one image's entropy-coded stream (Huffman + LZ77 is as serial as the boolean decoder). The
// work is split in two passes so that the host can size the per-image memory in between:
//   pass A  alph_parse_header : ALPH header byte, transforms (palette sub-image), colour-cache flag, meta-Huffman
//                               image -> AlphaHdr (+ the reader's state)
//   pass B  alph_decode_pixels: every group's five prefix codes -> lookup tables, then the pixel loop -> one ARGB
//                               word per coded pixel
//   then    alph_finish       : inverse transforms in place (predictor as a lag-2 row wavefront), palette /
//                               unbundling, green -> alpha, and the row unfilter -> the w x h alpha plane
// The whole VP8L syntax is accepted: all four transforms in any order (a COLOR_INDEXING transform that is not the first one
// widens the rows in place, al_inverse_transforms), colour cache, meta-Huffman groups up to the 65536 the syntax can name
// (tables only for the ones in use when the reference would have remapped them, alph_parse_header).
// Status mirrors the reference: a failure while reading the headers/codes surfaces as VP8_STATUS_OUT_OF_MEMORY
// (alpha_dec.c:190-196: the lossless decoder object was never attached), a failure in the pixel loop as
// VP8_STATUS_BITSTREAM_ERROR (frame_dec.c:452-460).
// Dual build like the other cores: nvcc for the product, g++ -DVP8_EMU for tests/emu.
#ifndef LIBWEBP_B200_VP8L_ALPHA_CORE_H_
#define LIBWEBP_B200_VP8L_ALPHA_CORE_H_

#include <stdint.h>

#if defined(__CUDACC__) && !defined(VP8_EMU)
#define AL_FN __device__ __forceinline__
#define AL_NOINLINE static __device__ __noinline__
#define AL_TABLE static __constant__ const
#define AL_POPC(x) __popc(x)
#else
#define AL_FN static inline
#define AL_NOINLINE static
#define AL_POPC(x) __builtin_popcount(x)
#define AL_TABLE static const
#endif

#define AL_OK 0
#define AL_OUT_OF_MEMORY 1
#define AL_BITSTREAM_ERROR 3
#define AL_UNSUPPORTED 4

#define AL_NUM_LITERAL 256
#define AL_NUM_LENGTH 24
#define AL_NUM_DISTANCE 40
#define AL_MAX_CODE_LEN 15
#define AL_ROOT_BITS 8
#define AL_LENGTHS_ROOT_BITS 7
#define AL_MAX_CACHE_BITS 11
#define AL_MAX_ALPHABET (AL_NUM_LITERAL + AL_NUM_LENGTH + (1 << AL_MAX_CACHE_BITS))   // 2328
#define AL_MAX_GROUPS 65536    // group numbers are 16 bits wide (the reference encoder never emits more than 2600, MAX_HUFF_IMAGE_SIZE)
// lookup-table entries of one group: 630 * 3 + 410 + the green table, whose worst case grows with the colour cache
// (kTableSize, vp8l_dec.c:81-96)
#define AL_FIXED_TABLE_ENTRIES (630 * 3 + 410)
AL_TABLE uint16_t kAlGreenTableSize[12] = { 654, 656, 658, 662, 670, 686, 718, 782, 912, 1168, 1680, 2704 };
#define AL_GROUP_ENTRIES(cache_bits) (AL_FIXED_TABLE_ENTRIES + (int)kAlGreenTableSize[cache_bits])
#define AL_SUB_TABLE_ENTRIES (AL_FIXED_TABLE_ENTRIES + 2704)   // a group with an 11-bit colour cache

AL_TABLE uint8_t kAlCodeLengthOrder[19] = { 17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15 };
AL_TABLE uint8_t kAlCodeToPlane[120] = {
  0x18, 0x07, 0x17, 0x19, 0x28, 0x06, 0x27, 0x29, 0x16, 0x1a, 0x26, 0x2a, 0x38, 0x05, 0x37, 0x39, 0x15, 0x1b, 0x36, 0x3a,
  0x25, 0x2b, 0x48, 0x04, 0x47, 0x49, 0x14, 0x1c, 0x35, 0x3b, 0x46, 0x4a, 0x24, 0x2c, 0x58, 0x45, 0x4b, 0x34, 0x3c, 0x03,
  0x57, 0x59, 0x13, 0x1d, 0x56, 0x5a, 0x23, 0x2d, 0x44, 0x4c, 0x55, 0x5b, 0x33, 0x3d, 0x68, 0x02, 0x67, 0x69, 0x12, 0x1e,
  0x66, 0x6a, 0x22, 0x2e, 0x54, 0x5c, 0x43, 0x4d, 0x65, 0x6b, 0x32, 0x3e, 0x78, 0x01, 0x77, 0x79, 0x53, 0x5d, 0x11, 0x1f,
  0x64, 0x6c, 0x42, 0x4e, 0x76, 0x7a, 0x21, 0x2f, 0x75, 0x7b, 0x31, 0x3f, 0x63, 0x6d, 0x52, 0x5e, 0x00, 0x74, 0x7c, 0x41,
  0x4f, 0x10, 0x20, 0x62, 0x6e, 0x30, 0x73, 0x7d, 0x51, 0x5f, 0x40, 0x72, 0x7e, 0x61, 0x6f, 0x50, 0x71, 0x7f, 0x60, 0x70
};

// ---------------------------------------------------------------------------------------------------------
// Bit reader: LSB first. State and end-of-stream rule follow the reference's VP8LBitReader literally (64-bit
// window `val`, `bit_pos` bits of it consumed, `pos` bytes of the buffer loaded), because whether a damaged
// stream decodes or fails is decided by exactly when `eos` rises.
struct LBits {
  uint64_t val;
  const uint8_t* buf;
  uint32_t len, pos;
  int bit_pos, eos;
};

AL_FN void lb_init(LBits& b, const uint8_t* buf, uint32_t len) {
  b.buf = buf; b.len = len; b.val = 0; b.bit_pos = 0; b.eos = 0;
  const uint32_t n = len > 8 ? 8 : len;
  for (uint32_t i = 0; i < n; ++i) b.val |= (uint64_t)buf[i] << (8 * i);
  b.pos = n;
}
AL_FN int lb_at_end(const LBits& b) { return b.eos || (b.pos == b.len && b.bit_pos > 64); }
AL_FN void lb_set_eos(LBits& b) { b.eos = 1; b.bit_pos = 0; }
AL_FN void lb_shift_bytes(LBits& b) {
  while (b.bit_pos >= 8 && b.pos < b.len) {
    b.val = (b.val >> 8) | ((uint64_t)b.buf[b.pos] << 56);
    ++b.pos;
    b.bit_pos -= 8;
  }
  if (lb_at_end(b)) lb_set_eos(b);
}
AL_FN void lb_fill(LBits& b) {   // VP8LFillBitWindow
  if (b.bit_pos >= 32) {
    if (b.pos + 8 < b.len) {
      const uint8_t* p = b.buf + b.pos;
      const uint64_t w = (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24);
      b.val = (b.val >> 32) | (w << 32);
      b.bit_pos -= 32;
      b.pos += 4;
    } else {
      lb_shift_bytes(b);
    }
  }
}
AL_FN uint32_t lb_peek(const LBits& b) { return (uint32_t)(b.val >> (b.bit_pos & 63)); }
AL_FN uint32_t lb_read(LBits& b, int n) {   // VP8LReadBits, n <= 24
  if (!b.eos) {
    const uint32_t v = lb_peek(b) & ((1u << n) - 1u);
    b.bit_pos += n;
    lb_shift_bytes(b);
    return v;
  }
  lb_set_eos(b);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Prefix-code lookup tables: entry = bits | value << 16; 8-bit root + second level (huffman_utils.c:80-209).
AL_FN uint32_t hc_next_key(uint32_t key, int len) {
  uint32_t step = 1u << (len - 1);
  while (key & step) step >>= 1;
  return step ? (key & (step - 1)) + step : key;
}
AL_FN void hc_replicate(uint32_t* table, int step, int end, uint32_t code) {
  do { end -= step; table[end] = code; } while (end > 0);
}
// code_lengths[n] (each <= 15) -> table (or NULL to validate only). `sorted` = n uint16 of scratch.
// Returns the number of entries used, 0 if the lengths do not describe a complete prefix code.
AL_NOINLINE int hc_build(uint32_t* root_table, int root_bits, const uint8_t* code_lengths, int n, uint16_t* sorted) {
  uint32_t* table = root_table;
  int total_size = 1 << root_bits;
  int count[AL_MAX_CODE_LEN + 1], offset[AL_MAX_CODE_LEN + 1];
  for (int i = 0; i <= AL_MAX_CODE_LEN; ++i) count[i] = 0;
  for (int s = 0; s < n; ++s) {
    if (code_lengths[s] > AL_MAX_CODE_LEN) return 0;
    ++count[code_lengths[s]];
  }
  if (count[0] == n) return 0;
  offset[1] = 0;
  for (int len = 1; len < AL_MAX_CODE_LEN; ++len) {
    if (count[len] > (1 << len)) return 0;
    offset[len + 1] = offset[len] + count[len];
  }
  for (int s = 0; s < n; ++s) {
    const int l = code_lengths[s];
    if (l > 0) sorted[offset[l]++] = (uint16_t)s;
  }
  if (offset[AL_MAX_CODE_LEN] == 1) {   // a single symbol: zero-bit code
    if (root_table != 0) hc_replicate(table, 1, total_size, (uint32_t)sorted[0] << 16);
    return total_size;
  }
  uint32_t low = 0xffffffffu, mask = (uint32_t)total_size - 1, key = 0;
  int num_nodes = 1, num_open = 1, table_bits = root_bits, table_size = 1 << table_bits, symbol = 0, step = 2;
  for (int len = 1; len <= root_bits; ++len, step <<= 1) {
    num_open <<= 1; num_nodes += num_open; num_open -= count[len];
    if (num_open < 0) return 0;
    for (; count[len] > 0; --count[len]) {
      if (root_table != 0) hc_replicate(&table[key], step, table_size, (uint32_t)len | ((uint32_t)sorted[symbol] << 16));
      ++symbol;
      key = hc_next_key(key, len);
    }
  }
  step = 2;
  for (int len = root_bits + 1; len <= AL_MAX_CODE_LEN; ++len, step <<= 1) {
    num_open <<= 1; num_nodes += num_open; num_open -= count[len];
    if (num_open < 0) return 0;
    for (; count[len] > 0; --count[len]) {
      if ((key & mask) != low) {
        if (root_table != 0) table += table_size;
        int left = 1 << (len - root_bits), l2 = len;   // NextTableBitSize
        while (l2 < AL_MAX_CODE_LEN) { left -= count[l2]; if (left <= 0) break; ++l2; left <<= 1; }
        table_bits = l2 - root_bits;
        table_size = 1 << table_bits;
        total_size += table_size;
        low = key & mask;
        if (root_table != 0) root_table[low] = (uint32_t)(table_bits + root_bits) | ((uint32_t)((table - root_table) - low) << 16);
      }
      if (root_table != 0) hc_replicate(&table[key >> root_bits], step, table_size, (uint32_t)(len - root_bits) | ((uint32_t)sorted[symbol] << 16));
      ++symbol;
      key = hc_next_key(key, len);
    }
  }
  if (num_nodes != 2 * offset[AL_MAX_CODE_LEN] - 1) return 0;
  return total_size;
}

AL_FN int hc_read_symbol(const uint32_t* table, LBits& b) {   // ReadSymbol, vp8l_dec.c:192-206
  uint32_t val = lb_peek(b);
  table += val & 0xff;
  const int nbits = (int)(*table & 0xff) - AL_ROOT_BITS;
  if (nbits > 0) {
    b.bit_pos += AL_ROOT_BITS;
    val = lb_peek(b);
    table += *table >> 16;
    table += val & ((1u << nbits) - 1u);
  }
  b.bit_pos += (int)(*table & 0xff);
  return (int)(*table >> 16);
}

// Per-thread scratch for reading one prefix code.
struct AlScratch {
  uint8_t code_lengths[AL_MAX_ALPHABET];
  uint16_t sorted[AL_MAX_ALPHABET];
  uint32_t lengths_table[1 << AL_LENGTHS_ROOT_BITS];
};

// ReadHuffmanCode (vp8l_dec.c:319-363): returns the table size, 0 on error. table may be NULL (validate only).
AL_NOINLINE int al_read_code(LBits& b, int alphabet_size, uint32_t* table, AlScratch* sc) {
  int ok = 0;
  const int simple = (int)lb_read(b, 1);
  for (int i = 0; i < alphabet_size; ++i) sc->code_lengths[i] = 0;
  if (simple) {
    const int num_symbols = (int)lb_read(b, 1) + 1;
    const int first_len_code = (int)lb_read(b, 1);
    int symbol = (int)lb_read(b, first_len_code == 0 ? 1 : 8);
    sc->code_lengths[symbol] = 1;
    if (num_symbols == 2) { symbol = (int)lb_read(b, 8); sc->code_lengths[symbol] = 1; }
    ok = 1;
  } else {
    uint8_t clcl[19];
    for (int i = 0; i < 19; ++i) clcl[i] = 0;
    const int num_codes = (int)lb_read(b, 4) + 4;
    for (int i = 0; i < num_codes; ++i) clcl[kAlCodeLengthOrder[i]] = (uint8_t)lb_read(b, 3);
    // ReadHuffmanCodeLengths (vp8l_dec.c:257-317)
    if (hc_build(sc->lengths_table, AL_LENGTHS_ROOT_BITS, clcl, 19, sc->sorted)) {
      int max_symbol, bad = 0;
      if (lb_read(b, 1)) {
        const int length_nbits = 2 + 2 * (int)lb_read(b, 3);
        max_symbol = 2 + (int)lb_read(b, length_nbits);
        if (max_symbol > alphabet_size) bad = 1;
      } else {
        max_symbol = alphabet_size;
      }
      if (!bad) {
        int symbol = 0, prev_code_len = 8;
        while (symbol < alphabet_size) {
          if (max_symbol-- == 0) break;
          lb_fill(b);
          const uint32_t e = sc->lengths_table[lb_peek(b) & ((1u << AL_LENGTHS_ROOT_BITS) - 1u)];
          b.bit_pos += (int)(e & 0xff);
          const int code_len = (int)(e >> 16);
          if (code_len < 16) {
            sc->code_lengths[symbol++] = (uint8_t)code_len;
            if (code_len != 0) prev_code_len = code_len;
          } else {
            const int slot = code_len - 16;
            const int extra_bits = (slot == 0) ? 2 : (slot == 1) ? 3 : 7;
            const int repeat_offset = (slot == 2) ? 11 : 3;
            int repeat = (int)lb_read(b, extra_bits) + repeat_offset;
            if (symbol + repeat > alphabet_size) { bad = 1; break; }
            const int length = (code_len == 16) ? prev_code_len : 0;
            while (repeat-- > 0) sc->code_lengths[symbol++] = (uint8_t)length;
          }
        }
        ok = !bad;
      }
    }
  }
  ok = ok && !b.eos;
  if (!ok) return 0;
  return hc_build(table, AL_ROOT_BITS, sc->code_lengths, alphabet_size, sc->sorted);
}

// Five tables of one group: offsets (in entries) from the group's arena base.
struct AlGroup {
  uint32_t off[5];      // GREEN, RED, BLUE, ALPHA, DIST
  uint32_t literal_arb; // alpha << 24 | red << 16 | blue when R, B and A are zero-bit codes
  uint8_t trivial_literal, trivial_code, pad[2];
};

// One group's codes (ReadHuffmanCodesHelper's inner loop, vp8l_dec.c:497-545). `arena`/`cap` = lookup-table
// memory still free; returns entries consumed, 0 on error. g may be NULL with arena NULL (unused group).
AL_NOINLINE int al_read_group(LBits& b, int cache_bits, uint32_t* arena, int cap, AlGroup* g, AlScratch* sc) {
  int used = 0, total_bits = 0, trivial_literal = 1;
  for (int j = 0; j < 5; ++j) {
    int alphabet = (j == 0) ? AL_NUM_LITERAL + AL_NUM_LENGTH + (cache_bits > 0 ? (1 << cache_bits) : 0)
                 : (j == 4) ? AL_NUM_DISTANCE : AL_NUM_LITERAL;
    // worst-case size of this table must fit before building it
    const int worst = (j == 0) ? (int)kAlGreenTableSize[cache_bits] : (j == 4) ? 410 : 630;
    if (arena != 0 && used + worst > cap) return 0;
    uint32_t* t = (arena != 0) ? arena + used : 0;
    const int size = al_read_code(b, alphabet, t, sc);
    if (size == 0) return 0;
    if (g != 0) {
      g->off[j] = (uint32_t)used;
      const int bits0 = (int)(t[0] & 0xff);
      if (j >= 1 && j <= 3 && trivial_literal) trivial_literal = (bits0 == 0);
      total_bits += bits0;
    }
    used += size;
  }
  if (g != 0) {
    g->trivial_literal = (uint8_t)trivial_literal;
    g->trivial_code = 0;
    g->literal_arb = 0;
    if (trivial_literal) {
      const uint32_t red = arena[g->off[1]] >> 16, blue = arena[g->off[2]] >> 16, alpha = arena[g->off[3]] >> 16;
      g->literal_arb = (alpha << 24) | (red << 16) | blue;
      if (total_bits == 0 && (arena[g->off[0]] >> 16) < AL_NUM_LITERAL) {
        g->trivial_code = 1;
        g->literal_arb |= (arena[g->off[0]] >> 16) << 8;
      }
    }
  }
  return used;
}

AL_FN int al_copy_value(int symbol, LBits& b) {   // GetCopyDistance / GetCopyLength, vp8l_dec.c:159-174
  if (symbol < 4) return symbol + 1;
  const int extra_bits = (symbol - 2) >> 1;
  const int offset = (2 + (symbol & 1)) << extra_bits;
  return offset + (int)lb_read(b, extra_bits) + 1;
}
AL_FN int al_plane_to_distance(int xsize, int plane_code) {   // vp8l_dec.c:176-186
  if (plane_code > 120) return plane_code - 120;
  const int dist_code = kAlCodeToPlane[plane_code - 1];
  const int yoffset = dist_code >> 4, xoffset = 8 - (dist_code & 0xf);
  const int dist = yoffset * xsize + xoffset;
  return dist >= 1 ? dist : 1;
}

// ---------------------------------------------------------------------------------------------------------
// Sub-image (level > 0 of DecodeImageStream, vp8l_dec.c:1455-1537 + DecodeImageData :1138-1293): colour-cache
// flag, one group of codes, xsize * ysize ARGB pixels into `out`. Scratch: `tables` (AL_SUB_TABLE_ENTRIES) and
// `cache` (1 << 11 words). Returns 1 on success.
AL_NOINLINE int al_decode_subimage(LBits& b, int xsize, int ysize, uint32_t* out, uint32_t* tables, uint32_t* cache, AlScratch* sc) {
  int cache_bits = 0;
  if (lb_read(b, 1)) {
    cache_bits = (int)lb_read(b, 4);
    if (cache_bits < 1 || cache_bits > AL_MAX_CACHE_BITS) return 0;
  }
  if (b.eos) return 0;
  AlGroup g;
  if (al_read_group(b, cache_bits, tables, AL_SUB_TABLE_ENTRIES, &g, sc) == 0) return 0;
  const int cache_size = cache_bits > 0 ? (1 << cache_bits) : 0;
  const int cache_shift = 32 - cache_bits;
  for (int i = 0; i < cache_size; ++i) cache[i] = 0;
  const int total = xsize * ysize;
  const int len_code_limit = AL_NUM_LITERAL + AL_NUM_LENGTH;
  int pos = 0;
  while (pos < total) {
    uint32_t px;
    if (g.trivial_code) {
      px = g.literal_arb;
    } else {
      lb_fill(b);
      const int code = hc_read_symbol(tables + g.off[0], b);
      if (lb_at_end(b)) break;
      if (code < AL_NUM_LITERAL) {
        if (g.trivial_literal) {
          px = g.literal_arb | ((uint32_t)code << 8);
        } else {
          const uint32_t red = (uint32_t)hc_read_symbol(tables + g.off[1], b);
          lb_fill(b);
          const uint32_t blue = (uint32_t)hc_read_symbol(tables + g.off[2], b);
          const uint32_t alpha = (uint32_t)hc_read_symbol(tables + g.off[3], b);
          if (lb_at_end(b)) break;
          px = (alpha << 24) | (red << 16) | ((uint32_t)code << 8) | blue;
        }
      } else if (code < len_code_limit) {
        const int length = al_copy_value(code - AL_NUM_LITERAL, b);
        const int dist_symbol = hc_read_symbol(tables + g.off[4], b);
        lb_fill(b);
        const int dist = al_plane_to_distance(xsize, al_copy_value(dist_symbol, b));
        if (lb_at_end(b)) break;
        if (pos < dist || total - pos < length) return 0;
        for (int k = 0; k < length; ++k) {
          const uint32_t v = out[pos - dist];
          out[pos++] = v;
          if (cache_size) cache[(v * 0x1e35a7bdu) >> cache_shift] = v;
        }
        continue;
      } else if (code < len_code_limit + cache_size) {
        px = cache[code - len_code_limit];
      } else {
        return 0;
      }
    }
    out[pos++] = px;
    if (cache_size) cache[(px * 0x1e35a7bdu) >> cache_shift] = px;
  }
  b.eos = lb_at_end(b);
  return !b.eos && pos >= total;
}

AL_FN int al_gradient_predict(int a, int b, int c) {   // GradientPredictor_C, filters.c:121-125
  const int g = a + b - c;
  return (g & ~0xff) == 0 ? g : (g < 0) ? 0 : 255;
}

// ---------------------------------------------------------------------------------------------------------
// What pass A leaves for the host and for pass B.
#define AL_T_PREDICTOR 0
#define AL_T_CROSS_COLOR 1
#define AL_T_SUBTRACT_GREEN 2
#define AL_T_COLOR_INDEXING 3
struct AlphaHdr {
  int32_t status;           // AL_OK or the image's failure
  uint8_t method;           // 0 raw, 1 lossless
  uint8_t filter;           // 0 none, 1 horizontal, 2 vertical, 3 gradient
  uint8_t ntrans;           // transforms in reading order
  uint8_t cache_bits;       // colour cache of the main image, 0 = none
  uint8_t huff_bits;        // meta-Huffman precision, 0 = one group
  uint8_t ttype[4];         // transform types
  uint8_t tbits[4];         // tile bits (predictor, cross colour) or bundling bits (colour indexing)
  uint8_t use_8b;           // set by the pixel pass: the reference would have taken its 8-bit path (DecodeAlphaData)
  uint8_t lossless;         // the stream is a whole VP8L picture (5-byte header instead of the ALPH byte), see vp8l_lossless_core.h
  uint8_t levels;           // ALPH header pre-processing = 1 (quantised levels): the reference then decodes every row at the first
                            // request (alpha_dec.c:196-203), whether or not alpha dithering was asked for
  int32_t fail_row;         // status != AL_OK: the alpha row whose decode fails (the reference decodes rows as they are asked for,
                            // alpha_dec.c:110-135), -1 = the header stage
  int32_t txsize[4];        // image width the transform applies to
  uint32_t tdata[4];        // word offset of the transform's tile image inside the transform-data area
  int32_t xsize;            // width of the coded image (after bundling)
  int32_t px_stride;        // words per row of the pixel buffer once the inverse transforms are done: xsize, or the picture's
                            // width when a palette transform that is not the first one has widened the rows on the way
  int32_t huff_xsize;       // width of the meta-Huffman image
  int32_t num_groups;       // groups whose codes follow in the stream
  int32_t used_groups;      // groups that get tables: num_groups, or only the ones the meta-Huffman image names when the
                            // reference would have remapped them (`mapped`, vp8l_dec.c:399-424)
  int32_t mapped;
  int32_t group_entries;    // lookup-table entries reserved per group
  // reader state after the meta-Huffman image
  uint64_t br_val;
  uint32_t br_pos;
  int32_t br_bit_pos;
  uint8_t palette_alpha[256];   // alpha (= green) of palette entry i, zero beyond the coded colours
  uint32_t palette[256];        // the expanded colour map (ExpandColorMap, vp8l_dec.c:1305-1328), zero beyond the coded colours
};

#define AL_PASSA_FIXED_WORDS (AL_SUB_TABLE_ENTRIES + (1 << AL_MAX_CACHE_BITS) + 256)
// bytes of per-image scratch: sub-image tables + colour cache + palette pixels + AlScratch (rounded up to 16)
#define AL_SCRATCH_FIXED_BYTES (4u * AL_PASSA_FIXED_WORDS + ((uint32_t)sizeof(AlScratch) + 15u) / 16u * 16u)
// ... + which group numbers the meta-Huffman image uses: one bit each, and the count of used numbers below every 32
#define AL_USED_BITS_WORDS (AL_MAX_GROUPS / 32)
#define AL_COPYQ 256   // queued backward references per flush (alph_decode_pixels_warp): 2 + 3 * AL_COPYQ words
#define AL_COPYQ_OFFSET (AL_SCRATCH_FIXED_BYTES + 4u * AL_USED_BITS_WORDS + 2u * AL_USED_BITS_WORDS)
#define AL_SCRATCH_BYTES (AL_COPYQ_OFFSET + 4u * (2u + 3u * AL_COPYQ) + 8u)
// (the queue lives in the image's scratch, not in shared memory: a kernel that declares shared memory gives up that much
// L1, and the one-lane pixel loop of the pictures that do not queue lives on L1 hits -- whole-picture VP8L 1125 -> 1375 ms
// with 3 KB of shared memory per block, profiles/r02y)
#define AL_COPYQ_PTR(scratch) ((uint32_t*)((uint8_t*)(scratch) + AL_COPYQ_OFFSET))
// dense index of group g among the used ones
#define AL_USED_BITS(scratch) ((uint32_t*)((uint8_t*)(scratch) + AL_SCRATCH_FIXED_BYTES))
#define AL_USED_BELOW(scratch) ((uint16_t*)(AL_USED_BITS(scratch) + AL_USED_BITS_WORDS))
// upper bound of a sub-sampled image (meta-Huffman, predictor or cross-colour tiles; precision >= 2 bits) of a
// w x h picture, in pixels
#define AL_META_PIXELS_BOUND(w, h) ((uint32_t)(((w) + 3) >> 2) * (uint32_t)(((h) + 3) >> 2))

// Pass A. alph = the ALPH chunk payload. scratch = AL_SCRATCH_BYTES bytes, 16-byte aligned.
// meta = 4 * AL_META_PIXELS_BOUND(w, h) bytes: decoded as 32-bit pixels, compacted in place to one uint16 group
// index per meta pixel (kept for pass B). tdata = 2 * AL_META_PIXELS_BOUND(w, h) words for the tile images of the
// predictor and cross-colour transforms (kept for alph_finish).
// lossless != 0: `alph` is a whole VP8L picture (VP8LDecodeHeader, vp8l_dec.c:1672-1704): the reader starts at the
// signature byte and ReadImageInfo (:115-124) takes 8 + 14 + 14 + 1 + 3 bits before the same image stream follows.
AL_NOINLINE void alph_parse_header(const uint8_t* alph, uint32_t alph_size, int w, int h, uint8_t* scratch, uint16_t* meta,
                                   uint32_t* tdata, AlphaHdr* hd, int lossless = 0) {
  hd->status = AL_OUT_OF_MEMORY;   // every failure in here is a header failure (see the top of this file)
  hd->fail_row = -1;
  hd->method = 0; hd->filter = 0; hd->ntrans = 0; hd->cache_bits = 0; hd->huff_bits = 0; hd->use_8b = 0;
  hd->lossless = (uint8_t)(lossless != 0);
  hd->levels = 0;
  hd->xsize = w; hd->px_stride = w; hd->huff_xsize = 0; hd->num_groups = 1; hd->used_groups = 1; hd->mapped = 0; hd->group_entries = AL_GROUP_ENTRIES(0);
  hd->br_val = 0; hd->br_pos = 0; hd->br_bit_pos = 0;
  uint32_t* tables = (uint32_t*)scratch;
  uint32_t* cache = tables + AL_SUB_TABLE_ENTRIES;
  uint32_t* pal = cache + (1 << AL_MAX_CACHE_BITS);
  AlScratch* sc = (AlScratch*)(pal + 256);
  LBits b;
  if (!lossless) {
    if (alph_size <= 1) return;
    const int method = alph[0] & 3, filter = (alph[0] >> 2) & 3, pre = (alph[0] >> 4) & 3, rsrv = (alph[0] >> 6) & 3;
    if (method > 1 || pre > 1 || rsrv != 0) return;
    hd->method = (uint8_t)method; hd->filter = (uint8_t)filter; hd->levels = (uint8_t)pre;
    if (method == 0) {
      if ((uint64_t)(alph_size - 1) >= (uint64_t)w * (uint64_t)h) hd->status = AL_OK;
      return;
    }
    lb_init(b, alph + 1, alph_size - 1);
  } else {
    hd->method = 1;
    lb_init(b, alph, alph_size);
    if (lb_read(b, 8) != 0x2f) return;
    const int sw = (int)lb_read(b, 14) + 1, sh = (int)lb_read(b, 14) + 1;
    lb_read(b, 1);   // alpha hint
    if (lb_read(b, 3) != 0 || b.eos) return;
    if (sw != w || sh != h) return;   // the container walk took the same fields (vp8_container.c)
  }
  // transforms (ReadTransform, vp8l_dec.c:1330-1384); each type at most once
  int xsize = w;
  uint32_t seen = 0, tdata_used = 0;
  int px_stride = 0;
  while (lb_read(b, 1)) {
    const uint32_t type = lb_read(b, 2);
    if (seen & (1u << type)) return;
    seen |= 1u << type;
    const int n = hd->ntrans++;
    hd->ttype[n] = (uint8_t)type; hd->tbits[n] = 0; hd->txsize[n] = xsize; hd->tdata[n] = tdata_used;
    if (type == AL_T_PREDICTOR || type == AL_T_CROSS_COLOR) {
      const int bits = (int)lb_read(b, 3) + 2;
      const int tx = (xsize + (1 << bits) - 1) >> bits, ty = (h + (1 << bits) - 1) >> bits;
      hd->tbits[n] = (uint8_t)bits;
      if (!al_decode_subimage(b, tx, ty, tdata + tdata_used, tables, cache, sc)) return;
      tdata_used += (uint32_t)(tx * ty);
    } else if (type == AL_T_COLOR_INDEXING) {
      const int num_colors = (int)lb_read(b, 8) + 1;
      const int bits = (num_colors > 16) ? 0 : (num_colors > 4) ? 1 : (num_colors > 2) ? 2 : 3;
      xsize = (xsize + (1 << bits) - 1) >> bits;
      hd->tbits[n] = (uint8_t)bits;
      if (!al_decode_subimage(b, num_colors, 1, pal, tables, cache, sc)) return;
      // ExpandColorMap (vp8l_dec.c:1305-1328): entries are byte-wise deltas; only green matters for alpha
      uint32_t g = 0, full = 0;
      for (int i = 0; i < 256; ++i) {
        if (i < num_colors) {
          g = (g + ((pal[i] >> 8) & 0xff)) & 0xff; hd->palette_alpha[i] = (uint8_t)g;
          full = (((full & 0xff00ff00u) + (pal[i] & 0xff00ff00u)) & 0xff00ff00u) | (((full & 0x00ff00ffu) + (pal[i] & 0x00ff00ffu)) & 0x00ff00ffu);
          hd->palette[i] = full;
        } else { hd->palette_alpha[i] = 0; hd->palette[i] = 0; }
      }
      if (n != 0) px_stride = hd->txsize[n];   // not the first transform (no encoder of the reference writes that): the pixels
                                               // widen to this width between two in-place transforms (al_inverse_transforms)
    }
  }
  // colour cache of the main image
  if (lb_read(b, 1)) {
    const int cache_bits = (int)lb_read(b, 4);
    if (cache_bits < 1 || cache_bits > AL_MAX_CACHE_BITS) return;
    hd->cache_bits = (uint8_t)cache_bits;
  }
  // meta-Huffman image (ReadHuffmanCodes, vp8l_dec.c:365-451)
  int num_groups = 1, used_groups = 1;
  hd->mapped = 0;
  if (lb_read(b, 1)) {
    const int precision = (int)lb_read(b, 3) + 2;
    const int hx = (xsize + (1 << precision) - 1) >> precision, hy = (h + (1 << precision) - 1) >> precision;
    uint32_t* img = (uint32_t*)meta;
    if (!al_decode_subimage(b, hx, hy, img, tables, cache, sc)) return;
    int max_group = 0;
    for (int i = 0; i < hx * hy; ++i) {
      const int group = (int)((img[i] >> 8) & 0xffff);
      meta[i] = (uint16_t)group;   // the i-th uint16 lies inside the (i/2)-th word: already consumed
      if (group > max_group) max_group = group;
    }
    num_groups = max_group + 1;
    hd->huff_bits = (uint8_t)precision;
    hd->huff_xsize = hx;
    // Many group numbers, or more than there are pixels: the reference keeps tables only for the numbers the image uses (it
    // still reads, and checks, the codes of the others) and looks at those alone when it picks its pixel loop
    // (vp8l_dec.c:399-424, 857-870). Dense numbers in increasing order here, in order of first appearance there: the same set.
    if (num_groups > 1000 || num_groups > xsize * h) {
      uint32_t* bits = AL_USED_BITS(scratch);
      uint16_t* below = AL_USED_BELOW(scratch);
      const int words = (num_groups + 31) >> 5;
      for (int k = 0; k < words; ++k) bits[k] = 0;
      for (int i = 0; i < hx * hy; ++i) bits[meta[i] >> 5] |= 1u << (meta[i] & 31);
      int used = 0;
      for (int k = 0; k < words; ++k) { below[k] = (uint16_t)used; used += AL_POPC(bits[k]); }
      for (int i = 0; i < hx * hy; ++i) meta[i] = (uint16_t)(below[meta[i] >> 5] + AL_POPC(bits[meta[i] >> 5] & ((1u << (meta[i] & 31)) - 1u)));
      used_groups = used;
      hd->mapped = 1;
    } else {
      used_groups = num_groups;
    }
  }
  if (b.eos) return;
  hd->xsize = xsize;
  hd->px_stride = px_stride ? px_stride : xsize;
  hd->num_groups = num_groups;
  hd->used_groups = used_groups;
  hd->group_entries = AL_GROUP_ENTRIES(hd->cache_bits);
  hd->br_val = b.val; hd->br_pos = b.pos; hd->br_bit_pos = b.bit_pos;
  hd->status = AL_OK;
}

// Pass B. tables = used_groups * group_entries words, groups = used_groups AlGroup, scratch as in pass A (its
// colour-cache area and AlScratch are reused), out = xsize * h ARGB words. last_row = rows the caller needs (the
// bottom of the crop window, h without cropping): like the reference, decoding stops there, so data missing further
// down is never noticed. Returns the image status.
// The codes of every group, the last thing the reference counts as header (VP8LDecodeHeader / VP8LDecodeAlphaHeader end here). The
// groups the meta-Huffman image never names are read and checked like the others but kept nowhere when the reference would have
// remapped them (alph_parse_header). Returns 0 on a bad code.
AL_NOINLINE int alph_read_groups(LBits& b, const AlphaHdr* hd, uint32_t* tables, AlGroup* groups, uint8_t* scratch, AlScratch* sc) {
  const int num_groups = hd->num_groups, stride = hd->group_entries, cache_bits = hd->cache_bits;
  const uint32_t* bits = AL_USED_BITS(scratch);
  const uint16_t* below = AL_USED_BELOW(scratch);
  AlGroup unused;
  for (int g = 0; g < num_groups; ++g) {
    int slot = g;
    if (hd->mapped) slot = ((bits[g >> 5] >> (g & 31)) & 1u) ? (int)(below[g >> 5] + AL_POPC(bits[g >> 5] & ((1u << (g & 31)) - 1u))) : -1;
    const int ok = slot >= 0 ? al_read_group(b, cache_bits, tables + (size_t)slot * stride, stride, &groups[slot], sc)
                             : al_read_group(b, cache_bits, (uint32_t*)scratch, stride, &unused, sc);   // pass A's table area is free now
    if (ok == 0) return 0;
  }
  return 1;
}

AL_NOINLINE int alph_decode_pixels(const uint8_t* alph, uint32_t alph_size, int h, int last_row, AlphaHdr* hd, const uint16_t* meta,
                                   uint32_t* tables, AlGroup* groups, uint8_t* scratch, uint32_t* out) {
  uint32_t* cache = (uint32_t*)scratch + AL_SUB_TABLE_ENTRIES;
  AlScratch* sc = (AlScratch*)(cache + (1 << AL_MAX_CACHE_BITS) + 256);
  LBits b;
  b.buf = alph + (hd->lossless ? 0 : 1); b.len = alph_size - (hd->lossless ? 0u : 1u); b.val = hd->br_val; b.pos = hd->br_pos; b.bit_pos = hd->br_bit_pos; b.eos = 0;
  const int used_groups = hd->used_groups, stride = hd->group_entries, cache_bits = hd->cache_bits;
  if (!alph_read_groups(b, hd, tables, groups, scratch, sc)) return AL_OUT_OF_MEMORY;   // (still part of the header as far as the status goes)
  // The reference runs DecodeAlphaData (vp8l_dec.c:1035-1116) when the only transform is the palette, there is no
  // colour cache and every group's R, B and A codes are zero-bit (Is8bOptimizable, :857-870), DecodeImageData
  // (:1138-1293) otherwise. Same symbols either way; what differs is when running out of data counts as a failure,
  // so keep both shapes. A whole VP8L picture always goes through DecodeImageData (VP8LDecodeImage, :1761-1765; the 8-bit
  // path is chosen by VP8LDecodeAlphaHeader alone, :1633-1641): there, data that runs out on the last symbol is an error.
  int use_8b = (!hd->lossless && hd->ntrans == 1 && hd->ttype[0] == AL_T_COLOR_INDEXING && cache_bits == 0);
  for (int g = 0; g < used_groups && use_8b; ++g) use_8b = groups[g].trivial_literal;
  hd->use_8b = (uint8_t)use_8b;
  const int cache_size = cache_bits ? (1 << cache_bits) : 0, cache_shift = 32 - cache_bits;
  for (int i = 0; i < cache_size; ++i) cache[i] = 0;
  const int width = hd->xsize;
  const int end = width * h, last = width * (last_row < h ? last_row : h);
  const int mask = hd->huff_bits ? (1 << hd->huff_bits) - 1 : -1;
  const int hbits = hd->huff_bits, hxs = hd->huff_xsize;
  const int len_code_limit = AL_NUM_LITERAL + AL_NUM_LENGTH;
  int pos = 0, col = 0, row = 0;
  int pos0 = 0;   // where the symbol being read starts: the reference meets a failure when asked for that row
  int ok = 1;
  const AlGroup* grp = &groups[hbits ? meta[0] : 0];
  const uint32_t* gt = tables + (size_t)(grp - groups) * stride;
  while (pos < last && !(use_8b && b.eos)) {
    pos0 = pos;
    if ((col & mask) == 0) {
      grp = &groups[hbits ? meta[hxs * (row >> hbits) + (col >> hbits)] : 0];
      gt = tables + (size_t)(grp - groups) * stride;
    }
    int code;
    uint32_t px = 0;
    if (!use_8b && grp->trivial_code) {
      code = 0;
      px = grp->literal_arb;
    } else {
      lb_fill(b);
      code = hc_read_symbol(gt + grp->off[0], b);
      if (!use_8b && lb_at_end(b)) break;
      if (code < AL_NUM_LITERAL) {
        if (use_8b || grp->trivial_literal) {
          px = grp->literal_arb | ((uint32_t)code << 8);
        } else {
          const uint32_t red = (uint32_t)hc_read_symbol(gt + grp->off[1], b);
          lb_fill(b);
          const uint32_t blue = (uint32_t)hc_read_symbol(gt + grp->off[2], b);
          const uint32_t alpha = (uint32_t)hc_read_symbol(gt + grp->off[3], b);
          if (lb_at_end(b)) break;
          px = (alpha << 24) | (red << 16) | ((uint32_t)code << 8) | blue;
        }
      }
    }
    if (code < AL_NUM_LITERAL) {
      out[pos++] = px;
      if (cache_size) cache[(px * 0x1e35a7bdu) >> cache_shift] = px;
      if (++col >= width) { col = 0; ++row; }
    } else if (code < len_code_limit) {
      const int length = al_copy_value(code - AL_NUM_LITERAL, b);
      const int dist_symbol = hc_read_symbol(gt + grp->off[4], b);
      lb_fill(b);
      const int dist = al_plane_to_distance(width, al_copy_value(dist_symbol, b));
      if (!use_8b && lb_at_end(b)) break;
      if (pos >= dist && end - pos >= length) {
        for (int k = 0; k < length; ++k) {
          const uint32_t v = out[pos + k - dist];
          out[pos + k] = v;
          if (cache_size) cache[(v * 0x1e35a7bdu) >> cache_shift] = v;
        }
      } else {
        ok = 0;
        break;
      }
      pos += length;
      col += length;
      while (col >= width) { col -= width; ++row; }
      if (pos < last && (col & mask)) {
        grp = &groups[hbits ? meta[hxs * (row >> hbits) + (col >> hbits)] : 0];
        gt = tables + (size_t)(grp - groups) * stride;
      }
    } else if (code < len_code_limit + cache_size) {
      px = cache[code - len_code_limit];
      out[pos++] = px;
      cache[(px * 0x1e35a7bdu) >> cache_shift] = px;
      if (++col >= width) { col = 0; ++row; }
    } else {
      ok = 0;
      break;
    }
    if (use_8b) b.eos = lb_at_end(b);
  }
  b.eos = lb_at_end(b);
  hd->fail_row = pos0 / width;
  if (!ok) return AL_BITSTREAM_ERROR;
  if (use_8b ? (b.eos && pos < end) : b.eos) return AL_BITSTREAM_ERROR;
  return AL_OK;
}

#if defined(__CUDACC__) && !defined(VP8_EMU)
// Backward references with more than one lane (the device): the serial decode (lane 0) does not need a single pixel VALUE as long
// as there is no colour cache -- symbols, positions, group switches and every failure rule depend on positions only -- so it
// writes its literals straight away, QUEUES the copies (position, distance, length) and runs ahead; when the queue is full or the
// stream ends, all lanes of the warp carry the queued copies out in order, each copy in parallel: element k of a copy reads
// out[pos - dist + k mod dist], which lies in the part written before the copy began whatever the overlap. The other lanes
// wait at the warp barrier meanwhile (no polling). With a colour cache the copies feed the cache in order: such pictures keep the
// one-lane loop above (alph_decode_pixels), and so does the host build. The two loops are the same text but for the queue; they
// are kept apart because the one-lane loop is a single dependent chain whose speed moved by 7-17 % with every change of the
// code around it (whole-picture VP8L, 1024 photographs: 1125 ms as it stands, 1209 / 1315 ms inside the queued form).
#define AL_COPY_INLINE 24   // shorter copies are done by lane 0 itself while the queue is empty

AL_NOINLINE int alph_decode_pixels_warp(const uint8_t* alph, uint32_t alph_size, int h, int last_row, AlphaHdr* hd, const uint16_t* meta,
                                   uint32_t* tables, AlGroup* groups, uint8_t* scratch, uint32_t* out,
                                   uint32_t* copyq, int lane) {
  uint32_t* cache = (uint32_t*)scratch + AL_SUB_TABLE_ENTRIES;
  AlScratch* sc = (AlScratch*)(cache + (1 << AL_MAX_CACHE_BITS) + 256);
  LBits b;
  b.buf = alph + (hd->lossless ? 0 : 1); b.len = alph_size - (hd->lossless ? 0u : 1u); b.val = hd->br_val; b.pos = hd->br_pos; b.bit_pos = hd->br_bit_pos; b.eos = 0;
  const int used_groups = hd->used_groups, stride = hd->group_entries, cache_bits = hd->cache_bits;
  int status = AL_OK, done = 0;
  int use_8b = 0;
  if (lane == 0) {
    if (!alph_read_groups(b, hd, tables, groups, scratch, sc)) { status = AL_OUT_OF_MEMORY; done = 1; }   // (still part of the header as far as the status goes)
    // The reference runs DecodeAlphaData (vp8l_dec.c:1035-1116) when the only transform is the palette, there is no
    // colour cache and every group's R, B and A codes are zero-bit (Is8bOptimizable, :857-870), DecodeImageData
    // (:1138-1293) otherwise. Same symbols either way; what differs is when running out of data counts as a failure,
    // so keep both shapes. A whole VP8L picture always goes through DecodeImageData (VP8LDecodeImage, :1761-1765; the 8-bit
    // path is chosen by VP8LDecodeAlphaHeader alone, :1633-1641): there, data that runs out on the last symbol is an error.
    if (!done) {
      use_8b = (!hd->lossless && hd->ntrans == 1 && hd->ttype[0] == AL_T_COLOR_INDEXING && cache_bits == 0);
      for (int g = 0; g < used_groups && use_8b; ++g) use_8b = groups[g].trivial_literal;
      hd->use_8b = (uint8_t)use_8b;
    }
  }
  const int cache_size = cache_bits ? (1 << cache_bits) : 0, cache_shift = 32 - cache_bits;
  if (lane == 0 && !done) for (int i = 0; i < cache_size; ++i) cache[i] = 0;
  const int width = hd->xsize;
  const int end = width * h, last = width * (last_row < h ? last_row : h);
  const int mask = hd->huff_bits ? (1 << hd->huff_bits) - 1 : -1;
  const int hbits = hd->huff_bits, hxs = hd->huff_xsize;
  const int len_code_limit = AL_NUM_LITERAL + AL_NUM_LENGTH;
  const int queue = copyq != 0 && cache_size == 0;
  int pos = 0, col = 0, row = 0;
  int pos0 = 0;   // where the symbol being read starts: the reference meets a failure when asked for that row
  int ok = 1;
  const AlGroup* grp = groups;
  const uint32_t* gt = tables;
  if (lane == 0 && !done) {
    grp = &groups[hbits ? meta[0] : 0];
    gt = tables + (size_t)(grp - groups) * stride;
  }
  for (;;) {
    int nq = 0;
    if (lane == 0 && !done) {
      int full = 0;   // left the loop because the queue filled up, not because the stream ended
      while (pos < last && !(use_8b && b.eos)) {
        pos0 = pos;
        if ((col & mask) == 0) {
          grp = &groups[hbits ? meta[hxs * (row >> hbits) + (col >> hbits)] : 0];
          gt = tables + (size_t)(grp - groups) * stride;
        }
        int code;
        uint32_t px = 0;
        if (!use_8b && grp->trivial_code) {
          code = 0;
          px = grp->literal_arb;
        } else {
          lb_fill(b);
          code = hc_read_symbol(gt + grp->off[0], b);
          if (!use_8b && lb_at_end(b)) break;
          if (code < AL_NUM_LITERAL) {
            if (use_8b || grp->trivial_literal) {
              px = grp->literal_arb | ((uint32_t)code << 8);
            } else {
              const uint32_t red = (uint32_t)hc_read_symbol(gt + grp->off[1], b);
              lb_fill(b);
              const uint32_t blue = (uint32_t)hc_read_symbol(gt + grp->off[2], b);
              const uint32_t alpha = (uint32_t)hc_read_symbol(gt + grp->off[3], b);
              if (lb_at_end(b)) break;
              px = (alpha << 24) | (red << 16) | ((uint32_t)code << 8) | blue;
            }
          }
        }
        if (code < AL_NUM_LITERAL) {
          out[pos++] = px;
          if (cache_size) cache[(px * 0x1e35a7bdu) >> cache_shift] = px;
          if (++col >= width) { col = 0; ++row; }
        } else if (code < len_code_limit) {
          const int length = al_copy_value(code - AL_NUM_LITERAL, b);
          const int dist_symbol = hc_read_symbol(gt + grp->off[4], b);
          lb_fill(b);
          const int dist = al_plane_to_distance(width, al_copy_value(dist_symbol, b));
          if (!use_8b && lb_at_end(b)) break;
          if (pos >= dist && end - pos >= length) {
            // a short copy with nothing queued before it is cheaper done on the spot (photographs: most copies are a few pixels
            // long); once something waits in the queue everything after it queues up behind it, to keep the order
            if (queue && (nq != 0 || length >= AL_COPY_INLINE)) {
              copyq[2 + 3 * nq] = (uint32_t)pos; copyq[3 + 3 * nq] = (uint32_t)dist; copyq[4 + 3 * nq] = (uint32_t)length;
              ++nq;
            } else {
              for (int k = 0; k < length; ++k) {
                const uint32_t v = out[pos + k - dist];
                out[pos + k] = v;
                if (cache_size) cache[(v * 0x1e35a7bdu) >> cache_shift] = v;
              }
            }
          } else {
            ok = 0;
            break;
          }
          pos += length;
          col += length;
          while (col >= width) { col -= width; ++row; }
          if (pos < last && (col & mask)) {
            grp = &groups[hbits ? meta[hxs * (row >> hbits) + (col >> hbits)] : 0];
            gt = tables + (size_t)(grp - groups) * stride;
          }
          if (nq == AL_COPYQ) {   // time to let the warp carry the queued copies out
            if (use_8b) b.eos = lb_at_end(b);
            full = 1;
            break;
          }
        } else if (code < len_code_limit + cache_size) {
          px = cache[code - len_code_limit];
          out[pos++] = px;
          cache[(px * 0x1e35a7bdu) >> cache_shift] = px;
          if (++col >= width) { col = 0; ++row; }
        } else {
          ok = 0;
          break;
        }
        if (use_8b) b.eos = lb_at_end(b);
      }
      if (!full) {   // the stream has ended, one way or another
        done = 1;
        b.eos = lb_at_end(b);
        hd->fail_row = pos0 / width;
        if (!ok) status = AL_BITSTREAM_ERROR;
        else if (use_8b ? (b.eos && pos < end) : b.eos) status = AL_BITSTREAM_ERROR;
      }
    }
    if (copyq == 0) break;   // one lane: nothing was queued, the loop above ran to the end
#if defined(__CUDACC__) && !defined(VP8_EMU)
    if (lane == 0) { copyq[0] = (uint32_t)nq; copyq[1] = (uint32_t)done; }
    __syncwarp();
    const int n = (int)copyq[0];
    const int all_done = (int)copyq[1];
    for (int j = 0; j < n; ++j) {
      const int p = (int)copyq[2 + 3 * j], d = (int)copyq[3 + 3 * j], l = (int)copyq[4 + 3 * j];
      const uint32_t* src = out + p - d;
      uint32_t* dst = out + p;
      if (d >= l) {
        for (int k = lane; k < l; k += 32) dst[k] = src[k];
      } else {
        int r = lane % d;
        const int step = 32 % d;
        for (int k = lane; k < l; k += 32) { dst[k] = src[r]; r += step; if (r >= d) r -= d; }
      }
      __syncwarp();   // the next copy may read what this one wrote
    }
    __syncwarp();
    if (all_done) break;
#else
    break;
#endif
  }
  return status;
}

#endif

// ---------------------------------------------------------------------------------------------------------
// Inverse transforms on ARGB words (src/dsp/lossless.c:28-340).
AL_FN uint32_t al_add_pixels(uint32_t a, uint32_t b) {
  return (((a & 0xff00ff00u) + (b & 0xff00ff00u)) & 0xff00ff00u) | (((a & 0x00ff00ffu) + (b & 0x00ff00ffu)) & 0x00ff00ffu);
}
AL_FN uint32_t al_avg2(uint32_t a, uint32_t b) { return (((a ^ b) & 0xfefefefeu) >> 1) + (a & b); }
AL_FN uint32_t al_clip255(uint32_t a) { return a < 256 ? a : (~a >> 24); }
AL_FN int al_sub3(int a, int b, int c) {
  const int pb = b - c, pa = a - c;
  return (pb < 0 ? -pb : pb) - (pa < 0 ? -pa : pa);
}
AL_FN uint32_t al_predict(int mode, uint32_t L, uint32_t T, uint32_t TL, uint32_t TR) {
  switch (mode) {
    case 1: return L;
    case 2: return T;
    case 3: return TR;
    case 4: return TL;
    case 5: return al_avg2(al_avg2(L, TR), T);
    case 6: return al_avg2(L, TL);
    case 7: return al_avg2(L, T);
    case 8: return al_avg2(TL, T);
    case 9: return al_avg2(T, TR);
    case 10: return al_avg2(al_avg2(L, TL), al_avg2(T, TR));
    case 11: {   // Select(top, left, top-left)
      const int d = al_sub3((int)(T >> 24), (int)(L >> 24), (int)(TL >> 24)) +
                    al_sub3((int)((T >> 16) & 0xff), (int)((L >> 16) & 0xff), (int)((TL >> 16) & 0xff)) +
                    al_sub3((int)((T >> 8) & 0xff), (int)((L >> 8) & 0xff), (int)((TL >> 8) & 0xff)) +
                    al_sub3((int)(T & 0xff), (int)(L & 0xff), (int)(TL & 0xff));
      return d <= 0 ? T : L;
    }
    case 12: {   // ClampedAddSubtractFull(left, top, top-left)
      uint32_t r = 0;
      for (int sh = 0; sh < 32; sh += 8)
        r |= al_clip255((uint32_t)((int)((L >> sh) & 0xff) + (int)((T >> sh) & 0xff) - (int)((TL >> sh) & 0xff))) << sh;
      return r;
    }
    case 13: {   // ClampedAddSubtractHalf(left, top, top-left)
      const uint32_t ave = al_avg2(L, T);
      uint32_t r = 0;
      for (int sh = 0; sh < 32; sh += 8) {
        const int a = (int)((ave >> sh) & 0xff), c = (int)((TL >> sh) & 0xff);
        r |= al_clip255((uint32_t)(a + (a - c) / 2)) << sh;
      }
      return r;
    }
    default: return 0xff000000u;   // mode 0, and the two unused codes 14 / 15
  }
}

#ifndef AL_BLOCK_SYNC
#define AL_BLOCK_SYNC() ((void)0)
#endif

// Inverse predictor transform in place (PredictorInverseTransform_C, lossless.c:186-231) as a wavefront: thread t
// of a band owns row y0 + t and trails the row above by two pixels (it needs its top-right neighbour). The
// top-right of a row's last pixel is the first pixel of the row itself: with rows stored back to back that is
// simply the next word.
AL_FN void al_inverse_predictor(uint32_t* px, int width, int h, const uint32_t* tiles, int bits, int tid, int nt) {
  const int tiles_per_row = (width + (1 << bits) - 1) >> bits;
  if (tid == 0) {   // first row: black for the first pixel, then the left neighbour
    uint32_t left = al_add_pixels(px[0], 0xff000000u);
    px[0] = left;
    for (int x = 1; x < width; ++x) { left = al_add_pixels(px[x], left); px[x] = left; }
  }
  AL_BLOCK_SYNC();
  for (int y0 = 1; y0 < h; y0 += nt) {
    const int y = y0 + tid;
    uint32_t* row = px + (size_t)y * width;
    const uint32_t* up = row - width;
    const uint32_t* trow = tiles + (size_t)(y >> bits) * tiles_per_row;
    uint32_t left = 0;
    for (int s = 0; s < width + 2 * (nt - 1); ++s) {
      const int x = s - 2 * tid;
      if (y < h && x >= 0 && x < width) {
        uint32_t pred;
        if (x == 0) pred = up[0];   // first pixel of a row: the pixel above
        else pred = al_predict((int)((trow[x >> bits] >> 8) & 0xf), left, up[x], up[x - 1], up[x + 1]);
        left = al_add_pixels(row[x], pred);
        row[x] = left;
      }
      AL_BLOCK_SYNC();
    }
  }
}

AL_FN int al_color_delta(int8_t pred, int8_t color) { return ((int)pred * (int)color) >> 5; }

// The inverse transforms of a VP8L stream, last read first, in place on the coded ARGB words (VP8LInverseTransform,
// lossless.c:391-442). COLOR_INDEXING is always transform 0 here and is folded into the caller's extraction.
AL_FN void al_inverse_transforms(const AlphaHdr* hd, uint32_t* px, const uint32_t* tdata, int h, int tid, int nt) {
  for (int n = (int)hd->ntrans - 1; n >= 0; --n) {
    const int type = hd->ttype[n], bits = hd->tbits[n], tw = hd->txsize[n];
    const uint32_t* tiles = tdata + hd->tdata[n];
    const size_t count = (size_t)tw * (size_t)h;
    if (type == AL_T_PREDICTOR) {
      al_inverse_predictor(px, tw, h, tiles, bits, tid, nt);
    } else if (type == AL_T_CROSS_COLOR) {   // ColorSpaceInverseTransform_C, lossless.c:284-338
      const int tiles_per_row = (tw + (1 << bits) - 1) >> bits;
      for (size_t i = (size_t)tid; i < count; i += (size_t)nt) {
        const int x = (int)(i % (size_t)tw), y = (int)(i / (size_t)tw);
        const uint32_t code = tiles[(size_t)(y >> bits) * tiles_per_row + (x >> bits)];
        const uint32_t argb = px[i];
        const int8_t green = (int8_t)(argb >> 8);
        int new_red = (int)((argb >> 16) & 0xff), new_blue = (int)(argb & 0xff);
        new_red = (new_red + al_color_delta((int8_t)(code & 0xff), green)) & 0xff;
        new_blue += al_color_delta((int8_t)((code >> 8) & 0xff), green);
        new_blue += al_color_delta((int8_t)((code >> 16) & 0xff), (int8_t)new_red);
        new_blue &= 0xff;
        px[i] = (argb & 0xff00ff00u) | ((uint32_t)new_red << 16) | (uint32_t)new_blue;
      }
    } else if (type == AL_T_SUBTRACT_GREEN) {   // VP8LAddGreenToBlueAndRed_C
      for (size_t i = (size_t)tid; i < count; i += (size_t)nt) {
        const uint32_t argb = px[i], green = (argb >> 8) & 0xff;
        px[i] = (argb & 0xff00ff00u) | (((argb & 0x00ff00ffu) + ((green << 16) | green)) & 0x00ff00ffu);
      }
    }
    else if (type == AL_T_COLOR_INDEXING && n != 0) {
      // VP8LColorIndexInverseTransform in place (lossless.c:341-385, vp8l_dec.c:... the reference moves the packed rows to the
      // end of the buffer first): rows of `xs` packed words become rows of `tw` pixels. Backwards, so that a word is read
      // before anything lands on it (the write position never falls below the read position). One thread: no encoder of the
      // reference writes this order.
      if (tid == 0) {
        const int xs = (tw + (1 << bits) - 1) >> bits, bpp = 8 >> bits;
        for (int y = h - 1; y >= 0; --y) {
          for (int x = tw - 1; x >= 0; --x) {
            const uint32_t packed = (px[(size_t)y * xs + (x >> bits)] >> 8) & 0xff;
            px[(size_t)y * tw + x] = hd->palette[(packed >> ((x & ((1 << bits) - 1)) * bpp)) & ((1u << bpp) - 1u)];
          }
        }
      }
    }
    // AL_T_COLOR_INDEXING as transform 0 is folded into the extraction below
    AL_BLOCK_SYNC();
  }
}

// Coded ARGB plane (or the raw payload, method 0) -> the final w x h alpha plane: inverse transforms in place in
// reverse reading order, palette/unbundle + green extraction, then the row unfilter in place. Runs on `nt`
// cooperating threads of one block (tid = 0..nt-1; AL_BLOCK_SYNC between phases), or on one host thread (nt = 1)
// in the emulation build.
//   none       : nothing to undo
//   horizontal : out[y][x] = in[y][x] + out[y][x-1], out[y][0] = in[y][0] + out[y-1][0]   -> column 0 first, then rows in parallel
//   vertical   : out[y][x] = in[y][x] + out[y-1][x], row 0 as horizontal                 -> row 0 first, then columns in parallel
//   gradient   : out[y][x] = in[y][x] + clip(left + top - topleft), row 0 as horizontal  -> skewed wavefront, one row per
//                thread, row y one column behind row y-1
// crop_top: first row of the output window. The reference's 8-bit path only unfilters from there when the filter
// is horizontal (ExtractPalettedAlphaRows, vp8l_dec.c:887-912), i.e. the window's first row is predicted from
// nothing instead of from the row above; reproduce that.
AL_FN void alph_finish(const AlphaHdr* hd, const uint8_t* raw /* method 0 */, uint32_t* px /* method 1 */, const uint32_t* tdata,
                       int w, int h, int crop_top, uint8_t* plane, int tid, int nt) {
  const size_t total = (size_t)w * (size_t)h;
  if (hd->method == 0) {
    for (size_t i = (size_t)tid; i < total; i += (size_t)nt) plane[i] = raw[i];
  } else {
    al_inverse_transforms(hd, px, tdata, h, tid, nt);
    const int has_palette = hd->ntrans > 0 && hd->ttype[0] == AL_T_COLOR_INDEXING;
    const int bits = has_palette ? hd->tbits[0] : 0, bpp = 8 >> bits;
    const int xs = hd->px_stride;
    for (size_t i = (size_t)tid; i < total; i += (size_t)nt) {
      const int x = (int)(i % (size_t)w), y = (int)(i / (size_t)w);
      if (has_palette) {   // VP8LColorIndexInverseTransform + green, lossless.c:341-385
        const uint32_t packed = (px[(size_t)y * xs + (x >> bits)] >> 8) & 0xff;
        plane[i] = hd->palette_alpha[(packed >> ((x & ((1 << bits) - 1)) * bpp)) & ((1u << bpp) - 1u)];
      } else {
        plane[i] = (uint8_t)(px[(size_t)y * xs + x] >> 8);
      }
    }
  }
  AL_BLOCK_SYNC();
  const int filter = hd->filter;
  if (filter == 0) return;
  if (filter == 1) {
    const int y_first = (hd->method == 1 && hd->use_8b) ? crop_top : 0;
    if (tid == 0) for (int y = y_first + 1; y < h; ++y) plane[(size_t)y * w] = (uint8_t)(plane[(size_t)y * w] + plane[(size_t)(y - 1) * w]);
    AL_BLOCK_SYNC();
    for (int y = tid; y < h; y += nt) {
      uint8_t* row = plane + (size_t)y * w;
      uint32_t acc = row[0];
      for (int x = 1; x < w; ++x) { acc = (acc + row[x]) & 0xff; row[x] = (uint8_t)acc; }
    }
    return;
  }
  // vertical and gradient: the first row is predicted horizontally (filters.c:205-233)
  if (tid == 0) {
    uint32_t acc = plane[0];
    for (int x = 1; x < w; ++x) { acc = (acc + plane[x]) & 0xff; plane[x] = (uint8_t)acc; }
  }
  AL_BLOCK_SYNC();
  if (filter == 2) {
    for (int x = tid; x < w; x += nt) {
      uint32_t acc = plane[x];
      for (int y = 1; y < h; ++y) { acc = (acc + plane[(size_t)y * w + x]) & 0xff; plane[(size_t)y * w + x] = (uint8_t)acc; }
    }
    return;
  }
  for (int y0 = 1; y0 < h; y0 += nt) {   // bands of nt rows; row y0 + t runs t columns behind row y0 + t - 1
    const int y = y0 + tid;
    uint8_t* row = plane + (size_t)y * w;
    const uint8_t* prev = row - w;
    int left = 0;
    for (int s = 0; s < w + nt - 1; ++s) {
      const int x = s - tid;
      if (y < h && x >= 0 && x < w) {
        const int pred = (x == 0) ? (int)prev[0] : al_gradient_predict(left, (int)prev[x], (int)prev[x - 1]);
        left = (row[x] + pred) & 0xff;
        row[x] = (uint8_t)left;
      }
      AL_BLOCK_SYNC();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// options.alpha_dithering_strength on planes whose levels the encoder quantised (ALPH header pre-processing = 1):
// WebPDequantizeLevels (src/utils/quant_levels_dec_utils.c:262-291; alpha_dec.c:200-230 applies it to the crop
// window of the finished plane). A (2r+1)^2 box average, r = 4 * strength / 100, pulls every pixel that is neither the
// plane's minimum nor its maximum towards the local mean through a correction curve that vanishes beyond the
// smallest distance between two used levels. The reference runs it as a rolling integral image in 16-bit modular
// arithmetic; the same numbers fall out of: (1) per row, prefix sums modulo 2^16 (`pre`, w x h uint16 of scratch);
// (2) per pixel, the sum of those over the 2r+1 rows around it (rows above the top replicate row 0; the last r rows
// are never filtered, exactly like the reference's loop) at the two or three columns its HFilter looks at.
// One thread block per plane; lut = 2047 int16 of shared memory, used = 256 bytes of shared memory.
AL_FN void alph_smooth(uint8_t* plane, int stride, int w, int h, int strength, uint16_t* pre, int16_t* lut, uint32_t* used,
                       int tid, int nt) {
  int radius = 4 * strength / 100;
  if (strength < 0 || strength > 100 || w <= 0 || h <= 0) return;
  if (2 * radius + 1 > w) radius = (w - 1) >> 1;
  if (2 * radius + 1 > h) radius = (h - 1) >> 1;
  if (radius <= 0) return;
  const int r = radius, R = 2 * r + 1;
  // ---- CountLevels (quant_levels_dec_utils.c:178-210)
  for (int k = tid; k < 256; k += nt) used[k] = 0;
  AL_BLOCK_SYNC();
  for (size_t i = (size_t)tid; i < (size_t)w * (size_t)h; i += (size_t)nt) used[plane[(i / (size_t)w) * (size_t)stride + (i % (size_t)w)]] = 1;
  AL_BLOCK_SYNC();
  int vmin = 255, vmax = 0, num_levels = 0, last = -1;
  for (int k = 0; k < 256; ++k) if (used[k]) { if (k < vmin) vmin = k; if (k > vmax) vmax = k; }
  int min_dist = vmax - vmin;
  for (int k = 0; k < 256; ++k) if (used[k]) { ++num_levels; if (last >= 0 && k - last < min_dist) min_dist = k - last; last = k; }
  if (num_levels <= 2) return;
  // ---- InitCorrectionLUT (:160-176), lut[i + 1023] = correction for a difference of i quarter-levels
  {
    const int threshold1 = min_dist << 2, threshold2 = (3 * threshold1) >> 2, delta = threshold1 - threshold2;
    for (int i = 1 + tid; i <= 1023; i += nt) {
      int c = (i <= threshold2) ? i : (i < threshold1) ? threshold2 * (threshold1 - i) / delta : 0;
      c >>= 2;
      lut[1023 + i] = (int16_t)c; lut[1023 - i] = (int16_t)-c;
    }
    if (tid == 0) lut[1023] = 0;
  }
  // ---- prefix sums of every row, modulo 2^16 (VFilter's `sum`, :79-91)
  for (int y = tid; y < h; y += nt) {
    const uint8_t* row = plane + (size_t)y * stride;
    uint16_t* o = pre + (size_t)y * w;
    uint32_t acc = 0;
    for (int x = 0; x < w; ++x) { acc += row[x]; o[x] = (uint16_t)acc; }
  }
  AL_BLOCK_SYNC();
  // ---- HFilter + ApplyFilter (:104-154) for the rows the reference's loop reaches
  const uint32_t scale = (uint32_t)((1 << 18) / (R * R));
  const int rows_out = h - r;
  for (size_t i = (size_t)tid; i < (size_t)w * (size_t)rows_out; i += (size_t)nt) {
    const int x = (int)(i % (size_t)w), d = (int)(i / (size_t)w);
    int xa, xb, xc = -1;   // delta = in[xa] + in[xb] (left), in[xa] - in[xb] (middle), 2 * in[w-1] - in[xb] - in[xc] (right)
    int kind;
    if (x <= r) { kind = 0; xa = x + r - 1; xb = r - x; }
    else if (x < w - r) { kind = 1; xa = x + r; xb = x - r - 1; }
    else { kind = 2; xa = w - 1; xb = 2 * w - 2 - r - x; xc = x - r - 1; }
    uint32_t sa = 0, sb = 0, sc = 0;
    for (int j = d - r; j <= d + r; ++j) {
      const uint16_t* prow = pre + (size_t)(j < 0 ? 0 : j) * w;
      sa += prow[xa]; sb += prow[xb];
      if (kind == 2) sc += prow[xc];
    }
    const uint16_t delta = (uint16_t)(kind == 0 ? sa + sb : kind == 1 ? sa - sb : 2u * sa - sb - sc);
    const uint16_t average = (uint16_t)(((uint32_t)delta * scale) >> 16);
    uint8_t* px = plane + (size_t)d * stride + x;
    const int v = *px;
    if (v < vmax && v > vmin) {
      const int c = v + (int)lut[1023 + (int)average - (v << 2)];
      *px = (uint8_t)(c < 0 ? 0 : c > 255 ? 255 : c);
    }
  }
  AL_BLOCK_SYNC();
}

#endif  // LIBWEBP_B200_VP8L_ALPHA_CORE_H_
