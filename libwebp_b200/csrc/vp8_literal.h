// Slow path for images one of whose partitions starts with byte 0xFF (ImgDesc flag VP8B_FLAG_LITERAL_READER, set by the host):
// the whole image is parsed again -- frame header, intra modes, tokens -- by ONE thread with the reference's reader taken
// literally (vp8_parse_core.h:RefBits), in the reference's own order (vp8_dec.c:646-674: rows top to bottom, row my from
// partition my % P), and leaves exactly what K1 and the token parser leave: FrameHdr, MbInfo, the token stream and MbTok.
// No encoder writes such a file; this exists so that a damaged one ends with the reference's status and pixels.
// Compiled by nvcc (k_parse_literal) and by g++ (tests/emu).
#ifndef LIBWEBP_B200_VP8_LITERAL_H_
#define LIBWEBP_B200_VP8_LITERAL_H_
#include "vp8_parse_core.h"
#include "vp8_tokens_fp.h"   // MbTok, TF_TOKENS_PER_MB, the token format

// scratch of one image (any memory): position-major probabilities | progress | one macroblock of levels | top modes | top contexts
#define LIT_POSPROB 0
#define LIT_PROGRESS 2256
#define LIT_LEVELS (LIT_PROGRESS + 48)
#define LIT_TOPMODES (LIT_LEVELS + 2 * VP8B_COEFFS_PER_MB)
#define LIT_SCRATCH_BYTES(mb_w) ((size_t)LIT_TOPMODES + 4u * (size_t)(mb_w) + 2u * (size_t)(VP8B_MAX_PARTS + 1) * (size_t)(mb_w))

// A macroblock's 400 levels -> tokens {sign 31, block 29:25, magnitude 24:13, position 9:6}, appended; the buffer is left zero.
struct LiteralSink {
  static const int kPerMb = 1;
  uint32_t* tokens;
  MbTok* mbtok;
  uint32_t* off;
  int mb_w;
  VP8_MFN void mb(int mx, int my, int16_t* lv) {
    const uint32_t first = *off;
    uint32_t o = first;
    for (int k = 0; k < VP8B_COEFFS_PER_MB; ++k) {
      const int v = lv[k];
      if (v == 0) continue;
      lv[k] = 0;
      tokens[o++] = (v < 0 ? 0x80000000u : 0u) | ((uint32_t)(k >> 4) << 25) | ((uint32_t)(v < 0 ? -v : v) << TF_ADD_SHIFT) | ((uint32_t)(k & 15) << 6);
    }
    MbTok t; t.first = first; t.count = o - first;
    mbtok[(size_t)my * mb_w + mx] = t;
    *off = o;
  }
};

// mbinfo / tokens / mbtok: this image's areas. bprob = kVp8BModeProba (900 bytes). scratch: LIT_SCRATCH_BYTES(im.mb_w), the
// level buffer inside it zeroed by the caller.
VP8_FN void parse_image_literal(const uint8_t* frame, const ImgDesc& im, FrameHdr* h, const uint8_t* bprob, uint32_t* mbinfo,
                                uint32_t* tokens, MbTok* mbtok, uint8_t* scratch) {
  uint8_t* posprob = scratch + LIT_POSPROB;
  int* progress = (int*)(scratch + LIT_PROGRESS);
  int16_t* lv = (int16_t*)(scratch + LIT_LEVELS);
  uint32_t* topmodes = (uint32_t*)(scratch + LIT_TOPMODES);
  uint16_t* topctx = (uint16_t*)(topmodes + im.mb_w);
  RefBits br;
  int st = parse_frame_header(br, frame, im, h);
  int fail_row = st == VP8B_OK ? VP8B_FAIL_NONE : VP8B_FAIL_HEADERS;
  if (st == VP8B_OK) st = parse_intra_modes(br, im, h, topmodes, bprob, mbinfo, &fail_row);
  h->modes_status = VP8B_OK;
  h->all_rows = h->rows;
  int rows = h->rows;
  if (st != VP8B_OK) {
    if (!(fail_row > 0 && fail_row != VP8B_FAIL_NONE)) { h->fail_row = fail_row; h->status = st; return; }
    rows = fail_row;   // the modes ran dry at this row: the tokens of the rows above still count (a failure there comes first)
  }
  for (int k = 0; k < VP8B_POSPROB_BYTES; ++k) posprob[k] = posprob_byte(h->prob, k);
  const int P = h->num_parts;
  TokenPartT<RefBits> tp[VP8B_MAX_PARTS];
  for (int p = 0; p < P; ++p) { token_part_init(tp[p], frame, h, p); progress[p] = 0; }
  uint32_t off = 0;
  LiteralSink sink; sink.tokens = tokens; sink.mbtok = mbtok; sink.off = &off; sink.mb_w = im.mb_w;
  for (int my = 0; my < rows; ++my) {
    const int p = my % P;
    parse_token_row(tp[p], im, h, p, my, posprob, topctx, progress, mbinfo, lv, sink);
    if (tp[p].status != VP8B_OK) { st = tp[p].status; fail_row = my; break; }
  }
  h->fail_row = fail_row;
  h->status = st;
}

#endif  // LIBWEBP_B200_VP8_LITERAL_H_
