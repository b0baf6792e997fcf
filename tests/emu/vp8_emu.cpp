// vp8_emu.cpp -- TEST INFRASTRUCTURE ONLY. Host build of the product's device logic (libwebp_b200/csrc/
// vp8_parse_core.h, vp8_pixel_core.h compiled with -DVP8_EMU): a warp phase becomes a loop over 32 lanes, a
// wavefront step a loop over its macroblocks. It lets tests/test_emu.py check the kernels' arithmetic,
// indexing and dependency analysis against the oracle where no GPU exists. It is never linked into the
// shipped library and is not a decode path of the product.
#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>
#include <cstdlib>
#include <cstdio>

#define VP8_EMU 1
#define VP8_WAIT_PROGRESS(ptr, need) assert(*(ptr) >= (need))
#define VP8_PUBLISH_PROGRESS(ptr, val) (*(ptr) = (val))
#include "vp8_container.h"
#include "vp8_parse_core.h"
#include "vp8_pixel_core.h"
#include "vp8_tokens_fsm.h"
#include "vp8_tokens_lockstep.h"
#include "vp8_tokens_fp.h"
#include "vp8_modes_lockstep.h"
#include "vp8_literal.h"
#include "vp8l_alpha_core.h"
#include "vp8l_lossless_core.h"
#include "vp8l_alpha_core.h"

// variant bit 0: visit the macroblocks of a wavefront step in reverse order
// variant bit 1: token parse with the row-at-a-time reference port (parse_token_row) instead of the lane FSM
// variant bit 2: lane FSM with a lazy ring producer
// variant bit 3: lockstep lane parser (vp8_tokens_lockstep.h), lanes advanced round-robin; with bit 4 its grouped event points
// variant bit 8: the lockstep mode parser (vp8_modes_lockstep.h) instead of parse_intra_modes
// variant bit 6: the fp parser (vp8_tokens_fp.h): fp32 boolean decoder, token stream out, read back by recon_load_tokens
static int emu_decode_crop(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                           int stride, int variant, uint8_t* unfiltered, int crop_x, int crop_y, int crop_w, int crop_h,
                           int scaled_w = 0, int scaled_h = 0);

extern "C" int emu_decode(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                          int stride, int variant, uint8_t* unfiltered /* optional: y|u|v padded */) {
  return emu_decode_crop(data, size, csp, flags, out, out_size, stride, variant, unfiltered, 0, 0, 0, 0);
}

// crop_w == 0: no cropping. flags bit 2 = flip. `stride` and `out` describe the (cropped) output.
extern "C" int emu_decode_window(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                                 int stride, int crop_x, int crop_y, int crop_w, int crop_h) {
  return emu_decode_crop(data, size, csp, flags, out, out_size, stride, 0, nullptr, crop_x, crop_y, crop_w, crop_h);
}

// options.use_scaling: the (cropped) picture rescaled to scaled_w x scaled_h (both given); `stride` and `out` describe
// the scaled picture.
// options.dithering_strength = strength (0..100) on the whole picture.
static int g_emu_dither_f = 0, g_emu_alpha_dither = 0;
// stage dump of the next emu_decode call (what the two parse stages leave: MbInfo, levels), for the device test hook's comparison
static uint32_t* g_dump_mbinfo = nullptr;
static int16_t* g_dump_levels = nullptr;
static size_t g_dump_max_mb = 0;
extern "C" void emu_set_stage_dump(uint32_t* mbinfo, int16_t* levels, size_t max_mb) { g_dump_mbinfo = mbinfo; g_dump_levels = levels; g_dump_max_mb = max_mb; }

// K2 with the fp parser (k_parse_tokens_fp): the image's partitions as lanes advanced round-robin, one group at a time.
template <int BAND>
static void emu_fp_tokens(FrameHdr& hdr, const uint8_t* frame, int mb_w, int rows, int variant, std::vector<uint32_t>& mbinfo,
                          std::vector<uint32_t>& tokens, std::vector<MbTok>& mbtok) {
  const int P = hdr.num_parts;
  std::vector<uint8_t> imgmem(TF_IMG_BYTES_B(BAND) + 1024 + 1024 + 128);   // (wandering lanes of the banded layout read a little below the rows)
  uint8_t* img = (uint8_t*)(((uintptr_t)imgmem.data() + 1024 + 1023) & ~(uintptr_t)1023);
  std::vector<uint64_t> tabmem((sizeof(TfTables) + 7) / 8);
  TfTables* ttab = (TfTables*)tabmem.data();
  tf_image_fill(img, &hdr, 0, 1, BAND);
  tf_tables_fill(ttab, 0, 1);
  std::vector<uint16_t> topctx((size_t)(P + 1) * mb_w, 0);
  std::vector<int> progress(VP8B_MAX_PARTS, 0);
  std::vector<TfLaneT<BAND>> lanes(P);
  std::vector<TfCtx> ctxs(P);
  std::vector<int> live(P, 0);
  for (int p = 0; p < P && p < rows; ++p) {
    TfCtx& cc = ctxs[p];
    cc.img_s = tk_saddr_of(img); cc.tab_s = tk_saddr_of(ttab);
    cc.k.mant_mask = 0x007fffffu; cc.k.exp46 = TF_EXP46;
    cc.topctx = topctx.data(); cc.progress = progress.data();
    cc.mbinfo = mbinfo.data(); cc.mbtok = mbtok.data(); cc.tokens = tokens.data();
    cc.mb_w = mb_w; cc.rows = rows; cc.P = P; cc.part = p; cc.use_skip = hdr.use_skip; cc.ctx_stride = mb_w;
    tf_lane_init(lanes[p], cc, frame, &hdr);
    live[p] = 1;
  }
  const bool inline_style = (variant & 16) != 0;   // variant bit 4: a branch per decode instead of the straight-line groups
  if (inline_style && P == 1 && live[0] && !tf_mb_next<0>(lanes[0], ctxs[0])) { tf_lane_park(lanes[0], ctxs[0]); live[0] = 0; }
  for (bool any = true; any;) {
    any = false;
    for (int p = P - 1; p >= 0; --p) {   // reverse order: exercises the wait-for-progress path
      if (!live[p]) continue;
      any = true;
      // parked lanes keep stepping, harmlessly, like on the device
      if (inline_style) { if (P > 1) tf_group_inline<1, 0>(lanes[p], ctxs[p]); else tf_group_inline<0, 0>(lanes[p], ctxs[p]); }
      else { if (P > 1) tf_group_flat<1, 0>(lanes[p], ctxs[p]); else tf_group_flat<0, 0>(lanes[p], ctxs[p]); }
      live[p] = lanes[p].alive;
    }
  }
  for (int p = 0; p < P && p < rows; ++p) if (lanes[p].status != VP8B_OK) hdr.status = lanes[p].status;
  if (hdr.status != VP8B_OK) {   // like the prologue of k_reconstruct
    const int row = tf_find_failed_row(mbtok.data(), mb_w, rows);
    if (row < hdr.fail_row) hdr.fail_row = row;
  }
}

// K6's two serial passes on an ALPH chunk (k_alpha_header, k_alpha_pixels): leaves the status and, after a failure, the alpha
// row the reference would have been asked for when it met it (AlphaHdr::fail_row).
static void emu_alpha_passes(const uint8_t* alph, uint32_t alph_size, const ImgDesc& im, AlphaHdr* ah, std::vector<uint32_t>& coded,
                             std::vector<uint32_t>& tdata) {
  std::vector<uint8_t> scratch(AL_SCRATCH_BYTES + 64);
  uint8_t* sc16 = (uint8_t*)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
  std::vector<uint32_t> meta(AL_META_PIXELS_BOUND(im.width, im.height) + 8);
  tdata.assign(2 * (size_t)AL_META_PIXELS_BOUND(im.width, im.height) + 8, 0);
  alph_parse_header(alph, alph_size, im.width, im.height, sc16, (uint16_t*)meta.data(), tdata.data(), ah);
  if (ah->status == AL_OK && ah->method == 1) {
    std::vector<uint32_t> tables((size_t)ah->used_groups * ah->group_entries);
    std::vector<AlGroup> groups(ah->used_groups);
    coded.assign((size_t)(ah->xsize > ah->px_stride ? ah->xsize : ah->px_stride) * im.height + 4, 0);
    ah->status = alph_decode_pixels(alph, alph_size, im.height, (int)im.crop_y + (int)im.out_h, ah, (const uint16_t*)meta.data(),
                                    tables.data(), groups.data(), sc16, coded.data());
  }
}
extern "C" int emu_decode_dithered(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                                   int stride, int strength, int crop_x, int crop_y, int crop_w, int crop_h);

extern "C" int emu_decode_scaled(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                                 int stride, int crop_x, int crop_y, int crop_w, int crop_h, int scaled_w, int scaled_h) {
  return emu_decode_crop(data, size, csp, flags, out, out_size, stride, 0, nullptr, crop_x, crop_y, crop_w, crop_h, scaled_w, scaled_h);
}

static int emu_decode_crop(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                           int stride, int variant, uint8_t* unfiltered, int crop_x, int crop_y, int crop_w, int crop_h,
                           int scaled_w, int scaled_h) {
  const int reverse_steps = variant & 1;
  Vp8Container c;
  int st = vp8b_parse_container(data, size, 1, &c);
  if (st != VP8_STATUS_OK) return st;
  if (c.has_animation) return VP8_STATUS_UNSUPPORTED_FEATURE;
  if (c.is_lossless) {   // whole-picture VP8L: passes A and B of the ALPH decoder, then vp8l_lossless_core.h (like k_lossless_finish)
    ImgDesc im;
    memset(&im, 0, sizeof(im));
    im.width = (uint16_t)c.width; im.height = (uint16_t)c.height;
    im.csp = (uint8_t)csp; im.flags = (uint8_t)(flags | VP8B_FLAG_LOSSLESS); im.out_stride = stride;
    im.out_w = im.width; im.out_h = im.height;
    if (crop_w > 0) { im.crop_x = (uint16_t)crop_x; im.crop_y = (uint16_t)crop_y; im.out_w = (uint16_t)crop_w; im.out_h = (uint16_t)crop_h; }
    if (scaled_w > 0) { im.dst_w = (uint32_t)scaled_w; im.dst_h = (uint32_t)scaled_h; }
    const uint8_t* bits = data + c.frame_offset;
    const uint32_t nbits = (uint32_t)c.frame_size;
    std::vector<uint8_t> scratch(AL_SCRATCH_BYTES + 64);
    uint8_t* sc16 = (uint8_t*)(((uintptr_t)scratch.data() + 15) & ~(uintptr_t)15);
    std::vector<uint32_t> meta(AL_META_PIXELS_BOUND(im.width, im.height) + 8);
    std::vector<uint32_t> tdata(2 * (size_t)AL_META_PIXELS_BOUND(im.width, im.height) + 8);
    AlphaHdr ah;
    alph_parse_header(bits, nbits, im.width, im.height, sc16, (uint16_t*)meta.data(), tdata.data(), &ah, 1);
    std::vector<uint32_t> coded;
    if (ah.status == AL_OK) {
      std::vector<uint32_t> tables((size_t)ah.used_groups * ah.group_entries);
      std::vector<AlGroup> groups(ah.used_groups);
      coded.assign((size_t)(ah.xsize > ah.px_stride ? ah.xsize : ah.px_stride) * im.height + 4, 0);
      ah.status = alph_decode_pixels(bits, nbits, im.height, (int)im.crop_y + (int)im.out_h, &ah, (const uint16_t*)meta.data(), tables.data(),
                                     groups.data(), sc16, coded.data());
    }
    // every failure of a whole-picture decode is a bitstream error (vp8l_dec.c:1292,1479-1488; nothing suspends outside idec)
    if (ah.status != AL_OK) return ah.status == AL_UNSUPPORTED ? VP8_STATUS_UNSUPPORTED_FEATURE : VP8_STATUS_BITSTREAM_ERROR;
    const int fw = im.dst_w ? im.dst_w : im.out_w, fh = im.dst_h ? im.dst_h : im.out_h;
    const size_t need = (csp == 11 || csp == 12)
        ? (size_t)im.out_stride * fh + 2 * (size_t)((fw + 1) / 2) * ((fh + 1) / 2) + (csp == 12 ? (size_t)fw * fh : 0)
        : (size_t)im.out_stride * (fh - 1) + (size_t)fw * ((csp == 0 || csp == 2) ? 3 : (csp == 5 || csp == 6 || csp == 10) ? 2 : 4);
    if (need > out_size) return VP8_STATUS_INVALID_PARAM;
    std::vector<uint8_t> pm(im.dst_w ? 4 * (size_t)im.out_w * im.out_h + 16 : 16);
    vp8l_finish_picture(&ah, im, coded.data(), tdata.data(), out, pm.data(), 0, 1);
    return VP8_STATUS_OK;
  }
  if (c.part0_size > c.frame_size - 10) return VP8_STATUS_NOT_ENOUGH_DATA;

  // input arena with padding on both sides, like the device arena
  std::vector<uint8_t> arena(size + 64 + 32768 + 64, 0xA5);
  memcpy(arena.data() + 64, data, size);
  ImgDesc im;
  memset(&im, 0, sizeof(im));
  im.in_off = 64 + c.frame_offset;
  im.vp8_size = (uint32_t)c.frame_size;
  im.part0_size = c.part0_size;
  im.width = (uint16_t)c.width; im.height = (uint16_t)c.height;
  im.mb_w = (uint16_t)((c.width + 15) >> 4); im.mb_h = (uint16_t)((c.height + 15) >> 4);
  im.csp = (uint8_t)csp; im.flags = (uint8_t)flags; im.out_stride = stride;
  im.out_w = im.width; im.out_h = im.height;
  if (crop_w > 0) { im.crop_x = (uint16_t)(crop_x & ~1); im.crop_y = (uint16_t)(crop_y & ~1); im.out_w = (uint16_t)crop_w; im.out_h = (uint16_t)crop_h; }
  if (scaled_w > 0) {   // like batch_build (vp8_batch.cu): scaled dimensions, and no loop filter for large downscaling ratios
    im.dst_w = (uint32_t)scaled_w; im.dst_h = (uint32_t)scaled_h;
    if (scaled_w < c.width * 3 / 4 && scaled_h < c.height * 3 / 4) im.flags |= VP8B_FLAG_BYPASS_FILTER;
  }
  im.dither_f = (uint8_t)g_emu_dither_f;
  im.num_parts = (uint8_t)vp8b_prescan_partitions(data + c.frame_offset + 10, c.part0_size);
  const int mb_w = im.mb_w, mb_h = im.mb_h;
  const size_t nmb = (size_t)mb_w * mb_h;
  const uint8_t* frame = arena.data() + im.in_off;

  FrameHdr hdr;
  memset(&hdr, 0, sizeof(hdr));
  std::vector<uint32_t> mbinfo(nmb * 4, 0);
  std::vector<int16_t> coeffs(nmb * VP8B_COEFFS_PER_MB, 0);
  std::vector<uint8_t> yuv(nmb * 384, 0);
  uint8_t* yp = yuv.data(); uint8_t* up = yp + nmb * 256; uint8_t* vp = up + nmb * 64;

  // A partition that starts with 0xFF: the whole image goes through the reference's reader taken literally instead (k_parse_literal)
  const bool literal = (variant & 64) && c.frame_size >= 10 &&
                       vp8b_partition_starts_with_ff(data + c.frame_offset + 10, c.part0_size, c.frame_size - 10, im.num_parts);
  std::vector<uint32_t> tokens;
  std::vector<MbTok> mbtok;
  // K1: header + intra modes
  if (literal) {
    tokens.assign(nmb * TF_TOKENS_PER_MB, 0xffffffffu);
    mbtok.assign(nmb, MbTok{ 0xffffffffu, 0xffffffffu });
    std::vector<uint8_t> scratch(LIT_SCRATCH_BYTES(mb_w) + 16, 0);
    parse_image_literal(frame, im, &hdr, kVp8BModeProba, mbinfo.data(), tokens.data(), mbtok.data(), scratch.data());
  } else {
    BoolDec br;
    std::vector<uint32_t> top(mb_w);
    hdr.status = parse_frame_header(br, frame, im, &hdr);
    hdr.fail_row = hdr.status == VP8B_OK ? VP8B_FAIL_NONE : VP8B_FAIL_HEADERS;
    if (hdr.status == VP8B_OK && (variant & 256)) {   // variant bit 8: the lockstep mode parser (vp8_modes_lockstep.h), this image as its one lane
      uint32_t tab[2 * ML_NODES];
      uint8_t row[ML_ROW_BYTES];
      ml_table_fill(tab, 0, 1);
      ml_row_fill(row, &hdr);
      MlCtx mc;
      mc.tab_s = tk_saddr_of(tab); mc.bprob_s = tk_saddr_of(kVp8BModeProba); mc.row_s = tk_saddr_of(row); mc.top_s = tk_saddr_of(top.data());
      mc.out = mbinfo.data(); mc.mb_w = mb_w; mc.mb_h = hdr.rows;
      mc.skip_node8 = hdr.use_skip ? 8u * ML_SKIP : 8u * ML_I16; mc.skip_off = hdr.use_skip ? ML_OFF_SKIP : ML_OFF_I16;
      mc.first_node8 = hdr.update_map ? 8u * ML_S0 : mc.skip_node8; mc.first_off = hdr.update_map ? ML_OFF_S0 : mc.skip_off;
      mc.k.mant_mask = 0x007fffffu; mc.k.exp46 = TF_EXP46;
      MlLane L;
      ml_start(L, mc, br);
      while (L.alive) ml_group(L, mc);
      hdr.status = L.status;
      if (L.status != VP8B_OK) hdr.fail_row = L.fail_row;
    } else
    if (hdr.status == VP8B_OK) hdr.status = parse_intra_modes(br, im, &hdr, top.data(), kVp8BModeProba, mbinfo.data(), &hdr.fail_row);
  }
  // Both chunks damaged: the fp parser keeps the row at which the VP8 stream fails, and the image's status is whichever
  // failure the reference's row loop meets first (vp8_dev.h:vp8b_vp8_failure_first, like batch_finish in vp8_batch.cu).
  const auto failed = [&](int vp8_status) -> int {
    if (!c.has_alph_chunk || !(variant & 64)) return vp8_status;
    AlphaHdr ah;
    std::vector<uint32_t> coded, tdata;
    emu_alpha_passes(data + c.alpha_offset, (uint32_t)c.alpha_size, im, &ah, coded, tdata);
    if (ah.status == AL_OK) return vp8_status;
    const int all_at_once = ah.levels != 0;
    if (getenv("VP8_EMU_DEBUG")) fprintf(stderr, "vp8 status %d row %d | alpha status %d row %d | filter %d rows %d crop_bottom %d\n", vp8_status, hdr.fail_row, ah.status, ah.fail_row, hdr.filter_type, hdr.rows, (int)im.crop_y + (int)im.out_h);
    return vp8b_vp8_failure_first(hdr.fail_row == VP8B_FAIL_NONE ? VP8B_FAIL_HEADERS : hdr.fail_row, ah.fail_row, hdr.filter_type, hdr.rows,
                                  (int)im.crop_y + (int)im.out_h, all_at_once) ? vp8_status : ah.status;
  };
  // intra modes that ran out at row r > 0: the fp parser still reads the tokens of the rows above (k_parse_tokens_fp)
  const bool modes_short = !literal && hdr.status == VP8B_NOT_ENOUGH_DATA && hdr.fail_row > 0 && hdr.fail_row != VP8B_FAIL_NONE && (variant & 64) && c.has_alph_chunk;
  if (hdr.status != VP8B_OK && !modes_short) return failed(hdr.status);
  if (!literal && hdr.num_parts != im.num_parts) return -100;   // host pre-scan disagrees with the device parse
  const int rows = modes_short && hdr.fail_row < hdr.rows ? hdr.fail_row : hdr.rows;   // macroblock rows that get decoded (all of them unless cropping)

  // K2: tokens
  if (literal) {   // done above
  } else if (variant & 64) {   // fp parser: one lane per partition, one decode per lane per round, levels as a token stream
    tokens.assign(nmb * TF_TOKENS_PER_MB, 0xffffffffu);
    mbtok.assign(nmb, MbTok{ 0xffffffffu, 0xffffffffu });
    // variant bit 7: the banded layout of the probability rows (the launches with very many small images)
    if (variant & 128) emu_fp_tokens<1>(hdr, frame, mb_w, rows, variant, mbinfo, tokens, mbtok);
    else emu_fp_tokens<0>(hdr, frame, mb_w, rows, variant, mbinfo, tokens, mbtok);
  } else if (variant & 8) {   // lockstep parser: one lane per partition, one decode per lane per round
    const int P = hdr.num_parts;
    std::vector<uint8_t> imgmem(TL_IMG_BYTES + 16 + 128);   // the look-ahead loads run up to 63 bytes past the rows
    uint8_t* img16 = (uint8_t*)(((uintptr_t)imgmem.data() + 15) & ~(uintptr_t)15);
    std::vector<uint64_t> tabmem((sizeof(TlTables) + 7) / 8);
    TlTables* ttab = (TlTables*)tabmem.data();
    tl_image_fill(img16, &hdr, 0, 1);
    tl_tables_fill(ttab, 0, 1);
    std::vector<uint16_t> topctx((size_t)(P + 1) * mb_w, 0);
    std::vector<int> progress(VP8B_MAX_PARTS, 0);
    std::vector<TlLane> lanes(P);
    std::vector<TlCtx> ctxs(P);
    std::vector<int> live(P, 0);
    for (int p = 0; p < P && p < rows; ++p) {
      TlCtx& cc = ctxs[p];
      cc.img_s = tk_saddr_of(img16); cc.tab_s = tk_saddr_of(ttab);
      cc.topctx = topctx.data(); cc.progress = progress.data();
      cc.mbinfo = mbinfo.data(); cc.coeffs = coeffs.data();
      cc.mb_w = mb_w; cc.rows = rows; cc.P = P; cc.part = p; cc.use_skip = hdr.use_skip; cc.ctx_stride = mb_w;
      tl_lane_init(lanes[p], cc, frame, &hdr);
      live[p] = 1;
    }
    const bool grouped = (variant & 16) != 0;   // variant bit 4: grouped event points instead of block ends on the spot
    if (!grouped && P == 1 && live[0] && !tl_mb_next<0>(lanes[0], ctxs[0])) { tl_lane_park(lanes[0], ctxs[0]); live[0] = 0; }
    for (bool any = true; any;) {
      any = false;
      for (int p = P - 1; p >= 0; --p) {   // reverse order: exercises the wait-for-progress path
        if (!live[p]) continue;
        any = true;
        if (grouped && (variant & 32)) {   // variant bit 5: straight-line groups
          if (P > 1) tl_group_flat<1>(lanes[p], ctxs[p]); else tl_group_flat<0>(lanes[p], ctxs[p]);
        } else if (grouped) {
          if (P > 1) tl_group<1>(lanes[p], ctxs[p]); else tl_group<0>(lanes[p], ctxs[p]);
        } else {
          bd_fill_lookahead(lanes[p].d);
          for (int k = 0; k < 4; ++k) {   // parked lanes keep stepping, harmlessly, like on the device
            if (P > 1) tl_step_inline<1>(lanes[p], ctxs[p]); else tl_step_inline<0>(lanes[p], ctxs[p]);
          }
        }
        live[p] = lanes[p].alive;
      }
    }
    for (int p = 0; p < P && p < rows; ++p) if (lanes[p].status != VP8B_OK) hdr.status = lanes[p].status;
  } else if (!(variant & 2)) {   // lane FSM: one lane per partition, lanes advanced round-robin one iteration at a time
    const int P = hdr.num_parts;
    TokImage timg;
    TokTables ttab;
    tk_image_fill(&timg, &hdr, P, 0, 1);
    tk_tables_fill(&ttab, 0, 1);
    std::vector<uint16_t> topctx((size_t)(P + 1) * mb_w, 0);
    std::vector<int> progress(P, 0);
    TokShared sh; sh.img = &timg; sh.img_s = tk_saddr_of(&timg); sh.tab_s = tk_saddr_of(&ttab);
    sh.topctx = topctx.data(); sh.progress = progress.data();
    std::vector<TokLane> lanes(P);
    // the arena buffer is only byte-aligned by the allocator's grace: use a 16-byte aligned view
    const uintptr_t abase = (uintptr_t)arena.data();
    const uint8_t* arena16 = (const uint8_t*)(abase & ~(uintptr_t)15);
    const uint64_t frame_off = (uint64_t)((const uint8_t*)frame - arena16);
    // one shared-memory ring + protocol block per stream, topped up by an emulated producer
    std::vector<uint8_t> ringmem((size_t)P * TK_RING_BYTES + 16);
    uint8_t* rings = (uint8_t*)(((uintptr_t)ringmem.data() + 15) & ~(uintptr_t)15);
    std::vector<TokStreamCtl> ctl(P);
    for (int p = 0; p < P; ++p) {
      if (p >= rows) { ctl[p].rd_w = TK_STREAM_DONE; ctl[p].filled_c = 0; continue; }
      const uint64_t a = tk_stream_start(frame_off, &hdr, p);
      for (int k = 0; k < TK_RING_CHUNKS; ++k) tk_stream_prefill(tk_saddr_of(rings + (size_t)p * TK_RING_BYTES), arena16, a, k);
      tk_stream_open(&ctl[p], a);
      tk_lane_init(lanes[p], tk_saddr_of(rings + (size_t)p * TK_RING_BYTES), tk_saddr_of(&ctl[p]), frame_off, &hdr, p,
                   mbinfo.data(), mb_w, rows);
    }
    for (int p = rows; p < P; ++p) lanes[p].phase = 2;
    const long topup_every = (variant & 4) ? 3000 : 40;   // variant bit 2: a lazy producer (readers find the ring dry)
    long iter = 0;
    for (bool any = true; any;) {
      any = false;
      if (iter++ % topup_every == 0) {
        for (int p = 0; p < P; ++p) {
          const uint32_t nf = tk_stream_topup(tk_saddr_of(&ctl[p]), tk_saddr_of(rings + (size_t)p * TK_RING_BYTES), arena16, ctl[p].filled_c);
          if (nf > ctl[p].filled_c) ctl[p].filled_c = nf;
        }
      }
      for (int p = P - 1; p >= 0; --p) {   // reverse order: exercises the wait-for-progress path
        TokLane& L = lanes[p];
        if (L.phase == 2) continue;
        any = true;
        if (L.phase == 0) tk_mb_start(L, sh, im, rows, P, mbinfo.data());
        if (L.phase == 1) tk_step(L, sh, im, P, mbinfo.data(), coeffs.data());
      }
    }
    for (int p = 0; p < P && p < rows; ++p) if (lanes[p].status != VP8B_OK) hdr.status = lanes[p].status;
  } else {   // rows interleaved over the partitions in dependency order
    const int P = hdr.num_parts;
    std::vector<TokenPart> tp(P);
    std::vector<uint8_t> posprob(VP8B_POSPROB_BYTES);
    for (int k = 0; k < VP8B_POSPROB_BYTES; ++k) posprob[k] = posprob_byte(hdr.prob, k);
    std::vector<uint16_t> topctx((size_t)(P + 1) * mb_w, 0);
    std::vector<int> progress(P, 0);
    for (int p = 0; p < P; ++p) token_part_init(tp[p], frame, &hdr, p);
    for (int my = 0; my < rows; ++my) {
      parse_token_row(tp[my % P], im, &hdr, my % P, my, posprob.data(), topctx.data(), progress.data(), mbinfo.data(), coeffs.data());
    }
    for (int p = 0; p < P && p < rows; ++p) if (tp[p].status != VP8B_OK) hdr.status = tp[p].status;
  }
  if (hdr.status != VP8B_OK) return failed(hdr.status);
  if (g_dump_mbinfo != nullptr && nmb <= g_dump_max_mb) {
    memcpy(g_dump_mbinfo, mbinfo.data(), 16 * nmb);
    memset(g_dump_levels, 0, sizeof(int16_t) * VP8B_COEFFS_PER_MB * nmb);
    for (size_t m = 0; m < nmb; ++m) {
      if (variant & 64) {
        for (uint32_t k = 0; k < mbtok[m].count; ++k) {
          const uint32_t t = tokens[mbtok[m].first + k];
          const int mag = (int)TF_TOK_MAG(t);
          g_dump_levels[m * VP8B_COEFFS_PER_MB + TF_TOK_BLOCK(t) * 16u + TF_TOK_POS(t)] = (int16_t)((t >> 31) ? -mag : mag);
        }
      } else {
        memcpy(g_dump_levels + m * VP8B_COEFFS_PER_MB, coeffs.data() + m * VP8B_COEFFS_PER_MB, sizeof(int16_t) * VP8B_COEFFS_PER_MB);
      }
    }
  }

  // K3: reconstruction wavefront (lag 2)
  {
    ReconWs ws;
    std::vector<uint8_t> ctxmem(recon_ctx_bytes(mb_w, mb_h) + 64, 0);
    ReconCtx cx;
    recon_ctx_bind(cx, ctxmem.data(), mb_w, mb_h);
    cx.pred4 = &kPred4x[0][0];
    const int steps = mb_w + 2 * (rows - 1);
    for (int d = 0; d < steps; ++d) {
      for (int k = 0; k < rows; ++k) {
        const int my = reverse_steps ? rows - 1 - k : k;
        const int mx = d - 2 * my;
        if (mx < 0 || mx >= mb_w) continue;
        const size_t idx = (size_t)my * mb_w + mx;
        const int16_t* dq6 = hdr.dq[(mbinfo[4 * idx + 3] >> MBW_SEG_SHIFT) & 3];
        if (variant & 64) {
          recon_macroblock(ws, cx, mx, my, mb_w, mbinfo.data() + 4 * idx, nullptr, dq6, yp, up, vp, tokens.data() + mbtok[idx].first, mbtok[idx].count);
        } else {
          recon_macroblock(ws, cx, mx, my, mb_w, mbinfo.data() + 4 * idx, coeffs.data() + idx * VP8B_COEFFS_PER_MB, dq6, yp, up, vp);
        }
      }
    }
  }
  if (unfiltered) memcpy(unfiltered, yuv.data(), yuv.size());

  // options.dithering_strength: the plan (one serial pass), then filter and dither in the reference's order (row by
  // row: FilterRow, then DitherRow, frame_dec.c:424-430); the device interleaves the two inside its wavefront
  const bool dithering = (hdr.dither[0] | hdr.dither[1] | hdr.dither[2] | hdr.dither[3]) != 0;
  if (dithering) {
    std::vector<int8_t> dplane(nmb * 128, 0);
    uint32_t tab[55];
    dither_plan_image(im, &hdr, mbinfo.data(), dplane.data(), tab);
    FilterWs ws;
    for (int my = 0; my < rows; ++my) {
      for (int mx = 0; mx < mb_w && hdr.filter_type > 0; ++mx) {
        const uint32_t w = mbinfo[4 * ((size_t)my * mb_w + mx) + 3];
        const uint8_t* fs = hdr.fstr[(w >> MBW_SEG_SHIFT) & 3][(w & MBW_I4X4) ? 1 : 0];
        filter_macroblock(ws, mx, my, mb_w, hdr.filter_type, fs, (w & MBW_INNER) != 0, yp, up, vp);
      }
      for (int mx = 0; mx < mb_w; ++mx) {
        if (mbinfo[4 * ((size_t)my * mb_w + mx) + 3] & MBW_DITHER) dither_macroblock(mx, my, mb_w, dplane.data(), up, vp);
      }
    }
  } else
  // K4: loop-filter wavefront
  if (hdr.filter_type > 0) {
    FilterWs ws;
    const int steps = mb_w + 2 * (rows - 1);
    for (int d = 0; d < steps; ++d) {
      for (int k = 0; k < rows; ++k) {
        const int my = reverse_steps ? rows - 1 - k : k;
        const int mx = d - 2 * my;
        if (mx < 0 || mx >= mb_w) continue;
        const uint32_t w = mbinfo[4 * ((size_t)my * mb_w + mx) + 3];
        const uint8_t* fs = hdr.fstr[(w >> MBW_SEG_SHIFT) & 3][(w & MBW_I4X4) ? 1 : 0];
        filter_macroblock(ws, mx, my, mb_w, hdr.filter_type, fs, (w & MBW_INNER) != 0, yp, up, vp);
      }
    }
  }

  // K6: ALPH chunk (header pass, pixel pass, inverse transforms + unfilter on one emulated thread)
  std::vector<uint8_t> alpha_plane;
  const uint8_t* alpha = nullptr;
  if (c.has_alph_chunk) {
    const uint8_t* alph = data + c.alpha_offset;
    const uint32_t alph_size = (uint32_t)c.alpha_size;
    std::vector<uint32_t> tdata;
    AlphaHdr ah;
    std::vector<uint32_t> coded;
    emu_alpha_passes(alph, alph_size, im, &ah, coded, tdata);
    if (ah.status != AL_OK) return ah.status;
    alpha_plane.assign((size_t)im.width * im.height, 0);
    alph_finish(&ah, alph + 1, coded.data(), tdata.data(), im.width, im.height, im.crop_y, alpha_plane.data(), 0, 1);
    if (g_emu_alpha_dither > 0 && ((alph[0] >> 4) & 3) == 1) {   // options.alpha_dithering_strength, alpha_dec.c:200-230
      std::vector<uint16_t> pre((size_t)im.out_w * im.out_h);
      std::vector<int16_t> lut(2047);
      uint32_t used[256];
      alph_smooth(alpha_plane.data() + (size_t)im.crop_y * im.width + im.crop_x, im.width, im.out_w, im.out_h,
                  g_emu_alpha_dither > 100 ? 100 : g_emu_alpha_dither, pre.data(), lut.data(), used, 0, 1);
    }
    alpha = alpha_plane.data();
  }

  // K5: output (window re-based like k_emit)
  const int w = im.out_w, h = im.out_h;
  const uint8_t* wy = yp + (size_t)im.crop_y * (16 * mb_w) + im.crop_x;
  const uint8_t* wu = up + (size_t)(im.crop_y >> 1) * (8 * mb_w) + (im.crop_x >> 1);
  const uint8_t* wv = vp + (size_t)(im.crop_y >> 1) * (8 * mb_w) + (im.crop_x >> 1);
  if (alpha) alpha += (size_t)im.crop_y * im.width + im.crop_x;
  if (im.dst_w != 0) {
    const int dw = im.dst_w, uvdw = (dw + 1) / 2;
    const int items = (csp == MODE_YUV || csp == MODE_YUVA) ? dw + 2 * uvdw + (csp == MODE_YUVA ? dw : 0) : dw;
    if (csp == MODE_YUV || csp == MODE_YUVA) im.out_stride = dw;
    for (int t = 0; t < items; ++t) emit_scaled_column(im, wy, wu, wv, alpha, out, t);
    return VP8_STATUS_OK;
  }
  if (csp == MODE_YUV || csp == MODE_YUVA) {
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    if (out_size < (size_t)w * h + 2 * (size_t)uvw * uvh + (csp == MODE_YUVA ? (size_t)w * h : 0)) return VP8_STATUS_INVALID_PARAM;
    im.out_stride = w;
    for (int plane = 0; plane < (csp == MODE_YUVA ? 4 : 3); ++plane) {
      const int pw = (plane == 1 || plane == 2) ? uvw : w, ph = (plane == 1 || plane == 2) ? uvh : h;
      for (int j = 0; j < ph; ++j) for (int q = 0; q < (pw + 15) / 16; ++q) emit_yuv_chunk(im, wy, wu, wv, alpha, out, plane, q, j);
    }
  } else if (emit_uses_pairs(csp, flags, im.crop_x)) {
    for (int t = 0; t <= h / 2; ++t) for (int q = 0; q < (w + 7) / 8; ++q) emit_rgba_pair8(im, wy, wu, wv, alpha, out, q, t);
  } else {
    for (int j = 0; j < h; ++j) for (int q = 0; q < (w + 3) / 4; ++q) emit_rgb_quad(im, wy, wu, wv, alpha, out, q, j);
  }
  return VP8_STATUS_OK;
}

extern "C" void emu_set_alpha_dithering(int strength) { g_emu_alpha_dither = strength; }

extern "C" int emu_decode_dithered(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                                   int stride, int strength, int crop_x, int crop_y, int crop_w, int crop_h) {
  g_emu_dither_f = strength < 0 ? 0 : strength > 100 ? 255 : strength * 255 / 100;
  const int st = emu_decode_crop(data, size, csp, flags, out, out_size, stride, 0, nullptr, crop_x, crop_y, crop_w, crop_h);
  g_emu_dither_f = 0;
  return st;
}
