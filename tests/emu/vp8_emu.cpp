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

#define VP8_EMU 1
#define VP8_WAIT_PROGRESS(ptr, need) assert(*(ptr) >= (need))
#define VP8_PUBLISH_PROGRESS(ptr, val) (*(ptr) = (val))
#include "vp8_container.h"
#include "vp8_parse_core.h"
#include "vp8_pixel_core.h"

extern "C" int emu_decode(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size,
                          int stride, int reverse_steps, uint8_t* unfiltered /* optional: y|u|v padded */) {
  Vp8Container c;
  int st = vp8b_parse_container(data, size, 1, &c);
  if (st != VP8_STATUS_OK) return st;
  if (c.has_animation || c.is_lossless || c.has_alph_chunk) return VP8_STATUS_UNSUPPORTED_FEATURE;
  if (c.part0_size > c.frame_size - 10) return VP8_STATUS_NOT_ENOUGH_DATA;

  // input arena with padding on both sides, like the device arena
  std::vector<uint8_t> arena(size + 256, 0xA5);
  memcpy(arena.data() + 64, data, size);
  ImgDesc im;
  memset(&im, 0, sizeof(im));
  im.in_off = 64 + c.frame_offset;
  im.vp8_size = (uint32_t)c.frame_size;
  im.part0_size = c.part0_size;
  im.width = (uint16_t)c.width; im.height = (uint16_t)c.height;
  im.mb_w = (uint16_t)((c.width + 15) >> 4); im.mb_h = (uint16_t)((c.height + 15) >> 4);
  im.csp = (uint8_t)csp; im.flags = (uint8_t)flags; im.out_stride = stride;
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

  // K1: header + intra modes
  {
    BoolDec br;
    std::vector<uint32_t> top(mb_w);
    hdr.status = parse_frame_header(br, frame, im, &hdr);
    if (hdr.status == VP8B_OK) hdr.status = parse_intra_modes(br, im, &hdr, top.data(), mbinfo.data());
  }
  if (hdr.status != VP8B_OK) return hdr.status;
  if (hdr.num_parts != im.num_parts) return -100;   // host pre-scan disagrees with the device parse

  // K2: tokens, rows interleaved over the partitions in dependency order
  {
    const int P = hdr.num_parts;
    std::vector<TokenPart> tp(P);
    std::vector<uint16_t> topctx((size_t)(P + 1) * mb_w, 0);
    std::vector<int> progress(P, 0);
    for (int p = 0; p < P; ++p) token_part_init(tp[p], frame, &hdr, p);
    for (int my = 0; my < mb_h; ++my) {
      parse_token_row(tp[my % P], im, &hdr, my % P, my, hdr.prob, topctx.data(), progress.data(), mbinfo.data(), coeffs.data());
    }
    for (int p = 0; p < P && p < mb_h; ++p) if (tp[p].status != VP8B_OK) hdr.status = tp[p].status;
  }
  if (hdr.status != VP8B_OK) return hdr.status;

  // K3: reconstruction wavefront (lag 2)
  {
    ReconWs ws;
    std::vector<uint8_t> ctxmem(recon_ctx_bytes(mb_w, mb_h) + 64, 0);
    ReconCtx cx;
    recon_ctx_bind(cx, ctxmem.data(), mb_w, mb_h);
    const int steps = mb_w + 2 * (mb_h - 1);
    for (int d = 0; d < steps; ++d) {
      for (int k = 0; k < mb_h; ++k) {
        const int my = reverse_steps ? mb_h - 1 - k : k;
        const int mx = d - 2 * my;
        if (mx < 0 || mx >= mb_w) continue;
        const size_t idx = (size_t)my * mb_w + mx;
        recon_macroblock(ws, cx, mx, my, mb_w, mbinfo.data() + 4 * idx, coeffs.data() + idx * VP8B_COEFFS_PER_MB, yp, up, vp);
      }
    }
  }
  if (unfiltered) memcpy(unfiltered, yuv.data(), yuv.size());

  // K4: loop-filter wavefront
  if (hdr.filter_type > 0) {
    FilterWs ws;
    const int steps = mb_w + 2 * (mb_h - 1);
    for (int d = 0; d < steps; ++d) {
      for (int k = 0; k < mb_h; ++k) {
        const int my = reverse_steps ? mb_h - 1 - k : k;
        const int mx = d - 2 * my;
        if (mx < 0 || mx >= mb_w) continue;
        const uint32_t w = mbinfo[4 * ((size_t)my * mb_w + mx) + 3];
        const uint8_t* fs = hdr.fstr[(w >> MBW_SEG_SHIFT) & 3][(w & MBW_I4X4) ? 1 : 0];
        filter_macroblock(ws, mx, my, mb_w, hdr.filter_type, fs, (w & MBW_INNER) != 0, yp, up, vp);
      }
    }
  }

  // K5: output
  const int w = im.width, h = im.height;
  if (csp == MODE_YUV) {
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    if (out_size < (size_t)w * h + 2 * (size_t)uvw * uvh) return VP8_STATUS_INVALID_PARAM;
    im.out_stride = w;
    for (int plane = 0; plane < 3; ++plane) {
      const int pw = plane ? uvw : w, ph = plane ? uvh : h;
      for (int j = 0; j < ph; ++j) for (int q = 0; q < (pw + 15) / 16; ++q) emit_yuv_chunk(im, yp, up, vp, out, plane, q, j);
    }
  } else {
    for (int j = 0; j < h; ++j) for (int q = 0; q < (w + 3) / 4; ++q) emit_rgb_quad(im, yp, up, vp, out, q, j);
  }
  return VP8_STATUS_OK;
}
