// Host build (g++ -DVP8_EMU, like tests/emu) of the two HEADER parses, for one question only: a call whose output request
// is illegal (crop window, scaling, colourspace, caller's buffer) AND whose file is damaged -- which status comes first?
// The reference parses the frame header (VP8GetHeaders) or the whole VP8L header (VP8LDecodeHeader: transforms, colour
// cache, meta prefix image, every group's codes) before it looks at the request (webp_dec.c:469-481), so a header failure
// wins over VP8_STATUS_INVALID_PARAM; anything that goes wrong later in the data loses. plan_item (vp8_batch.cu) asks here
// only after it has found the request illegal: no pixel is decoded, nothing of this runs for a call the device will see.
#define VP8_EMU
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "vp8_parse_core.h"
#include "vp8l_alpha_core.h"

extern "C" __attribute__((visibility("hidden")))
int vp8b_host_headers_status(const uint8_t* frame, size_t frame_size, uint32_t part0_size, int width, int height, int is_lossless) {
  if (!is_lossless) {
    ImgDesc im;
    memset(&im, 0, sizeof(im));
    im.vp8_size = (uint32_t)frame_size; im.part0_size = part0_size;
    im.width = (uint16_t)width; im.height = (uint16_t)height;
    im.mb_w = (uint16_t)((width + 15) >> 4); im.mb_h = (uint16_t)((height + 15) >> 4);
    im.out_w = im.width; im.out_h = im.height;
    FrameHdr* h = (FrameHdr*)calloc(1, sizeof(FrameHdr));
    if (h == NULL) return VP8B_OK;
    RefBits br;   // byte-wise, exact on every stream, never reads outside [frame, frame + frame_size)
    const int st = parse_frame_header(br, frame, im, h);
    free(h);
    return st;
  }
  // VP8L: pass A + the group codes of pass B, tables built into a throw-away area
  const size_t meta_px = AL_META_PIXELS_BOUND(width, height);
  uint8_t* scratch = (uint8_t*)malloc(AL_SCRATCH_BYTES + 64);
  uint32_t* meta = (uint32_t*)malloc(4 * (meta_px + 8));
  uint32_t* tdata = (uint32_t*)malloc(4 * (2 * meta_px + 8));
  AlphaHdr* hd = (AlphaHdr*)malloc(sizeof(AlphaHdr));
  int st = AL_OK;
  if (scratch != NULL && meta != NULL && tdata != NULL && hd != NULL) {
    uint8_t* sc16 = (uint8_t*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);
    alph_parse_header(frame, (uint32_t)frame_size, width, height, sc16, (uint16_t*)meta, tdata, hd, 1);
    st = hd->status;
    if (st == AL_OK) {
      uint32_t* cache = (uint32_t*)sc16 + AL_SUB_TABLE_ENTRIES;
      AlScratch* sc = (AlScratch*)(cache + (1 << AL_MAX_CACHE_BITS) + 256);
      LBits b;
      b.buf = frame; b.len = (uint32_t)frame_size; b.val = hd->br_val; b.pos = hd->br_pos; b.bit_pos = hd->br_bit_pos; b.eos = 0;
      AlGroup unused;
      for (int g = 0; g < hd->num_groups && st == AL_OK; ++g) {
        if (al_read_group(b, hd->cache_bits, (uint32_t*)sc16, hd->group_entries, &unused, sc) == 0) st = AL_BITSTREAM_ERROR;
      }
    }
    // every failure of a whole-picture VP8L header is a bitstream error (vp8l_dec.c:1672-1704)
    if (st != AL_OK) st = AL_BITSTREAM_ERROR;
  }
  free(scratch); free(meta); free(tdata); free(hd);
  return st;
}
