// vp8_dev.h -- device-side data layout of the batched VP8 decoder (shared by host driver and kernels).
//
// HBM layout of one wave of images (all arrays batch-wide, indexed through ImgDesc::mb_base):
//   input arena   : the compressed files exactly as the caller laid them out (one or few H2D copies)
//   ImgDesc[n]    : host-built geometry + where the VP8 frame of image i starts in the arena
//   FrameHdr[n]   : written by the header/mode kernel: quantisers, probabilities, filter strengths, partitions
//   MbInfo[M]     : 16 B per macroblock: 16 sub-block modes (4 bit each), nz codes, flags
//   coeffs[M*400] : int16 coefficient LEVELS exactly as parsed (not yet dequantised), in parse (zigzag) order inside
//                   each 4x4 block; blocks 0-15 Y, 16-19 U, 20-23 V, 24 = Y2 (the WHT input of i16 macroblocks);
//                   dequantisation + zigzag scatter happen in the reconstruction kernel      (800 B / macroblock)
//   yuv[M*384]    : per image Y (16mb_w x 16mb_h) | U | V, macroblock-padded planes     (1.5 B / pixel)
//   output arena  : RGBA/RGB/... or Y|U|V per image, tight strides
//   alpha arena   : images with an ALPH chunk only: AlphaHdr, lookup tables, coded bytes, w x h alpha plane
//                   (vp8l_alpha_core.h)
#ifndef LIBWEBP_B200_VP8_DEV_H_
#define LIBWEBP_B200_VP8_DEV_H_

#include <stdint.h>

#define VP8B_COEFFS_PER_MB 400
#define VP8B_TOKENS_PER_MB 388   // token stream: at most 16 x 15 + 16 (or 16 x 16) luma + 128 chroma = 384 non-zero levels per macroblock,
                                 // + 4 of slack (a lane whose block has ended writes up to three stray tokens past its last one, vp8_tokens_fp.h)
#define VP8B_MAX_PARTS 8

// VP8StatusCode values used on the device (include/webp/decode.h).
#define VP8B_OK 0
#define VP8B_BITSTREAM_ERROR 3
#define VP8B_UNSUPPORTED 4
#define VP8B_NOT_ENOUGH_DATA 7
#define VP8B_NOT_A_VP8_FRAME (-2)   // FrameHdr::status of a VP8B_FLAG_LOSSLESS image (never leaves the library)

// ImgDesc::flags
#define VP8B_FLAG_BYPASS_FILTER 1
#define VP8B_FLAG_NO_FANCY 2
#define VP8B_FLAG_FLIP 4     // options.flip: output rows bottom-up (WebPFlipBuffer, buffer_dec.c:152-175)
#define VP8B_FLAG_LITERAL_READER 16   // some partition starts with byte 0xFF: parsed again by k_parse_literal (vp8_parse_core.h:RefBits)
#define VP8B_FLAG_LOSSLESS 8 // a whole VP8L picture: no macroblocks (mb_w = mb_h = 0, the VP8 kernels pass it by); alpha_in /
                             // alpha_size locate the VP8L bitstream and the picture leaves through vp8l_lossless_core.h

typedef struct ImgDesc {
  uint64_t in_off;     // byte offset of the VP8 frame tag inside the input arena
  uint64_t out_off;    // byte offset of this image's pixels inside the output arena
  uint32_t vp8_size;   // bytes from the frame tag to the end of the file (reference: io.data_size)
  uint32_t part0_size; // first-partition length from the frame tag
  uint32_t mb_base;    // first macroblock of this image in the per-macroblock arrays
  int32_t out_stride;  // bytes per output row (RGB family) or Y stride (MODE_YUV; U/V stride = (w+1)/2)
  uint16_t width, height, mb_w, mb_h;
  uint8_t csp;         // WEBP_CSP_MODE
  uint8_t flags;
  uint8_t num_parts;   // host pre-scan of the partition count (launch geometry only; FrameHdr is authoritative)
  uint8_t dither_f;    // options.dithering_strength mapped to 0..255 (0 = off), see parse_frame_header
  uint64_t alpha_in;   // byte offset of the ALPH chunk payload inside the input arena (alpha_size == 0: no alpha)
  uint64_t alpha_plane;// byte offset of this image's decoded w x h alpha plane inside the alpha arena (host fills it
                       // after the alpha header pass); VP8B_NO_ALPHA when the image has none
  uint32_t alpha_size; // ALPH payload bytes
  uint32_t alpha_index;// index among the batch's alpha images
  // output window (options.use_cropping, webp_dec.c:802-829; the whole picture otherwise). crop_x / crop_y are even.
  // The window is upsampled and emitted as if it were the picture (edges replicate at the window, io_dec.c:57-109).
  uint16_t crop_x, crop_y, out_w, out_h;
  // options.use_scaling: the window above is rescaled to dst_w x dst_h (io_dec.c:239-556, src/dsp/rescaler.c); 0 = no scaling.
  // out_stride / out_off then describe the scaled picture.
  uint32_t dst_w, dst_h;  // (32 bits: the reference rescales to anything its allocator accepts, rescaler_utils.c:86-118)
  uint8_t alpha_dither;   // options.alpha_dithering_strength (1..100) when the ALPH chunk's levels were quantised, else 0
  uint8_t pad_[7];
} ImgDesc;
#define VP8B_NO_ALPHA 0xffffffffffffffffull

typedef struct FrameHdr {
  int32_t status;               // VP8B_OK or the failure of the header / mode parse; token parse may overwrite
  uint8_t filter_type;          // 0 none, 1 simple, 2 normal (after bypass_filtering)
  uint8_t num_parts;
  uint8_t use_skip, skip_p;
  uint8_t update_map;
  uint8_t seg_prob[3];
  int16_t dq[4][6];             // per segment: y1 dc/ac, y2 dc/ac, uv dc/ac
  uint8_t fstr[4][2][4];        // per segment, per is_i4x4: limit, ilevel, inner, hev_thresh
  uint32_t part_off[VP8B_MAX_PARTS];   // token partitions, relative to the frame tag
  uint32_t part_size[VP8B_MAX_PARTS];
  uint8_t prob[4 * 8 * 3 * 11]; // [type][band][ctx][node]
  uint8_t dither[4];            // per segment: amplitude of the random chroma dithering (VP8InitDithering, frame_dec.c:328-349), 0 = none
  int32_t rows;                 // macroblock rows that get decoded: all of them, or down to the bottom of the crop
                                // window plus the filter's reach (VP8EnterCritical, frame_dec.c:571-596)
  int32_t fail_row;             // where the reference's row loop (vp8_dec.c:646-674) would meet this frame's failure:
                                // VP8B_FAIL_NONE, VP8B_FAIL_HEADERS (before any row), or the macroblock row whose intra
                                // modes (K1) or tokens (K2, the smallest over the partitions) ran out of data
  int32_t modes_status;         // intra modes that ran out of data at a row r > 0: K1 leaves status OK, rows = r and the
                                // failure here, so that K2 still parses the rows above (a token failure there comes first);
                                // every later stage treats the frame as lost (vp8b_frame_lost)
  int32_t all_rows;             // `rows` before K1 lowered it
} FrameHdr;
#define vp8b_frame_lost(h) ((h)->status != VP8B_OK || (h)->modes_status != VP8B_OK)
#define vp8b_frame_status(h) ((h)->status != VP8B_OK ? (h)->status : (h)->modes_status)
#define VP8B_FAIL_NONE 0x7fffffff
#define VP8B_FAIL_HEADERS (-1)

// A file whose ALPH chunk AND VP8 payload are both damaged: the reference decodes alpha rows as the macroblock rows above
// them finish (FinishRow, frame_dec.c:440-460: rows [y_start, y_end) of macroblock row r right after its tokens) and reports
// whichever failure its row loop meets first. vp8_fail_row as in FrameHdr; alpha_fail_row = the ALPHA row whose decode fails
// (-1: the ALPH header, met by the first FinishRow); alpha_all_at_once: the plane's levels were quantised, which makes the
// reference decode all of it at the first request (alpha_dec.c:196-203). Returns 1 when the VP8 failure comes first.
static inline int vp8b_vp8_failure_first(int vp8_fail_row, int alpha_fail_row, int filter_type, int rows, int crop_bottom,
                                         int alpha_all_at_once) {
  const int extra = (filter_type == 2) ? 8 : (filter_type == 1) ? 2 : 0;   /* kFilterExtraRows, frame_dec.c:201 */
  int r;
  if (vp8_fail_row < 0) return 1;                       /* VP8GetHeaders: before the first row */
  if (alpha_fail_row < 0 || alpha_all_at_once) return vp8_fail_row <= 0;   /* alpha fails in FinishRow(0) */
  for (r = 0; r < rows; ++r) {                          /* first macroblock row whose FinishRow asks for the failing alpha row */
    int y_end = 16 * (r + 1) - ((r >= rows - 1) ? 0 : extra);
    if (y_end > crop_bottom) y_end = crop_bottom;
    if (y_end > alpha_fail_row) break;
  }
  return vp8_fail_row <= r;
}

// MbInfo words (one uint4 per macroblock):
//   x, y : sixteen 4-bit sub-block modes, mode n at bits 4*(n&7) of word n>>3 (i16: mode in the low nibble of x)
//   z    : non_zero_y, 2 bits per luma block, block 0 in the top bits (reference: vp8i_dec.h:150-158)
//   w    : bits 0-15 non_zero_uv | bit16 is_i4x4 | bits17-18 uvmode | bit19 skip | bits20-21 segment
//          | bit22 has_y2 (i16 with a non-empty Y2 block) | bit23 filter-inner (set by the reconstruction)
//          | bit24 dithered (set by the dither plan)
#define MBW_I4X4 (1u << 16)
#define MBW_UVMODE_SHIFT 17
#define MBW_SKIP (1u << 19)
#define MBW_SEG_SHIFT 20
#define MBW_HAS_Y2 (1u << 22)
#define MBW_INNER (1u << 23)
#define MBW_DITHER (1u << 24)   // chroma gets dithered; its 128 offsets sit in the dither plane (dither_plan_image)

// intra modes (RFC 6386 numbering as used by the reference, src/dec/common_dec.h:18-40)
enum { M_DC = 0, M_TM = 1, M_VE = 2, M_HE = 3, M_RD = 4, M_VR = 5, M_LD = 6, M_VL = 7, M_HD = 8, M_HU = 9 };

#endif  // LIBWEBP_B200_VP8_DEV_H_
