// vp8_oracle.h -- TEST INFRASTRUCTURE ONLY (the checker, never the product).
//
// Plain-C restatement of the reference's lossy (VP8 key-frame) decode path, whole-frame and sequential:
//   container walk   src/dec/webp_dec.c:54-412      frame/segment/filter/partition/quant/proba headers
//                    src/dec/vp8_dec.c:162-395, src/dec/quant_dec.c:62-112, src/dec/tree_dec.c:515-538
//   boolean decoder  src/utils/bit_reader_inl_utils.h:107-157, bit_reader_utils.c:35-118
//   intra modes      src/dec/tree_dec.c:290-367       coefficients  src/dec/vp8_dec.c:400-635
//   reconstruction   src/dec/frame_dec.c:71-196 + src/dsp/dec.c:22-474
//   loop filter      src/dec/frame_dec.c:203-313 + src/dsp/dec.c:484-693
//   output           src/dec/io_dec.c:25-109, src/dsp/upsampling.c:37-93, src/dsp/yuv.h:59-144, yuv.c:22-63
// Pinned against the compiled reference (oracle/_ref/libwebp_ref.so) by tests/test_oracle_pin.py: byte-equal
// RGBA/RGB/BGRA/YUV outputs (+bypass_filtering, no_fancy_upsampling) on examples/test.webp and on seeded
// synthetic corpora covering simple/normal filter, 1/4 segments, 1..8 token partitions and odd sizes.
// Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load this library.
#ifndef ORACLE_VP8_ORACLE_H_
#define ORACLE_VP8_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// Status values = VP8StatusCode (src/webp/decode.h:235-244).
enum { VP8O_OK = 0, VP8O_OUT_OF_MEMORY, VP8O_INVALID_PARAM, VP8O_BITSTREAM_ERROR, VP8O_UNSUPPORTED_FEATURE,
       VP8O_SUSPENDED, VP8O_USER_ABORT, VP8O_NOT_ENOUGH_DATA };

// WEBP_CSP_MODE subset (src/webp/decode.h:150-163).
enum { VP8O_RGB = 0, VP8O_RGBA = 1, VP8O_BGR = 2, VP8O_BGRA = 3, VP8O_ARGB = 4, VP8O_rgbA = 7, VP8O_bgrA = 8,
       VP8O_Argb = 9, VP8O_YUV = 11 };

enum { VP8O_FLAG_BYPASS_FILTER = 1, VP8O_FLAG_NO_FANCY = 2 };

// feat5 = {width, height, has_alpha, has_animation, format}; mirrors WebPGetFeatures.
int vp8o_features(const uint8_t* data, size_t size, int* feat5);

// Mirrors WebPDecode into caller memory. RGB-family: out has stride*h bytes. VP8O_YUV: y|u|v, tight strides.
int vp8o_decode(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size, int stride);

// Stage dump for kernel debugging. All arrays are malloc'ed by vp8o_dump and released by vp8o_dump_free.
typedef struct {
  int status;
  int width, height, mb_w, mb_h;
  int filter_type;          // 0 off, 1 simple, 2 normal
  int num_parts;
  int dq[4][6];             // per segment: y1 dc/ac, y2 dc/ac, uv dc/ac
  uint8_t* modes;           // mb_w*mb_h*20: imodes[16], is_i4x4, uvmode, skip, segment
  int16_t* coeffs;          // mb_w*mb_h*384, dequantised, WHT already folded into the luma DCs
  uint32_t* nz;             // mb_w*mb_h*2: non_zero_y, non_zero_uv (src/dec/vp8i_dec.h:150-158)
  uint8_t* finfo;           // mb_w*mb_h*4: limit, ilevel, inner, hev_thresh
  int y_stride, uv_stride;  // = 16*mb_w, 8*mb_w
  uint8_t* unfiltered;      // y (16mb_w x 16mb_h) | u | v planes before the loop filter
  uint8_t* filtered;        // same layout after the loop filter
} Vp8oDump;

int vp8o_dump(const uint8_t* data, size_t size, Vp8oDump* d);
// Length of every dependent chain of the frame, in boolean decodes: out[0] first partition, out[1..8] token
// partitions, out[9] their number.
int vp8o_count_decodes(const uint8_t* data, size_t size, uint64_t out[10]);
void vp8o_dump_free(Vp8oDump* d);

#ifdef __cplusplus
}
#endif
#endif  // ORACLE_VP8_ORACLE_H_
