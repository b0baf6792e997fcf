// vp8_kernels.h -- launchers of the kernels in vp8_kernels.cu (C interface for the host driver).
#ifndef LIBWEBP_B200_VP8_KERNELS_H_
#define LIBWEBP_B200_VP8_KERNELS_H_

#include <cuda_runtime.h>

#include "vp8_dev.h"

#ifdef __cplusplus
extern "C" {
#endif

// Once per device (the calling thread's current one): opts every kernel that sizes its shared memory at launch into the
// full 227 KB. The attribute belongs to the device's context, so nothing about it may live in process-wide statics.
cudaError_t vp8k_init_device(void);

// Images [first, first+count) of the wave. All launches are asynchronous on `s`.
void vp8k_parse_modes(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                      int first, int count, int max_mb_w);
// `ids` = device array of `count` image indices that all have P token partitions.
void vp8k_parse_tokens(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                       int16_t* coeffs, const int* ids, int count, int P, int max_mb_w);
// Row bands: the token parse, the reconstruction, the loop filter and (for the 4-byte RGB family with fancy upsampling)
// the output stage can each take a range of macroblock rows / output row pairs of every image, so that the pixel stages
// and the download of one band run while the next band is parsed. `resume` (one TokResume per image), `resume_ctx`
// (max_mb_w uint16 per image) and `band_ctx` (32 * max_mb_w bytes per image) carry the state across the band launches.
// Whole images: rows [0, INT_MAX), pairs [0, INT_MAX).
typedef struct TokResume { int32_t wp_off; uint32_t V, vlo, nxt, R24; int32_t nbits, last_shift, pad; } TokResume;
int vp8k_tokens_take_bands(int count, int P);
void vp8k_parse_tokens_band(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                            int16_t* coeffs, const int* ids, int count, int max_mb_w, int row_begin, int row_end,
                            TokResume* resume, uint16_t* resume_ctx);
// `tokens` != NULL: the levels come from the token stream (VP8B_TOKENS_PER_MB words reserved per macroblock) and `mbtok`
// (two words per macroblock: first token inside the image's area, count), as vp8k_parse_tokens_stream left them; else
// from the dense plane `coeffs` of the older parsers.
void vp8k_reconstruct(cudaStream_t s, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo, const int16_t* coeffs,
                      uint8_t* yuv, int first, int count, int max_mb_w, int max_mb_h, int row_begin, int row_end, uint8_t* band_ctx,
                      const uint32_t* tokens, const void* mbtok);
// Images flagged VP8B_FLAG_LITERAL_READER (a partition that starts with 0xFF), parsed once more with the reference's reader taken
// literally (vp8_literal.h); after vp8k_parse_modes and vp8k_parse_tokens_stream of the wave, before vp8k_reconstruct.
void vp8k_parse_literal(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                        uint32_t* tokens, void* mbtok, int first, int count, int max_mb_w);
// The default token parser (vp8_tokens_fp.h): lockstep lanes, fp32 boolean decoder, one 32-bit token per non-zero level.
int vp8k_tokens_use_stream(void);
void vp8k_parse_tokens_stream(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                              uint32_t* tokens, void* mbtok, const int* ids, int count, int P, int max_mb_w);
void vp8k_loop_filter(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint32_t* mbinfo, uint8_t* yuv,
                      int first, int count, int max_mb_h, int row_begin, int row_end, const int8_t* dither_plane);
// FrameHdr::status of images [0, count) into mapped page-locked host memory (no copy engine involved).
void vp8k_collect_status(cudaStream_t s, const FrameHdr* hdrs, int* host_statuses, int count);
void vp8k_copy_to_host(cudaStream_t s, const void* src, void* host_dst, size_t bytes);
// options.dithering_strength: after the token parse of the whole image, before the loop filter. dither_plane = 128 bytes
// per macroblock of the wave (NULL to vp8k_loop_filter: no image of the wave asked for dithering).
void vp8k_dither_plan(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, uint32_t* mbinfo, int8_t* dither_plane,
                      int first, int count);
// max_units = largest per-image work-item count (RGB: ceil(w/4)*h; YUV: 16-byte chunks of the three planes).
void vp8k_emit(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, const uint8_t* alpha_arena,
               uint8_t* out, int first, int count, int max_units, int pair_begin, int pair_end);

// Images with options.use_scaling (ImgDesc::dst_w != 0; vp8k_emit skips them). max_items = largest per-image work-item
// count: dst_w for the RGB family, dst_w + 2 * ((dst_w + 1) / 2) (+ dst_w for MODE_YUVA) for the planar modes.
void vp8k_emit_scaled(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, const uint8_t* alpha_arena,
                      uint8_t* out, int first, int count, int max_items);

// ALPH chunks (vp8l_alpha_core.h): `aimgs` = the `count` image indices that carry one, `plans` = where each of them
// works (device addresses). Header pass first; the host then sizes tables / planes from the AlphaHdr it
// reads back, fills the rest of the plans and ImgDesc::alpha_plane, and runs the decode (pixels + unfilter).
struct AlphaHdr;
typedef struct AlphaPlan {
  uint64_t scratch;   // AL_SCRATCH_BYTES: sub-image tables, colour cache, palette, AlScratch
  uint64_t meta;      // 4 bytes per meta-Huffman pixel (upper bound)
  uint64_t tdata;     // tile images of the predictor / cross-colour transforms (upper bound)
  uint64_t tables;    // num_groups * group_entries words                 (after the header pass)
  uint64_t groups;    // num_groups AlGroup                               (after the header pass)
  uint64_t coded;     // xsize * height ARGB words                        (after the header pass)
  uint64_t smooth;    // 2 * out_w * out_h bytes for alpha de-banding (ImgDesc::alpha_dither), 0 = none;
                      // lossless pictures with options.use_scaling: 4 * out_w * out_h bytes, the premultiplied window
} AlphaPlan;
void vp8k_lossless_finish(cudaStream_t s, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans, const struct AlphaHdr* ahdrs,
                          uint8_t* out, int count);
void vp8k_alpha_header(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans,
                       struct AlphaHdr* ahdrs, int count);
void vp8k_alpha_decode(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans,
                       struct AlphaHdr* ahdrs, uint8_t* alpha_arena, int count);


#ifdef __cplusplus
}
#endif
#endif  // LIBWEBP_B200_VP8_KERNELS_H_
