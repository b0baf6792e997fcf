// vp8_kernels.h -- launchers of the kernels in vp8_kernels.cu (C interface for the host driver).
#ifndef LIBWEBP_B200_VP8_KERNELS_H_
#define LIBWEBP_B200_VP8_KERNELS_H_

#include <cuda_runtime.h>

#include "vp8_dev.h"

#ifdef __cplusplus
extern "C" {
#endif

// Raises the dynamic shared-memory limits for the largest image of the batch (in macroblocks).
cudaError_t vp8k_configure(int max_mb_w, int max_mb_h);

// Images [first, first+count) of the wave. All launches are asynchronous on `s`.
void vp8k_parse_modes(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                      int first, int count, int max_mb_w);
// `ids` = device array of `count` image indices that all have P token partitions.
void vp8k_parse_tokens(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                       int16_t* coeffs, const int* ids, int count, int P, int max_mb_w);
void vp8k_reconstruct(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, uint32_t* mbinfo, const int16_t* coeffs,
                      uint8_t* yuv, int first, int count, int max_mb_w, int max_mb_h);
void vp8k_loop_filter(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint32_t* mbinfo, uint8_t* yuv,
                      int first, int count);
// max_units = largest per-image work-item count (RGB: ceil(w/4)*h; YUV: 16-byte chunks of the three planes).
void vp8k_emit(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, uint8_t* out, int first,
               int count, int max_units);

#ifdef __cplusplus
}
#endif
#endif  // LIBWEBP_B200_VP8_KERNELS_H_
