/* webp/decode.h -- decode API of libwebp_b200, the B200-native batched WebP-lossy decoder.
 *
 * Drop-in for the reference's src/webp/decode.h (ABI 0x0209): every struct below has the reference's field
 * order and sizes, every enum the reference's values, every function the reference's name and signature, so
 * that examples/dwebp.c + imageio/ relink against this library unchanged. Each declaration cites the
 * reference line it replaces. What actually runs underneath is different: WebPDecode() is a batch of one on
 * the GPU (see webp/decode_batch.h); there is no CPU decode path.
 *
 * Supported here: everything the reference's decoder takes through this header -- VP8 lossy key frames with or without
 * an ALPH chunk, whole-picture VP8L, all thirteen colourspaces, bypass_filtering, no_fancy_upsampling, cropping,
 * scaling, flip, dithering_strength, alpha_dithering_strength, internal or external output memory, the incremental
 * API (as a buffering shim: pixels appear when the last byte has arrived). A top-level animated file is
 * VP8_STATUS_UNSUPPORTED_FEATURE here as in the reference (webp_dec.c:427-429). Nothing falls back to the host:
 * without a usable CUDA device every decode fails with VP8_STATUS_USER_ABORT. */
#ifndef WEBP_WEBP_DECODE_H_
#define WEBP_WEBP_DECODE_H_

#include "./types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WEBP_DECODER_ABI_VERSION 0x0209 /* reference: decode.h:23 */

typedef struct WebPRGBABuffer WebPRGBABuffer;
typedef struct WebPYUVABuffer WebPYUVABuffer;
typedef struct WebPDecBuffer WebPDecBuffer;
typedef struct WebPIDecoder WebPIDecoder;
typedef struct WebPBitstreamFeatures WebPBitstreamFeatures;
typedef struct WebPDecoderOptions WebPDecoderOptions;
typedef struct WebPDecoderConfig WebPDecoderConfig;

/* 0x010302, like the reference build this replaces (decode.h:41). */
WEBP_EXTERN int WebPGetDecoderVersion(void);

/* Header probe: 1 and the dimensions on success, 0 otherwise (decode.h:49). */
WEBP_EXTERN int WebPGetInfo(const uint8_t* data, size_t data_size, int* width, int* height);

/* Simple API (decode.h:57-96): decode into library-allocated host memory (free with WebPFree). */
WEBP_EXTERN uint8_t* WebPDecodeRGBA(const uint8_t* data, size_t data_size, int* width, int* height);
WEBP_EXTERN uint8_t* WebPDecodeARGB(const uint8_t* data, size_t data_size, int* width, int* height);
WEBP_EXTERN uint8_t* WebPDecodeBGRA(const uint8_t* data, size_t data_size, int* width, int* height);
WEBP_EXTERN uint8_t* WebPDecodeRGB(const uint8_t* data, size_t data_size, int* width, int* height);
WEBP_EXTERN uint8_t* WebPDecodeBGR(const uint8_t* data, size_t data_size, int* width, int* height);
WEBP_EXTERN uint8_t* WebPDecodeYUV(const uint8_t* data, size_t data_size, int* width, int* height,
                                   uint8_t** u, uint8_t** v, int* stride, int* uv_stride);

/* "Into" variants (decode.h:107-133): caller-owned host memory; NULL on failure. */
WEBP_EXTERN uint8_t* WebPDecodeRGBAInto(const uint8_t* data, size_t data_size, uint8_t* output_buffer,
                                        size_t output_buffer_size, int output_stride);
WEBP_EXTERN uint8_t* WebPDecodeARGBInto(const uint8_t* data, size_t data_size, uint8_t* output_buffer,
                                        size_t output_buffer_size, int output_stride);
WEBP_EXTERN uint8_t* WebPDecodeBGRAInto(const uint8_t* data, size_t data_size, uint8_t* output_buffer,
                                        size_t output_buffer_size, int output_stride);
WEBP_EXTERN uint8_t* WebPDecodeRGBInto(const uint8_t* data, size_t data_size, uint8_t* output_buffer,
                                       size_t output_buffer_size, int output_stride);
WEBP_EXTERN uint8_t* WebPDecodeBGRInto(const uint8_t* data, size_t data_size, uint8_t* output_buffer,
                                       size_t output_buffer_size, int output_stride);
WEBP_EXTERN uint8_t* WebPDecodeYUVInto(const uint8_t* data, size_t data_size, uint8_t* luma, size_t luma_size,
                                       int luma_stride, uint8_t* u, size_t u_size, int u_stride, uint8_t* v,
                                       size_t v_size, int v_stride);

/* Output colourspaces; numbering of decode.h:150-163. Lower-case = premultiplied by alpha. */
typedef enum WEBP_CSP_MODE {
  MODE_RGB = 0, MODE_RGBA = 1, MODE_BGR = 2, MODE_BGRA = 3, MODE_ARGB = 4, MODE_RGBA_4444 = 5, MODE_RGB_565 = 6,
  MODE_rgbA = 7, MODE_bgrA = 8, MODE_Argb = 9, MODE_rgbA_4444 = 10,
  MODE_YUV = 11, MODE_YUVA = 12,
  MODE_LAST = 13
} WEBP_CSP_MODE;

static WEBP_INLINE int WebPIsPremultipliedMode(WEBP_CSP_MODE mode) { /* decode.h:166 */
  return mode == MODE_rgbA || mode == MODE_bgrA || mode == MODE_Argb || mode == MODE_rgbA_4444;
}
static WEBP_INLINE int WebPIsAlphaMode(WEBP_CSP_MODE mode) { /* decode.h:171 */
  return mode == MODE_RGBA || mode == MODE_BGRA || mode == MODE_ARGB || mode == MODE_RGBA_4444 ||
         mode == MODE_YUVA || WebPIsPremultipliedMode(mode);
}
static WEBP_INLINE int WebPIsRGBMode(WEBP_CSP_MODE mode) { return mode < MODE_YUV; } /* decode.h:177 */

struct WebPRGBABuffer { /* decode.h:184-188 */
  uint8_t* rgba;
  int stride;  /* bytes between rows */
  size_t size; /* bytes available at rgba */
};

struct WebPYUVABuffer { /* decode.h:190-198 */
  uint8_t *y, *u, *v, *a;
  int y_stride;
  int u_stride, v_stride;
  int a_stride;
  size_t y_size;
  size_t u_size, v_size;
  size_t a_size;
};

struct WebPDecBuffer { /* decode.h:201-217 */
  WEBP_CSP_MODE colorspace;
  int width, height;
  int is_external_memory; /* >0: caller memory described in u, never freed by the library */
  union {
    WebPRGBABuffer RGBA;
    WebPYUVABuffer YUVA;
  } u;
  uint32_t pad[4];
  uint8_t* private_memory; /* library-owned block when is_external_memory == 0 */
};

WEBP_EXTERN int WebPInitDecBufferInternal(WebPDecBuffer*, int); /* decode.h:220 */
static WEBP_INLINE int WebPInitDecBuffer(WebPDecBuffer* buffer) {
  return WebPInitDecBufferInternal(buffer, WEBP_DECODER_ABI_VERSION);
}
WEBP_EXTERN void WebPFreeDecBuffer(WebPDecBuffer* buffer); /* decode.h:230 */

typedef enum VP8StatusCode { /* decode.h:235-244 */
  VP8_STATUS_OK = 0,
  VP8_STATUS_OUT_OF_MEMORY,
  VP8_STATUS_INVALID_PARAM,
  VP8_STATUS_BITSTREAM_ERROR,
  VP8_STATUS_UNSUPPORTED_FEATURE,
  VP8_STATUS_SUSPENDED,
  VP8_STATUS_USER_ABORT,
  VP8_STATUS_NOT_ENOUGH_DATA
} VP8StatusCode;

/* Incremental API (decode.h:281-372), as a buffering shim: same calls, same status protocol
 * (VP8_STATUS_SUSPENDED until the file is complete), one device decode when the last byte has arrived. */
WEBP_EXTERN WebPIDecoder* WebPINewDecoder(WebPDecBuffer* output_buffer);
WEBP_EXTERN WebPIDecoder* WebPIDecode(const uint8_t* data, size_t data_size, WebPDecoderConfig* config);
WEBP_EXTERN void WebPIDelete(WebPIDecoder* idec);
WEBP_EXTERN VP8StatusCode WebPIAppend(WebPIDecoder* idec, const uint8_t* data, size_t data_size);
WEBP_EXTERN VP8StatusCode WebPIUpdate(WebPIDecoder* idec, const uint8_t* data, size_t data_size);
WEBP_EXTERN WebPIDecoder* WebPINewRGB(WEBP_CSP_MODE csp, uint8_t* output_buffer, size_t output_buffer_size, int output_stride); /* decode.h:297 */
WEBP_EXTERN WebPIDecoder* WebPINewYUVA(uint8_t* luma, size_t luma_size, int luma_stride, uint8_t* u, size_t u_size, int u_stride,
                                       uint8_t* v, size_t v_size, int v_stride, uint8_t* a, size_t a_size, int a_stride); /* decode.h:312 */
WEBP_EXTERN WebPIDecoder* WebPINewYUV(uint8_t* luma, size_t luma_size, int luma_stride, uint8_t* u, size_t u_size, int u_stride,
                                      uint8_t* v, size_t v_size, int v_stride); /* decode.h:320 */
WEBP_EXTERN uint8_t* WebPIDecGetRGB(const WebPIDecoder* idec, int* last_y, int* width, int* height, int* stride); /* decode.h:350 */
WEBP_EXTERN uint8_t* WebPIDecGetYUVA(const WebPIDecoder* idec, int* last_y, uint8_t** u, uint8_t** v, uint8_t** a,
                                     int* width, int* height, int* stride, int* uv_stride, int* a_stride); /* decode.h:357 */
static WEBP_INLINE uint8_t* WebPIDecGetYUV(const WebPIDecoder* idec, int* last_y, uint8_t** u, uint8_t** v,
                                           int* width, int* height, int* stride, int* uv_stride) { /* decode.h:364 */
  return WebPIDecGetYUVA(idec, last_y, u, v, NULL, width, height, stride, uv_stride, NULL);
}
WEBP_EXTERN const WebPDecBuffer* WebPIDecodedArea(const WebPIDecoder* idec, int* left, int* top, int* width, int* height); /* decode.h:377 */

struct WebPBitstreamFeatures { /* decode.h:414-422 */
  int width;
  int height;
  int has_alpha;
  int has_animation;
  int format; /* 0 undefined/mixed, 1 lossy, 2 lossless */
  uint32_t pad[5];
};

WEBP_EXTERN VP8StatusCode WebPGetFeaturesInternal(const uint8_t*, size_t, WebPBitstreamFeatures*, int);
static WEBP_INLINE VP8StatusCode WebPGetFeatures(const uint8_t* data, size_t data_size,
                                                 WebPBitstreamFeatures* features) { /* decode.h:439 */
  return WebPGetFeaturesInternal(data, data_size, features, WEBP_DECODER_ABI_VERSION);
}

struct WebPDecoderOptions { /* decode.h:447-462 */
  int bypass_filtering;
  int no_fancy_upsampling;
  int use_cropping;
  int crop_left, crop_top;
  int crop_width, crop_height;
  int use_scaling;
  int scaled_width, scaled_height;
  int use_threads;
  int dithering_strength;
  int flip;
  int alpha_dithering_strength;
  uint32_t pad[5];
};

struct WebPDecoderConfig { /* decode.h:465-469 */
  WebPBitstreamFeatures input;
  WebPDecBuffer output;
  WebPDecoderOptions options;
};

WEBP_EXTERN int WebPInitDecoderConfigInternal(WebPDecoderConfig*, int); /* decode.h:472 */
static WEBP_INLINE int WebPInitDecoderConfig(WebPDecoderConfig* config) {
  return WebPInitDecoderConfigInternal(config, WEBP_DECODER_ABI_VERSION);
}

/* Full decode with options; `config->output` describes host memory (decode.h:497). Batch of one. */
WEBP_EXTERN VP8StatusCode WebPDecode(const uint8_t* data, size_t data_size, WebPDecoderConfig* config);

#ifdef __cplusplus
}
#endif
#endif /* WEBP_WEBP_DECODE_H_ */
