/* webp_api.c -- the reference's public decode entry points (include/webp/decode.h), host side, plain C.
 * Every decode is a batch of one handed to WebPDecodeBatch() (vp8_batch.cu); there is no CPU decode path.
 *
 * Mirrors: src/dec/webp_dec.c:526-800 (simple API, WebPGetInfo, WebPGetFeatures, WebPInitDecoderConfig,
 * WebPDecode), src/dec/buffer_dec.c:229-260 (WebPInitDecBuffer, WebPFreeDecBuffer), src/utils/utils.c
 * (WebPMalloc/WebPFree), src/dsp/cpu.c:176 (the VP8GetCPUInfo hook that dwebp -noasm clears). */
#include <stdlib.h>
#include <string.h>

#include "vp8_container.h"
#include "webp/decode_batch.h"

/* dwebp writes NULL here for -noasm (examples/dwebp.c:285-287). Nothing in this library reads it: there is
 * no host dsp code to select. The symbol exists so the reference tools link unchanged. */
typedef int (*VP8CPUInfo)(int feature);
static int NoCpuFeature(int feature) { (void)feature; return 0; }
WEBP_EXTERN VP8CPUInfo VP8GetCPUInfo;
VP8CPUInfo VP8GetCPUInfo = NoCpuFeature;

int WebPGetDecoderVersion(void) { return (1 << 16) | (3 << 8) | 2; }

void* WebPMalloc(size_t size) { return malloc(size); }
void WebPFree(void* ptr) { free(ptr); }

int WebPInitDecBufferInternal(WebPDecBuffer* buffer, int version) {
  if (WEBP_ABI_IS_INCOMPATIBLE(version, WEBP_DECODER_ABI_VERSION)) return 0;
  if (buffer == NULL) return 0;
  memset(buffer, 0, sizeof(*buffer));
  return 1;
}

void WebPFreeDecBuffer(WebPDecBuffer* buffer) {
  if (buffer == NULL) return;
  if (buffer->is_external_memory <= 0) free(buffer->private_memory);
  buffer->private_memory = NULL;
}

VP8StatusCode WebPGetFeaturesInternal(const uint8_t* data, size_t data_size, WebPBitstreamFeatures* features,
                                      int version) {
  if (WEBP_ABI_IS_INCOMPATIBLE(version, WEBP_DECODER_ABI_VERSION)) return VP8_STATUS_INVALID_PARAM;
  if (features == NULL || data == NULL) return VP8_STATUS_INVALID_PARAM;
  return vp8b_get_features(data, data_size, features);
}

int WebPGetInfo(const uint8_t* data, size_t data_size, int* width, int* height) {
  WebPBitstreamFeatures f;
  if (data == NULL || vp8b_get_features(data, data_size, &f) != VP8_STATUS_OK) return 0;
  if (width != NULL) *width = f.width;
  if (height != NULL) *height = f.height;
  return 1;
}

int WebPInitDecoderConfigInternal(WebPDecoderConfig* config, int version) {
  if (WEBP_ABI_IS_INCOMPATIBLE(version, WEBP_DECODER_ABI_VERSION)) return 0;
  if (config == NULL) return 0;
  memset(config, 0, sizeof(*config));
  return 1;
}

VP8StatusCode WebPDecode(const uint8_t* data, size_t data_size, WebPDecoderConfig* config) {
  WebPBatchItem item;
  if (config == NULL) return VP8_STATUS_INVALID_PARAM;
  item.data = data;
  item.data_size = data_size;
  item.config = config;
  item.status = VP8_STATUS_OK;
  WebPDecodeBatch(&item, 1, NULL);
  return item.status;
}

/* ---- simple API ------------------------------------------------------------------------------------ */
static uint8_t* DecodeAlloc(WEBP_CSP_MODE mode, const uint8_t* data, size_t size, int* width, int* height,
                            WebPDecBuffer* keep) {
  WebPDecoderConfig config;
  if (!WebPInitDecoderConfig(&config)) return NULL;
  config.output.colorspace = mode;
  if (WebPDecode(data, size, &config) != VP8_STATUS_OK) return NULL;
  if (width != NULL) *width = config.output.width;
  if (height != NULL) *height = config.output.height;
  if (keep != NULL) *keep = config.output;
  return WebPIsRGBMode(mode) ? config.output.u.RGBA.rgba : config.output.u.YUVA.y;
}

uint8_t* WebPDecodeRGBA(const uint8_t* d, size_t n, int* w, int* h) { return DecodeAlloc(MODE_RGBA, d, n, w, h, NULL); }
uint8_t* WebPDecodeARGB(const uint8_t* d, size_t n, int* w, int* h) { return DecodeAlloc(MODE_ARGB, d, n, w, h, NULL); }
uint8_t* WebPDecodeBGRA(const uint8_t* d, size_t n, int* w, int* h) { return DecodeAlloc(MODE_BGRA, d, n, w, h, NULL); }
uint8_t* WebPDecodeRGB(const uint8_t* d, size_t n, int* w, int* h) { return DecodeAlloc(MODE_RGB, d, n, w, h, NULL); }
uint8_t* WebPDecodeBGR(const uint8_t* d, size_t n, int* w, int* h) { return DecodeAlloc(MODE_BGR, d, n, w, h, NULL); }

uint8_t* WebPDecodeYUV(const uint8_t* data, size_t size, int* width, int* height, uint8_t** u, uint8_t** v,
                       int* stride, int* uv_stride) {
  WebPDecBuffer out;
  uint8_t* const y = DecodeAlloc(MODE_YUV, data, size, width, height, &out);
  if (y != NULL) {
    if (u != NULL) *u = out.u.YUVA.u;
    if (v != NULL) *v = out.u.YUVA.v;
    if (stride != NULL) *stride = out.u.YUVA.y_stride;
    if (uv_stride != NULL) *uv_stride = out.u.YUVA.u_stride;
  }
  return y;
}

static uint8_t* DecodeIntoRGB(WEBP_CSP_MODE mode, const uint8_t* data, size_t size, uint8_t* out, size_t out_size,
                              int stride) {
  WebPDecoderConfig config;
  if (out == NULL || !WebPInitDecoderConfig(&config)) return NULL;
  config.output.colorspace = mode;
  config.output.is_external_memory = 1;
  config.output.u.RGBA.rgba = out;
  config.output.u.RGBA.stride = stride;
  config.output.u.RGBA.size = out_size;
  return (WebPDecode(data, size, &config) == VP8_STATUS_OK) ? out : NULL;
}

uint8_t* WebPDecodeRGBAInto(const uint8_t* d, size_t n, uint8_t* o, size_t os, int st) { return DecodeIntoRGB(MODE_RGBA, d, n, o, os, st); }
uint8_t* WebPDecodeARGBInto(const uint8_t* d, size_t n, uint8_t* o, size_t os, int st) { return DecodeIntoRGB(MODE_ARGB, d, n, o, os, st); }
uint8_t* WebPDecodeBGRAInto(const uint8_t* d, size_t n, uint8_t* o, size_t os, int st) { return DecodeIntoRGB(MODE_BGRA, d, n, o, os, st); }
uint8_t* WebPDecodeRGBInto(const uint8_t* d, size_t n, uint8_t* o, size_t os, int st) { return DecodeIntoRGB(MODE_RGB, d, n, o, os, st); }
uint8_t* WebPDecodeBGRInto(const uint8_t* d, size_t n, uint8_t* o, size_t os, int st) { return DecodeIntoRGB(MODE_BGR, d, n, o, os, st); }

uint8_t* WebPDecodeYUVInto(const uint8_t* data, size_t size, uint8_t* luma, size_t luma_size, int luma_stride,
                           uint8_t* u, size_t u_size, int u_stride, uint8_t* v, size_t v_size, int v_stride) {
  WebPDecoderConfig config;
  if (luma == NULL || !WebPInitDecoderConfig(&config)) return NULL;
  config.output.colorspace = MODE_YUV;
  config.output.is_external_memory = 1;
  config.output.u.YUVA.y = luma; config.output.u.YUVA.y_stride = luma_stride; config.output.u.YUVA.y_size = luma_size;
  config.output.u.YUVA.u = u; config.output.u.YUVA.u_stride = u_stride; config.output.u.YUVA.u_size = u_size;
  config.output.u.YUVA.v = v; config.output.u.YUVA.v_stride = v_stride; config.output.u.YUVA.v_size = v_size;
  return (WebPDecode(data, size, &config) == VP8_STATUS_OK) ? luma : NULL;
}

/* ---- incremental API: link-compatible, intentionally inert (see include/webp/decode.h) ----------------- */
WebPIDecoder* WebPINewDecoder(WebPDecBuffer* output_buffer) { (void)output_buffer; return NULL; }
WebPIDecoder* WebPIDecode(const uint8_t* data, size_t data_size, WebPDecoderConfig* config) {
  (void)data; (void)data_size; (void)config;
  return NULL;
}
void WebPIDelete(WebPIDecoder* idec) { (void)idec; }
VP8StatusCode WebPIAppend(WebPIDecoder* idec, const uint8_t* data, size_t data_size) {
  (void)idec; (void)data; (void)data_size;
  return VP8_STATUS_INVALID_PARAM;
}
VP8StatusCode WebPIUpdate(WebPIDecoder* idec, const uint8_t* data, size_t data_size) {
  (void)idec; (void)data; (void)data_size;
  return VP8_STATUS_INVALID_PARAM;
}
