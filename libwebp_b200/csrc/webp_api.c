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

/* src/utils/utils.c:171-231. The three WebPSafe* entry points are WEBP_EXTERN in the reference (src/utils/utils.h:52-61)
 * because libwebpdemux and libwebpmux allocate through the library they are linked beside; exported here so that the
 * reference's demuxer / WebPAnimDecoder link against this library alone. */
#define WEBP_B200_MAX_ALLOCABLE_MEMORY (1ULL << 34)   /* utils.h:44: WEBP_MAX_ALLOCABLE_MEMORY on 64-bit hosts */
static int alloc_args_ok(uint64_t nmemb, size_t size) {
  if (nmemb == 0) return 1;
  if ((uint64_t)size > WEBP_B200_MAX_ALLOCABLE_MEMORY / nmemb) return 0;
  return 1;   /* size_t is 64-bit on every host this library builds for: nmemb * size <= 2^34 fits */
}
WEBP_EXTERN void* WebPSafeMalloc(uint64_t nmemb, size_t size);
WEBP_EXTERN void* WebPSafeCalloc(uint64_t nmemb, size_t size);
WEBP_EXTERN void WebPSafeFree(void* const ptr);
void* WebPSafeMalloc(uint64_t nmemb, size_t size) { return alloc_args_ok(nmemb, size) ? malloc((size_t)(nmemb * size)) : NULL; }
void* WebPSafeCalloc(uint64_t nmemb, size_t size) { return alloc_args_ok(nmemb, size) ? calloc((size_t)nmemb, size) : NULL; }
void WebPSafeFree(void* const ptr) { free(ptr); }
void* WebPMalloc(size_t size) { return WebPSafeMalloc(1, size); }
void WebPFree(void* ptr) { WebPSafeFree(ptr); }

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
  if (config == NULL || data == NULL) return VP8_STATUS_INVALID_PARAM;   /* webp_dec.c:756; GetFeatures webp_dec.c:693-695 */
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

/* ---- incremental API (src/dec/idec_dec.c:604-909) as a buffering shim ------------------------------------
 * A batch device decoder has nothing to gain from decoding row by row as bytes arrive, so the shim keeps the
 * reference's interface and status protocol and does the work in one go: WebPIAppend() accumulates, WebPIUpdate()
 * looks at the caller's growing buffer, both answer VP8_STATUS_SUSPENDED until the whole file is there (the RIFF
 * size says when; a bare VP8 stream is tried on every call and "not enough data" reads as suspended), then the
 * image is decoded once on the device and the status is VP8_STATUS_OK or the decode error, which sticks.
 * The getters follow the reference's protocol as far as a decoder that produces all rows at once can: NULL while the
 * reference would still be reading headers (idec_dec.c:843-851: state <= STATE_VP8_PARTS0), then the caller's buffer with
 * "no rows yet" (last_y = 0) until the picture is complete, then all rows. A caller that draws data[0 .. stride * last_y)
 * after every append (the prefix comparisons of src/tests.zig:676-686) sees nothing wrong, only later. */
struct WebPIDecoder {
  WebPDecoderConfig* config;     /* caller's config (WebPIDecode) or NULL */
  WebPDecoderConfig own_config;  /* used when the caller gave only an output buffer, or nothing */
  WebPDecBuffer* final_output;   /* WebPINewDecoder(output_buffer): receives the result */
  uint8_t* buf; size_t size, cap; int mode;   /* 0 unset, 1 append, 2 update */
  VP8StatusCode status;          /* SUSPENDED until decoded, then final */
  int early;                     /* 1: the output area exists (the caller's buffer), no rows in it yet; 2: and its dimensions are known */
};

int vp8b_prepare_host_buffer(int w, int h, WebPDecBuffer* b);   /* vp8_batch.cu */
int vp8b_host_headers_status(const uint8_t* frame, size_t frame_size, uint32_t part0_size, int width, int height, int is_lossless);   /* vp8_host_probe.cpp */

static WebPIDecoder* NewIDecoder(WebPDecBuffer* output_buffer, WebPDecoderConfig* config) {
  WebPIDecoder* const idec = (WebPIDecoder*)calloc(1, sizeof(*idec));
  if (idec == NULL) return NULL;
  WebPInitDecoderConfig(&idec->own_config);
  idec->own_config.output.colorspace = MODE_RGB;   /* idec_dec.c: default output when none is given */
  idec->config = config;
  idec->final_output = output_buffer;
  idec->status = VP8_STATUS_SUSPENDED;
  return idec;
}

WebPIDecoder* WebPINewDecoder(WebPDecBuffer* output_buffer) { return NewIDecoder(output_buffer, NULL); }

WebPIDecoder* WebPIDecode(const uint8_t* data, size_t data_size, WebPDecoderConfig* config) {
  WebPBitstreamFeatures tmp;
  WebPBitstreamFeatures* const features = (config == NULL) ? &tmp : &config->input;
  memset(&tmp, 0, sizeof(tmp));
  if (data != NULL && data_size > 0) {
    if (WebPGetFeatures(data, data_size, features) != VP8_STATUS_OK) return NULL;
  }
  return NewIDecoder(NULL, config);
}

void WebPIDelete(WebPIDecoder* idec) {
  if (idec == NULL) return;
  free(idec->buf);
  WebPFreeDecBuffer(&idec->own_config.output);
  free(idec);
}

/* Bytes the container announces for the whole file, 0 when it does not say (bare bitstream / header incomplete). */
static size_t AnnouncedSize(const uint8_t* data, size_t size) {
  if (size >= 12 && !memcmp(data, "RIFF", 4) && !memcmp(data + 8, "WEBP", 4)) {
    const uint32_t riff = (uint32_t)data[4] | ((uint32_t)data[5] << 8) | ((uint32_t)data[6] << 16) | ((uint32_t)data[7] << 24);
    return (size_t)riff + 8;
  }
  return 0;
}

static VP8StatusCode IDecodeNow(WebPIDecoder* idec, const uint8_t* data, size_t size) {
  WebPDecoderConfig* cfg = (idec->config != NULL) ? idec->config : &idec->own_config;
  VP8StatusCode st;
  size_t total;
  if (idec->status != VP8_STATUS_SUSPENDED) return idec->status;
  if (size < 12) return VP8_STATUS_SUSPENDED;
  total = AnnouncedSize(data, size);
  if (total != 0 && size < total) {
    /* still arriving; a header that is already wrong is reported at once, like the reference does */
    WebPBitstreamFeatures f;
    st = WebPGetFeatures(data, size, &f);
    if (st != VP8_STATUS_OK && st != VP8_STATUS_NOT_ENOUGH_DATA) { idec->status = st; return st; }
    if (st == VP8_STATUS_OK && idec->early < 2) {
      /* Past the point where the reference starts answering with an output area (the first partition of a VP8 frame is in,
       * idec_dec.c:DecodePartition0; a VP8L stream: its container header)? Only for a caller's own buffer and a plain request:
       * the size of an area that depends on crop or scaling options is settled by the one decode. */
      Vp8Container c;
      WebPDecBuffer* const out = (idec->config != NULL) ? &idec->config->output
                               : (idec->final_output != NULL) ? idec->final_output : &idec->own_config.output;
      const WebPDecoderOptions* const o = &cfg->options;
      if (out->is_external_memory > 0 && !o->use_cropping && !o->use_scaling &&
          vp8b_parse_container(data, size, 0, &c) == VP8_STATUS_OK && !c.has_animation) {
        int dims = 0;
        if (c.is_lossless) {
          /* the area is there from the container header on (idec state VP8L_HEADER), its dimensions once the VP8L header --
           * transforms, colour cache, every group's codes -- has been decoded from what has arrived (idec_dec.c:DecodeVP8LHeader) */
          idec->early = 1;
          dims = size > c.frame_offset &&
                 vp8b_host_headers_status(data + c.frame_offset, size - c.frame_offset, 0, f.width, f.height, 1) == 0;
        } else {
          dims = size >= c.frame_offset + 10 + (size_t)c.part0_size &&
                 vp8b_last_partition_begun(data + c.frame_offset + 10, c.part0_size, size - c.frame_offset - 10,
                                           vp8b_prescan_partitions(data + c.frame_offset + 10, c.part0_size));
        }
        if (dims && vp8b_prepare_host_buffer(f.width, f.height, out) == VP8_STATUS_OK) idec->early = 2;
      }
    }
    return VP8_STATUS_SUSPENDED;
  }
  if (idec->config == NULL && idec->final_output != NULL) {   /* decode straight into the caller's buffer */
    cfg->output = *idec->final_output;
  }
  st = WebPDecode(data, size, cfg);
  if (st == VP8_STATUS_NOT_ENOUGH_DATA && total == 0) return VP8_STATUS_SUSPENDED;   /* bare stream, more may come */
  if (st == VP8_STATUS_OK && idec->config == NULL && idec->final_output != NULL) {
    *idec->final_output = cfg->output;
    memset(&cfg->output, 0, sizeof(cfg->output));   /* ownership moved */
  }
  idec->status = st;
  return st;
}

VP8StatusCode WebPIAppend(WebPIDecoder* idec, const uint8_t* data, size_t data_size) {
  if (idec == NULL || data == NULL) return VP8_STATUS_INVALID_PARAM;
  if (idec->status != VP8_STATUS_SUSPENDED) return idec->status;
  if (idec->mode == 2) return VP8_STATUS_INVALID_PARAM;   /* no mixing of append and update (idec_dec.c:94) */
  idec->mode = 1;
  if (idec->size + data_size > idec->cap) {
    size_t ncap = idec->cap ? idec->cap : 4096;
    uint8_t* nb;
    while (ncap < idec->size + data_size) ncap *= 2;
    nb = (uint8_t*)realloc(idec->buf, ncap);
    if (nb == NULL) return VP8_STATUS_OUT_OF_MEMORY;
    idec->buf = nb; idec->cap = ncap;
  }
  memcpy(idec->buf + idec->size, data, data_size);
  idec->size += data_size;
  return IDecodeNow(idec, idec->buf, idec->size);
}

VP8StatusCode WebPIUpdate(WebPIDecoder* idec, const uint8_t* data, size_t data_size) {
  if (idec == NULL || data == NULL) return VP8_STATUS_INVALID_PARAM;
  if (idec->status != VP8_STATUS_SUSPENDED) return idec->status;
  if (idec->mode == 1) return VP8_STATUS_INVALID_PARAM;
  idec->mode = 2;
  return IDecodeNow(idec, data, data_size);
}

/* ---- the rest of the incremental interface (idec_dec.c:626-700, 840-909). Constructors that describe the output up
 * front, and the getters: before the image has been decoded there is no displayable area (the reference answers NULL
 * until the first rows are out), afterwards it is the whole picture (last_y = height). */
WebPIDecoder* WebPINewRGB(WEBP_CSP_MODE csp, uint8_t* output_buffer, size_t output_buffer_size, int output_stride) {
  const int is_external_memory = (output_buffer != NULL) ? 1 : 0;
  WebPIDecoder* idec;
  if (csp >= MODE_YUV) return NULL;
  if (is_external_memory == 0) {      /* overrule the size and stride of an absent buffer */
    output_buffer_size = 0;
    output_stride = 0;
  } else if (output_stride == 0 || output_buffer_size == 0) {
    return NULL;
  }
  idec = WebPINewDecoder(NULL);
  if (idec == NULL) return NULL;
  idec->own_config.output.colorspace = csp;
  idec->own_config.output.is_external_memory = is_external_memory;
  idec->own_config.output.u.RGBA.rgba = output_buffer;
  idec->own_config.output.u.RGBA.stride = output_stride;
  idec->own_config.output.u.RGBA.size = output_buffer_size;
  return idec;
}

WebPIDecoder* WebPINewYUVA(uint8_t* luma, size_t luma_size, int luma_stride, uint8_t* u, size_t u_size, int u_stride,
                           uint8_t* v, size_t v_size, int v_stride, uint8_t* a, size_t a_size, int a_stride) {
  const int is_external_memory = (luma != NULL) ? 1 : 0;
  WebPIDecoder* idec;
  WEBP_CSP_MODE colorspace;
  if (is_external_memory == 0) {      /* overrule everything: the library allocates */
    luma_size = u_size = v_size = a_size = 0;
    luma_stride = u_stride = v_stride = a_stride = 0;
    u = v = a = NULL;
    colorspace = MODE_YUVA;
  } else {
    if (u == NULL || v == NULL) return NULL;
    if (luma_size == 0 || u_size == 0 || v_size == 0) return NULL;
    if (luma_stride == 0 || u_stride == 0 || v_stride == 0) return NULL;
    if (a != NULL && (a_size == 0 || a_stride == 0)) return NULL;
    colorspace = (a == NULL) ? MODE_YUV : MODE_YUVA;
  }
  idec = WebPINewDecoder(NULL);
  if (idec == NULL) return NULL;
  idec->own_config.output.colorspace = colorspace;
  idec->own_config.output.is_external_memory = is_external_memory;
  idec->own_config.output.u.YUVA.y = luma; idec->own_config.output.u.YUVA.y_stride = luma_stride; idec->own_config.output.u.YUVA.y_size = luma_size;
  idec->own_config.output.u.YUVA.u = u; idec->own_config.output.u.YUVA.u_stride = u_stride; idec->own_config.output.u.YUVA.u_size = u_size;
  idec->own_config.output.u.YUVA.v = v; idec->own_config.output.u.YUVA.v_stride = v_stride; idec->own_config.output.u.YUVA.v_size = v_size;
  idec->own_config.output.u.YUVA.a = a; idec->own_config.output.u.YUVA.a_stride = a_stride; idec->own_config.output.u.YUVA.a_size = a_size;
  return idec;
}

WebPIDecoder* WebPINewYUV(uint8_t* luma, size_t luma_size, int luma_stride, uint8_t* u, size_t u_size, int u_stride,
                          uint8_t* v, size_t v_size, int v_stride) {
  return WebPINewYUVA(luma, luma_size, luma_stride, u, u_size, u_stride, v, v_size, v_stride, NULL, 0, 0);
}

/* Rows of the output area that hold pixels: none until the one decode, then all of them. */
static int IDecRows(const WebPIDecoder* idec, const WebPDecBuffer* output) {
  return (idec->status == VP8_STATUS_OK) ? output->height : 0;
}

static const WebPDecBuffer* IDecOutput(const WebPIDecoder* idec) {
  if (idec == NULL) return NULL;
  if (idec->status != VP8_STATUS_OK && !(idec->status == VP8_STATUS_SUSPENDED && idec->early)) return NULL;   /* no area yet */
  if (idec->config != NULL) return &idec->config->output;
  if (idec->final_output != NULL) return idec->final_output;
  return &idec->own_config.output;
}

const WebPDecBuffer* WebPIDecodedArea(const WebPIDecoder* idec, int* left, int* top, int* width, int* height) {
  const WebPDecBuffer* const src = IDecOutput(idec);
  if (left != NULL) *left = 0;
  if (top != NULL) *top = 0;
  if (src != NULL) {
    if (width != NULL) *width = src->width;
    if (height != NULL) *height = IDecRows(idec, src);   /* idec_dec.c:869: the decoded area ends at last_y */
  } else {
    if (width != NULL) *width = 0;
    if (height != NULL) *height = 0;
  }
  return src;
}

uint8_t* WebPIDecGetRGB(const WebPIDecoder* idec, int* last_y, int* width, int* height, int* stride) {
  const WebPDecBuffer* const output = IDecOutput(idec);
  if (output == NULL || output->colorspace >= MODE_YUV) return NULL;
  if (last_y != NULL) *last_y = IDecRows(idec, output);
  if (width != NULL) *width = output->width;
  if (height != NULL) *height = output->height;
  if (stride != NULL) *stride = output->u.RGBA.stride;
  return output->u.RGBA.rgba;
}

uint8_t* WebPIDecGetYUVA(const WebPIDecoder* idec, int* last_y, uint8_t** u, uint8_t** v, uint8_t** a,
                         int* width, int* height, int* stride, int* uv_stride, int* a_stride) {
  const WebPDecBuffer* const output = IDecOutput(idec);
  if (output == NULL || output->colorspace < MODE_YUV) return NULL;
  if (last_y != NULL) *last_y = IDecRows(idec, output);
  if (u != NULL) *u = output->u.YUVA.u;
  if (v != NULL) *v = output->u.YUVA.v;
  if (a != NULL) *a = output->u.YUVA.a;
  if (width != NULL) *width = output->width;
  if (height != NULL) *height = output->height;
  if (stride != NULL) *stride = output->u.YUVA.y_stride;
  if (uv_stride != NULL) *uv_stride = output->u.YUVA.u_stride;
  if (a_stride != NULL) *a_stride = output->u.YUVA.a_stride;
  return output->u.YUVA.y;
}
