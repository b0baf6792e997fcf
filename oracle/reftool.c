// reftool.c -- TEST INFRASTRUCTURE ONLY. Helpers compiled INTO oracle/_ref/libwebp_ref.so next to the
// unmodified reference objects (see oracle/Makefile). Nothing here is linked into, loaded by or called from
// the product library (libwebp_b200/). Users: tests/, bench.py's cpu_baseline / --impl reference legs and
// __graft_entry__.smoke(), through ctypes.
//
// What it provides on top of the reference's public API (src/webp/decode.h, src/webp/encode.h):
//   * a seeded procedural image generator (the corpus recipe of SURVEY.md 8(d)),
//   * WebPEncode with an explicit WebPConfig (cwebp has no -partitions flag; see src/enc/webp_enc.c:115-121),
//   * single decodes with options (golden outputs for the parity tests),
//   * the CPU baseline: N pthreads, one image per thread at a time, WebPDecode into preallocated external
//     buffers.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "src/webp/decode.h"
#include "src/webp/encode.h"
#include "src/dsp/cpu.h"

extern VP8CPUInfo VP8GetCPUInfo;   // src/dsp/cpu.c:176 (dwebp -noasm nulls it, examples/dwebp.c:285-287)
static VP8CPUInfo g_saved_cpuinfo = NULL;
static int g_saved = 0;

// on=0: plain-C dsp path, on=1: the reference's runtime SIMD dispatch.
void reft_set_simd(int on) {
  if (!g_saved) { g_saved_cpuinfo = VP8GetCPUInfo; g_saved = 1; }
  VP8GetCPUInfo = on ? g_saved_cpuinfo : NULL;
}

// ---------------------------------------------------------------------------------------------------------
// Procedural images: smooth sinusoid fields + noise + flat rectangles, so i16 and i4x4 macroblocks, skipped
// macroblocks and every token category occur. Fully determined by (w, h, seed).
static uint64_t sm64(uint64_t* s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
static int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

// bpp = 3 (RGB) or 4 (RGBA). Alpha (when bpp==4) = f(x)+g(y) random walks: with -alpha_filter best the
// reference encoder then picks the GRADIENT filter.
void reft_synth(int w, int h, uint32_t seed, int bpp, uint8_t* out) {
  uint64_t s = 0x1234567ull + (uint64_t)seed * 0x9e3779b97f4a7c15ull;
  const double p0 = (double)(sm64(&s) % 1000) / 159.0, p1 = (double)(sm64(&s) % 1000) / 159.0;
  const double p2 = (double)(sm64(&s) % 1000) / 159.0;
  const double fx0 = 30.0 + (double)(sm64(&s) % 20), fy0 = 80.0 + (double)(sm64(&s) % 30);
  const double fx1 = 45.0 + (double)(sm64(&s) % 20), fy1 = 25.0 + (double)(sm64(&s) % 10);
  const double sigma = 12.0;
  int x, y, i;
  for (y = 0; y < h; ++y) {
    uint8_t* row = out + (size_t)y * w * bpp;
    for (x = 0; x < w; ++x) {
      const uint64_t r = sm64(&s);
      // Irwin-Hall(4) ~ N(0,1) after scaling: sum of four 16-bit uniforms.
      const double u = (double)((r & 0xffff) + ((r >> 16) & 0xffff) + ((r >> 32) & 0xffff) + (r >> 48));
      const double n = (u / 65536.0 - 2.0) * 1.7320508 * sigma;
      const uint64_t r2 = sm64(&s);
      const double n1 = ((double)((r2 & 0xffff) + ((r2 >> 16) & 0xffff)) / 65536.0 - 1.0) * 2.449 * sigma;
      const double n2 = ((double)(((r2 >> 32) & 0xffff) + (r2 >> 48)) / 65536.0 - 1.0) * 2.449 * sigma;
      const double a = 128.0 + 100.0 * sin(x / fx0 + y / fy0 + p0);
      const double b = 128.0 + 90.0 * cos(x / fx1 - y / fy1 + p1);
      const double c = 128.0 + 80.0 * sin((double)x * y / 60000.0 + p2);
      row[x * bpp + 0] = (uint8_t)clip255((int)lrint(a + n));
      row[x * bpp + 1] = (uint8_t)clip255((int)lrint(b + n1));
      row[x * bpp + 2] = (uint8_t)clip255((int)lrint(c + n2));
    }
  }
  {  // flat rectangles
    const int nrect = (int)((int64_t)w * h / 35000) + 1;
    for (i = 0; i < nrect; ++i) {
      const int rw = 8 + (int)(sm64(&s) % (uint64_t)(w / 10 + 8)), rh = 8 + (int)(sm64(&s) % (uint64_t)(h / 10 + 8));
      const int rx = (int)(sm64(&s) % (uint64_t)w), ry = (int)(sm64(&s) % (uint64_t)h);
      const uint64_t col = sm64(&s);
      const int x1 = rx + rw < w ? rx + rw : w, y1 = ry + rh < h ? ry + rh : h;
      for (y = ry; y < y1; ++y) {
        uint8_t* row = out + (size_t)y * w * bpp;
        for (x = rx; x < x1; ++x) {
          row[x * bpp + 0] = (uint8_t)col; row[x * bpp + 1] = (uint8_t)(col >> 8); row[x * bpp + 2] = (uint8_t)(col >> 16);
        }
      }
    }
  }
  if (bpp == 4) {
    int* f = (int*)malloc(sizeof(int) * (size_t)(w + h));
    int* g = f + w;
    int v = 0, mn = 0, mxf, mxg;
    for (x = 0; x < w; ++x) { v += (int)(sm64(&s) % 7) - 3; f[x] = v; if (v < mn) mn = v; }
    mxf = 1; for (x = 0; x < w; ++x) { f[x] -= mn; if (f[x] > mxf) mxf = f[x]; }
    v = 0; mn = 0;
    for (y = 0; y < h; ++y) { v += (int)(sm64(&s) % 7) - 3; g[y] = v; if (v < mn) mn = v; }
    mxg = 1; for (y = 0; y < h; ++y) { g[y] -= mn; if (g[y] > mxg) mxg = g[y]; }
    for (y = 0; y < h; ++y) {
      uint8_t* row = out + (size_t)y * w * 4;
      for (x = 0; x < w; ++x) {
        row[x * 4 + 3] = (uint8_t)clip255(8 + (120 * f[x] + mxf / 2) / mxf + (120 * g[y] + mxg / 2) / mxg);
      }
    }
    free(f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Encoding with an explicit WebPConfig. Negative values keep the encoder default (WebPConfigInit).
typedef struct {
  float quality;
  int method, segments, filter_type, filter_strength, filter_sharpness, partitions, low_memory;
  int alpha_filtering, alpha_quality, sns_strength;
  int lossless;   // 1: VP8L (quality / method then steer the lossless encoder's effort)
} ReftEncCfg;

size_t reft_encode(const uint8_t* pix, int w, int h, int bpp, const ReftEncCfg* c, uint8_t** out) {
  WebPConfig config;
  WebPPicture pic;
  WebPMemoryWriter wr;
  size_t size = 0;
  *out = NULL;
  if (!WebPConfigInit(&config) || !WebPPictureInit(&pic)) return 0;
  if (c->quality >= 0) config.quality = c->quality;
  if (c->method >= 0) config.method = c->method;
  if (c->segments >= 0) config.segments = c->segments;
  if (c->filter_type >= 0) config.filter_type = c->filter_type;
  if (c->filter_strength >= 0) config.filter_strength = c->filter_strength;
  if (c->filter_sharpness >= 0) config.filter_sharpness = c->filter_sharpness;
  if (c->partitions >= 0) config.partitions = c->partitions;
  if (c->low_memory >= 0) config.low_memory = c->low_memory;
  if (c->alpha_filtering >= 0) config.alpha_filtering = c->alpha_filtering;
  if (c->alpha_quality >= 0) config.alpha_quality = c->alpha_quality;
  if (c->sns_strength >= 0) config.sns_strength = c->sns_strength;
  if (c->lossless > 0) config.lossless = 1;
  if (!WebPValidateConfig(&config)) return 0;
  pic.width = w; pic.height = h;
  if (c->lossless > 0) pic.use_argb = 1;
  if (bpp == 4 ? !WebPPictureImportRGBA(&pic, pix, w * 4) : !WebPPictureImportRGB(&pic, pix, w * 3)) return 0;
  WebPMemoryWriterInit(&wr);
  pic.writer = WebPMemoryWrite;
  pic.custom_ptr = &wr;
  if (WebPEncode(&config, &pic)) {
    *out = (uint8_t*)malloc(wr.size);
    if (*out != NULL) { memcpy(*out, wr.mem, wr.size); size = wr.size; }
  }
  WebPMemoryWriterClear(&wr);
  WebPPictureFree(&pic);
  return size;
}

void reft_free(void* p) { free(p); }

typedef struct {
  int n, w, h, bpp, next, nthreads; uint32_t seed0; const ReftEncCfg* cfg;
  uint8_t** outs; size_t* sizes; pthread_mutex_t mu; int failed;
} CorpusJob;

static void* corpus_worker(void* arg) {
  CorpusJob* j = (CorpusJob*)arg;
  uint8_t* pix = (uint8_t*)malloc((size_t)j->w * j->h * j->bpp);
  for (;;) {
    int i;
    pthread_mutex_lock(&j->mu); i = j->next++; pthread_mutex_unlock(&j->mu);
    if (i >= j->n || pix == NULL) break;
    reft_synth(j->w, j->h, j->seed0 + (uint32_t)i, j->bpp, pix);
    j->sizes[i] = reft_encode(pix, j->w, j->h, j->bpp, j->cfg, &j->outs[i]);
    if (j->sizes[i] == 0) j->failed = 1;
  }
  free(pix);
  return NULL;
}

// Image k of the corpus = reft_synth(seed0 + k) encoded with *cfg. Returns 0 on success.
int reft_encode_corpus(int n, int w, int h, int bpp, uint32_t seed0, const ReftEncCfg* cfg, int nthreads,
                       uint8_t** outs, size_t* sizes) {
  CorpusJob j;
  pthread_t th[256];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  memset(&j, 0, sizeof(j));
  j.n = n; j.w = w; j.h = h; j.bpp = bpp; j.seed0 = seed0; j.cfg = cfg; j.outs = outs; j.sizes = sizes;
  pthread_mutex_init(&j.mu, NULL);
  for (t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, corpus_worker, &j);
  for (t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&j.mu);
  return j.failed;
}

// ---------------------------------------------------------------------------------------------------------
// One reference decode through WebPDecode (src/dec/webp_dec.c:752) into caller memory.
// RGB-family modes: out = stride*h bytes. MODE_YUV: out = y (w*h) | u | v with tight strides.
// flags: bit0 bypass_filtering, bit1 no_fancy_upsampling, bit2 use_threads, bit3 flip.
// Returns the VP8StatusCode.
int reft_decode(const uint8_t* data, size_t size, int csp, int flags, uint8_t* out, size_t out_size, int stride) {
  WebPDecoderConfig cfg;
  int st;
  if (!WebPInitDecoderConfig(&cfg)) return -1;
  st = WebPGetFeatures(data, size, &cfg.input);
  if (st != VP8_STATUS_OK) return WebPDecode(data, size, &cfg);   /* the status WebPDecode itself reports */
  cfg.options.bypass_filtering = flags & 1;
  cfg.options.no_fancy_upsampling = (flags >> 1) & 1;
  cfg.options.use_threads = (flags >> 2) & 1;
  cfg.options.flip = (flags >> 3) & 1;
  cfg.output.colorspace = (WEBP_CSP_MODE)csp;
  cfg.output.is_external_memory = 1;
  if (csp == MODE_YUV || csp == MODE_YUVA) {
    const int w = cfg.input.width, h = cfg.input.height;
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    const size_t need = (size_t)w * h + 2 * (size_t)uvw * uvh + (csp == MODE_YUVA ? (size_t)w * h : 0);
    if (out_size < need) return -2;
    cfg.output.u.YUVA.y = out; cfg.output.u.YUVA.y_stride = w; cfg.output.u.YUVA.y_size = (size_t)w * h;
    cfg.output.u.YUVA.u = out + (size_t)w * h; cfg.output.u.YUVA.u_stride = uvw; cfg.output.u.YUVA.u_size = (size_t)uvw * uvh;
    cfg.output.u.YUVA.v = cfg.output.u.YUVA.u + (size_t)uvw * uvh; cfg.output.u.YUVA.v_stride = uvw; cfg.output.u.YUVA.v_size = (size_t)uvw * uvh;
    if (csp == MODE_YUVA) {
      cfg.output.u.YUVA.a = cfg.output.u.YUVA.v + (size_t)uvw * uvh; cfg.output.u.YUVA.a_stride = w; cfg.output.u.YUVA.a_size = (size_t)w * h;
    }
  } else {
    cfg.output.u.RGBA.rgba = out; cfg.output.u.RGBA.stride = stride; cfg.output.u.RGBA.size = out_size;
  }
  return WebPDecode(data, size, &cfg);
}

/* WebPDecode with a crop window (crop4 = left, top, width, height) and/or flip, into a tight external buffer of the
 * cropped size. RGB-family and MODE_YUV. */
int reft_decode_window(const uint8_t* data, size_t size, int csp, int flags, const int* crop4, uint8_t* out, size_t out_size) {
  WebPDecoderConfig cfg;
  int st, w, h;
  if (!WebPInitDecoderConfig(&cfg)) return -1;
  st = WebPGetFeatures(data, size, &cfg.input);
  if (st != VP8_STATUS_OK) return WebPDecode(data, size, &cfg);
  cfg.options.bypass_filtering = flags & 1;
  cfg.options.no_fancy_upsampling = (flags >> 1) & 1;
  cfg.options.flip = (flags >> 3) & 1;
  w = cfg.input.width; h = cfg.input.height;
  if (crop4 != NULL && crop4[2] > 0) {
    cfg.options.use_cropping = 1;
    cfg.options.crop_left = crop4[0]; cfg.options.crop_top = crop4[1];
    cfg.options.crop_width = crop4[2]; cfg.options.crop_height = crop4[3];
    w = crop4[2]; h = crop4[3];
  }
  cfg.output.colorspace = (WEBP_CSP_MODE)csp;
  cfg.output.is_external_memory = 1;
  if (csp == MODE_YUV || csp == MODE_YUVA) {
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    if (out_size < (size_t)w * h + 2 * (size_t)uvw * uvh + (csp == MODE_YUVA ? (size_t)w * h : 0)) return -2;
    cfg.output.u.YUVA.y = out; cfg.output.u.YUVA.y_stride = w; cfg.output.u.YUVA.y_size = (size_t)w * h;
    cfg.output.u.YUVA.u = out + (size_t)w * h; cfg.output.u.YUVA.u_stride = uvw; cfg.output.u.YUVA.u_size = (size_t)uvw * uvh;
    cfg.output.u.YUVA.v = cfg.output.u.YUVA.u + (size_t)uvw * uvh; cfg.output.u.YUVA.v_stride = uvw; cfg.output.u.YUVA.v_size = (size_t)uvw * uvh;
    if (csp == MODE_YUVA) {
      cfg.output.u.YUVA.a = cfg.output.u.YUVA.v + (size_t)uvw * uvh; cfg.output.u.YUVA.a_stride = w; cfg.output.u.YUVA.a_size = (size_t)w * h;
    }
  } else {
    const int bpp = (csp == MODE_RGB || csp == MODE_BGR) ? 3
                  : (csp == MODE_RGBA_4444 || csp == MODE_RGB_565 || csp == MODE_rgbA_4444) ? 2 : 4;
    if (out_size < (size_t)w * h * bpp) return -2;
    cfg.output.u.RGBA.rgba = out; cfg.output.u.RGBA.stride = w * bpp; cfg.output.u.RGBA.size = out_size;
  }
  return WebPDecode(data, size, &cfg);
}

/* WebPDecode with options.use_scaling (scaled2 = requested width, height; 0 = keep the ratio), optionally on a crop
 * window, into a tight external buffer; dims2 receives the scaled dimensions the decoder settled on. */
int reft_decode_scaled(const uint8_t* data, size_t size, int csp, int flags, const int* crop4, const int* scaled2, int* dims2,
                       uint8_t* out, size_t out_size) {
  WebPDecoderConfig cfg;
  int st, w, h, sw, sh;
  if (!WebPInitDecoderConfig(&cfg)) return -1;
  st = WebPGetFeatures(data, size, &cfg.input);
  if (st != VP8_STATUS_OK) return WebPDecode(data, size, &cfg);
  cfg.options.bypass_filtering = flags & 1;
  cfg.options.no_fancy_upsampling = (flags >> 1) & 1;
  cfg.options.flip = (flags >> 3) & 1;
  w = cfg.input.width; h = cfg.input.height;
  if (crop4 != NULL && crop4[2] > 0) {
    cfg.options.use_cropping = 1;
    cfg.options.crop_left = crop4[0]; cfg.options.crop_top = crop4[1];
    cfg.options.crop_width = crop4[2]; cfg.options.crop_height = crop4[3];
    w = crop4[2]; h = crop4[3];
  }
  cfg.options.use_scaling = 1;
  cfg.options.scaled_width = scaled2[0]; cfg.options.scaled_height = scaled2[1];
  sw = scaled2[0]; sh = scaled2[1];
  if (sw == 0 && h > 0) sw = (int)(((uint64_t)w * sh + h - 1) / h);
  if (sh == 0 && w > 0) sh = (int)(((uint64_t)h * sw + w - 1) / w);
  dims2[0] = sw; dims2[1] = sh;
  cfg.output.colorspace = (WEBP_CSP_MODE)csp;
  if (sw <= 0 || sh <= 0) return WebPDecode(data, size, &cfg);   /* the status the decoder reports for this request */
  cfg.output.is_external_memory = 1;
  if (csp == MODE_YUV || csp == MODE_YUVA) {
    const int uvw = (sw + 1) / 2, uvh = (sh + 1) / 2;
    if (out_size < (size_t)sw * sh + 2 * (size_t)uvw * uvh + (csp == MODE_YUVA ? (size_t)sw * sh : 0)) return -2;
    cfg.output.u.YUVA.y = out; cfg.output.u.YUVA.y_stride = sw; cfg.output.u.YUVA.y_size = (size_t)sw * sh;
    cfg.output.u.YUVA.u = out + (size_t)sw * sh; cfg.output.u.YUVA.u_stride = uvw; cfg.output.u.YUVA.u_size = (size_t)uvw * uvh;
    cfg.output.u.YUVA.v = cfg.output.u.YUVA.u + (size_t)uvw * uvh; cfg.output.u.YUVA.v_stride = uvw; cfg.output.u.YUVA.v_size = (size_t)uvw * uvh;
    if (csp == MODE_YUVA) {
      cfg.output.u.YUVA.a = cfg.output.u.YUVA.v + (size_t)uvw * uvh; cfg.output.u.YUVA.a_stride = sw; cfg.output.u.YUVA.a_size = (size_t)sw * sh;
    }
  } else {
    const int bpp = (csp == MODE_RGB || csp == MODE_BGR) ? 3
                  : (csp == MODE_RGBA_4444 || csp == MODE_RGB_565 || csp == MODE_rgbA_4444) ? 2 : 4;
    if (out_size < (size_t)sw * sh * bpp) return -2;
    cfg.output.u.RGBA.rgba = out; cfg.output.u.RGBA.stride = sw * bpp; cfg.output.u.RGBA.size = out_size;
  }
  return WebPDecode(data, size, &cfg);
}

/* reft_decode_window with options.dithering_strength = strength and options.alpha_dithering_strength = alpha_strength. */
int reft_decode_dithered(const uint8_t* data, size_t size, int csp, int flags, const int* crop4, int strength, int alpha_strength,
                         uint8_t* out, size_t out_size) {
  WebPDecoderConfig cfg;
  int st, w, h;
  if (!WebPInitDecoderConfig(&cfg)) return -1;
  st = WebPGetFeatures(data, size, &cfg.input);
  if (st != VP8_STATUS_OK) return WebPDecode(data, size, &cfg);
  cfg.options.bypass_filtering = flags & 1;
  cfg.options.no_fancy_upsampling = (flags >> 1) & 1;
  cfg.options.flip = (flags >> 3) & 1;
  cfg.options.dithering_strength = strength;
  cfg.options.alpha_dithering_strength = alpha_strength;
  w = cfg.input.width; h = cfg.input.height;
  if (crop4 != NULL && crop4[2] > 0) {
    cfg.options.use_cropping = 1;
    cfg.options.crop_left = crop4[0]; cfg.options.crop_top = crop4[1];
    cfg.options.crop_width = crop4[2]; cfg.options.crop_height = crop4[3];
    w = crop4[2]; h = crop4[3];
  }
  cfg.output.colorspace = (WEBP_CSP_MODE)csp;
  cfg.output.is_external_memory = 1;
  if (csp == MODE_YUV || csp == MODE_YUVA) {
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    if (out_size < (size_t)w * h + 2 * (size_t)uvw * uvh + (csp == MODE_YUVA ? (size_t)w * h : 0)) return -2;
    cfg.output.u.YUVA.y = out; cfg.output.u.YUVA.y_stride = w; cfg.output.u.YUVA.y_size = (size_t)w * h;
    cfg.output.u.YUVA.u = out + (size_t)w * h; cfg.output.u.YUVA.u_stride = uvw; cfg.output.u.YUVA.u_size = (size_t)uvw * uvh;
    cfg.output.u.YUVA.v = cfg.output.u.YUVA.u + (size_t)uvw * uvh; cfg.output.u.YUVA.v_stride = uvw; cfg.output.u.YUVA.v_size = (size_t)uvw * uvh;
    if (csp == MODE_YUVA) {
      cfg.output.u.YUVA.a = cfg.output.u.YUVA.v + (size_t)uvw * uvh; cfg.output.u.YUVA.a_stride = w; cfg.output.u.YUVA.a_size = (size_t)w * h;
    }
  } else {
    const int bpp = (csp == MODE_RGB || csp == MODE_BGR) ? 3
                  : (csp == MODE_RGBA_4444 || csp == MODE_RGB_565 || csp == MODE_rgbA_4444) ? 2 : 4;
    if (out_size < (size_t)w * h * bpp) return -2;
    cfg.output.u.RGBA.rgba = out; cfg.output.u.RGBA.stride = w * bpp; cfg.output.u.RGBA.size = out_size;
  }
  return WebPDecode(data, size, &cfg);
}

int reft_features(const uint8_t* data, size_t size, int* feat5) {
  WebPBitstreamFeatures f;
  const int st = WebPGetFeatures(data, size, &f);
  feat5[0] = f.width; feat5[1] = f.height; feat5[2] = f.has_alpha; feat5[3] = f.has_animation; feat5[4] = f.format;
  return st;
}

// ---------------------------------------------------------------------------------------------------------
// CPU baseline: nthreads pthreads, static interleaved split of the list, one image per thread at a time,
// each thread decoding into its own preallocated external buffer. Returns wall seconds for `passes` passes
// over the list (after one untimed warm-up pass); *mpix_out gets decoded megapixels per pass.
typedef struct {
  const uint8_t* const* datas; const size_t* sizes; int n, tid, nthreads, csp, passes; size_t max_out;
  int errors; pthread_barrier_t* bar; struct timespec* t0; struct timespec* t1;
} BenchJob;

static void decode_pass(BenchJob* j, uint8_t* buf) {
  int i;
  for (i = j->tid; i < j->n; i += j->nthreads) {
    int f[5];
    if (reft_features(j->datas[i], j->sizes[i], f) != VP8_STATUS_OK) { j->errors++; continue; }
    if (reft_decode(j->datas[i], j->sizes[i], j->csp, 0, buf, j->max_out,
                    f[0] * (j->csp == MODE_RGB || j->csp == MODE_BGR ? 3 : 4)) != VP8_STATUS_OK) j->errors++;
  }
}

static void* bench_worker(void* arg) {
  BenchJob* j = (BenchJob*)arg;
  uint8_t* buf = (uint8_t*)malloc(j->max_out);
  int p;
  memset(buf, 0, j->max_out);        // fault the pages in before timing
  decode_pass(j, buf);               // warm-up
  pthread_barrier_wait(j->bar);
  if (j->tid == 0) clock_gettime(CLOCK_MONOTONIC, j->t0);
  for (p = 0; p < j->passes; ++p) decode_pass(j, buf);
  pthread_barrier_wait(j->bar);
  if (j->tid == 0) clock_gettime(CLOCK_MONOTONIC, j->t1);
  free(buf);
  return NULL;
}

double reft_decode_bench(const uint8_t* const* datas, const size_t* sizes, int n, int nthreads, int csp,
                         int passes, double* mpix_out, int* errors_out) {
  pthread_t th[512];
  BenchJob jobs[512];
  pthread_barrier_t bar;
  struct timespec t0, t1;
  size_t max_out = 0;
  double mpix = 0;
  int i, t, errors = 0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 512) nthreads = 512;
  for (i = 0; i < n; ++i) {
    int f[5];
    if (reft_features(datas[i], sizes[i], f) == VP8_STATUS_OK) {
      const size_t need = (size_t)f[0] * f[1] * 4;
      if (need > max_out) max_out = need;
      mpix += (double)f[0] * f[1] * 1e-6;
    }
  }
  pthread_barrier_init(&bar, NULL, (unsigned)nthreads);
  for (t = 0; t < nthreads; ++t) {
    BenchJob* j = &jobs[t];
    j->datas = datas; j->sizes = sizes; j->n = n; j->tid = t; j->nthreads = nthreads; j->csp = csp;
    j->passes = passes; j->max_out = max_out; j->errors = 0; j->bar = &bar; j->t0 = &t0; j->t1 = &t1;
    pthread_create(&th[t], NULL, bench_worker, j);
  }
  for (t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); errors += jobs[t].errors; }
  pthread_barrier_destroy(&bar);
  if (mpix_out) *mpix_out = mpix;
  if (errors_out) *errors_out = errors;
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
