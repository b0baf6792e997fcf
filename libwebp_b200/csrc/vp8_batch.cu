// vp8_batch.cu -- host driver of the batched decoder: per-item validation, device memory, uploads, kernel
// sequencing per wave, status/pixel downloads. Public entry points: include/webp/decode_batch.h.
//
// Replaces the per-image control flow of DecodeInto (src/dec/webp_dec.c:447-523): WebPParseHeaders ->
// VP8GetHeaders -> WebPAllocateDecBuffer -> VP8Decode, and the output-buffer contract of
// src/dec/buffer_dec.c:41-228 (CheckDecBuffer / AllocateBuffer). There is no CPU decode path: without a
// usable CUDA device every item fails and WebPBatchLastError() says why.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "vp8_container.h"
#include "vp8_kernels.h"
#include "vp8l_alpha_core.h"   // AlphaHdr, AlGroup and the sizing macros only: nothing of it runs on the host
#include "webp/decode_batch.h"

// ---------------------------------------------------------------------------------------------------------
static thread_local char g_last_error[256] = "";

static void set_error(const char* what, cudaError_t e) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, e == cudaSuccess ? "failed" : cudaGetErrorString(e));
}

#define CU_TRY(call, what)            \
  do {                                \
    cudaError_t e_ = (call);          \
    if (e_ != cudaSuccess) {          \
      set_error((what), e_);          \
      return false;                   \
    }                                 \
  } while (0)

extern "C" const char* WebPBatchLastError(void) { return g_last_error; }

extern "C" int WebPBatchDeviceCount(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void* WebPBatchHostAlloc(size_t size) {
  void* p = NULL;
  if (cudaHostAlloc(&p, size, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return NULL; }
  return p;
}
extern "C" void WebPBatchHostFree(void* ptr) { if (ptr != NULL) cudaFreeHost(ptr); }

extern "C" int WebPBatchOptionsInitInternal(WebPBatchOptions* o, int version) {
  if (o == NULL || WEBP_ABI_IS_INCOMPATIBLE(version, WEBP_BATCH_ABI_VERSION)) return 0;
  memset(o, 0, sizeof(*o));
  o->device = -1;
  o->output = WEBP_BATCH_HOST;
  return 1;
}

// ---------------------------------------------------------------------------------------------------------
// Per-device state: two streams and a small cache of device allocations so that repeated WebPDecode() calls
// do not pay cudaMalloc/cudaFree every time.
struct CachedBlock { void* p; size_t cap; };

struct DeviceCtx {
  int device = -1;
  bool ok = false;
  cudaStream_t stream = nullptr;        // kernels, uploads
  cudaStream_t copy_stream = nullptr;   // pixel downloads, overlapped with the kernels of the next wave
  cudaStream_t pixel_stream = nullptr;  // row bands: reconstruction / filter / output of one band while the next is parsed
  std::mutex mu;          // one batch at a time per device
  std::vector<CachedBlock> cache;
  size_t cached_bytes = 0;
  size_t cache_limit = 0;   // released blocks are kept for the next batch up to this many bytes
};

static std::mutex g_ctx_mu;
static DeviceCtx* g_ctx[64];

static DeviceCtx* get_ctx(int device) {
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) { set_error("cudaGetDevice (no CUDA device; this library has no CPU path)", cudaGetLastError()); return nullptr; }
  }
  if (device >= 64) { set_error("device ordinal out of range", cudaSuccess); return nullptr; }
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (g_ctx[device] == nullptr) {
    DeviceCtx* c = new DeviceCtx();
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->pixel_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error("CUDA device init (no usable GPU; this library has no CPU path)", e); cudaGetLastError(); delete c; return nullptr; }
    {
      size_t free_b = 0, total_b = 0;
      const char* env = getenv("WEBP_B200_CACHE_GB");
      if (env != NULL) c->cache_limit = (size_t)atof(env) << 30;
      else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c->cache_limit = (size_t)((double)free_b * 0.80);
      cudaGetLastError();
    }
    c->ok = true;
    g_ctx[device] = c;
  }
  return g_ctx[device];
}

static void* dev_alloc(DeviceCtx* c, size_t bytes) {
  if (bytes == 0) bytes = 256;
  int best = -1;
  for (size_t i = 0; i < c->cache.size(); ++i) {
    if (c->cache[i].cap >= bytes && c->cache[i].cap <= 2 * bytes + (1 << 20) &&
        (best < 0 || c->cache[i].cap < c->cache[best].cap)) best = (int)i;
  }
  if (best >= 0) {
    void* p = c->cache[best].p;
    c->cached_bytes -= c->cache[best].cap;
    c->cache.erase(c->cache.begin() + best);
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {   // drop the cache and retry once
    cudaGetLastError();
    for (auto& b : c->cache) cudaFree(b.p);
    c->cache.clear(); c->cached_bytes = 0;
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc", e); cudaGetLastError(); return nullptr; }
  }
  return p;
}

struct Owned { void* p = nullptr; size_t cap = 0; };

static bool own_alloc(DeviceCtx* c, Owned& o, size_t bytes) {
  o.cap = bytes ? bytes : 256;
  o.p = dev_alloc(c, o.cap);
  return o.p != nullptr;
}

static void own_free(DeviceCtx* c, Owned& o) {
  if (o.p == nullptr) return;
  if (c->cached_bytes + o.cap <= c->cache_limit && c->cache.size() < 64) {   // cudaFree/cudaMalloc of tens of GB cost 0.1-1 s
    c->cache.push_back({ o.p, o.cap });
    c->cached_bytes += o.cap;
  } else {
    cudaFree(o.p);
  }
  o.p = nullptr; o.cap = 0;
}

// ---------------------------------------------------------------------------------------------------------
static const int kBpp[MODE_LAST] = { 3, 4, 3, 4, 4, 2, 2, 4, 4, 4, 2, 1, 1 };   // buffer_dec.c:24-27

static bool csp_supported(int csp) { return csp >= MODE_RGB && csp < MODE_LAST; }   // all thirteen of decode.h:150-163

// CheckDecBuffer (buffer_dec.c:41-84) for host output.
static VP8StatusCode check_host_buffer(const WebPDecBuffer* b) {
  const int w = b->width, h = b->height;
  bool ok = true;
  if (!WebPIsRGBMode(b->colorspace)) {
    const WebPYUVABuffer* y = &b->u.YUVA;
    const int uvw = (w + 1) / 2, uvh = (h + 1) / 2;
    const uint64_t ys = (uint64_t)abs(y->y_stride), us = (uint64_t)abs(y->u_stride), vs = (uint64_t)abs(y->v_stride);
    ok &= (ys * (h - 1) + w <= y->y_size) && (us * (uvh - 1) + uvw <= y->u_size) && (vs * (uvh - 1) + uvw <= y->v_size);
    ok &= ((int)ys >= w) && ((int)us >= uvw) && ((int)vs >= uvw);
    ok &= (y->y != NULL) && (y->u != NULL) && (y->v != NULL);
    if (b->colorspace == MODE_YUVA) {
      const uint64_t as = (uint64_t)abs(y->a_stride);
      ok &= (as * (h - 1) + w <= y->a_size) && ((int)as >= w) && (y->a != NULL);
    }
  } else {
    const WebPRGBABuffer* r = &b->u.RGBA;
    const uint64_t st = (uint64_t)abs(r->stride);
    const uint64_t row = (uint64_t)w * kBpp[b->colorspace];
    ok &= (st * (h - 1) + row <= r->size) && (st >= row) && (r->rgba != NULL);
  }
  return ok ? VP8_STATUS_OK : VP8_STATUS_INVALID_PARAM;
}

// WebPAllocateDecBuffer / AllocateBuffer (buffer_dec.c:87-227) for host output, without crop/scale/flip.
static VP8StatusCode prepare_host_buffer(int w, int h, WebPDecBuffer* b) {
  if (b == NULL || w <= 0 || h <= 0) return VP8_STATUS_INVALID_PARAM;
  b->width = w; b->height = h;
  const int csp = b->colorspace;
  if (csp < MODE_RGB || csp >= MODE_LAST) return VP8_STATUS_INVALID_PARAM;
  if (b->is_external_memory <= 0 && b->private_memory == NULL) {
    if ((uint64_t)w * kBpp[csp] >= (1ull << 31)) return VP8_STATUS_INVALID_PARAM;
    const int stride = w * kBpp[csp];
    const uint64_t size = (uint64_t)stride * h;
    uint64_t uv_size = 0, a_size = 0;
    int uv_stride = 0;
    if (!WebPIsRGBMode((WEBP_CSP_MODE)csp)) { uv_stride = (w + 1) / 2; uv_size = (uint64_t)uv_stride * ((h + 1) / 2); }
    if (csp == MODE_YUVA) a_size = (uint64_t)w * h;
    const uint64_t total = size + 2 * uv_size + a_size;
    if (total >= (1ull << 34)) return VP8_STATUS_OUT_OF_MEMORY;   // WEBP_MAX_ALLOCABLE_MEMORY, utils.h:34-41
    uint8_t* mem = (uint8_t*)malloc((size_t)total);
    if (mem == NULL) return VP8_STATUS_OUT_OF_MEMORY;
    b->private_memory = mem;
    if (!WebPIsRGBMode((WEBP_CSP_MODE)csp)) {
      WebPYUVABuffer* y = &b->u.YUVA;
      y->y = mem; y->y_stride = stride; y->y_size = (size_t)size;
      y->u = mem + size; y->u_stride = uv_stride; y->u_size = (size_t)uv_size;
      y->v = mem + size + uv_size; y->v_stride = uv_stride; y->v_size = (size_t)uv_size;
      y->a = (csp == MODE_YUVA) ? mem + size + 2 * uv_size : NULL; y->a_size = (size_t)a_size; y->a_stride = (csp == MODE_YUVA) ? w : 0;
    } else {
      b->u.RGBA.rgba = mem; b->u.RGBA.stride = stride; b->u.RGBA.size = (size_t)size;
    }
  }
  return check_host_buffer(b);
}

// ---------------------------------------------------------------------------------------------------------
struct ItemPlan {
  int img = -1;            // index among the device images, -1 when the item failed on the host
  size_t frame_offset = 0; // VP8 frame tag inside the file
  size_t out_bytes = 0;    // bytes of this image in the device output arena (tight strides)
};

struct Wave {
  int first = 0, count = 0;
  size_t mbs = 0;
  int max_mb_w = 0, max_mb_h = 0, max_units = 0;
  int max_scaled_items = 0;   // images with options.use_scaling: work items of vp8k_emit_scaled (0 = none in this wave)
  int ids_off[4] = { 0, 0, 0, 0 }, ids_cnt[4] = { 0, 0, 0, 0 };   // per log2(P) slice of the ids array
};

struct HostRange { const uint8_t* base; size_t size; size_t dev_off; };

struct WebPBatch {
  DeviceCtx* ctx = nullptr;
  WebPBatchItem* items = nullptr;
  int n = 0;
  WebPBatchOptions opt;
  std::vector<ItemPlan> plan;
  std::vector<ImgDesc> imgs;       // device images, wave-major
  std::vector<int> img_item;       // device image -> item
  std::vector<Wave> waves;
  std::vector<int> ids;            // token-parse launch lists
  std::vector<int> statuses;       // host copy of FrameHdr::status
  Owned d_in, d_imgs, d_hdrs, d_ids, d_mbinfo, d_coeffs, d_yuv, d_out;
  Owned d_dither; // options.dithering_strength: 128 offsets per macroblock of a wave (allocated when an item asks for it)
  bool any_dither = false;
  bool any_lossless = false;   // whole-picture VP8L images ride the ALPH machinery (VP8B_FLAG_LOSSLESS)
  Owned d_band;   // row bands: TokResume[m] | uint16 top contexts [m][max_mb_w] | unfiltered top pixels [m][32 * max_mb_w]
  // images with an ALPH chunk
  std::vector<int> aimgs;              // their image indices
  std::vector<AlphaPlan> aplans;
  std::vector<AlphaHdr> ahdrs;         // host copy after the header pass / after the decode
  Owned d_aimgs, d_aplans, d_ahdrs, d_awork, d_awork2, d_alpha;
  bool alpha_planned = false;          // work areas sized from the headers (kept across repeated decodes)
  size_t out_total = 0;
  int max_mb_w = 1, max_mb_h = 1;
  std::vector<cudaEvent_t> ev;     // pool of timing / hand-off events, grown on demand
  size_t ev_used = 0;
  struct Span { int stage, a, b; };
  std::vector<Span> spans;         // (stage, first event, second event) of every timed launch of the last decode
  WebPBatchTimings timings;
  bool decoded = false;
};

static void fail_all(WebPBatchItem* items, int n, VP8StatusCode st) {
  for (int i = 0; i < n; ++i) if (items[i].status == VP8_STATUS_OK) items[i].status = st;
}

// Host-side part of one item: container walk, option/colourspace screening, output buffer.
static VP8StatusCode plan_item(WebPBatchItem* it, const WebPBatchOptions& opt, Vp8Container* c) {
  WebPDecoderConfig* cfg = it->config;
  if (cfg == NULL) return VP8_STATUS_INVALID_PARAM;
  VP8StatusCode st = vp8b_get_features(it->data, it->data_size, &cfg->input);
  if (st != VP8_STATUS_OK) return st == VP8_STATUS_NOT_ENOUGH_DATA ? VP8_STATUS_BITSTREAM_ERROR : st;   // webp_dec.c:761-767
  st = (VP8StatusCode)vp8b_parse_container(it->data, it->data_size, 1, c);
  if (st != VP8_STATUS_OK) return st;
  if (c->has_animation) return VP8_STATUS_UNSUPPORTED_FEATURE;             // webp_dec.c:427-429
  if (!c->is_lossless && c->part0_size > c->frame_size - 10) return VP8_STATUS_NOT_ENOUGH_DATA;   // vp8_dec.c:345-348
  const WebPDecoderOptions* o = &cfg->options;
  const int csp = cfg->output.colorspace;
  if (csp < MODE_RGB || csp >= MODE_LAST) return VP8_STATUS_INVALID_PARAM;
  if (!csp_supported(csp)) return VP8_STATUS_UNSUPPORTED_FEATURE;
  int ow = c->width, oh = c->height;
  if (o->use_cropping) {   // WebPAllocateDecBuffer, buffer_dec.c:184-195 (x, y snapped to even like the decoder's own io)
    const int x = o->crop_left & ~1, y = o->crop_top & ~1;
    const int cw = o->crop_width, ch = o->crop_height;
    if (x < 0 || y < 0 || cw <= 0 || ch <= 0 || x >= ow || cw > ow || cw > ow - x || y >= oh || ch > oh || ch > oh - y)
      return VP8_STATUS_INVALID_PARAM;
    // a lossless picture is cropped at the offsets as given, not snapped (WebPIoInitFromOptions snaps for YUV420 sources
    // only, webp_dec.c:809-817; VP8LDecodeImage hands it MODE_BGRA, vp8l_dec.c:1722-1726): the window must fit there too
    if (c->is_lossless && (o->crop_left >= ow || cw > ow - o->crop_left || o->crop_top >= oh || ch > oh - o->crop_top))
      return VP8_STATUS_INVALID_PARAM;
    ow = cw; oh = ch;
  }
  if (o->use_scaling) {   // WebPAllocateDecBuffer, buffer_dec.c:197-205; WebPRescalerGetScaledDimensions, rescaler_utils.c:86-118
    int sw = o->scaled_width, sh = o->scaled_height;
    if (sw == 0 && oh > 0) sw = (int)(((uint64_t)ow * sh + oh - 1) / oh);
    if (sh == 0 && ow > 0) sh = (int)(((uint64_t)oh * sw + ow - 1) / ow);
    if (sw <= 0 || sh <= 0 || sw > 0x3fffffff || sh > 0x3fffffff) return VP8_STATUS_INVALID_PARAM;
    if (sw > 16383 || sh > 16383) return VP8_STATUS_UNSUPPORTED_FEATURE;   // ImgDesc keeps 16-bit dimensions
    ow = sw; oh = sh;
  }
  if (opt.output == WEBP_BATCH_HOST) return prepare_host_buffer(ow, oh, &cfg->output);
  cfg->output.width = ow; cfg->output.height = oh;
  return VP8_STATUS_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int final_w(const ImgDesc& d) { return d.dst_w ? d.dst_w : d.out_w; }   // the picture that leaves the device
static inline int final_h(const ImgDesc& d) { return d.dst_h ? d.dst_h : d.out_h; }

// Host pass: everything that can be decided without a GPU. Returns the number of items still alive.
static int batch_plan(WebPBatch* b, std::vector<Vp8Container>& cont) {
  const int n = b->n;
  int alive = 0;
  b->plan.resize(n);
  cont.resize(n);
  for (int i = 0; i < n; ++i) {
    b->items[i].status = plan_item(&b->items[i], b->opt, &cont[i]);
    if (b->items[i].status == VP8_STATUS_OK) ++alive;
    else if (b->items[i].config != NULL && b->opt.output == WEBP_BATCH_HOST) {
      WebPFreeDecBuffer(&b->items[i].config->output);   // webp_dec.c:513-515
    }
  }
  return alive;
}

static bool batch_build(WebPBatch* b, const std::vector<Vp8Container>& cont) {
  DeviceCtx* ctx = b->ctx;
  const int n = b->n;
  // ---- input ranges: merge host buffers that sit (almost) next to each other into single H2D copies
  std::vector<int> order;
  for (int i = 0; i < n; ++i) if (b->items[i].status == VP8_STATUS_OK) order.push_back(i);
  std::vector<int> by_addr(order);
  std::sort(by_addr.begin(), by_addr.end(), [&](int a, int c2) { return b->items[a].data < b->items[c2].data; });
  std::vector<HostRange> ranges;
  std::vector<int> item_range(n, -1);
  size_t in_total = 256;
  for (int i : by_addr) {
    const uint8_t* p = b->items[i].data;
    const size_t sz = b->items[i].data_size;
    if (!ranges.empty()) {
      HostRange& r = ranges.back();
      if (p >= r.base && p <= r.base + r.size + 4096) {
        const size_t end = (size_t)(p - r.base) + sz;
        if (end > r.size) r.size = end;
        item_range[i] = (int)ranges.size() - 1;
        continue;
      }
    }
    ranges.push_back({ p, sz, 0 });
    item_range[i] = (int)ranges.size() - 1;
  }
  for (auto& r : ranges) { r.dev_off = in_total; in_total += align_up(r.size, 256) + 256; }
  // ---- device images, output offsets
  size_t total_mbs = 0;
  for (int i : order) {
    const Vp8Container& c = cont[i];
    const WebPDecoderConfig* cfg = b->items[i].config;
    ImgDesc d;
    memset(&d, 0, sizeof(d));
    d.in_off = ranges[item_range[i]].dev_off + (size_t)(b->items[i].data - ranges[item_range[i]].base) + c.frame_offset;
    d.vp8_size = (uint32_t)c.frame_size;
    d.part0_size = c.part0_size;
    d.width = (uint16_t)c.width; d.height = (uint16_t)c.height;
    if (!c.is_lossless) { d.mb_w = (uint16_t)((c.width + 15) >> 4); d.mb_h = (uint16_t)((c.height + 15) >> 4); }
    d.csp = (uint8_t)cfg->output.colorspace;
    d.flags = (uint8_t)((cfg->options.bypass_filtering ? VP8B_FLAG_BYPASS_FILTER : 0) |
                        (cfg->options.no_fancy_upsampling ? VP8B_FLAG_NO_FANCY : 0) |
                        (cfg->options.flip ? VP8B_FLAG_FLIP : 0));
    d.out_w = d.width; d.out_h = d.height;
    if (cfg->options.use_cropping) {
      d.crop_x = (uint16_t)(cfg->options.crop_left & ~1); d.crop_y = (uint16_t)(cfg->options.crop_top & ~1);
      d.out_w = (uint16_t)cfg->options.crop_width; d.out_h = (uint16_t)cfg->options.crop_height;
      // a lossless picture is cropped where the caller said (WebPIoInitFromOptions only snaps YUV420 sources, webp_dec.c:816-820)
      if (c.is_lossless) { d.crop_x = (uint16_t)cfg->options.crop_left; d.crop_y = (uint16_t)cfg->options.crop_top; }
    }
    int fw = d.out_w, fh = d.out_h;   // the picture that leaves the device
    if (cfg->options.use_scaling) {
      fw = cfg->output.width; fh = cfg->output.height;   // set by plan_item
      d.dst_w = (uint16_t)fw; d.dst_h = (uint16_t)fh;
      // WebPIoInitFromOptions, webp_dec.c:851-856: no loop filter for large downscaling ratios (against the whole picture)
      if (fw < c.width * 3 / 4 && fh < c.height * 3 / 4) d.flags |= VP8B_FLAG_BYPASS_FILTER;
    }
    const int ds = cfg->options.dithering_strength;
    d.dither_f = (uint8_t)(ds < 0 ? 0 : ds > 100 ? 255 : ds * 255 / 100);
    if (d.dither_f != 0) b->any_dither = true;
    d.num_parts = c.is_lossless ? 0 : (uint8_t)vp8b_prescan_partitions(b->items[i].data + c.frame_offset + 10, c.part0_size);
    d.alpha_plane = VP8B_NO_ALPHA;
    if (c.is_lossless) {   // the VP8L passes find the bitstream through the alpha fields; no VP8 kernel touches the image
      d.flags |= VP8B_FLAG_LOSSLESS;
      d.dither_f = 0;
      d.alpha_in = d.in_off;
      d.alpha_size = (uint32_t)c.frame_size;
      d.alpha_index = (uint32_t)b->aimgs.size();
      b->aimgs.push_back((int)b->imgs.size());
      b->any_lossless = true;
    } else if (c.has_alph_chunk) {
      d.alpha_in = ranges[item_range[i]].dev_off + (size_t)(b->items[i].data - ranges[item_range[i]].base) + c.alpha_offset;
      d.alpha_size = (uint32_t)c.alpha_size;
      d.alpha_index = (uint32_t)b->aimgs.size();
      // alpha de-banding only ever runs on planes whose levels were quantised by the encoder (ALPH header:
      // pre-processing = 1, alpha_dec.c:71,200-210); dwebp asks for it by default
      const int ad = cfg->options.alpha_dithering_strength;
      if (ad > 0 && c.alpha_size > 0 && ((b->items[i].data[c.alpha_offset] >> 4) & 3) == 1) d.alpha_dither = (uint8_t)(ad > 100 ? 100 : ad);
      b->aimgs.push_back((int)b->imgs.size());
    }
    size_t bytes;
    if (d.csp == MODE_YUV || d.csp == MODE_YUVA) {
      d.out_stride = fw;
      bytes = (size_t)fw * fh + 2 * (size_t)((fw + 1) / 2) * ((fh + 1) / 2) + (d.csp == MODE_YUVA ? (size_t)fw * fh : 0);
    } else {
      d.out_stride = fw * kBpp[d.csp];
      bytes = (size_t)d.out_stride * fh;
    }
    d.out_off = b->out_total;
    b->out_total += align_up(bytes, 256);
    b->plan[i].img = (int)b->imgs.size();
    b->plan[i].frame_offset = c.frame_offset;
    b->plan[i].out_bytes = bytes;
    b->imgs.push_back(d);
    b->img_item.push_back(i);
    total_mbs += (size_t)d.mb_w * d.mb_h;
    b->max_mb_w = std::max(b->max_mb_w, (int)d.mb_w);
    b->max_mb_h = std::max(b->max_mb_h, (int)d.mb_h);
  }
  const int m = (int)b->imgs.size();
  b->statuses.assign(m, 0);
  if (m == 0) return true;

  CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
  // ---- resident allocations: input (64 KB tail padding, see vp8_tokens_fsm.h:tk_lane_init), descriptors, headers, output
  if (!own_alloc(ctx, b->d_in, in_total + 65536) || !own_alloc(ctx, b->d_imgs, sizeof(ImgDesc) * m) ||
      !own_alloc(ctx, b->d_hdrs, sizeof(FrameHdr) * m) || !own_alloc(ctx, b->d_ids, sizeof(int) * m) ||
      !own_alloc(ctx, b->d_out, b->out_total + 256)) return false;
  // ---- waves: per-macroblock scratch = 16 (MbInfo) + 800 (coefficients) + 384 (planes) bytes
  const size_t per_mb = 16 + 2 * VP8B_COEFFS_PER_MB + 384;
  size_t budget = b->opt.scratch_bytes;
  if (budget == 0) {
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
    budget = (size_t)((double)(free_b + ctx->cached_bytes) * 0.85);
  }
  size_t wave_cap_mbs = std::max<size_t>(budget / per_mb, (size_t)b->max_mb_w * b->max_mb_h);
  {
    // More than one wave only on request (or when scratch memory forces it): the token parse wants every stream
    // it can get in flight, and measured end to end one wave + chunked pixel stages beats two waves whose
    // downloads overlap the second parse (profiles/r01*_e2e_sweep.log).
    int waves = b->opt.pipeline_waves;
    if (waves <= 0) {
      static int env_waves = -1;
      if (env_waves < 0) { const char* e = getenv("WEBP_B200_HOST_WAVES"); env_waves = e ? atoi(e) : 0; }
      waves = env_waves > 0 ? env_waves : 1;
    }
    if (waves > 1) wave_cap_mbs = std::min(wave_cap_mbs, std::max<size_t>((total_mbs + waves - 1) / waves, (size_t)b->max_mb_w * b->max_mb_h));
  }
  {
    Wave w;
    for (int k = 0; k < m; ++k) {
      const size_t mbs = (size_t)b->imgs[k].mb_w * b->imgs[k].mb_h;
      if (w.count > 0 && w.mbs + mbs > wave_cap_mbs) { b->waves.push_back(w); w = Wave(); w.first = k; }
      b->imgs[k].mb_base = (uint32_t)w.mbs;
      w.mbs += mbs; w.count++;
      w.max_mb_w = std::max(w.max_mb_w, (int)b->imgs[k].mb_w);
      w.max_mb_h = std::max(w.max_mb_h, (int)b->imgs[k].mb_h);
      const ImgDesc& d = b->imgs[k];
      // work items of the output kernel (must match k_emit / emit_uses_pairs in vp8_pixel_core.h)
      const bool pairs = !(d.flags & VP8B_FLAG_NO_FANCY) && (d.crop_x & 7) == 0 && kBpp[d.csp] == 4;
      const int units = (d.csp == MODE_YUV || d.csp == MODE_YUVA)
                            ? (d.csp == MODE_YUVA ? 2 : 1) * ((d.out_w + 15) / 16) * d.out_h + 2 * ((((d.out_w + 1) / 2) + 15) / 16) * ((d.out_h + 1) / 2)
                            : pairs ? ((d.out_w + 7) / 8) * (d.out_h / 2 + 1)
                                    : ((d.out_w + 3) / 4) * d.out_h;
      w.max_units = std::max(w.max_units, units);
      if (d.dst_w != 0) {
        const int uvdw = (d.dst_w + 1) / 2;
        w.max_scaled_items = std::max(w.max_scaled_items, (d.csp == MODE_YUV || d.csp == MODE_YUVA) ? d.dst_w + 2 * uvdw + (d.csp == MODE_YUVA ? d.dst_w : 0) : (int)d.dst_w);
      }
    }
    b->waves.push_back(w);
  }
  size_t max_wave_mbs = 0;
  b->ids.resize(m);
  for (auto& w : b->waves) {
    max_wave_mbs = std::max(max_wave_mbs, w.mbs);
    int pos = w.first;
    for (int lg = 0; lg < 4; ++lg) {
      w.ids_off[lg] = pos;
      for (int k = w.first; k < w.first + w.count; ++k) if (b->imgs[k].num_parts == (1 << lg)) b->ids[pos++] = k;
      w.ids_cnt[lg] = pos - w.ids_off[lg];
    }
  }
  if (!own_alloc(ctx, b->d_mbinfo, max_wave_mbs * 16) || !own_alloc(ctx, b->d_coeffs, max_wave_mbs * 2 * VP8B_COEFFS_PER_MB) ||
      !own_alloc(ctx, b->d_yuv, max_wave_mbs * 384)) return false;
  if (b->any_dither && !own_alloc(ctx, b->d_dither, max_wave_mbs * 128 + 256)) return false;
  if (!own_alloc(ctx, b->d_band, align_up((size_t)m * sizeof(TokResume), 256) + align_up((size_t)m * 2 * b->max_mb_w, 256) +
                                     (size_t)m * 32 * b->max_mb_w + 256)) return false;
  // ---- uploads
  cudaStream_t s = ctx->stream;
  for (const auto& r : ranges) {
    CU_TRY(cudaMemcpyAsync((uint8_t*)b->d_in.p + r.dev_off, r.base, r.size, cudaMemcpyHostToDevice, s), "H2D input");
  }
  CU_TRY(cudaMemcpyAsync(b->d_imgs.p, b->imgs.data(), sizeof(ImgDesc) * m, cudaMemcpyHostToDevice, s), "H2D descriptors");
  CU_TRY(cudaMemcpyAsync(b->d_ids.p, b->ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice, s), "H2D ids");
  CU_TRY(vp8k_configure(b->max_mb_w, b->max_mb_h), "cudaFuncSetAttribute");
  CU_TRY(cudaStreamSynchronize(s), "upload sync");
  return true;
}

static void batch_release(WebPBatch* b) {
  if (b == nullptr) return;
  if (b->ctx != nullptr) {
    DeviceCtx* c = b->ctx;
    own_free(c, b->d_in); own_free(c, b->d_imgs); own_free(c, b->d_hdrs); own_free(c, b->d_ids);
    own_free(c, b->d_mbinfo); own_free(c, b->d_coeffs); own_free(c, b->d_yuv); own_free(c, b->d_out); own_free(c, b->d_band); own_free(c, b->d_dither);
    own_free(c, b->d_aimgs); own_free(c, b->d_aplans); own_free(c, b->d_ahdrs); own_free(c, b->d_awork); own_free(c, b->d_awork2); own_free(c, b->d_alpha);
    for (auto e : b->ev) if (e) cudaEventDestroy(e);
  }
  delete b;
}

extern "C" WebPBatch* WebPBatchCreate(WebPBatchItem* items, int num_items, const WebPBatchOptions* options,
                                      VP8StatusCode* status) {
  VP8StatusCode dummy;
  if (status == NULL) status = &dummy;
  *status = VP8_STATUS_INVALID_PARAM;
  if (items == NULL || num_items <= 0) return NULL;
  WebPBatchOptions opt;
  if (options != NULL) opt = *options; else WebPBatchOptionsInit(&opt);
  g_last_error[0] = 0;
  WebPBatch* b = new WebPBatch();
  b->items = items; b->n = num_items; b->opt = opt;
  memset(&b->timings, 0, sizeof(b->timings));
  std::vector<Vp8Container> cont;
  if (batch_plan(b, cont) == 0) {   // nothing survived the host checks: no device work, no device needed
    *status = VP8_STATUS_OK;
    b->decoded = true;
    return b;
  }
  DeviceCtx* ctx = get_ctx(opt.device);
  if (ctx == nullptr) {
    for (int i = 0; i < num_items; ++i) {
      if (items[i].status != VP8_STATUS_OK) continue;
      items[i].status = VP8_STATUS_USER_ABORT;
      if (items[i].config != NULL && opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&items[i].config->output);
    }
    *status = VP8_STATUS_USER_ABORT;
    fprintf(stderr, "libwebp_b200: %s\n", g_last_error);
    delete b;
    return NULL;
  }
  b->ctx = ctx;
  ctx->mu.lock();
  const bool ok = batch_build(b, cont);
  ctx->mu.unlock();
  if (!ok) {
    const VP8StatusCode st = strstr(g_last_error, "cudaMalloc") ? VP8_STATUS_OUT_OF_MEMORY : VP8_STATUS_USER_ABORT;
    for (int i = 0; i < num_items; ++i) {
      if (items[i].status == VP8_STATUS_OK) {
        items[i].status = st;
        if (items[i].config != NULL && opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&items[i].config->output);
      }
    }
    *status = st;
    ctx->mu.lock(); batch_release(b); ctx->mu.unlock();
    return NULL;
  }
  *status = VP8_STATUS_OK;
  return b;
}

// Queues the device->host copies of images [first, first + count) on stream `s`. Consecutive images whose
// host buffers are contiguous and tightly packed travel in one cudaMemcpyAsync. With check_status the images
// that failed are skipped (their buffers are already released); without it the statuses are not known yet and
// every image still owns a valid destination.
static bool enqueue_download(WebPBatch* b, int first, int count, cudaStream_t s, bool check_status) {
  const uint8_t* dout = (const uint8_t*)b->d_out.p;
  uint8_t* run_host = nullptr; const uint8_t* run_dev = nullptr; size_t run_bytes = 0;
  auto flush = [&]() -> bool {
    if (run_bytes > 0) CU_TRY(cudaMemcpyAsync(run_host, run_dev, run_bytes, cudaMemcpyDeviceToHost, s), "D2H pixels");
    run_bytes = 0;
    return true;
  };
  for (int k = first; k < first + count; ++k) {
    const WebPBatchItem* it = &b->items[b->img_item[k]];
    if (check_status && it->status != VP8_STATUS_OK) continue;
    const ImgDesc& d = b->imgs[k];
    const WebPDecBuffer* o = &it->config->output;
    const uint8_t* src = dout + d.out_off;
    if (d.csp == MODE_YUV || d.csp == MODE_YUVA) {
      if (!flush()) return false;
      const int w = final_w(d), h = final_h(d), uvw = (w + 1) / 2, uvh = (h + 1) / 2;
      const WebPYUVABuffer* y = &o->u.YUVA;
      CU_TRY(cudaMemcpy2DAsync(y->y, y->y_stride, src, w, w, h, cudaMemcpyDeviceToHost, s), "D2H Y");
      CU_TRY(cudaMemcpy2DAsync(y->u, y->u_stride, src + (size_t)w * h, uvw, uvw, uvh, cudaMemcpyDeviceToHost, s), "D2H U");
      CU_TRY(cudaMemcpy2DAsync(y->v, y->v_stride, src + (size_t)w * h + (size_t)uvw * uvh, uvw, uvw, uvh, cudaMemcpyDeviceToHost, s), "D2H V");
      if (d.csp == MODE_YUVA) {
        CU_TRY(cudaMemcpy2DAsync(y->a, y->a_stride, src + (size_t)w * h + 2 * (size_t)uvw * uvh, w, w, h, cudaMemcpyDeviceToHost, s), "D2H A");
      }
      continue;
    }
    const size_t row = (size_t)d.out_stride;
    const size_t bytes = row * final_h(d);
    if ((size_t)o->u.RGBA.stride == row) {
      if (run_bytes > 0 && o->u.RGBA.rgba == run_host + run_bytes && src == run_dev + run_bytes) {
        run_bytes += bytes;
      } else {
        if (!flush()) return false;
        run_host = o->u.RGBA.rgba; run_dev = src; run_bytes = bytes;
      }
      if (run_bytes >= ((size_t)256 << 20)) { if (!flush()) return false; }
    } else {
      if (!flush()) return false;
      CU_TRY(cudaMemcpy2DAsync(o->u.RGBA.rgba, (size_t)o->u.RGBA.stride, src, row, row, final_h(d), cudaMemcpyDeviceToHost, s), "D2H pixels 2D");
    }
  }
  return flush();
}

// Output rows [row_lo(k), row_hi(k)) of images [first, first + count) (4-byte RGB family, not flipped): one 2-D copy
// when the images are equally sized and equally spaced on both sides, else one copy per image.
static bool enqueue_download_rows(WebPBatch* b, int first, int count, int pair_begin, int pair_end, cudaStream_t s) {
  const uint8_t* dout = (const uint8_t*)b->d_out.p;
  auto rows_of = [&](const ImgDesc& d, int* lo, int* hi) {
    const long l = pair_begin > 0 ? 2L * pair_begin - 1 : 0, h = 2L * pair_end - 1;
    *lo = (int)std::min<long>(l, d.out_h); *hi = (int)std::min<long>(h, d.out_h);
  };
  bool uniform = count > 1;
  const ImgDesc& d0 = b->imgs[first];
  const WebPDecBuffer* o0 = &b->items[b->img_item[first]].config->output;
  ptrdiff_t hpitch = 0, dpitch = 0;
  for (int k = first; k < first + count && uniform; ++k) {
    const ImgDesc& d = b->imgs[k];
    const WebPDecBuffer* o = &b->items[b->img_item[k]].config->output;
    if (d.out_h != d0.out_h || d.out_stride != d0.out_stride || o->u.RGBA.stride != d.out_stride) uniform = false;
    if (k == first + 1) { hpitch = o->u.RGBA.rgba - o0->u.RGBA.rgba; dpitch = (ptrdiff_t)(d.out_off - d0.out_off); }
    if (k > first && (o->u.RGBA.rgba - o0->u.RGBA.rgba != hpitch * (k - first) || (ptrdiff_t)(d.out_off - d0.out_off) != dpitch * (k - first)))
      uniform = false;
  }
  if (uniform && hpitch > 0 && dpitch > 0 && (size_t)hpitch < ((size_t)1 << 31) && (size_t)dpitch < ((size_t)1 << 31)) {
    int lo, hi; rows_of(d0, &lo, &hi);
    if (hi <= lo) return true;
    const size_t row = (size_t)d0.out_stride;
    CU_TRY(cudaMemcpy2DAsync(o0->u.RGBA.rgba + lo * row, (size_t)hpitch, dout + d0.out_off + lo * row, (size_t)dpitch, (hi - lo) * row,
                             (size_t)count, cudaMemcpyDeviceToHost, s), "D2H pixel band 2D");
    return true;
  }
  for (int k = first; k < first + count; ++k) {
    const ImgDesc& d = b->imgs[k];
    const WebPDecBuffer* o = &b->items[b->img_item[k]].config->output;
    int lo, hi; rows_of(d, &lo, &hi);
    if (hi <= lo) continue;
    const size_t row = (size_t)d.out_stride;
    CU_TRY(cudaMemcpy2DAsync(o->u.RGBA.rgba + (size_t)lo * o->u.RGBA.stride, (size_t)o->u.RGBA.stride, dout + d.out_off + lo * row, row, row,
                             (size_t)(hi - lo), cudaMemcpyDeviceToHost, s), "D2H pixel band");
  }
  return true;
}

enum { ST_MODES = 0, ST_TOKENS, ST_RECON, ST_FILTER, ST_EMIT, ST_ALPHA, ST_COUNT };

static int ev_mark(WebPBatch* b, cudaStream_t s) {   // records the next pooled event on `s`; -1 on failure
  if (b->ev_used == b->ev.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { set_error("cudaEventCreate", cudaGetLastError()); return -1; }
    b->ev.push_back(e);
  }
  if (cudaEventRecord(b->ev[b->ev_used], s) != cudaSuccess) { set_error("cudaEventRecord", cudaGetLastError()); return -1; }
  return (int)b->ev_used++;
}

// ALPH chunks of the batch: header pass, host-side sizing of the work areas (first decode only), then the pixel
// pass + unfilter, all queued on the compute stream ahead of the VP8 kernels. Leaves every image's w x h alpha
// plane in d_alpha; k_emit picks it up through ImgDesc::alpha_plane.
static bool batch_alpha(WebPBatch* b) {
  DeviceCtx* ctx = b->ctx;
  const int na = (int)b->aimgs.size();
  if (na == 0) return true;
  cudaStream_t s = ctx->stream;
  const uint8_t* arena = (const uint8_t*)b->d_in.p;
  if (!b->alpha_planned) {
    b->aplans.assign(na, AlphaPlan());
    b->ahdrs.resize(na);
    size_t work1 = 0;
    for (int a = 0; a < na; ++a) {
      const ImgDesc& d = b->imgs[b->aimgs[a]];
      b->aplans[a].scratch = work1; work1 += align_up(AL_SCRATCH_BYTES, 256);
      b->aplans[a].meta = work1; work1 += align_up(4 * (size_t)AL_META_PIXELS_BOUND(d.width, d.height) + 16, 256);
      b->aplans[a].tdata = work1; work1 += align_up(8 * (size_t)AL_META_PIXELS_BOUND(d.width, d.height) + 16, 256);
    }
    if (!own_alloc(ctx, b->d_aimgs, sizeof(int) * na) || !own_alloc(ctx, b->d_aplans, sizeof(AlphaPlan) * na) ||
        !own_alloc(ctx, b->d_ahdrs, sizeof(AlphaHdr) * na) || !own_alloc(ctx, b->d_awork, work1)) return false;
    for (int a = 0; a < na; ++a) {
      b->aplans[a].scratch += (uint64_t)(uintptr_t)b->d_awork.p;
      b->aplans[a].meta += (uint64_t)(uintptr_t)b->d_awork.p;
      b->aplans[a].tdata += (uint64_t)(uintptr_t)b->d_awork.p;
    }
    CU_TRY(cudaMemcpyAsync(b->d_aimgs.p, b->aimgs.data(), sizeof(int) * na, cudaMemcpyHostToDevice, s), "H2D alpha image list");
    CU_TRY(cudaMemcpyAsync(b->d_aplans.p, b->aplans.data(), sizeof(AlphaPlan) * na, cudaMemcpyHostToDevice, s), "H2D alpha plans");
  }
  vp8k_alpha_header(s, arena, (const ImgDesc*)b->d_imgs.p, (const int*)b->d_aimgs.p, (const AlphaPlan*)b->d_aplans.p,
                    (AlphaHdr*)b->d_ahdrs.p, na);
  if (!b->alpha_planned) {
    CU_TRY(cudaMemcpyAsync(b->ahdrs.data(), b->d_ahdrs.p, sizeof(AlphaHdr) * na, cudaMemcpyDeviceToHost, s), "D2H alpha headers");
    CU_TRY(cudaStreamSynchronize(s), "alpha header pass");
    size_t work2 = 0, planes = 0;
    std::vector<size_t> tab(na, 0), grp(na, 0), cod(na, 0), smo(na, 0);
    for (int a = 0; a < na; ++a) {
      const AlphaHdr& h = b->ahdrs[a];
      ImgDesc& d = b->imgs[b->aimgs[a]];
      if (h.status != AL_OK) continue;
      if (h.method == 1) {
        tab[a] = work2; work2 += align_up((size_t)h.num_groups * (size_t)h.group_entries * 4, 256);
        grp[a] = work2; work2 += align_up((size_t)h.num_groups * sizeof(AlGroup), 256);
        cod[a] = work2; work2 += align_up(4 * ((size_t)h.xsize * d.height + 4), 256);
      }
      if (d.flags & VP8B_FLAG_LOSSLESS) {   // its pixels go straight to the output arena (vp8k_lossless_finish)
        if (d.dst_w != 0) { smo[a] = work2 + 1; work2 += align_up(4 * (size_t)d.out_w * d.out_h + 16, 256); }
        continue;
      }
      if (d.alpha_dither != 0) { smo[a] = work2 + 1; work2 += align_up(2 * (size_t)d.out_w * d.out_h + 16, 256); }   // +1: 0 means none
      d.alpha_plane = planes;
      planes += align_up((size_t)d.width * d.height, 256);
    }
    Owned w2;
    if (!own_alloc(ctx, w2, work2 + 256) || !own_alloc(ctx, b->d_alpha, planes + 256)) { own_free(ctx, w2); return false; }
    b->d_awork2 = w2;   // lives as long as the batch
    for (int a = 0; a < na; ++a) {
      const uint64_t base = (uint64_t)(uintptr_t)b->d_awork2.p;
      b->aplans[a].tables = base + tab[a]; b->aplans[a].groups = base + grp[a]; b->aplans[a].coded = base + cod[a];
      b->aplans[a].smooth = smo[a] ? base + smo[a] - 1 : 0;
    }
    CU_TRY(cudaMemcpyAsync(b->d_aplans.p, b->aplans.data(), sizeof(AlphaPlan) * na, cudaMemcpyHostToDevice, s), "H2D alpha plans");
    CU_TRY(cudaMemcpyAsync(b->d_imgs.p, b->imgs.data(), sizeof(ImgDesc) * b->imgs.size(), cudaMemcpyHostToDevice, s), "H2D descriptors");
    b->alpha_planned = true;
  }
  vp8k_alpha_decode(s, arena, (const ImgDesc*)b->d_imgs.p, (const int*)b->d_aimgs.p, (const AlphaPlan*)b->d_aplans.p,
                    (AlphaHdr*)b->d_ahdrs.p, (uint8_t*)b->d_alpha.p, na);
  if (b->any_lossless) vp8k_lossless_finish(s, (const ImgDesc*)b->d_imgs.p, (const int*)b->d_aimgs.p, (const AlphaPlan*)b->d_aplans.p,
                                            (const AlphaHdr*)b->d_ahdrs.p, (uint8_t*)b->d_out.p, na);
  CU_TRY(cudaMemcpyAsync(b->ahdrs.data(), b->d_ahdrs.p, sizeof(AlphaHdr) * na, cudaMemcpyDeviceToHost, s), "D2H alpha status");
  return true;
}

// Runs the kernels on the device's compute stream. The two parse kernels take a whole wave at a time (the
// serial entropy decode needs every stream it can get in flight); the pixel stages then walk the wave in
// chunks, and with `download` (one-shot host-output path) each chunk's pixels start their way back on the copy
// stream as soon as its emit kernel has finished, while the next chunk is reconstructed / the next wave parsed.
static bool batch_decode(WebPBatch* b, bool download) {
  DeviceCtx* ctx = b->ctx;
  const int m = (int)b->imgs.size();
  memset(&b->timings, 0, sizeof(b->timings));
  if (m == 0) return true;
  CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
  cudaStream_t s = ctx->stream;
  const uint8_t* arena = (const uint8_t*)b->d_in.p;
  const ImgDesc* imgs = (const ImgDesc*)b->d_imgs.p;
  FrameHdr* hdrs = (FrameHdr*)b->d_hdrs.p;
  uint32_t* mbinfo = (uint32_t*)b->d_mbinfo.p;
  int16_t* coeffs = (int16_t*)b->d_coeffs.p;
  uint8_t* yuv = (uint8_t*)b->d_yuv.p;
  int launches = 0;
  b->ev_used = 0;
  b->spans.clear();
  // chunk of the pixel stages when downloads ride along: ~2 GiB of output (256 full-HD images) per chunk
  size_t chunk_bytes = (size_t)2 << 30;
  { const char* e = getenv("WEBP_B200_CHUNK_MB"); if (e != NULL && atoi(e) > 0) chunk_bytes = (size_t)atoi(e) << 20; }
#define MARK(var) const int var = ev_mark(b, s); if (var < 0) return false
  MARK(e_begin);
  if (!b->aimgs.empty()) {
    MARK(ea0);
    if (!batch_alpha(b)) return false;
    MARK(ea1);
    b->spans.push_back({ ST_ALPHA, ea0, ea1 });
    launches += 3;
  }
  for (const Wave& w : b->waves) {
    MARK(e0);
    vp8k_parse_modes(s, arena, imgs, hdrs, mbinfo, w.first, w.count, w.max_mb_w);
    ++launches;
    MARK(e1);
    CU_TRY(cudaMemsetAsync(coeffs, 0, w.mbs * 2 * VP8B_COEFFS_PER_MB, s), "memset coefficients");
    for (int lg = 0; lg < 4; ++lg) {
      if (w.ids_cnt[lg] == 0) continue;
      vp8k_parse_tokens(s, arena, imgs, hdrs, mbinfo, coeffs, (const int*)b->d_ids.p + w.ids_off[lg], w.ids_cnt[lg], 1 << lg, w.max_mb_w);
      ++launches;
    }
    // ---- row bands: see vp8_kernels.h. The wave qualifies when every image has one token partition, the lockstep
    // parser takes the launch, and every image goes through the row-pair output path unflipped and uncropped.
    int bands = 1;
    {
      static int env_bands = -1;
      if (env_bands < 0) { const char* e = getenv("WEBP_B200_BANDS"); env_bands = e ? atoi(e) : 1; }
      bool ok = env_bands > 1 && !b->any_dither && b->aimgs.empty() && w.ids_cnt[0] == w.count && vp8k_tokens_take_bands(w.count, 1) && w.max_mb_h >= 16;
      for (int k = w.first; ok && k < w.first + w.count; ++k) {
        const ImgDesc& d = b->imgs[k];
        const bool pairs = !(d.flags & VP8B_FLAG_NO_FANCY) && kBpp[d.csp] == 4 && d.csp != MODE_YUVA;
        if (!pairs || d.dst_w != 0 || (d.flags & VP8B_FLAG_FLIP) || d.crop_x != 0 || d.crop_y != 0 || d.out_w != d.width || d.out_h != d.height) ok = false;
      }
      // Off unless asked for (WEBP_B200_BANDS=n). Measured on 4096 full-HD images (profiles/r01m_row_bands.log): the
      // download of one band during the parse of the next slows the parse as much as it hides (a kernel running beside a
      // 55 GB/s device-to-host copy takes 2-3x as long on this box: tokens 312 -> 632 ms, end to end 1003 -> 1013 ms), and
      // running the pixel KERNELS beside the parse (WEBP_B200_BAND_OVERLAP=1) is worse still: the lockstep parser follows
      // one dependent chain per warp and loses more from sharing its issue port than the overlap gains (tokens 653 ms).
      if (ok) bands = std::min(env_bands, w.max_mb_h / 8);
    }
    if (bands > 1) {
      static int overlap = -1;
      if (overlap < 0) { const char* e = getenv("WEBP_B200_BAND_OVERLAP"); overlap = (e != NULL && atoi(e) > 0) ? 1 : 0; }
      cudaStream_t ps = overlap ? ctx->pixel_stream : s;
      TokResume* resume = (TokResume*)b->d_band.p;
      uint16_t* resume_ctx = (uint16_t*)((uint8_t*)b->d_band.p + align_up((size_t)m * sizeof(TokResume), 256));
      uint8_t* band_ctx = (uint8_t*)resume_ctx + align_up((size_t)m * 2 * b->max_mb_w, 256);
      b->spans.push_back({ ST_MODES, e0, e1 });
      int prev_tok = e1, last_px = -1;
      for (int k = 0; k < bands; ++k) {
        const int r0 = (int)((long)w.max_mb_h * k / bands), r1 = (k == bands - 1) ? 0x7fffffff : (int)((long)w.max_mb_h * (k + 1) / bands);
        // the loop filter of the next band still touches the last 3 luma / 6 chroma-covered rows of this one, and the
        // upsampler looks one chroma row ahead: hold back the last 8 + 2 pixel rows (kFilterExtraRows, frame_dec.c:201)
        const int p0 = (k == 0) ? 0 : 8 * r0 - 4, p1 = (k == bands - 1) ? 0x7fffffff : 8 * r1 - 4;
        vp8k_parse_tokens_band(s, arena, imgs, hdrs, mbinfo, coeffs, (const int*)b->d_ids.p + w.ids_off[0], w.count, w.max_mb_w, r0, r1,
                               resume, resume_ctx);
        ++launches;
        MARK(et);
        b->spans.push_back({ ST_TOKENS, prev_tok, et });
        if (ps != s) CU_TRY(cudaStreamWaitEvent(ps, b->ev[et], 0), "cudaStreamWaitEvent");
        const int q0 = ev_mark(b, ps); if (q0 < 0) return false;
        vp8k_reconstruct(ps, imgs, hdrs, mbinfo, coeffs, yuv, w.first, w.count, w.max_mb_w, w.max_mb_h, r0, r1, band_ctx);
        const int q1 = ev_mark(b, ps); if (q1 < 0) return false;
        vp8k_loop_filter(ps, imgs, hdrs, mbinfo, yuv, w.first, w.count, w.max_mb_h, r0, r1, nullptr);
        const int q2 = ev_mark(b, ps); if (q2 < 0) return false;
        const int band_pairs = (p1 == 0x7fffffff ? (16 * w.max_mb_h) / 2 + 1 : p1) - p0;
        vp8k_emit(ps, imgs, hdrs, yuv, (const uint8_t*)b->d_alpha.p, (uint8_t*)b->d_out.p, w.first, w.count,
                  ((16 * w.max_mb_w + 7) / 8) * band_pairs, p0, p1);
        const int q3 = ev_mark(b, ps); if (q3 < 0) return false;
        launches += 3;
        b->spans.push_back({ ST_RECON, q0, q1 });
        b->spans.push_back({ ST_FILTER, q1, q2 });
        b->spans.push_back({ ST_EMIT, q2, q3 });
        last_px = q3;
        prev_tok = (ps == s) ? q3 : et;
        if (download) {
          CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, b->ev[q3], 0), "cudaStreamWaitEvent");
          if (!enqueue_download_rows(b, w.first, w.count, p0, p1, ctx->copy_stream)) return false;
        }
      }
      if (ps != s) CU_TRY(cudaStreamWaitEvent(s, b->ev[last_px], 0), "cudaStreamWaitEvent");   // the scratch arrays are free again
      continue;
    }
    if (b->any_dither) { vp8k_dither_plan(s, imgs, hdrs, mbinfo, (int8_t*)b->d_dither.p, w.first, w.count); ++launches; }
    MARK(e2);
    b->spans.push_back({ ST_MODES, e0, e1 });
    b->spans.push_back({ ST_TOKENS, e1, e2 });
    int prev = e2;
    for (int c0 = w.first; c0 < w.first + w.count;) {
      int c1 = w.first + w.count;
      if (download) {
        size_t acc = 0;
        for (c1 = c0; c1 < w.first + w.count && (c1 == c0 || acc + b->plan[b->img_item[c1]].out_bytes <= chunk_bytes); ++c1)
          acc += b->plan[b->img_item[c1]].out_bytes;
      }
      const int cnt = c1 - c0;
      vp8k_reconstruct(s, imgs, hdrs, mbinfo, coeffs, yuv, c0, cnt, w.max_mb_w, w.max_mb_h, 0, 0x7fffffff, (uint8_t*)b->d_band.p);
      MARK(e3);
      vp8k_loop_filter(s, imgs, hdrs, mbinfo, yuv, c0, cnt, w.max_mb_h, 0, 0x7fffffff, (const int8_t*)b->d_dither.p);
      MARK(e4);
      vp8k_emit(s, imgs, hdrs, yuv, (const uint8_t*)b->d_alpha.p, (uint8_t*)b->d_out.p, c0, cnt, w.max_units, 0, 0x7fffffff);
      if (w.max_scaled_items > 0) { vp8k_emit_scaled(s, imgs, hdrs, yuv, (const uint8_t*)b->d_alpha.p, (uint8_t*)b->d_out.p, c0, cnt, w.max_scaled_items); ++launches; }
      MARK(e5);
      launches += 3;
      b->spans.push_back({ ST_RECON, prev, e3 });
      b->spans.push_back({ ST_FILTER, e3, e4 });
      b->spans.push_back({ ST_EMIT, e4, e5 });
      prev = e5;
      if (download) {
        CU_TRY(cudaStreamWaitEvent(ctx->copy_stream, b->ev[e5], 0), "cudaStreamWaitEvent");
        if (!enqueue_download(b, c0, cnt, ctx->copy_stream, false)) return false;
      }
      c0 = c1;
    }
  }
#undef MARK
  const int e_end = ev_mark(b, s); if (e_end < 0) return false;
  // per-image status words: FrameHdr::status is the first field
  CU_TRY(cudaMemcpy2DAsync(b->statuses.data(), sizeof(int), hdrs, sizeof(FrameHdr), sizeof(int), m, cudaMemcpyDeviceToHost, s),
         "D2H status");
  CU_TRY(cudaStreamSynchronize(s), "kernel execution");
  CU_TRY(cudaGetLastError(), "kernel launch");
  if (download) CU_TRY(cudaStreamSynchronize(ctx->copy_stream), "download sync");
  float acc[ST_COUNT] = { 0, 0, 0, 0, 0, 0 };
  for (const auto& sp : b->spans) {
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, b->ev[sp.a], b->ev[sp.b]), "cudaEventElapsedTime");
    acc[sp.stage] += ms;
  }
  b->timings.modes_ms = acc[ST_MODES]; b->timings.tokens_ms = acc[ST_TOKENS]; b->timings.recon_ms = acc[ST_RECON];
  b->timings.filter_ms = acc[ST_FILTER]; b->timings.emit_ms = acc[ST_EMIT];
  b->timings.alpha_ms = acc[ST_ALPHA];
  // with row bands the pixel stages of one band overlap the parse of the next: the step is what the stream saw end to end
  { float ms = 0; CU_TRY(cudaEventElapsedTime(&ms, b->ev[e_begin], b->ev[e_end]), "cudaEventElapsedTime"); b->timings.total_ms = ms; }
  b->timings.launches = launches;
  for (int k = 0; k < m; ++k) {
    WebPBatchItem* it = &b->items[b->img_item[k]];
    it->status = (VP8StatusCode)b->statuses[k];
    if (b->imgs[k].flags & VP8B_FLAG_LOSSLESS) {
      // every failure of a whole-picture VP8L decode is a bitstream error (vp8l_dec.c:1292,1479-1488: nothing suspends
      // outside the incremental decoder); the two limits of vp8l_alpha_core.h stay UNSUPPORTED_FEATURE
      const int ls = b->ahdrs[b->imgs[k].alpha_index].status;
      it->status = ls == AL_OK ? VP8_STATUS_OK : ls == AL_UNSUPPORTED ? VP8_STATUS_UNSUPPORTED_FEATURE : VP8_STATUS_BITSTREAM_ERROR;
    } else
    // a lost alpha plane loses the image (frame_dec.c:452-460), unless the VP8 stream had already failed
    if (it->status == VP8_STATUS_OK && b->imgs[k].alpha_size != 0) it->status = (VP8StatusCode)b->ahdrs[b->imgs[k].alpha_index].status;
    if (it->status != VP8_STATUS_OK && b->opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&it->config->output);
  }
  b->decoded = true;
  return true;
}

static VP8StatusCode batch_decode_locked(WebPBatch* b, bool download) {
  bool ok = true;
  if (b->ctx != nullptr) {
    b->ctx->mu.lock();
    ok = batch_decode(b, download);
    if (!ok) { cudaStreamSynchronize(b->ctx->stream); cudaStreamSynchronize(b->ctx->pixel_stream); cudaStreamSynchronize(b->ctx->copy_stream); cudaGetLastError(); }
    b->ctx->mu.unlock();
  }
  if (!ok) {
    fprintf(stderr, "libwebp_b200: %s\n", g_last_error);
    fail_all(b->items, b->n, VP8_STATUS_USER_ABORT);
    return VP8_STATUS_USER_ABORT;
  }
  for (int i = 0; i < b->n; ++i) if (b->items[i].status != VP8_STATUS_OK) return b->items[i].status;
  return VP8_STATUS_OK;
}

extern "C" VP8StatusCode WebPBatchDecode(WebPBatch* b) {
  if (b == NULL) return VP8_STATUS_INVALID_PARAM;
  return batch_decode_locked(b, false);
}

static bool batch_download(WebPBatch* b) {
  DeviceCtx* ctx = b->ctx;
  if (b->opt.output != WEBP_BATCH_HOST || b->imgs.empty()) return true;
  CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
  if (!enqueue_download(b, 0, (int)b->imgs.size(), ctx->copy_stream, true)) return false;
  CU_TRY(cudaStreamSynchronize(ctx->copy_stream), "download sync");
  return true;
}

extern "C" VP8StatusCode WebPBatchDownload(WebPBatch* b) {
  if (b == NULL || !b->decoded) return VP8_STATUS_INVALID_PARAM;
  if (b->ctx == nullptr) return VP8_STATUS_OK;
  b->ctx->mu.lock();
  const bool ok = batch_download(b);
  if (!ok) { cudaStreamSynchronize(b->ctx->copy_stream); cudaGetLastError(); }
  b->ctx->mu.unlock();
  if (!ok) { fprintf(stderr, "libwebp_b200: %s\n", g_last_error); return VP8_STATUS_USER_ABORT; }
  return VP8_STATUS_OK;
}

extern "C" void WebPBatchDestroy(WebPBatch* b) {
  if (b == NULL) return;
  DeviceCtx* c = b->ctx;
  if (c == nullptr) { delete b; return; }
  c->mu.lock();
  batch_release(b);
  c->mu.unlock();
}

extern "C" int WebPBatchOutput(const WebPBatch* b, int index, WebPBatchPlane* p) {
  if (b == NULL || p == NULL || index < 0 || index >= b->n || b->plan[index].img < 0) return 0;
  const ImgDesc& d = b->imgs[b->plan[index].img];
  uint8_t* base = (uint8_t*)b->d_out.p + d.out_off;
  memset(p, 0, sizeof(*p));
  const int fw = final_w(d), fh = final_h(d);
  p->width = fw; p->height = fh;
  p->y_or_rgba = base; p->stride = d.out_stride;
  if (d.csp == MODE_YUV || d.csp == MODE_YUVA) {
    const int uvw = (fw + 1) / 2, uvh = (fh + 1) / 2;
    p->u = base + (size_t)fw * fh;
    p->v = base + (size_t)fw * fh + (size_t)uvw * uvh;
    p->uv_stride = uvw;
  }
  return 1;
}

extern "C" int WebPBatchGetTimings(const WebPBatch* b, WebPBatchTimings* t) {
  if (b == NULL || t == NULL) return 0;
  *t = b->timings;
  return 1;
}

extern "C" VP8StatusCode WebPDecodeBatch(WebPBatchItem* items, int num_items, const WebPBatchOptions* options) {
  static int trace = -1;
  if (trace < 0) trace = getenv("WEBP_B200_TRACE") != NULL;
  const auto t0 = std::chrono::steady_clock::now();
  VP8StatusCode st;
  WebPBatch* b = WebPBatchCreate(items, num_items, options, &st);
  if (b == NULL) {
    if (st != VP8_STATUS_OK) return st;
    return VP8_STATUS_INVALID_PARAM;
  }
  const auto t1 = std::chrono::steady_clock::now();
  st = batch_decode_locked(b, b->opt.output == WEBP_BATCH_HOST);   // downloads ride along, wave by wave
  const auto t2 = std::chrono::steady_clock::now();
  const WebPBatchTimings tm = b->timings;
  const size_t nwaves = b->waves.size();
  WebPBatchDestroy(b);
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) {
      return std::chrono::duration<double, std::milli>(c - a).count();
    };
    fprintf(stderr, "libwebp_b200 trace: %d items, %zu waves: create(plan+alloc+H2D) %.1f ms, decode+download %.1f ms "
            "(kernels %.1f = modes %.1f tokens %.1f recon %.1f filter %.1f emit %.1f), destroy %.1f ms\n",
            num_items, nwaves, ms(t0, t1), ms(t1, t2), tm.total_ms, tm.modes_ms, tm.tokens_ms, tm.recon_ms, tm.filter_ms,
            tm.emit_ms, ms(t2, t3));
  }
  for (int i = 0; i < num_items; ++i) if (items[i].status != VP8_STATUS_OK) return items[i].status;
  return st;
}
