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
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
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
struct Owned { void* p = nullptr; size_t cap = 0; };

// One per device. `mu` guards the ORDER in which work is queued on the streams, the block caches and the wave scratch;
// it is held while a batch is planned into the streams, never while the host waits for the device. Several batches can
// therefore be in flight on one device: their kernels run back to back on `stream`, and the pixels of batch k travel
// to the host on `copy_stream` while the kernels of batch k+1 run (WebPBatchSubmit / WebPBatchWait, or simply two
// caller threads in WebPDecodeBatch).
struct DeviceCtx {
  int device = -1;
  bool ok = false;
  cudaStream_t stream = nullptr;        // kernels, uploads
  cudaStream_t copy_stream = nullptr;   // pixel downloads, overlapped with the kernels of the next chunk / wave / batch
  cudaStream_t pixel_stream = nullptr;  // row bands: reconstruction / filter / output of one band while the next is parsed
  std::mutex mu;
  std::vector<CachedBlock> cache;       // device blocks released by finished batches
  size_t cached_bytes = 0;
  size_t cache_limit = 0;               // released blocks are kept for the next batch up to this many bytes
  std::vector<CachedBlock> pinned;      // small page-locked host blocks (descriptors up, status words down)
  // Per-wave scratch (MbInfo, coefficient tokens, planes, ...) is shared by every batch on the device: only the kernels
  // touch it, and kernels of different batches are ordered (same stream, or through `scratch_free` when a caller
  // brought a stream of its own).
  Owned s_mbinfo, s_coeffs, s_tokens, s_mbtok, s_yuv, s_dither, s_band;
  const void* scratch_owner = nullptr;   // the batch whose kernels wrote the scratch last (WebPBatchDebugStages)
  cudaEvent_t scratch_free = nullptr;   // recorded behind the last kernel queued so far
  bool scratch_used = false;
  // Pixel downloads are queued by a thread of their own. cudaMemcpyAsync returns at once only while the driver's queue
  // has room: the 34 GB of a full batch are far more than it holds, and the caller that queued them inline sat in the
  // call until most of them had MOVED -- which is why, in round 1, kernels launched after a download looked slow.
  std::thread copier;
  std::mutex copy_mu;
  std::condition_variable copy_cv;
  std::deque<std::function<void()>> copy_jobs;
};

static void copier_main(DeviceCtx* c) {
  cudaSetDevice(c->device);
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(c->copy_mu);
      c->copy_cv.wait(lk, [&]() { return !c->copy_jobs.empty(); });
      job = std::move(c->copy_jobs.front());
      c->copy_jobs.pop_front();
    }
    job();
  }
}

static void copier_post(DeviceCtx* c, std::function<void()> job) {
  { std::lock_guard<std::mutex> lk(c->copy_mu); c->copy_jobs.push_back(std::move(job)); }
  c->copy_cv.notify_one();
}

static std::mutex g_ctx_mu;
static DeviceCtx* g_ctx[64];

// The calling thread's current device is put back when a library call returns.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != device) cudaSetDevice(device);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static DeviceCtx* get_ctx(int device) {
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) { set_error("cudaGetDevice (no CUDA device; this library has no CPU path)", cudaGetLastError()); return nullptr; }
  }
  if (device >= 64) { set_error("device ordinal out of range", cudaSuccess); return nullptr; }
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if (g_ctx[device] == nullptr) {
    DeviceCtx* c = new DeviceCtx();
    c->device = device;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->pixel_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->scratch_free, cudaEventDisableTiming);
    if (e == cudaSuccess) e = vp8k_init_device();   // per-device kernel attributes (dynamic shared memory opt-in)
    if (e != cudaSuccess) {
      set_error("CUDA device init (no usable GPU; this library has no CPU path)", e); cudaGetLastError();
      if (prev >= 0) cudaSetDevice(prev);
      delete c;
      return nullptr;
    }
    {
      // Released device blocks are kept for the next batch (cudaMalloc / cudaFree of tens of GB cost 0.1-1 s) up to a
      // quarter of the device's memory by default; WebPBatchSetCacheLimit() / WEBP_B200_CACHE_GB change it and
      // WebPBatchTrimCache() gives everything back.
      size_t free_b = 0, total_b = 0;
      const char* env = getenv("WEBP_B200_CACHE_GB");
      if (env != NULL) c->cache_limit = (size_t)(atof(env) * (double)(1u << 30));
      else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c->cache_limit = total_b / 4;
      cudaGetLastError();
    }
    if (prev >= 0) cudaSetDevice(prev);
    c->copier = std::thread(copier_main, c);
    c->copier.detach();   // lives as long as the process, like the context
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

// Small page-locked host blocks: what the library itself sends up (descriptors, launch lists) and brings down (status
// words) must not sit in pageable memory, or cudaMemcpyAsync turns into a host-side wait in the middle of the queue.
static void* pinned_alloc(DeviceCtx* c, size_t bytes, size_t* cap) {
  bytes = (bytes + 4095) & ~(size_t)4095;
  int best = -1;
  for (size_t i = 0; i < c->pinned.size(); ++i) {
    if (c->pinned[i].cap >= bytes && c->pinned[i].cap <= 4 * bytes && (best < 0 || c->pinned[i].cap < c->pinned[best].cap)) best = (int)i;
  }
  if (best >= 0) {
    void* p = c->pinned[best].p; *cap = c->pinned[best].cap;
    c->pinned.erase(c->pinned.begin() + best);
    return p;
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { set_error("cudaHostAlloc", cudaGetLastError()); return nullptr; }
  *cap = bytes;
  return p;
}
static void pinned_free(DeviceCtx* c, void* p, size_t cap) {
  if (p == nullptr) return;
  if (c->pinned.size() < 32 && cap <= ((size_t)64 << 20)) c->pinned.push_back({ p, cap });
  else cudaFreeHost(p);
}

static int resolve_device(int device) {
  if (device >= 0) return device;
  if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); return -1; }
  return device;
}

// Gives every cached block of `device` (-1: the calling thread's current device) back to the driver. Blocks of batches
// that are still alive are not touched. Returns the number of bytes released.
extern "C" size_t WebPBatchTrimCache(int device) {
  device = resolve_device(device);
  if (device < 0 || device >= 64) return 0;
  DeviceCtx* c;
  { std::lock_guard<std::mutex> lk(g_ctx_mu); c = g_ctx[device]; }
  if (c == nullptr) return 0;
  DeviceGuard g(c->device);
  std::lock_guard<std::mutex> lk(c->mu);
  size_t released = c->cached_bytes;
  cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->pixel_stream); cudaStreamSynchronize(c->copy_stream);
  for (auto& b : c->cache) cudaFree(b.p);
  c->cache.clear(); c->cached_bytes = 0;
  Owned* scratch[] = { &c->s_mbinfo, &c->s_coeffs, &c->s_tokens, &c->s_mbtok, &c->s_yuv, &c->s_dither, &c->s_band };
  for (Owned* o : scratch) { if (o->p != nullptr) { released += o->cap; cudaFree(o->p); o->p = nullptr; o->cap = 0; } }
  for (auto& b : c->pinned) cudaFreeHost(b.p);
  c->pinned.clear();
  cudaGetLastError();
  return released;
}

// Upper bound on the device memory the library keeps between batches on `device`; returns the previous bound.
extern "C" size_t WebPBatchSetCacheLimit(int device, size_t bytes) {
  DeviceCtx* c = get_ctx(device);
  if (c == nullptr) return 0;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t prev = c->cache_limit;
  c->cache_limit = bytes;
  return prev;
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

// For the incremental shim (webp_api.c): is this caller's buffer usable for a w x h picture? (hidden symbol)
extern "C" int vp8b_prepare_host_buffer(int w, int h, WebPDecBuffer* b) { return (int)prepare_host_buffer(w, h, b); }

// ---------------------------------------------------------------------------------------------------------
struct ItemPlan {
  int img = -1;            // index among the device images, -1 when the item failed on the host
  size_t frame_offset = 0; // VP8 frame tag inside the file
  size_t out_bytes = 0;    // bytes of this image in the device output arena (tight strides)
  bool discard = false;    // decode, but copy nothing out and end with VP8_STATUS_INVALID_PARAM unless the decode itself fails: a
                           // caller's buffer that is unusable, in the one case where the reference only finds out AFTER decoding
                           // (is_external_memory >= 2 + premultiplied output + a file with alpha: WebPDecode goes through a buffer
                           // of its own and copies at the end, webp_dec.c:769-786, buffer_dec.c:270-310)
};

struct Wave {
  int first = 0, count = 0;
  size_t mbs = 0;
  int max_mb_w = 0, max_mb_h = 0, max_units = 0;
  int max_scaled_items = 0;   // images with options.use_scaling: work items of vp8k_emit_scaled (0 = none in this wave)
  int literal = 0;            // images flagged VP8B_FLAG_LITERAL_READER: vp8k_parse_literal goes over the wave
  int ids_off[4] = { 0, 0, 0, 0 }, ids_cnt[4] = { 0, 0, 0, 0 };   // per log2(P) slice of the ids array
};

struct HostRange { const uint8_t* base; size_t size; size_t dev_off; };

struct WebPBatch {
  DeviceCtx* ctx = nullptr;
  WebPBatchItem* items = nullptr;
  int n = 0;
  WebPBatchOptions opt;
  cudaStream_t stream = nullptr;   // compute stream of this batch: the device's own, or the caller's (WebPBatchOptions::stream)
  std::vector<ItemPlan> plan;
  std::vector<ImgDesc> imgs;       // device images, wave-major
  std::vector<int> img_item;       // device image -> item
  std::vector<Wave> waves;
  std::vector<int> ids;            // token-parse launch lists
  int* statuses = nullptr;         // page-locked: status, fail_row, rows << 8 | filter_type of every FrameHdr (3 m ints, k_collect_status)
  void* h_stage = nullptr;         // page-locked staging of the descriptors and launch lists on their way up
  size_t h_status_cap = 0, h_stage_cap = 0;
  Owned d_in, d_imgs, d_hdrs, d_ids, d_out;
  size_t max_wave_mbs = 0;
  bool any_dither = false;     // options.dithering_strength: 128 offsets per macroblock of a wave
  bool any_lossless = false;   // whole-picture VP8L images ride the ALPH machinery (VP8B_FLAG_LOSSLESS)
  // images with an ALPH chunk
  std::vector<int> aimgs;              // their image indices
  std::vector<AlphaPlan> aplans;
  AlphaHdr* ahdrs = nullptr;           // page-locked host copy after the header pass / after the decode
  size_t h_ahdrs_cap = 0;
  Owned d_aimgs, d_aplans, d_ahdrs, d_awork, d_awork2, d_alpha;
  bool alpha_planned = false;          // work areas sized from the headers (kept across repeated decodes)
  size_t out_total = 0;
  int max_mb_w = 1, max_mb_h = 1;
  std::vector<cudaEvent_t> ev;     // pool of timing / hand-off events, grown on demand
  size_t ev_used = 0;
  struct Span { int stage, a, b; };
  std::vector<Span> spans;         // (stage, first event, second event) of every timed launch of the last decode
  int e_begin = -1, e_end = -1;    // first / last event of the last decode on the compute stream
  cudaEvent_t copy_done = nullptr; // behind the last pixel download of the last decode
  bool copies_queued = false;
  std::mutex job_mu;               // download jobs handed to the device's copier thread and not yet queued by it
  std::condition_variable job_cv;
  int jobs_pending = 0;
  bool job_failed = false;
  char job_error[256] = "";
  int launches = 0;
  WebPBatchTimings timings;
  bool queued = false;             // kernels are in the streams, results not yet collected (WebPBatchWait)
  bool decoded = false;
  // WebPBatchOptions::devices: the parent only shards. Child g owns items g, g + G, ... (copies of the caller's
  // WebPBatchItem; the configs they point to are the caller's) on devices[g].
  std::vector<WebPBatch*> shards;
  std::vector<WebPBatchItem> shard_items;
};

static void fail_all(WebPBatchItem* items, int n, VP8StatusCode st) {
  for (int i = 0; i < n; ++i) if (items[i].status == VP8_STATUS_OK) items[i].status = st;
}

// Host-side part of one item: container walk, option/colourspace screening, output buffer.
extern "C" int vp8b_host_headers_status(const uint8_t* frame, size_t frame_size, uint32_t part0_size, int width, int height, int is_lossless);

// The output request of one item against the picture's dimensions: what WebPAllocateDecBuffer would say (buffer_dec.c:41-227).
static VP8StatusCode check_request(WebPBatchItem* it, const WebPBatchOptions& opt, const Vp8Container* c, bool* discard);

static VP8StatusCode plan_item(WebPBatchItem* it, const WebPBatchOptions& opt, Vp8Container* c, bool* discard) {
  WebPDecoderConfig* cfg = it->config;
  if (cfg == NULL || it->data == NULL) return VP8_STATUS_INVALID_PARAM;   // webp_dec.c:756, GetFeatures webp_dec.c:693-695
  VP8StatusCode st = vp8b_get_features(it->data, it->data_size, &cfg->input);
  if (st != VP8_STATUS_OK) return st == VP8_STATUS_NOT_ENOUGH_DATA ? VP8_STATUS_BITSTREAM_ERROR : st;   // webp_dec.c:761-767
  st = (VP8StatusCode)vp8b_parse_container(it->data, it->data_size, 1, c);
  if (st != VP8_STATUS_OK) return st;
  if (c->has_animation) return VP8_STATUS_UNSUPPORTED_FEATURE;             // webp_dec.c:427-429
  if (!c->is_lossless && c->part0_size > c->frame_size - 10) return VP8_STATUS_NOT_ENOUGH_DATA;   // vp8_dec.c:345-348
  st = check_request(it, opt, c, discard);
  if (st != VP8_STATUS_OK) {
    // The reference parses the frame header / the VP8L header BEFORE it looks at the request (webp_dec.c:469-481): a file
    // whose header is damaged reports that, whatever was asked for. Decided on the host (vp8_host_probe.cpp): nothing of
    // this item reaches the device either way.
    const int hs = vp8b_host_headers_status(it->data + c->frame_offset, c->frame_size, c->part0_size, c->width, c->height, c->is_lossless);
    if (hs != 0) return (VP8StatusCode)hs;
  }
  return st;
}

static VP8StatusCode check_request(WebPBatchItem* it, const WebPBatchOptions& opt, const Vp8Container* c, bool* discard) {
  WebPDecoderConfig* cfg = it->config;
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
    ow = sw; oh = sh;
  }
  if (opt.output == WEBP_BATCH_HOST) {
    const VP8StatusCode bs = prepare_host_buffer(ow, oh, &cfg->output);
    if (bs == VP8_STATUS_INVALID_PARAM && cfg->output.is_external_memory >= 2 && WebPIsPremultipliedMode((WEBP_CSP_MODE)csp) && cfg->input.has_alpha) {
      *discard = true;   // see ItemPlan::discard
      return VP8_STATUS_OK;
    }
    return bs;
  }
  // device-resident output: the limits the reference's own allocator would have applied (buffer_dec.c:88-116, utils.h:34-41)
  if ((uint64_t)ow * 4 >= (1ull << 31)) return VP8_STATUS_INVALID_PARAM;
  if ((uint64_t)ow * 4 * (uint64_t)oh >= (1ull << 34)) return VP8_STATUS_OUT_OF_MEMORY;
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
    b->items[i].status = plan_item(&b->items[i], b->opt, &cont[i], &b->plan[i].discard);
    if (b->items[i].status == VP8_STATUS_OK) ++alive;
    else if (b->items[i].config != NULL && b->opt.output == WEBP_BATCH_HOST) {
      WebPFreeDecBuffer(&b->items[i].config->output);   // webp_dec.c:513-515
    }
  }
  return alive;
}

// The CUDA allocation (page-locked or registered host memory) that contains `p`, if CUDA knows one. One cudaMemcpyAsync
// must not reach outside the allocation its first byte lies in: the bytes between two caller buffers are only ours to
// read when both buffers sit in the same allocation (merged ranges used to span allocations).
struct HostAlloc { const uint8_t* base = nullptr; size_t size = 0; bool known = false; };
typedef int (*PointerGetAttributeFn)(void* data, int attribute, unsigned long long ptr);   // cuPointerGetAttribute
static PointerGetAttributeFn pointer_get_attribute() {
  static PointerGetAttributeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuPointerGetAttribute", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (PointerGetAttributeFn)f;
    cudaGetLastError();
  });
  return fn;
}
static HostAlloc host_alloc_of(const void* p) {
  HostAlloc a;
  PointerGetAttributeFn fn = pointer_get_attribute();
  if (fn == nullptr) return a;
  unsigned long long start = 0; size_t size = 0;
  // CU_POINTER_ATTRIBUTE_RANGE_START_ADDR = 11, CU_POINTER_ATTRIBUTE_RANGE_SIZE = 12
  if (fn(&start, 11, (unsigned long long)(uintptr_t)p) != 0 || fn(&size, 12, (unsigned long long)(uintptr_t)p) != 0 || size == 0) return a;
  a.base = (const uint8_t*)(uintptr_t)start; a.size = size; a.known = true;
  return a;
}
static inline uintptr_t page_of(const uint8_t* p) { return (uintptr_t)p >> 12; }

static bool batch_build(WebPBatch* b, const std::vector<Vp8Container>& cont) {
  DeviceCtx* ctx = b->ctx;
  const int n = b->n;
  // ---- input ranges: host buffers that lie in the same allocation, close to each other, travel in one H2D copy
  std::vector<int> order;
  for (int i = 0; i < n; ++i) if (b->items[i].status == VP8_STATUS_OK) order.push_back(i);
  std::vector<int> by_addr(order);
  std::sort(by_addr.begin(), by_addr.end(), [&](int a, int c2) { return b->items[a].data < b->items[c2].data; });
  std::vector<HostRange> ranges;
  std::vector<int> item_range(n, -1);
  size_t in_total = 256;
  HostAlloc cur_alloc;
  for (int i : by_addr) {
    const uint8_t* p = b->items[i].data;
    const size_t sz = b->items[i].data_size;
    if (!ranges.empty()) {
      HostRange& r = ranges.back();
      const uint8_t* rend = r.base + r.size;
      bool merge = false;
      if (p >= r.base && p + sz <= rend) merge = true;                          // inside what is copied anyway
      else if (cur_alloc.known) merge = p >= r.base && p <= rend + 65536 && p + sz <= cur_alloc.base + cur_alloc.size;
      else if (p >= r.base && p <= rend + 4095 && (p <= rend || page_of(p) - page_of(rend - 1) <= 1)) {
        // memory CUDA does not know (pageable): the gap is readable when each of its bytes shares a page with a byte of
        // one of the two buffers; the neighbour must be pageable as well
        merge = !host_alloc_of(p).known;
      }
      if (merge) {
        const size_t end = (size_t)(p - r.base) + sz;
        if (end > r.size) r.size = end;
        item_range[i] = (int)ranges.size() - 1;
        continue;
      }
    }
    ranges.push_back({ p, sz, 0 });
    cur_alloc = host_alloc_of(p);
    if (cur_alloc.known && p + sz > cur_alloc.base + cur_alloc.size) cur_alloc.known = false;   // never expected
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
      d.dst_w = (uint32_t)fw; d.dst_h = (uint32_t)fh;
      // WebPIoInitFromOptions, webp_dec.c:851-856: no loop filter for large downscaling ratios (against the whole picture)
      if (fw < c.width * 3 / 4 && fh < c.height * 3 / 4) d.flags |= VP8B_FLAG_BYPASS_FILTER;
    }
    const int ds = cfg->options.dithering_strength;
    d.dither_f = (uint8_t)(ds < 0 ? 0 : ds > 100 ? 255 : ds * 255 / 100);
    if (d.dither_f != 0) b->any_dither = true;
    d.num_parts = c.is_lossless ? 0 : (uint8_t)vp8b_prescan_partitions(b->items[i].data + c.frame_offset + 10, c.part0_size);
    // a partition that starts with 0xFF (never encoded): the reference's reader leaves its range, see vp8_literal.h
    if (!c.is_lossless && c.frame_size >= 10 &&
        vp8b_partition_starts_with_ff(b->items[i].data + c.frame_offset + 10, c.part0_size, c.frame_size - 10, d.num_parts)) d.flags |= VP8B_FLAG_LITERAL_READER;
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
  if (m == 0) return true;

  // ---- page-locked staging for what the library itself moves: descriptors + launch lists up, status words down
  b->statuses = (int*)pinned_alloc(ctx, 3 * sizeof(int) * (size_t)m, &b->h_status_cap);
  b->h_stage = pinned_alloc(ctx, (sizeof(ImgDesc) + sizeof(int)) * (size_t)m, &b->h_stage_cap);
  if (b->statuses == nullptr || b->h_stage == nullptr) return false;
  memset(b->statuses, 0, 3 * sizeof(int) * (size_t)m);
  // ---- resident allocations: input (64 KB tail padding, see vp8_tokens_fsm.h:tk_lane_init), descriptors, headers, output
  if (!own_alloc(ctx, b->d_in, in_total + 65536) || !own_alloc(ctx, b->d_imgs, sizeof(ImgDesc) * m) ||
      !own_alloc(ctx, b->d_hdrs, sizeof(FrameHdr) * m) || !own_alloc(ctx, b->d_ids, sizeof(int) * m) ||
      !own_alloc(ctx, b->d_out, b->out_total + 256)) return false;
  // ---- waves: per-macroblock scratch = 16 (MbInfo) + 384 (planes) bytes + the levels: 8 (MbTok) + 4 * 384 reserved for
  // the token stream (only what a macroblock really has is ever touched), or the 800-byte dense plane of the older parsers
  const size_t per_mb = 16 + 384 + (vp8k_tokens_use_stream() ? 8 + 4 * VP8B_TOKENS_PER_MB : 2 * VP8B_COEFFS_PER_MB);
  size_t budget = b->opt.scratch_bytes;
  if (budget == 0) {
    size_t free_b = 0, total_b = 0;
    CU_TRY(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
    const size_t scratch_now = ctx->s_mbinfo.cap + ctx->s_coeffs.cap + ctx->s_tokens.cap + ctx->s_mbtok.cap + ctx->s_yuv.cap +
                               ctx->s_dither.cap + ctx->s_band.cap;
    budget = (size_t)((double)(free_b + ctx->cached_bytes + scratch_now) * 0.85);
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
      if (b->imgs[k].flags & VP8B_FLAG_LITERAL_READER) ++w.literal;
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
        const int dw = (int)d.dst_w, uvdw = (dw + 1) / 2;
        w.max_scaled_items = std::max(w.max_scaled_items, (d.csp == MODE_YUV || d.csp == MODE_YUVA) ? dw + 2 * uvdw + (d.csp == MODE_YUVA ? dw : 0) : dw);
      }
    }
    b->waves.push_back(w);
  }
  b->max_wave_mbs = 0;
  b->ids.resize(m);
  for (auto& w : b->waves) {
    b->max_wave_mbs = std::max(b->max_wave_mbs, w.mbs);
    int pos = w.first;
    for (int lg = 0; lg < 4; ++lg) {
      w.ids_off[lg] = pos;
      for (int k = w.first; k < w.first + w.count; ++k) if (b->imgs[k].num_parts == (1 << lg)) b->ids[pos++] = k;
      w.ids_cnt[lg] = pos - w.ids_off[lg];
    }
  }
  // ---- uploads: queued, not waited for (the caller's input buffers stay untouched until the batch is collected)
  cudaStream_t s = b->stream;
  for (const auto& r : ranges) {
    CU_TRY(cudaMemcpyAsync((uint8_t*)b->d_in.p + r.dev_off, r.base, r.size, cudaMemcpyHostToDevice, s), "H2D input");
  }
  memcpy(b->h_stage, b->imgs.data(), sizeof(ImgDesc) * (size_t)m);
  memcpy((uint8_t*)b->h_stage + sizeof(ImgDesc) * (size_t)m, b->ids.data(), sizeof(int) * (size_t)m);
  CU_TRY(cudaMemcpyAsync(b->d_imgs.p, b->h_stage, sizeof(ImgDesc) * (size_t)m, cudaMemcpyHostToDevice, s), "H2D descriptors");
  CU_TRY(cudaMemcpyAsync(b->d_ids.p, (uint8_t*)b->h_stage + sizeof(ImgDesc) * (size_t)m, sizeof(int) * (size_t)m, cudaMemcpyHostToDevice, s), "H2D ids");
  return true;
}

// The device's wave scratch, grown to what this batch needs. Called with ctx->mu held, before the batch's kernels are
// queued. A block that is replaced goes back to the cache (or to the driver, which waits for the device first), and
// whoever gets it next uses it behind the kernels already queued: same stream, or behind `scratch_free`.
static bool ensure_scratch(WebPBatch* b) {
  DeviceCtx* c = b->ctx;
  const int m = (int)b->imgs.size();
  auto grow = [&](Owned& o, size_t bytes) -> bool {
    if (o.p != nullptr && o.cap >= bytes) return true;
    own_free(c, o);
    return own_alloc(c, o, bytes);
  };
  if (!grow(c->s_mbinfo, b->max_wave_mbs * 16 + 256) || !grow(c->s_yuv, b->max_wave_mbs * 384 + 256)) return false;
  if (vp8k_tokens_use_stream()) {
    if (!grow(c->s_tokens, b->max_wave_mbs * 4 * VP8B_TOKENS_PER_MB + 256) || !grow(c->s_mbtok, b->max_wave_mbs * 8 + 256)) return false;
  } else if (!grow(c->s_coeffs, b->max_wave_mbs * 2 * VP8B_COEFFS_PER_MB + 256)) return false;
  if (b->any_dither && !grow(c->s_dither, b->max_wave_mbs * 128 + 256)) return false;
  return grow(c->s_band, align_up((size_t)m * sizeof(TokResume), 256) + align_up((size_t)m * 2 * b->max_mb_w, 256) +
                             (size_t)m * 32 * b->max_mb_w + 256);
}

static void batch_release(WebPBatch* b) {
  if (b == nullptr) return;
  if (b->ctx != nullptr) {
    DeviceCtx* c = b->ctx;
    own_free(c, b->d_in); own_free(c, b->d_imgs); own_free(c, b->d_hdrs); own_free(c, b->d_ids); own_free(c, b->d_out);
    pinned_free(c, b->statuses, b->h_status_cap); pinned_free(c, b->h_stage, b->h_stage_cap); pinned_free(c, b->ahdrs, b->h_ahdrs_cap);
    if (b->copy_done != nullptr) cudaEventDestroy(b->copy_done);
    own_free(c, b->d_aimgs); own_free(c, b->d_aplans); own_free(c, b->d_ahdrs); own_free(c, b->d_awork); own_free(c, b->d_awork2); own_free(c, b->d_alpha);
    for (auto e : b->ev) if (e) cudaEventDestroy(e);
  }
  delete b;
}

// One device: host checks, device memory, uploads queued. *status = OK with a batch, or why there is none.
static WebPBatch* create_single(WebPBatchItem* items, int num_items, const WebPBatchOptions& opt, VP8StatusCode* status) {
  *status = VP8_STATUS_INVALID_PARAM;
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
  b->stream = opt.stream != NULL ? (cudaStream_t)opt.stream : ctx->stream;
  DeviceGuard guard(ctx->device);
  ctx->mu.lock();
  const bool ok = batch_build(b, cont);
  ctx->mu.unlock();
  if (!ok) {
    const VP8StatusCode st = (strstr(g_last_error, "cudaMalloc") || strstr(g_last_error, "cudaHostAlloc")) ? VP8_STATUS_OUT_OF_MEMORY : VP8_STATUS_USER_ABORT;
    for (int i = 0; i < num_items; ++i) {
      if (items[i].status == VP8_STATUS_OK) {
        items[i].status = st;
        if (items[i].config != NULL && opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&items[i].config->output);
      }
    }
    *status = st;
    cudaStreamSynchronize(b->stream); cudaGetLastError();   // uploads already queued read the caller's memory
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
  HostAlloc run_alloc;   // a run never leaves the allocation it started in (see batch_build)
  auto flush = [&]() -> bool {
    if (run_bytes > 0) CU_TRY(cudaMemcpyAsync(run_host, run_dev, run_bytes, cudaMemcpyDeviceToHost, s), "D2H pixels");
    run_bytes = 0;
    return true;
  };
  for (int k = first; k < first + count; ++k) {
    const WebPBatchItem* it = &b->items[b->img_item[k]];
    if (check_status && it->status != VP8_STATUS_OK) continue;
    if (b->plan[b->img_item[k]].discard) continue;
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
      if (run_bytes > 0 && o->u.RGBA.rgba == run_host + run_bytes && src == run_dev + run_bytes &&
          (run_alloc.known ? o->u.RGBA.rgba + bytes <= run_alloc.base + run_alloc.size : !host_alloc_of(o->u.RGBA.rgba).known)) {
        run_bytes += bytes;
      } else {
        if (!flush()) return false;
        run_host = o->u.RGBA.rgba; run_dev = src; run_bytes = bytes;
        run_alloc = host_alloc_of(run_host);
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
    if (d.out_h != d0.out_h || d.out_stride != d0.out_stride || o->u.RGBA.stride != d.out_stride || b->plan[b->img_item[k]].discard) uniform = false;
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
    if (hi <= lo || b->plan[b->img_item[k]].discard) continue;
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
static bool batch_alpha(WebPBatch* b, cudaStream_t s) {
  DeviceCtx* ctx = b->ctx;
  const int na = (int)b->aimgs.size();
  if (na == 0) return true;
  const uint8_t* arena = (const uint8_t*)b->d_in.p;
  if (!b->alpha_planned) {
    b->aplans.assign(na, AlphaPlan());
    if (b->ahdrs == nullptr) b->ahdrs = (AlphaHdr*)pinned_alloc(ctx, sizeof(AlphaHdr) * (size_t)na, &b->h_ahdrs_cap);
    if (b->ahdrs == nullptr) return false;
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
    vp8k_copy_to_host(s, b->d_ahdrs.p, b->ahdrs, sizeof(AlphaHdr) * na);   // page-locked + mapped: no copy engine (see k_collect_status)
    CU_TRY(cudaStreamSynchronize(s), "alpha header pass");
    size_t work2 = 0, planes = 0;
    std::vector<size_t> tab(na, 0), grp(na, 0), cod(na, 0), smo(na, 0);
    for (int a = 0; a < na; ++a) {
      const AlphaHdr& h = b->ahdrs[a];
      ImgDesc& d = b->imgs[b->aimgs[a]];
      if (h.status != AL_OK) continue;
      if (h.method == 1) {
        tab[a] = work2; work2 += align_up((size_t)h.used_groups * (size_t)h.group_entries * 4, 256);
        grp[a] = work2; work2 += align_up((size_t)h.used_groups * sizeof(AlGroup), 256);
        cod[a] = work2; work2 += align_up(4 * ((size_t)std::max(h.xsize, h.px_stride) * d.height + 4), 256);
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
  vp8k_copy_to_host(s, b->d_ahdrs.p, b->ahdrs, sizeof(AlphaHdr) * na);
  return true;
}

// Runs the kernels on the device's compute stream. The two parse kernels take a whole wave at a time (the
// serial entropy decode needs every stream it can get in flight); the pixel stages then walk the wave in
// chunks, and with `download` (one-shot host-output path) each chunk's pixels start their way back on the copy
// stream as soon as its emit kernel has finished, while the next chunk is reconstructed / the next wave parsed.
// Hands one download step to the device's copier thread: wait (on the copy stream) for `after`, then queue the copies.
// Jobs of all batches run in the order they were posted, so the copy stream sees the same order as before.
static void post_download(WebPBatch* b, cudaEvent_t after, std::function<bool()> copies) {
  DeviceCtx* ctx = b->ctx;
  { std::lock_guard<std::mutex> lk(b->job_mu); ++b->jobs_pending; }
  copier_post(ctx, [=]() {
    bool ok = true;
    g_last_error[0] = 0;
    if (after != nullptr && cudaStreamWaitEvent(ctx->copy_stream, after, 0) != cudaSuccess) { set_error("cudaStreamWaitEvent", cudaGetLastError()); ok = false; }
    if (ok) ok = copies();
    std::lock_guard<std::mutex> lk(b->job_mu);
    if (!ok && !b->job_failed) { b->job_failed = true; snprintf(b->job_error, sizeof(b->job_error), "%s", g_last_error); }
    --b->jobs_pending;
    b->job_cv.notify_all();
  });
}

static bool jobs_drain(WebPBatch* b) {   // every download job of the batch has been queued on the copy stream
  std::unique_lock<std::mutex> lk(b->job_mu);
  b->job_cv.wait(lk, [&]() { return b->jobs_pending == 0; });
  if (b->job_failed) { snprintf(g_last_error, sizeof(g_last_error), "%s", b->job_error); b->job_failed = false; return false; }
  return true;
}

static bool batch_enqueue(WebPBatch* b, bool download) {
  DeviceCtx* ctx = b->ctx;
  const int m = (int)b->imgs.size();
  memset(&b->timings, 0, sizeof(b->timings));
  b->e_begin = b->e_end = -1;
  b->copies_queued = false;
  b->launches = 0;
  if (m == 0) return true;
  if (!ensure_scratch(b)) return false;
  cudaStream_t s = b->stream;
  if (ctx->scratch_used) CU_TRY(cudaStreamWaitEvent(s, ctx->scratch_free, 0), "cudaStreamWaitEvent");   // no-op on the device's own stream
  const uint8_t* arena = (const uint8_t*)b->d_in.p;
  const ImgDesc* imgs = (const ImgDesc*)b->d_imgs.p;
  FrameHdr* hdrs = (FrameHdr*)b->d_hdrs.p;
  uint32_t* mbinfo = (uint32_t*)ctx->s_mbinfo.p;
  const bool stream_tokens = vp8k_tokens_use_stream() != 0;
  int16_t* coeffs = (int16_t*)ctx->s_coeffs.p;
  uint32_t* tokens = stream_tokens ? (uint32_t*)ctx->s_tokens.p : nullptr;
  void* mbtok = stream_tokens ? ctx->s_mbtok.p : nullptr;
  uint8_t* yuv = (uint8_t*)ctx->s_yuv.p;
  int8_t* dither = b->any_dither ? (int8_t*)ctx->s_dither.p : nullptr;
  uint8_t* band = (uint8_t*)ctx->s_band.p;
  int launches = 0;
  b->ev_used = 0;
  b->spans.clear();
  // chunk of the pixel stages when downloads ride along: ~2 GiB of output (256 full-HD images) per chunk
  size_t chunk_bytes = (size_t)2 << 30;
  { const char* e = getenv("WEBP_B200_CHUNK_MB"); if (e != NULL && atoi(e) > 0) chunk_bytes = (size_t)atoi(e) << 20; }
#define MARK(var) const int var = ev_mark(b, s); if (var < 0) return false
  MARK(e_begin);
  b->e_begin = e_begin;
  // ALPH chunks and whole VP8L pictures first, on the compute stream. (Decoding them on a stream of their own beside the
  // VP8 parse was measured and dropped, profiles/r02h: the VP8L loops are a few hundred single-thread warps, and a second
  // warp on a sub-partition more than doubles the time of the latency-bound parser -- 4096x4096 + ALPH: tokens 1324 ->
  // 2932 ms, modes 93 -> 184 ms, for 309 ms of alpha work hidden.)
  int alpha_done = -1;
  if (!b->aimgs.empty()) {
    MARK(ea0);
    if (!batch_alpha(b, s)) return false;
    MARK(ea1);
    b->spans.push_back({ ST_ALPHA, ea0, ea1 });
    launches += 3;
  }
  for (const Wave& w : b->waves) {
    MARK(e0);
    vp8k_parse_modes(s, arena, imgs, hdrs, mbinfo, w.first, w.count, w.max_mb_w);
    ++launches;
    MARK(e1);
    if (!stream_tokens) CU_TRY(cudaMemsetAsync(coeffs, 0, w.mbs * 2 * VP8B_COEFFS_PER_MB, s), "memset coefficients");
    for (int lg = 0; lg < 4; ++lg) {
      if (w.ids_cnt[lg] == 0) continue;
      if (stream_tokens) vp8k_parse_tokens_stream(s, arena, imgs, hdrs, mbinfo, tokens, mbtok, (const int*)b->d_ids.p + w.ids_off[lg], w.ids_cnt[lg], 1 << lg, w.max_mb_w);
      else vp8k_parse_tokens(s, arena, imgs, hdrs, mbinfo, coeffs, (const int*)b->d_ids.p + w.ids_off[lg], w.ids_cnt[lg], 1 << lg, w.max_mb_w);
      ++launches;
    }
    if (stream_tokens && w.literal > 0) { vp8k_parse_literal(s, arena, imgs, hdrs, mbinfo, tokens, mbtok, w.first, w.count, w.max_mb_w); ++launches; }
    // ---- row bands: see vp8_kernels.h. The wave qualifies when every image has one token partition, the lockstep
    // parser takes the launch, and every image goes through the row-pair output path unflipped and uncropped.
    int bands = 1;
    {
      static int env_bands = -1;
      if (env_bands < 0) { const char* e = getenv("WEBP_B200_BANDS"); env_bands = e ? atoi(e) : 1; }
      bool ok = env_bands > 1 && !stream_tokens && !b->any_dither && b->aimgs.empty() && w.ids_cnt[0] == w.count && vp8k_tokens_take_bands(w.count, 1) && w.max_mb_h >= 16;
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
      if (alpha_done >= 0) { CU_TRY(cudaStreamWaitEvent(s, b->ev[alpha_done], 0), "cudaStreamWaitEvent"); alpha_done = -1; }
      static int overlap = -1;
      if (overlap < 0) { const char* e = getenv("WEBP_B200_BAND_OVERLAP"); overlap = (e != NULL && atoi(e) > 0) ? 1 : 0; }
      cudaStream_t ps = overlap ? ctx->pixel_stream : s;
      TokResume* resume = (TokResume*)band;
      uint16_t* resume_ctx = (uint16_t*)(band + align_up((size_t)m * sizeof(TokResume), 256));
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
        vp8k_reconstruct(ps, imgs, hdrs, mbinfo, coeffs, yuv, w.first, w.count, w.max_mb_w, w.max_mb_h, r0, r1, band_ctx, nullptr, nullptr);
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
          const int first = w.first, count = w.count;
          post_download(b, b->ev[q3], [=]() { return enqueue_download_rows(b, first, count, p0, p1, ctx->copy_stream); });
        }
      }
      if (ps != s) CU_TRY(cudaStreamWaitEvent(s, b->ev[last_px], 0), "cudaStreamWaitEvent");   // the scratch arrays are free again
      continue;
    }
    if (b->any_dither) { vp8k_dither_plan(s, imgs, hdrs, mbinfo, dither, w.first, w.count); ++launches; }
    if (alpha_done >= 0) { CU_TRY(cudaStreamWaitEvent(s, b->ev[alpha_done], 0), "cudaStreamWaitEvent"); alpha_done = -1; }
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
      vp8k_reconstruct(s, imgs, hdrs, mbinfo, coeffs, yuv, c0, cnt, w.max_mb_w, w.max_mb_h, 0, 0x7fffffff, band, tokens, mbtok);
      MARK(e3);
      vp8k_loop_filter(s, imgs, hdrs, mbinfo, yuv, c0, cnt, w.max_mb_h, 0, 0x7fffffff, dither);
      MARK(e4);
      vp8k_emit(s, imgs, hdrs, yuv, (const uint8_t*)b->d_alpha.p, (uint8_t*)b->d_out.p, c0, cnt, w.max_units, 0, 0x7fffffff);
      if (w.max_scaled_items > 0) { vp8k_emit_scaled(s, imgs, hdrs, yuv, (const uint8_t*)b->d_alpha.p, (uint8_t*)b->d_out.p, c0, cnt, w.max_scaled_items); ++launches; }
      MARK(e5);
      launches += 3;
      b->spans.push_back({ ST_RECON, prev, e3 });
      b->spans.push_back({ ST_FILTER, e3, e4 });
      b->spans.push_back({ ST_EMIT, e4, e5 });
      prev = e5;
      if (download) post_download(b, b->ev[e5], [=]() { return enqueue_download(b, c0, cnt, ctx->copy_stream, false); });
      c0 = c1;
    }
  }
#undef MARK
  if (alpha_done >= 0) CU_TRY(cudaStreamWaitEvent(s, b->ev[alpha_done], 0), "cudaStreamWaitEvent");
  const int e_end = ev_mark(b, s); if (e_end < 0) return false;
  b->e_end = e_end;
  b->launches = launches;
  CU_TRY(cudaEventRecord(ctx->scratch_free, s), "cudaEventRecord");
  ctx->scratch_used = true;
  ctx->scratch_owner = b;
  // per-image status words, written into page-locked host memory by the device (see k_collect_status)
  {
    int* dev_view = nullptr;
    CU_TRY(cudaHostGetDevicePointer((void**)&dev_view, b->statuses, 0), "cudaHostGetDevicePointer");
    vp8k_collect_status(s, hdrs, dev_view, m);
  }
  const int e_status = ev_mark(b, s); if (e_status < 0) return false;
  (void)e_status;   // the last event of the pool on the compute stream: batch_finish waits for it
  if (download) {
    if (b->copy_done == nullptr) CU_TRY(cudaEventCreateWithFlags(&b->copy_done, cudaEventDisableTiming), "cudaEventCreate");
    cudaEvent_t done = b->copy_done;
    post_download(b, nullptr, [=]() {
      if (cudaEventRecord(done, ctx->copy_stream) != cudaSuccess) { set_error("cudaEventRecord", cudaGetLastError()); return false; }
      return true;
    });
    b->copies_queued = true;
  }
  b->queued = true;
  return true;
}

// Waits for what batch_enqueue put into the streams (no lock held: other batches may be queued meanwhile), then turns
// the status words into per-item VP8StatusCodes and the events into stage times.
static bool batch_finish(WebPBatch* b) {
  if (!b->queued) return true;
  b->queued = false;
  const int m = (int)b->imgs.size();
  CU_TRY(cudaEventSynchronize(b->ev[b->ev_used - 1]), "kernel execution");
  if (b->copies_queued) {
    if (!jobs_drain(b)) return false;
    CU_TRY(cudaEventSynchronize(b->copy_done), "download sync");
  }
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
  { float ms = 0; CU_TRY(cudaEventElapsedTime(&ms, b->ev[b->e_begin], b->ev[b->e_end]), "cudaEventElapsedTime"); b->timings.total_ms = ms; }
  b->timings.launches = b->launches;
  for (int k = 0; k < m; ++k) {
    WebPBatchItem* it = &b->items[b->img_item[k]];
    it->status = (VP8StatusCode)b->statuses[3 * k];
    if (b->imgs[k].flags & VP8B_FLAG_LOSSLESS) {
      // every failure of a whole-picture VP8L decode is a bitstream error (vp8l_dec.c:1292,1479-1488: nothing suspends
      // outside the incremental decoder); the two limits of vp8l_alpha_core.h stay UNSUPPORTED_FEATURE
      const int ls = b->ahdrs[b->imgs[k].alpha_index].status;
      it->status = ls == AL_OK ? VP8_STATUS_OK : ls == AL_UNSUPPORTED ? VP8_STATUS_UNSUPPORTED_FEATURE : VP8_STATUS_BITSTREAM_ERROR;
    } else
    // a lost alpha plane loses the image (frame_dec.c:452-460), unless the VP8 stream fails first: the reference decodes alpha
    // rows as the macroblock rows above them finish, so with both chunks damaged the rows of the two failures decide
    if (b->imgs[k].alpha_size != 0) {
      const ImgDesc& d = b->imgs[k];
      const AlphaHdr& ah = b->ahdrs[d.alpha_index];
      if (it->status == VP8_STATUS_OK) it->status = (VP8StatusCode)ah.status;
      else if (ah.status != AL_OK) {
        int vrow = b->statuses[3 * k + 1];
        if (vrow == VP8B_FAIL_NONE) vrow = VP8B_FAIL_HEADERS;   // a parser that keeps no rows (the WEBP_B200_TOKEN_MAP debug paths)
        const int meta = b->statuses[3 * k + 2];
        if (!vp8b_vp8_failure_first(vrow, ah.fail_row, meta & 3, meta >> 8, (int)d.crop_y + (int)d.out_h, ah.levels != 0))
          it->status = (VP8StatusCode)ah.status;
      }
    }
    if (it->status == VP8_STATUS_OK && b->plan[b->img_item[k]].discard) it->status = VP8_STATUS_INVALID_PARAM;
    if (it->status != VP8_STATUS_OK && b->opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&it->config->output);
  }
  b->decoded = true;
  return true;
}

static VP8StatusCode first_failure(const WebPBatchItem* items, int n) {
  for (int i = 0; i < n; ++i) if (items[i].status != VP8_STATUS_OK) return items[i].status;
  return VP8_STATUS_OK;
}

static void drain_after_failure(WebPBatch* b) {
  jobs_drain(b);
  cudaStreamSynchronize(b->stream); cudaStreamSynchronize(b->ctx->pixel_stream); cudaStreamSynchronize(b->ctx->copy_stream);
  cudaGetLastError();
  b->queued = false;
}

// Queues one decode of a single-device batch (kernels, and with `download` the pixel copies).
static bool submit_single(WebPBatch* b, bool download) {
  if (b->ctx == nullptr) return true;   // nothing reached the device
  DeviceGuard guard(b->ctx->device);
  b->ctx->mu.lock();
  const bool ok = batch_enqueue(b, download);
  if (!ok) drain_after_failure(b);
  b->ctx->mu.unlock();
  if (!ok) {
    fprintf(stderr, "libwebp_b200: %s\n", g_last_error);
    fail_all(b->items, b->n, strstr(g_last_error, "cudaMalloc") ? VP8_STATUS_OUT_OF_MEMORY : VP8_STATUS_USER_ABORT);
  }
  return ok;
}

static VP8StatusCode wait_single(WebPBatch* b) {
  if (b->ctx != nullptr && b->queued) {
    DeviceGuard guard(b->ctx->device);
    if (!batch_finish(b)) {
      b->ctx->mu.lock(); drain_after_failure(b); b->ctx->mu.unlock();
      fprintf(stderr, "libwebp_b200: %s\n", g_last_error);
      fail_all(b->items, b->n, VP8_STATUS_USER_ABORT);
      return VP8_STATUS_USER_ABORT;
    }
  }
  return first_failure(b->items, b->n);
}

static void destroy_single(WebPBatch* b) {
  DeviceCtx* c = b->ctx;
  if (c == nullptr) { delete b; return; }
  DeviceGuard guard(c->device);
  if (b->queued) { batch_finish(b); }   // never free what the streams still use
  c->mu.lock();
  batch_release(b);
  c->mu.unlock();
}

// ---------------------------------------------------------------------------------------------------------
// WebPBatchOptions::devices: item i goes to devices[i % num_devices]. One host thread per device plans, uploads and
// queues its shard; nothing is exchanged between the devices (images are independent: SURVEY.md 8e).
static bool wants_shards(const WebPBatchOptions& o) { return o.devices != NULL && o.num_devices > 1; }

template <typename F>
static void for_each_shard(WebPBatch* parent, F fn) {
  std::vector<std::thread> th;
  for (size_t g = 1; g < parent->shards.size(); ++g) th.emplace_back([&, g]() { fn(parent->shards[g], (int)g); });
  fn(parent->shards[0], 0);
  for (auto& t : th) t.join();
}

static void shards_pull_statuses(WebPBatch* parent) {
  const int G = (int)parent->shards.size();
  std::vector<int> pos(G, 0);
  for (int i = 0; i < parent->n; ++i) { const int g = i % G; parent->items[i].status = parent->shards[g]->items[pos[g]++].status; }
}

static WebPBatch* create_any(WebPBatchItem* items, int num_items, const WebPBatchOptions* options, VP8StatusCode* status) {
  VP8StatusCode dummy;
  if (status == NULL) status = &dummy;
  *status = VP8_STATUS_INVALID_PARAM;
  if (items == NULL || num_items <= 0) return NULL;
  WebPBatchOptions opt;
  if (options != NULL) opt = *options; else WebPBatchOptionsInit(&opt);
  if (opt.num_devices < 0 || (opt.num_devices > 0 && opt.devices == NULL)) return NULL;
  if (opt.num_devices == 1) { opt.device = opt.devices[0]; }
  if (!wants_shards(opt)) { opt.devices = NULL; opt.num_devices = 0; return create_single(items, num_items, opt, status); }
  if (opt.stream != NULL) return NULL;   // a caller's stream belongs to one device
  const int G = std::min(opt.num_devices, num_items);
  WebPBatch* parent = new WebPBatch();
  parent->items = items; parent->n = num_items; parent->opt = opt;
  memset(&parent->timings, 0, sizeof(parent->timings));
  parent->shard_items.resize(num_items);
  std::vector<int> first(G + 1, 0);
  for (int g = 0; g < G; ++g) first[g + 1] = first[g] + (num_items - g + G - 1) / G;
  { std::vector<int> pos(first.begin(), first.end() - 1); for (int i = 0; i < num_items; ++i) parent->shard_items[pos[i % G]++] = items[i]; }
  parent->shards.assign(G, nullptr);
  std::vector<VP8StatusCode> sts(G, VP8_STATUS_OK);
  std::vector<std::string> errs(G);
  {
    std::vector<std::thread> th;
    auto make = [&](int g) {
      WebPBatchOptions o = opt;
      o.device = opt.devices[g]; o.devices = NULL; o.num_devices = 0;
      parent->shards[g] = create_single(parent->shard_items.data() + first[g], first[g + 1] - first[g], o, &sts[g]);
      errs[g] = g_last_error;
    };
    for (int g = 1; g < G; ++g) th.emplace_back(make, g);
    make(0);
    for (auto& t : th) t.join();
  }
  bool ok = true;
  for (int g = 0; g < G; ++g) if (parent->shards[g] == nullptr) { ok = false; *status = sts[g]; snprintf(g_last_error, sizeof(g_last_error), "%s", errs[g].c_str()); }
  if (!ok) {   // a shard without a batch has already marked its items; the others are given up as well (one call, one answer)
    for (int g = 0; g < G; ++g) {
      if (parent->shards[g] == nullptr) continue;
      WebPBatch* sb = parent->shards[g];
      for (int i = 0; i < sb->n; ++i) if (sb->items[i].status == VP8_STATUS_OK) {
        sb->items[i].status = *status;
        if (sb->items[i].config != NULL && opt.output == WEBP_BATCH_HOST) WebPFreeDecBuffer(&sb->items[i].config->output);
      }
    }
    std::vector<int> pos(first.begin(), first.end() - 1);
    for (int i = 0; i < num_items; ++i) items[i].status = parent->shard_items[pos[i % G]++].status;
    for (int g = 0; g < G; ++g) if (parent->shards[g] != nullptr) destroy_single(parent->shards[g]);
    delete parent;
    return NULL;
  }
  shards_pull_statuses(parent);
  *status = VP8_STATUS_OK;
  return parent;
}

static bool submit_any(WebPBatch* b, bool download) {
  if (b->shards.empty()) return submit_single(b, download);
  std::vector<char> ok(b->shards.size(), 1);
  for_each_shard(b, [&](WebPBatch* sb, int g) { ok[g] = submit_single(sb, download) ? 1 : 0; });
  b->queued = true;
  for (char c : ok) if (!c) return false;
  return true;
}

static VP8StatusCode wait_any(WebPBatch* b) {
  if (b->shards.empty()) return wait_single(b);
  for_each_shard(b, [&](WebPBatch* sb, int) { wait_single(sb); });
  b->queued = false; b->decoded = true;
  shards_pull_statuses(b);
  // stage times: the slowest shard's (the devices run side by side)
  memset(&b->timings, 0, sizeof(b->timings));
  for (WebPBatch* sb : b->shards) {
    if (sb->timings.total_ms >= b->timings.total_ms) { const int l = b->timings.launches; b->timings = sb->timings; b->timings.launches = l; }
    b->timings.launches += sb->timings.launches;
  }
  return first_failure(b->items, b->n);
}

extern "C" WebPBatch* WebPBatchCreate(WebPBatchItem* items, int num_items, const WebPBatchOptions* options,
                                      VP8StatusCode* status) {
  return create_any(items, num_items, options, status);
}

extern "C" VP8StatusCode WebPBatchDecode(WebPBatch* b) {
  if (b == NULL) return VP8_STATUS_INVALID_PARAM;
  if (b->queued) wait_any(b);
  submit_any(b, false);
  return wait_any(b);
}

// Asynchronous pair: WebPBatchSubmit = WebPBatchCreate + everything queued (uploads, kernels and, for host output, the
// pixel copies) without waiting; WebPBatchWait collects the statuses. Between the two the caller's input and output
// buffers belong to the library. Several submitted batches per device pipeline: see DeviceCtx.
extern "C" WebPBatch* WebPBatchSubmit(WebPBatchItem* items, int num_items, const WebPBatchOptions* options, VP8StatusCode* status) {
  VP8StatusCode dummy;
  if (status == NULL) status = &dummy;
  WebPBatch* b = create_any(items, num_items, options, status);
  if (b == NULL) return NULL;
  submit_any(b, b->opt.output == WEBP_BATCH_HOST);
  return b;
}

extern "C" VP8StatusCode WebPBatchWait(WebPBatch* b) {
  if (b == NULL) return VP8_STATUS_INVALID_PARAM;
  return wait_any(b);
}

static bool batch_download(WebPBatch* b) {
  DeviceCtx* ctx = b->ctx;
  if (b->opt.output != WEBP_BATCH_HOST || b->imgs.empty()) return true;
  if (!enqueue_download(b, 0, (int)b->imgs.size(), ctx->copy_stream, true)) return false;
  if (b->copy_done == nullptr) CU_TRY(cudaEventCreateWithFlags(&b->copy_done, cudaEventDisableTiming), "cudaEventCreate");
  CU_TRY(cudaEventRecord(b->copy_done, ctx->copy_stream), "cudaEventRecord");
  return true;
}

static VP8StatusCode download_single(WebPBatch* b) {
  if (b->ctx == nullptr) return VP8_STATUS_OK;
  DeviceGuard guard(b->ctx->device);
  b->ctx->mu.lock();
  bool ok = batch_download(b);
  b->ctx->mu.unlock();
  if (ok && b->copy_done != nullptr && cudaEventSynchronize(b->copy_done) != cudaSuccess) { set_error("download sync", cudaGetLastError()); ok = false; }
  if (!ok) { cudaStreamSynchronize(b->ctx->copy_stream); cudaGetLastError(); fprintf(stderr, "libwebp_b200: %s\n", g_last_error); return VP8_STATUS_USER_ABORT; }
  return VP8_STATUS_OK;
}

extern "C" VP8StatusCode WebPBatchDownload(WebPBatch* b) {
  if (b == NULL || !b->decoded) return VP8_STATUS_INVALID_PARAM;
  if (b->shards.empty()) return download_single(b);
  std::vector<VP8StatusCode> st(b->shards.size(), VP8_STATUS_OK);
  for_each_shard(b, [&](WebPBatch* sb, int g) { st[g] = download_single(sb); });
  for (VP8StatusCode s : st) if (s != VP8_STATUS_OK) return s;
  return VP8_STATUS_OK;
}

extern "C" void WebPBatchDestroy(WebPBatch* b) {
  if (b == NULL) return;
  if (b->shards.empty()) { destroy_single(b); return; }
  for (WebPBatch* sb : b->shards) destroy_single(sb);
  delete b;
}

extern "C" int WebPBatchOutput(const WebPBatch* b, int index, WebPBatchPlane* p) {
  if (b == NULL || p == NULL || index < 0 || index >= b->n) return 0;
  if (!b->shards.empty()) { const int G = (int)b->shards.size(); return WebPBatchOutput(b->shards[index % G], index / G, p); }
  if (b->plan[index].img < 0) return 0;
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
  p->device = b->ctx != nullptr ? b->ctx->device : -1;
  return 1;
}

extern "C" int WebPBatchGetTimings(const WebPBatch* b, WebPBatchTimings* t) {
  if (b == NULL || t == NULL) return 0;
  *t = b->timings;
  return 1;
}

// Test hook: see decode_batch.h. The scratch arrays are the device's, not the batch's: `scratch_owner` says whose data they hold.
extern "C" int WebPBatchDebugStages(WebPBatch* b, int item, uint32_t* mbinfo, int16_t* levels, size_t max_mb) {
  if (b == NULL || mbinfo == NULL || levels == NULL || item < 0 || item >= b->n) return -1;
  if (!b->shards.empty() || b->ctx == nullptr || !b->decoded || b->waves.size() != 1 || !vp8k_tokens_use_stream()) return -1;
  const int img = b->plan[item].img;
  if (img < 0) return -1;
  DeviceCtx* ctx = b->ctx;
  DeviceGuard guard(ctx->device);
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (ctx->scratch_owner != b) return -1;
  const ImgDesc& d = b->imgs[img];
  const size_t n = (size_t)d.mb_w * d.mb_h;
  if (n > max_mb) return -1;
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return -1;
  std::vector<uint32_t> toks(n * VP8B_TOKENS_PER_MB);
  std::vector<uint32_t> mt(2 * n);
  if (cudaMemcpy(mbinfo, (const uint8_t*)ctx->s_mbinfo.p + 16 * (size_t)d.mb_base, 16 * n, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(mt.data(), (const uint8_t*)ctx->s_mbtok.p + 8 * (size_t)d.mb_base, 8 * n, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(toks.data(), (const uint8_t*)ctx->s_tokens.p + 4 * (size_t)d.mb_base * VP8B_TOKENS_PER_MB, 4 * toks.size(), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  memset(levels, 0, sizeof(int16_t) * VP8B_COEFFS_PER_MB * n);
  for (size_t m = 0; m < n; ++m) {
    const uint32_t first = mt[2 * m], count = mt[2 * m + 1];
    if (count > VP8B_TOKENS_PER_MB || (size_t)first + count > toks.size()) return -1;
    for (uint32_t k = 0; k < count; ++k) {
      const uint32_t t = toks[first + k];
      const int mag = (int)((t >> 13) & 0xfffu);
      levels[m * VP8B_COEFFS_PER_MB + ((t >> 25) & 31u) * 16u + ((t >> 6) & 15u)] = (int16_t)((t >> 31) ? -mag : mag);
    }
  }
  return (int)n;
}

// ---------------------------------------------------------------------------------------------------------
// One-shot entry point. A batch whose device memory would not fit is cut into groups that run one after the other, and
// an image that cannot fit on its own fails alone (VP8_STATUS_OUT_OF_MEMORY): "one bad image never poisons the others"
// also holds for files that merely DECLARE huge dimensions.
static size_t item_device_bytes(const WebPBatchItem* it) {
  WebPBitstreamFeatures f;
  if (it->data == NULL || it->config == NULL || vp8b_get_features(it->data, it->data_size, &f) != VP8_STATUS_OK) return 0;
  const size_t w = (size_t)f.width, h = (size_t)f.height;
  size_t ow = w, oh = h;
  const WebPDecoderOptions* o = &it->config->options;
  if (o->use_cropping && o->crop_width > 0 && o->crop_height > 0) { ow = std::min(ow, (size_t)o->crop_width); oh = std::min(oh, (size_t)o->crop_height); }
  if (o->use_scaling && o->scaled_width > 0 && o->scaled_height > 0) { ow = (size_t)o->scaled_width; oh = (size_t)o->scaled_height; }
  size_t bytes = it->data_size + 4 * ow * oh + 4096;                       // input + output (at most 4 bytes per pixel)
  if (f.has_alpha || f.format == 2) bytes += 9 * w * h + (1u << 20);      // alpha plane, coded ARGB, de-banding / tile work
  return bytes;
}

static size_t device_room(int device) {   // free + what the library could give back on that device
  DeviceCtx* c = get_ctx(device);
  if (c == nullptr) return 0;
  DeviceGuard guard(c->device);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
  std::lock_guard<std::mutex> lk(c->mu);
  return free_b + c->cached_bytes;
}

static VP8StatusCode decode_group(WebPBatchItem* items, int n, const WebPBatchOptions* options, WebPBatchTimings* tm, size_t* nwaves) {
  VP8StatusCode st;
  WebPBatch* b = create_any(items, n, options, &st);
  if (b == NULL) {
    if (st == VP8_STATUS_OUT_OF_MEMORY && n > 1) {   // halve and retry: whatever cannot fit ends up alone and fails alone
      for (int i = 0; i < n; ++i) if (items[i].status == VP8_STATUS_OUT_OF_MEMORY) items[i].status = VP8_STATUS_OK;
      const int h = n / 2;
      const VP8StatusCode a = decode_group(items, h, options, tm, nwaves), c = decode_group(items + h, n - h, options, tm, nwaves);
      return a != VP8_STATUS_OK ? a : c;
    }
    return st != VP8_STATUS_OK ? st : VP8_STATUS_INVALID_PARAM;
  }
  submit_any(b, b->opt.output == WEBP_BATCH_HOST);   // downloads ride along, chunk by chunk
  st = wait_any(b);
  if (tm != NULL) *tm = b->timings;
  if (nwaves != NULL) *nwaves = b->waves.size();
  WebPBatchDestroy(b);
  return st;
}

extern "C" VP8StatusCode WebPDecodeBatch(WebPBatchItem* items, int num_items, const WebPBatchOptions* options) {
  static int trace = -1;
  if (trace < 0) trace = getenv("WEBP_B200_TRACE") != NULL;
  if (items == NULL || num_items <= 0) return VP8_STATUS_INVALID_PARAM;
  const auto t0 = std::chrono::steady_clock::now();
  WebPBatchTimings tm;
  memset(&tm, 0, sizeof(tm));
  size_t nwaves = 0;
  for (int i = 0; i < num_items; ++i) items[i].status = VP8_STATUS_OK;
  // ---- how much device memory the resident part of each item takes; groups of at most half of what is there
  const int ndev = (options != NULL && options->devices != NULL && options->num_devices > 1) ? options->num_devices : 1;
  size_t room = (size_t)-1;
  bool sized = false;
  std::vector<size_t> need;
  size_t total = 0;
  need.resize(num_items);
  for (int i = 0; i < num_items; ++i) { need[i] = item_device_bytes(&items[i]); total += need[i]; }
  if (total > ((size_t)1 << 30)) {   // small batches never come near the limit: skip the device query
    for (int g = 0; g < ndev; ++g) {
      const int dev = ndev > 1 ? options->devices[g] : (options != NULL ? options->device : -1);
      const size_t r = device_room(dev);
      if (r > 0) { room = std::min(room, r); sized = true; }
    }
  }
  VP8StatusCode result = VP8_STATUS_OK;
  if (!sized || total / ndev <= room / 2) {
    result = decode_group(items, num_items, options, &tm, &nwaves);
  } else {
    const size_t cap = room / 2 * ndev;
    int i0 = 0;
    while (i0 < num_items) {
      if (need[i0] > room / 2) {   // cannot fit even alone
        WebPBatchItem* it = &items[i0];
        it->status = VP8_STATUS_OUT_OF_MEMORY;
        WebPBitstreamFeatures f;   // the reference fills config->input before it allocates (webp_dec.c:761-767)
        if (it->config != NULL && vp8b_get_features(it->data, it->data_size, &f) == VP8_STATUS_OK) it->config->input = f;
        ++i0;
        continue;
      }
      int i1 = i0; size_t acc = 0;
      while (i1 < num_items && need[i1] <= room / 2 && (i1 == i0 || acc + need[i1] <= cap)) acc += need[i1++];
      decode_group(items + i0, i1 - i0, options, &tm, &nwaves);
      i0 = i1;
    }
    result = first_failure(items, num_items);
  }
  if (trace) {
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "libwebp_b200 trace: %d items, %zu waves: %.1f ms in the call (kernels %.1f = modes %.1f tokens %.1f recon %.1f filter %.1f emit %.1f alpha %.1f)\n",
            num_items, nwaves, std::chrono::duration<double, std::milli>(t1 - t0).count(), tm.total_ms, tm.modes_ms, tm.tokens_ms, tm.recon_ms,
            tm.filter_ms, tm.emit_ms, tm.alpha_ms);
  }
  const VP8StatusCode ff = first_failure(items, num_items);
  return ff != VP8_STATUS_OK ? ff : result;
}
