/* webp/decode_batch.h -- batched entry points of libwebp_b200 (new; no counterpart in the reference).
 *
 * The reference decodes one image per call on one host thread (WebPDecode, src/dec/webp_dec.c:752, which
 * loops VP8ParseIntraModeRow / VP8DecodeMB / VP8ProcessRow per macroblock row, src/dec/vp8_dec.c:646-674).
 * Here the unit of work is a batch: every stage runs as one CUDA kernel over all images of the batch, and
 * WebPDecode() is the batch of one. Per-item semantics (status codes, output buffer contract, options) are
 * those of WebPDecode. Images are independent, so a multi-GPU caller shards the item array by index and
 * gives each shard to one device (one process or host thread per device); there is no collective.
 *
 * Plain C ABI: pointers and sizes only. This is the surface a binding of the reference's language would
 * wrap. */
#ifndef WEBP_WEBP_DECODE_BATCH_H_
#define WEBP_WEBP_DECODE_BATCH_H_

#include "./decode.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WEBP_BATCH_ABI_VERSION 0x0101   /* 0x0101: devices / num_devices / stream carved out of WebPBatchOptions::pad (same size) */

typedef struct WebPBatchItem {
  const uint8_t* data;       /* one complete .webp file (RIFF or bare VP8), host memory */
  size_t data_size;
  WebPDecoderConfig* config; /* as for WebPDecode(); `input` is filled, `output` describes the destination */
  VP8StatusCode status;      /* out: per-item result, same codes WebPDecode() would return */
} WebPBatchItem;

typedef enum WebPBatchMemory {
  WEBP_BATCH_HOST = 0,  /* config->output is host memory, exactly as for WebPDecode() */
  WEBP_BATCH_DEVICE = 1 /* decoded pixels stay in device memory owned by the batch (see WebPBatchOutput) */
} WebPBatchMemory;

typedef struct WebPBatchOptions {
  int device;                /* CUDA device ordinal, -1 = the calling thread's current device */
  WebPBatchMemory output;    /* where decoded pixels end up */
  size_t scratch_bytes;      /* cap on device scratch per wave (0 = default: 85% of free memory) */
  int pipeline_waves;        /* split the batch into this many waves (0 = default: one wave unless scratch memory
                                forces more; downloads overlap the pixel stages chunk by chunk either way) */
  int num_devices;           /* > 1: shard the items over `devices`, item i on devices[i % num_devices]; one host
                                thread per device inside the call, no collective, no peer access (images are
                                independent). 0 = `device` alone. */
  const int* devices;        /* num_devices CUDA ordinals (read during the call only) */
  void* stream;              /* cudaStream_t of the caller on `device` (single device only), NULL = the library's own:
                                uploads and kernels are queued on it, so the batch runs behind whatever the caller
                                queued there before and ahead of what it queues after WebPBatchSubmit()/WebPBatchDecode()
                                returned; host-output pixel copies travel on the library's copy stream either way */
  uint32_t pad[2];
} WebPBatchOptions;

WEBP_EXTERN int WebPBatchOptionsInitInternal(WebPBatchOptions*, int);
static WEBP_INLINE int WebPBatchOptionsInit(WebPBatchOptions* o) {
  return WebPBatchOptionsInitInternal(o, WEBP_BATCH_ABI_VERSION);
}

/* One-shot: parse, upload, decode, and (WEBP_BATCH_HOST) download. Returns VP8_STATUS_OK when every item
 * decoded, else the first failing item's status; items[i].status always holds the per-item code. One bad
 * image never poisons the others. `options` may be NULL (device -1, host output). */
WEBP_EXTERN VP8StatusCode WebPDecodeBatch(WebPBatchItem* items, int num_items, const WebPBatchOptions* options);

/* Resident form, for callers that keep compressed data and pixels on the device between steps:
 *   WebPBatchCreate   parses containers on the host, allocates device memory, uploads the compressed bytes
 *   WebPBatchDecode   runs the kernels (can be repeated; inputs stay resident), waits for completion
 *   WebPBatchDownload copies the pixels into every item's config->output (host memory)
 *   WebPBatchOutput   device address + layout of item i's pixels
 *   WebPBatchDestroy  releases everything
 * `items` must stay valid for the lifetime of the batch. */
typedef struct WebPBatch WebPBatch;

WEBP_EXTERN WebPBatch* WebPBatchCreate(WebPBatchItem* items, int num_items, const WebPBatchOptions* options,
                                       VP8StatusCode* status);
WEBP_EXTERN VP8StatusCode WebPBatchDecode(WebPBatch* batch);
WEBP_EXTERN VP8StatusCode WebPBatchDownload(WebPBatch* batch);
WEBP_EXTERN void WebPBatchDestroy(WebPBatch* batch);

/* Asynchronous form of WebPDecodeBatch(), for callers that keep the device busy:
 *   WebPBatchSubmit  = WebPBatchCreate + everything queued on the device (uploads, kernels and, for WEBP_BATCH_HOST,
 *                      the pixel copies) WITHOUT waiting for any of it;
 *   WebPBatchWait    waits, fills items[i].status, returns what WebPDecodeBatch() would have returned;
 *   WebPBatchDestroy releases the batch.
 * Between Submit and Wait the input bytes and the output buffers of the items belong to the library (page-locked
 * memory keeps the copies asynchronous). Batches submitted to one device run in submission order, and the pixels of
 * batch k travel to the host while the kernels of batch k+1 run: two batches in flight hide the whole download.
 * Two host threads calling WebPDecodeBatch() on one device pipeline the same way. */
WEBP_EXTERN WebPBatch* WebPBatchSubmit(WebPBatchItem* items, int num_items, const WebPBatchOptions* options,
                                       VP8StatusCode* status);
WEBP_EXTERN VP8StatusCode WebPBatchWait(WebPBatch* batch);

typedef struct WebPBatchPlane {
  void* y_or_rgba; /* device pointer: RGB-family pixels, or the Y plane for MODE_YUV */
  void* u;
  void* v;
  int stride, uv_stride;
  int width, height;
  int device;      /* CUDA ordinal the pointers belong to (differs per item when the batch was sharded) */
} WebPBatchPlane;
WEBP_EXTERN int WebPBatchOutput(const WebPBatch* batch, int index, WebPBatchPlane* plane);

/* Animated files: every frame of the file decoded as ONE batch, then the reference's canvas reconstruction
 * (WebPAnimDecoderGetNext, src/demux/anim_decode.c:325-440: key-frame rule, blending, disposal) replayed on the host.
 * canvases receives frame_count canvases of 4 * canvas_width * canvas_height bytes one after the other, exactly what
 * successive WebPAnimDecoderGetNext() calls return; timestamps (may be NULL) their end times in ms. mode is one of
 * MODE_RGBA, MODE_BGRA, MODE_rgbA, MODE_bgrA (WebPAnimDecoderOptions::color_mode). Not an animated file, or a file the
 * reference's demuxer would refuse: VP8_STATUS_BITSTREAM_ERROR. (The reference's own WebPAnimDecoder also works on top
 * of this library, one GPU round trip per frame.) */
typedef struct WebPAnimBatchInfo {
  int canvas_width, canvas_height, frame_count, loop_count;
  uint32_t bgcolor;
  uint32_t pad[3];
} WebPAnimBatchInfo;
WEBP_EXTERN int WebPAnimBatchGetInfo(const uint8_t* data, size_t data_size, WebPAnimBatchInfo* info);
WEBP_EXTERN VP8StatusCode WebPAnimDecodeBatch(const uint8_t* data, size_t data_size, WEBP_CSP_MODE mode, uint8_t* canvases,
                                              size_t canvases_size, int* timestamps, const WebPBatchOptions* options);

/* Per-stage device time of the last WebPBatchDecode() on this batch, from CUDA events recorded on the
 * batch's own stream around each kernel (milliseconds; `launches` = kernels launched). */
typedef struct WebPBatchTimings {
  float total_ms;
  float modes_ms;    /* header + intra-mode parse (partition 0) */
  float tokens_ms;   /* coefficient token parse (incl. zeroing the coefficient planes) */
  float recon_ms;    /* dequantised IDCT/WHT + intra prediction wavefront */
  float filter_ms;   /* in-loop deblocking wavefront */
  float emit_ms;     /* upsample + YUV->RGB (or plane copy) */
  int launches;
  float alpha_ms;    /* ALPH chunks: VP8L header + pixel passes, inverse transforms, unfilter (0 without alpha) */
  uint32_t pad[4];
} WebPBatchTimings;
WEBP_EXTERN int WebPBatchGetTimings(const WebPBatch* batch, WebPBatchTimings* t);

/* Test hook (not part of the stable interface): what the two parse kernels left for item `item` of a batch that was just
   decoded -- MbInfo (4 words per macroblock: intra modes, non-zero codes, flags) and the coefficient LEVELS (400 per
   macroblock: blocks 0-23 and the Y2 block, parse order inside a block), read back from the device's scratch. Only valid
   straight after WebPBatchDecode / WebPBatchWait of a single-wave batch, before anything else runs on that device.
   Returns the number of macroblocks written (<= max_mb), -1 when the scratch is not this batch's any more or cannot be read. */
WEBP_EXTERN int WebPBatchDebugStages(WebPBatch* batch, int item, uint32_t* mbinfo, int16_t* levels, size_t max_mb);

/* Page-locked host memory for inputs/outputs (plain malloc works too, just slower over PCIe). */
WEBP_EXTERN void* WebPBatchHostAlloc(size_t size);
WEBP_EXTERN void WebPBatchHostFree(void* ptr);

/* Device memory kept between calls. Blocks released by a finished batch are cached (cudaMalloc/cudaFree of tens of GB
 * cost 0.1-1 s) up to a limit, by default a quarter of the device's memory (environment: WEBP_B200_CACHE_GB).
 * WebPBatchSetCacheLimit returns the previous limit; WebPBatchTrimCache gives every cached block, the shared wave
 * scratch and the library's page-locked staging back to the driver and returns the device bytes released.
 * device -1 = the calling thread's current device. */
WEBP_EXTERN size_t WebPBatchSetCacheLimit(int device, size_t bytes);
WEBP_EXTERN size_t WebPBatchTrimCache(int device);

/* Number of CUDA devices visible to the library; 0 means WebPDecode*() will fail (there is no CPU path). */
WEBP_EXTERN int WebPBatchDeviceCount(void);
/* Human-readable text for the last CUDA/runtime error seen by the calling thread ("" if none). */
WEBP_EXTERN const char* WebPBatchLastError(void);

#ifdef __cplusplus
}
#endif
#endif /* WEBP_WEBP_DECODE_BATCH_H_ */
