// vp8_kernels.cu -- the sm_100a kernels of the batched VP8 decoder and their launchers.
//
//   k_parse_modes   one warp per image: frame header + intra modes from partition 0 (serial bool decoding)
//   k_parse_tokens  one warp per token partition, the partitions of an image in one thread block, coupled
//                   through shared-memory progress counters (top non-zero context, vp8_dec.c:524-535)
//   k_reconstruct   one thread block per image, one warp per macroblock on a lag-2 wavefront; neighbour pixels
//                   live in shared memory, the HBM planes are write-only
//   k_loop_filter   same wavefront, macroblock tile staged through shared memory, in place in HBM
//   k_emit_rgb/yuv  fancy upsampling + YUV->RGB with 16-byte stores / plane copies
//
// All arithmetic lives in vp8_parse_core.h / vp8_pixel_core.h; this file is launch geometry, shared-memory
// carving and the inter-warp synchronisation.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define VP8_WAIT_PROGRESS(ptr, need)                         \
  do {                                                       \
    while (*(ptr) < (need)) { __nanosleep(40); }             \
    __threadfence_block();                                   \
  } while (0)
#define VP8_PUBLISH_PROGRESS(ptr, val) \
  do {                                 \
    __threadfence_block();             \
    *(ptr) = (val);                    \
  } while (0)

#include "vp8_kernels.h"
#include "vp8_parse_core.h"
#include "vp8_pixel_core.h"
#include "vp8_tokens_fsm.h"

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_parse_modes(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                    FrameHdr* hdrs, uint32_t* mbinfo, int first, int count) {
  extern __shared__ uint32_t top_modes[];   // mb_w words
  const int i = blockIdx.x;
  if (i >= count || threadIdx.x != 0) return;
  const ImgDesc im = imgs[first + i];
  FrameHdr* h = &hdrs[first + i];
  BoolDec br;
  int st = parse_frame_header(br, arena + im.in_off, im, h);
  if (st == VP8B_OK) st = parse_intra_modes(br, im, h, top_modes, mbinfo + 4 * (size_t)im.mb_base);
  h->status = st;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_parse_tokens(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs, FrameHdr* hdrs,
                               uint32_t* mbinfo, int16_t* coeffs, const int* __restrict__ ids, int P, int ctx_stride) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* probs = smem;                                   // 1056 B
  volatile int* progress = (volatile int*)(smem + 1056);   // P ints (+ status word)
  uint16_t* topctx = (uint16_t*)(smem + 1056 + 4 * (VP8B_MAX_PARTS + 1) + 12);   // (P+1) * ctx_stride
  const int img = ids[blockIdx.x];
  const ImgDesc im = imgs[img];
  FrameHdr* h = &hdrs[img];
  const int tid = threadIdx.x, lane = tid & 31, part = tid >> 5;
  for (int k = tid; k < 264; k += blockDim.x) ((uint32_t*)probs)[k] = ((const uint32_t*)h->prob)[k];
  if (tid < P) progress[tid] = 0;
  if (tid == 0) progress[VP8B_MAX_PARTS] = (h->status == VP8B_OK && h->num_parts == P) ? 1 : 0;
  __syncthreads();
  if (!progress[VP8B_MAX_PARTS]) {   // header failed (or, never expected, the host pre-scan disagreed)
    if (tid == 0 && h->status == VP8B_OK) h->status = VP8B_BITSTREAM_ERROR;
    return;
  }
  if (lane != 0 || part >= im.mb_h) return;
  (void)ctx_stride;
  TokenPart tp;
  token_part_init(tp, arena + im.in_off, h, part);
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  for (int my = part; my < im.mb_h; my += P) {
    parse_token_row(tp, im, h, part, my, probs, topctx, progress, mbi, cf);
  }
  if (tp.status != VP8B_OK) h->status = tp.status;
}

// ---------------------------------------------------------------------------------------------------------
// Lane-parallel token parser (vp8_tokens_fsm.h). A block owns IPB images x P partitions = IPB*P streams; stream j
// sits in warp j / LPW, lane j % LPW (lanes >= LPW of a warp stay idle: fewer streams per warp means fewer
// divergent block-end branches per iteration, more streams per warp means fewer issue slots per decode).
#define TOK_IMG_BYTES 1152   // sizeof(TokImage) rounded up to 16
#define TOK_TAB_BYTES 192    // sizeof(TokTables)

template <int LPW>
__global__ void __launch_bounds__(256) k_parse_tokens_fsm(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                          FrameHdr* hdrs, uint32_t* mbinfo, int16_t* coeffs,
                                                          const int* __restrict__ ids, int count, int P, int ipb, int ctx_stride) {
  extern __shared__ __align__(16) uint8_t smem[];
  TokTables* tables = reinterpret_cast<TokTables*>(smem);
  uint8_t* img_mem = smem + TOK_TAB_BYTES;                                         // ipb * TOK_IMG_BYTES
  int* progress = reinterpret_cast<int*>(img_mem + (size_t)ipb * TOK_IMG_BYTES);   // ipb * 8 ints
  uint16_t* ctx_mem = reinterpret_cast<uint16_t*>(progress + ipb * VP8B_MAX_PARTS); // ipb * (P+1) * ctx_stride
  const int tid = threadIdx.x, nthreads = blockDim.x;
  tk_tables_fill(tables, tid, nthreads);
  for (int k = tid; k < ipb * VP8B_MAX_PARTS; k += nthreads) progress[k] = 0;
  for (int slot = 0; slot < ipb; ++slot) {
    const int g = blockIdx.x * ipb + slot;
    if (g < count) tk_image_fill(reinterpret_cast<TokImage*>(img_mem + (size_t)slot * TOK_IMG_BYTES), &hdrs[ids[g]], P, tid, nthreads);
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int j = warp * LPW + lane;          // stream inside the block
  if (lane >= LPW || j >= ipb * P) return;
  const int slot = j / P, part = j % P;
  const int g = blockIdx.x * ipb + slot;
  if (g >= count) return;
  const int img = ids[g];
  const TokImage* timg = reinterpret_cast<const TokImage*>(img_mem + (size_t)slot * TOK_IMG_BYTES);
  FrameHdr* h = &hdrs[img];
  if (!timg->ok) {   // header failed (or, never expected, the host pre-scan disagreed with the device parse)
    if (part == 0 && h->status == VP8B_OK) h->status = VP8B_BITSTREAM_ERROR;
    return;
  }
  const ImgDesc im = imgs[img];
  if (part >= im.mb_h) return;
  TokShared sh;
  sh.img = timg;
  sh.img_s = tk_saddr_of(timg);
  sh.tab_s = tk_saddr_of(tables);
  asm volatile("" : "+r"(sh.img_s), "+r"(sh.tab_s));   // keep both as plain registers (no per-iteration cvta)
  sh.topctx = ctx_mem + (size_t)slot * (P + 1) * ctx_stride;
  sh.progress = progress + slot * VP8B_MAX_PARTS;
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  const uint32_t* arena32 = reinterpret_cast<const uint32_t*>(arena);
  TokLane L;
  tk_lane_init(L, arena32, im.in_off, h, part);
  while (L.phase != 2) {
    if (L.phase == 0) tk_mb_start(L, sh, im, P, mbi);
    if (L.phase == 1) tk_step(L, sh, im, P, arena32, mbi, cf);
  }
  if (L.status != VP8B_OK) h->status = L.status;
}

// ---------------------------------------------------------------------------------------------------------
#define RECON_WARPS 8

__global__ void __launch_bounds__(32 * RECON_WARPS) k_reconstruct(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                                  uint32_t* mbinfo, const int16_t* __restrict__ coeffs,
                                                                  uint8_t* yuv, int first) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int img = first + blockIdx.x;
  if (hdrs[img].status != VP8B_OK) return;
  const ImgDesc im = imgs[img];
  const int mb_w = im.mb_w, mb_h = im.mb_h;
  const int warp = threadIdx.x >> 5;
  ReconWs& ws = *reinterpret_cast<ReconWs*>(smem + sizeof(ReconWs) * warp);
  ReconCtx cx;
  recon_ctx_bind(cx, smem + sizeof(ReconWs) * RECON_WARPS, mb_w, mb_h);
  const size_t nmb = (size_t)mb_w * mb_h;
  uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  uint8_t* up = yp + nmb * 256;
  uint8_t* vp = up + nmb * 64;
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  const int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  const int steps = mb_w + 2 * (mb_h - 1);
  for (int d = 0; d < steps; ++d) {
    // rows with a macroblock on this anti-diagonal: mx = d - 2*my in [0, mb_w)
    const int my_lo = (d - mb_w + 2 > 0) ? (d - mb_w + 2) >> 1 : 0;
    const int my_hi = (d >> 1) < mb_h - 1 ? (d >> 1) : mb_h - 1;
    for (int my = my_lo + warp; my <= my_hi; my += RECON_WARPS) {
      const int mx = d - 2 * my;
      const size_t idx = (size_t)my * mb_w + mx;
      recon_macroblock(ws, cx, mx, my, mb_w, mbi + 4 * idx, cf + idx * VP8B_COEFFS_PER_MB, yp, up, vp);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
#define FILTER_WARPS 8

__global__ void __launch_bounds__(32 * FILTER_WARPS) k_loop_filter(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                                   const uint32_t* __restrict__ mbinfo, uint8_t* yuv, int first) {
  __shared__ __align__(16) FilterWs wss[FILTER_WARPS];
  __shared__ uint8_t fstr[32];
  const int img = first + blockIdx.x;
  const FrameHdr* h = &hdrs[img];
  if (h->status != VP8B_OK || h->filter_type == 0) return;
  const ImgDesc im = imgs[img];
  const int mb_w = im.mb_w, mb_h = im.mb_h;
  const int filter_type = h->filter_type;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) fstr[threadIdx.x] = ((const uint8_t*)h->fstr)[threadIdx.x];
  __syncthreads();
  FilterWs& ws = wss[warp];
  const size_t nmb = (size_t)mb_w * mb_h;
  uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  uint8_t* up = yp + nmb * 256;
  uint8_t* vp = up + nmb * 64;
  const uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  const int steps = mb_w + 2 * (mb_h - 1);
  for (int d = 0; d < steps; ++d) {
    const int my_lo = (d - mb_w + 2 > 0) ? (d - mb_w + 2) >> 1 : 0;
    const int my_hi = (d >> 1) < mb_h - 1 ? (d >> 1) : mb_h - 1;
    for (int my = my_lo + warp; my <= my_hi; my += FILTER_WARPS) {
      const int mx = d - 2 * my;
      const uint32_t w = mbi[4 * ((size_t)my * mb_w + mx) + 3];
      const uint8_t* fs = fstr + 8 * ((w >> MBW_SEG_SHIFT) & 3) + ((w & MBW_I4X4) ? 4 : 0);
      filter_macroblock(ws, mx, my, mb_w, filter_type, fs, (w & MBW_INNER) != 0, yp, up, vp);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
#define EMIT_THREADS 256

__global__ void __launch_bounds__(EMIT_THREADS) k_emit(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                       const uint8_t* __restrict__ yuv, uint8_t* out, int first,
                                                       int blocks_per_image) {
  const int img = first + blockIdx.x / blocks_per_image;
  const int chunk = blockIdx.x % blocks_per_image;
  if (hdrs[img].status != VP8B_OK) return;
  const ImgDesc im = imgs[img];
  const size_t nmb = (size_t)im.mb_w * im.mb_h;
  const uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  const uint8_t* up = yp + nmb * 256;
  const uint8_t* vp = up + nmb * 64;
  uint8_t* o = out + im.out_off;
  const int t = chunk * EMIT_THREADS + threadIdx.x;
  if (im.csp == 11) {   // MODE_YUV: 16-byte chunks of Y rows, then U rows, then V rows
    const int w = im.width, h = im.height, uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
    const int qy = (w + 15) >> 4, quv = (uvw + 15) >> 4;
    const int ny = qy * h, nuv = quv * uvh;
    if (t < ny) emit_yuv_chunk(im, yp, up, vp, o, 0, t % qy, t / qy);
    else if (t < ny + nuv) emit_yuv_chunk(im, yp, up, vp, o, 1, (t - ny) % quv, (t - ny) / quv);
    else if (t < ny + 2 * nuv) emit_yuv_chunk(im, yp, up, vp, o, 2, (t - ny - nuv) % quv, (t - ny - nuv) / quv);
  } else {
    const int qw = (im.width + 3) >> 2;
    if (t < qw * im.height) emit_rgb_quad(im, yp, up, vp, o, t % qw, t / qw);
  }
}

// =========================================================================================================
// Launchers (plain C interface for vp8_batch.cu).
static size_t recon_smem_bytes(int max_mb_w, int max_mb_h) {
  return sizeof(ReconWs) * RECON_WARPS + ((recon_ctx_bytes(max_mb_w, max_mb_h) + 15) & ~(size_t)15);
}

static size_t tokens_smem_bytes(int P, int max_mb_w) {
  return 1056 + 4 * (VP8B_MAX_PARTS + 1) + 12 + (size_t)(P + 1) * max_mb_w * 2;
}

extern "C" cudaError_t vp8k_configure(int max_mb_w, int max_mb_h) {
  cudaError_t e = cudaFuncSetAttribute(k_reconstruct, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)recon_smem_bytes(max_mb_w, max_mb_h));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_parse_tokens, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)tokens_smem_bytes(VP8B_MAX_PARTS, max_mb_w));
}

extern "C" void vp8k_parse_modes(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                 int first, int count, int max_mb_w) {
  k_parse_modes<<<count, 32, (size_t)max_mb_w * 4, s>>>(arena, imgs, hdrs, mbinfo, first, count);
}

static size_t tokens_fsm_smem_bytes(int P, int ipb, int max_mb_w) {
  return TOK_TAB_BYTES + (size_t)ipb * TOK_IMG_BYTES + (size_t)ipb * VP8B_MAX_PARTS * 4 + (size_t)ipb * (P + 1) * max_mb_w * 2;
}

template <int LPW>
static void launch_tokens_fsm(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                              int16_t* coeffs, const int* ids, int count, int P, int max_mb_w, int warps) {
  int ipb = (LPW * warps) / P;                 // images per block
  if (ipb < 1) ipb = 1;
  size_t smem = tokens_fsm_smem_bytes(P, ipb, max_mb_w);
  while (ipb > 1 && smem > 200 * 1024) { ipb >>= 1; smem = tokens_fsm_smem_bytes(P, ipb, max_mb_w); }
  const int streams = ipb * P;
  const int threads = ((streams + LPW - 1) / LPW) * 32;
  const int blocks = (count + ipb - 1) / ipb;
  cudaFuncSetAttribute(k_parse_tokens_fsm<LPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_parse_tokens_fsm<LPW><<<blocks, threads, smem, s>>>(arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, ipb, max_mb_w);
}

// Streams per warp: enough warps to give every SM sub-partition one or two, as few lanes per warp as that allows.
static int pick_lpw(int streams) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("WEBP_B200_TOKEN_LPW"); forced = e ? atoi(e) : 0; }
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
  int lpw = 1;
  while (lpw < 32 && streams / lpw > 148 * 4 * 2) lpw <<= 1;
  return lpw;
}

extern "C" void vp8k_parse_tokens(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                  int16_t* coeffs, const int* ids, int count, int P, int max_mb_w) {
  // Two mappings of the same parse. Few streams (a few per SM sub-partition): one warp per partition, straight-line
  // code on one lane, latency hidden by the other warps. Many streams: the lane-parallel state machine, which
  // spends ~8x fewer issue slots per decode. WEBP_B200_TOKEN_MAP=warp|lanes forces one of them (A/B timing).
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("WEBP_B200_TOKEN_MAP");
    forced = (e && e[0] == 'w') ? 1 : (e && e[0] == 'l') ? 2 : 0;
  }
  const int use_warp_map = forced ? (forced == 1) : (count * P < 16384);
  if (use_warp_map) {
    k_parse_tokens<<<count, 32 * P, tokens_smem_bytes(P, max_mb_w), s>>>(arena, imgs, hdrs, mbinfo, coeffs, ids, P, max_mb_w);
    return;
  }
  const int lpw = pick_lpw(count * P);
  const int warps = 2;
  switch (lpw) {
    case 1: launch_tokens_fsm<1>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
    case 2: launch_tokens_fsm<2>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
    case 4: launch_tokens_fsm<4>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
    case 8: launch_tokens_fsm<8>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
    case 16: launch_tokens_fsm<16>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
    default: launch_tokens_fsm<32>(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, warps); break;
  }
}

extern "C" void vp8k_reconstruct(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, uint32_t* mbinfo, const int16_t* coeffs,
                                 uint8_t* yuv, int first, int count, int max_mb_w, int max_mb_h) {
  k_reconstruct<<<count, 32 * RECON_WARPS, recon_smem_bytes(max_mb_w, max_mb_h), s>>>(imgs, hdrs, mbinfo, coeffs, yuv, first);
}

extern "C" void vp8k_loop_filter(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint32_t* mbinfo, uint8_t* yuv,
                                 int first, int count) {
  k_loop_filter<<<count, 32 * FILTER_WARPS, 0, s>>>(imgs, hdrs, mbinfo, yuv, first);
}

extern "C" void vp8k_emit(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, uint8_t* out, int first,
                          int count, int max_units) {
  const int bpi = (max_units + EMIT_THREADS - 1) / EMIT_THREADS;
  k_emit<<<(unsigned)count * (unsigned)bpi, EMIT_THREADS, 0, s>>>(imgs, hdrs, yuv, out, first, bpi);
}
