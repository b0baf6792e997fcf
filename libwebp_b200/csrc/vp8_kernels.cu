// vp8_kernels.cu -- the sm_100a kernels of the batched VP8 decoder and their launchers.
//
//   k_parse_modes   one warp per image: frame header + intra modes from partition 0 (serial bool decoding);
//                   k_parse_modes_lockstep: the same as lockstep lanes of a table of tree nodes (A/B, slower)
//   k_parse_tokens_fp  (default token parse) one LANE per token partition, fp32 boolean decoder, one decode per lane and
//                   step, token stream out; k_parse_tokens / _lockstep / _fsm: the older mappings (dense level plane)
//   k_reconstruct   one thread block per image, one warp per macroblock ROW: rows trail each other by two macroblocks,
//                   ordered by progress counters in shared memory; neighbour pixels live in shared memory, the HBM
//                   planes are write-only (ROWS = 0: anti-diagonals with a block-wide barrier, A/B)
//   k_loop_filter   same row-per-warp wavefront, macroblock tile staged through shared memory, in place in HBM
//   k_emit          fancy upsampling + YUV->RGB with 16-byte stores / plane copies
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
#include "vp8_tokens_lockstep.h"
#include "vp8_tokens_fp.h"
#include "vp8_literal.h"
#define AL_BLOCK_SYNC() __syncthreads()
#include "vp8_modes_lockstep.h"
#include "vp8l_alpha_core.h"
#include "vp8l_lossless_core.h"

// ---------------------------------------------------------------------------------------------------------
#define MODES_WARPS 28   // one image per warp, seven per SM sub-partition (see k_parse_tokens)
// `lanes` images per warp (lanes 0 .. lanes-1 each parse their own image, SIMT-divergent where their syntax differs).
__global__ void __launch_bounds__(32 * MODES_WARPS, 1) k_parse_modes(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                                     FrameHdr* hdrs, uint32_t* mbinfo, int first, int count, int ipb,
                                                                     int max_mb_w, int lanes) {
  extern __shared__ uint32_t top_modes_all[];   // ipb x max_mb_w words
  __shared__ uint8_t bprob[900];                // kVp8BModeProba, out of the constant bank (indexed per decode)
  const int lane = threadIdx.x & 31;
  const int warp = (threadIdx.x >> 5) * lanes + lane;   // slot of this lane's image inside the block
  const int i = blockIdx.x * ipb + warp;
  for (int k = threadIdx.x; k < 900; k += blockDim.x) bprob[k] = kVp8BModeProba[k];
  __syncthreads();
  if (i >= count || lane >= lanes || warp >= ipb) return;
  const ImgDesc im = imgs[first + i];
  FrameHdr* h = &hdrs[first + i];
  // a whole-picture VP8L image has no VP8 frame: a status that is not OK makes every pixel kernel pass it by; its own
  // status comes from the VP8L passes (k_alpha_header / k_alpha_pixels / k_lossless_finish)
  if (im.flags & VP8B_FLAG_LOSSLESS) { h->status = VP8B_NOT_A_VP8_FRAME; return; }
  BoolDec br;
  int st = parse_frame_header(br, arena + im.in_off, im, h);
  int fail_row = st == VP8B_OK ? VP8B_FAIL_NONE : VP8B_FAIL_HEADERS;
  if (st == VP8B_OK) st = parse_intra_modes(br, im, h, top_modes_all + (size_t)warp * max_mb_w, bprob, mbinfo + 4 * (size_t)im.mb_base, &fail_row);
  h->fail_row = fail_row;
  h->modes_status = VP8B_OK;
  h->all_rows = h->rows;
  if (st != VP8B_OK && fail_row > 0 && fail_row != VP8B_FAIL_NONE) {   // see FrameHdr::modes_status
    h->modes_status = st;
    h->rows = fail_row;
    st = VP8B_OK;
  }
  h->status = st;
}

// ---------------------------------------------------------------------------------------------------------
// K1, second mapping (vp8_modes_lockstep.h): `lanes` images per warp as lockstep lanes of one table-driven state machine,
// four warps per block (one per SM sub-partition when the launch has one block per SM). Shared memory: node table |
// kVp8BModeProba | one 16-byte row of fixed probabilities per image | the images' top-mode rows.
#define MODESL_WARPS 4
#define MODESL_TAB_BYTES 176     // ML_TAB_BYTES padded to 16
#define MODESL_BPROB_BYTES 912   // 900 padded to 16
#ifndef ML_GROUPS_PER_VOTE
#define ML_GROUPS_PER_VOTE 8
#endif
__global__ void __launch_bounds__(32 * MODESL_WARPS, 1) k_parse_modes_lockstep(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                                               FrameHdr* hdrs, uint32_t* mbinfo, int first, int count,
                                                                               int max_mb_w, int lanes) {
  extern __shared__ __align__(16) uint8_t msm[];
  uint32_t* tab = reinterpret_cast<uint32_t*>(msm);
  uint8_t* bprob = msm + MODESL_TAB_BYTES;
  const int ipb = MODESL_WARPS * lanes;
  uint8_t* rows = bprob + MODESL_BPROB_BYTES;
  uint32_t* tops = reinterpret_cast<uint32_t*>(rows + (size_t)ipb * ML_ROW_BYTES);
  ml_table_fill(tab, threadIdx.x, blockDim.x);
  for (int k = threadIdx.x; k < 900; k += blockDim.x) bprob[k] = kVp8BModeProba[k];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int slot = (threadIdx.x >> 5) * lanes + lane;
  const int i = blockIdx.x * ipb + slot;
  const bool mine = lane < lanes && i < count;
  FrameHdr* h = mine ? &hdrs[first + i] : nullptr;
  ImgDesc im;
  BoolDec br;
  int st = VP8B_NOT_A_VP8_FRAME;
  if (mine) {
    im = imgs[first + i];
    // (a whole-picture VP8L image has no VP8 frame, see k_parse_modes)
    if (!(im.flags & VP8B_FLAG_LOSSLESS)) st = parse_frame_header(br, arena + im.in_off, im, h);
  }
  const bool run = mine && st == VP8B_OK;
  const unsigned mask = __ballot_sync(0xffffffffu, run);
  if (!run) {
    if (mine) {
      h->status = st;
      if (!(im.flags & VP8B_FLAG_LOSSLESS)) { h->fail_row = VP8B_FAIL_HEADERS; h->modes_status = VP8B_OK; h->all_rows = h->rows; }
    }
    return;
  }
  MlCtx c;
  c.tab_s = tk_saddr_of(tab); c.bprob_s = tk_saddr_of(bprob);
  c.row_s = tk_saddr_of(rows + (size_t)slot * ML_ROW_BYTES);
  c.top_s = tk_saddr_of(tops + (size_t)slot * max_mb_w);
  c.out = mbinfo + 4 * (size_t)im.mb_base;
  c.mb_w = im.mb_w; c.mb_h = h->rows;
  c.skip_node8 = h->use_skip ? 8u * ML_SKIP : 8u * ML_I16; c.skip_off = h->use_skip ? ML_OFF_SKIP : ML_OFF_I16;
  c.first_node8 = h->update_map ? 8u * ML_S0 : c.skip_node8; c.first_off = h->update_map ? ML_OFF_S0 : c.skip_off;
  c.k.mant_mask = 0x007fffffu; c.k.exp46 = TF_EXP46;
  ml_row_fill(rows + (size_t)slot * ML_ROW_BYTES, h);
  MlLane L;
  ml_start(L, c, br);
  while (__any_sync(mask, L.alive)) {
    for (int r = 0; r < ML_GROUPS_PER_VOTE; ++r) ml_group(L, c);
  }
  // FrameHdr bookkeeping as k_parse_modes leaves it
  int stm = L.status;
  const int fail_row = L.fail_row;
  h->fail_row = fail_row;
  h->modes_status = VP8B_OK;
  h->all_rows = h->rows;
  if (stm != VP8B_OK && fail_row > 0 && fail_row != VP8B_FAIL_NONE) {   // see FrameHdr::modes_status
    h->modes_status = stm;
    h->rows = fail_row;
    stm = VP8B_OK;
  }
  h->status = stm;
}

// ---------------------------------------------------------------------------------------------------------
// Shared memory: probabilities by position 2244 B (padded to 2256) | progress P+1 ints (padded to 48 B) | top contexts.
#define TOKW_PROGRESS 2256
#define TOKW_CTX (TOKW_PROGRESS + 48)
#define TOKW_MAX_WARPS 28   // 7 per SM sub-partition: what the register file holds at 64 registers per thread
// A block owns ipb images x P partitions; warp w parses partition w % P of the block's image w / P. Packing 28
// warps into one block per SM puts exactly seven streams on every sub-partition (single-warp blocks are spread by
// the hardware as it sees fit, and the sub-partition that gets an eighth stream finishes last).
__global__ void __launch_bounds__(32 * TOKW_MAX_WARPS, 1) k_parse_tokens(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                                      FrameHdr* hdrs, uint32_t* mbinfo, int16_t* coeffs,
                                                                      const int* __restrict__ ids, int count, int P, int ipb, int slot_bytes) {
  extern __shared__ __align__(16) uint8_t smem_all[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = warp / P, part = warp % P;
  const int g = blockIdx.x * ipb + slot;
  const int live = slot < ipb && g < count;
  uint8_t* smem = smem_all + (size_t)(live ? slot : 0) * slot_bytes;
  uint8_t* probs = smem;
  volatile int* progress = (volatile int*)(smem + TOKW_PROGRESS);   // P ints (+ status word)
  uint16_t* topctx = (uint16_t*)(smem + TOKW_CTX);                  // (P+1) * mb_w
  const int img = live ? ids[g] : 0;
  const ImgDesc im = imgs[img];
  FrameHdr* h = &hdrs[img];
  if (live) {
    const int t = part * 32 + lane;   // this image's threads
    for (int k = t; k < VP8B_POSPROB_BYTES; k += 32 * P) probs[k] = posprob_byte(h->prob, k);
    if (t < P) progress[t] = 0;
    if (t == 0) progress[VP8B_MAX_PARTS] = (h->status == VP8B_OK && h->num_parts == P) ? 1 : 0;
  }
  __syncthreads();
  if (!live) return;
  if (!progress[VP8B_MAX_PARTS]) {   // header failed (or, never expected, the host pre-scan disagreed)
    if (part == 0 && lane == 0 && h->status == VP8B_OK) h->status = VP8B_BITSTREAM_ERROR;
    return;
  }
  const int rows = h->rows;
  if (lane != 0 || part >= rows) return;
  TokenPart tp;
  token_part_init(tp, arena + im.in_off, h, part);
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  for (int my = part; my < rows; my += P) {
    parse_token_row(tp, im, h, part, my, probs, topctx, progress, mbi, cf);
  }
  if (tp.status != VP8B_OK) h->status = tp.status;
}

// ---------------------------------------------------------------------------------------------------------
// Lane-parallel token parser (vp8_tokens_fsm.h). A block owns ipb images x P partitions = S streams. Stream j
// sits in parsing warp j / lpw, lane j % lpw (lanes >= lpw of a warp stay idle: fewer streams per warp means
// fewer divergent block-end branches per iteration, more streams per warp means fewer issue slots per decode).
// The warp after the `cw` parsing warps is the producer: it keeps the streams' shared-memory rings topped up
// with cp.async copies from the input arena.
struct TokLayout {   // byte offsets inside the block's dynamic shared memory
  uint32_t images, progress, ctl, rings, ctx, total;
};
__host__ __device__ static inline TokLayout tok_layout(int P, int ipb, int ctx_stride) {
  TokLayout t;
  const uint32_t S = (uint32_t)(ipb * P);
  t.images = (TOK_TAB_BYTES + 15u) & ~15u;
  t.progress = t.images + (uint32_t)ipb * TOK_IMG_BYTES;
  t.ctl = t.progress + (uint32_t)ipb * VP8B_MAX_PARTS * 4u;
  t.rings = (t.ctl + S * 8u + 15u) & ~15u;
  t.ctx = t.rings + S * TK_RING_BYTES;
  t.total = t.ctx + (uint32_t)ipb * (uint32_t)(P + 1) * (uint32_t)ctx_stride * 2u;
  return t;
}

__global__ void __launch_bounds__(32 * 9) k_parse_tokens_fsm(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                             FrameHdr* hdrs, uint32_t* mbinfo, int16_t* coeffs,
                                                             const int* __restrict__ ids, int count, int P, int ipb, int lpw,
                                                             int cw, int ctx_stride) {
  extern __shared__ __align__(16) uint8_t smem[];
  const TokLayout lay = tok_layout(P, ipb, ctx_stride);
  TokTables* tables = reinterpret_cast<TokTables*>(smem);
  uint8_t* img_mem = smem + lay.images;
  int* progress = reinterpret_cast<int*>(smem + lay.progress);
  TokStreamCtl* ctl = reinterpret_cast<TokStreamCtl*>(smem + lay.ctl);
  uint16_t* ctx_mem = reinterpret_cast<uint16_t*>(smem + lay.ctx);
  const tk_saddr smem_s = tk_saddr_of(smem);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int S = ipb * P;
  // ---- prologue 1: tables, per-image blocks, progress counters
  tk_tables_fill(tables, tid, nthreads);
  for (int k = tid; k < ipb * VP8B_MAX_PARTS; k += nthreads) progress[k] = 0;
  for (int slot = 0; slot < ipb; ++slot) {
    const int g = blockIdx.x * ipb + slot;
    TokImage* ti = reinterpret_cast<TokImage*>(img_mem + (size_t)slot * TOK_IMG_BYTES);
    if (g < count) tk_image_fill(ti, &hdrs[ids[g]], P, tid, nthreads);
    else if (tid == 0) ti->ok = 0;
  }
  __syncthreads();
  // ---- prologue 2: first fill of every live stream's ring
  for (int k = tid; k < S * TK_RING_CHUNKS; k += nthreads) {
    const int j = k / TK_RING_CHUNKS, slot = j / P, part = j % P;
    const TokImage* ti = reinterpret_cast<const TokImage*>(img_mem + (size_t)slot * TOK_IMG_BYTES);
    if (!ti->ok) continue;
    const int img = ids[blockIdx.x * ipb + slot];
    if (part >= hdrs[img].rows) continue;
    tk_stream_prefill(smem_s + lay.rings + (uint32_t)j * TK_RING_BYTES, arena, tk_stream_start(imgs[img].in_off, &hdrs[img], part),
                      k % TK_RING_CHUNKS);
  }
  tk_copy_wait();
  for (int j = tid; j < S; j += nthreads) {
    const int slot = j / P, part = j % P;
    const TokImage* ti = reinterpret_cast<const TokImage*>(img_mem + (size_t)slot * TOK_IMG_BYTES);
    int live = ti->ok;
    if (live) {
      const int img = ids[blockIdx.x * ipb + slot];
      live = part < hdrs[img].rows;
      if (live) tk_stream_open(&ctl[j], tk_stream_start(imgs[img].in_off, &hdrs[img], part));
    }
    if (!live) { ctl[j].rd_w = TK_STREAM_DONE; ctl[j].filled_c = 0; }
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  if (warp == cw) {
    // ---- producer warp: lane l serves streams l, l + 32, ...
    for (;;) {
      int live = 0;
      for (int j = lane; j < S; j += 32) {
        const tk_saddr ctl_s = smem_s + lay.ctl + (uint32_t)j * 8u;
        const uint32_t filled = tk_ldsv_u32(ctl_s + 4);
        const uint32_t nf = tk_stream_topup(ctl_s, smem_s + lay.rings + (uint32_t)j * TK_RING_BYTES, arena, filled);
        if (nf != 0) live = 1;
        if (nf > filled) {
          tk_copy_wait();
          __threadfence_block();
          tk_stsv_u32(ctl_s + 4, nf);
        }
      }
      if (!__any_sync(0xffffffffu, live)) break;
      __nanosleep(400);
    }
    return;
  }
  if (warp > cw) return;
  // ---- parsing warps
  const int j = warp * lpw + lane;          // stream inside the block
  if (lane >= lpw || j >= S) return;
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
  const int rows = h->rows;
  if (part >= rows) return;
  TokShared sh;
  sh.img = timg;
  sh.img_s = tk_saddr_of(timg);
  sh.tab_s = tk_saddr_of(tables);
  asm volatile("" : "+r"(sh.img_s), "+r"(sh.tab_s));   // keep both as plain registers (no per-iteration cvta)
  sh.topctx = ctx_mem + (size_t)slot * (P + 1) * ctx_stride;
  sh.progress = progress + slot * VP8B_MAX_PARTS;
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  TokLane L;
  tk_lane_init(L, smem_s + lay.rings + (uint32_t)j * TK_RING_BYTES, smem_s + lay.ctl + (uint32_t)j * 8u, im.in_off, h, part,
               mbi, im.mb_w, rows);
  while (L.phase != 2) {
    if (L.phase == 0) tk_mb_start(L, sh, im, rows, P, mbi);
    if (L.phase == 1) tk_step(L, sh, im, P, mbi, cf);
  }
  if (L.status != VP8B_OK) h->status = L.status;
}

// ---------------------------------------------------------------------------------------------------------
// Lockstep token parser (vp8_tokens_lockstep.h). A block owns ipb images x P partitions = S streams; stream j
// sits in warp j % cw, lane j / cw (so the partitions of one image sit in different warps), lanes >= lpw idle.
// With one stream per lane and ~7 lanes per warp, ONE warp per SM sub-partition carries the whole of BASELINE
// config 2 (4096 streams over 592 sub-partitions) and the loop runs at its dependency latency, not at the issue
// rate shared with six other warps.
struct TlLayout {   // byte offsets inside the block's dynamic shared memory
  uint32_t images, progress, ctx, total;
};
__host__ __device__ static inline TlLayout tl_layout(int P, int ipb, int ctx_stride) {
  TlLayout t;
  t.images = (TL_TAB_BYTES + 15u) & ~15u;
  t.progress = t.images + (uint32_t)ipb * TL_IMG_STRIDE;
  t.ctx = t.progress + (uint32_t)ipb * VP8B_MAX_PARTS * 4u;
  t.total = t.ctx + (uint32_t)ipb * (uint32_t)(P + 1) * (uint32_t)ctx_stride * 2u;
  return t;
}

__global__ void __launch_bounds__(32 * 16, 1) k_parse_tokens_lockstep(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                                  FrameHdr* hdrs, uint32_t* mbinfo, int16_t* coeffs,
                                                                  const int* __restrict__ ids, int count, int P, int ipb, int lpw,
                                                                  int cw, int ctx_stride, int grouped, int row_begin, int row_end,
                                                                  TokResume* resume, uint16_t* resume_ctx) {
  extern __shared__ __align__(16) uint8_t smem[];
  const TlLayout lay = tl_layout(P, ipb, ctx_stride);
  TlTables* tables = reinterpret_cast<TlTables*>(smem);
  int* progress = reinterpret_cast<int*>(smem + lay.progress);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  tl_tables_fill(tables, tid, nthreads);
  for (int k = tid; k < ipb * VP8B_MAX_PARTS; k += nthreads) progress[k] = 0;
  for (int slot = 0; slot < ipb; ++slot) {
    const int g = blockIdx.x * ipb + slot;
    if (g < count) tl_image_fill(smem + lay.images + (size_t)slot * TL_IMG_STRIDE, &hdrs[ids[g]], tid, nthreads);
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int j = lane * cw + warp;           // stream inside the block
  const int slot = j / P, part = j % P;
  const int g = blockIdx.x * ipb + slot;
  int have = lane < lpw && j < ipb * P && g < count;
  const int img = have ? ids[g] : 0;
  FrameHdr* h = &hdrs[img];
  if (have && !(h->status == VP8B_OK && h->num_parts == P)) {   // header failed (or, never expected, the host pre-scan disagreed)
    if (part == 0 && h->status == VP8B_OK) h->status = VP8B_BITSTREAM_ERROR;
    have = 0;
  }
  const ImgDesc im = imgs[img];
  TlCtx c;
  c.img_s = tk_saddr_of(smem + lay.images + (size_t)(have ? slot : 0) * TL_IMG_STRIDE);
  c.tab_s = tk_saddr_of(tables);
  asm volatile("" : "+r"(c.img_s), "+r"(c.tab_s));   // keep both as plain registers (no per-iteration cvta)
  c.topctx = reinterpret_cast<uint16_t*>(smem + lay.ctx) + (size_t)(have ? slot : 0) * (P + 1) * ctx_stride;
  c.progress = progress + (have ? slot : 0) * VP8B_MAX_PARTS;
  c.mbinfo = mbinfo + 4 * (size_t)im.mb_base;
  c.coeffs = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  // Row bands (single-partition images only): this launch parses macroblock rows [row_begin, row_end) and leaves the
  // reader and the top contexts of each stream in `resume` / `resume_ctx` for the launch that takes the next band.
  const int rows_total = have ? h->rows : 0;
  c.mb_w = im.mb_w; c.rows = rows_total < row_end ? rows_total : row_end; c.P = P; c.part = part; c.use_skip = h->use_skip;
  c.ctx_stride = ctx_stride;
  if (part >= c.rows || row_begin >= c.rows) have = 0;
  // per-lane copies of what the rare paths need: left as kernel parameters they are re-read from the constant bank
  // in every iteration (the loads get hoisted above the branches that need them)
  asm volatile("" : "+r"(c.P), "+r"(c.ctx_stride), "+r"(c.mb_w), "+l"(c.mbinfo), "+l"(c.coeffs));
  __builtin_assume(__isGlobal(c.mbinfo));
  __builtin_assume(__isGlobal(c.coeffs));
  TlLane L;
  if (have) {
    tl_lane_init(L, c, arena + im.in_off, h);
    if (row_begin > 0) {
      const TokResume rs = resume[img];
      L.d.wp = L.d.wbase + rs.wp_off; L.d.V = rs.V; L.d.vlo = rs.vlo; L.d.nxt = rs.nxt; L.d.R24 = rs.R24;
      L.d.nbits = rs.nbits; L.d.last_shift = rs.last_shift;
      L.my = row_begin;
      L.w_next = __ldg(c.mbinfo + 4 * ((size_t)row_begin * c.mb_w) + 3);
      uint16_t* ring = c.topctx + (size_t)((row_begin - 1) & 1) * ctx_stride;
      const uint16_t* saved = resume_ctx + (size_t)img * ctx_stride;
      for (int x = 0; x < c.mb_w; ++x) ring[x] = saved[x];
    }
  } else {
    tl_lane_idle(L, c, arena);
  }
  // The counters of the loops below start from a per-thread value (always 0: sink is an XOR of bytes), for a
  // warp-uniform one makes the compiler fence the loop body with WARPSYNC.ALL; the vote costs as much as half a step
  // and is taken every 32 steps.
  const int r0 = (int)(L.sink >> 31);
  if (grouped == 2) {
    // Straight-line groups (vp8_tokens_lockstep.h:tl_group_flat): lanes that are not running execute a stand-in.
    if (have) tl_set_aside(L);   // "needs a macroblock": the first event point takes the reader back
    if (P > 1) {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < 8; ++r) tl_group_flat<1>(L, c); }
    } else {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < 8; ++r) tl_group_flat<0>(L, c); }
    }
  } else if (grouped) {
    // Groups of four decodes per lane between event points (vp8_tokens_lockstep.h:tl_group). A lane starts as "needs a
    // macroblock", which the first event point resolves.
    if (P > 1) {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < 8; ++r) tl_group<1>(L, c); }
    } else {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < 8; ++r) tl_group<0>(L, c); }
    }
  } else {
    // One decode per lane per step with the block ends handled on the spot; the whole warp (parked lanes included)
    // meets again at the end of every step. (Left to itself the compiler turns "decode until the block ends" into an
    // inner loop, and a lane whose block has ended would wait at that loop's exit for the longest block in the warp.)
    if (P > 1) {
      while (__any_sync(0xffffffffu, L.alive)) {
        for (int r = r0; r < 8; ++r) {
          bd_fill_lookahead(L.d);
          tl_step_inline<1>(L, c); tl_step_inline<1>(L, c); tl_step_inline<1>(L, c); tl_step_inline<1>(L, c);
        }
      }
    } else {
      if (have && !tl_mb_next<0>(L, c)) tl_lane_park(L, c);
      while (__any_sync(0xffffffffu, L.alive)) {
        for (int r = r0; r < 8; ++r) {
          bd_fill_lookahead(L.d);
          tl_step_inline<0>(L, c); tl_step_inline<0>(L, c); tl_step_inline<0>(L, c); tl_step_inline<0>(L, c);
        }
      }
    }
  }
  if (have && L.status != VP8B_OK) h->status = L.status;
  if (have && L.status == VP8B_OK && row_end < rows_total) {   // more bands to come
    TokResume rs;
    const BoolDec& d = L.parked;
    rs.wp_off = (int32_t)(d.wp - d.wbase); rs.V = d.V; rs.vlo = d.vlo; rs.nxt = d.nxt; rs.R24 = d.R24;
    rs.nbits = d.nbits; rs.last_shift = d.last_shift; rs.pad = 0;
    resume[img] = rs;
    const uint16_t* ring = c.topctx + (size_t)((row_end - 1) & 1) * ctx_stride;
    uint16_t* saved = resume_ctx + (size_t)img * ctx_stride;
    for (int x = 0; x < c.mb_w; ++x) saved[x] = ring[x];
  }
  if (L.sink == 0xffffffffu) h->status = VP8B_BITSTREAM_ERROR;   // never true (an XOR of bytes): keeps TlLane::sink alive
}

// ---------------------------------------------------------------------------------------------------------
// Lockstep lanes with the fp32 boolean decoder and token-stream output (vp8_tokens_fp.h): the default token parser.
// A block owns ipb images x P partitions = S streams; stream j sits in warp j % cw, lane j / cw.
struct TfLayout {   // byte offsets from the 1024-byte aligned start of the block's dynamic shared memory
  uint32_t images, rings, bars, progress, ctx, total;
};
__host__ __device__ static inline TfLayout tf_layout(int P, int ipb, int ctx_stride, int ring, int band) {
  TfLayout t;
  t.images = 1024;                                         // TfTables in front
  t.rings = t.images + (uint32_t)ipb * TF_IMG_BYTES_B(band);   // RING: per stream a 256-byte input ring (256-byte aligned)
  t.bars = t.rings + (ring ? (uint32_t)(ipb * P) * 256u : 0u);          // ... and two mbarriers
  t.progress = t.bars + (ring ? (uint32_t)(ipb * P) * 16u : 0u);
  t.ctx = t.progress + (uint32_t)ipb * VP8B_MAX_PARTS * 4u;
  t.total = t.ctx + (uint32_t)ipb * (uint32_t)(P + 1) * (uint32_t)ctx_stride * 2u + 1024u;   // + alignment slack
  return t;
}

#ifndef TF_GROUPS_PER_VOTE
#define TF_GROUPS_PER_VOTE 8   // groups of four decodes between two votes on "any lane still alive"
#endif
template <int RING, int BAND>
__global__ void __launch_bounds__(32 * 16, 1) k_parse_tokens_fp(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                            FrameHdr* hdrs, uint32_t* mbinfo, uint32_t* tokens, MbTok* mbtok,
                                                            const int* __restrict__ ids, int count, int P, int ipb, int lpw,
                                                            int cw, int ctx_stride, int flat) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tk_saddr_of(smem_raw) & 1023u)) & 1023u);   // the rows' addresses carry the position in bits 6-9
  const TfLayout lay = tf_layout(P, ipb, ctx_stride, RING, BAND);
  TfTables* tables = reinterpret_cast<TfTables*>(smem);
  int* progress = reinterpret_cast<int*>(smem + lay.progress);
  const int tid = threadIdx.x, nthreads = blockDim.x;
  tf_tables_fill(tables, tid, nthreads);
  for (int k = tid; k < ipb * VP8B_MAX_PARTS; k += nthreads) progress[k] = 0;
  for (int slot = 0; slot < ipb; ++slot) {
    const int g = blockIdx.x * ipb + slot;
    if (g < count) tf_image_fill(smem + lay.images + (size_t)slot * TF_IMG_BYTES_B(BAND), &hdrs[ids[g]], tid, nthreads, BAND);
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int j = lane * cw + warp;           // stream inside the block
  const int slot = j / P, part = j % P;
  const int g = blockIdx.x * ipb + slot;
  int have = lane < lpw && j < ipb * P && g < count;
  const int img = have ? ids[g] : 0;
  FrameHdr* h = &hdrs[img];
  if (have && !(h->status == VP8B_OK && h->num_parts == P)) {   // header failed (or, never expected, the host pre-scan disagreed)
    if (part == 0 && h->status == VP8B_OK) h->status = VP8B_BITSTREAM_ERROR;
    have = 0;
  }
  const ImgDesc im = imgs[img];
  TfCtx c;
  c.img_s = tk_saddr_of(smem + lay.images + (size_t)(have ? slot : 0) * TF_IMG_BYTES_B(BAND));
  c.tab_s = tk_saddr_of(tables);
  c.k.mant_mask = 0x007fffffu; c.k.exp46 = TF_EXP46;
  asm volatile("" : "+r"(c.img_s), "+r"(c.tab_s), "+r"(c.k.mant_mask), "+r"(c.k.exp46));   // plain registers: no per-step cvta, one LOP3 in fd_bit
  c.topctx = reinterpret_cast<uint16_t*>(smem + lay.ctx) + (size_t)(have ? slot : 0) * (P + 1) * ctx_stride;
  c.progress = progress + (have ? slot : 0) * VP8B_MAX_PARTS;
  c.mbinfo = mbinfo + 4 * (size_t)im.mb_base;
  c.mbtok = mbtok + (size_t)im.mb_base;
  c.tokens = tokens + (size_t)im.mb_base * TF_TOKENS_PER_MB;
  c.mb_w = im.mb_w; c.rows = have ? h->rows : 0; c.P = P; c.part = part; c.use_skip = h->use_skip;
  c.ctx_stride = ctx_stride;
  if (part >= c.rows) have = 0;
  asm volatile("" : "+r"(c.P), "+r"(c.ctx_stride), "+r"(c.mb_w), "+l"(c.mbinfo), "+l"(c.tokens), "+l"(c.mbtok));
  __builtin_assume(__isGlobal(c.mbinfo));
  __builtin_assume(__isGlobal(c.tokens));
  __builtin_assume(__isGlobal(c.mbtok));
  TfLaneT<BAND> L;
  if (have) {
    tf_lane_init(L, c, arena + im.in_off, h);
    if (RING) fd_ring_open(L.d, tk_saddr_of(smem + lay.rings) + (uint32_t)j * 256u, tk_saddr_of(smem + lay.bars) + (uint32_t)j * 16u);
  } else {
    tf_lane_idle(L, c, arena);
  }
  // The loop counter starts from a per-thread value (always 0) so that the compiler does not fence the body with
  // WARPSYNC.ALL; the vote costs as much as half a step and is taken every 32 steps.
  const int r0 = (int)(L.sink >> 31);
  if (flat) {
    // Straight-line groups of four decodes with one event point (vp8_tokens_fp.h:tf_group_flat): the default.
    if (P > 1) {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < TF_GROUPS_PER_VOTE; ++r) tf_group_flat<1, RING>(L, c); }
    } else {
      while (__any_sync(0xffffffffu, L.alive)) { for (int r = r0; r < TF_GROUPS_PER_VOTE; ++r) tf_group_flat<0, RING>(L, c); }
    }
  } else if (P > 1) {
    while (__any_sync(0xffffffffu, L.alive)) {
      for (int r = r0; r < TF_GROUPS_PER_VOTE; ++r) tf_group_inline<1, RING>(L, c);
    }
  } else {
    if (have && !tf_mb_next<0>(L, c)) tf_lane_park(L, c);
    while (__any_sync(0xffffffffu, L.alive)) {
      for (int r = r0; r < TF_GROUPS_PER_VOTE; ++r) tf_group_inline<0, RING>(L, c);
    }
  }
  if (have && L.status != VP8B_OK) h->status = L.status;
  if (L.sink == 0xffffffffu) h->status = VP8B_BITSTREAM_ERROR;   // never true (an XOR of bytes): keeps TfLane::sink alive
}

// ---------------------------------------------------------------------------------------------------------
// Images flagged VP8B_FLAG_LITERAL_READER, after K1 and the token parser have been over them: parsed once more, header to last
// token, by lane 0 of a warp with the reference's reader (vp8_literal.h), overwriting what those kernels left. One block of
// one warp per image of the wave; launched only when the wave holds such an image.
__global__ void __launch_bounds__(32) k_parse_literal(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs, FrameHdr* hdrs,
                                                      uint32_t* mbinfo, uint32_t* tokens, MbTok* mbtok, int first) {
  extern __shared__ __align__(16) uint8_t lit_scratch[];
  __shared__ uint8_t bprob[900];
  const int img = first + blockIdx.x;
  const ImgDesc im = imgs[img];
  if (!(im.flags & VP8B_FLAG_LITERAL_READER) || (im.flags & VP8B_FLAG_LOSSLESS)) return;
  for (int k = threadIdx.x; k < 900; k += 32) bprob[k] = kVp8BModeProba[k];
  for (int k = threadIdx.x; k < VP8B_COEFFS_PER_MB; k += 32) ((int16_t*)(lit_scratch + LIT_LEVELS))[k] = 0;
  __syncwarp();
  if (threadIdx.x == 0) {
    parse_image_literal(arena + im.in_off, im, &hdrs[img], bprob, mbinfo + 4 * (size_t)im.mb_base,
                        tokens + (size_t)im.mb_base * TF_TOKENS_PER_MB, mbtok + (size_t)im.mb_base, lit_scratch);
  }
}

// ---------------------------------------------------------------------------------------------------------
// RECON_WARPS warps per image: 8, 4 for small pictures (pixel_warps_for); 16 is compiled for WEBP_B200_RECON_WARPS, an A/B
// switch. ROWS = 1 (default): one warp per macroblock row; ROWS = 0: all warps on one anti-diagonal at a time.
template <int RECON_WARPS, int ROWS>
__global__ void __launch_bounds__(32 * RECON_WARPS, 32 / RECON_WARPS) k_reconstruct(const ImgDesc* __restrict__ imgs, FrameHdr* hdrs,
                                                                  uint32_t* mbinfo, const int16_t* __restrict__ coeffs,
                                                                  uint8_t* yuv, int first, int row_begin, int row_end,
                                                                  uint8_t* band_ctx, int band_ctx_stride,
                                                                  const uint32_t* __restrict__ tokens, const MbTok* __restrict__ mbtok) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int img = first + blockIdx.x;
  if (vp8b_frame_lost(&hdrs[img])) {
    // Nothing to reconstruct; but a frame whose tokens ran out and which carries an ALPH chunk needs the row at which they did
    // (FrameHdr::fail_row), and this is the last launch that still finds the wave's MbTok entries in place. tf_mb_store marked the
    // macroblock where each partition stopped; stale marks can only sit below or to the right of a fresh one, so the smallest
    // marked row is the answer. (Keeping the row inside the parser -- a field of TfLane, a pointer in TfCtx, two spare bytes
    // of the shared-memory rows, or (mx, my) read back after the loops -- cost its decode loop 2-4 %: 264.7 -> 272-275 ms per
    // 4096 full-HD images, profiles/r02p_variants.log.)
    FrameHdr* h = &hdrs[img];
    if (tokens != nullptr && row_begin == 0 && h->status != VP8B_OK && h->fail_row > 0 && imgs[img].alpha_size != 0 &&
        !(imgs[img].flags & VP8B_FLAG_LITERAL_READER)) {   // (k_parse_literal leaves the row itself)   // K1's own failures leave fail_row <= 0
      __shared__ int found;
      if (threadIdx.x == 0) found = VP8B_FAIL_NONE;
      __syncthreads();
      const MbTok* mt = mbtok + (size_t)imgs[img].mb_base;
      const int w = imgs[img].mb_w, n = w * h->rows;
      for (int i = threadIdx.x; i < n; i += blockDim.x) if (mt[i].count == TF_MBTOK_FAILED) atomicMin(&found, i / w);
      __syncthreads();
      if (threadIdx.x == 0 && found < h->fail_row) h->fail_row = found;
    }
    return;
  }
  const ImgDesc im = imgs[img];
  const int mb_w = im.mb_w, mb_h = hdrs[img].rows;
  // macroblock rows [r0, r1) of this launch (row bands: the top-neighbour pixels of row r0 come from band_ctx, where the
  // launch that reconstructed row r0 - 1 left them; they must be the UNFILTERED ones, and the planes get filtered in between)
  const int r0 = row_begin, r1 = mb_h < row_end ? mb_h : row_end;
  if (r0 >= r1) return;
  const int warp = threadIdx.x >> 5;
  ReconWs& ws = *reinterpret_cast<ReconWs*>(smem + sizeof(ReconWs) * warp);
  ReconCtx cx;
  recon_ctx_bind(cx, smem + sizeof(ReconWs) * RECON_WARPS, mb_w, mb_h);
  __shared__ int16_t dqs[24];   // the frame's dequantisers, [segment][y1 dc/ac, y2 dc/ac, uv dc/ac]
  if (threadIdx.x < 24) dqs[threadIdx.x] = (&hdrs[img].dq[0][0])[threadIdx.x];
  __shared__ uint32_t pred4[160];   // every lane of a sub-block reads a different word: shared memory, not the constant bank
  for (int k = threadIdx.x; k < 160; k += blockDim.x) pred4[k] = (&kPred4x[0][0])[k];
  cx.pred4 = pred4;
  uint8_t* saved = band_ctx + (size_t)img * band_ctx_stride;
  if (r0 > 0) for (int k = threadIdx.x; k < 8 * mb_w; k += blockDim.x) ((uint32_t*)cx.top)[k] = ((const uint32_t*)saved)[k];
  __syncthreads();
  const size_t nmb = (size_t)mb_w * im.mb_h;
  uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  uint8_t* up = yp + nmb * 256;
  uint8_t* vp = up + nmb * 64;
  uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  const int16_t* cf = coeffs + (size_t)im.mb_base * VP8B_COEFFS_PER_MB;
  const uint32_t* tk = tokens != nullptr ? tokens + (size_t)im.mb_base * TF_TOKENS_PER_MB : nullptr;   // token stream instead of the dense plane
  const MbTok* mt = mbtok + (size_t)im.mb_base;
  const int nrows = r1 - r0;
  if (ROWS) {
    // One warp per macroblock row, no block-wide barrier: warp w owns rows r0 + w, r0 + w + RECON_WARPS, ... and walks each
    // from left to right; macroblock (mx, my) needs (mx + 1, my - 1) finished (its above-right pixels), whose owner
    // publishes how many macroblocks of its row are done. Every context entry has one writer and one reader that is also
    // its next writer, so nothing else orders the rows. The next macroblock's MbInfo and MbTok are fetched while the
    // current one is reconstructed.
    volatile int* row_done = reinterpret_cast<volatile int*>(smem + sizeof(ReconWs) * RECON_WARPS + ((recon_ctx_bytes(mb_w, mb_h) + 15) & ~(size_t)15));
    for (int k = threadIdx.x; k < nrows; k += blockDim.x) row_done[k] = 0;
    __syncthreads();
    for (int ly = warp; ly < nrows; ly += RECON_WARPS) {
      const int my = r0 + ly;
      const size_t row0 = (size_t)my * mb_w;
      uint4 iw = *(const uint4*)(mbi + 4 * row0);
      MbTok t; t.first = 0; t.count = 0;
      ReconTok rt;
      if (tk != nullptr) { t = mt[row0]; recon_fetch_tokens(rt, tk + t.first, t.count); }
      for (int mx = 0; mx < mb_w; ++mx) {
        const size_t idx = row0 + mx;
        const uint4 iw_cur = iw;
        const MbTok t_cur = t;
        t.count = 0;
        if (mx + 1 < mb_w) { iw = *(const uint4*)(mbi + 4 * (idx + 1)); if (tk != nullptr) t = mt[idx + 1]; }
        if (ly > 0) {
          const int need = mx + 2 < mb_w ? mx + 2 : mb_w;
          while (row_done[ly - 1] < need) __nanosleep(40);
          __threadfence_block();
        }
        const int16_t* dq6 = dqs + 6 * ((iw_cur.w >> MBW_SEG_SHIFT) & 3);
        if (tk != nullptr) recon_macroblock(ws, cx, mx, my, mb_w, iw_cur, mbi + 4 * idx, nullptr, dq6, yp, up, vp, tk + t_cur.first, t_cur.count, rt, tk + t.first, t.count);
        else recon_macroblock(ws, cx, mx, my, mb_w, iw_cur, mbi + 4 * idx, cf + idx * VP8B_COEFFS_PER_MB, dq6, yp, up, vp, nullptr, 0, rt, nullptr, 0);
        __threadfence_block();
        if ((threadIdx.x & 31) == 0) row_done[ly] = mx + 1;
      }
    }
    __syncthreads();
  } else {
  // Lag-2 anti-diagonal wavefront with a block-wide barrier per step.
  const int steps = mb_w + 2 * (nrows - 1);
  for (int d = 0; d < steps; ++d) {
    // rows with a macroblock on this anti-diagonal: mx = d - 2*(my - r0) in [0, mb_w)
    const int ly_lo = (d - mb_w + 2 > 0) ? (d - mb_w + 2) >> 1 : 0;
    const int ly_hi = (d >> 1) < nrows - 1 ? (d >> 1) : nrows - 1;
    for (int ly = ly_lo + warp; ly <= ly_hi; ly += RECON_WARPS) {
      const int mx = d - 2 * ly, my = r0 + ly;
      const size_t idx = (size_t)my * mb_w + mx;
      const int16_t* dq6 = dqs + 6 * ((mbi[4 * idx + 3] >> MBW_SEG_SHIFT) & 3);
      if (tk != nullptr) {
        const MbTok t = mt[idx];
        recon_macroblock(ws, cx, mx, my, mb_w, mbi + 4 * idx, nullptr, dq6, yp, up, vp, tk + t.first, t.count);
      } else {
        recon_macroblock(ws, cx, mx, my, mb_w, mbi + 4 * idx, cf + idx * VP8B_COEFFS_PER_MB, dq6, yp, up, vp);
      }
    }
    __syncthreads();
  }
  }
  if (r1 < mb_h) for (int k = threadIdx.x; k < 8 * mb_w; k += blockDim.x) ((uint32_t*)saved)[k] = ((const uint32_t*)cx.top)[k];
}

// ---------------------------------------------------------------------------------------------------------
// FILTER_WARPS warps per image: 8, or 4 for small pictures (a 16-row thumbnail keeps 8 warps waiting on each other's first
// macroblocks for as long as they work; see vp8k_loop_filter).
template <int FILTER_WARPS>
__global__ void __launch_bounds__(32 * FILTER_WARPS, 40 / FILTER_WARPS) k_loop_filter(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                                   const uint32_t* __restrict__ mbinfo, uint8_t* yuv, int first,
                                                                   int row_begin, int row_end, const int8_t* __restrict__ dither_plane) {
  __shared__ __align__(16) FilterWs wss[FILTER_WARPS];
  __shared__ uint8_t fstr[32];
  const int img = first + blockIdx.x;
  const FrameHdr* h = &hdrs[img];
  // options.dithering_strength: the planned offsets (dither_plan_image) go onto a macroblock's chroma once its own row
  // has finished with it, i.e. after the macroblock to its right has been filtered, and before the row below may touch it
  const int dithering = dither_plane != nullptr && (h->dither[0] | h->dither[1] | h->dither[2] | h->dither[3]) != 0;
  if (vp8b_frame_lost(h) || (h->filter_type == 0 && !dithering)) return;
  const ImgDesc im = imgs[img];
  const int mb_w = im.mb_w, mb_h = h->rows;
  const int filter_type = h->filter_type;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x < 32) fstr[threadIdx.x] = ((const uint8_t*)h->fstr)[threadIdx.x];
  __syncthreads();
  FilterWs& ws = wss[warp];
  const size_t nmb = (size_t)mb_w * im.mb_h;
  uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  uint8_t* up = yp + nmb * 256;
  uint8_t* vp = up + nmb * 64;
  const uint32_t* mbi = mbinfo + 4 * (size_t)im.mb_base;
  const int r0 = row_begin, r1 = mb_h < row_end ? mb_h : row_end;   // macroblock rows of this launch (row bands)
  const int nrows = r1 - r0;
  // Wavefront without block-wide barriers: warp w owns rows r0 + w, r0 + w + 8, ... and walks each left to right; macroblock
  // (mx, my) follows (mx - 1, my) (same warp) and (mx + 1, my - 1), whose owner publishes the count of macroblocks it has
  // finished (frame_dec.c:203-262 filters in raster order and reaches 3 pixels up and left, 4 with the ones it only reads).
  // Measured against a barrier per anti-diagonal: simple filter 27.4 -> 26.0 ms, normal filter 64 -> 50 ms per 4096 full-HD images.
  extern __shared__ __align__(16) uint8_t fsmem[];
  volatile int* row_done = reinterpret_cast<volatile int*>(fsmem);
  for (int k = threadIdx.x; k < nrows; k += blockDim.x) row_done[k] = 0;
  __syncthreads();
  const int normal = filter_type == 2;
  for (int ly = warp; ly < nrows; ly += FILTER_WARPS) {
    const int my = r0 + ly;
    // The warp runs one macroblock ahead of itself: the rows above macroblock mx + 1 must be final (its own need) before
    // its columns are fetched, which happens while macroblock mx is filtered (filter_fetch, vp8_pixel_core.h).
    FilterPre pre;
    if (ly > 0) { const int need = 2 < mb_w ? 2 : mb_w; while (row_done[ly - 1] < need) __nanosleep(20); __threadfence_block(); }
    if (filter_type != 0) filter_fetch(pre, 0, my, mb_w, normal, yp, up, vp);
    for (int mx = 0; mx < mb_w; ++mx) {
      const uint32_t w = mbi[4 * ((size_t)my * mb_w + mx) + 3];
      const uint8_t* fs = fstr + 8 * ((w >> MBW_SEG_SHIFT) & 3) + ((w & MBW_I4X4) ? 4 : 0);
      if (filter_type != 0) filter_fill(ws, pre, normal);
      if (mx + 1 < mb_w) {
        if (ly > 0) {
          const int need = mx + 3 < mb_w ? mx + 3 : mb_w;
          while (row_done[ly - 1] < need) __nanosleep(20);
          __threadfence_block();
        }
        if (filter_type != 0) filter_fetch(pre, mx + 1, my, mb_w, normal, yp, up, vp);
      }
      if (filter_type != 0) {
        if (fs[0] != 0) filter_tile(ws, mx, my, mb_w, filter_type, fs, (w & MBW_INNER) != 0, yp, up, vp);
        filter_shift(ws, normal);
      }
      if (dithering) {
        const int8_t* dp = dither_plane + (size_t)im.mb_base * 128;
        __syncwarp();
        if (mx > 0 && (mbi[4 * ((size_t)my * mb_w + mx - 1) + 3] & MBW_DITHER)) dither_macroblock(mx - 1, my, mb_w, dp, up, vp);
        if (mx == mb_w - 1 && (w & MBW_DITHER)) dither_macroblock(mx, my, mb_w, dp, up, vp);
      }
      __syncwarp();
      __threadfence_block();
      if ((threadIdx.x & 31) == 0) row_done[ly] = mx + 1;
    }
  }
}

// options.dithering_strength: one thread per image lays the random offsets down (vp8_pixel_core.h:dither_plan_image).
__global__ void __launch_bounds__(32) k_dither_plan(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                    uint32_t* mbinfo, int8_t* dither_plane, int first, int count) {
  __shared__ uint32_t tabs[32][55];
  const int k = blockIdx.x * 32 + threadIdx.x;
  if (k >= count) return;
  const int img = first + k;
  if (vp8b_frame_lost(&hdrs[img])) return;
  const ImgDesc im = imgs[img];
  dither_plan_image(im, &hdrs[img], mbinfo + 4 * (size_t)im.mb_base, dither_plane + (size_t)im.mb_base * 128, tabs[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------
#define EMIT_THREADS 256

__global__ void __launch_bounds__(EMIT_THREADS, 5) k_emit(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                       const uint8_t* __restrict__ yuv, const uint8_t* __restrict__ alpha_arena,
                                                       uint8_t* out, int first, int blocks_per_image, int pair_begin, int pair_end) {
  const int img = first + blockIdx.x / blocks_per_image;
  const int chunk = blockIdx.x % blocks_per_image;
  if (vp8b_frame_lost(&hdrs[img])) return;
  const ImgDesc im = imgs[img];
  if (im.dst_w != 0) return;   // options.use_scaling: k_emit_scaled
  const size_t nmb = (size_t)im.mb_w * im.mb_h;
  // planes and alpha re-based at the output window (crop_x, crop_y are even)
  const uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  const uint8_t* up = yp + nmb * 256 + (size_t)(im.crop_y >> 1) * (8 * im.mb_w) + (im.crop_x >> 1);
  const uint8_t* vp = yp + nmb * 256 + nmb * 64 + (size_t)(im.crop_y >> 1) * (8 * im.mb_w) + (im.crop_x >> 1);
  yp += (size_t)im.crop_y * (16 * im.mb_w) + im.crop_x;
  uint8_t* o = out + im.out_off;
  const uint8_t* alpha = (im.alpha_plane != VP8B_NO_ALPHA) ? alpha_arena + im.alpha_plane + (size_t)im.crop_y * im.width + im.crop_x : nullptr;
  const int t = chunk * EMIT_THREADS + threadIdx.x;
  const int w = im.out_w, h = im.out_h;
  if (im.csp == 11 || im.csp == 12) {   // MODE_YUV / MODE_YUVA: 16-byte chunks of Y rows, then U, V (and alpha) rows
    const int uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
    const int qy = (w + 15) >> 4, quv = (uvw + 15) >> 4;
    const int ny = qy * h, nuv = quv * uvh;
    if (t < ny) emit_yuv_chunk(im, yp, up, vp, alpha, o, 0, t % qy, t / qy);
    else if (t < ny + nuv) emit_yuv_chunk(im, yp, up, vp, alpha, o, 1, (t - ny) % quv, (t - ny) / quv);
    else if (t < ny + 2 * nuv) emit_yuv_chunk(im, yp, up, vp, alpha, o, 2, (t - ny - nuv) % quv, (t - ny - nuv) / quv);
    else if (im.csp == 12 && t < 2 * ny + 2 * nuv) emit_yuv_chunk(im, yp, up, vp, alpha, o, 3, (t - ny - 2 * nuv) % qy, (t - ny - 2 * nuv) / qy);
  } else if (emit_uses_pairs(im.csp, im.flags, im.crop_x)) {   // 8 pixels x 2 rows per thread
    const int qw = (w + 7) >> 3;   // row pairs [pair_begin, pair_end) of this launch (row bands; everything otherwise)
    const int pairs = (h >> 1) + 1, pt = t / qw + pair_begin;
    if (pt < pairs && pt < pair_end) emit_rgba_pair8(im, yp, up, vp, alpha, o, t % qw, pt);
  } else {
    const int qw = (w + 3) >> 2;
    if (t < qw * h) emit_rgb_quad(im, yp, up, vp, alpha, o, t % qw, t / qw);
  }
}

// options.use_scaling: one thread per output column of a plane (vp8_pixel_core.h:emit_scaled_column).
__global__ void __launch_bounds__(EMIT_THREADS) k_emit_scaled(const ImgDesc* __restrict__ imgs, const FrameHdr* __restrict__ hdrs,
                                                              const uint8_t* __restrict__ yuv, const uint8_t* __restrict__ alpha_arena,
                                                              uint8_t* out, int first, int blocks_per_image) {
  const int img = first + blockIdx.x / blocks_per_image;
  const int chunk = blockIdx.x % blocks_per_image;
  if (vp8b_frame_lost(&hdrs[img])) return;
  const ImgDesc im = imgs[img];
  if (im.dst_w == 0) return;
  const size_t nmb = (size_t)im.mb_w * im.mb_h;
  const uint8_t* yp = yuv + (size_t)im.mb_base * 384;
  const uint8_t* up = yp + nmb * 256 + (size_t)(im.crop_y >> 1) * (8 * im.mb_w) + (im.crop_x >> 1);
  const uint8_t* vp = yp + nmb * 256 + nmb * 64 + (size_t)(im.crop_y >> 1) * (8 * im.mb_w) + (im.crop_x >> 1);
  yp += (size_t)im.crop_y * (16 * im.mb_w) + im.crop_x;
  const uint8_t* alpha = (im.alpha_plane != VP8B_NO_ALPHA) ? alpha_arena + im.alpha_plane + (size_t)im.crop_y * im.width + im.crop_x : nullptr;
  emit_scaled_column(im, yp, up, vp, alpha, out + im.out_off, chunk * EMIT_THREADS + threadIdx.x);
}

// Per-image status words -> page-locked host memory, written by the device itself. (A cudaMemcpyAsync here would queue
// behind the pixel downloads on the device-to-host copy engine and hold the COMPUTE stream up until they are through:
// that is what kept batch k+1 from starting under batch k's download.)
// Three words per image: status, fail_row, rows << 8 | filter_type (what vp8b_vp8_failure_first needs on the host).
__global__ void __launch_bounds__(256) k_collect_status(const FrameHdr* __restrict__ hdrs, int* host_statuses, int count) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) {
    host_statuses[3 * k] = vp8b_frame_status(&hdrs[k]);
    host_statuses[3 * k + 1] = hdrs[k].fail_row;
    host_statuses[3 * k + 2] = (hdrs[k].all_rows << 8) | hdrs[k].filter_type;
  }
}

// Small device -> mapped host copies of the same kind (the ALPH headers the host sizes its work areas from).
__global__ void __launch_bounds__(256) k_copy_to_host(const uint8_t* __restrict__ src, uint8_t* host_dst, size_t bytes) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < bytes; k += (size_t)gridDim.x * blockDim.x) host_dst[k] = src[k];
}

// =========================================================================================================
// Launchers (plain C interface for vp8_batch.cu).
static size_t recon_smem_bytes(int warps, int max_mb_w, int max_mb_h) {
  return sizeof(ReconWs) * warps + ((recon_ctx_bytes(max_mb_w, max_mb_h) + 15) & ~(size_t)15) + 4 * (size_t)max_mb_h + 16;   // + row_done
}

static size_t tokens_slot_bytes(int P, int max_mb_w) {
  return (TOKW_CTX + (size_t)(P + 1) * max_mb_w * 2 + 15) & ~(size_t)15;
}

#define VP8K_MAX_DYN_SMEM (227 * 1024)   // opt-in ceiling per block on sm_100
extern "C" cudaError_t vp8k_init_device(void) {
  const void* kernels[] = { (const void*)k_parse_modes, (const void*)k_parse_modes_lockstep, (const void*)k_parse_tokens, (const void*)k_parse_tokens_fsm,
                            (const void*)k_parse_tokens_lockstep, (const void*)k_parse_tokens_fp<0, 0>, (const void*)k_parse_tokens_fp<1, 0>,
                            (const void*)k_parse_tokens_fp<0, 1>,
                            (const void*)k_reconstruct<4, 0>, (const void*)k_reconstruct<8, 0>, (const void*)k_reconstruct<16, 0>,
                            (const void*)k_reconstruct<2, 1>, (const void*)k_reconstruct<4, 1>, (const void*)k_reconstruct<8, 1>, (const void*)k_reconstruct<16, 1>,
                            (const void*)k_loop_filter<8>, (const void*)k_loop_filter<4>, (const void*)k_loop_filter<2> };
  for (const void* k : kernels) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, k);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, VP8K_MAX_DYN_SMEM - (int)fa.sharedSizeBytes);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

static int pack_per_block(int count, int max_per_block) {   // as many as fit, but no fewer than 148 blocks where the launch allows
  int ipb = max_per_block;
  while (ipb > 1 && (count + ipb - 2) / (ipb - 1) <= 148) --ipb;
  return ipb;
}

// Images per warp of the mode parse. The kernel is issue-bound (one dependent chain per image, ~27 instructions per decode),
// so with enough images to keep ~3.5 warps on every SM sub-partition anyway, several images share one warp's instruction
// stream as SIMT lanes: they diverge where their syntax differs (an i4x4 macroblock beside an i16 one) and meet again at
// every macroblock end. Measured (profiles/r01t_modes_lanes.log): 65536 thumbnails 25.6 -> 3.9 ms at 32 lanes, 16384: 6.4 ->
// 2.5 ms at 8, 8192: 3.2 -> 2.0 ms at 4; 4096 full-HD images are best left at one image per warp (30.3 ms; 31.3 at 2 lanes,
// 43.7 at 7: too few warps left to hide the chain's latency). WEBP_B200_MODES_LANES forces a value.
static int modes_lanes_for(int count) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("WEBP_B200_MODES_LANES"); forced = (e != NULL && atoi(e) >= 1 && atoi(e) <= 32) ? atoi(e) : 0; }
  if (forced > 0) return forced;
  if (count < 6144) return 1;
  int want = (count + 1036) / 2072, lanes = 1;   // 2072 = 148 SMs x 4 sub-partitions x 3.5 warps
  while (lanes * 2 <= want && lanes < 32) lanes *= 2;
  return lanes;
}

extern "C" void vp8k_parse_modes(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                 int first, int count, int max_mb_w) {
  // Mapping: one image per warp (k_parse_modes) unless WEBP_B200_MODES=lockstep asks for the lockstep lanes
  // (k_parse_modes_lockstep, vp8_modes_lockstep.h). Measured on the B200 (profiles/r03c, r03d): the table-driven machine pays
  // two dependent shared-memory loads and a divergent leaf path per decode where the straight-line code pays neither --
  // 4096 full-HD images 54.4 ms at 7 lanes per warp, 44.5 at 2, 62.6 at 1, against 30.2 ms here; 65536 thumbnails 5.8 against
  // 4.1 ms. Kept as a second, parity-tested instantiation (GPU test_every_mode_mapping, emulation variants +256), not the default.
  static int mapping = -1;   // process-wide A/B switch, read once
  if (mapping < 0) { const char* e = getenv("WEBP_B200_MODES"); mapping = (e != NULL && !strcmp(e, "lockstep")) ? 2 : 1; }
  if (mapping != 1) {
    int ll = (count + 148 * MODESL_WARPS - 1) / (148 * MODESL_WARPS);   // one warp per sub-partition, as many lanes as that takes
    if (ll > 32) ll = 32;
    static int forced_ll = -1;
    if (forced_ll < 0) { const char* e = getenv("WEBP_B200_MODES_LANES"); forced_ll = (e != NULL && atoi(e) >= 1 && atoi(e) <= 32) ? atoi(e) : 0; }
    if (forced_ll > 0) ll = forced_ll;
    const size_t per_image = (size_t)max_mb_w * 4 + ML_ROW_BYTES;
    while (ll > 1 && (size_t)MODESL_WARPS * ll * per_image > 190u * 1024u) --ll;
    const size_t msm = MODESL_TAB_BYTES + MODESL_BPROB_BYTES + (size_t)MODESL_WARPS * ll * per_image;
    if (msm <= 200u * 1024u && (mapping == 2 || ll >= 2)) {
      const int ipb = MODESL_WARPS * ll;
      k_parse_modes_lockstep<<<(count + ipb - 1) / ipb, 32 * MODESL_WARPS, msm, s>>>(arena, imgs, hdrs, mbinfo, first, count, max_mb_w, ll);
      return;
    }
  }
  int lanes = modes_lanes_for(count);
  if ((size_t)lanes * max_mb_w * 4 > 200u * 1024u) lanes = 1;   // the lanes' top-mode rows share the block's shared memory
  int ipb;
  if (lanes == 1) {
    ipb = MODES_WARPS;
    while (ipb > 1 && (size_t)ipb * max_mb_w * 4 > 200u * 1024u) --ipb;
    ipb = pack_per_block(count, ipb);
  } else {
    ipb = 4 * lanes;   // one warp per sub-partition and block; many blocks per SM
    while (ipb > lanes && (size_t)ipb * max_mb_w * 4 > 200u * 1024u) ipb -= lanes;
  }
  const size_t smem = (size_t)ipb * max_mb_w * 4;
  const int warps = (ipb + lanes - 1) / lanes;
  k_parse_modes<<<(count + ipb - 1) / ipb, 32 * warps, smem, s>>>(arena, imgs, hdrs, mbinfo, first, count, ipb, max_mb_w, lanes);
}

static int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

// Launch geometry of the lane-parallel parser for `count` images of P partitions: cw parsing warps per block
// (one per SM sub-partition by default), lpw streams per warp chosen so that the whole launch is resident at
// once where it can be (148 SMs), and as many images per block as its streams hold.
static void launch_tokens_fsm(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                              int16_t* coeffs, const int* ids, int count, int P, int max_mb_w) {
  static int f_lpw = -1, f_cw = -1;
  if (f_lpw < 0) { f_lpw = env_int("WEBP_B200_TOKEN_LPW"); f_cw = env_int("WEBP_B200_TOKEN_CW"); }
  int cw = (f_cw >= 1 && f_cw <= 8) ? f_cw : 4;
  const long streams = (long)count * P;
  int lpw = (int)((streams + 148L * cw - 1) / (148L * cw));
  if (f_lpw >= 1 && f_lpw <= 32) lpw = f_lpw;
  if (lpw < 1) lpw = 1;
  if (lpw > 32) lpw = 32;
  while (cw * lpw < P) ++lpw;               // a block holds at least one image
  int ipb = (cw * lpw) / P;                  // images per block
  while (ipb > 1 && tok_layout(P, ipb, max_mb_w).total > 200u * 1024u) --ipb;
  const TokLayout lay = tok_layout(P, ipb, max_mb_w);
  const int blocks = (count + ipb - 1) / ipb;
  k_parse_tokens_fsm<<<blocks, 32 * (cw + 1), lay.total, s>>>(arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, ipb, lpw, cw, max_mb_w);
}

// Launch geometry of the lockstep parser: cw warps per block (one per SM sub-partition), lpw streams per warp so
// that one block per SM holds the whole launch where shared memory allows it.
static void launch_tokens_lockstep(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                   int16_t* coeffs, const int* ids, int count, int P, int max_mb_w, int row_begin, int row_end,
                                   TokResume* resume, uint16_t* resume_ctx) {
  static int f_lpw = -1, f_cw = -1;
  if (f_lpw < 0) { f_lpw = env_int("WEBP_B200_TOKEN_LPW"); f_cw = env_int("WEBP_B200_TOKEN_CW"); }
  const long streams = (long)count * P;
  // one warp per SM sub-partition while the streams fit seven to a warp, then more warps (shared memory caps a block at
  // ~43 images' probability rows, so many small images want their lanes spread over more warps)
  int cw = (f_cw >= 1 && f_cw <= 16) ? f_cw : (streams <= 148L * 4 * 8 ? 4 : 8);
  int lpw = (int)((streams + 148L * cw - 1) / (148L * cw));
  if (f_lpw >= 1 && f_lpw <= 32) lpw = f_lpw;
  if (lpw < 1) lpw = 1;
  if (lpw > 32) lpw = 32;
  while (cw * lpw < P) ++lpw;               // a block holds at least one image
  int ipb = (cw * lpw) / P;                  // images per block
  while (ipb > 1 && tl_layout(P, ipb, max_mb_w).total > 200u * 1024u) --ipb;
  const TlLayout lay = tl_layout(P, ipb, max_mb_w);
  const int blocks = (count + ipb - 1) / ipb;
  // How the lanes are run (vp8_tokens_lockstep.h): 0 = block ends on the spot, while a warp has few lanes; 2 = groups of four
  // straight-line steps with one event point, when it has many (the event point's cost is shared by all the lanes that have
  // a block end pending); 1 = the same with branches around the steps of lanes that are not running. Measured per 4096
  // full-HD images, styles 0 / 1 / 2: 1 partition (7 lanes per warp) 300 / 385 / 303 ms, 8 partitions (28 lanes) 160 / 91 /
  // 83 ms; 65536 thumbnails (shared memory holds 43 images per block: 5 lanes) 170 / 201 / 205 ms.
  static int f_grouped = -2;
  if (f_grouped == -2) { const char* e = getenv("WEBP_B200_TOKEN_GROUPED"); f_grouped = e ? atoi(e) : -1; }
  const int lanes_per_warp = (ipb * P + cw - 1) / cw;
  const int grouped = f_grouped >= 0 ? f_grouped : (lanes_per_warp >= 16 ? 2 : 0);
  k_parse_tokens_lockstep<<<blocks, 32 * cw, lay.total, s>>>(arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, ipb, lpw, cw, max_mb_w, grouped,
                                                             row_begin, row_end, resume, resume_ctx);
}

// Launch geometry of the fp parser: cw warps per block (one per SM sub-partition while the streams fit eight to a warp, then
// more), lpw streams per warp so that one block per SM holds the whole launch where shared memory allows it.
static void launch_tokens_fp(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                             uint32_t* tokens, MbTok* mbtok, const int* ids, int count, int P, int max_mb_w) {
  const int f_lpw = env_int("WEBP_B200_TOKEN_LPW"), f_cw = env_int("WEBP_B200_TOKEN_CW");
  const long streams = (long)count * P;
  int cw = (f_cw >= 1 && f_cw <= 16) ? f_cw : (streams <= 148L * 4 * 8 ? 4 : 8);
  int lpw = (int)((streams + 148L * cw - 1) / (148L * cw));
  if (f_lpw >= 1 && f_lpw <= 32) lpw = f_lpw;
  if (lpw < 1) lpw = 1;
  if (lpw > 32) lpw = 32;
  while (cw * lpw < P) ++lpw;               // a block holds at least one image
  int ipb = (cw * lpw) / P;                  // images per block
  const char* er = getenv("WEBP_B200_TOKEN_RING");
  const int ring = (er != NULL && atoi(er) != 0) ? 1 : 0;
  const int want = ipb;
  while (ipb > 1 && tf_layout(P, ipb, max_mb_w, ring, 0).total > (uint32_t)VP8K_MAX_DYN_SMEM - 1024u) --ipb;
  // Shared memory seats fewer images than there are lanes for them (tens of thousands of small images): the banded layout of the
  // probability rows (2 KB per image instead of 4, vp8_tokens_fp.h:TF_TYPE_BYTES_B) seats twice as many, in half as many warps
  // (one per sub-partition, up to 32 lanes each: the instruction stream is shared by all of a warp's lanes).
  // WEBP_B200_TOKEN_BAND=0|1 forces it. Measured on 65536 256x256 thumbnails: profiles/r02u.
  const char* eb = getenv("WEBP_B200_TOKEN_BAND");
  const int band = ring ? 0 : (eb != NULL ? (atoi(eb) != 0) : (ipb < want && streams >= 148L * 4 * 16));
  if (band) {
    cw = (f_cw >= 1 && f_cw <= 16) ? f_cw : 4;
    lpw = (int)((streams + 148L * cw - 1) / (148L * cw));
    if (f_lpw >= 1 && f_lpw <= 32) lpw = f_lpw;
    lpw = lpw < 1 ? 1 : lpw > 32 ? 32 : lpw;
    while (cw * lpw < P) ++lpw;
    ipb = (cw * lpw) / P;
    while (ipb > 1 && tf_layout(P, ipb, max_mb_w, 0, 1).total > (uint32_t)VP8K_MAX_DYN_SMEM - 1024u) --ipb;
  }
  const TfLayout lay = tf_layout(P, ipb, max_mb_w, ring, band);
  const int blocks = (count + ipb - 1) / ipb;
  // How the lanes are run (vp8_tokens_fp.h): a branch per decode with the block ends handled on the spot while a warp has
  // few lanes, straight-line groups of four decodes with one event point when it has many (the event point's cost is
  // shared by all the lanes that have a block end pending, and lanes that sit out the rest of a group cost nothing extra).
  // Measured per 4096 full-HD images (profiles/r02g): 1 partition (7 lanes per warp) 267 / 335 ms, 8 partitions (28 lanes)
  // 163 / 82 ms. WEBP_B200_TOKEN_GROUPED=0|1 forces one.
  const char* eg = getenv("WEBP_B200_TOKEN_GROUPED");
  const int lanes_per_warp = (ipb * P + cw - 1) / cw;
  const int flat = eg != NULL ? (atoi(eg) != 0) : (lanes_per_warp >= 16);
  // How the compressed bytes reach the reader: one read-only global load per 32-bit refill, issued one refill ahead of
  // its use (default), or (WEBP_B200_TOKEN_RING=1) per-stream shared-memory rings filled 128 bytes at a time by
  // cp.async.bulk with mbarrier completion (vp8_tokens_fp.h:FpDec). Both are built and parity-tested; measured per 4096
  // full-HD images (profiles/r02k): 1 partition 267 ms / 347 ms, 8 partitions 82 / 195, 65536 thumbnails 132 / 162 -- a
  // stream consumes 0.8 bits per decode, so the ring saves one load per ~40 decodes and pays for it with a longer refill path
  // inside a warp whose every instruction is on the critical path.
  if (ring) {
    k_parse_tokens_fp<1, 0><<<blocks, 32 * cw, lay.total, s>>>(arena, imgs, hdrs, mbinfo, tokens, mbtok, ids, count, P, ipb, lpw, cw, max_mb_w, flat);
  } else if (band) {
    k_parse_tokens_fp<0, 1><<<blocks, 32 * cw, lay.total, s>>>(arena, imgs, hdrs, mbinfo, tokens, mbtok, ids, count, P, ipb, lpw, cw, max_mb_w, flat);
  } else {
    k_parse_tokens_fp<0, 0><<<blocks, 32 * cw, lay.total, s>>>(arena, imgs, hdrs, mbinfo, tokens, mbtok, ids, count, P, ipb, lpw, cw, max_mb_w, flat);
  }
}

// Which token parser a wave takes: 1 = the fp parser (token stream out), 0 = one of the older mappings (dense level plane),
// forced by WEBP_B200_TOKEN_MAP=warp|k|lanes (A/B runs, and the row-band pipeline, which only the older lockstep parser has).
extern "C" int vp8k_tokens_use_stream(void) {
  const char* e = getenv("WEBP_B200_TOKEN_MAP");
  return e == NULL || e[0] == 'f';
}

extern "C" void vp8k_parse_tokens_stream(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                         uint32_t* tokens, void* mbtok, const int* ids, int count, int P, int max_mb_w) {
  launch_tokens_fp(s, arena, imgs, hdrs, mbinfo, tokens, (MbTok*)mbtok, ids, count, P, max_mb_w);
}

extern "C" int vp8k_tokens_take_bands(int count, int P) {   // does vp8k_parse_tokens pick the mapping that can parse by row bands?
  const char* e = getenv("WEBP_B200_TOKEN_MAP");
  if (e != NULL && e[0] != 'k') return 0;
  return P == 1 && ((e != NULL && e[0] == 'k') || (long)count * P >= 148L * 4 * 2);
}

extern "C" void vp8k_parse_tokens_band(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                       int16_t* coeffs, const int* ids, int count, int max_mb_w, int row_begin, int row_end,
                                       TokResume* resume, uint16_t* resume_ctx) {
  launch_tokens_lockstep(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, 1, max_mb_w, row_begin, row_end, resume, resume_ctx);
}

extern "C" void vp8k_parse_tokens(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                  int16_t* coeffs, const int* ids, int count, int P, int max_mb_w) {
  // Three mappings of the same parse: one warp per partition (straight-line code on one lane; seven warps per SM
  // sub-partition keep its issue port busy), the lockstep lanes (vp8_tokens_lockstep.h: a third of the issue slots
  // per decode, but one iteration costs a whole dependent chain, so it needs several warps per sub-partition to
  // pay), and the older table-driven state machine (vp8_tokens_fsm.h). WEBP_B200_TOKEN_MAP=warp|k|lanes forces one.
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("WEBP_B200_TOKEN_MAP");
    forced = (e && e[0] == 'w') ? 1 : (e && e[0] == 'l') ? 2 : (e && e[0] == 'k') ? 3 : 0;
  }
  // Measured per 4096 full-HD images (profiles/r01*_token_map_sweep.log): 4096 streams: warp 377 ms, lockstep 331;
  // 32768 streams (8 partitions): warp 420, state machine 247, lockstep 162; 65536 thumbnails: warp 274, lockstep 178.
  // Below two streams per SM sub-partition a lane-per-stream warp has nothing to share its instructions with.
  const int many = (long)count * P >= 148L * 4 * 2;
  if (forced == 3 || (forced == 0 && many)) {
    launch_tokens_lockstep(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w, 0, 0x7fffffff, nullptr, nullptr);
    return;
  }
  const int use_warp_map = forced ? (forced == 1) : 1;
  if (use_warp_map) {
    // one block per SM where the launch fits in one wave, else as many images per block as the block may hold
    static int f_ipb = -1;
    if (f_ipb < 0) f_ipb = env_int("WEBP_B200_TOKEN_IPB");
    const int slot = (int)tokens_slot_bytes(P, max_mb_w);
    int ipb = TOKW_MAX_WARPS / P;
    if (f_ipb >= 1 && f_ipb <= ipb) ipb = f_ipb;
    while (ipb > 1 && (size_t)ipb * slot > 200u * 1024u) --ipb;
    ipb = pack_per_block(count, ipb);
    k_parse_tokens<<<(count + ipb - 1) / ipb, 32 * ipb * P, (size_t)ipb * slot, s>>>(arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, ipb, slot);
    return;
  }
  launch_tokens_fsm(s, arena, imgs, hdrs, mbinfo, coeffs, ids, count, P, max_mb_w);
}

extern "C" void vp8k_parse_literal(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo,
                                   uint32_t* tokens, void* mbtok, int first, int count, int max_mb_w) {
  k_parse_literal<<<count, 32, LIT_SCRATCH_BYTES(max_mb_w), s>>>(arena, imgs, hdrs, mbinfo, tokens, (MbTok*)mbtok, first);
}

// Warps per image of the two wavefront kernels (K3, K4): a warp owns every eighth (fourth) macroblock row and trails the row
// above by two macroblocks, so a launch of small pictures spends much of its time filling and draining that pipeline with 8
// warps (a 16 x 16-macroblock thumbnail: 14 macroblock times of stagger on 32 of work per warp); 4 warps halve the stagger and
// twice as many images are resident, 2 warps once more for pictures of up to 16 rows. WEBP_B200_PIXEL_WARPS=2|4|8 forces one
// (A/B: profiles/r03i, r03r: 65536 thumbnails K3 24.0 / 20.1 / 19.3 ms, K4 27.2 / 19.9 / 19.2 ms at 8 / 4 / 2 warps).
static int pixel_warps_for(int max_mb_h) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("WEBP_B200_PIXEL_WARPS"); const int v = e != NULL ? atoi(e) : 0; forced = (v == 2 || v == 4 || v == 8) ? v : 0; }
  if (forced) return forced;
  return max_mb_h <= 16 ? 2 : max_mb_h <= 32 ? 4 : 8;
}

extern "C" void vp8k_reconstruct(cudaStream_t s, const ImgDesc* imgs, FrameHdr* hdrs, uint32_t* mbinfo, const int16_t* coeffs,
                                 uint8_t* yuv, int first, int count, int max_mb_w, int max_mb_h, int row_begin, int row_end,
                                 uint8_t* band_ctx, const uint32_t* tokens, const void* mbtok) {
  static int forced_warps = -1;   // process-wide A/B switch, read once
  if (forced_warps < 0) { const char* e = getenv("WEBP_B200_RECON_WARPS"); const int v = e != NULL ? atoi(e) : 0; forced_warps = (v == 2 || v == 4 || v == 8 || v == 16) ? v : 0; }
  const int warps = forced_warps ? forced_warps : pixel_warps_for(max_mb_h);
  const size_t smem = recon_smem_bytes(warps, max_mb_w, max_mb_h);
  const int bctx = 32 * max_mb_w;
  const MbTok* mt = (const MbTok*)mbtok;
  static int rows = -1;
  if (rows < 0) { const char* e = getenv("WEBP_B200_RECON_ROWS"); rows = e != NULL ? (atoi(e) != 0) : 1; }
#define RECON_LAUNCH(W, R) k_reconstruct<W, R><<<count, 32 * W, smem, s>>>(imgs, hdrs, mbinfo, coeffs, yuv, first, row_begin, row_end, band_ctx, bctx, tokens, mt)
  if (rows) { if (warps == 2) RECON_LAUNCH(2, 1); else if (warps == 4) RECON_LAUNCH(4, 1); else if (warps == 16) RECON_LAUNCH(16, 1); else RECON_LAUNCH(8, 1); }
  else { if (warps <= 4) RECON_LAUNCH(4, 0); else if (warps == 16) RECON_LAUNCH(16, 0); else RECON_LAUNCH(8, 0); }
#undef RECON_LAUNCH
}

extern "C" void vp8k_loop_filter(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint32_t* mbinfo, uint8_t* yuv,
                                 int first, int count, int max_mb_h, int row_begin, int row_end, const int8_t* dither_plane) {
  const int fw = pixel_warps_for(max_mb_h);
  if (fw == 2) k_loop_filter<2><<<count, 64, (size_t)max_mb_h * 4 + 16, s>>>(imgs, hdrs, mbinfo, yuv, first, row_begin, row_end, dither_plane);
  else if (fw == 4) k_loop_filter<4><<<count, 128, (size_t)max_mb_h * 4 + 16, s>>>(imgs, hdrs, mbinfo, yuv, first, row_begin, row_end, dither_plane);
  else k_loop_filter<8><<<count, 256, (size_t)max_mb_h * 4 + 16, s>>>(imgs, hdrs, mbinfo, yuv, first, row_begin, row_end, dither_plane);
}

extern "C" void vp8k_collect_status(cudaStream_t s, const FrameHdr* hdrs, int* host_statuses, int count) {
  k_collect_status<<<(count + 255) / 256, 256, 0, s>>>(hdrs, host_statuses, count);
}

extern "C" void vp8k_copy_to_host(cudaStream_t s, const void* src, void* host_dst, size_t bytes) {
  const size_t blocks = (bytes + 255) / 256;
  k_copy_to_host<<<(unsigned)(blocks < 1024 ? (blocks ? blocks : 1) : 1024), 256, 0, s>>>((const uint8_t*)src, (uint8_t*)host_dst, bytes);
}

extern "C" void vp8k_dither_plan(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, uint32_t* mbinfo, int8_t* dither_plane,
                                 int first, int count) {
  k_dither_plan<<<(count + 31) / 32, 32, 0, s>>>(imgs, hdrs, mbinfo, dither_plane, first, count);
}

extern "C" void vp8k_emit(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, const uint8_t* alpha_arena,
                          uint8_t* out, int first, int count, int max_units, int pair_begin, int pair_end) {
  const int bpi = (max_units + EMIT_THREADS - 1) / EMIT_THREADS;
  k_emit<<<(unsigned)count * (unsigned)bpi, EMIT_THREADS, 0, s>>>(imgs, hdrs, yuv, alpha_arena, out, first, bpi, pair_begin, pair_end);
}

extern "C" void vp8k_emit_scaled(cudaStream_t s, const ImgDesc* imgs, const FrameHdr* hdrs, const uint8_t* yuv, const uint8_t* alpha_arena,
                                 uint8_t* out, int first, int count, int max_items) {
  const int bpi = (max_items + EMIT_THREADS - 1) / EMIT_THREADS;
  k_emit_scaled<<<(unsigned)count * (unsigned)bpi, EMIT_THREADS, 0, s>>>(imgs, hdrs, yuv, alpha_arena, out, first, bpi);
}

// ---------------------------------------------------------------------------------------------------------
// ALPH chunks (vp8l_alpha_core.h). `aimgs` = indices of the images that carry one. Work areas are byte offsets
// into `work` (host-planned, AlphaPlan).
__global__ void __launch_bounds__(32) k_alpha_header(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                     const int* __restrict__ aimgs, const AlphaPlan* __restrict__ plans,
                                                     AlphaHdr* ahdrs) {
  if (threadIdx.x != 0) return;
  const int a = blockIdx.x;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  alph_parse_header(arena + im.alpha_in, im.alpha_size, im.width, im.height, (uint8_t*)pl.scratch, (uint16_t*)pl.meta,
                    (uint32_t*)pl.tdata, &ahdrs[a], (im.flags & VP8B_FLAG_LOSSLESS) ? 1 : 0);
}

// The pixel loop of a picture is serial (prefix codes, LZ77): one lane per picture. Two kernels over the same list, each passing
// by the pictures of the other: ALPH planes without a colour cache let lane 0 queue their backward references and the whole warp
// carry them out (alph_decode_pixels_warp: alpha planes are long runs, 512 planes of 4096x4096 384 -> 150 ms); pictures with
// a colour cache (the copies feed it in order) and whole VP8L pictures (photographs: literals and short copies, the queued form
// costs them 20 %) keep the one-lane loop, in a kernel of its own because that loop is one dependent chain whose speed moves by
// 10-20 % with the code compiled around it (1024 full-HD photographs: 1126 ms alone, 1296 ms sharing a kernel with the other
// loop; profiles/r02y).
AL_FN int alpha_takes_the_warp(const AlphaHdr* hd) { return hd->cache_bits == 0 && !hd->lossless; }

__global__ void __launch_bounds__(32) k_alpha_pixels(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                     const int* __restrict__ aimgs, const AlphaPlan* __restrict__ plans,
                                                     AlphaHdr* ahdrs) {
  if (threadIdx.x != 0) return;
  const int a = blockIdx.x;
  AlphaHdr* hd = &ahdrs[a];
  if (hd->status != AL_OK || hd->method == 0 || alpha_takes_the_warp(hd)) return;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  hd->status = alph_decode_pixels(arena + im.alpha_in, im.alpha_size, im.height, (int)im.crop_y + (int)im.out_h, hd, (const uint16_t*)pl.meta,
                                  (uint32_t*)pl.tables, (AlGroup*)pl.groups, (uint8_t*)pl.scratch, (uint32_t*)pl.coded);
}

__global__ void __launch_bounds__(32) k_alpha_pixels_warp(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                          const int* __restrict__ aimgs, const AlphaPlan* __restrict__ plans,
                                                          AlphaHdr* ahdrs) {
  const int a = blockIdx.x;
  AlphaHdr* hd = &ahdrs[a];
  if (hd->status != AL_OK || hd->method == 0 || !alpha_takes_the_warp(hd)) return;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  const int st = alph_decode_pixels_warp(arena + im.alpha_in, im.alpha_size, im.height, (int)im.crop_y + (int)im.out_h, hd, (const uint16_t*)pl.meta,
                                         (uint32_t*)pl.tables, (AlGroup*)pl.groups, (uint8_t*)pl.scratch, (uint32_t*)pl.coded,
                                         AL_COPYQ_PTR(pl.scratch), (int)threadIdx.x);
  __syncwarp();
  if (threadIdx.x == 0) hd->status = st;
}

#define ALPHA_FINISH_THREADS 1024
__global__ void __launch_bounds__(ALPHA_FINISH_THREADS) k_alpha_finish(const uint8_t* __restrict__ arena, const ImgDesc* __restrict__ imgs,
                                                                       const int* __restrict__ aimgs, const AlphaPlan* __restrict__ plans,
                                                                       const AlphaHdr* __restrict__ ahdrs, uint8_t* alpha_arena) {
  const int a = blockIdx.x;
  const AlphaHdr* hd = &ahdrs[a];
  if (hd->status != AL_OK) return;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  if (im.alpha_plane == VP8B_NO_ALPHA) return;
  alph_finish(hd, arena + im.alpha_in + 1, (uint32_t*)pl.coded, (const uint32_t*)pl.tdata, im.width, im.height, im.crop_y,
              alpha_arena + im.alpha_plane, (int)threadIdx.x, (int)blockDim.x);
}

// options.alpha_dithering_strength (vp8l_alpha_core.h:alph_smooth) on the crop window of the finished plane.
__global__ void __launch_bounds__(ALPHA_FINISH_THREADS) k_alpha_smooth(const ImgDesc* __restrict__ imgs, const int* __restrict__ aimgs,
                                                                       const AlphaPlan* __restrict__ plans, const AlphaHdr* __restrict__ ahdrs,
                                                                       uint8_t* alpha_arena) {
  __shared__ int16_t lut[2048];
  __shared__ uint32_t used[256];
  const int a = blockIdx.x;
  if (ahdrs[a].status != AL_OK) return;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  if (im.alpha_dither == 0 || pl.smooth == 0 || im.alpha_plane == VP8B_NO_ALPHA) return;
  alph_smooth(alpha_arena + im.alpha_plane + (size_t)im.crop_y * im.width + im.crop_x, im.width, im.out_w, im.out_h, im.alpha_dither,
              (uint16_t*)pl.smooth, lut, used, (int)threadIdx.x, (int)blockDim.x);
}

// Whole-picture VP8L (vp8l_lossless_core.h): inverse transforms, palette, colourspace conversion into the output arena.
__global__ void __launch_bounds__(ALPHA_FINISH_THREADS) k_lossless_finish(const ImgDesc* __restrict__ imgs, const int* __restrict__ aimgs,
                                                                          const AlphaPlan* __restrict__ plans, const AlphaHdr* __restrict__ ahdrs,
                                                                          uint8_t* out) {
  const int a = blockIdx.x;
  const AlphaHdr* hd = &ahdrs[a];
  if (hd->status != AL_OK || !hd->lossless) return;
  const ImgDesc im = imgs[aimgs[a]];
  const AlphaPlan pl = plans[a];
  vp8l_finish_picture(hd, im, (uint32_t*)pl.coded, (const uint32_t*)pl.tdata, out + im.out_off, (uint8_t*)pl.smooth, (int)threadIdx.x, (int)blockDim.x);
}

extern "C" void vp8k_alpha_header(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans,
                                  AlphaHdr* ahdrs, int count) {
  k_alpha_header<<<count, 32, 0, s>>>(arena, imgs, aimgs, plans, ahdrs);
}
extern "C" void vp8k_alpha_decode(cudaStream_t s, const uint8_t* arena, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans,
                                  AlphaHdr* ahdrs, uint8_t* alpha_arena, int count) {
  k_alpha_pixels<<<count, 32, 0, s>>>(arena, imgs, aimgs, plans, ahdrs);
  k_alpha_pixels_warp<<<count, 32, 0, s>>>(arena, imgs, aimgs, plans, ahdrs);
  k_alpha_finish<<<count, ALPHA_FINISH_THREADS, 0, s>>>(arena, imgs, aimgs, plans, ahdrs, alpha_arena);
  k_alpha_smooth<<<count, ALPHA_FINISH_THREADS, 0, s>>>(imgs, aimgs, plans, ahdrs, alpha_arena);
}
extern "C" void vp8k_lossless_finish(cudaStream_t s, const ImgDesc* imgs, const int* aimgs, const AlphaPlan* plans, const AlphaHdr* ahdrs,
                                     uint8_t* out, int count) {
  k_lossless_finish<<<count, ALPHA_FINISH_THREADS, 0, s>>>(imgs, aimgs, plans, ahdrs, out);
}
