// vp8_pixel_core.h -- pixel side of the batched VP8 decoder: inverse WHT/DCT + intra prediction (one warp per
// macroblock inside a lag-2 macroblock wavefront), the in-loop deblocking filter (same wavefront shape), and
// the output stage (fancy chroma upsampling + YUV->RGB, or plane copies).
//
// Replaces, for a whole batch at once:
//   ReconstructRow src/dec/frame_dec.c:71-196; TransformOne/AC3/DC/UV/DCUV/WHT src/dsp/dec.c:44-162;
//   VP8PredLuma4/16, VP8PredChroma8 src/dsp/dec.c:173-474
//   DoFilter/FilterRow src/dec/frame_dec.c:203-260; Simple*/[VH]Filter* src/dsp/dec.c:484-693
//   EmitFancyRGB/EmitSampledRGB/EmitYUV src/dec/io_dec.c:25-109; UPSAMPLE_FUNC src/dsp/upsampling.c:37-93;
//   VP8YuvToRgb* src/dsp/yuv.h:59-144
//
// Lanes of a warp cooperate through a per-warp shared-memory workspace; every exchange point is a warp
// barrier (WARP_PHASE ... WARP_PHASE_END). Under -DVP8_EMU (tests/emu) a phase becomes a loop over 32 lanes
// so the same logic runs on a host without a GPU for unit tests; that build is never shipped.
#ifndef LIBWEBP_B200_VP8_PIXEL_CORE_H_
#define LIBWEBP_B200_VP8_PIXEL_CORE_H_

#include "vp8_dev.h"

#if defined(__CUDACC__) && !defined(VP8_EMU)
#define VP8_PFN __device__ __forceinline__
#define VP8_PTABLE static __constant__ const
#define WARP_PHASE(lane) { const int lane = (int)(threadIdx.x & 31);
#define WARP_PHASE_END } __syncwarp();
#define VP8_PCLZ(x) __clz((int)(x))
#define VP8_UNROLL _Pragma("unroll")
#else
#define VP8_PFN static inline
#define VP8_PTABLE static const
#define WARP_PHASE(lane) for (int lane = 0; lane < 32; ++lane) {
#define WARP_PHASE_END }
#define VP8_PCLZ(x) __builtin_clz((unsigned)(x))
#define VP8_UNROLL
#ifndef VP8_EMU_VECTORS
#define VP8_EMU_VECTORS
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
#endif
#endif

VP8_PFN int clip8i(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
VP8_PFN int sum4(uint32_t v) {   // the four bytes of a word added up
#if defined(__CUDACC__) && !defined(VP8_EMU)
  return (int)__dp4a(v, 0x01010101u, 0u);
#else
  return (int)((v & 255u) + ((v >> 8) & 255u) + ((v >> 16) & 255u) + (v >> 24));
#endif
}
VP8_PFN int mul1(int a) { return ((a * 20091) >> 16) + a; }
VP8_PFN int mul2(int a) { return (a * 35468) >> 16; }

// ---------------------------------------------------------------------------------------------------------
// Reconstruction workspace of one warp (shared memory). Tile coordinates: luma pixel (r, c) of the macroblock
// lives at y[(r + 1) * 32 + c + 4]; row 0 is the row above, column 3 the column to the left, columns 20-23
// the four pixels above-right (replicated at rows 4, 8, 12 for the right-most sub-blocks, frame_dec.c:131-141).
// Chroma: u at uv[(r + 1) * 32 + c + 4], v at uv[(r + 1) * 32 + c + 20].
struct ReconWs {
  int16_t lv[25 * 16];                // the macroblock's coefficient LEVELS, parse order inside a block (blocks 0-23, 24 = Y2): the
                                      // token parser's stream scattered out by recon_load_tokens (16-byte aligned: first member)
  int16_t res[24 * 16];               // residual (IDCT output, already >> 3) of blocks 0-23, raster order inside a block
  uint8_t y[17 * 32];
  uint8_t uv[9 * 32];
  int16_t dc[16];                     // inverse WHT of the Y2 block: DC coefficient of the 16 luma blocks (i16 only)
  uint8_t spare[32];
  uint32_t nz;                        // some block of the macroblock has a non-zero coefficient
  uint32_t pad[3];
};

// Wavefront context of one image (shared memory of the block that owns the image): the unfiltered pixels each
// macroblock needs from its neighbours. Lag 2 between rows keeps every entry valid exactly while it is read.
// Laid out so that a macroblock finds everything at one constant offset from its column / its row:
//   top  : 32 bytes per macroblock column, the bottom row of the macroblock above: 16 luma | 8 U | 8 V
//   left : 36 bytes per macroblock row, the right column of the macroblock to the left: 16 luma | 8 U | 8 V, then the
//          pixel above-left of the next macroblock of the row for y, u, v (+ 1 byte of padding)
struct ReconCtx {
  uint8_t* top;
  uint8_t* left;
  const uint32_t* pred4;   // kPred4x where every lane can read its own word at once (shared memory on the device)
};

#define recon_ctx_bytes(mb_w, mb_h) ((size_t)32 * (size_t)(mb_w) + (size_t)36 * (size_t)(mb_h))

VP8_PFN void recon_ctx_bind(ReconCtx& c, uint8_t* mem, int mb_w, int mb_h) {
  (void)mb_h;
  c.top = mem; c.left = mem + 32 * (size_t)mb_w;
}

// 4x4 predictors, one word per (mode, pixel): three tile offsets (bytes 0-2, each + 64, relative to the sub-block's
// origin in the 32-byte-stride tile) of the pixels a, b, c the prediction reads directly from the tile:
// (a + 2b + c + 2) >> 2 for VE..HU (the two-tap averages (a + b + 1) >> 1 are the same formula with c = a), a + b - c
// clipped for TM. Neighbours by offset: left column I J K L = -1, 31, 63, 95 (L also stands for the pixels below it),
// corner X = -33, row above A..H = -32..-25 (dsp/dec.c:259-440 restated over one table). Row 0 (DC) is not read.
VP8_PTABLE uint32_t kPred4x[10][16] = {
  /* DC */ { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 },
  /* TM */ { 0x1f3f20, 0x1f3f21, 0x1f3f22, 0x1f3f23, 0x1f5f20, 0x1f5f21, 0x1f5f22, 0x1f5f23, 0x1f7f20, 0x1f7f21, 0x1f7f22, 0x1f7f23, 0x1f9f20, 0x1f9f21, 0x1f9f22, 0x1f9f23 },
  /* VE */ { 0x21201f, 0x222120, 0x232221, 0x242322, 0x21201f, 0x222120, 0x232221, 0x242322, 0x21201f, 0x222120, 0x232221, 0x242322, 0x21201f, 0x222120, 0x232221, 0x242322 },
  /* HE */ { 0x1f3f5f, 0x1f3f5f, 0x1f3f5f, 0x1f3f5f, 0x3f5f7f, 0x3f5f7f, 0x3f5f7f, 0x3f5f7f, 0x5f7f9f, 0x5f7f9f, 0x5f7f9f, 0x5f7f9f, 0x7f9f9f, 0x7f9f9f, 0x7f9f9f, 0x7f9f9f },
  /* RD */ { 0x201f3f, 0x21201f, 0x222120, 0x232221, 0x1f3f5f, 0x201f3f, 0x21201f, 0x222120, 0x3f5f7f, 0x1f3f5f, 0x201f3f, 0x21201f, 0x5f7f9f, 0x3f5f7f, 0x1f3f5f, 0x201f3f },
  /* VR */ { 0x1f201f, 0x202120, 0x212221, 0x222322, 0x201f3f, 0x21201f, 0x222120, 0x232221, 0x1f3f5f, 0x1f201f, 0x202120, 0x212221, 0x3f5f7f, 0x201f3f, 0x21201f, 0x222120 },
  /* LD */ { 0x222120, 0x232221, 0x242322, 0x252423, 0x232221, 0x242322, 0x252423, 0x262524, 0x242322, 0x252423, 0x262524, 0x272625, 0x252423, 0x262524, 0x272625, 0x272726 },
  /* VL */ { 0x202120, 0x212221, 0x222322, 0x232423, 0x222120, 0x232221, 0x242322, 0x252423, 0x212221, 0x222322, 0x232423, 0x262524, 0x232221, 0x242322, 0x252423, 0x272625 },
  /* HD */ { 0x3f1f3f, 0x201f3f, 0x21201f, 0x222120, 0x5f3f5f, 0x1f3f5f, 0x3f1f3f, 0x201f3f, 0x7f5f7f, 0x3f5f7f, 0x5f3f5f, 0x1f3f5f, 0x9f7f9f, 0x5f7f9f, 0x7f5f7f, 0x3f5f7f },
  /* HU */ { 0x5f3f5f, 0x3f5f7f, 0x7f5f7f, 0x5f7f9f, 0x7f5f7f, 0x5f7f9f, 0x9f7f9f, 0x7f9f9f, 0x9f7f9f, 0x7f9f9f, 0x9f9f9f, 0x9f9f9f, 0x9f9f9f, 0x9f9f9f, 0x9f9f9f, 0x9f9f9f }
};

// 16 coefficient levels of one block (parse order, two 16-byte words of HBM) -> dequantised int16 values in raster
// order (GetCoeffs' `out[kZigzag[n]] = level * dq[n > 0]` with the int16 store, vp8_dec.c:463-466).
VP8_PFN void load_block_coeffs(const int16_t* levels, int q_dc, int q_ac, int in[16]) {
  const uint4 a = ((const uint4*)levels)[0], b = ((const uint4*)levels)[1];
  const uint32_t wd[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
  // raster position of parse position n (kZigzag, vp8_dec.c:406)
  const int zz[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
  VP8_UNROLL
  for (int n = 0; n < 16; ++n) {
    const int lvl = (n & 1) ? ((int)wd[n >> 1] >> 16) : (int)(int16_t)(wd[n >> 1] & 0xffffu);
    in[zz[n]] = (int)(int16_t)(lvl * (n == 0 ? q_dc : q_ac));
  }
}

// A macroblock's tokens (vp8_tokens_fp.h: {sign 31, block 29:25, magnitude 24:13, position 9:6}, `count` words at
// `toks`) -> ws.lv, the dense 25 x 16 level array load_block_coeffs reads. One warp, in two parts so that the latency of
// the loads runs beside other work: recon_fetch_tokens issues the loads of the first 96 tokens into registers (the
// row-per-warp kernel does that for macroblock k + 1 while it finishes macroblock k), recon_scatter_tokens writes them out
// once ws.lv has been cleared and loops over whatever is left.
#if defined(__CUDACC__) && !defined(VP8_EMU)
#define VP8_LANE_SLOTS 1          // a value that lives in a lane across phases: a register on the device,
#define VP8_LANE_SLOT(lane) 0
#else
#define VP8_LANE_SLOTS 32         // one slot per lane in the host build
#define VP8_LANE_SLOT(lane) (lane)
#endif
struct ReconTok { uint32_t t[3][VP8_LANE_SLOTS]; };

VP8_PFN void recon_put_token(ReconWs& ws, uint32_t t) {
  const int mag = (int)((t >> 13) & 0xfffu);
  ws.lv[((t >> 25) & 31u) * 16u + ((t >> 6) & 15u)] = (int16_t)((t >> 31) ? -mag : mag);
}

VP8_PFN void recon_fetch_tokens(ReconTok& rt, const uint32_t* toks, uint32_t count) {   // loads only: nothing waits for them here
  WARP_PHASE(lane)
    VP8_UNROLL
    for (int j = 0; j < 3; ++j) {
      const uint32_t k = (uint32_t)lane + 32u * (uint32_t)j;
      rt.t[j][VP8_LANE_SLOT(lane)] = (k < count) ? toks[k] : 0u;
    }
  WARP_PHASE_END
}

VP8_PFN void recon_clear_levels(ReconWs& ws) {
  WARP_PHASE(lane)
    uint4 z; z.x = 0; z.y = 0; z.z = 0; z.w = 0;
    uint4* dst = (uint4*)ws.lv;
    if (lane < 25) { dst[2 * lane] = z; dst[2 * lane + 1] = z; }
  WARP_PHASE_END
}

VP8_PFN void recon_scatter_tokens(ReconWs& ws, const ReconTok& rt, const uint32_t* toks, uint32_t count) {
  WARP_PHASE(lane)
    VP8_UNROLL
    for (int j = 0; j < 3; ++j) {
      if ((uint32_t)lane + 32u * (uint32_t)j < count) recon_put_token(ws, rt.t[j][VP8_LANE_SLOT(lane)]);
    }
    for (uint32_t k = (uint32_t)lane + 96u; k < count; k += 32) recon_put_token(ws, toks[k]);
  WARP_PHASE_END
}

// TransformOne_C (dsp/dec.c:44-82) without the prediction: out = the 16 residuals to add, raster order.
VP8_PFN void idct_block(const int in[16], int out[16]) {
  int tmp[16];
  VP8_UNROLL
  for (int i = 0; i < 4; ++i) {   // vertical pass
    const int a = in[i] + in[8 + i], b = in[i] - in[8 + i];
    const int c = mul2(in[4 + i]) - mul1(in[12 + i]), d = mul1(in[4 + i]) + mul2(in[12 + i]);
    tmp[4 * i + 0] = a + d; tmp[4 * i + 1] = b + c; tmp[4 * i + 2] = b - c; tmp[4 * i + 3] = a - d;
  }
  VP8_UNROLL
  for (int i = 0; i < 4; ++i) {   // horizontal pass: output row i
    const int dc = tmp[i] + 4;
    const int a = dc + tmp[8 + i], b = dc - tmp[8 + i];
    const int c = mul2(tmp[4 + i]) - mul1(tmp[12 + i]), d = mul1(tmp[4 + i]) + mul2(tmp[12 + i]);
    out[4 * i + 0] = (a + d) >> 3; out[4 * i + 1] = (b + c) >> 3; out[4 * i + 2] = (b - c) >> 3; out[4 * i + 3] = (a - d) >> 3;
  }
}

// TransformWHT_C (dsp/dec.c:137-162): 16 dequantised Y2 coefficients (raster order) -> DC of luma block k in dc[k].
VP8_PFN void wht_block(const int in[16], int16_t dc[16]) {
  int tmp[16];
  VP8_UNROLL
  for (int i = 0; i < 4; ++i) {
    const int a0 = in[0 + i] + in[12 + i], a1 = in[4 + i] + in[8 + i];
    const int a2 = in[4 + i] - in[8 + i], a3 = in[0 + i] - in[12 + i];
    tmp[0 + i] = a0 + a1; tmp[8 + i] = a0 - a1; tmp[4 + i] = a3 + a2; tmp[12 + i] = a3 - a2;
  }
  VP8_UNROLL
  for (int i = 0; i < 4; ++i) {
    const int d0 = tmp[0 + i * 4] + 3;
    const int a0 = d0 + tmp[3 + i * 4], a1 = tmp[1 + i * 4] + tmp[2 + i * 4];
    const int a2 = tmp[1 + i * 4] - tmp[2 + i * 4], a3 = d0 - tmp[3 + i * 4];
    dc[4 * i + 0] = (int16_t)((a0 + a1) >> 3); dc[4 * i + 1] = (int16_t)((a3 + a2) >> 3);
    dc[4 * i + 2] = (int16_t)((a0 - a1) >> 3); dc[4 * i + 3] = (int16_t)((a3 - a2) >> 3);
  }
}

// 16x16 / 8x8 prediction of one pixel. t = pointer to the block origin inside a 32-byte-stride tile.
// `mode` after the border substitution of CheckMode (frame_dec.c:28-37): 0 DC, 1 TM, 2 V, 3 H, 4 DC without
// top, 5 DC without left, 6 DC without both. `dc` = precomputed DC value for the DC modes.
// N pixels of row y from column x0 on: the mode (the same for the whole macroblock) is looked at once, not per pixel.
template <int N>
VP8_PFN void pred_big_run(const uint8_t* t, int mode, int dc, int x0, int y, int v[N]) {
  if (mode == 1) {
    const int d = (int)t[y * 32 - 1] - (int)t[-33];
    VP8_UNROLL
    for (int k = 0; k < N; ++k) v[k] = clip8i((int)t[x0 + k - 32] + d);
  } else if (mode == 2) {
    VP8_UNROLL
    for (int k = 0; k < N; ++k) v[k] = t[x0 + k - 32];
  } else {
    const int c = (mode == 3) ? (int)t[y * 32 - 1] : dc;
    VP8_UNROLL
    for (int k = 0; k < N; ++k) v[k] = c;
  }
}

VP8_PFN int dc_big(const uint8_t* t, int mode, int size) {   // dsp/dec.c:215-244, 445-474
  const int sh = (size == 16) ? 4 : 3;
  int s = 0;
  const uint32_t* const above = (const uint32_t*)(t - 32);   // word-aligned in both tiles
  if (mode == 0) {
    for (int i = 0; i < size / 4; ++i) s += sum4(above[i]);
    for (int i = 0; i < size; ++i) s += t[i * 32 - 1];
    return (s + size) >> (sh + 1);
  } else if (mode == 4) {
    for (int i = 0; i < size; ++i) s += t[i * 32 - 1];
    return (s + (size >> 1)) >> sh;
  } else if (mode == 5) {
    for (int i = 0; i < size / 4; ++i) s += sum4(above[i]);
    return (s + (size >> 1)) >> sh;
  }
  return 0x80;
}

VP8_PFN int check_mode(int mx, int my, int mode) {
  if (mode == M_DC) return (mx == 0) ? ((my == 0) ? 6 : 5) : ((my == 0) ? 4 : 0);
  return mode;
}

// Reconstructs macroblock (mx, my) of one image with one warp. `info` = this macroblock's MbInfo (4 words),
// `coeffs` = its 400 coefficient LEVELS in HBM (parse order inside each block, as the token parser left them),
// dq6 = the dequantisers of the macroblock's segment {y1 dc, ac, y2 dc, ac, uv dc, ac} (VP8ParseQuant,
// quant_dec.c:62-112), planes = the image's padded Y/U/V in HBM (write-only here).
// Order of work: (0) neighbour pixels into the tile, inverse WHT by one lane; (1) one lane per 4x4 block:
// dequantise + inverse DCT in registers, residuals to shared; (2) luma prediction + residual: all 256 pixels at
// once for i16, a 10-step sub-block wavefront (two sub-blocks per step) for i4x4; (3) chroma; (4) tile -> planes.
// Also finalises MbInfo: the filter-inner bit (vp8_dec.c:629-633), which needs the lone-DC rule
// NzCodeBits(nz, dst[0] != 0) (vp8_dec.c:511-515) evaluated on the dequantised int16.
// `iw` = the four MbInfo words (the caller may have fetched them ahead of time), `info` = where they live (word 3 is rewritten).
// `rt` = this macroblock's first tokens, fetched by the caller (recon_fetch_tokens); on return it holds those of the
// macroblock the same warp takes next (`next_toks`, `next_ntok`; 0 for none), their loads still in flight.
VP8_PFN void recon_macroblock(ReconWs& ws, const ReconCtx& cx, int mx, int my, int mb_w, const uint4 iw, uint32_t* info,
                              const int16_t* coeffs, const int16_t* dq6, uint8_t* yplane, uint8_t* uplane, uint8_t* vplane,
                              const uint32_t* toks, uint32_t ntok, ReconTok& rt, const uint32_t* next_toks, uint32_t next_ntok) {
  const uint32_t w = iw.w;
  const uint32_t m0 = iw.x, m1 = iw.y;
  const uint32_t nzy = iw.z;
  const uint32_t nzuv = w & 0xffffu;
  const int is_i4 = (w & MBW_I4X4) != 0;
  const int has_y2 = (w & MBW_HAS_Y2) != 0;
  const int any_coef = (nzy | nzuv) != 0 || has_y2;
  const int ys = 16 * mb_w, uvs = 8 * mb_w;
  // levels: the dense plane of the older token parsers, or this macroblock's slice of the token stream
  const int from_tokens = toks != nullptr && any_coef;
  if (toks != nullptr) coeffs = ws.lv;
  if (from_tokens) recon_clear_levels(ws);

  // ---- phase 0: neighbour pixels -> tile (the token loads are in flight); then the levels, then the inverse WHT of the Y2 block
  // Left columns: one byte per lane (16 luma, 8 U, 8 V). Rows above: one word per lane (lanes 0-3 luma, 4 the four pixels
  // above-right, 5-6 U, 7-8 V), corners lanes 9-11. Outside the frame: 127 above, 129 to the left, the corner like the
  // row above unless that exists and the left does not (frame_dec.c:91-121).
  WARP_PHASE(lane)
    if (lane == 0) ws.nz = 0;
    {
      uint8_t* dst = lane < 16 ? ws.y + (lane + 1) * 32 + 3 : ws.uv + ((lane & 7) + 1) * 32 + (lane < 24 ? 3 : 19);
      *dst = (mx > 0) ? cx.left[36 * my + lane] : 129;
    }
    if (lane < 9) {
      uint32_t v = 0x7f7f7f7fu;
      uint32_t* dst = (lane < 5) ? (uint32_t*)(ws.y + 4 + 4 * lane) : (uint32_t*)(ws.uv + 4 + 4 * ((lane - 5) & 1) + 16 * ((lane - 5) >> 1));
      if (my > 0) {
        // word of the column's 32 bytes: 0-3 luma, 4-5 U, 6-7 V; the four pixels above-right are the next column's first word
        const uint32_t* col = (const uint32_t*)(cx.top + 32 * mx);
        v = (lane == 4 && mx == mb_w - 1) ? cx.top[32 * mx + 15] * 0x01010101u : col[lane < 4 ? lane : lane == 4 ? 8 : lane - 1];
      }
      if (lane == 4 && is_i4) { dst[32] = v; dst[64] = v; dst[96] = v; }   // rows 4, 8, 12 of the tile: above-right of the right-most sub-blocks
      *dst = v;
    } else if (lane < 12) {
      const int k = lane - 9;   // corner of y, u, v
      (k == 0 ? ws.y : ws.uv)[k == 2 ? 19 : 3] = (my > 0) ? ((mx > 0) ? cx.left[36 * my + 32 + k] : 129) : 127;
    }
  WARP_PHASE_END
  if (from_tokens) recon_scatter_tokens(ws, rt, toks, ntok);
  if (!is_i4 && has_y2) {
    WARP_PHASE(lane)
      if (lane == 0) {
        int in[16];
        load_block_coeffs(coeffs + 24 * 16, dq6[2], dq6[3], in);
        wht_block(in, ws.dc);
      }
    WARP_PHASE_END
  }

  // ---- phase 1: residuals of all 24 blocks, one lane per block
  // No block of the macroblock with more than a DC level (2-bit codes all 0 or 1): the inverse transform of a lone DC is
  // the constant (dc + 4) >> 3 (TransformDC_C, dsp/dec.c:105-114, equal to TransformOne on such a block), and the lanes
  // move in step, so the short form only pays when the whole macroblock can take it.
  const int dc_only = ((nzy & 0xaaaaaaaau) | (nzuv & 0xaaaau)) == 0;
  if (any_coef && dc_only) {
    WARP_PHASE(lane)
      if (lane < 24) {
        const int blk = lane;
        const int luma = blk < 16;
        const uint32_t code = luma ? ((nzy >> (30 - 2 * blk)) & 3u)
                                   : ((nzuv >> (8 * ((blk - 16) >> 2) + 6 - 2 * ((blk - 16) & 3))) & 3u);
        const int i16_luma = luma && !is_i4;
        int in0 = 0;
        if (code != 0) in0 = (int)(int16_t)((int)coeffs[blk * 16] * (int)(luma ? dq6[0] : dq6[4]));
        if (i16_luma) in0 = has_y2 ? (int)ws.dc[blk] : 0;
        const uint32_t r2 = (uint32_t)(uint16_t)((in0 + 4) >> 3) * 0x00010001u;
        uint4 v; v.x = r2; v.y = r2; v.z = r2; v.w = r2;
        uint4* dst = (uint4*)(ws.res + blk * 16);
        dst[0] = v; dst[1] = v;
        if (in0 != 0) ws.nz = 1;   // a lone DC level whose int16 product is 0 leaves the block uncoded
      }
    WARP_PHASE_END
  } else if (any_coef) {
    WARP_PHASE(lane)
      if (lane < 24) {
        const int blk = lane;
        const int luma = blk < 16;
        const uint32_t code = luma ? ((nzy >> (30 - 2 * blk)) & 3u)
                                   : ((nzuv >> (8 * ((blk - 16) >> 2) + 6 - 2 * ((blk - 16) & 3))) & 3u);
        const int i16_luma = luma && !is_i4;
        const int wht_dc = (i16_luma && has_y2) ? (int)ws.dc[blk] : 0;
        int out[16];
        int coded = 0;
        if (code != 0 || wht_dc != 0) {
          int in[16];
          if (code != 0) {
            load_block_coeffs(coeffs + blk * 16, luma ? dq6[0] : dq6[4], luma ? dq6[1] : dq6[5], in);
          } else {
            VP8_UNROLL
            for (int k = 0; k < 16; ++k) in[k] = 0;
          }
          if (i16_luma) in[0] = wht_dc;
          coded = (code >= 2) || in[0] != 0;   // a lone DC level whose int16 product is 0 leaves the block uncoded
          idct_block(in, out);
        } else {
          VP8_UNROLL
          for (int k = 0; k < 16; ++k) out[k] = 0;
        }
        uint4 lo, hi;
        lo.x = (uint32_t)(uint16_t)out[0] | ((uint32_t)out[1] << 16); lo.y = (uint32_t)(uint16_t)out[2] | ((uint32_t)out[3] << 16);
        lo.z = (uint32_t)(uint16_t)out[4] | ((uint32_t)out[5] << 16); lo.w = (uint32_t)(uint16_t)out[6] | ((uint32_t)out[7] << 16);
        hi.x = (uint32_t)(uint16_t)out[8] | ((uint32_t)out[9] << 16); hi.y = (uint32_t)(uint16_t)out[10] | ((uint32_t)out[11] << 16);
        hi.z = (uint32_t)(uint16_t)out[12] | ((uint32_t)out[13] << 16); hi.w = (uint32_t)(uint16_t)out[14] | ((uint32_t)out[15] << 16);
        uint4* dst = (uint4*)(ws.res + blk * 16);
        dst[0] = lo; dst[1] = hi;
        if (coded) ws.nz = 1;
      }
    WARP_PHASE_END
  }

  // ---- phase 2: luma
  if (is_i4) {
    // Ten anti-diagonals bx + 2*by = step, two sub-blocks on each: lanes 0-15 take (bx0, by0), lanes 16-31 the one a row
    // down and two columns left (n + 2; 120 bytes further on in the tile), where there is one. Both read only pixels
    // outside themselves (finished in earlier steps), straight from the tile, and write only their own 16: one phase per
    // step. Unrolled, everything but the lane's half is a constant of the step.
    VP8_UNROLL
    for (int step = 0; step < 10; ++step) {
      const int by0 = (step > 2) ? (step - 2) >> 1 : 0, bx0 = step - 2 * by0;
      const int n0 = 4 * by0 + bx0;
      const int second = (by0 < 3) && (bx0 >= 2);           // lanes 16-31 have a sub-block on this step
      const int origin = (by0 * 4 + 1) * 32 + bx0 * 4 + 4;  // of sub-block n0 in the tile
      WARP_PHASE(lane)
        const int half = lane >> 4, l = lane & 15;
        if (half == 0 || second) {
          // modes of sub-blocks n0 and n0 + 2: the 64-bit mode word moved down by two nibbles for the second half
          const uint32_t ma = half ? (m0 >> 8) | (m1 << 24) : m0, mb = half ? (m1 >> 8) : m1;
          const int mode = (int)(((n0 < 8) ? (ma >> (4 * (n0 & 7))) : (mb >> (4 * (n0 & 7)))) & 15);
          uint8_t* const t = ws.y + 120 * half + origin;
          int p;
          if (mode == M_DC) {
            p = (sum4(*(const uint32_t*)(t - 32)) + t[-1] + t[31] + t[63] + t[95] + 4) >> 3;
          } else {
            const uint32_t e = cx.pred4[mode * 16 + l];
            const uint8_t* const tb = t - 64;
            const int a = tb[e & 255u], b = tb[(e >> 8) & 255u], c = tb[e >> 16];
            p = (mode == M_TM) ? clip8i(a + b - c) : (a + 2 * b + c + 2) >> 2;
          }
          if (any_coef) p = clip8i(p + (int)ws.res[(n0 + 2 * half) * 16 + l]);
          t[(l >> 2) * 32 + (l & 3)] = (uint8_t)p;
        }
      WARP_PHASE_END
    }
  } else {
    const int mode = check_mode(mx, my, (int)(m0 & 15));
    uint8_t* const t = ws.y + 32 + 4;
    WARP_PHASE(lane)
      const int dc = dc_big(t, mode, 16);
      const int r = lane >> 1, c0 = (lane & 1) * 8;
      int v[8];
      pred_big_run<8>(t, mode, dc, c0, r, v);
      if (any_coef) {
        const int16_t* ra = ws.res + ((r >> 2) * 4 + (c0 >> 2)) * 16 + (r & 3) * 4;   // blocks (c0>>2) and (c0>>2)+1 of block row r>>2
        for (int k = 0; k < 4; ++k) { v[k] = clip8i(v[k] + ra[k]); v[4 + k] = clip8i(v[4 + k] + ra[16 + k]); }
      }
      // reads above touch only the border row/column, writes only the block interior: no hazard in a phase
      for (int k = 0; k < 8; ++k) t[r * 32 + c0 + k] = (uint8_t)v[k];
    WARP_PHASE_END
  }

  // ---- phase 3: chroma (lanes 0-15 U, 16-31 V)
  {
    const int mode = check_mode(mx, my, (int)((w >> MBW_UVMODE_SHIFT) & 3));
    WARP_PHASE(lane)
      const int ch = lane >> 4;
      uint8_t* const t = ws.uv + 32 + 4 + 16 * ch;
      const int dc = dc_big(t, mode, 8);
      const int r = (lane & 15) >> 1, c0 = (lane & 1) * 4;
      int v[4];
      pred_big_run<4>(t, mode, dc, c0, r, v);
      if (any_coef) {
        const int16_t* ra = ws.res + (16 + 4 * ch + (r >> 2) * 2 + (c0 >> 2)) * 16 + (r & 3) * 4;
        for (int k = 0; k < 4; ++k) v[k] = clip8i(v[k] + ra[k]);
      }
      for (int k = 0; k < 4; ++k) t[r * 32 + c0 + k] = (uint8_t)v[k];
    WARP_PHASE_END
  }

  if (next_ntok != 0) recon_fetch_tokens(rt, next_toks, next_ntok);
  // ---- phase 4: tile -> HBM planes, neighbour context for the macroblocks to the right and below, MbInfo
  WARP_PHASE(lane)
    if (lane < 16) {
      const uint32_t* row = (const uint32_t*)(ws.y + (lane + 1) * 32 + 4);
      uint4 v; v.x = row[0]; v.y = row[1]; v.z = row[2]; v.w = row[3];
      *(uint4*)(yplane + (size_t)(16 * my + lane) * ys + 16 * mx) = v;
    } else {
      const int ch = (lane - 16) >> 3, r = lane & 7;
      const uint32_t* row = (const uint32_t*)(ws.uv + (r + 1) * 32 + 4 + 16 * ch);
      uint2 v; v.x = row[0]; v.y = row[1];
      *(uint2*)((ch ? vplane : uplane) + (size_t)(8 * my + r) * uvs + 8 * mx) = v;
    }
    // right column -> left context of the next macroblock of the row; bottom row -> top context of the macroblock below
    cx.left[36 * my + lane] = lane < 16 ? ws.y[(lane + 1) * 32 + 19] : ws.uv[((lane & 7) + 1) * 32 + (lane < 24 ? 11 : 27)];
    if (lane < 8) {
      const uint32_t* src = lane < 4 ? (const uint32_t*)(ws.y + 16 * 32 + 4 + 4 * lane)
                                     : (const uint32_t*)(ws.uv + 8 * 32 + 4 + 4 * ((lane - 4) & 1) + 16 * ((lane - 4) >> 1));
      ((uint32_t*)(cx.top + 32 * mx))[lane] = *src;
    }
    if (lane == 0) {
      cx.left[36 * my + 32] = ws.y[19];
      cx.left[36 * my + 33] = ws.uv[11];
      cx.left[36 * my + 34] = ws.uv[27];
      info[3] = (is_i4 || ws.nz != 0) ? (w | MBW_INNER) : (w & ~MBW_INNER);
    }
  WARP_PHASE_END
}

VP8_PFN void recon_macroblock(ReconWs& ws, const ReconCtx& cx, int mx, int my, int mb_w, uint32_t* info,
                              const int16_t* coeffs, const int16_t* dq6, uint8_t* yplane, uint8_t* uplane, uint8_t* vplane,
                              const uint32_t* toks = nullptr, uint32_t ntok = 0) {
  uint4 iw; iw.x = info[0]; iw.y = info[1]; iw.z = info[2]; iw.w = info[3];
  ReconTok rt;
  if (toks != nullptr) recon_fetch_tokens(rt, toks, ntok);
  recon_macroblock(ws, cx, mx, my, mb_w, iw, info, coeffs, dq6, yplane, uplane, vplane, toks, ntok, rt, nullptr, 0);
}

// =========================================================================================================
// In-loop deblocking filter. Thresholds as in the reference (clip tables of dec_clip_tables.c written as
// clamps). One lane filters one line of 8 pixels across an edge; `p` addresses the first pixel after the
// edge (q0), `step` is the distance between pixels across the edge, inside the warp's shared-memory tile.
VP8_PFN int iabs_(int v) { return v < 0 ? -v : v; }
VP8_PFN int sclip1_(int v) { return v < -128 ? -128 : v > 127 ? 127 : v; }
VP8_PFN int sclip2_(int v) { return v < -16 ? -16 : v > 15 ? 15 : v; }

VP8_PFN void lf_filter2(uint8_t* p, int step, int p1, int p0, int q0, int q1) {   // DoFilter2_C
  const int a = 3 * (q0 - p0) + sclip1_(p1 - q1);
  const int a1 = sclip2_((a + 4) >> 3), a2 = sclip2_((a + 3) >> 3);
  p[-step] = (uint8_t)clip8i(p0 + a2);
  p[0] = (uint8_t)clip8i(q0 - a1);
}

// kind 0: simple filter; 1: normal filter on a macroblock edge; 2: normal filter on an inner edge.
VP8_PFN void lf_line(uint8_t* p, int step, int kind, int thresh, int ithresh, int hev_t) {
  const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
  const int t2 = 2 * thresh + 1;
  if (4 * iabs_(p0 - q0) + iabs_(p1 - q1) > t2) return;
  if (kind == 0) { lf_filter2(p, step, p1, p0, q0, q1); return; }
  const int p3 = p[-4 * step], p2 = p[-3 * step], q2 = p[2 * step], q3 = p[3 * step];
  if (iabs_(p3 - p2) > ithresh || iabs_(p2 - p1) > ithresh || iabs_(p1 - p0) > ithresh ||
      iabs_(q3 - q2) > ithresh || iabs_(q2 - q1) > ithresh || iabs_(q1 - q0) > ithresh) return;
  if (iabs_(p1 - p0) > hev_t || iabs_(q1 - q0) > hev_t) { lf_filter2(p, step, p1, p0, q0, q1); return; }
  if (kind == 1) {   // DoFilter6_C
    const int a = sclip1_(3 * (q0 - p0) + sclip1_(p1 - q1));
    const int a1 = (27 * a + 63) >> 7, a2 = (18 * a + 63) >> 7, a3 = (9 * a + 63) >> 7;
    p[-3 * step] = (uint8_t)clip8i(p2 + a3); p[-2 * step] = (uint8_t)clip8i(p1 + a2); p[-step] = (uint8_t)clip8i(p0 + a1);
    p[0] = (uint8_t)clip8i(q0 - a1); p[step] = (uint8_t)clip8i(q1 - a2); p[2 * step] = (uint8_t)clip8i(q2 - a3);
  } else {           // DoFilter4_C
    const int a = 3 * (q0 - p0);
    const int a1 = sclip2_((a + 4) >> 3), a2 = sclip2_((a + 3) >> 3), a3 = (a1 + 1) >> 1;
    p[-2 * step] = (uint8_t)clip8i(p1 + a3); p[-step] = (uint8_t)clip8i(p0 + a2);
    p[0] = (uint8_t)clip8i(q0 - a1); p[step] = (uint8_t)clip8i(q1 - a3);
  }
}

// Filter workspace of one warp: luma rows -4..15 / cols -4..15 at y[(r + 4) * FW_STRIDE + c + 4]; chroma rows -4..7,
// U at uv[(r + 4) * FW_STRIDE + c + 4], V at uv[(r + 4) * FW_STRIDE + c + 20].
// Rows are 36 bytes apart, not 32: the sixteen lines of a vertical edge then read sixteen different banks (at 32 the rows
// r, r + 4, r + 8, r + 12 share one: ncu counted 0.5 G conflicts per 1024 full-HD images, short-scoreboard stalls on top).
#define FW_STRIDE 36
struct FilterWs {
  uint8_t y[20 * FW_STRIDE];
  uint8_t uv[12 * FW_STRIDE];
};

// Filters macroblock (mx, my) in place in the image's HBM planes, in the reference's edge order
// (DoFilter, frame_dec.c:203-250). fs = {limit, ilevel, inner(unused), hev_thresh} for this macroblock's
// segment / block type; `inner` = MbInfo filter-inner bit; filter_type 1 simple, 2 normal.
// filter_tile: the edges of one macroblock on the warp's tile (already loaded), then the tile back into the planes;
// filter_macroblock (below) loads the tile first.
VP8_PFN void filter_tile(FilterWs& ws, int mx, int my, int mb_w, int filter_type, const uint8_t* fs, int inner,
                         uint8_t* yplane, uint8_t* uplane, uint8_t* vplane) {
  const int limit = fs[0], ilevel = fs[1], hev_t = fs[3];
  const int ys = 16 * mb_w, uvs = 8 * mb_w;
  const int normal = (filter_type == 2);
  if (!normal) {
    // Simple filter (luma only): a line reads p1 p0 q0 q1 and writes p0 q0, i.e. columns 4k-2 .. 4k+1 of edge k, so the
    // four edges of one direction touch disjoint pixels and the reference's edge-after-edge order (frame_dec.c:216-231)
    // equals all of them at once: two edges x 16 lines per warp pass instead of one. Direction order stays.
    for (int dir = 0; dir < 2; ++dir) {
      const int outer = dir == 0 ? (mx > 0) : (my > 0);
      for (int half = 0; half < 2; ++half) {
        if (half == 0 ? (outer || inner) : inner) {
          WARP_PHASE(lane)
            const int k = 2 * half + (lane >> 4), i = lane & 15;
            if (k == 0 ? outer : inner) {
              uint8_t* p = dir == 0 ? ws.y + (4 + i) * FW_STRIDE + 4 + 4 * k : ws.y + (4 + 4 * k) * FW_STRIDE + 4 + i;
              lf_line(p, dir == 0 ? 1 : FW_STRIDE, 0, (k == 0) ? limit + 4 : limit, ilevel, hev_t);
            }
          WARP_PHASE_END
        }
      }
    }
  }
  // ---- vertical edges (filtering across columns): macroblock edge, then the three inner edges
  for (int k = 0; normal && k < 4; ++k) {
    if (k == 0 ? (mx > 0) : inner) {
      const int thresh = (k == 0) ? limit + 4 : limit;
      const int kind = normal ? ((k == 0) ? 1 : 2) : 0;
      WARP_PHASE(lane)
        if (lane < 16) {
          lf_line(ws.y + (4 + lane) * FW_STRIDE + 4 + 4 * k, 1, kind, thresh, ilevel, hev_t);
        } else if (normal && k < 2) {
          const int r = lane & 7, ch = (lane - 16) >> 3;
          lf_line(ws.uv + (4 + r) * FW_STRIDE + 4 + 16 * ch + 4 * k, 1, kind, thresh, ilevel, hev_t);
        }
      WARP_PHASE_END
    }
  }
  // ---- horizontal edges (filtering across rows)
  for (int k = 0; normal && k < 4; ++k) {
    if (k == 0 ? (my > 0) : inner) {
      const int thresh = (k == 0) ? limit + 4 : limit;
      const int kind = normal ? ((k == 0) ? 1 : 2) : 0;
      WARP_PHASE(lane)
        if (lane < 16) {
          lf_line(ws.y + (4 + 4 * k) * FW_STRIDE + 4 + lane, FW_STRIDE, kind, thresh, ilevel, hev_t);
        } else if (normal && k < 2) {
          const int c = lane & 7, ch = (lane - 16) >> 3;
          lf_line(ws.uv + (4 + 4 * k) * FW_STRIDE + 4 + 16 * ch + c, FW_STRIDE, kind, thresh, ilevel, hev_t);
        }
      WARP_PHASE_END
    }
  }
  // ---- store (everything that was loaded: no other macroblock of the same wavefront step touches it)
  WARP_PHASE(lane)
    if (lane < 20) {
      const int gy = 16 * my - 4 + lane;
      if (gy >= 0) {
        uint8_t* dst = yplane + (size_t)gy * ys + 16 * mx;
        const uint32_t* s = (const uint32_t*)(ws.y + lane * FW_STRIDE + 4);
        uint4 v; v.x = s[0]; v.y = s[1]; v.z = s[2]; v.w = s[3];
        *(uint4*)dst = v;
        if (mx > 0) *(uint32_t*)(dst - 4) = *(const uint32_t*)(ws.y + lane * FW_STRIDE);
      }
    } else if (normal) {
      const int r = lane - 20;
      const int gy = 8 * my - 4 + r;
      if (gy >= 0) {
        uint8_t* du = uplane + (size_t)gy * uvs + 8 * mx;
        uint8_t* dv = vplane + (size_t)gy * uvs + 8 * mx;
        const uint32_t* s = (const uint32_t*)(ws.uv + r * FW_STRIDE);
        uint2 a, b; a.x = s[1]; a.y = s[2]; b.x = s[5]; b.y = s[6];
        *(uint2*)du = a; *(uint2*)dv = b;
        if (mx > 0) { *(uint32_t*)(du - 4) = s[0]; *(uint32_t*)(dv - 4) = s[4]; }
      }
    }
  WARP_PHASE_END
}

VP8_PFN void filter_macroblock(FilterWs& ws, int mx, int my, int mb_w, int filter_type, const uint8_t* fs, int inner,
                               uint8_t* yplane, uint8_t* uplane, uint8_t* vplane) {
  const int ys = 16 * mb_w, uvs = 8 * mb_w;
  const int normal = (filter_type == 2);
  if (fs[0] == 0) return;
  // ---- load
  WARP_PHASE(lane)
    if (lane < 20) {
      const int gy = 16 * my - 4 + lane;
      if (gy >= 0) {
        const uint8_t* src = yplane + (size_t)gy * ys + 16 * mx;
        const uint4 v = *(const uint4*)src;
        uint32_t* d = (uint32_t*)(ws.y + lane * FW_STRIDE + 4);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        if (mx > 0) *(uint32_t*)(ws.y + lane * FW_STRIDE) = *(const uint32_t*)(src - 4);
      }
    } else if (normal) {
      const int r = lane - 20;
      const int gy = 8 * my - 4 + r;
      if (gy >= 0) {
        const uint8_t* su = uplane + (size_t)gy * uvs + 8 * mx;
        const uint8_t* sv = vplane + (size_t)gy * uvs + 8 * mx;
        const uint2 a = *(const uint2*)su, b = *(const uint2*)sv;
        uint32_t* d = (uint32_t*)(ws.uv + r * FW_STRIDE);
        d[1] = a.x; d[2] = a.y; d[5] = b.x; d[6] = b.y;
        if (mx > 0) { d[0] = *(const uint32_t*)(su - 4); d[4] = *(const uint32_t*)(sv - 4); }
      }
    }
  WARP_PHASE_END
  filter_tile(ws, mx, my, mb_w, filter_type, fs, inner, yplane, uplane, vplane);
}

// The same for a warp that walks a macroblock row from left to right (k_loop_filter): a macroblock's own columns are
// fetched into registers while the macroblock before it is filtered (filter_fetch; its rows above must be final by then),
// put into the tile (filter_fill), and the tile's four right-most columns -- filtered -- stay behind as the next
// macroblock's left columns (filter_shift) instead of coming back from the planes.
struct FilterPre { uint32_t v[4][VP8_LANE_SLOTS]; };

VP8_PFN void filter_fetch(FilterPre& pre, int mx, int my, int mb_w, int normal, const uint8_t* yplane, const uint8_t* uplane,
                          const uint8_t* vplane) {
  const int ys = 16 * mb_w, uvs = 8 * mb_w;
  WARP_PHASE(lane)
    uint32_t a = 0, b = 0, c = 0, d = 0;
    if (lane < 20) {
      const int gy = 16 * my - 4 + lane;
      if (gy >= 0) { const uint4 v = *(const uint4*)(yplane + (size_t)gy * ys + 16 * mx); a = v.x; b = v.y; c = v.z; d = v.w; }
    } else if (normal) {
      const int gy = 8 * my - 4 + (lane - 20);
      if (gy >= 0) {
        const uint2 p = *(const uint2*)(uplane + (size_t)gy * uvs + 8 * mx), q = *(const uint2*)(vplane + (size_t)gy * uvs + 8 * mx);
        a = p.x; b = p.y; c = q.x; d = q.y;
      }
    }
    pre.v[0][VP8_LANE_SLOT(lane)] = a; pre.v[1][VP8_LANE_SLOT(lane)] = b; pre.v[2][VP8_LANE_SLOT(lane)] = c; pre.v[3][VP8_LANE_SLOT(lane)] = d;
  WARP_PHASE_END
}

VP8_PFN void filter_fill(FilterWs& ws, const FilterPre& pre, int normal) {
  WARP_PHASE(lane)
    const uint32_t a = pre.v[0][VP8_LANE_SLOT(lane)], b = pre.v[1][VP8_LANE_SLOT(lane)], c = pre.v[2][VP8_LANE_SLOT(lane)], d = pre.v[3][VP8_LANE_SLOT(lane)];
    if (lane < 20) {
      uint32_t* t = (uint32_t*)(ws.y + lane * FW_STRIDE + 4);
      t[0] = a; t[1] = b; t[2] = c; t[3] = d;
    } else if (normal) {
      uint32_t* t = (uint32_t*)(ws.uv + (lane - 20) * FW_STRIDE);
      t[1] = a; t[2] = b; t[5] = c; t[6] = d;
    }
  WARP_PHASE_END
}

VP8_PFN void filter_shift(FilterWs& ws, int normal) {
  WARP_PHASE(lane)
    if (lane < 20) {
      uint32_t* t = (uint32_t*)(ws.y + lane * FW_STRIDE);
      t[0] = t[4];                 // luma columns 12-15 -> columns -4..-1
    } else if (normal) {
      uint32_t* t = (uint32_t*)(ws.uv + (lane - 20) * FW_STRIDE);
      t[0] = t[2]; t[4] = t[6];    // U and V columns 4-7 -> columns -4..-1
    }
  WARP_PHASE_END
}

// =========================================================================================================
// Output stage. One thread converts four horizontally adjacent pixels.
VP8_PFN int yuv_clip6(int v) { return ((v & ~16383) == 0) ? (v >> 6) : (v < 0) ? 0 : 255; }

VP8_PFN void yuv_to_rgb(int y, int u, int v, int* r, int* g, int* b) {   // yuv.h:59-77
  const int yy = (y * 19077) >> 8;
  *r = yuv_clip6(yy + ((v * 26149) >> 8) - 14234);
  *g = yuv_clip6(yy - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708);
  *b = yuv_clip6(yy + ((u * 33050) >> 8) - 17685);
}

// Fancy-upsampled chroma of pixel (i, j) of a w x h picture (UPSAMPLE_FUNC, upsampling.c:37-93 applied the way
// EmitFancyRGB does over a whole frame, io_dec.c:57-109). `pl` = one chroma plane, stride uvs.
// near/far = the chroma rows that weigh 3/4 and 1/4 for this luma row.
VP8_PFN int fancy_chroma(const uint8_t* near, const uint8_t* far, int i, int w) {
  if (i == 0) return (3 * near[0] + far[0] + 2) >> 2;
  const int x = (i + 1) >> 1;
  if (!(w & 1) && i == w - 1) return (3 * near[x - 1] + far[x - 1] + 2) >> 2;
  const int nl = near[x - 1], nr = near[x], fl = far[x - 1], fr = far[x];
  const int avg = nl + nr + fl + fr + 8;
  return (i & 1) ? ((((avg + 2 * (nr + fl)) >> 3) + nl) >> 1) : ((((avg + 2 * (nl + fr)) >> 3) + nr) >> 1);
}

// a = the pixel's alpha (0xff for opaque images). Premultiplied modes scale the colour by alpha exactly like
// ApplyAlphaMultiply_C (alpha_processing.c:214-241): untouched when a == 0xff, else (c * a * 32897) >> 23.
VP8_PFN uint32_t pack_pixel4(int csp, int r, int g, int b, int a) {   // little-endian byte order in memory
  if (csp >= 7 && a != 0xff) {
    const uint32_t m = (uint32_t)a * 32897u;
    r = (int)(((uint32_t)r * m) >> 23); g = (int)(((uint32_t)g * m) >> 23); b = (int)(((uint32_t)b * m) >> 23);
  }
  switch (csp) {
    case 1: case 7: return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24);   // RGBA / rgbA
    case 3: case 8: return (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16) | ((uint32_t)a << 24);   // BGRA / bgrA
    default: return (uint32_t)a | ((uint32_t)r << 8) | ((uint32_t)g << 16) | ((uint32_t)b << 24);           // ARGB / Argb
  }
}

// 16-bit colourspaces (yuv.h:93-123, bytes in memory order rg | gb resp. rg | ba; WEBP_SWAP_16BIT_CSP is off in the
// reference build). a = 8-bit alpha or 0xff. rgbA_4444 scales by the 4-bit alpha like ApplyAlphaMultiply4444_C
// (alpha_processing.c:244-283): nibbles replicated to bytes, times a * 0x1111, >> 16; alpha 15 leaves the pixel alone.
VP8_PFN uint32_t pack_pixel2(int csp, int r, int g, int b, int a) {
  if (csp == 6) return (uint32_t)((r & 0xf8) | (g >> 5)) | ((uint32_t)(((g << 3) & 0xe0) | (b >> 3)) << 8);   // RGB_565
  const uint32_t a4 = (uint32_t)a >> 4;
  uint32_t rg = (uint32_t)((r & 0xf0) | (g >> 4)), ba = (uint32_t)(b & 0xf0) | a4;
  if (csp == 10 && a4 != 15) {
    const uint32_t m = a4 * 0x1111u;
    const uint32_t r2 = ((((rg & 0xf0) | (rg >> 4)) * m) >> 16) & 0xff;
    const uint32_t g2 = ((((rg & 0x0f) | ((rg << 4) & 0xf0)) * m) >> 16) & 0xff;
    const uint32_t b2 = ((((ba & 0xf0) | (ba >> 4)) * m) >> 16) & 0xff;
    rg = (r2 & 0xf0) | ((g2 >> 4) & 0x0f);
    ba = (b2 & 0xf0) | a4;
  }
  return rg | (ba << 8);
}

// Pixels 4*q .. 4*q+3 of output row j of image `im`. yuv = the image's padded planes.
VP8_PFN void emit_rgb_quad(const ImgDesc& im, const uint8_t* yplane, const uint8_t* uplane, const uint8_t* vplane,
                           const uint8_t* alpha /* window origin inside the frame-wide plane, or NULL */, uint8_t* out, int q, int j) {
  const int w = im.out_w, h = im.out_h;
  const int ys = 16 * im.mb_w, uvs = 8 * im.mb_w;
  const int uvh = (h + 1) >> 1;
  const int csp = im.csp;
  const int fancy = !(im.flags & VP8B_FLAG_NO_FANCY);
  int rn, rf;   // near / far chroma rows
  if (!fancy) { rn = rf = j >> 1; }
  else if (j == 0) { rn = rf = 0; }
  else if (j & 1) { rn = (j - 1) >> 1; rf = rn + 1 < uvh ? rn + 1 : uvh - 1; }
  else { rn = j >> 1; rf = rn - 1; }
  const uint8_t* yrow = yplane + (size_t)j * ys;
  const uint8_t* un = uplane + (size_t)rn * uvs; const uint8_t* uf = uplane + (size_t)rf * uvs;
  const uint8_t* vn = vplane + (size_t)rn * uvs; const uint8_t* vf = vplane + (size_t)rf * uvs;
  uint8_t* orow = out + (size_t)((im.flags & VP8B_FLAG_FLIP) ? h - 1 - j : j) * im.out_stride;
  const int i0 = 4 * q;
  const int n = (w - i0 < 4) ? w - i0 : 4;
  const int bpp = (csp == 0 || csp == 2) ? 3 : (csp == 5 || csp == 6 || csp == 10) ? 2 : 4;
  uint32_t px[4];
  for (int k = 0; k < n; ++k) {
    const int i = i0 + k;
    int u, v, r, g, b;
    if (fancy) { u = fancy_chroma(un, uf, i, w); v = fancy_chroma(vn, vf, i, w); }
    else { u = un[i >> 1]; v = vn[i >> 1]; }
    yuv_to_rgb(yrow[i], u, v, &r, &g, &b);
    if (bpp == 4) {
      px[k] = pack_pixel4(csp, r, g, b, alpha ? (int)alpha[(size_t)j * im.width + i] : 0xff);
    } else if (bpp == 2) {
      const uint32_t p2 = pack_pixel2(csp, r, g, b, alpha ? (int)alpha[(size_t)j * im.width + i] : 0xff);
      orow[2 * i] = (uint8_t)p2; orow[2 * i + 1] = (uint8_t)(p2 >> 8);
    } else {
      uint8_t* o = orow + 3 * i;
      if (csp == 0) { o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b; }
      else { o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r; }
    }
  }
  if (bpp == 4) {
    uint8_t* o = orow + 4 * i0;
    if (n == 4 && (((uintptr_t)o) & 15) == 0) {
      uint4 v4; v4.x = px[0]; v4.y = px[1]; v4.z = px[2]; v4.w = px[3];
      *(uint4*)o = v4;
    } else if ((((uintptr_t)o) & 3) == 0) {
      for (int k = 0; k < n; ++k) ((uint32_t*)o)[k] = px[k];
    } else {
      for (int k = 0; k < n; ++k) { o[4 * k] = (uint8_t)px[k]; o[4 * k + 1] = (uint8_t)(px[k] >> 8); o[4 * k + 2] = (uint8_t)(px[k] >> 16); o[4 * k + 3] = (uint8_t)(px[k] >> 24); }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fast path of the output stage for the 4-byte RGB family with fancy upsampling: one thread converts 8 pixels
// of the two output rows 2t-1 and 2t, which share the chroma rows t-1 and t (UPSAMPLE_FUNC processes exactly
// these row pairs, upsampling.c:37-93; the first and, for even heights, the last row stand alone,
// io_dec.c:72-109). Edge replication = clamping the chroma column / row (verified against the closed form in
// fancy_chroma()). Loads are 8-byte (Y) and 4-byte (chroma) words, stores 16-byte.
#if defined(__CUDACC__) && !defined(VP8_EMU)
VP8_PFN int sat_u8(int v) { int r; asm("cvt.sat.u8.s32 %0, %1;" : "=r"(r) : "r"(v)); return r; }
#else
VP8_PFN int sat_u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
#endif

VP8_PFN uint32_t yuv_to_px4(int csp, int y, int u, int v, int a) {
  const int yy = (y * 19077) >> 8;
  const int r = sat_u8((yy + ((v * 26149) >> 8) - 14234) >> 6);
  const int g = sat_u8((yy - ((u * 6419) >> 8) - ((v * 13320) >> 8) + 8708) >> 6);
  const int b = sat_u8((yy + ((u * 33050) >> 8) - 17685) >> 6);
  return pack_pixel4(csp, r, g, b, a);
}

// Six chroma samples of plane row `row` for output pixels 8q..8q+7: columns 4q-1 .. 4q+4, clamped to [0, uvw).
VP8_PFN void load_chroma6(const uint8_t* row, int q, int uvw, int c[6]) {
  const int x0 = 4 * q;
  if (x0 + 4 < uvw) {   // interior: one aligned word + two neighbours
    const uint32_t wd = *(const uint32_t*)(row + x0);
    c[0] = row[x0 > 0 ? x0 - 1 : 0];
    c[1] = wd & 0xff; c[2] = (wd >> 8) & 0xff; c[3] = (wd >> 16) & 0xff; c[4] = wd >> 24;
    c[5] = row[x0 + 4];
  } else {
    for (int k = 0; k < 6; ++k) { int x = x0 - 1 + k; x = x < 0 ? 0 : x >= uvw ? uvw - 1 : x; c[k] = row[x]; }
  }
}

// Upsampled chroma of the 8 pixels for the row whose nearer chroma row is `n` (the other is `f`).
VP8_PFN void upsample8(const int n[6], const int f[6], int out[8]) {
VP8_UNROLL
  for (int k = 0; k < 5; ++k) {   // column pair (k, k+1) of the six: pixels 2k-1 (odd member) and 2k (even member)
    const int nl = n[k], nr = n[k + 1], fl = f[k], fr = f[k + 1];
    const int s = nl + nr + fl + fr + 8;
    if (k > 0) out[2 * k - 1] = (((s + 2 * (nr + fl)) >> 3) + nl) >> 1;
    if (k < 4) out[2 * k] = (((s + 2 * (nl + fr)) >> 3) + nr) >> 1;
  }
}

// Opaque pixel of the 4-byte family from y << 16 (as extracted from the packed word) and 8-bit u, v; the same
// arithmetic as yuv_to_rgb() with the products taken as high halves where the operand arrives pre-shifted, and the
// three clips + the byte packing done by two saturating pack conversions. ORDER: 0 = R,G,B,A  1 = B,G,R,A  2 = A,R,G,B.
#if defined(__CUDACC__) && !defined(VP8_EMU)
VP8_PFN uint32_t pack_sat2(int hi, int lo, uint32_t upper) {   // sat_u8(hi) << 8 | sat_u8(lo) | upper << 16
  uint32_t d; asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper)); return d;
}
VP8_PFN int mulhi_s(int a, int b) { return __mulhi(a, b); }
#else
VP8_PFN uint32_t pack_sat2(int hi, int lo, uint32_t upper) { return ((uint32_t)sat_u8(hi) << 8) | (uint32_t)sat_u8(lo) | (upper << 16); }
VP8_PFN int mulhi_s(int a, int b) { return (int)(((int64_t)a * (int64_t)b) >> 32); }
#endif

template <int ORDER>
VP8_PFN uint32_t yuv_to_px4_opaque(int y16, int u, int v) {
  const int yy = mulhi_s(y16, 19077 << 8);                       // (y * 19077) >> 8
  const int r = (yy + ((v * 26149) >> 8) - 14234) >> 6;
  const int g = (yy + (8708 - ((u * 6419) >> 8)) - ((v * 13320) >> 8)) >> 6;
  const int b = (yy + ((u * 33050) >> 8) - 17685) >> 6;
  if (ORDER == 0) return pack_sat2(g, r, pack_sat2(255, b, 0));
  if (ORDER == 1) return pack_sat2(g, b, pack_sat2(255, r, 0));
  return pack_sat2(r, 255, pack_sat2(b, g, 0));
}

// emit_rgba_pair8() for images without an alpha plane (premultiplied modes equal the plain ones there), aligned rows.
template <int ORDER>
VP8_PFN void emit_opaque_pair8(const ImgDesc& im, const uint8_t* yplane, const uint8_t* uplane, const uint8_t* vplane,
                               uint8_t* out, int q, int t) {
  const int w = im.out_w, h = im.out_h;
  const int ys = 16 * im.mb_w, uvs = 8 * im.mb_w;
  const int uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
  const int ra = t > 0 ? t - 1 : 0, rb = t < uvh ? t : uvh - 1;
  int ua[6], ub[6], va[6], vb[6];
  load_chroma6(uplane + (size_t)ra * uvs, q, uvw, ua);
  load_chroma6(uplane + (size_t)rb * uvs, q, uvw, ub);
  load_chroma6(vplane + (size_t)ra * uvs, q, uvw, va);
  load_chroma6(vplane + (size_t)rb * uvs, q, uvw, vb);
  const int i0 = 8 * q;
VP8_UNROLL
  for (int half = 0; half < 2; ++half) {
    const int j = 2 * t - 1 + half;
    if (j < 0 || j >= h) continue;
    int u8[8], v8[8];
    if (half == 0) { upsample8(ua, ub, u8); upsample8(va, vb, v8); }
    else { upsample8(ub, ua, u8); upsample8(vb, va, v8); }
    const uint2 yw = *(const uint2*)(yplane + (size_t)j * ys + i0);
    uint32_t px[8];
VP8_UNROLL
    for (int k = 0; k < 8; ++k) {
      const uint32_t wd = k < 4 ? yw.x : yw.y;
      const int sh = 8 * (k & 3);
      const int y16 = (int)(sh <= 16 ? (wd << (16 - sh)) & 0xff0000u : (wd >> 8) & 0xff0000u);
      px[k] = yuv_to_px4_opaque<ORDER>(y16, u8[k], v8[k]);
    }
    uint4* o = (uint4*)(out + (size_t)((im.flags & VP8B_FLAG_FLIP) ? h - 1 - j : j) * im.out_stride + 4 * (size_t)i0);
    uint4 a, b;
    a.x = px[0]; a.y = px[1]; a.z = px[2]; a.w = px[3];
    b.x = px[4]; b.y = px[5]; b.z = px[6]; b.w = px[7];
    o[0] = a; o[1] = b;
  }
}

// Which images take the fast path (host driver and kernel must agree on the work-item count).
VP8_PFN int emit_uses_pairs(int csp, int flags, int crop_x) {   // crop_x: the word loads need the window 8-pixel aligned
  return !(flags & VP8B_FLAG_NO_FANCY) && (crop_x & 7) == 0 && (csp == 1 || csp == 3 || csp == 4 || csp == 7 || csp == 8 || csp == 9);
}

// Pixels 8q..8q+7 of output rows 2t-1 and 2t.
VP8_PFN void emit_rgba_pair8(const ImgDesc& im, const uint8_t* yplane, const uint8_t* uplane, const uint8_t* vplane,
                             const uint8_t* alpha /* window origin inside the frame-wide plane, or NULL */, uint8_t* out, int q, int t) {
  const int w = im.out_w, h = im.out_h;
  const int ys = 16 * im.mb_w, uvs = 8 * im.mb_w;
  const int uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
  const int csp = im.csp;
  if (alpha == 0 && w - 8 * q >= 8 && ((((uintptr_t)out) | (uintptr_t)im.out_stride) & 15) == 0) {   // opaque, whole, 16-byte aligned
    if (csp == 1 || csp == 7) emit_opaque_pair8<0>(im, yplane, uplane, vplane, out, q, t);
    else if (csp == 3 || csp == 8) emit_opaque_pair8<1>(im, yplane, uplane, vplane, out, q, t);
    else emit_opaque_pair8<2>(im, yplane, uplane, vplane, out, q, t);
    return;
  }
  const int ra = t > 0 ? t - 1 : 0, rb = t < uvh ? t : uvh - 1;
  int ua[6], ub[6], va[6], vb[6];
  load_chroma6(uplane + (size_t)ra * uvs, q, uvw, ua);
  load_chroma6(uplane + (size_t)rb * uvs, q, uvw, ub);
  load_chroma6(vplane + (size_t)ra * uvs, q, uvw, va);
  load_chroma6(vplane + (size_t)rb * uvs, q, uvw, vb);
  const int i0 = 8 * q;
  const int n = (w - i0 < 8) ? w - i0 : 8;
VP8_UNROLL
  for (int half = 0; half < 2; ++half) {
    const int j = 2 * t - 1 + half;
    if (j < 0 || j >= h) continue;
    int u8[8], v8[8];
    if (half == 0) { upsample8(ua, ub, u8); upsample8(va, vb, v8); }
    else { upsample8(ub, ua, u8); upsample8(vb, va, v8); }
    const uint2 yw = *(const uint2*)(yplane + (size_t)j * ys + i0);
    uint32_t px[8];
VP8_UNROLL
    for (int k = 0; k < 8; ++k) {
      const int y = (int)(((k < 4 ? yw.x : yw.y) >> (8 * (k & 3))) & 0xff);
      const int a = (alpha != 0 && k < n) ? (int)alpha[(size_t)j * im.width + i0 + k] : 0xff;
      px[k] = yuv_to_px4(csp, y, u8[k], v8[k], a);
    }
    uint8_t* o = out + (size_t)((im.flags & VP8B_FLAG_FLIP) ? h - 1 - j : j) * im.out_stride + 4 * (size_t)i0;
    if (n == 8 && (((uintptr_t)o) & 15) == 0) {
      uint4 a, b;
      a.x = px[0]; a.y = px[1]; a.z = px[2]; a.w = px[3];
      b.x = px[4]; b.y = px[5]; b.z = px[6]; b.w = px[7];
      ((uint4*)o)[0] = a; ((uint4*)o)[1] = b;
    } else if ((((uintptr_t)o) & 3) == 0) {
VP8_UNROLL
      for (int k = 0; k < 8; ++k) if (k < n) ((uint32_t*)o)[k] = px[k];
    } else {
VP8_UNROLL
      for (int k = 0; k < 8; ++k) if (k < n) {
        o[4 * k] = (uint8_t)px[k]; o[4 * k + 1] = (uint8_t)(px[k] >> 8); o[4 * k + 2] = (uint8_t)(px[k] >> 16); o[4 * k + 3] = (uint8_t)(px[k] >> 24);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// options.dithering_strength (frame_dec.c:319-386): the chroma of macroblocks without AC chroma coefficients gets a
// pseudo-random offset per pixel, amplitude by the segment's quantiser. The generator (D. Knuth's subtractive one,
// src/utils/random_utils.[ch]) is ONE sequence per picture, consumed in raster order by the macroblocks that are
// dithered (128 numbers each, U then V), so a serial pass per image lays the offsets down first (dither_plan_image) and
// the loop-filter kernel adds them where the reference does: after a macroblock row has been filtered, before the row
// below filters across their common edge.
VP8_PTABLE uint32_t kDitherRandomTable[55] = {
  0x0de15230, 0x03b31886, 0x775faccb, 0x1c88626a, 0x68385c55, 0x14b3b828, 0x4a85fef8, 0x49ddb84b, 0x64fcf397, 0x5c550289,
  0x4a290000, 0x0d7ec1da, 0x5940b7ab, 0x5492577d, 0x4e19ca72, 0x38d38c69, 0x0c01ee65, 0x32a1755f, 0x5437f652, 0x5abb2c32,
  0x0faa57b1, 0x73f533e7, 0x685feeda, 0x7563cce2, 0x6e990e83, 0x4730a7ed, 0x4fc0d9c6, 0x496b153c, 0x4f1403fa, 0x541afb0c,
  0x73990b32, 0x26d7cb1c, 0x6fcc3706, 0x2cbb77d8, 0x75762f2a, 0x6425ccdd, 0x24b35461, 0x0a7d8715, 0x220414a8, 0x141ebf67,
  0x56b41583, 0x73e502e3, 0x44cab16f, 0x28264d42, 0x73baaefb, 0x0a50ebed, 0x1d6ab6fb, 0x0d3ad40b, 0x35db3b68, 0x2b081e83,
  0x77ce6b95, 0x5181e5f0, 0x78853bbc, 0x009f9494, 0x27e5ed3c
};
#define VP8B_MIN_DITHER_AMP 4

// One image, one thread. `tab` = 55 words of scratch (shared memory on the device). Marks the dithered macroblocks in
// MbInfo (MBW_DITHER) and writes their 128 offsets (what DitherCombine8x8_C adds, dsp/dec.c:698-711) to `plane`
// (128 int8 per macroblock of the image). Rows [0, rows), columns [tl_x, br_x) (VP8EnterCritical, frame_dec.c:571-596).
VP8_PFN void dither_plan_image(const ImgDesc& im, const FrameHdr* h, uint32_t* mbinfo, int8_t* plane, uint32_t* tab) {
  if (!(h->dither[0] | h->dither[1] | h->dither[2] | h->dither[3])) return;
  const int mb_w = im.mb_w, rows = h->rows;
  const int extra = (h->filter_type == 2) ? 8 : (h->filter_type == 1) ? 2 : 0;   // kFilterExtraRows
  int tl_x = (h->filter_type == 2) ? 0 : ((int)im.crop_x - extra) >> 4;
  if (tl_x < 0) tl_x = 0;
  int br_x = ((int)im.crop_x + (int)im.out_w + 15 + extra) >> 4;
  if (br_x > mb_w) br_x = mb_w;
  for (int k = 0; k < 55; ++k) tab[k] = kDitherRandomTable[k];
  int i1 = 0, i2 = 31;
  for (int my = 0; my < rows; ++my) {
    for (int mx = tl_x; mx < br_x; ++mx) {
      const size_t idx = (size_t)my * mb_w + mx;
      const uint32_t w = mbinfo[4 * idx + 3];
      // VP8DecodeMB, vp8_dec.c:603,626: skipped macroblocks and those with AC chroma coefficients are left alone
      const int skipped = h->use_skip && (w & MBW_SKIP);
      const int amp = (skipped || (w & 0xaaaau)) ? 0 : (int)h->dither[(w >> MBW_SEG_SHIFT) & 3];
      if (amp < VP8B_MIN_DITHER_AMP) continue;
      mbinfo[4 * idx + 3] = w | MBW_DITHER;
      int8_t* out = plane + idx * 128;
      for (int k = 0; k < 128; ++k) {   // VP8RandomBits2(rg, 8, amp), random_utils.h:39-53
        int diff = (int)(tab[i1] - tab[i2]);
        if (diff < 0) diff += (int)(1u << 31);
        tab[i1] = (uint32_t)diff;
        if (++i1 == 55) i1 = 0;
        if (++i2 == 55) i2 = 0;
        diff = (int)((uint32_t)diff << 1) >> (32 - 8);
        diff = (diff * amp) >> 8;
        diff += 1 << 7;
        out[k] = (int8_t)(((diff - 128) + 8) >> 4);   // delta1 of DitherCombine8x8_C
      }
    }
  }
}

// Adds the planned offsets to the U and V blocks of macroblock (mx, my): lane l handles row l & 7 of U (l < 8) / V (l < 16).
VP8_PFN void dither_macroblock(int mx, int my, int mb_w, const int8_t* plane, uint8_t* uplane, uint8_t* vplane) {
  const int uvs = 8 * mb_w;
  const int8_t* d = plane + ((size_t)my * mb_w + mx) * 128;
  WARP_PHASE(lane)
    if (lane < 16) {
      uint8_t* p = (lane < 8 ? uplane : vplane) + (size_t)(8 * my + (lane & 7)) * uvs + 8 * mx;
      const int8_t* dd = d + 8 * lane;   // U rows 0-7 then V rows 0-7: exactly 64 + 64 offsets in generation order
      for (int k = 0; k < 8; ++k) p[k] = (uint8_t)clip8i((int)p[k] + (int)dd[k]);
    }
  WARP_PHASE_END
}

// ---------------------------------------------------------------------------------------------------------
// options.use_scaling. The reference rescales Y, U and V with one WebPRescaler each (src/utils/rescaler_utils.c,
// src/dsp/rescaler.c: fixed point, 32 fractional bits) as the rows arrive and converts YUV444 -> RGB row by row
// (EmitRescaledRGB / ExportRGB, io_dec.c:357-412), or stores the planes (EmitRescaledYUV, io_dec.c:252-270).
// Here one thread owns one OUTPUT COLUMN of one plane: the horizontal pass of a source row is a closed form of
// the column index (import_row_*), the vertical pass is the reference's state machine with the column's two
// accumulators in registers. Everything is uint32 / uint64 arithmetic exactly as in the reference.
#define RS_ONE (1ull << 32)
#define RS_FRAC(x, y) ((uint32_t)((((uint64_t)(x)) << 32) / (uint64_t)(y)))
#define RS_MULT_FIX(x, y) ((((uint64_t)(x)) * (uint64_t)(y) + (RS_ONE >> 1)) >> 32)
#define RS_MULT_FIX_FLOOR(x, y) ((((uint64_t)(x)) * (uint64_t)(y)) >> 32)

// WebPMultRow_C (alpha_processing.c:160-175): x * a / 255 (inverse = 0) or x * 255 / a (inverse = 1) in 24-bit fixed point.
VP8_PFN uint32_t mult_by_alpha(uint32_t x, uint32_t a, int inverse) {
  if (a == 255) return x;
  if (a == 0) return 0;
  const uint32_t scale = inverse ? (255u << 24) / a : a * ((1u << 24) / 255u);
  return ((x * scale + (1u << 23)) >> 24) & 0xffu;
}

struct Rescaler {   // one plane of one image, as seen by one output column (WebPRescalerInit, rescaler_utils.c:24-84)
  const uint8_t* src; int src_stride;
  int px_step;                            // bytes between horizontally adjacent samples (4: one channel of interleaved BGRA)
  const uint8_t* asrc; int asrc_stride;   // MODE_YUVA: the luma is multiplied by this alpha plane on its way in (io_dec.c:258-265)
  int src_w, src_h, dst_w, dst_h;
  int x_expand, y_expand;
  int x_add, x_sub, y_add, y_sub, y_accum;
  uint32_t fx_scale, fy_scale, fxy_scale;
  int src_y, dst_y;
  uint32_t irow, frow;   // this column's entries of the two work rows
};

VP8_PFN void rescaler_init(Rescaler& r, const uint8_t* src, int src_stride, int src_w, int src_h, int dst_w, int dst_h) {
  r.src = src; r.src_stride = src_stride; r.px_step = 1; r.asrc = 0; r.asrc_stride = 0;
  r.src_w = src_w; r.src_h = src_h; r.dst_w = dst_w; r.dst_h = dst_h;
  r.x_expand = src_w < dst_w; r.y_expand = src_h < dst_h;
  r.x_add = r.x_expand ? dst_w - 1 : src_w;
  r.x_sub = r.x_expand ? src_w - 1 : dst_w;
  r.fx_scale = r.x_expand ? 0u : RS_FRAC(1, r.x_sub);
  r.y_add = r.y_expand ? src_h - 1 : src_h;
  r.y_sub = r.y_expand ? dst_h - 1 : dst_h;
  r.y_accum = r.y_expand ? r.y_sub : r.y_add;
  r.fxy_scale = 0;
  if (!r.y_expand) {
    const uint64_t ratio = ((uint64_t)dst_h * RS_ONE) / ((uint64_t)r.x_add * (uint64_t)r.y_add);
    r.fxy_scale = (ratio != (uint32_t)ratio) ? 0u : (uint32_t)ratio;
    r.fy_scale = RS_FRAC(1, r.y_sub);
  } else {
    r.fy_scale = RS_FRAC(1, r.x_add);
  }
  r.src_y = 0; r.dst_y = 0; r.irow = 0; r.frow = 0;
}

// frow[x] of WebPRescalerImportRowShrink_C (rescaler.c:62-95) for one source row: output k consumes the inputs
// n(k-1) .. n(k)-1 with n(k) = ceil((k+1) * x_add / x_sub), starting from the fraction the previous output left over.
#define RS_PX(i) (arow ? mult_by_alpha(row[i], arow[i], 0) : (uint32_t)row[(i) * r.px_step])
VP8_PFN uint32_t import_row_shrink(const Rescaler& r, const uint8_t* row, const uint8_t* arow, int x) {
  const int64_t t1 = (int64_t)(x + 1) * r.x_add;
  const int n1 = (int)((t1 + r.x_sub - 1) / r.x_sub);
  const int acc1 = (int)(t1 - (int64_t)n1 * r.x_sub);          // accum after this output (<= 0)
  int n0 = 0;
  uint32_t sum = 0;
  if (x > 0) {
    const int64_t t0 = (int64_t)x * r.x_add;
    n0 = (int)((t0 + r.x_sub - 1) / r.x_sub);
    const int acc0 = (int)(t0 - (int64_t)n0 * r.x_sub);
    const uint32_t frac0 = RS_PX(n0 - 1) * (uint32_t)(-acc0);
    sum = (uint32_t)(int)RS_MULT_FIX(frac0, r.fx_scale);
  }
  for (int i = n0; i < n1; ++i) sum += RS_PX(i);
  const uint32_t frac = RS_PX(n1 - 1) * (uint32_t)(-acc1);
  return sum * (uint32_t)r.x_sub - frac;
}

// frow[x] of WebPRescalerImportRowExpand_C (rescaler.c:29-60): bilinear, the input position advanced a(k) times before
// output k, a(k) the smallest a >= 0 with x_add - k * x_sub + a * x_add >= 0.
VP8_PFN uint32_t import_row_expand(const Rescaler& r, const uint8_t* row, const uint8_t* arow, int x) {
  const int64_t d = (int64_t)x * r.x_sub - r.x_add;
  const int a = (d > 0) ? (int)((d + r.x_add - 1) / r.x_add) : 0;
  const int accum = (int)((int64_t)r.x_add - (int64_t)x * r.x_sub + (int64_t)a * r.x_add);
  const uint32_t left = RS_PX(a);
  const uint32_t right = (r.src_w > 1) ? RS_PX(a + 1 < r.src_w ? a + 1 : r.src_w - 1) : left;
  return right * (uint32_t)r.x_add + (left - right) * (uint32_t)accum;
}

// The next output sample of this column: import source rows until one is due (WebPRescalerImport, rescaler_utils.c:
// 132-155), then export it (WebPRescalerExportRow + ExportRowExpand_C / ExportRowShrink_C, rescaler.c:100-196).
VP8_PFN int rescaler_next(Rescaler& r, int x) {
  while (r.y_accum > 0 && r.src_y < r.src_h) {
    const uint8_t* row = r.src + (size_t)r.src_y * r.src_stride;
    const uint8_t* arow = r.asrc ? r.asrc + (size_t)r.src_y * r.asrc_stride : 0;
    if (r.y_expand) r.irow = r.frow;   // the two work rows swap roles
    r.frow = r.x_expand ? import_row_expand(r, row, arow, x) : import_row_shrink(r, row, arow, x);
    if (!r.y_expand) r.irow += r.frow;
    ++r.src_y;
    r.y_accum -= r.y_sub;
  }
  int v;
  if (r.y_expand) {
    if (r.y_accum == 0) {
      v = (int)RS_MULT_FIX(r.frow, r.fy_scale);
    } else {
      const uint32_t B = RS_FRAC(-r.y_accum, r.y_sub);
      const uint32_t A = (uint32_t)(RS_ONE - B);
      const uint64_t I = (uint64_t)A * r.frow + (uint64_t)B * r.irow;
      const uint32_t J = (uint32_t)((I + (RS_ONE >> 1)) >> 32);
      v = (int)RS_MULT_FIX(J, r.fy_scale);
    }
  } else if (r.fxy_scale) {
    const uint32_t yscale = r.fy_scale * (uint32_t)(-r.y_accum);
    if (yscale) {
      const uint32_t frac = (uint32_t)RS_MULT_FIX_FLOOR(r.frow, yscale);
      v = (int)RS_MULT_FIX(r.irow - frac, r.fxy_scale);
      r.irow = frac;
    } else {
      v = (int)RS_MULT_FIX(r.irow, r.fxy_scale);
      r.irow = 0;
    }
  } else {   // src_width == 1, same height: the accumulated row as it stands (rescaler.c:208-216)
    v = (int)(r.irow & 0xffu);
    r.irow = 0;
    r.y_accum += r.y_add; ++r.dst_y;
    return v;
  }
  r.y_accum += r.y_add;
  ++r.dst_y;
  return v > 255 ? 255 : v;
}

// One pixel of any RGB-family colourspace at column x of output row `orow`; a = its alpha (0xff for opaque images; the
// premultiplied modes scale the colour by it like ExportAlpha + WebPApplyAlphaMultiply, io_dec.c:414-450).
VP8_PFN void store_rgb_pixel(int csp, int y, int u, int v, int a, uint8_t* orow, int x) {
  int r, g, b;
  yuv_to_rgb(y, u, v, &r, &g, &b);
  if (csp == 0 || csp == 2) {
    uint8_t* o = orow + 3 * x;
    if (csp == 0) { o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b; } else { o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r; }
  } else if (csp == 5 || csp == 6 || csp == 10) {
    const uint32_t p2 = pack_pixel2(csp, r, g, b, a);
    orow[2 * x] = (uint8_t)p2; orow[2 * x + 1] = (uint8_t)(p2 >> 8);
  } else {
    const uint32_t p4 = pack_pixel4(csp, r, g, b, a);
    uint8_t* o = orow + 4 * x;
    o[0] = (uint8_t)p4; o[1] = (uint8_t)(p4 >> 8); o[2] = (uint8_t)(p4 >> 16); o[3] = (uint8_t)(p4 >> 24);
  }
}

// Work item `t` of a scaled image. RGB family: t = output column, all rows (Y, U, V each rescaled to dst_w x dst_h,
// then converted). MODE_YUV / MODE_YUVA: t in [0, dst_w) = a Y column, then (dst_w+1)/2 U columns, as many V columns,
// and for MODE_YUVA dst_w columns of the (opaque) alpha plane. yplane/uplane/vplane = the window's origin.
VP8_PFN void emit_scaled_column(const ImgDesc& im, const uint8_t* yplane, const uint8_t* uplane, const uint8_t* vplane,
                                const uint8_t* alpha /* window origin inside the frame-wide plane, or NULL */, uint8_t* out, int t) {
  const int sw = im.out_w, sh = im.out_h, dw = im.dst_w, dh = im.dst_h;
  const int ys = 16 * im.mb_w, uvs = 8 * im.mb_w;
  const int uv_sw = (sw + 1) >> 1, uv_sh = (sh + 1) >> 1;
  const int flip = (im.flags & VP8B_FLAG_FLIP) != 0;
  if (im.csp == 11 || im.csp == 12) {
    const int uv_dw = (dw + 1) >> 1, uv_dh = (dh + 1) >> 1;
    uint8_t* uplane_out = out + (size_t)im.out_stride * dh;
    uint8_t* aplane_out = uplane_out + 2 * (size_t)uv_dw * uv_dh;
    Rescaler r;
    uint8_t* dst; int stride, rows, x;
    if (t < dw) {   // a luma column; MODE_YUVA with an alpha plane: multiplied by alpha on the way in, divided by the
                    // rescaled alpha on the way out (EmitRescaledYUV / EmitRescaledAlphaYUV, io_dec.c:252-300)
      rescaler_init(r, yplane, ys, sw, sh, dw, dh);
      if (im.csp == 12 && alpha != 0) {
        Rescaler ra;
        rescaler_init(ra, alpha, im.width, sw, sh, dw, dh);
        r.asrc = alpha; r.asrc_stride = im.width;
        for (int k = 0; k < dh; ++k) {
          const uint32_t y = (uint32_t)rescaler_next(r, t), a = (uint32_t)rescaler_next(ra, t);
          const int kd = flip ? dh - 1 - k : k;
          out[(size_t)kd * im.out_stride + t] = (uint8_t)mult_by_alpha(y, a, 1);
          aplane_out[(size_t)kd * dw + t] = (uint8_t)a;
        }
        return;
      }
      dst = out; stride = im.out_stride; rows = dh; x = t;
    }
    else if (t < dw + uv_dw) { rescaler_init(r, uplane, uvs, uv_sw, uv_sh, uv_dw, uv_dh); dst = uplane_out; stride = uv_dw; rows = uv_dh; x = t - dw; }
    else if (t < dw + 2 * uv_dw) { rescaler_init(r, vplane, uvs, uv_sw, uv_sh, uv_dw, uv_dh); dst = uplane_out + (size_t)uv_dw * uv_dh; stride = uv_dw; rows = uv_dh; x = t - dw - uv_dw; }
    else if (im.csp == 12 && alpha == 0 && t < 2 * dw + 2 * uv_dw) {
      for (int k = 0; k < dh; ++k) aplane_out[(size_t)k * dw + (t - dw - 2 * uv_dw)] = 0xff;   // FillAlphaPlane, io_dec.c:283-288
      return;
    } else return;
    for (int k = 0; k < rows; ++k) dst[(size_t)(flip ? rows - 1 - k : k) * stride + x] = (uint8_t)rescaler_next(r, x);
    return;
  }
  if (t >= dw) return;
  const int with_alpha = alpha != 0 && (im.csp == 1 || im.csp == 3 || im.csp == 4 || im.csp == 5 || (im.csp >= 7 && im.csp <= 10));
  Rescaler ry, ru, rv, ra;
  rescaler_init(ry, yplane, ys, sw, sh, dw, dh);
  rescaler_init(ru, uplane, uvs, uv_sw, uv_sh, dw, dh);
  rescaler_init(rv, vplane, uvs, uv_sw, uv_sh, dw, dh);
  rescaler_init(ra, with_alpha ? alpha : yplane, with_alpha ? im.width : ys, sw, sh, dw, dh);
  for (int k = 0; k < dh; ++k) {
    const int y = rescaler_next(ry, t), u = rescaler_next(ru, t), v = rescaler_next(rv, t);
    const int a = with_alpha ? rescaler_next(ra, t) : 0xff;
    store_rgb_pixel(im.csp, y, u, v, a, out + (size_t)(flip ? dh - 1 - k : k) * im.out_stride, t);
  }
}

// MODE_YUV (EmitYUV, io_dec.c:25-40): 16 bytes of one row of one plane. plane 0 = Y (w x h), 1 = U, 2 = V
// ((w+1)/2 x (h+1)/2). Output = y | u | v at out_off with strides out_stride / (w+1)/2.
// plane 3 (MODE_YUVA only, EmitAlphaYUV io_dec.c:131-152) = the alpha plane, 0xff where the file has none.
VP8_PFN void emit_yuv_chunk(const ImgDesc& im, const uint8_t* yplane, const uint8_t* uplane, const uint8_t* vplane,
                            const uint8_t* alpha, uint8_t* out, int plane, int q, int j) {
  const int w = im.out_w, h = im.out_h;
  const int uvw = (w + 1) >> 1, uvh = (h + 1) >> 1;
  if (plane == 3) {
    const int jd = (im.flags & VP8B_FLAG_FLIP) ? h - 1 - j : j;
    uint8_t* dst = out + (size_t)im.out_stride * h + 2 * (size_t)uvw * uvh + (size_t)jd * w + 16 * q;
    const int n = (w - 16 * q < 16) ? w - 16 * q : 16;
    for (int k = 0; k < n; ++k) dst[k] = alpha ? alpha[(size_t)j * im.width + 16 * q + k] : 0xff;
    return;
  }
  const int pw = plane ? uvw : w;
  const int jd = (im.flags & VP8B_FLAG_FLIP) ? (plane ? uvh : h) - 1 - j : j;   // destination row
  const int sstride = plane ? 8 * im.mb_w : 16 * im.mb_w;
  const int dstride = plane ? uvw : im.out_stride;
  const uint8_t* src = (plane == 0 ? yplane : plane == 1 ? uplane : vplane) + (size_t)j * sstride + 16 * q;
  uint8_t* dst = out + (plane == 0 ? 0 : (size_t)im.out_stride * h + (plane == 2 ? (size_t)uvw * uvh : 0)) +
                 (size_t)jd * dstride + 16 * q;
  const int n = (pw - 16 * q < 16) ? pw - 16 * q : 16;
  if (n == 16 && ((((uintptr_t)dst) | ((uintptr_t)src)) & 15) == 0) {
    *(uint4*)dst = *(const uint4*)src;
  } else {
    for (int k = 0; k < n; ++k) dst[k] = src[k];
  }
}

#endif  // LIBWEBP_B200_VP8_PIXEL_CORE_H_
