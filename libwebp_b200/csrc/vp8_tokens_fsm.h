// vp8_tokens_fsm.h -- coefficient-token parser as a lane-parallel finite-state machine.
//
// One LANE owns one token partition (one image when partitions = 1) and keeps the boolean decoder, the token
// tree position and the block/macroblock bookkeeping in registers. Every loop iteration performs exactly one
// boolean decode per lane, and what happens next is read from a 16-state x {0,1} transition table, so the
// lanes of a warp execute the same instruction stream although each is somewhere else in its own bitstream
// (GetCoeffs' nested loops and GetLargeValue's branches, src/dec/vp8_dec.c:411-469, become table rows).
// Only block ends (every ~18 decodes), the rare large-value categories and macroblock starts branch.
//
// Everything on the per-decode path is 32-bit: shared-window addresses for the probabilities and tables,
// a word index into the input arena, an element index into the coefficient plane.
//
// Replaces VP8DecodeMB / ParseResiduals / GetCoeffs / GetLargeValue (src/dec/vp8_dec.c:400-635) for a batch.
// Dual build like the other cores: nvcc for the product, g++ -DVP8_EMU for tests/emu.
#ifndef LIBWEBP_B200_VP8_TOKENS_FSM_H_
#define LIBWEBP_B200_VP8_TOKENS_FSM_H_

#include "vp8_dev.h"

#if defined(__CUDACC__) && !defined(VP8_EMU)
#define TK_FN __device__ __forceinline__
#define TK_CLZ(x) __clz((int)(x))
#define TK_BSWAP(x) __byte_perm((x), 0, 0x0123)
#define TK_FENCE() __threadfence_block()
typedef uint32_t tk_saddr;   // address inside the shared-memory window
TK_FN tk_saddr tk_saddr_of(const void* p) { return (tk_saddr)__cvta_generic_to_shared(p); }
TK_FN uint32_t tk_lds_u8(tk_saddr a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
TK_FN uint32_t tk_lds_u16(tk_saddr a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
TK_FN void tk_lds_v2(tk_saddr a, uint32_t& x, uint32_t& y) { asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
TK_FN uint32_t tk_shr_clamp(uint32_t w, int n) { return __funnelshift_rc(w, 0u, (uint32_t)n); }   // w >> n, 0 when n >= 32
TK_FN uint32_t tk_shl_pair(uint32_t hi, uint32_t lo, int n) { return __funnelshift_l(lo, hi, (uint32_t)n); }
#else
#define TK_FN static inline
#define TK_CLZ(x) __builtin_clz((unsigned)(x))
#define TK_BSWAP(x) __builtin_bswap32(x)
#define TK_FENCE() ((void)0)
typedef uintptr_t tk_saddr;
TK_FN tk_saddr tk_saddr_of(const void* p) { return (tk_saddr)p; }
TK_FN uint32_t tk_lds_u8(tk_saddr a) { return *(const uint8_t*)a; }
TK_FN uint32_t tk_lds_u16(tk_saddr a) { return *(const uint16_t*)a; }
TK_FN void tk_lds_v2(tk_saddr a, uint32_t& x, uint32_t& y) { x = ((const uint32_t*)a)[0]; y = ((const uint32_t*)a)[1]; }
TK_FN uint32_t tk_shr_clamp(uint32_t w, int n) { return n >= 32 ? 0u : (w >> n); }
TK_FN uint32_t tk_shl_pair(uint32_t hi, uint32_t lo, int n) { return n == 0 ? hi : ((hi << n) | (lo >> (32 - n))); }
#endif

// ---- token-tree states
enum { S_P0 = 0, S_P1, S_P2, S_P3, S_P4, S_P5, S_P6, S_P7, S_P8, S_P9, S_P10, S_C159, S_C165, S_C145, S_EXTRA, S_SIGN };

// ---- transition word
//  [6:3] next state * 8   [11:7] offset (relative move in the probability row, or index into the constants)
//  12 ABS  13 NEWCOEF  14 ZERO  15 EMIT  16 NEG  17 EOB  [19:18] context of the next coefficient (0 = keep)
//  20 CAT  21 EXTRA  [23:22] category  [25:24] v multiplier  [29:26] v addend
#define TE_STATE8(e) ((e) & 0x78u)
#define TE_OFF(e) (((e) >> 7) & 31u)
#define TE_ABS (1u << 12)
#define TE_NEWCOEF (1u << 13)
#define TE_ZERO (1u << 14)
#define TE_EMIT (1u << 15)
#define TE_NEG (1u << 16)
#define TE_EOB (1u << 17)
#define TE_CTX(e) (((e) >> 18) & 3u)
#define TE_CAT (1u << 20)
#define TE_EXTRA (1u << 21)
#define TE_CATV(e) (((e) >> 22) & 3u)
#define TE_VMUL(e) (((e) >> 24) & 3u)
#define TE_VADD(e) (((e) >> 26) & 15u)

#define TE_(next, off, vmul, vadd, flags) (((uint32_t)(next) << 3) | ((uint32_t)(off) << 7) | ((uint32_t)(vmul) << 24) | ((uint32_t)(vadd) << 26) | (flags))
#define TE_C1 (1u << 18)
#define TE_C2 (2u << 18)
#define TE_CATN(c) (TE_CAT | ((uint32_t)(c) << 22))

// constants region appended to each image's probability block (byte offsets from TokImage::consts)
#define TKC_SIGN 0
#define TKC_159 1
#define TKC_165 2
#define TKC_145 3
#define TKC_CAT3 4
#define TKC_CAT4 7
#define TKC_CAT5 11
#define TKC_CAT6 16
#define TKC_BYTES 28
#define TK_CONST_INIT { 128, 159, 165, 145, 173, 148, 140, 176, 155, 140, 135, 180, 157, 141, 134, 130, \
                        254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0 }

// [state][bit]
#define TK_TABLE_INIT {                                                                                          \
  /* P0   */ { TE_(S_P0, 0, 1, 0, TE_EOB),                           TE_(S_P1, 1, 1, 0, 0) },                    \
  /* P1   */ { TE_(S_P1, 0, 1, 0, TE_ZERO | TE_NEWCOEF),             TE_(S_P2, 1, 1, 0, 0) },                    \
  /* P2   */ { TE_(S_SIGN, TKC_SIGN, 0, 1, TE_ABS | TE_C1),          TE_(S_P3, 1, 1, 0, TE_C2) },                \
  /* P3   */ { TE_(S_P4, 1, 1, 0, 0),                                TE_(S_P6, 3, 1, 0, 0) },                    \
  /* P4   */ { TE_(S_SIGN, TKC_SIGN, 0, 2, TE_ABS),                  TE_(S_P5, 1, 1, 0, 0) },                    \
  /* P5   */ { TE_(S_SIGN, TKC_SIGN, 0, 3, TE_ABS),                  TE_(S_SIGN, TKC_SIGN, 0, 4, TE_ABS) },      \
  /* P6   */ { TE_(S_P7, 1, 1, 0, 0),                                TE_(S_P8, 2, 1, 0, 0) },                    \
  /* P7   */ { TE_(S_C159, TKC_159, 1, 0, TE_ABS),                   TE_(S_C165, TKC_165, 1, 0, TE_ABS) },       \
  /* P8   */ { TE_(S_P9, 1, 1, 0, 0),                                TE_(S_P10, 2, 1, 0, 0) },                   \
  /* P9   */ { TE_(S_EXTRA, TKC_CAT3, 0, 0, TE_ABS | TE_CATN(0)),    TE_(S_EXTRA, TKC_CAT4, 0, 0, TE_ABS | TE_CATN(1)) }, \
  /* P10  */ { TE_(S_EXTRA, TKC_CAT5, 0, 0, TE_ABS | TE_CATN(2)),    TE_(S_EXTRA, TKC_CAT6, 0, 0, TE_ABS | TE_CATN(3)) }, \
  /* C159 */ { TE_(S_SIGN, TKC_SIGN, 0, 5, TE_ABS),                  TE_(S_SIGN, TKC_SIGN, 0, 6, TE_ABS) },      \
  /* C165 */ { TE_(S_C145, TKC_145, 0, 7, TE_ABS),                   TE_(S_C145, TKC_145, 0, 9, TE_ABS) },       \
  /* C145 */ { TE_(S_SIGN, TKC_SIGN, 1, 0, TE_ABS),                  TE_(S_SIGN, TKC_SIGN, 1, 1, TE_ABS) },      \
  /* EXTRA*/ { TE_(S_EXTRA, 1, 2, 0, TE_EXTRA),                      TE_(S_EXTRA, 1, 2, 1, TE_EXTRA) },          \
  /* SIGN */ { TE_(S_P0, 0, 1, 0, TE_EMIT | TE_NEWCOEF),             TE_(S_P0, 0, 1, 0, TE_EMIT | TE_NEG | TE_NEWCOEF) } }

// Block-wide constant tables in shared memory (filled by tk_tables_fill).
struct TokTables {
  uint32_t trans[16][2];   // 128 B
  uint16_t band_off[18];   // band(n) * 33 for n = 0..16 (+ one spare), at byte offset 128
  uint8_t zigzag[16];      // at byte offset 164
  uint8_t pad[12];
};
#define TKT_BAND 128
#define TKT_ZIGZAG 164

// Shared-memory block of one image: probabilities, token constants, dequantisers, flags.
struct TokImage {
  uint8_t prob[1056];          // [type][band][ctx][node]
  uint8_t consts[TKC_BYTES];   // at offset 1056
  int16_t dq[4][6];            // per segment: y1 dc/ac, y2 dc/ac, uv dc/ac
  int32_t ok;                  // header parsed and the partition count matches the launch
  int32_t use_skip;
  int32_t pad[2];
};

// Per-lane state (registers). phase: 0 = needs a macroblock, 1 = decoding, 2 = finished.
struct TokLane {
  // boolean decoder: 64-bit left-aligned window vhi:vlo, range stored minus one
  uint32_t vhi, vlo;
  int nbits;
  uint32_t range;
  uint32_t wi;        // next word of the stream, as an index into the arena viewed as uint32[]
  int shift;          // renormalisation shift of the most recent decode
  // token tree
  uint32_t state8;    // state * 8
  tk_saddr pp;        // shared address of the next probability
  tk_saddr pnb;       // shared address of the row for band(n+1), ctx 0, of the current block type
  tk_saddr pbase;     // shared address of the current block type's [8][3][11] table
  int v, n;
  uint32_t nc11;      // 11 * context of the next coefficient
  int extra_left, cat;
  // block
  int seq;            // 0 = Y2, 1..16 luma, 17..24 chroma; 25 = macroblock done
  int first, dcnz;
  uint32_t tnz, lnz;  // bit 0-3 luma, 4-5 U, 6-7 V, 8 Y2
  uint32_t nzy, nzuv;
  uint32_t dq;        // dc | ac << 16
  uint32_t outi;      // index of the current block's first coefficient inside the image's coefficient plane
  // macroblock / partition
  int mx, my, part, done_mbs, phase, status;
  uint32_t w;
  int seg;
  // stream geometry for the end-of-data test (bits)
  int64_t pos_bias;   // stream bits loaded = 32 * wi - pos_bias
  int64_t limit;      // 8 * size - 8
};

// Per-image shared state handed to the lanes.
struct TokShared {
  const TokImage* img;       // shared memory (generic pointer, rare paths)
  tk_saddr img_s;            // same, as a shared address (per-decode path)
  tk_saddr tab_s;            // TokTables
  uint16_t* topctx;          // (P+1) x mb_w ring
  volatile int* progress;    // P counters
};

TK_FN void tk_tables_fill(TokTables* t, int tid, int nthreads) {
  const uint32_t trans[16][2] = TK_TABLE_INIT;
  const uint8_t bands[17] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0 };
  const uint8_t zz[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
  for (int k = tid; k < 32; k += nthreads) t->trans[k >> 1][k & 1] = trans[k >> 1][k & 1];
  for (int k = tid; k < 18; k += nthreads) t->band_off[k] = (uint16_t)((k < 17 ? bands[k] : 0) * 33);
  for (int k = tid; k < 16; k += nthreads) t->zigzag[k] = zz[k];
}

// Fills the shared per-image block from the parsed frame header (any thread subset may call with its slice).
TK_FN void tk_image_fill(TokImage* im, const FrameHdr* h, int P, int tid, int nthreads) {
  const uint8_t consts[TKC_BYTES] = TK_CONST_INIT;
  for (int k = tid; k < 264; k += nthreads) ((uint32_t*)im->prob)[k] = ((const uint32_t*)h->prob)[k];
  for (int k = tid; k < TKC_BYTES; k += nthreads) im->consts[k] = consts[k];
  for (int k = tid; k < 24; k += nthreads) im->dq[k / 6][k % 6] = h->dq[k / 6][k % 6];
  if (tid == 0) { im->ok = (h->status == VP8B_OK && h->num_parts == P) ? 1 : 0; im->use_skip = h->use_skip; }
}

// Lane for token partition `part` of an image. `arena32` = the input arena as words, frame_off = byte offset
// of the frame tag inside it. The arena is padded by >= 32 KB so that a lane that ran past the end of its
// stream (detected at the next macroblock boundary) never leaves the allocation.
TK_FN void tk_lane_init(TokLane& L, const uint32_t* arena32, uint64_t frame_off, const FrameHdr* h, int part) {
  const uint64_t a = frame_off + h->part_off[part];
  const uint32_t size = h->part_size[part];
  const int off = (int)(a & 3);
  L.wi = (uint32_t)(a >> 2);
  const uint32_t w0 = TK_BSWAP(arena32[L.wi]);
  L.wi++;
  L.vhi = (off == 0) ? w0 : (w0 << (8 * off));
  L.vlo = 0;
  L.nbits = 32 - 8 * off;
  L.range = 254;
  L.shift = 0;
  L.pos_bias = 32 * (int64_t)(a >> 2) + 8 * off;
  L.limit = 8 * (int64_t)size - 8;
  L.state8 = 0; L.pp = 0; L.pnb = 0; L.pbase = 0;
  L.v = 0; L.n = 0; L.nc11 = 0; L.extra_left = 0; L.cat = 0;
  L.seq = 0; L.first = 0; L.dcnz = 0; L.tnz = 0; L.lnz = 0; L.nzy = 0; L.nzuv = 0; L.dq = 0; L.outi = 0;
  L.mx = 0; L.my = part; L.part = part; L.done_mbs = 0; L.phase = 0; L.status = VP8B_OK; L.w = 0; L.seg = 0;
}

// The reference's eof_ flag from bit positions (see vp8_parse_core.h:bd_eof).
TK_FN int tk_eof(const TokLane& L) {
  const int64_t loaded = 32 * (int64_t)L.wi - L.pos_bias;
  return (loaded - L.nbits - L.shift) > L.limit;
}

// Sets up block `seq` of the current macroblock (context, probabilities, dequantisers, output index).
TK_FN void tk_block_setup(TokLane& L, const TokShared& sh, uint32_t mb_coef_index) {
  const int seq = L.seq;
  const int is_i4 = (L.w & MBW_I4X4) != 0;
  int blk, tb, lb, type, qi;
  if (seq == 0) { blk = 24; tb = 8; lb = 8; type = 1; qi = 2; }
  else {
    blk = seq - 1;
    if (blk < 16) { tb = blk & 3; lb = blk >> 2; type = is_i4 ? 3 : 0; qi = 0; }
    else { const int c = blk - 16; tb = 4 + (c & 1) + 2 * (c >> 2); lb = 4 + ((c >> 1) & 1) + 2 * (c >> 2); type = 2; qi = 4; }
  }
  const uint32_t ctx = ((L.tnz >> tb) & 1) + ((L.lnz >> lb) & 1);
  const int16_t* q = sh.img->dq[L.seg];
  L.dq = (uint32_t)(uint16_t)q[qi] | ((uint32_t)(uint16_t)q[qi + 1] << 16);
  L.first = (blk < 16 && !is_i4) ? 1 : 0;
  L.n = L.first;
  L.pbase = sh.img_s + (uint32_t)type * 264u;
  L.pp = L.pbase + (uint32_t)L.first * 33u + ctx * 11u;
  L.pnb = L.pbase + tk_lds_u16(sh.tab_s + TKT_BAND + 2 * (L.n + 1));
  L.state8 = S_P0 * 8;
  L.dcnz = 0;
  L.outi = mb_coef_index + (uint32_t)blk * 16u;
}

// Finishes block `seq` with return value nz (GetCoeffs), updates contexts and nz codes (vp8_dec.c:517-609).
TK_FN void tk_block_end(TokLane& L, int nz) {
  const int seq = L.seq;
  if (seq == 0) {
    const uint32_t f = (nz > 0) ? 0x100u : 0u;
    L.tnz = (L.tnz & 0xffu) | f;
    L.lnz = (L.lnz & 0xffu) | f;
    if (nz > 0) L.w |= MBW_HAS_Y2;
  } else {
    const int blk = seq - 1;
    const uint32_t code = (nz > 3) ? 3u : (nz > 1) ? 2u : (uint32_t)L.dcnz;
    const uint32_t l = (nz > L.first) ? 1u : 0u;
    int tb, lb;
    if (blk < 16) {
      tb = blk & 3; lb = blk >> 2;
      L.nzy |= code << (30 - 2 * blk);
    } else {
      const int c = blk - 16;
      tb = 4 + (c & 1) + 2 * (c >> 2); lb = 4 + ((c >> 1) & 1) + 2 * (c >> 2);
      L.nzuv |= code << (8 * (c >> 2) + 6 - 2 * (c & 3));
    }
    L.tnz = (L.tnz & ~(1u << tb)) | (l << tb);
    L.lnz = (L.lnz & ~(1u << lb)) | (l << lb);
  }
  L.seq = seq + 1;
}

// Writes the macroblock's results and advances to the next one.
TK_FN void tk_mb_finish(TokLane& L, const TokShared& sh, const ImgDesc& im, int P, uint32_t* mbinfo) {
  const int mb_w = im.mb_w;
  const size_t idx = (size_t)L.my * mb_w + L.mx;
  mbinfo[4 * idx + 2] = L.nzy;
  mbinfo[4 * idx + 3] = (L.w & 0xffff0000u) | L.nzuv;
  sh.topctx[(size_t)(L.my % (P + 1)) * mb_w + L.mx] = (uint16_t)L.tnz;
  L.done_mbs++;
  L.phase = 0;
  if (++L.mx == mb_w) { L.mx = 0; L.my += P; }
  if (tk_eof(L)) {
    // Ran past the end of the partition: the image is lost (vp8_dec.c:651-659). Stop reading the bitstream and
    // release every partition that waits on this one.
    L.status = VP8B_NOT_ENOUGH_DATA;
    L.phase = 2;
    if (P > 1) { TK_FENCE(); sh.progress[L.part] = 0x7fffffff; }
    return;
  }
  if (P > 1) { TK_FENCE(); sh.progress[L.part] = L.done_mbs; }
}

// Starts the next macroblock of this lane's partition, or finishes the lane. Returns without doing anything
// when the partition owning the row above has not got far enough yet (the caller simply retries).
TK_FN void tk_mb_start(TokLane& L, const TokShared& sh, const ImgDesc& im, int P, uint32_t* mbinfo) {
  const int mb_w = im.mb_w;
  if (L.my >= im.mb_h) { L.phase = 2; return; }
  uint32_t tctx = 0;
  if (L.my > 0) {
    if (P > 1) {
      const int prev = (L.part + P - 1) % P;
      const int need = ((L.my - 1 - prev) / P) * mb_w + L.mx + 1;
      if (sh.progress[prev] < need) return;   // not yet: stay in phase 0
      TK_FENCE();
    }
    tctx = sh.topctx[(size_t)((L.my + P) % (P + 1)) * mb_w + L.mx];
  }
  const size_t idx = (size_t)L.my * mb_w + L.mx;
  L.w = mbinfo[4 * idx + 3];
  L.seg = (int)((L.w >> MBW_SEG_SHIFT) & 3);
  L.tnz = tctx;
  if (L.mx == 0) L.lnz = 0;
  L.nzy = 0; L.nzuv = 0;
  const int is_i4 = (L.w & MBW_I4X4) != 0;
  if (sh.img->use_skip && (L.w & MBW_SKIP)) {
    L.tnz &= is_i4 ? 0x100u : 0u;
    L.lnz &= is_i4 ? 0x100u : 0u;
    tk_mb_finish(L, sh, im, P, mbinfo);   // nothing to parse; stays in phase 0 for the next macroblock
  } else {
    L.seq = is_i4 ? 1 : 0;
    tk_block_setup(L, sh, (uint32_t)idx * VP8B_COEFFS_PER_MB);
    L.phase = 1;
  }
}

// One iteration of a lane in phase 1: one boolean decode and its consequences.
//   arena32 : input arena as words          coeffs : this image's coefficient plane (int16, pre-zeroed)
TK_FN void tk_step(TokLane& L, const TokShared& sh, const ImgDesc& im, int P, const uint32_t* arena32,
                   uint32_t* mbinfo, int16_t* coeffs) {
  const uint32_t prob = tk_lds_u8(L.pp);
  uint32_t e0, e1;
  tk_lds_v2(sh.tab_s + L.state8, e0, e1);
  // ---- boolean decode (bit_reader_inl_utils.h:107-136, range kept minus one)
  if (L.nbits <= 32) {   // vlo is empty: append one big-endian word behind the valid bits
    const uint32_t w = TK_BSWAP(arena32[L.wi]);
    L.wi++;
    L.vhi |= tk_shr_clamp(w, L.nbits);
    L.vlo = w << (32 - L.nbits);
    L.nbits += 32;
  }
  const uint32_t split = (L.range * prob) >> 8;
  const int bit = (L.vhi >> 24) > split;
  const uint32_t r = bit ? (L.range - split) : (split + 1);
  L.vhi -= bit ? ((split + 1) << 24) : 0u;
  const int shift = TK_CLZ(r) - 24;
  L.range = (r << shift) - 1;
  L.vhi = tk_shl_pair(L.vhi, L.vlo, shift);
  L.vlo <<= shift;
  L.nbits -= shift;
  L.shift = shift;
  // ---- transition
  const uint32_t e = bit ? e1 : e0;
  L.v = L.v * (int)TE_VMUL(e) + (int)TE_VADD(e);
  L.state8 = TE_STATE8(e);
  tk_saddr pp = ((e & TE_ABS) ? (sh.img_s + 1056u) : L.pp) + TE_OFF(e);
  const uint32_t cs = TE_CTX(e);
  L.nc11 = cs ? cs * 11u : L.nc11;
  if (e & (TE_CAT | TE_EXTRA)) {   // DCT_CAT3..6: rare
    if (e & TE_CAT) {
      L.cat = (int)TE_CATV(e);
      L.extra_left = (L.cat == 3) ? 11 : 3 + L.cat;
    } else if (--L.extra_left == 0) {
      L.v += 3 + (8 << L.cat);
      L.state8 = S_SIGN * 8;
      pp = sh.img_s + 1056u + TKC_SIGN;
    }
  }
  int done = -1;
  if (e & TE_EMIT) {
    const int val = (e & TE_NEG) ? -L.v : L.v;
    const int q = (int)((L.n > 0) ? (L.dq >> 16) : (L.dq & 0xffffu));
    const int16_t c = (int16_t)(val * q);
    coeffs[L.outi + tk_lds_u8(sh.tab_s + TKT_ZIGZAG + L.n)] = c;
    L.dcnz |= (L.n == 0 && c != 0) ? 1 : 0;
  }
  if (e & TE_NEWCOEF) {   // a zero (row ctx 0, node 1) or a finished coefficient (row nextctx, node 0)
    pp = L.pnb + ((e & TE_ZERO) ? 1u : L.nc11);
    L.n++;
    L.pnb = L.pbase + tk_lds_u16(sh.tab_s + TKT_BAND + 2 * (L.n + 1));
    if (L.n == 16) done = 16;
  }
  if (e & TE_EOB) done = L.n;
  L.pp = pp;
  if (done >= 0) {
    tk_block_end(L, done);
    if (L.seq < 25) {
      const size_t idx = (size_t)L.my * im.mb_w + L.mx;
      tk_block_setup(L, sh, (uint32_t)idx * VP8B_COEFFS_PER_MB);
    } else {
      tk_mb_finish(L, sh, im, P, mbinfo);
    }
  }
}

#endif  // LIBWEBP_B200_VP8_TOKENS_FSM_H_
