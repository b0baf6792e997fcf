// vp8_tokens_lockstep.h -- coefficient-token parser, third mapping: few lanes per warp in strict lockstep.
//
// One LANE owns one token partition. Every loop iteration is exactly one boolean decode per lane, and the whole
// token syntax of GetCoeffs / GetLargeValue (src/dec/vp8_dec.c:411-469) is a 61-state x {0,1} transition table.
// What makes the iteration short (about 40 SASS instructions, against ~25 per decode for a warp that serves one
// stream only) is that the probability address IS the state:
//
//   row   shared address of a 64-byte row, one per (block type, coefficient position n); walking a block is
//         `row += 64`, and n == 16 is `row == rowend`
//   s     index inside the row: 0..32 = [ctx 3][node 11] of that position's band, 33..60 = the fixed
//         probabilities (sign, 159/165/145, the DCT_CAT3..6 extra bits), replicated in every row so that
//         `prob = row[s]` holds in every state
//
// and a transition entry is only {next s, advance row?, emit?, end of block?, addend of the value}: the context
// of the next coefficient, the category and the extra-bit weights are all folded into which state comes next.
// Block ends (every ~17 decodes per lane) and macroblock ends branch; finished lanes leave the loop.
//
// The compressed bytes are read straight from HBM through the read-only path, one aligned word per refill,
// fetched one refill (>= 32 payload bits, ~40 iterations) ahead of its use.
//
// Replaces VP8DecodeMB / ParseResiduals / GetCoeffs / GetLargeValue (src/dec/vp8_dec.c:400-635) for a batch.
// Dual build like the other cores: nvcc for the product, g++ -DVP8_EMU for tests/emu.
#ifndef LIBWEBP_B200_VP8_TOKENS_LOCKSTEP_H_
#define LIBWEBP_B200_VP8_TOKENS_LOCKSTEP_H_

#include "vp8_dev.h"
#include "vp8_parse_core.h"
#include "vp8_tokens_fsm.h"   // TK_FN, tk_saddr and the shared-memory access helpers

#if defined(__GNUC__) || defined(__CUDACC__)
#define TL_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define TL_UNLIKELY(x) (x)
#endif

// ---- states
#define TL_SIGN1 33    // sign of a coefficient of magnitude 1 (next context 1)
#define TL_SIGN2 34    // sign of a larger one (next context 2)
#define TL_C159 35
#define TL_C165 36
#define TL_C145 37
#define TL_CAT3 38     // 3 extra bits
#define TL_CAT4 41     // 4
#define TL_CAT5 45     // 5
#define TL_CAT6 50     // 11
#define TL_STATES 61
#define TL_ROW_BYTES 64
#define TL_TYPE_BYTES (16 * TL_ROW_BYTES)
#define TL_IMG_BYTES (4 * TL_TYPE_BYTES)
#define TL_IMG_STRIDE (TL_IMG_BYTES + 20)   // odd multiple of 4 banks: the lanes of a warp spread over the banks

// ---- transition entry. o = next s + 64 * (advance the row, i.e. n++): the offset of the next probability from the current
// row, and the index of the next state's entries in the table (whose two halves are alike).
//   [31:25] o          so that the probability's address is row + (e >> 25): one shift-and-add
//   [24:14] addend of v
//   [9:3]   o again    so that the entries' address is table + (e & 0x3f8): one AND, the table base rides in the load
//   1       end of block      0  emit +-v at n
#define TL_ADV 64u     // flag arguments of TL_E only
#define TL_EMIT 1u
#define TL_EOB 2u
#define TL_E(next, flags, add) \
  (((((uint32_t)(next)) | ((uint32_t)(flags) & TL_ADV)) << 25) | ((uint32_t)(add) << 14) | \
   ((((uint32_t)(next)) | ((uint32_t)(flags) & TL_ADV)) << 3) | ((uint32_t)(flags) & 3u))
#define TL_E_OFS(e) ((e) >> 25)
#define TL_E_TAB(e) ((e) & 0x3f8u)
#define TL_E_ADD(e) (((e) >> 14) & 0x7ffu)
#define TL_E_ROWSTEP(e) (((e) >> 31) << 6)

TK_FN uint32_t tl_trans_entry(int s, int b) {
  if (s < 33) {
    const int base = (s / 11) * 11, k = s % 11;
    switch (k) {
      case 0: return b ? TL_E(s + 1, 0, 0) : TL_E(s, TL_EOB, 0);
      case 1: return b ? TL_E(s + 1, 0, 0) : TL_E(1, TL_ADV, 0);            // a zero: next position, ctx 0, node 1
      case 2: return b ? TL_E(s + 1, 0, 0) : TL_E(TL_SIGN1, 0, 1);
      case 3: return b ? TL_E(base + 6, 0, 0) : TL_E(base + 4, 0, 0);
      case 4: return b ? TL_E(base + 5, 0, 0) : TL_E(TL_SIGN2, 0, 2);
      case 5: return b ? TL_E(TL_SIGN2, 0, 4) : TL_E(TL_SIGN2, 0, 3);
      case 6: return b ? TL_E(base + 8, 0, 0) : TL_E(base + 7, 0, 0);
      case 7: return b ? TL_E(TL_C165, 0, 7) : TL_E(TL_C159, 0, 5);
      case 8: return b ? TL_E(base + 10, 0, 0) : TL_E(base + 9, 0, 0);
      case 9: return b ? TL_E(TL_CAT4, 0, 3 + 16) : TL_E(TL_CAT3, 0, 3 + 8);
      default: return b ? TL_E(TL_CAT6, 0, 3 + 64) : TL_E(TL_CAT5, 0, 3 + 32);
    }
  }
  if (s == TL_SIGN1) return TL_E(11, TL_ADV | TL_EMIT, 0);
  if (s == TL_SIGN2) return TL_E(22, TL_ADV | TL_EMIT, 0);
  if (s == TL_C159) return TL_E(TL_SIGN2, 0, b);
  if (s == TL_C165) return TL_E(TL_C145, 0, 2 * b);
  if (s == TL_C145) return TL_E(TL_SIGN2, 0, b);
  int first, nb;
  if (s < TL_CAT4) { first = TL_CAT3; nb = 3; }
  else if (s < TL_CAT5) { first = TL_CAT4; nb = 4; }
  else if (s < TL_CAT6) { first = TL_CAT5; nb = 5; }
  else { first = TL_CAT6; nb = 11; }
  const int i = s - first;
  return TL_E(i == nb - 1 ? TL_SIGN2 : s + 1, 0, b << (nb - 1 - i));
}

// byte s of the row of (type t, position n)
TK_FN uint8_t tl_row_byte(const uint8_t* prob /* [4][8][3][11] */, int t, int n, int s) {
  const uint8_t fixed[28] = { 128, 128, 159, 165, 145, 173, 148, 140, 176, 155, 140, 135, 180, 157, 141, 134, 130,
                              254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129 };
  const uint8_t bands[16] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7 };
  if (s < 33) return prob[t * 264 + bands[n] * 33 + s];
  if (s < TL_STATES) return fixed[s - 33];
  return 0;
}

// Block-wide tables in shared memory.
struct TlTables {
  uint32_t trans[128][2];  // [s + 64 * adv][bit]: indexed by the low 7 bits of the entry that led here (both halves alike)
  uint32_t seqmask[28];    // block seq (0 = Y2, 1..16 luma, 17..24 chroma): its two context bits inside TlLane::cx
};
#define TLT_SEQMASK 1024
#define TL_TAB_BYTES 1136    // sizeof(TlTables)

TK_FN uint32_t tl_seqmask(int k) {
  if (k == 0) return (1u << 8) | (1u << 24);
  if (k <= 16) { const uint32_t blk = (uint32_t)k - 1; return (1u << (blk & 3)) | (1u << (16 + (blk >> 2))); }
  if (k <= 24) {
    const uint32_t c = (uint32_t)k - 17;
    return (1u << (4 + (c & 1) + 2 * (c >> 2))) | (1u << (16 + 4 + ((c >> 1) & 1) + 2 * (c >> 2)));
  }
  return 0;
}

TK_FN void tl_tables_fill(TlTables* t, int tid, int nthreads) {
  for (int k = tid; k < 256; k += nthreads) {
    const int st = (k >> 1) & 63;
    t->trans[k >> 1][k & 1] = st < TL_STATES ? tl_trans_entry(st, k & 1) : TL_E(63, 0, 0);   // 63 = TL_DEAD
  }
  for (int k = tid; k < 28; k += nthreads) t->seqmask[k] = tl_seqmask(k);
}

// One image's rows (TL_IMG_BYTES at `dst`, 4-byte aligned) from the parsed header; any thread subset.
TK_FN void tl_image_fill(uint8_t* dst, const FrameHdr* h, int tid, int nthreads) {
  for (int k = tid; k < TL_IMG_BYTES / 4; k += nthreads) {
    const int t = k >> 8, n = (k >> 4) & 15, s0 = (k & 15) * 4;
    uint32_t w = 0;
    for (int j = 0; j < 4; ++j) w |= (uint32_t)tl_row_byte(h->prob, t, n, s0 + j) << (8 * j);
    ((uint32_t*)dst)[k] = w;
  }
}

// Lane phases (TlLane::pend). Only running lanes execute a step; the others sit out the rest of their group and
// are looked after at the warp's next event point.
#define TL_RUN 0        // a block is set up, the pending decode is loaded
#define TL_BLOCK_END 1  // the block has ended: bookkeeping and the next block / macroblock are due
#define TL_NEED_MB 2    // the next macroblock has to be started (first one, or the row above is not ready yet)
#define TL_FINISHED 3

// ---- per-lane state (registers)
struct TlLane {
  BoolDec d;
  tk_saddr row, rowend;   // current row, row of position 16 of the current block type
  uint32_t s;             // state the current block started in (tl_prime); the walk itself lives in e0/e1
  uint32_t sink;          // see tl_step (never meaningful)
  uint32_t prob, e0, e1;  // the pending decode: its probability and the transition entries of its two outcomes
  uint32_t v;             // magnitude under construction
  uint32_t ofs;           // coefficient n of the current block goes to byte (row >> 5) + ofs of the macroblock's 800
  uint32_t cx;            // non-zero contexts: top in bits 0-8 (0-3 luma, 4-5 U, 6-7 V, 8 Y2), left in bits 16-24
  uint32_t acc_lo, acc_hi;// 2-bit nz codes of the macroblock's blocks shifted in, in parse order
  uint32_t m, m_next;     // context bits of the current / the next block (seqmask)
  uint32_t lut;           // nz -> 2-bit code of the current block: 2-bit fields indexed by min(nz, 4)
  int seq;                // 0 = Y2, 1..16 luma, 17..24 chroma
  // macroblock
  tk_saddr yrow, yend;    // luma rows of this macroblock: first one parsed, end
  uint32_t yofs, ylut;    // luma blocks: ofs - 32 * seq, lut
  int16_t* mbcoef;
  uint32_t w, w_next;     // MbInfo word 3 of this / the partition's next macroblock
  int mx, my;
  int done_mbs;
  int pend;               // TL_RUN / TL_BLOCK_END / TL_NEED_MB / TL_FINISHED
  BoolDec parked;         // see tl_lane_park
  uint32_t sv_V, sv_vlo, sv_R24; int sv_nbits, sv_shift;   // see tl_set_aside
  int waiting;            // P > 1: the partition owning the row above has not got far enough yet
  int alive;              // 0 once parked
  int status;
};
#define TL_LUT_FROM0 0x3a4u   // nz 0 -> 0, 1 -> 1 (a lone DC level: re-examined after dequantisation, recon_macroblock), 2,3 -> 2, >= 4 -> 3
#define TL_LUT_FROM1 0x3a0u   // luma blocks of i16 macroblocks start at coefficient 1: nz = 1 means empty

// Per-lane constants.
struct TlCtx {
  tk_saddr img_s;         // this image's rows
  tk_saddr tab_s;         // TlTables
  uint16_t* topctx;       // (P + 1) x ctx_stride ring
  volatile int* progress; // P counters
  uint32_t* mbinfo;       // this image's MbInfo
  int16_t* coeffs;        // this image's coefficient plane
  int mb_w, rows, P, part, use_skip, ctx_stride;
};

TK_FN void tl_lane_reset(TlLane& L, const TlCtx& c) {
  L.row = c.img_s; L.rowend = 0; L.s = 0; L.sink = 0; L.prob = 0; L.e0 = 0; L.e1 = 0; L.v = 0; L.ofs = 0; L.cx = 0;
  L.acc_lo = 0; L.acc_hi = 0; L.m = 0; L.m_next = 0; L.lut = 0; L.seq = 0;
  L.yrow = 0; L.yend = 0; L.yofs = 0; L.ylut = 0; L.mbcoef = c.coeffs;
  L.mx = 0; L.my = c.part; L.done_mbs = 0; L.waiting = 1; L.alive = 1; L.status = VP8B_OK; L.parked = L.d;
  L.sv_V = L.d.V; L.sv_vlo = L.d.vlo; L.sv_R24 = L.d.R24; L.sv_nbits = L.d.nbits; L.sv_shift = L.d.last_shift;
  L.pend = TL_NEED_MB;
  L.w = 0; L.w_next = 0;
}

TK_FN void tl_lane_init(TlLane& L, const TlCtx& c, const uint8_t* frame, const FrameHdr* h) {
  bd_init(L.d, frame + h->part_off[c.part], h->part_size[c.part]);
  tl_lane_reset(L, c);
  L.w_next = (c.part < c.rows) ? VP8_LDG(c.mbinfo + 4 * ((size_t)c.part * c.mb_w) + 3) : 0;
}

// Loads the pending decode (probability, both transition entries) of state L.s in row L.row.
TK_FN void tl_prime(TlLane& L, const TlCtx& c) {
  L.prob = tk_lds_u8(L.row + L.s);
  tk_lds_v2(c.tab_s + L.s * 8u, L.e0, L.e1);
}

TK_FN uint32_t tl_popc(uint32_t x) {
#if defined(__CUDACC__) && !defined(VP8_EMU)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}

// Sets up block L.seq >= 1 (contexts in L.cx are final for it). Luma blocks sit at 32 * (seq - 1) bytes of the
// macroblock's coefficients, chroma blocks follow them (blocks 16..23 = seq 17..24).
TK_FN void tl_block_setup(TlLane& L, const TlCtx& c) {
  const int chroma = L.seq >= 17;
  const tk_saddr crow = c.img_s + 2 * TL_TYPE_BYTES;
  L.m = L.m_next;   // fetched while the previous block was parsed: one shared-memory round trip less at the block end
  L.m_next = tk_lds_u32(c.tab_s + TLT_SEQMASK + 4u * (uint32_t)L.seq + 4u);
  L.rowend = chroma ? crow + TL_TYPE_BYTES : L.yend;
  L.row = chroma ? crow : L.yrow;
  L.lut = chroma ? TL_LUT_FROM0 : L.ylut;
  L.ofs = (uint32_t)L.seq * 32u + (chroma ? 0u - 32u - (uint32_t)(crow >> 5) : L.yofs);
  L.s = tl_popc(L.cx & L.m) * 11u;
  tl_prime(L, c);
}

// The Y2 block of an i16 macroblock (seq 0): type 1, block 24.
TK_FN void tl_y2_setup(TlLane& L, const TlCtx& c) {
  L.m = (1u << 8) | (1u << 24);
  L.m_next = (1u << 0) | (1u << 16);   // seq 1
  L.row = c.img_s + 1 * TL_TYPE_BYTES;
  L.rowend = L.row + TL_TYPE_BYTES;
  L.lut = TL_LUT_FROM0;
  L.ofs = 24u * 32u - (uint32_t)(L.row >> 5);
  L.s = tl_popc(L.cx & L.m) * 11u;
  tl_prime(L, c);
}

// Stores a finished (or skipped) macroblock's results and steps to the partition's next macroblock.
template <int MULTI>
TK_FN void tl_mb_store(TlLane& L, const TlCtx& c, uint32_t nzy, uint32_t w3) {
  const int P = MULTI ? c.P : 1, mb_w = c.mb_w;
  const size_t idx = (size_t)L.my * mb_w + L.mx;
  c.mbinfo[4 * idx + 2] = nzy;
  c.mbinfo[4 * idx + 3] = w3;
  const int ring_row = MULTI ? L.my % (P + 1) : (L.my & 1);
  c.topctx[(size_t)ring_row * c.ctx_stride + L.mx] = (uint16_t)(L.cx & 0x1ffu);
  L.done_mbs++;
  if (++L.mx == mb_w) { L.mx = 0; L.my += P; }
  if (bd_eof(L.d)) {
    // Ran past the end of the partition: the image is lost (vp8_dec.c:651-659); release whoever waits on us.
    L.status = VP8B_NOT_ENOUGH_DATA;
    if (MULTI) { TK_FENCE(); c.progress[c.part] = 0x7fffffff; }
  } else if (MULTI) { TK_FENCE(); c.progress[c.part] = L.done_mbs; }
}

// Leaves the lane either with a block set up (returns 1), waiting for the row above (returns 0, L.waiting = 1) or
// finished (returns 0, L.waiting = 0, L.my >= rows or L.status != OK). Skipped macroblocks are consumed here.
template <int MULTI>
TK_FN int tl_mb_next(TlLane& L, const TlCtx& c) {
  const int P = MULTI ? c.P : 1, mb_w = c.mb_w;
  for (;;) {
    if (L.my >= c.rows || L.status != VP8B_OK) { L.waiting = 0; return 0; }
    uint32_t tctx = 0;
    if (L.my > 0) {
      if (MULTI) {
        const int prev = (c.part + P - 1) % P;
        const int need = ((L.my - 1 - prev) / P) * mb_w + L.mx + 1;
        if (c.progress[prev] < need) { L.waiting = 1; return 0; }
        TK_FENCE();
      }
      const int ring_row = MULTI ? (L.my + P) % (P + 1) : ((L.my + 1) & 1);
      tctx = c.topctx[(size_t)ring_row * c.ctx_stride + L.mx];
    }
    L.waiting = 0;
    L.w = L.w_next;
    {   // flags of this partition's next macroblock: needed one macroblock from here
      int nx = L.mx + 1, ny = L.my;
      if (nx == mb_w) { nx = 0; ny += P; }
      if (ny < c.rows) L.w_next = VP8_LDG(c.mbinfo + 4 * ((size_t)ny * mb_w + nx) + 3);
    }
    if (L.mx == 0) L.cx = 0;
    L.cx = (L.cx & 0xffff0000u) | tctx;
    const int is_i4 = (L.w & MBW_I4X4) != 0;
    if (!(c.use_skip && (L.w & MBW_SKIP))) {
      L.acc_lo = 0; L.acc_hi = 0;
      L.mbcoef = c.coeffs + ((size_t)L.my * mb_w + L.mx) * VP8B_COEFFS_PER_MB;
      // luma: type 3 from coefficient 0 (i4x4) or type 0 from coefficient 1 (after the Y2 block)
      const tk_saddr ybase = c.img_s + (is_i4 ? 3u : 0u) * TL_TYPE_BYTES;
      L.yend = ybase + TL_TYPE_BYTES;
      L.yrow = ybase + (is_i4 ? 0u : (uint32_t)TL_ROW_BYTES);
      L.ylut = is_i4 ? TL_LUT_FROM0 : TL_LUT_FROM1;
      L.yofs = 0u - 32u - (uint32_t)(ybase >> 5);
      if (is_i4) { L.seq = 1; L.m_next = (1u << 0) | (1u << 16); tl_block_setup(L, c); } else { L.seq = 0; tl_y2_setup(L, c); }
      return 1;
    }
    L.cx &= is_i4 ? 0x01000100u : 0u;
    tl_mb_store<MULTI>(L, c, 0u, L.w & 0xffff0000u);
  }
}

// The macroblock's last block has ended: store its results, move on.
template <int MULTI>
TK_FN void tl_mb_finish(TlLane& L, const TlCtx& c) {
  const uint32_t nzy = (L.acc_hi << 16) | (L.acc_lo >> 16);
  const uint32_t uv = L.acc_lo & 0xffffu;                      // U codes in bits 15-8, V in 7-0
  const uint32_t nzuv = (uv >> 8) | ((uv & 0xffu) << 8);       // reference order: U bits 0-7, V bits 8-15
  uint32_t w = L.w;
  if ((L.acc_hi >> 16) & 3u) w |= MBW_HAS_Y2;                  // the Y2 block's code, shifted in first
  tl_mb_store<MULTI>(L, c, nzy, (w & 0xffff0000u) | nzuv);
}

// Parks a lane that has nothing (more) to do. Where every lane of the warp executes every step (tl_step_inline) a
// parked lane keeps decoding in state 63: a zero probability for ever, never emits, never ends a block, and its
// reader only shifts in zeros.
#define TL_DEAD 63u
TK_FN void tl_lane_park(TlLane& L, const TlCtx& c) {
  L.pend = TL_FINISHED; L.alive = 0; L.waiting = 0;
  L.parked = L.d;   // the reader as the lane left it (a launch that parses a band of rows hands it to the next one)
  L.s = TL_DEAD; L.row = c.img_s; L.rowend = 0;
  tl_prime(L, c);
}

// A lane without a stream: finished from the start (`any` = some valid address for its reader).
TK_FN void tl_lane_idle(TlLane& L, const TlCtx& c, const uint8_t* any) {
  bd_init(L.d, any, 0);
  tl_lane_reset(L, c);
  tl_lane_park(L, c);
}

// One boolean decode and the transition it selects; returns the transition entry. The caller has topped the window up
// (bd_fill_lookahead) within the last three decodes.
TK_FN uint32_t tl_decode(TlLane& L, const TlCtx& c) {
  // ---- what the NEXT decode needs, fetched for both outcomes of this one before its bit is known: the dependent
  // chain of an iteration is then select -> multiply -> compare, and the shared-memory latency runs beside it.
  // (ADV is bit 6 of o = the row stride, so `row + o` is the next probability's address as it stands.)
  uint32_t e00, e01, e10, e11;
  tk_lds_v2_pinned(c.tab_s + TL_E_TAB(L.e0), e00, e01);
  tk_lds_v2_pinned(c.tab_s + TL_E_TAB(L.e1), e10, e11);
  const uint32_t p0 = tk_lds_u8_pinned(L.row + TL_E_OFS(L.e0)), p1 = tk_lds_u8_pinned(L.row + TL_E_OFS(L.e1));
  // ---- boolean decode (bit_reader_inl_utils.h:107-136)
  const int bit = bd_bit_nofill(L.d, L.prob);
  // ---- transition
  const uint32_t e = bit ? L.e1 : L.e0;
  L.prob = bit ? p1 : p0;
  L.sink ^= p1;   // an unconditional use: without it the assembler folds the select into a load of p1 predicated on the
                  // bit, which puts the shared-memory latency straight back on the dependent chain
  L.e0 = bit ? e10 : e00;
  L.e1 = bit ? e11 : e01;
  L.v += TL_E_ADD(e);
  if (e & TL_EMIT) {   // level, parse order
    *(int16_t*)((uint8_t*)L.mbcoef + (uint32_t)((L.row >> 5) + L.ofs)) = (int16_t)(bit ? -(int)L.v : (int)L.v);
    L.v = 0;
  }
  L.row += TL_E_ROWSTEP(e);
  return e;
}

// ParseResiduals' bookkeeping at the end of a block (vp8_dec.c:517-609) with GetCoeffs' return value nz, then the next
// block. Returns 1 when the macroblock's last block has ended instead (tl_mb_finish + tl_mb_next are due).
TK_FN int tl_block_end(TlLane& L, const TlCtx& c) {
  const uint32_t nz2 = 32u - (uint32_t)((L.rowend - L.row) >> 5);   // 2 * nz
  const uint32_t code = (L.lut >> (nz2 < 8u ? nz2 : 8u)) & 3u;       // non-zero exactly when the block counts as non-empty
  L.acc_hi = (L.acc_hi << 2) | (L.acc_lo >> 30);
  L.acc_lo = (L.acc_lo << 2) | code;
  L.cx = (L.cx & ~L.m) | (code ? L.m : 0u);
  L.seq++;
  if (TL_UNLIKELY(L.seq == 25)) return 1;
  tl_block_setup(L, c);
  return 0;
}

// ---- first way of running the lanes: every lane executes every step, and a lane whose block has ended does its
// bookkeeping on the spot while the others wait. Best when a warp is alone on its SM sub-partition (one dependent
// chain to follow, nothing to overlap the bookkeeping with). Returns 0 once the lane has finished (parked).
template <int MULTI>
TK_FN int tl_step_inline(TlLane& L, const TlCtx& c) {
  if (MULTI) {
    if (L.waiting) {
      if (!tl_mb_next<MULTI>(L, c)) {
        if (!L.waiting) tl_lane_park(L, c);
        return L.alive;
      }
    }
  }
  const uint32_t e = tl_decode(L, c);
  if (TL_UNLIKELY((e & TL_EOB) || L.row == L.rowend)) {
    if (tl_block_end(L, c)) {
      tl_mb_finish<MULTI>(L, c);
      if (!tl_mb_next<MULTI>(L, c)) {
        if (!(MULTI && L.waiting)) tl_lane_park(L, c);
        return L.alive;
      }
    }
  }
  return 1;
}

// ---- second way: a step is branch-free and a block end only marks the lane, which then sits out the rest of its
// group of four steps; the bookkeeping of all marked lanes happens together at the group's event point. Fewer
// instructions per decode; best when several warps share the sub-partition.
TK_FN void tl_step(TlLane& L, const TlCtx& c) {
  const uint32_t e = tl_decode(L, c);
  if ((e & TL_EOB) || L.row == L.rowend) L.pend = TL_BLOCK_END;
}

// The warp's event point, once per group of steps: lanes whose block has ended do ParseResiduals' bookkeeping
// (vp8_dec.c:517-609) with GetCoeffs' return value nz and start their next block or macroblock; lanes that need a
// macroblock (MULTI = several token partitions: the row above may not be ready) try again.
template <int MULTI>
TK_FN void tl_events(TlLane& L, const TlCtx& c) {
  if (L.pend == TL_BLOCK_END || L.pend == TL_NEED_MB) {
    int need_mb = (L.pend == TL_NEED_MB);
    L.pend = TL_RUN;
    if (!need_mb && tl_block_end(L, c)) {
      tl_mb_finish<MULTI>(L, c);
      need_mb = 1;
    }
    if (TL_UNLIKELY(need_mb)) {
      if (!tl_mb_next<MULTI>(L, c)) {
        if (MULTI && L.waiting) L.pend = TL_NEED_MB; else tl_lane_park(L, c);
      }
    }
  }
}

// ---- third way: like the second, but a lane that is not running keeps executing the steps on a stand-in state
// instead of branching around them, so that the four steps of a group are straight-line code the scheduler can
// overlap. The stand-in: dead-state entries (no emit, no advance, no block end) and a reader whose real registers
// are set aside in TlLane::sv (the words it has fetched stay put: the top-up is skipped for such lanes).
TK_FN void tl_set_aside(TlLane& L) {
  L.sv_V = L.d.V; L.sv_vlo = L.d.vlo; L.sv_R24 = L.d.R24; L.sv_nbits = L.d.nbits; L.sv_shift = L.d.last_shift;
  L.prob = 0; L.e0 = TL_E(TL_DEAD, 0, 0); L.e1 = TL_E(TL_DEAD, 0, 0);
}
TK_FN void tl_take_back(TlLane& L) {
  L.d.V = L.sv_V; L.d.vlo = L.sv_vlo; L.d.R24 = L.sv_R24; L.d.nbits = L.sv_nbits; L.d.last_shift = L.sv_shift;
}

TK_FN void tl_step_flat(TlLane& L, const TlCtx& c) {
  const uint32_t e = tl_decode(L, c);
  // tl_set_aside() on a block end, written as selects: a branch here would cut the group into four basic blocks
  const bool done = (L.pend == TL_RUN) & (((e & TL_EOB) != 0u) | (L.row == L.rowend));
  L.pend = done ? TL_BLOCK_END : L.pend;
  L.sv_V = done ? L.d.V : L.sv_V; L.sv_vlo = done ? L.d.vlo : L.sv_vlo; L.sv_R24 = done ? L.d.R24 : L.sv_R24;
  L.sv_nbits = done ? L.d.nbits : L.sv_nbits; L.sv_shift = done ? L.d.last_shift : L.sv_shift;
  L.prob = done ? 0u : L.prob; L.e0 = done ? TL_E(TL_DEAD, 0, 0) : L.e0; L.e1 = done ? TL_E(TL_DEAD, 0, 0) : L.e1;
}

template <int MULTI>
TK_FN void tl_group_flat(TlLane& L, const TlCtx& c) {
  if (L.pend == TL_BLOCK_END || L.pend == TL_NEED_MB) {
    tl_take_back(L);
    tl_events<MULTI>(L, c);
    if (L.pend != TL_RUN) tl_set_aside(L);   // still waiting for the row above, or finished
  }
  if (L.pend == TL_RUN) bd_fill_lookahead(L.d);
  tl_step_flat(L, c); tl_step_flat(L, c); tl_step_flat(L, c); tl_step_flat(L, c);
}

// One group: the event point, a top-up of the window, four decodes.
template <int MULTI>
TK_FN void tl_group(TlLane& L, const TlCtx& c) {
  tl_events<MULTI>(L, c);
  if (L.pend == TL_RUN) bd_fill_lookahead(L.d);
  if (L.pend == TL_RUN) tl_step(L, c);
  if (L.pend == TL_RUN) tl_step(L, c);
  if (L.pend == TL_RUN) tl_step(L, c);
  if (L.pend == TL_RUN) tl_step(L, c);
}

#endif  // LIBWEBP_B200_VP8_TOKENS_LOCKSTEP_H_
