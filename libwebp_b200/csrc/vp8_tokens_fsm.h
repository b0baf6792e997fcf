// vp8_tokens_fsm.h -- coefficient-token parser as a lane-parallel finite-state machine.
//
// One LANE owns one token partition (one image when partitions = 1) and keeps the boolean decoder, the token
// tree position and the block/macroblock bookkeeping in registers. Every loop iteration performs exactly one
// boolean decode per lane, and what happens next is read from a 16-state x {0,1} transition table, so the
// lanes of a warp execute the same instruction stream although each is somewhere else in its own bitstream
// (GetCoeffs' nested loops and GetLargeValue's branches, src/dec/vp8_dec.c:411-469, become table rows).
// Only block ends (every ~18 decodes), the rare large-value categories and macroblock starts branch.
//
// The parsing warps never touch the compressed bytes in HBM: a PRODUCER warp of the same thread block streams
// every partition through a 256-byte shared-memory ring with 16-byte cp.async copies, far ahead of the reader,
// so no parsing warp ever waits on a global load (a warp-wide scoreboard stall would hold back all its lanes).
// Ring protocol per stream: the reader publishes the next word it will read (rd_w), the producer the first
// 16-byte chunk that is not in the ring yet (filled_c); chunk c lives in slot c & 15.
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
TK_FN int32_t tk_lds_s16(tk_saddr a) { int32_t v; asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
TK_FN uint32_t tk_lds_u32(tk_saddr a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
TK_FN void tk_lds_v2(tk_saddr a, uint32_t& x, uint32_t& y) { asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
// ring words and the ring protocol words change under the reader's feet: volatile accesses
TK_FN uint32_t tk_ldsv_u32(tk_saddr a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
TK_FN void tk_stsv_u32(tk_saddr a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
TK_FN uint32_t tk_ldsv_u8(tk_saddr a) { uint32_t v; asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
TK_FN void tk_stsv_u8(tk_saddr a, uint32_t v) { asm volatile("st.volatile.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// loads that must stay where they are written (look-ahead loads: the compiler would otherwise sink them under the
// condition that selects their result)
TK_FN uint32_t tk_lds_u8_pinned(tk_saddr a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
TK_FN void tk_lds_v2_pinned(tk_saddr a, uint32_t& x, uint32_t& y) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
TK_FN uint32_t tk_shr_clamp(uint32_t w, int n) { return __funnelshift_rc(w, 0u, (uint32_t)n); }   // w >> n, 0 when n >= 32
TK_FN uint32_t tk_shl_clamp(uint32_t w, int n) { return __funnelshift_lc(0u, w, (uint32_t)n); }   // w << n, 0 when n >= 32
TK_FN uint32_t tk_shl_pair(uint32_t hi, uint32_t lo, int n) { return __funnelshift_l(lo, hi, (uint32_t)n); }
TK_FN void tk_copy16(tk_saddr dst, const uint8_t* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
TK_FN void tk_copy_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TK_FN uint32_t tk_ldg_u32(const uint32_t* p) { return __ldg(p); }
#else
#include <string.h>
#define TK_FN static inline
#define TK_CLZ(x) __builtin_clz((unsigned)(x))
#define TK_BSWAP(x) __builtin_bswap32(x)
#define TK_FENCE() ((void)0)
typedef uintptr_t tk_saddr;
TK_FN tk_saddr tk_saddr_of(const void* p) { return (tk_saddr)p; }
TK_FN uint32_t tk_lds_u8(tk_saddr a) { return *(const uint8_t*)a; }
TK_FN uint32_t tk_lds_u16(tk_saddr a) { return *(const uint16_t*)a; }
TK_FN int32_t tk_lds_s16(tk_saddr a) { return *(const int16_t*)a; }
TK_FN uint32_t tk_lds_u32(tk_saddr a) { return *(const uint32_t*)a; }
TK_FN void tk_lds_v2(tk_saddr a, uint32_t& x, uint32_t& y) { x = ((const uint32_t*)a)[0]; y = ((const uint32_t*)a)[1]; }
TK_FN uint32_t tk_ldsv_u32(tk_saddr a) { return *(volatile const uint32_t*)a; }
TK_FN void tk_stsv_u32(tk_saddr a, uint32_t v) { *(volatile uint32_t*)a = v; }
TK_FN uint32_t tk_ldsv_u8(tk_saddr a) { return *(volatile const uint8_t*)a; }
TK_FN void tk_stsv_u8(tk_saddr a, uint32_t v) { *(volatile uint8_t*)a = (uint8_t)v; }
TK_FN uint32_t tk_lds_u8_pinned(tk_saddr a) { return *(const uint8_t*)a; }
TK_FN void tk_lds_v2_pinned(tk_saddr a, uint32_t& x, uint32_t& y) { x = ((const uint32_t*)a)[0]; y = ((const uint32_t*)a)[1]; }
TK_FN uint32_t tk_shr_clamp(uint32_t w, int n) { return n >= 32 ? 0u : (w >> n); }
TK_FN uint32_t tk_shl_clamp(uint32_t w, int n) { return n >= 32 ? 0u : (w << n); }
TK_FN uint32_t tk_shl_pair(uint32_t hi, uint32_t lo, int n) { return n == 0 ? hi : ((hi << n) | (lo >> (32 - n))); }
TK_FN void tk_copy16(tk_saddr dst, const uint8_t* src) { memcpy((void*)dst, src, 16); }
TK_FN void tk_copy_wait() {}
TK_FN uint32_t tk_ldg_u32(const uint32_t* p) { return *p; }
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

// ---- block sequence word, one per block of a macroblock in parse order (seq 0 = Y2, 1..16 luma, 17..24 chroma)
//  [3:0] bit of the top context  [7:4] bit of the left context  [12:8] shift of the 2-bit nz code
//  13 chroma  14 luma  [16:15] block type when the macroblock is i16 (luma of i4x4 macroblocks: type 3)
//  [18:17] dequantiser pair  [23:19] block index inside the macroblock's coefficients
#define TQ_TB(q) ((q) & 15u)
#define TQ_LB(q) (((q) >> 4) & 15u)
#define TQ_NZSH(q) (((q) >> 8) & 31u)
#define TQ_CHROMA (1u << 13)
#define TQ_LUMA (1u << 14)
#define TQ_TYPE(q) (((q) >> 15) & 3u)
#define TQ_QPAIR(q) (((q) >> 17) & 3u)
#define TQ_BLK(q) (((q) >> 19) & 31u)

// Block-wide constant tables in shared memory (filled by tk_tables_fill).
struct TokTables {
  uint32_t trans[16][2];   // 128 B
  uint16_t band_off[18];   // band(n) * 33 for n = 0..16 (+ one spare), at byte offset 128
  uint8_t zigzag[16];      // at byte offset 164
  uint8_t pad[12];
  uint32_t seq[28];        // at byte offset 192
};
#define TKT_BAND 128
#define TKT_ZIGZAG 164
#define TKT_SEQ 192
#define TOK_TAB_BYTES 304    // sizeof(TokTables)

// Shared-memory block of one image: probabilities, token constants, dequantisers, flags.
struct TokImage {
  uint8_t prob[1056];          // [type][band][ctx][node]
  uint8_t consts[TKC_BYTES];   // at offset 1056
  int16_t dq[4][6];            // per segment: y1 dc/ac, y2 dc/ac, uv dc/ac; at offset 1084
  int32_t ok;                  // header parsed and the partition count matches the launch
  int32_t use_skip;
  int32_t pad[3];
};
#define TKI_CONSTS 1056
#define TKI_DQ 1084
#define TOK_IMG_BYTES 1152   // sizeof(TokImage)

// Shared-memory ring of one stream and its protocol words.
#define TK_RING_BYTES 256
#define TK_RING_CHUNKS 16
#define TK_STREAM_DONE 0xffffffffu
struct TokStreamCtl {
  uint32_t rd_w;       // reader -> producer: next arena word the reader will fetch; TK_STREAM_DONE when finished
  uint32_t filled_c;   // producer -> reader: first 16-byte chunk (absolute arena chunk index) not yet in the ring
};

// Per-lane state (registers). phase: 0 = needs a macroblock, 1 = decoding, 2 = finished.
struct TokLane {
  // boolean decoder: 64-bit left-aligned window vhi:vlo, range stored minus one
  uint32_t vhi, vlo;
  int nbits;
  uint32_t range;
  uint32_t wi;        // next word of the stream, as an index into the arena viewed as uint32[]
  uint32_t avail_w;   // words below this index are known to be in the ring
  int shift;          // renormalisation shift of the most recent decode
  tk_saddr ring_s;    // this stream's ring
  tk_saddr ctl_s;     // this stream's TokStreamCtl
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
  uint32_t sq;        // block sequence word of `seq`
  int first;
  uint32_t tnz, lnz;  // bit 0-3 luma, 4-5 U, 6-7 V, 8 Y2
  uint32_t nzy, nzuv;
  uint32_t outi;      // index of the current block's first coefficient inside the image's coefficient plane
  // macroblock / partition
  int mx, my, part, done_mbs, phase, status;
  uint32_t w;         // MbInfo word 3 of the current macroblock
  uint32_t w_next;    // ... of this partition's next macroblock (fetched one macroblock ahead)
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
  for (int k = tid; k < 28; k += nthreads) {
    uint32_t q = 0;
    if (k == 0) {                      // Y2: context bit 8 on both sides, type 1, y2 dequantisers, block 24
      q = 8u | (8u << 4) | (1u << 15) | (1u << 17) | (24u << 19);
    } else if (k <= 16) {              // luma block k-1
      const uint32_t blk = (uint32_t)k - 1;
      q = (blk & 3) | ((blk >> 2) << 4) | ((30 - 2 * blk) << 8) | TQ_LUMA | (0u << 15) | (0u << 17) | (blk << 19);
    } else if (k <= 24) {              // chroma block c: U 0-3, V 4-7
      const uint32_t c = (uint32_t)k - 17, blk = 16 + c;
      const uint32_t tb = 4 + (c & 1) + 2 * (c >> 2), lb = 4 + ((c >> 1) & 1) + 2 * (c >> 2);
      q = tb | (lb << 4) | ((8 * (c >> 2) + 6 - 2 * (c & 3)) << 8) | TQ_CHROMA | (2u << 15) | (2u << 17) | (blk << 19);
    }
    t->seq[k] = q;
  }
}

// Fills the shared per-image block from the parsed frame header (any thread subset may call with its slice).
TK_FN void tk_image_fill(TokImage* im, const FrameHdr* h, int P, int tid, int nthreads) {
  const uint8_t consts[TKC_BYTES] = TK_CONST_INIT;
  for (int k = tid; k < 264; k += nthreads) ((uint32_t*)im->prob)[k] = ((const uint32_t*)h->prob)[k];
  for (int k = tid; k < TKC_BYTES; k += nthreads) im->consts[k] = consts[k];
  for (int k = tid; k < 24; k += nthreads) im->dq[k / 6][k % 6] = h->dq[k / 6][k % 6];
  if (tid == 0) { im->ok = (h->status == VP8B_OK && h->num_parts == P) ? 1 : 0; im->use_skip = h->use_skip; }
}

// ---- producer side of the ring -------------------------------------------------------------------------
// First fill of one chunk slot (k = 0..15) of a stream starting at arena byte `a`; every thread of the block
// takes some (stream, k) pairs, then tk_copy_wait() + a block barrier, then tk_stream_open() by one thread.
TK_FN void tk_stream_prefill(tk_saddr ring_s, const uint8_t* arena, uint64_t a, int k) {
  const uint64_t c = (a >> 4) + (uint64_t)k;
  tk_copy16(ring_s + (uint32_t)((c & (TK_RING_CHUNKS - 1)) * 16), arena + 16 * c);
}
TK_FN void tk_stream_open(TokStreamCtl* ctl, uint64_t a) {
  ctl->rd_w = (uint32_t)(a >> 2);
  ctl->filled_c = (uint32_t)(a >> 4) + TK_RING_CHUNKS;
}
// One top-up of one stream by its producer lane: issues the copies of every chunk whose slot the reader has
// left, returns the new filled_c (to be published with tk_stsv_u32 after tk_copy_wait() and a fence), or 0
// when the reader has finished. *filled = the value published last time.
TK_FN uint32_t tk_stream_topup(tk_saddr ctl_s, tk_saddr ring_s, const uint8_t* arena, uint32_t filled) {
  const uint32_t rd_w = tk_ldsv_u32(ctl_s);
  if (rd_w == TK_STREAM_DONE) return 0;
  const uint32_t lim = (rd_w >> 2) + TK_RING_CHUNKS;   // chunk c overwrites chunk c-16, which must lie behind the reader
  uint32_t c = filled;
  for (; c < lim; ++c) tk_copy16(ring_s + (c & (TK_RING_CHUNKS - 1)) * 16, arena + 16 * (uint64_t)c);
  return c;
}

// ---- reader side -----------------------------------------------------------------------------------------
// Lane for token partition `part` of an image. frame_off = byte offset of the frame tag inside the arena.
// The arena is padded by >= 32 KB so that neither a lane that ran past the end of its stream (detected at the
// next macroblock boundary) nor the producer's read-ahead ever leaves the allocation.
TK_FN uint64_t tk_stream_start(uint64_t frame_off, const FrameHdr* h, int part) { return frame_off + h->part_off[part]; }

TK_FN void tk_lane_init(TokLane& L, tk_saddr ring_s, tk_saddr ctl_s, uint64_t frame_off, const FrameHdr* h, int part,
                        const uint32_t* mbinfo, int mb_w, int mb_h) {
  const uint64_t a = tk_stream_start(frame_off, h, part);
  const uint32_t size = h->part_size[part];
  const int off = (int)(a & 3);
  L.ring_s = ring_s; L.ctl_s = ctl_s;
  L.wi = (uint32_t)(a >> 2);
  L.avail_w = ((uint32_t)(a >> 4) + TK_RING_CHUNKS) * 4;
  const uint32_t w0 = TK_BSWAP(tk_ldsv_u32(ring_s + ((L.wi & (TK_RING_BYTES / 4 - 1)) << 2)));
  L.wi++;
  tk_stsv_u32(ctl_s, L.wi);
  L.vhi = (off == 0) ? w0 : (w0 << (8 * off));
  L.vlo = 0;
  L.nbits = 32 - 8 * off;
  L.range = 254;
  L.shift = 0;
  L.pos_bias = 32 * (int64_t)(a >> 2) + 8 * off;
  L.limit = 8 * (int64_t)size - 8;
  L.state8 = 0; L.pp = 0; L.pnb = 0; L.pbase = 0;
  L.v = 0; L.n = 0; L.nc11 = 0; L.extra_left = 0; L.cat = 0;
  L.seq = 0; L.sq = 0; L.first = 0; L.tnz = 0; L.lnz = 0; L.nzy = 0; L.nzuv = 0; L.outi = 0;
  L.mx = 0; L.my = part; L.part = part; L.done_mbs = 0; L.phase = 0; L.status = VP8B_OK; L.w = 0; L.seg = 0;
  L.w_next = (part < mb_h) ? tk_ldg_u32(mbinfo + 4 * ((size_t)part * mb_w) + 3) : 0;
}

// The reference's eof_ flag from bit positions (see vp8_parse_core.h:bd_eof).
TK_FN int tk_eof(const TokLane& L) {
  const int64_t loaded = 32 * (int64_t)L.wi - L.pos_bias;
  return (loaded - L.nbits - L.shift) > L.limit;
}

TK_FN void tk_lane_finish(TokLane& L) {
  L.phase = 2;
  tk_stsv_u32(L.ctl_s, TK_STREAM_DONE);
}

// Sets up block L.seq of the current macroblock (context, probabilities, dequantisers, output index).
TK_FN void tk_block_setup(TokLane& L, const TokShared& sh, uint32_t mb_coef_index) {
  const uint32_t q = tk_lds_u32(sh.tab_s + TKT_SEQ + 4u * (uint32_t)L.seq);
  const uint32_t i4_luma = ((q & TQ_LUMA) && (L.w & MBW_I4X4)) ? 1u : 0u;
  const uint32_t i16_luma = (q & TQ_LUMA) ? 1u - i4_luma : 0u;
  const uint32_t type = i4_luma ? 3u : TQ_TYPE(q);
  const uint32_t ctx = ((L.tnz >> TQ_TB(q)) & 1u) + ((L.lnz >> TQ_LB(q)) & 1u);
  L.sq = q;
  L.first = (int)i16_luma;
  L.n = (int)i16_luma;
  L.pbase = sh.img_s + type * 264u;
  L.pp = L.pbase + i16_luma * 33u + ctx * 11u;
  L.pnb = L.pbase + (i16_luma ? 66u : 33u);   // band(n + 1) * 33 for n = 1 / n = 0
  L.state8 = S_P0 * 8;
  L.outi = mb_coef_index + TQ_BLK(q) * 16u;
}

// Finishes block L.seq with return value nz (GetCoeffs), updates contexts and nz codes (vp8_dec.c:517-609).
TK_FN void tk_block_end(TokLane& L, int nz) {
  const uint32_t q = L.sq;
  const uint32_t tb = TQ_TB(q), lb = TQ_LB(q);
  const uint32_t l = (nz > L.first) ? 1u : 0u;
  const uint32_t code = ((nz > 3) ? 3u : (nz > 1) ? 2u : l) << TQ_NZSH(q);   // lone DC: re-examined by recon_macroblock
  if (q & TQ_LUMA) L.nzy |= code;
  if (q & TQ_CHROMA) L.nzuv |= code;
  if (!(q & (TQ_LUMA | TQ_CHROMA)) && nz > 0) L.w |= MBW_HAS_Y2;
  L.tnz = (L.tnz & ~(1u << tb)) | (l << tb);
  L.lnz = (L.lnz & ~(1u << lb)) | (l << lb);
  L.seq++;
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
    tk_lane_finish(L);
    if (P > 1) { TK_FENCE(); sh.progress[L.part] = 0x7fffffff; }
    return;
  }
  if (P > 1) { TK_FENCE(); sh.progress[L.part] = L.done_mbs; }
}

// Starts the next macroblock of this lane's partition, or finishes the lane. Returns without doing anything
// when the partition owning the row above has not got far enough yet (the caller simply retries).
TK_FN void tk_mb_start(TokLane& L, const TokShared& sh, const ImgDesc& im, int rows, int P, uint32_t* mbinfo) {
  const int mb_w = im.mb_w;
  if (L.my >= rows) { tk_lane_finish(L); return; }
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
  L.w = L.w_next;
  {   // fetch the flags of this partition's next macroblock now; they are needed one macroblock from here
    int nx = L.mx + 1, ny = L.my;
    if (nx == mb_w) { nx = 0; ny += P; }
    if (ny < rows) L.w_next = tk_ldg_u32(mbinfo + 4 * ((size_t)ny * mb_w + nx) + 3);
  }
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
//   coeffs : this image's coefficient plane (int16 levels in parse order, pre-zeroed)
TK_FN void tk_step(TokLane& L, const TokShared& sh, const ImgDesc& im, int P, uint32_t* mbinfo, int16_t* coeffs) {
  const uint32_t prob = tk_lds_u8(L.pp);
  uint32_t e0, e1;
  tk_lds_v2(sh.tab_s + L.state8, e0, e1);
  // ---- boolean decode (bit_reader_inl_utils.h:107-136, range kept minus one)
  if (L.nbits <= 32) {   // vlo is empty: append one big-endian word behind the valid bits
    if (L.wi >= L.avail_w) {   // first look at what the producer has published since
      L.avail_w = tk_ldsv_u32(L.ctl_s + 4) * 4u;
      if (L.wi >= L.avail_w) return;   // the ring has run dry (start-up only): try again next iteration
      TK_FENCE();
    }
    const uint32_t w = TK_BSWAP(tk_ldsv_u32(L.ring_s + ((L.wi & (TK_RING_BYTES / 4 - 1)) << 2)));
    L.wi++;
    tk_stsv_u32(L.ctl_s, L.wi);
    L.vhi |= tk_shr_clamp(w, L.nbits);
    L.vlo = tk_shl_clamp(w, 32 - L.nbits);
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
  tk_saddr pp = ((e & TE_ABS) ? (sh.img_s + TKI_CONSTS) : L.pp) + TE_OFF(e);
  const uint32_t cs = TE_CTX(e);
  L.nc11 = cs ? cs * 11u : L.nc11;
  if (e & (TE_CAT | TE_EXTRA)) {   // DCT_CAT3..6: rare
    if (e & TE_CAT) {
      L.cat = (int)TE_CATV(e);
      L.extra_left = (L.cat == 3) ? 11 : 3 + L.cat;
    } else if (--L.extra_left == 0) {
      L.v += 3 + (8 << L.cat);
      L.state8 = S_SIGN * 8;
      pp = sh.img_s + TKI_CONSTS + TKC_SIGN;
    }
  }
  int done = -1;
  if (e & TE_EMIT) coeffs[L.outi + (uint32_t)L.n] = (int16_t)((e & TE_NEG) ? -L.v : L.v);   // level, parse order
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
