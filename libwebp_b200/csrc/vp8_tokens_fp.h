// vp8_tokens_fp.h -- coefficient-token parser, fourth mapping: lockstep lanes, fp32 boolean decoder, token stream out.
//
// Same shape as vp8_tokens_lockstep.h (one LANE per token partition, one boolean decode per lane per step, the token
// syntax of GetCoeffs / GetLargeValue, src/dec/vp8_dec.c:411-469, as a table of states), rebuilt around what
// tools/chain_floor.cu measured on the B200 (profiles/r02b_chain_floor.json):
//
//  * One warp alone on an SM sub-partition issues an integer-pipe instruction every 2 cycles whatever the number of
//    active lanes, and the chain of one decode through IMAD.HI (9 cycles) and FLO (17 cycles) takes 58 cycles. The
//    decoder here keeps the range as two floats and gets the split from one FFMA.RZ and the renormalisation shift
//    from the exponent field (bit-exact: every quantity is an integer below 2^24; 3 * 10^9 random decodes against
//    the integer form, zero differences):
//        everything is kept scaled by S = 2^-102, so that the exponent field of the new range IS the shift count:
//        Ra = (range-1) * 2^39, Rp = S * ((range-1) + 2^23), pd = the probability byte read as a DENORMAL float (p * 2^-149:
//        no integer-to-float conversion at all; FFMA takes denormals at full speed)
//        m  = fma.rz(Ra, pd, S * 2^23)      = S * (2^23 + split)           (rz truncates: floor; 2^39 * 2^-149 = S / 256)
//        s1 = bits(m) * 2^24 + 2^24         = (split + 1) << 24            (the exponent byte falls off the top)
//        bit = V >= s1;  V = min(V, V - s1) (the subtraction wraps exactly when the bit is 0)
//        f  = bit ? Rp - m : m - S*(2^23-1) = S * new range                (range-1-split, or split+1)
//        k  = bits(f) >> 23 = 32 - shift:   V:vlo = (V:vlo) >> k as a clamped funnel shift, no subtraction
//        new Ra = (mantissa of f under exponent 2^46) - 2^39, new Rp = fma(that, 2^-141, S * (2^23-1))
//  * A step spends its integer-pipe slots on selects, so the rest is moved to the FMA pipe or dropped:
//      - the magnitude is not assembled per step: `acc += entry` (one IMAD.IADD) adds the whole transition entry, whose
//        fields are laid out so that the junk of the low fields never carries into the addend field; the magnitude
//        is masked out at the emit
//      - no row pointer: the address of the pending probability is the state, an entry carries the distance to the
//        next one (rows are 64 bytes apart, so the coefficient position is bits 6-9 of that address)
//  * Levels are not stored into a zero-filled 800-byte plane any more. Every non-zero level is appended as one
//    32-bit token {sign, block, position, magnitude} to the partition's token stream (worst case 384 per macroblock
//    reserved, only what is written is ever touched), and MbTok[mb] = {first token, count}: the reconstruction
//    kernel scatters a macroblock's tokens into shared memory (vp8_pixel_core.h:recon_load_tokens).
//
// Replaces VP8DecodeMB / ParseResiduals / GetCoeffs / GetLargeValue (src/dec/vp8_dec.c:400-635) for a batch.
// Dual build like the other cores: nvcc for the product, g++ -DVP8_EMU for tests/emu.
#ifndef LIBWEBP_B200_VP8_TOKENS_FP_H_
#define LIBWEBP_B200_VP8_TOKENS_FP_H_

#include "vp8_dev.h"
#include "vp8_parse_core.h"
#include "vp8_tokens_fsm.h"   // TK_FN, tk_saddr and the shared-memory access helpers

#if defined(__GNUC__) || defined(__CUDACC__)
#define TF_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define TF_UNLIKELY(x) (x)
#endif

// ---- states: 0..32 = [ctx 3][node 11] of the position's band; then the fixed probabilities. Numbered so that every
// transition that stays at its position moves UP (the distance field of an entry is unsigned).
#define TF_C159 33
#define TF_C165 34
#define TF_C145 35
#define TF_CAT3 36     // 3 extra bits
#define TF_CAT4 39     // 4
#define TF_CAT5 43     // 5
#define TF_CAT6 48     // 11
#define TF_SIGN1 59    // sign of a coefficient of magnitude 1 (next context 1)
#define TF_SIGN2 60    // sign of a larger one (next context 2)
#define TF_STATES 61
#define TF_DEAD 63u
#define TF_ROW_BYTES 64
#define TF_TYPE_BYTES (16 * TF_ROW_BYTES)   // 1024: the position of a probability is bits 6-9 of its address
#define TF_IMG_BYTES (4 * TF_TYPE_BYTES)    // per image, 1024-byte aligned
// Second layout (BAND = 1 below): one row per probability BAND instead of per position -- 8 rows per type, 2 KB per image.
// Positions 7..14 (and 4) share band 6, so the 16-row layout holds that row nine times; with tens of thousands of small
// images in a batch the 4 KB per image are what limits the lanes per SM (227 KB: 54 images), and half the size doubles
// them. The price: the position no longer sits in the address, so the lane counts it (pos6), and the step to the next
// position's row is no longer a constant 64 bytes but 64 * (band[n+1] - band[n]), looked up per position (TfTables::delta)
// one decode ahead. All of it off the dependent chain (both outcomes' addresses are formed before the bit is known), but
// six more instructions per decode: only launches with more streams than the 16-row layout can seat use it.
#define TF_TYPE_BYTES_B(BAND) ((BAND) ? 8 * TF_ROW_BYTES : 16 * TF_ROW_BYTES)
#define TF_IMG_BYTES_B(BAND) (4 * TF_TYPE_BYTES_B(BAND))

// ---- transition entry
//   [31:25] distance in bytes from the pending probability to the next one: next s + 64 * (n advances) - s
//   [24:13] addend of the magnitude          (the sums of the fields below stay under 2^13 between two resets of acc)
//   [8:3]   next s, i.e. the offset of the next state's entry pair in the table
//   2       the walk moves to the next position (the 64 of the distance; only the banded layout looks at it)
//   1       end of block      0  emit at the current position
#define TF_EMIT 1u
#define TF_EOB 2u
#define TF_ADV 4u
#define TF_ADD_SHIFT 13
#define TF_ADD_MASK 0x01ffe000u
#define TF_E(s, next, adv, flags, add) \
  ((((uint32_t)(next) + ((adv) ? 64u : 0u) - (uint32_t)(s)) << 25) | ((uint32_t)(add) << TF_ADD_SHIFT) | ((uint32_t)(next) << 3) | ((adv) ? TF_ADV : 0u) | (uint32_t)(flags))
#define TF_E_DIST(e) ((e) >> 25)
#define TF_E_TAB(e) ((e) & 0x1f8u)

TK_FN uint32_t tf_trans_entry(int s, int b) {
  if (s < 33) {
    const int base = (s / 11) * 11, k = s % 11;
    switch (k) {
      case 0: return b ? TF_E(s, s + 1, 0, 0, 0) : TF_E(s, s, 0, TF_EOB, 0);
      case 1: return b ? TF_E(s, s + 1, 0, 0, 0) : TF_E(s, 1, 1, 0, 0);            // a zero: next position, ctx 0, node 1
      case 2: return b ? TF_E(s, s + 1, 0, 0, 0) : TF_E(s, TF_SIGN1, 0, 0, 1);
      case 3: return b ? TF_E(s, base + 6, 0, 0, 0) : TF_E(s, base + 4, 0, 0, 0);
      case 4: return b ? TF_E(s, base + 5, 0, 0, 0) : TF_E(s, TF_SIGN2, 0, 0, 2);
      case 5: return b ? TF_E(s, TF_SIGN2, 0, 0, 4) : TF_E(s, TF_SIGN2, 0, 0, 3);
      case 6: return b ? TF_E(s, base + 8, 0, 0, 0) : TF_E(s, base + 7, 0, 0, 0);
      case 7: return b ? TF_E(s, TF_C165, 0, 0, 7) : TF_E(s, TF_C159, 0, 0, 5);
      case 8: return b ? TF_E(s, base + 10, 0, 0, 0) : TF_E(s, base + 9, 0, 0, 0);
      case 9: return b ? TF_E(s, TF_CAT4, 0, 0, 3 + 16) : TF_E(s, TF_CAT3, 0, 0, 3 + 8);
      default: return b ? TF_E(s, TF_CAT6, 0, 0, 3 + 64) : TF_E(s, TF_CAT5, 0, 0, 3 + 32);
    }
  }
  if (s == TF_SIGN1) return TF_E(s, 11, 1, TF_EMIT, 0);
  if (s == TF_SIGN2) return TF_E(s, 22, 1, TF_EMIT, 0);
  if (s == TF_C159) return TF_E(s, TF_SIGN2, 0, 0, b);
  if (s == TF_C165) return TF_E(s, TF_C145, 0, 0, 2 * b);
  if (s == TF_C145) return TF_E(s, TF_SIGN2, 0, 0, b);
  int first, nb;
  if (s < TF_CAT4) { first = TF_CAT3; nb = 3; }
  else if (s < TF_CAT5) { first = TF_CAT4; nb = 4; }
  else if (s < TF_CAT6) { first = TF_CAT5; nb = 5; }
  else { first = TF_CAT6; nb = 11; }
  const int i = s - first;
  return TF_E(s, i == nb - 1 ? TF_SIGN2 : s + 1, 0, 0, b << (nb - 1 - i));
}

// byte s of the row of (type t, position n)
TK_FN uint8_t tf_row_byte(const uint8_t* prob /* [4][8][3][11] */, int t, int n, int s) {
  const uint8_t fixed[28] = { 159, 165, 145, 173, 148, 140, 176, 155, 140, 135, 180, 157, 141, 134, 130,
                              254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 128, 128 };
  const uint8_t bands[16] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7 };
  if (s < 33) return prob[t * 264 + (n < 0 ? -1 - n : bands[n]) * 33 + s];   // n < 0: band -1 - n itself (banded layout)
  if (s < TF_STATES) return fixed[s - 33];
  return 0;
}

// Block-wide tables in shared memory.
struct TfTables {
  uint32_t trans[64][2];   // [s][bit]; states 61..63 are dead (probability 0 for ever, nothing emitted, no block end)
  uint32_t seqmask[28];    // block seq (0 = Y2, 1..16 luma, 17..24 chroma): its two context bits inside TfLane::cx
  int16_t delta[16];       // banded layout: bytes from the row of position n to the row of position n + 1, minus the 64 the
                           // transition entries already carry
};
#define TFT_SEQMASK 512
#define TFT_DELTA 624
#define TF_TAB_BYTES 656    // sizeof(TfTables)

TK_FN uint32_t tf_seqmask(int k) {
  if (k == 0) return (1u << 8) | (1u << 24);
  if (k <= 16) { const uint32_t blk = (uint32_t)k - 1; return (1u << (blk & 3)) | (1u << (16 + (blk >> 2))); }
  if (k <= 24) {
    const uint32_t c = (uint32_t)k - 17;
    return (1u << (4 + (c & 1) + 2 * (c >> 2))) | (1u << (16 + 4 + ((c >> 1) & 1) + 2 * (c >> 2)));
  }
  return 0;
}

TK_FN void tf_tables_fill(TfTables* t, int tid, int nthreads) {
  for (int k = tid; k < 128; k += nthreads) {
    const int st = k >> 1;
    t->trans[st][k & 1] = st < TF_STATES ? tf_trans_entry(st, k & 1) : TF_E(st, st, 0, 0, 0);
  }
  for (int k = tid; k < 28; k += nthreads) t->seqmask[k] = tf_seqmask(k);
  for (int k = tid; k < 16; k += nthreads) {
    const int bands[17] = { 0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 8 };   // "position 16" lies one row past band 7
    t->delta[k] = (int16_t)(TF_ROW_BYTES * (bands[k + 1] - bands[k]) - 64);
  }
}

// One image's rows (TF_IMG_BYTES at `dst`, 1024-byte aligned) from the parsed header; any thread subset.
TK_FN void tf_image_fill(uint8_t* dst, const FrameHdr* h, int tid, int nthreads, int band = 0) {
  for (int k = tid; k < TF_IMG_BYTES_B(band) / 4; k += nthreads) {
    const int t = band ? k >> 7 : k >> 8, n = band ? -1 - ((k >> 4) & 7) : (k >> 4) & 15, s0 = (k & 15) * 4;
    uint32_t w = 0;
    for (int j = 0; j < 4; ++j) w |= (uint32_t)tf_row_byte(h->prob, t, n, s0 + j) << (8 * j);
    ((uint32_t*)dst)[k] = w;
  }
}

// ---- fp32 boolean decoder (see the file header). The window, its refill and the end-of-stream rule are BoolDec's.
#if defined(__CUDACC__) && !defined(VP8_EMU)
TK_FN float tf_fma_rz(float a, float b, float c) { return __fmaf_rz(a, b, c); }
TK_FN float tf_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
TK_FN float tf_as_float(uint32_t u) { return __uint_as_float(u); }
TK_FN uint32_t tf_as_uint(float f) { return __float_as_uint(f); }
// (x & m) | o with the two constants in registers: ONE LOP3 (with immediates the assembler needs two)
TK_FN uint32_t tf_and_or(uint32_t x, uint32_t m, uint32_t o) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(m), "r"(o)); return r; }
TK_FN uint32_t tf_mad(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
TK_FN uint32_t tf_umin(uint32_t a, uint32_t b) { return min(a, b); }
TK_FN uint32_t tf_shr_pair(uint32_t hi, uint32_t lo, uint32_t n) { return __funnelshift_rc(lo, hi, n); }   // low word of (hi:lo) >> min(n, 32)
#else
#include <math.h>
TK_FN float tf_fma_rz(float a, float b, float c) {   // a*b + c is exact in double here (at most 32 significant bits)
  const double d = (double)a * (double)b + (double)c;
  float f = (float)d;
  if ((double)f > d) f = nextafterf(f, 0.0f);
  return f;
}
TK_FN float tf_fma(float a, float b, float c) { return (float)((double)a * (double)b + (double)c); }
TK_FN float tf_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
TK_FN uint32_t tf_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
TK_FN uint32_t tf_and_or(uint32_t x, uint32_t m, uint32_t o) { return (x & m) | o; }
TK_FN uint32_t tf_mad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
TK_FN uint32_t tf_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
TK_FN uint32_t tf_shr_pair(uint32_t hi, uint32_t lo, uint32_t n) { return n >= 32 ? hi : n == 0 ? lo : ((hi << (32 - n)) | (lo >> n)); }
#endif

#define TF_S_2P23 0x1p-79f          // S * 2^23
#define TF_S_2P23M1 0x1.fffffcp-80f // S * (2^23 - 1)
#define TF_2P39 0x1p39f
#define TF_EXP46 0x56800000u        // exponent field of 2^46
#define TF_2PM141_BITS 0x00000100u  // 2^-141 (a denormal)

struct FpDec {
  const uint32_t* wp;    // word after `nxt`
  const uint32_t* wend;  // first word wholly past the stream: from there on zeros are shifted in
  const uint32_t* wbase; // stream bits moved into the window so far = 32 * (wp - wbase) - bias8
  uint32_t V, vlo, nxt;  // 64-bit left-aligned window V:vlo, nxt = the next raw (little-endian) word
  int nb;                // valid bits in V:vlo, plus 32 for every decode since the last fd_fill (a step adds k = 32 - shift)
  int nb_prev;           // nb before the most recent decode
  float Ra, Rp;          // (range - 1) * 2^39 and S * ((range - 1) + 2^23)
  int bias8;
  int64_t limit;         // 8*size - 8: a decode starting beyond this bit position reads past the end
  // The compressed bytes reach the reader through a 256-byte shared-memory ring per stream, filled 128 bytes at a time by
  // bulk asynchronous copies (cp.async.bulk, the 1-D form of TMA: one instruction per 128 bytes, completion on an mbarrier
  // per ring half) that the lane itself issues a whole chunk ahead of its reads: no per-word global load on the parser's
  // path. Chunk k (counted from the 128-byte aligned chunk of the stream's first word, `gbase`) lives in half k & 1 and
  // completes phase (k >> 1) & 1 of that half's barrier. (Host build: plain loads.)
  tk_saddr ring, bars;   // 256-byte aligned ring, two mbarriers
  uintptr_t gbase;       // global address of chunk 0
  uint32_t roff;         // ring offset (0..255) of the next word the reader takes
  int chunk;             // chunk that word lies in; chunk + 1 has been requested
};

#if defined(__CUDACC__) && !defined(VP8_EMU)
TK_FN void fd_ring_fetch(const FpDec& d, int chunk) {
  const tk_saddr dst = d.ring + (uint32_t)(chunk & 1) * 128u, bar = d.bars + (uint32_t)(chunk & 1) * 8u;
  const uintptr_t src = d.gbase + (uintptr_t)chunk * 128u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the half's earlier reads (generic proxy) before the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 128;" :: "r"(bar) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
               :: "r"(dst), "l"(src), "r"(bar) : "memory");
}
TK_FN int fd_ring_wait(const FpDec& d, int chunk) {   // 0 if the copy never lands (never expected: the caller fails the stream)
  const tk_saddr bar = d.bars + (uint32_t)(chunk & 1) * 8u;
  const uint32_t parity = (uint32_t)(chunk >> 1) & 1u;
  uint32_t done = 0;
  int spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done && ++spins < (1 << 22));
  return (int)done;
}
// After fd_init: the words from d.wp on come through the ring.
TK_FN void fd_ring_open(FpDec& d, tk_saddr ring, tk_saddr bars) {
  d.ring = ring; d.bars = bars; d.chunk = 0;
  d.gbase = (uintptr_t)d.wp & ~(uintptr_t)127;
  d.roff = (uint32_t)((uintptr_t)d.wp & 127u);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bars) : "memory");
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bars + 8u) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (d.wp < d.wend) {
    fd_ring_fetch(d, 0); fd_ring_fetch(d, 1);
    if (!fd_ring_wait(d, 0)) { d.limit = -0x4000000000000000ll; d.wend = d.wp; }
  }
}
// The reader steps into the next chunk (every 32nd word): it must have landed; the half behind it takes the chunk after.
TK_FN void fd_ring_cross(FpDec& d, const uint32_t* p) {
  d.chunk += 1;
  if (!fd_ring_wait(d, d.chunk)) { d.limit = -0x4000000000000000ll; d.wend = p; return; }   // given up: reads as "past the end"
  fd_ring_fetch(d, d.chunk + 1);
}
#endif

TK_FN void fd_init(FpDec& d, const uint8_t* start, uint32_t size) {
  BoolDec b;
  bd_init(b, start, size);
  d.ring = 0; d.bars = 0; d.gbase = 0; d.roff = 0; d.chunk = 0;
  d.wp = b.wp; d.wend = b.wend; d.wbase = b.wbase; d.V = b.V; d.vlo = b.vlo; d.nxt = b.nxt; d.nb = b.nbits; d.nb_prev = b.nbits;
  d.Ra = 254.0f * TF_2P39; d.Rp = tf_fma(254.0f * TF_2P39, tf_as_float(TF_2PM141_BITS), TF_S_2P23);
  d.bias8 = b.bias8; d.limit = b.limit;
}

// True when the reference's reader would have raised eof_ (bd_eof): the most recent decode STARTED with fewer than 8
// real bits left. Only valid right after a decode that followed an fd_fill by at most four steps; `since` = decodes
// since that fill (nb carries 32 extra per decode).
TK_FN int fd_eof(const FpDec& d, int since) {
  const int64_t loaded = 32 * (int64_t)(d.wp - d.wbase) - d.bias8;
  return (loaded - (d.nb_prev - 32 * (since - 1))) > d.limit;
}

// bd_fill_lookahead: tops the window up to > 32 valid bits; the word fetched here is not looked at before the next fill.
// Call with nb holding the true count (fd_settle). RING = 1: the words come through the shared-memory ring (device only).
template <int RING>
TK_FN void fd_fill(FpDec& d) {
  if (d.nb <= 32) {
    const uint32_t* p = d.wp;   // d.nxt came from p - 1
    const uint32_t w = (p - 1 < d.wend) ? VP8_BSWAP(d.nxt) : 0u;
#if defined(__CUDACC__) && !defined(VP8_EMU)
    if (RING) {
      if (p < d.wend) {   // a word at or past wend is never looked at (w above)
        if (TF_UNLIKELY((d.roff & 127u) == 0u && p != (const uint32_t*)d.gbase)) fd_ring_cross(d, p);
        if (p < d.wend) d.nxt = tk_lds_u32(d.ring + d.roff);
        d.roff = (d.roff + 4u) & 255u;
      }
    } else
#endif
    d.nxt = VP8_LDG(p < d.wend ? p : d.wend);   // straight from HBM through the read-only path; wend itself lies inside the arena's tail padding
    d.wp = p + 1;
    d.V |= vp8_shr_clamp(w, d.nb);
    d.vlo = vp8_shl_clamp(w, 32 - d.nb);
    d.nb += 32;
  }
}
TK_FN void fd_settle(FpDec& d, int decodes) { d.nb -= 32 * decodes; }   // after `decodes` steps: nb is the true count again

// Constants of the decode step that must sit in registers (see tf_and_or).
struct FpConst { uint32_t mant_mask, exp46; };

// fd_bit for the straight-line groups (tf_group_flat): the decode is computed, the reader only moves when `run`.
// nb stays the true count here (one three-input add).
TK_FN int fd_bit_guarded(FpDec& d, uint32_t prob_bits, const FpConst& k, bool run) {
  const float m = tf_fma_rz(d.Ra, tf_as_float(prob_bits), TF_S_2P23);
  const uint32_t s1 = tf_mad(tf_as_uint(m), 0x01000000u, 0x01000000u);
  const int bit = d.V >= s1;
  const float f0 = m - TF_S_2P23M1, f1 = d.Rp - m;
  const float f = bit ? f1 : f0;
  const uint32_t fb = tf_as_uint(f);
  const float fa = tf_as_float(tf_and_or(fb, k.mant_mask, k.exp46));
  const uint32_t kk = fb >> 23;
  const uint32_t vs = tf_umin(d.V, d.V - s1);
  if (run) {
    d.Ra = fa - TF_2P39;
    d.Rp = tf_fma(fa, tf_as_float(TF_2PM141_BITS), TF_S_2P23M1);
    d.V = tf_shr_pair(vs, d.vlo, kk);
    d.vlo = tf_shr_pair(d.vlo, 0u, kk);
    d.nb_prev = d.nb;
    d.nb = d.nb + (int)kk - 32;
  }
  return bit;
}

// One decode; pd = the probability byte as float BITS (a denormal). Needs >= 8 valid bits in the window.
TK_FN int fd_bit(FpDec& d, uint32_t prob_bits, const FpConst& k) {
  const float m = tf_fma_rz(d.Ra, tf_as_float(prob_bits), TF_S_2P23);
  const uint32_t s1 = tf_mad(tf_as_uint(m), 0x01000000u, 0x01000000u);
  const int bit = d.V >= s1;
  const float f0 = m - TF_S_2P23M1, f1 = d.Rp - m;
  const float f = bit ? f1 : f0;
  const uint32_t fb = tf_as_uint(f);
  const float fa = tf_as_float(tf_and_or(fb, k.mant_mask, k.exp46));   // 2^39 * the new range normalised into [128, 256)
  const uint32_t kk = fb >> 23;                                          // 32 - shift
  d.Ra = fa - TF_2P39;
  d.Rp = tf_fma(fa, tf_as_float(TF_2PM141_BITS), TF_S_2P23M1);
  const uint32_t vs = tf_umin(d.V, d.V - s1);
  d.V = tf_shr_pair(vs, d.vlo, kk);
  d.vlo = tf_shr_pair(d.vlo, 0u, kk);
  d.nb_prev = d.nb;
  d.nb += (int)kk;
  return bit;
}

// ---- one token as the reconstruction kernel reads it
//   31 sign   [29:25] block (0..15 luma, 16..23 chroma, 24 = Y2)   [24:13] magnitude   [9:6] position in parse (zigzag) order
#define TF_TOK_BLOCK(t) (((t) >> 25) & 31u)
#define TF_TOK_POS(t) (((t) >> 6) & 15u)
#define TF_TOK_MAG(t) (((t) >> TF_ADD_SHIFT) & 0xfffu)
#define TF_TOKENS_PER_MB VP8B_TOKENS_PER_MB
#define TF_MBTOK_FAILED 0x80000000u   // MbTok::count of the macroblock at which a partition ran out of data

// Lane phases as in vp8_tokens_lockstep.h.
#define TF_RUN 0
#define TF_BLOCK_END 1
#define TF_NEED_MB 2
#define TF_FINISHED 3

// ---- per-lane state (registers). BAND: the banded layout (see TF_TYPE_BYTES_B).
template <int BAND>
struct TfLaneT {
  FpDec d;
  tk_saddr a;             // address of the pending probability: row of (type, position) + state
  tk_saddr rowend;        // first address of position 16 of the current block type
  tk_saddr a_end;         // straight-line groups: where the block ended (L.a wanders on afterwards)
  uint32_t tok_end;       // ... and the token count at that point
  uint32_t pos6, pos_end; // BAND: position of the pending probability << 6 (the 16-row layout reads it off the address); at a_end
  int32_t dm64;           // BAND: TfTables::delta of that position
  uint32_t pb;            // the pending probability: the byte, i.e. the bits of a denormal float
  uint32_t tag, tag_s;    // block index << 25, and the same with the sign bit set
  int eofs;               // fd_eof() of the most recent decode, taken where a macroblock ends
  uint32_t e0, e1;        // the transition entries of the pending decode's two outcomes
  uint32_t sink;          // see tf_decode (never meaningful)
  uint32_t acc;           // sum of the entries taken since the last emit / block start: magnitude in TF_ADD_MASK
  uint32_t tokoff;        // next token of this partition, counted from the image's first token
  uint32_t mbtok0;        // first token of the current macroblock
  uint32_t cx;            // non-zero contexts: top in bits 0-8 (0-3 luma, 4-5 U, 6-7 V, 8 Y2), left in bits 16-24
  uint32_t acc_lo, acc_hi;// 2-bit nz codes of the macroblock's blocks shifted in, in parse order
  uint32_t m, m_next;     // context bits of the current / the next block (seqmask)
  uint32_t lut;           // nz -> 2-bit code of the current block: 2-bit fields indexed by min(nz, 4)
  int seq;                // 0 = Y2, 1..16 luma, 17..24 chroma
  // macroblock
  tk_saddr yrow, yend;    // luma rows of this macroblock: first one parsed, end
  uint32_t ylut;
  uint32_t w, w_next;     // MbInfo word 3 of this / the partition's next macroblock
  int mx, my;
  int done_mbs;
  int pend;
  int waiting;            // P > 1: the partition owning the row above has not got far enough yet
  int alive;              // 0 once parked
  int status;
};
typedef TfLaneT<0> TfLane;
#define TF_LUT_FROM0 0x3a4u   // nz 0 -> 0, 1 -> 1 (a lone DC level: re-examined after dequantisation, recon_macroblock), 2,3 -> 2, >= 4 -> 3
#define TF_LUT_FROM1 0x3a0u   // luma blocks of i16 macroblocks start at coefficient 1: nz = 1 means empty

// One macroblock's slice of its partition's token stream.
struct MbTok { uint32_t first, count; };

// Per-lane constants.
struct TfCtx {
  tk_saddr img_s;         // this image's rows
  tk_saddr tab_s;         // TfTables
  uint16_t* topctx;       // (P + 1) x ctx_stride ring
  volatile int* progress; // P counters
  uint32_t* mbinfo;       // this image's MbInfo
  MbTok* mbtok;           // this image's MbTok
  uint32_t* tokens;       // this image's token area (TF_TOKENS_PER_MB per macroblock)
  int mb_w, rows, P, part, use_skip, ctx_stride;
  FpConst k;
};

template <int BAND>
TK_FN void tf_lane_reset(TfLaneT<BAND>& L, const TfCtx& c) {
  L.pos6 = 0; L.pos_end = 0; L.dm64 = 0;
  L.a = c.img_s; L.rowend = 0; L.a_end = 0; L.tok_end = 0; L.pb = 0; L.e0 = 0; L.e1 = 0; L.acc = 0; L.sink = 0; L.tag = 0; L.tag_s = 0; L.cx = 0;
  L.acc_lo = 0; L.acc_hi = 0; L.m = 0; L.m_next = 0; L.lut = 0; L.seq = 0;
  L.yrow = 0; L.yend = 0; L.ylut = 0;
  L.mx = 0; L.my = c.part; L.done_mbs = 0; L.waiting = 1; L.alive = 1; L.status = VP8B_OK;
  L.pend = TF_NEED_MB;
  L.w = 0; L.w_next = 0;
  // partition p owns rows p, p + P, ...: its tokens follow those of the partitions before it
  uint32_t rows_before = 0;
  for (int q = 0; q < c.part; ++q) rows_before += (uint32_t)((c.rows - q + c.P - 1) / c.P);
  L.tokoff = rows_before * (uint32_t)c.mb_w * TF_TOKENS_PER_MB;
  L.mbtok0 = L.tokoff;
}

template <int BAND>
TK_FN void tf_lane_init(TfLaneT<BAND>& L, const TfCtx& c, const uint8_t* frame, const FrameHdr* h) {
  fd_init(L.d, frame + h->part_off[c.part], h->part_size[c.part]);
  tf_lane_reset(L, c);
  L.eofs = fd_eof(L.d, 1);
  L.w_next = (c.part < c.rows) ? VP8_LDG(c.mbinfo + 4 * ((size_t)c.part * c.mb_w) + 3) : 0;
}

// Loads the pending decode (probability, both transition entries) of state s at address L.a.
template <int BAND>
TK_FN void tf_prime(TfLaneT<BAND>& L, const TfCtx& c, uint32_t s) {
  L.pb = tk_lds_u8(L.a);
  tk_lds_v2(c.tab_s + s * 8u, L.e0, L.e1);
  if (BAND) L.dm64 = tk_lds_s16(c.tab_s + TFT_DELTA + (L.pos6 >> 5));
}

TK_FN uint32_t tf_popc(uint32_t x) {
#if defined(__CUDACC__) && !defined(VP8_EMU)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}

// Sets up block L.seq >= 1 (contexts in L.cx are final for it). Luma block k is block k - 1, chroma blocks follow.
// (Preparing the next block's start state while the current one is parsed -- everything of its context but the one bit
// this block sets -- takes the POPC and a load off the block end's critical path, but costs 16 registers and four
// instructions per block: measured 267 -> 276 ms per 4096 full-HD images, profiles/r02l, and dropped.)
template <int BAND>
TK_FN void tf_block_setup(TfLaneT<BAND>& L, const TfCtx& c) {
  const int chroma = L.seq >= 17;
  const tk_saddr crow = c.img_s + 2 * TF_TYPE_BYTES_B(BAND);
  L.m = L.m_next;   // fetched while the previous block was parsed
  L.m_next = tk_lds_u32(c.tab_s + TFT_SEQMASK + 4u * (uint32_t)L.seq + 4u);
  L.rowend = chroma ? crow + TF_TYPE_BYTES_B(BAND) : L.yend;
  if (BAND) L.pos6 = chroma ? 0u : (L.ylut == TF_LUT_FROM1 ? 64u : 0u);   // luma of an i16 macroblock starts at coefficient 1
  L.lut = chroma ? TF_LUT_FROM0 : L.ylut;
  L.tag = ((uint32_t)L.seq - 1u) << 25; L.tag_s = L.tag | 0x80000000u;
  L.acc = 0;
  const uint32_t s = tf_popc(L.cx & L.m) * 11u;
  L.a = (chroma ? crow : L.yrow) + s;
  tf_prime(L, c, s);
}

// The Y2 block of an i16 macroblock (seq 0): type 1, block 24.
template <int BAND>
TK_FN void tf_y2_setup(TfLaneT<BAND>& L, const TfCtx& c) {
  L.m = (1u << 8) | (1u << 24);
  L.m_next = (1u << 0) | (1u << 16);   // seq 1
  const tk_saddr row = c.img_s + 1 * TF_TYPE_BYTES_B(BAND);
  L.rowend = row + TF_TYPE_BYTES_B(BAND);
  if (BAND) L.pos6 = 0;
  L.lut = TF_LUT_FROM0;
  L.tag = 24u << 25; L.tag_s = L.tag | 0x80000000u;
  L.acc = 0;
  const uint32_t s = tf_popc(L.cx & L.m) * 11u;
  L.a = row + s;
  tf_prime(L, c, s);
}

// Stores a finished (or skipped) macroblock's results and steps to the partition's next macroblock.
template <int MULTI, int BAND>
TK_FN void tf_mb_store(TfLaneT<BAND>& L, const TfCtx& c, uint32_t nzy, uint32_t w3) {
  const int P = MULTI ? c.P : 1, mb_w = c.mb_w;
  const size_t idx = (size_t)L.my * mb_w + L.mx;
  c.mbinfo[4 * idx + 2] = nzy;
  c.mbinfo[4 * idx + 3] = w3;
  MbTok mt; mt.first = L.mbtok0; mt.count = L.tokoff - L.mbtok0;
  c.mbtok[idx] = mt;
  L.mbtok0 = L.tokoff;
  const int ring_row = MULTI ? L.my % (P + 1) : (L.my & 1);
  c.topctx[(size_t)ring_row * c.ctx_stride + L.mx] = (uint16_t)(L.cx & 0x1ffu);
  L.done_mbs++;
  if (++L.mx == mb_w) { L.mx = 0; L.my += P; }
  if (L.eofs) {
    // Ran past the end of the partition: the image is lost (vp8_dec.c:651-659); release whoever waits on us. The row goes on
    // record (tf_failed_row): with a damaged ALPH chunk as well, the reference reports the failure its row loop meets first.
    c.mbtok[idx].count = TF_MBTOK_FAILED;
    L.status = VP8B_NOT_ENOUGH_DATA;
    if (MULTI) { TK_FENCE(); c.progress[c.part] = 0x7fffffff; }
  } else if (MULTI) { TK_FENCE(); c.progress[c.part] = L.done_mbs; }
}

// The macroblock row at which a frame's tokens ran out of data (FrameHdr::fail_row), found afterwards from the marks
// tf_mb_store leaves (TF_MBTOK_FAILED): k_reconstruct does this block-wide on the device, the emulation with this loop.
TK_FN int tf_find_failed_row(const MbTok* mbtok, int mb_w, int rows) {
  for (int i = 0; i < mb_w * rows; ++i) if (mbtok[i].count == TF_MBTOK_FAILED) return i / mb_w;
  return VP8B_FAIL_NONE;
}

// Leaves the lane either with a block set up (returns 1), waiting for the row above (returns 0, L.waiting = 1) or
// finished (returns 0, L.waiting = 0, L.my >= rows or L.status != OK). Skipped macroblocks are consumed here.
template <int MULTI, int BAND>
TK_FN int tf_mb_next(TfLaneT<BAND>& L, const TfCtx& c) {
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
      // luma: type 3 from coefficient 0 (i4x4) or type 0 from coefficient 1 (after the Y2 block)
      const tk_saddr ybase = c.img_s + (is_i4 ? 3u : 0u) * TF_TYPE_BYTES_B(BAND);
      L.yend = ybase + TF_TYPE_BYTES_B(BAND);
      L.yrow = ybase + (is_i4 ? 0u : (uint32_t)TF_ROW_BYTES);
      L.ylut = is_i4 ? TF_LUT_FROM0 : TF_LUT_FROM1;
      if (is_i4) { L.seq = 1; L.m_next = (1u << 0) | (1u << 16); tf_block_setup(L, c); } else { L.seq = 0; tf_y2_setup(L, c); }
      return 1;
    }
    L.cx &= is_i4 ? 0x01000100u : 0u;
    tf_mb_store<MULTI>(L, c, 0u, L.w & 0xffff0000u);
  }
}

// The macroblock's last block has ended: store its results, move on.
template <int MULTI, int BAND>
TK_FN void tf_mb_finish(TfLaneT<BAND>& L, const TfCtx& c, int since) {
  L.eofs = fd_eof(L.d, since);
  const uint32_t nzy = (L.acc_hi << 16) | (L.acc_lo >> 16);
  const uint32_t uv = L.acc_lo & 0xffffu;                      // U codes in bits 15-8, V in 7-0
  const uint32_t nzuv = (uv >> 8) | ((uv & 0xffu) << 8);       // reference order: U bits 0-7, V bits 8-15
  uint32_t w = L.w;
  if ((L.acc_hi >> 16) & 3u) w |= MBW_HAS_Y2;                  // the Y2 block's code, shifted in first
  tf_mb_store<MULTI>(L, c, nzy, (w & 0xffff0000u) | nzuv);
}

// Parks a lane that has nothing (more) to do: it keeps decoding in the dead state (probability 0 for ever, never
// emits, never ends a block) and its reader only shifts in zeros.
template <int BAND>
TK_FN void tf_lane_park(TfLaneT<BAND>& L, const TfCtx& c) {
  L.pend = TF_FINISHED; L.alive = 0; L.waiting = 0;
  L.a = c.img_s + TF_DEAD; L.rowend = ~(tk_saddr)0;
  L.pos6 = 0;
  tf_prime(L, c, TF_DEAD);
  L.pb = 0; L.dm64 = 0;
}

// A lane without a stream: finished from the start (`any` = some valid address for its reader).
template <int BAND>
TK_FN void tf_lane_idle(TfLaneT<BAND>& L, const TfCtx& c, const uint8_t* any) {
  fd_init(L.d, any, 0);
  tf_lane_reset(L, c);
  L.eofs = 0;
  tf_lane_park(L, c);
}

// One boolean decode and the transition it selects; returns the transition entry. The caller has topped the window up
// (fd_fill) within the last three decodes.
//
// What the next decode needs (its probability, the entries of its two outcomes) is loaded AFTER the bit is known, at
// an address picked by the bit: one shared-memory latency per step on the dependent chain (23 cycles, about half of
// it). The alternative -- loading for both outcomes ahead of the bit and selecting -- takes the latency off the chain
// but costs ten more integer-pipe instructions per step, and one warp alone on a sub-partition gets an integer-pipe
// slot only every other cycle: measured 102 cycles per step against this form's (profiles/r02e, r02f).
template <int BAND>
TK_FN uint32_t tf_decode(TfLaneT<BAND>& L, const TfCtx& c) {
  tk_saddr a0 = L.a + TF_E_DIST(L.e0), a1 = L.a + TF_E_DIST(L.e1);
  if (BAND) {   // an advancing outcome lands on the next position's BAND row, not 64 bytes further
    const tk_saddr dm = (tk_saddr)(long long)L.dm64;   // (signed: some steps go back)
    a0 += (L.e0 & TF_ADV) ? dm : (tk_saddr)0;
    a1 += (L.e1 & TF_ADV) ? dm : (tk_saddr)0;
  }
  const tk_saddr t0 = c.tab_s + TF_E_TAB(L.e0), t1 = c.tab_s + TF_E_TAB(L.e1);
  // ---- boolean decode
  const int bit = fd_bit(L.d, L.pb, c.k);
  // ---- transition
  const uint32_t e = bit ? L.e1 : L.e0;
  const uint32_t pos_emit = BAND ? L.pos6 : ((uint32_t)L.a & 0x3c0u);
  L.a = bit ? a1 : a0;
  L.pb = tk_lds_u8(L.a);
  tk_lds_v2(bit ? t1 : t0, L.e0, L.e1);
  if (BAND) {
    L.pos6 += (e & TF_ADV) << 4;
    L.dm64 = tk_lds_s16(c.tab_s + TFT_DELTA + (L.pos6 >> 5));   // for the decode after this one
  }
  L.acc += e;
  if (e & TF_EMIT) {   // the sign has just been decoded: one token
    const uint32_t t = (L.acc & TF_ADD_MASK) | (bit ? L.tag_s : L.tag);
    c.tokens[L.tokoff] = t | pos_emit;
    L.tokoff += 1;
    L.acc = 0;
  }
  return e;
}

// ParseResiduals' bookkeeping at the end of a block (vp8_dec.c:517-609) with GetCoeffs' return value nz, then the next
// block. Returns 1 when the macroblock's last block has ended instead (tf_mb_finish + tf_mb_next are due).
// nz = position of the last decoded coefficient + 1 = the position the walk stands at: after an emit the address has
// already moved to the next position; at an end-of-block decision it still stands at the position that was asked.
template <int BAND>
TK_FN int tf_block_end(TfLaneT<BAND>& L, const TfCtx& c) {
  const uint32_t nz2 = BAND ? (L.pos6 >> 5) : 32u - (uint32_t)((L.rowend - (L.a & ~(tk_saddr)63)) >> 5);   // 2 * nz
  const uint32_t code = (L.lut >> (nz2 < 8u ? nz2 : 8u)) & 3u;       // non-zero exactly when the block counts as non-empty
  L.acc_hi = (L.acc_hi << 2) | (L.acc_lo >> 30);
  L.acc_lo = (L.acc_lo << 2) | code;
  L.cx = (L.cx & ~L.m) | (code ? L.m : 0u);
  L.seq++;
  if (TF_UNLIKELY(L.seq == 25)) return 1;
  tf_block_setup(L, c);
  return 0;
}

// ---- every lane executes every step; a lane whose block has ended does its bookkeeping on the spot while the others
// wait. Returns 0 once the lane has finished (parked).
template <int MULTI, int BAND>
TK_FN int tf_step_inline(TfLaneT<BAND>& L, const TfCtx& c, int since /* decodes since the last fd_fill, this one included */) {
  if (MULTI) {
    if (L.waiting) {
      if (!tf_mb_next<MULTI>(L, c)) {
        if (!L.waiting) tf_lane_park(L, c);
        L.d.nb += 32;   // a step without a decode: fd_settle counts 32 per step
        return L.alive;
      }
    }
  }
  const uint32_t e = tf_decode(L, c);
  if (TF_UNLIKELY((e & TF_EOB) || (BAND ? L.pos6 >= 1024u : L.a >= L.rowend))) {
    if (tf_block_end(L, c)) {
      tf_mb_finish<MULTI>(L, c, since);
      if (!tf_mb_next<MULTI>(L, c)) {
        if (!(MULTI && L.waiting)) tf_lane_park(L, c);
        return L.alive;
      }
    }
  }
  return 1;
}

// One group: top the window up, four decodes.
template <int MULTI, int RING, int BAND>
TK_FN void tf_group_inline(TfLaneT<BAND>& L, const TfCtx& c) {
  fd_fill<RING>(L.d);
  tf_step_inline<MULTI>(L, c, 1); tf_step_inline<MULTI>(L, c, 2); tf_step_inline<MULTI>(L, c, 3); tf_step_inline<MULTI>(L, c, 4);
  fd_settle(L.d, 4);
}

// ---- second way of running the lanes: groups of four straight-line steps, nothing branches between them.
// One warp alone on a sub-partition pays ~18 cycles for every branch it executes (resolve + reconverge) and cannot
// overlap anything across it, so the end-of-block test per step cost more than the decode it guards (profiles/r02f).
// Here a lane whose block ends inside a group is switched to the dead state on the spot (probability 0, entries that
// neither emit nor end a block) and its READER is frozen by predication (fd_bit_guarded: `run`); it sits out the rest
// of the group, and all lanes that ended do their bookkeeping together at the group's one event point. A lane that has
// to wait for the row above (several partitions) or has finished is in the dead state from the start of the group.
#define TF_DEAD_ENTRY TF_E(TF_DEAD, TF_DEAD, 0, 0, 0)

template <int BAND>
TK_FN void tf_go_dead(TfLaneT<BAND>& L, const TfCtx& c) {   // the decode state of a lane that is not running
  L.a = c.img_s + TF_DEAD; L.rowend = ~(tk_saddr)0;
  L.pos6 = 0; L.dm64 = 0;
  L.pb = 0; L.e0 = TF_DEAD_ENTRY; L.e1 = TF_DEAD_ENTRY;
}

template <int BAND>
TK_FN bool tf_step_flat(TfLaneT<BAND>& L, const TfCtx& c, const bool run) {
  tk_saddr a0 = L.a + TF_E_DIST(L.e0), a1 = L.a + TF_E_DIST(L.e1);
  if (BAND) {
    const tk_saddr dm = (tk_saddr)(long long)L.dm64;   // (signed: some steps go back)
    a0 += (L.e0 & TF_ADV) ? dm : (tk_saddr)0;
    a1 += (L.e1 & TF_ADV) ? dm : (tk_saddr)0;
  }
  const tk_saddr t0 = c.tab_s + TF_E_TAB(L.e0), t1 = c.tab_s + TF_E_TAB(L.e1);
  const int bit = fd_bit_guarded(L.d, L.pb, c.k, run);
  const uint32_t e = bit ? L.e1 : L.e0;
  const uint32_t pos_emit = BAND ? (L.pos6 & 0x3c0u) : ((uint32_t)L.a & 0x3c0u);
  const tk_saddr an = bit ? a1 : a0;
  L.a = an;
  L.pb = tk_lds_u8(an);
  tk_lds_v2(bit ? t1 : t0, L.e0, L.e1);
  if (BAND) {   // (a lane that is not running wanders: the mask keeps its position inside the delta table)
    L.pos6 = (L.pos6 + ((e & TF_ADV) << 4)) & 0x7c0u;
    L.dm64 = tk_lds_s16(c.tab_s + TFT_DELTA + ((L.pos6 & 0x3c0u) >> 5));
  }
  L.acc += e;
  if (e & TF_EMIT) {   // the sign has just been decoded: one token. A lane that is not running writes its garbage into the
                       // slots after its last real token, which the next real tokens overwrite (tf_events puts tokoff back)
    const uint32_t t = (L.acc & TF_ADD_MASK) | (bit ? L.tag_s : L.tag);
    c.tokens[L.tokoff] = t | pos_emit;
    L.tokoff += 1;
    L.acc = 0;
  }
  const bool end = ((e & TF_EOB) != 0u) || (BAND ? L.pos6 >= 1024u : an >= L.rowend);
  if (run && end) { L.a_end = an; L.tok_end = L.tokoff; if (BAND) L.pos_end = L.pos6; }   // the block has ended here: the position gives nz
  return run && !end;
}

// The group's event point: lanes whose block has ended do ParseResiduals' bookkeeping and start their next block or
// macroblock; lanes that wait for the row above try again.
template <int MULTI, int BAND>
TK_FN void tf_events(TfLaneT<BAND>& L, const TfCtx& c, int ended) {
  int need_mb = (L.pend == TF_NEED_MB);
  L.pend = TF_RUN;
  if (ended) {
    L.a = L.a_end; L.tokoff = L.tok_end;
    if (BAND) L.pos6 = L.pos_end;
    if (tf_block_end(L, c)) { tf_mb_finish<MULTI>(L, c, 1); need_mb = 1; }
  }
  if (TF_UNLIKELY(need_mb)) {
    if (!tf_mb_next<MULTI>(L, c)) {
      if (MULTI && L.waiting) { L.pend = TF_NEED_MB; tf_go_dead(L, c); } else tf_lane_park(L, c);
    }
  }
}

// One group: four decodes, then the event point for the lanes that need it. A running lane whose block ends inside the
// group keeps stepping on garbage with its reader frozen (fd_bit_guarded) until the event point puts it right.
template <int MULTI, int RING, int BAND>
TK_FN void tf_group_flat(TfLaneT<BAND>& L, const TfCtx& c) {
  const bool run0 = L.pend == TF_RUN;
  if (run0) fd_fill<RING>(L.d);
  bool run = run0;
  run = tf_step_flat(L, c, run); run = tf_step_flat(L, c, run); run = tf_step_flat(L, c, run); run = tf_step_flat(L, c, run);
  if ((run0 && !run) || L.pend == TF_NEED_MB) tf_events<MULTI>(L, c, run0 && !run);
}

#endif  // LIBWEBP_B200_VP8_TOKENS_FP_H_
