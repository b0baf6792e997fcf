// vp8_modes_lockstep.h -- intra-mode parse (K1), second mapping: images as lockstep lanes of a warp.
//
// k_parse_modes gives every image a warp of its own and runs VP8ParseIntraModeRow (src/dec/tree_dec.c:290-367) as
// straight-line code in ONE lane: ~32 instructions per boolean decode, seven such warps per SM sub-partition, issue-bound
// (85 % issue-active, 30 ms per 4096 full-HD images). Here the images of a warp are its lanes (7 for that batch) and share
// one instruction stream, as the token parser does (vp8_tokens_fp.h): every step is one boolean decode per lane with the
// fp32 decoder, and the syntax is ONE table of tree nodes -- segment id, skip flag, block size, the 16x16 mode tree, the
// ten-leaf sub-block mode tree (kYModesIntra4, tree_dec.c:28-36) and the chroma mode tree laid end to end. An entry either
// names the next node and where its probability sits relative to the tree's row (the sub-block tree's row is the context
// row kBModesProba[top][left], the others share a per-image row of fixed probabilities), or it is a leaf: the lane then
// does that leaf's bookkeeping, everybody else sitting the branch out. Contexts rotate instead of being indexed: the four
// modes above and to the left live in one word each, the next one always in the low byte.
//
// Replaces VP8ParseIntraModeRow / ParseIntraMode for a batch; the frame header (VP8GetHeaders, VP8ParseProba) stays with
// parse_frame_header, run by the same lanes before the loop. Dual build: nvcc for the product, g++ -DVP8_EMU for tests/emu.
#ifndef LIBWEBP_B200_VP8_MODES_LOCKSTEP_H_
#define LIBWEBP_B200_VP8_MODES_LOCKSTEP_H_

#include "vp8_tokens_fp.h"

// ---- nodes. A node's two entries sit at 8 * node (+ 4 for a decoded 1).
#define ML_B0 0      // sub-block mode tree: nodes 0-8, probability = context row[node]
#define ML_Y0 9      // 16x16 mode: 156, then 128 (after a 1) or 163 (after a 0)
#define ML_Y1 10
#define ML_Y2 11
#define ML_U0 12     // chroma mode: 142, 114, 183
#define ML_U1 13
#define ML_U2 14
#define ML_S0 15     // segment id: the image's three probabilities
#define ML_S1 16
#define ML_S2 17
#define ML_SKIP 18
#define ML_I16 19    // 145: 1 = one 16x16 mode, 0 = sixteen sub-block modes
#define ML_DEAD 20   // a lane that has finished or failed: probability 0, loops to itself
#define ML_NODES 21
#define ML_TAB_BYTES (ML_NODES * 8)

// ---- the per-image row of fixed probabilities (16 bytes): offsets of the nodes above in it
#define ML_ROW_BYTES 16
#define ML_OFF_S0 0
#define ML_OFF_SKIP 3
#define ML_OFF_I16 4
#define ML_OFF_Y0 5
#define ML_OFF_U0 8
#define ML_OFF_DEAD 11

// ---- entry: [2:0] leaf kind (0 = inner node)   [7:3] inner: 8 * next node >> 3, leaf: the value   [15:8] inner: offset of
// the next node's probability in the tree's row
#define ML_INNER(next, off) (((uint32_t)(off) << 8) | ((uint32_t)(next) << 3))
#define ML_LEAF(kind, value) (((uint32_t)(value) << 3) | (uint32_t)(kind))
#define ML_K_B 1
#define ML_K_Y 2
#define ML_K_UV 3
#define ML_K_SEG 4
#define ML_K_SKIP 5
#define ML_K_I16 6

TK_FN uint32_t ml_entry(int node, int b) {
  switch (node) {
    // ParseIntraMode's sub-block tree (tree_dec.c:28-36, 330-352), the decisions of parse_bmode (vp8_parse_core.h)
    case 0: return b ? ML_INNER(1, 1) : ML_LEAF(ML_K_B, M_DC);
    case 1: return b ? ML_INNER(2, 2) : ML_LEAF(ML_K_B, M_TM);
    case 2: return b ? ML_INNER(3, 3) : ML_LEAF(ML_K_B, M_VE);
    case 3: return b ? ML_INNER(6, 6) : ML_INNER(4, 4);
    case 4: return b ? ML_INNER(5, 5) : ML_LEAF(ML_K_B, M_HE);
    case 5: return b ? ML_LEAF(ML_K_B, M_VR) : ML_LEAF(ML_K_B, M_RD);
    case 6: return b ? ML_INNER(7, 7) : ML_LEAF(ML_K_B, M_LD);
    case 7: return b ? ML_INNER(8, 8) : ML_LEAF(ML_K_B, M_VL);
    case 8: return b ? ML_LEAF(ML_K_B, M_HU) : ML_LEAF(ML_K_B, M_HD);
    // 16x16 mode (tree_dec.c:319-323)
    case ML_Y0: return b ? ML_INNER(ML_Y1, ML_OFF_Y0 + 1) : ML_INNER(ML_Y2, ML_OFF_Y0 + 2);
    case ML_Y1: return ML_LEAF(ML_K_Y, b ? M_TM : M_HE);
    case ML_Y2: return ML_LEAF(ML_K_Y, b ? M_VE : M_DC);
    // chroma mode (tree_dec.c:355-358)
    case ML_U0: return b ? ML_INNER(ML_U1, ML_OFF_U0 + 1) : ML_LEAF(ML_K_UV, M_DC);
    case ML_U1: return b ? ML_INNER(ML_U2, ML_OFF_U0 + 2) : ML_LEAF(ML_K_UV, M_VE);
    case ML_U2: return ML_LEAF(ML_K_UV, b ? M_TM : M_HE);
    // segment id (tree_dec.c:303-309)
    case ML_S0: return b ? ML_INNER(ML_S2, ML_OFF_S0 + 2) : ML_INNER(ML_S1, ML_OFF_S0 + 1);
    case ML_S1: return ML_LEAF(ML_K_SEG, b);
    case ML_S2: return ML_LEAF(ML_K_SEG, 2 + b);
    case ML_SKIP: return ML_LEAF(ML_K_SKIP, b);
    case ML_I16: return ML_LEAF(ML_K_I16, b);
    default: return ML_INNER(ML_DEAD, ML_OFF_DEAD);
  }
}

TK_FN void ml_table_fill(uint32_t* tab /* ML_NODES x 2 */, int tid, int nthreads) {
  for (int k = tid; k < 2 * ML_NODES; k += nthreads) tab[k] = ml_entry(k >> 1, k & 1);
}

// One image's row of fixed probabilities from its parsed header (one thread).
TK_FN void ml_row_fill(uint8_t* row, const FrameHdr* h) {
  row[0] = h->seg_prob[0]; row[1] = h->seg_prob[1]; row[2] = h->seg_prob[2];
  row[ML_OFF_SKIP] = h->skip_p; row[ML_OFF_I16] = 145;
  row[ML_OFF_Y0] = 156; row[ML_OFF_Y0 + 1] = 128; row[ML_OFF_Y0 + 2] = 163;
  row[ML_OFF_U0] = 142; row[ML_OFF_U0 + 1] = 114; row[ML_OFF_U0 + 2] = 183;
  for (int k = ML_OFF_DEAD; k < ML_ROW_BYTES; ++k) row[k] = 0;
}

struct MlCtx {
  tk_saddr tab_s;      // the node table
  tk_saddr bprob_s;    // kVp8BModeProba [10 above][10 left][9]
  tk_saddr row_s;      // this image's fixed probabilities
  tk_saddr top_s;      // this image's sub-block modes of the macroblock row above, one word per macroblock column
  uint32_t* out;       // this image's MbInfo (4 words per macroblock)
  int mb_w, mb_h;      // mb_h = the macroblock rows that get parsed (FrameHdr::rows)
  uint32_t first_node8, first_off;     // where a macroblock starts: segment id, skip flag or block size, as the header says
  uint32_t skip_node8, skip_off;       // what follows the segment id: the skip flag or the block size
  FpConst k;
};

struct MlLane {
  FpDec d;
  tk_saddr row;        // row the current tree's offsets are relative to
  uint32_t pb;         // the pending probability (the byte = the bits of a denormal float)
  uint32_t node8;      // 8 * the pending node
  uint32_t t, left;    // modes above / to the left of the sub-blocks still to come, the next one in the low byte
  uint32_t m0, m1, w;  // the macroblock's MbInfo words as they fill up
  int n;               // sub-blocks of the current macroblock parsed so far
  int mx, my;
  uint32_t mb;         // index of the current macroblock
  int alive;
  int status, fail_row;
};

TK_FN void ml_goto(MlLane& L, const MlCtx& c, uint32_t node8, tk_saddr row, uint32_t off) {
  L.node8 = node8; L.row = row; L.pb = tk_lds_u8(row + off);
}

TK_FN void ml_park(MlLane& L, const MlCtx& c) {
  L.alive = 0;
  ml_goto(L, c, 8u * ML_DEAD, c.row_s, ML_OFF_DEAD);
}

// Takes over from the header parse: the window and the range of `br` as the fp decoder keeps them (fd_init).
TK_FN void ml_start(MlLane& L, const MlCtx& c, const BoolDec& br) {
  FpDec& d = L.d;
  d.ring = 0; d.bars = 0; d.gbase = 0; d.roff = 0; d.chunk = 0;
  d.wp = br.wp; d.wend = br.wend; d.wbase = br.wbase; d.V = br.V; d.vlo = br.vlo; d.nxt = br.nxt; d.nb = br.nbits; d.nb_prev = br.nbits;
  d.Ra = (float)(br.R24 >> 24) * TF_2P39;
  d.Rp = tf_fma(d.Ra, tf_as_float(TF_2PM141_BITS), TF_S_2P23);
  d.bias8 = br.bias8; d.limit = br.limit;
  L.t = 0; L.left = 0; L.m0 = 0; L.m1 = 0; L.w = 0; L.n = 0; L.mx = 0; L.my = 0; L.mb = 0;
  L.alive = 1; L.status = VP8B_OK; L.fail_row = VP8B_FAIL_NONE;
  for (int mx = 0; mx < c.mb_w; ++mx) tk_stsv_u32(c.top_s + 4u * (uint32_t)mx, 0u);   // M_DC
  ml_goto(L, c, c.first_node8, c.row_s, c.first_off);
  if (c.mb_h <= 0) ml_park(L, c);
}

// The sub-block tree of the next sub-block: the row of (mode above, mode to the left).
TK_FN void ml_goto_subblock(MlLane& L, const MlCtx& c) {
  const tk_saddr row = c.bprob_s + (L.t & 255u) * 90u + (L.left & 255u) * 9u;
  ml_goto(L, c, 8u * ML_B0, row, 0);
}

// One step: one boolean decode, then the node table. `since` = decodes since the last fd_fill, this one included.
TK_FN void ml_step(MlLane& L, const MlCtx& c, int since) {
  const int bit = fd_bit(L.d, L.pb, c.k);
  const uint32_t e = tk_lds_u32(c.tab_s + L.node8 + 4u * (uint32_t)bit);
  const uint32_t kind = e & 7u;
  if (kind == 0) {
    L.node8 = e & 0xf8u;
    L.pb = tk_lds_u8(L.row + (e >> 8));
    return;
  }
  const uint32_t v = e >> 3;
  if (kind == ML_K_B) {
    // one more sub-block mode: into the MbInfo nibbles (the first one ends up lowest), over the mode above it, and -- at
    // the end of a sub-block row -- over the left mode of that row; the contexts of the next sub-block turn up in the low bytes
    L.m0 = (L.m0 >> 4) | (L.m1 << 28);
    L.m1 = (L.m1 >> 4) | (v << 28);
    L.t = (L.t >> 8) | (v << 24);
    L.n += 1;
    if ((L.n & 3) == 0) L.left = (L.left >> 8) | (v << 24);
    else L.left = (L.left & ~255u) | v;
    if (L.n < 16) { ml_goto_subblock(L, c); return; }
    // (after four rows `left` has the four right-most modes in row order again, the low byte being scratch until then:
    //  it was overwritten by every mode of the row, and the row's last rotation moved it out)
    tk_stsv_u32(c.top_s + 4u * (uint32_t)L.mx, L.t);
    ml_goto(L, c, 8u * ML_U0, c.row_s, ML_OFF_U0);
    return;
  }
  if (kind == ML_K_I16) {
    if (v) { ml_goto(L, c, 8u * ML_Y0, c.row_s, ML_OFF_Y0); return; }
    L.w |= MBW_I4X4;
    L.n = 0;
    L.t = tk_ldsv_u32(c.top_s + 4u * (uint32_t)L.mx);
    ml_goto_subblock(L, c);
    return;
  }
  if (kind == ML_K_Y) {
    L.m0 = v; L.m1 = 0;
    L.t = v * 0x01010101u; L.left = L.t;
    tk_stsv_u32(c.top_s + 4u * (uint32_t)L.mx, L.t);
    ml_goto(L, c, 8u * ML_U0, c.row_s, ML_OFF_U0);
    return;
  }
  if (kind == ML_K_SEG) { L.w |= v << MBW_SEG_SHIFT; ml_goto(L, c, c.skip_node8, c.row_s, c.skip_off); return; }
  if (kind == ML_K_SKIP) { if (v) L.w |= MBW_SKIP; ml_goto(L, c, 8u * ML_I16, c.row_s, ML_OFF_I16); return; }
  // chroma mode: the macroblock is complete
  L.w |= v << MBW_UVMODE_SHIFT;
  uint4 o4; o4.x = L.m0; o4.y = L.m1; o4.z = 0; o4.w = L.w;
  *(uint4*)(c.out + 4 * (size_t)L.mb) = o4;
  L.mb += 1; L.w = 0; L.m0 = 0; L.m1 = 0;
  ml_goto(L, c, c.first_node8, c.row_s, c.first_off);
  if (++L.mx == c.mb_w) {
    // end of a macroblock row: the reference looks at its reader's end-of-data flag here (vp8_dec.c:651-654)
    if (fd_eof(L.d, since)) { L.status = VP8B_NOT_ENOUGH_DATA; L.fail_row = L.my; ml_park(L, c); return; }
    L.mx = 0; L.left = 0;
    if (++L.my == c.mb_h) ml_park(L, c);
  }
}

// One group: top the window up, four decodes.
TK_FN void ml_group(MlLane& L, const MlCtx& c) {
  fd_fill<0>(L.d);
  ml_step(L, c, 1); ml_step(L, c, 2); ml_step(L, c, 3); ml_step(L, c, 4);
  fd_settle(L.d, 4);
}

#endif  // LIBWEBP_B200_VP8_MODES_LOCKSTEP_H_
