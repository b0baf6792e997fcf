// chain_floor.cu -- microbenchmarks behind the token parser's roofline.
//
// The token parse is one dependent chain per stream: what bounds it is not HBM but (a) the latency of the boolean
// decoder's dependent chain and (b) how fast ONE warp can issue into the pipes of its SM sub-partition. This tool
// measures both on the device it runs on, with one (or two) warps per sub-partition -- the parser's launch shape on
// BASELINE config 2 -- and `lanes` of the 32 lanes active:
//
//   lat_X      cycles per op of a dependent chain of X (32 ops per loop iteration, inline PTX so that nothing is
//              folded or hoisted; `cuobjdump -sass` of this binary shows which SASS op each became)
//   thr_X      cycles per op of 8 independent chains of X issued by one warp: the issue rate one warp gets from a pipe
//   booldec_int   one boolean decode (bit_reader_inl_utils.h:107-136 as restated in vp8_parse_core.h:bd_decode):
//                 multiply-high -> multiply-add -> compare -> select -> find-leading-one -> shift -> add
//   booldec_fp    the same decode with the split and the renormalisation shift taken from fp32 arithmetic
//                 (fma.rz, exponent field) instead of IMAD.HI + FLO; bit-exact (checked below against the integer form)
//   *_lds      the same with the next probability fetched from shared memory at an address chosen by the bit
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o chain_floor chain_floor.cu
// Run:   ./chain_floor            (prints one JSON object)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 512

#define R4(x) x x x x
#define R8(x) R4(x) R4(x)
#define R32(x) R8(x) R8(x) R8(x) R8(x)

#define TIMED(body, ops)                                                  \
  do {                                                                    \
    __syncwarp();                                                         \
    const long long t0_ = clock64();                                      \
    _Pragma("unroll 1") for (int it = 0; it < ITERS; ++it) { body }      \
    const long long t1_ = clock64();                                      \
    cyc = (float)(t1_ - t0_) / (float)(ITERS * (ops));                    \
  } while (0)

// ---- the two boolean decoders under test (state: V window, vlo, range, nbits left out) ----------------------------
struct BdInt { uint32_t V, vlo, R24; };
__device__ __forceinline__ int bd_int_step(BdInt& d, uint32_t prob) {
  const uint32_t s1 = (__umulhi(d.R24, prob) << 24) + (1u << 24);
  const int bit = d.V >= s1;
  uint32_t r = s1;
  if (bit) { r = d.R24 + (1u << 24) - s1; d.V -= s1; }
  uint32_t sh; asm("bfind.shiftamt.u32 %0, %1;" : "=r"(sh) : "r"(r));
  d.R24 = (r << sh) - (1u << 24);
  d.V = __funnelshift_l(d.vlo, d.V, sh);
  d.vlo <<= sh;
  return bit;
}
// fp32 form. Rs = (range - 1) / 256 (exact), Rp = (range - 1) + 2^23. prob as a float.
//   m  = fma.rz(Rs, prob, 2^23)        = 2^23 + floor((range-1) * prob / 256) = 2^23 + split          (exact: rz truncates)
//   f0 = m - (2^23 - 1)                = float(split + 1)        = new range if bit == 0
//   f1 = Rp - m                        = float(range-1 - split)  = new range if bit == 1
//   shift = 134 - exponent(f), normalised range - 1 = mantissa with the exponent forced to 2^7, minus 1
struct BdFp { uint32_t V, vlo; float Rs, Rp; };
__device__ __forceinline__ int bd_fp_step(BdFp& d, float prob) {
  const float m = __fmaf_rz(d.Rs, prob, 8388608.0f);
  const uint32_t s1 = (__float_as_uint(m) << 24) + (1u << 24);
  const int bit = d.V >= s1;
  const float f0 = m - 8388607.0f, f1 = d.Rp - m;
  const float f = bit ? f1 : f0;
  if (bit) d.V -= s1;
  const uint32_t fb = __float_as_uint(f);
  const uint32_t sh = 134u - (fb >> 23);
  const float fn = __uint_as_float((fb & 0x007fffffu) | 0x43000000u);   // range in [128, 256)
  d.Rs = (fn - 1.0f) * 0.00390625f;
  d.Rp = fn + 8388607.0f;
  d.V = __funnelshift_l(d.vlo, d.V, sh);
  d.vlo <<= sh;
  return bit;
}

__global__ void k_floor(int test, int lanes, uint32_t seed, float* out, uint32_t* sink) {
  __shared__ uint32_t tab[1024 + 8];
  for (int k = threadIdx.x; k < 1024 + 8; k += blockDim.x) tab[k] = (uint32_t)(((k * 37 + 11) & 1023) * 4u);   // pointer-chase permutation (byte offsets)
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (lane >= lanes) return;
  uint32_t a = seed + lane, b = seed * 3 + 1, c = seed ^ 0x55, d = seed + 7, e = a ^ b, f = b ^ c, g = c ^ d, h = d ^ a;
  const uint32_t k1 = seed | 1, k2 = (seed >> 3) | 3;
  float fa = (float)(seed & 255) + 1.f, fk = 1.0000001f;
  float cyc = 0.f;
  switch (test) {
    case 0: TIMED(R32(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(k1), "r"(k2));), 32); break;                 // lat_lop3
    case 1: TIMED(R32(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(k1), "r"(k2));), 32); break;                      // lat_imad
    case 2: TIMED(R32(asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(k1), "r"(k2));), 32); break;                  // lat_shf
    case 3: TIMED(R32(asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(k1), "r"(k2));), 32); break;                      // lat_imadhi
    case 4: TIMED(R32(asm volatile("bfind.shiftamt.u32 %0, %0;" : "+r"(a));), 32); break;                                         // lat_flo
    case 5: TIMED(R32(asm volatile("popc.b32 %0, %0;" : "+r"(a));), 32); break;                                                   // lat_popc
    case 6: TIMED(R32(asm volatile("fma.rz.f32 %0, %0, %1, %1;" : "+f"(fa) : "f"(fk));), 32); a = __float_as_uint(fa); break;     // lat_ffma
    case 7: TIMED(R32(asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a) : "r"(k1), "r"(k2));), 32); break;   // lat_setp_selp (2 ops)
    case 8: TIMED(R32(asm volatile("mad.lo.u32 %0, %0, %1, %2; lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(k1), "r"(k2));), 64); break;   // lat_cross (imad <-> lop3)
    case 9:   // lat_lds
      a = (a & 1023u) * 4u;
      TIMED(R32(a = *(const volatile uint32_t*)((const char*)tab + a);), 32);
      break;
    case 10:  // thr_lop3: 8 independent chains
      TIMED(R4(asm volatile("lop3.b32 %0, %0, %8, %9, 0x96; lop3.b32 %1, %1, %8, %9, 0x96; lop3.b32 %2, %2, %8, %9, 0x96; lop3.b32 %3, %3, %8, %9, 0x96;"
                            "lop3.b32 %4, %4, %8, %9, 0x96; lop3.b32 %5, %5, %8, %9, 0x96; lop3.b32 %6, %6, %8, %9, 0x96; lop3.b32 %7, %7, %8, %9, 0x96;"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k1), "r"(k2));), 32);
      break;
    case 11:  // thr_imad
      TIMED(R4(asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                            "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k1), "r"(k2));), 32);
      break;
    case 12:  // thr_mix: 4 lop3 + 4 imad, interleaved
      TIMED(R4(asm volatile("lop3.b32 %0, %0, %8, %9, 0x96; mad.lo.u32 %1, %1, %8, %9; lop3.b32 %2, %2, %8, %9, 0x96; mad.lo.u32 %3, %3, %8, %9;"
                            "lop3.b32 %4, %4, %8, %9, 0x96; mad.lo.u32 %5, %5, %8, %9; lop3.b32 %6, %6, %8, %9, 0x96; mad.lo.u32 %7, %7, %8, %9;"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k1), "r"(k2));), 32);
      break;
    case 13:  // thr_shf
      TIMED(R4(asm volatile("shf.l.wrap.b32 %0, %0, %8, %9; shf.l.wrap.b32 %1, %1, %8, %9; shf.l.wrap.b32 %2, %2, %8, %9; shf.l.wrap.b32 %3, %3, %8, %9;"
                            "shf.l.wrap.b32 %4, %4, %8, %9; shf.l.wrap.b32 %5, %5, %8, %9; shf.l.wrap.b32 %6, %6, %8, %9; shf.l.wrap.b32 %7, %7, %8, %9;"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k1), "r"(k2));), 32);
      break;
    case 14: {  // thr_selp: 8 independent selects on one predicate
      TIMED(R4(asm volatile("{.reg .pred p; setp.ge.u32 p, %8, %9; selp.u32 %0, %9, %0, p; selp.u32 %1, %9, %1, p; selp.u32 %2, %9, %2, p; selp.u32 %3, %9, %3, p;"
                            "selp.u32 %4, %9, %4, p; selp.u32 %5, %9, %5, p; selp.u32 %6, %9, %6, p; selp.u32 %7, %9, %7, p;}"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k1), "r"(k2));), 36);
      break;
    }
    case 15: {  // booldec_int, probability in a register
      BdInt s; s.V = a | 0x80000000u; s.vlo = b; s.R24 = 254u << 24;
      const uint32_t prob = (seed & 127u) + 64u;
      TIMED(R8(a += bd_int_step(s, prob); s.vlo |= 0x10101u;), 8);
      a ^= s.V ^ s.R24;
      break;
    }
    case 16: {  // booldec_fp
      BdFp s; s.V = a | 0x80000000u; s.vlo = b; s.Rs = 254.f / 256.f; s.Rp = 254.f + 8388608.f;
      const float prob = (float)((seed & 127u) + 64u);
      TIMED(R8(a += bd_fp_step(s, prob); s.vlo |= 0x10101u;), 8);
      a ^= s.V ^ __float_as_uint(s.Rs);
      break;
    }
    case 17: {  // booldec_int_lds: next probability from shared memory at an address chosen by the bit
      BdInt s; s.V = a | 0x80000000u; s.vlo = b; s.R24 = 254u << 24;
      uint32_t st = (a & 1023u) * 4u, prob = (seed & 127u) + 64u;
      TIMED(R8({ const int bit = bd_int_step(s, prob); st = *(const volatile uint32_t*)((const char*)tab + st + (bit ? 4u : 0u));
                 prob = ((st >> 4) & 127u) + 64u; s.vlo |= 0x10101u; }), 8);
      a ^= s.V ^ s.R24 ^ st;
      break;
    }
    case 18: {  // booldec_fp_lds
      BdFp s; s.V = a | 0x80000000u; s.vlo = b; s.Rs = 254.f / 256.f; s.Rp = 254.f + 8388608.f;
      uint32_t st = (a & 1023u) * 4u; float prob = (float)((seed & 127u) + 64u);
      TIMED(R8({ const int bit = bd_fp_step(s, prob); st = *(const volatile uint32_t*)((const char*)tab + st + (bit ? 4u : 0u));
                 prob = __uint_as_float(0x4b000000u | (((st >> 4) & 127u) + 64u)) - 8388608.0f; s.vlo |= 0x10101u; }), 8);
      a ^= s.V ^ __float_as_uint(s.Rs) ^ st;
      break;
    }
    case 19: {  // booldec_fp_pair: both outcomes' successors already loaded (look-ahead): the chain is select -> fma only
      BdFp s; s.V = a | 0x80000000u; s.vlo = b; s.Rs = 254.f / 256.f; s.Rp = 254.f + 8388608.f;
      uint32_t st = (a & 1023u) * 4u; float p0 = 100.f, p1 = 190.f, prob = 128.f; uint32_t n0 = st, n1 = st ^ 64u;
      TIMED(R8({ const int bit = bd_fp_step(s, prob); prob = bit ? p1 : p0; st = bit ? n1 : n0;
                 const uint2 nx = *(const uint2*)((const char*)tab + (st & 0xff8u));
                 n0 = nx.x; n1 = nx.y; p0 = __uint_as_float(0x43000000u | ((nx.x & 0x7fu) << 16)); p1 = __uint_as_float(0x42800000u | ((nx.y & 0x7fu) << 16));
                 s.vlo |= 0x10101u; }), 8);
      a ^= s.V ^ __float_as_uint(s.Rs) ^ st;
      break;
    }
  }
  if (lane == 0 && (threadIdx.x >> 5) == 0) out[blockIdx.x] = cyc;
  if ((a ^ b ^ c ^ d ^ e ^ f ^ g ^ h) == 0x12345678u) sink[0] = a;
}

// ---- bit-exactness of the fp32 decoder against the integer one: random probabilities, random bit windows
__global__ void k_check(uint32_t seed, int steps, unsigned long long* mismatches) {
  uint32_t x = seed + 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  BdInt a; BdFp f;
  a.V = f.V = x * 2654435761u; a.vlo = f.vlo = x ^ 0xdeadbeefu; a.R24 = 254u << 24; f.Rs = 254.f / 256.f; f.Rp = 254.f + 8388608.f;
  unsigned long long bad = 0;
  for (int k = 0; k < steps; ++k) {
    x = x * 1664525u + 1013904223u;
    const uint32_t prob = (x >> 24);               // 0..255 (0 occurs in the reference's tables too)
    x = x * 1664525u + 1013904223u;
    const int b0 = bd_int_step(a, prob), b1 = bd_fp_step(f, (float)prob);
    a.vlo |= x & 0xffffu; f.vlo |= x & 0xffffu;    // keep feeding bits
    const uint32_t r_int = a.R24 >> 24, r_fp = (uint32_t)(f.Rs * 256.f);
    if (b0 != b1 || a.V != f.V || a.vlo != f.vlo || r_int != r_fp || f.Rp != (float)r_fp + 8388608.f) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

int main() {
  const char* names[] = { "lat_lop3", "lat_imad", "lat_shf", "lat_imadhi", "lat_flo", "lat_popc", "lat_ffma_rz", "lat_setp_selp_pair", "lat_cross_imad_lop3",
                          "lat_lds", "thr_lop3", "thr_imad", "thr_mix_lop3_imad", "thr_shf", "thr_selp", "booldec_int", "booldec_fp", "booldec_int_lds",
                          "booldec_fp_lds", "booldec_fp_lookahead" };
  const int ntests = 20;
  float* d_out; uint32_t* d_sink; unsigned long long* d_bad;
  cudaMalloc(&d_out, 256 * sizeof(float)); cudaMalloc(&d_sink, 4); cudaMalloc(&d_bad, 8);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int nsm = pr.multiProcessorCount < 256 ? pr.multiProcessorCount : 256;
  // bit-exactness first
  cudaMemset(d_bad, 0, 8);
  k_check<<<nsm * 4, 256>>>(20261018u, 20000, d_bad);
  unsigned long long bad = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "k_check failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"iters\": %d, \"fp_vs_int_decodes_checked\": %llu, \"fp_vs_int_mismatches\": %llu, "
         "\"unit\": \"cycles per op (per decode for booldec*); key = test.w<warps per SM sub-partition>.l<active lanes>\", \"results\": {",
         pr.name, pr.multiProcessorCount, pr.clockRate, ITERS, (unsigned long long)nsm * 4 * 256 * 20000ull, bad);
  const int lane_set[] = { 1, 7, 32 };
  const int warp_set[] = { 1, 2 };
  bool first = true;
  for (int t = 0; t < ntests; ++t) {
    for (int wi = 0; wi < 2; ++wi) {
      for (int li = 0; li < 3; ++li) {
        float host[256];
        for (int rep = 0; rep < 2; ++rep) {   // the second run is the measurement (the first warms the instruction cache)
          k_floor<<<nsm, 128 * warp_set[wi]>>>(t, lane_set[li], 12345u + rep, d_out, d_sink);
          if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "test %s failed: %s\n", names[t], cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        cudaMemcpy(host, d_out, sizeof(float) * nsm, cudaMemcpyDeviceToHost);
        float sum = 0;
        for (int k = 0; k < nsm; ++k) sum += host[k];
        printf("%s\"%s.w%d.l%d\": %.2f", first ? "" : ", ", names[t], warp_set[wi], lane_set[li], sum / nsm);
        first = false;
      }
    }
  }
  printf("}}\n");
  return 0;
}
