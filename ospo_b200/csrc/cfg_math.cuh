// Arithmetic shared by the CFG sampler kernels (head_kernels.cuh) and the fused decode-GEMM epilogue
// (epilogues.cuh): the op-by-op CFG merge, the power-of-two-relative softmax weights and the fixed summation
// order.  oracle/cfg_sample.c restates every function here operation for operation.
#pragma once

#include <cstring>

#include "ptx.cuh"

namespace ospo {

constexpr int SAMPLE_THREADS = 512;
constexpr int SAMPLE_SEG = 32;    // codes per segment (one thread / one warp lane set)
constexpr int SAMPLE_GRP = 32;    // segments per group
constexpr int SAMPLE_TILE = 128;  // codes per tile (= accumulator rows of one decode-GEMM CTA)

// e^t = P(r) * 2^n; returns P(r), writes n (integer-valued float).  One IEEE fp32 op per line.
__device__ __forceinline__ float exp_parts(float t, float& n) {
  float y = __fmul_rn(t, 1.4426950408889634f);
  y = fmaxf(y, -1.0e4f);
  y = fminf(y, 1.0e4f);
  n = rintf(y);
  float r = __fmaf_rn(n, -0.693145751953125f, t);
  r = __fmaf_rn(n, -1.42860682030941723212e-6f, r);
  float p = 1.3888888888888889e-03f;
  p = __fmaf_rn(p, r, 8.3333333333333332e-03f);
  p = __fmaf_rn(p, r, 4.1666666666666664e-02f);
  p = __fmaf_rn(p, r, 1.6666666666666666e-01f);
  p = __fmaf_rn(p, r, 0.5f);
  p = __fmaf_rn(p, r, 1.0f);
  p = __fmaf_rn(p, r, 1.0f);
  return p;
}
__device__ __forceinline__ float exp_n_only(float t) {
  float y = __fmul_rn(t, 1.4426950408889634f);
  y = fmaxf(y, -1.0e4f);
  y = fminf(y, 1.0e4f);
  return rintf(y);
}
// 2^e for integer-valued e <= 0; 0 below -120
__device__ __forceinline__ float pow2_factor(float e) {
  const float f = __int_as_float((static_cast<int>(e) + 127) << 23);
  return (e < -120.0f) ? 0.0f : f;
}

// ---- the same weight for two codes at once on the packed fp32 pipe (FFMA2 / FMUL2 / FADD2, sm_100) ----------
// Every lane of an f32x2 instruction is the IEEE operation of its scalar form, so the results are the bits
// exp_parts + pow2_factor give.  Two more changes keep the conversion unit out of the loop: rint(y) is
// (y + 1.5 * 2^23) - 1.5 * 2^23 (round-to-nearest-even for |y| < 2^22; y is clamped to 1e4), and n - kt is taken
// from the low mantissa bits of y + 1.5 * 2^23 as an integer (exp_koff(kt) = the bits of kt + 1.5 * 2^23).
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// a + b for a product a that must keep its own rounding.  ptxas 12.9 contracts mul.rn.f32x2 followed by add.rn.f32x2
// into one FFMA2 (the "explicit rounding mode is never fused" rule of the scalar forms is not applied to the packed
// ones, with or without --fmad=false), which rounds t log2 e + 1.5 * 2^23 once instead of twice and can move n by one
// at a tie.  Written as fma(a, 1, b) -- exactly a + b, one rounding -- there is no multiply-add pair left to contract.
// tests/test_abi_cpu.py checks the library's SASS for a fused form of this step.
__device__ __forceinline__ uint64_t f2_add_nofuse(uint64_t a, uint64_t b) { return f2_fma(a, f2_pack(1.0f, 1.0f), b); }
constexpr float kRintMagic = 12582912.0f;  // 1.5 * 2^23 = 0x4B400000
__device__ __forceinline__ int exp_koff(float kt) { return __float_as_int(__fadd_rn(kt, kRintMagic)); }
// 2^e for an integer e <= 0; 0 below -120 (pow2_factor on an integer)
__device__ __forceinline__ float pow2_factor_i(int e) {
  const float f = __int_as_float((e << 23) + 0x3F800000);
  return (e < -120) ? 0.0f : f;
}
// w = P(r) 2^(n - kt) for two codes (packed in, packed out); koff = exp_koff(kt) of each code's tile
__device__ __forceinline__ uint64_t exp_weight2p(uint64_t t, int koff0, int koff1) {
  float y0, y1;
  f2_unpack(f2_mul(t, f2_pack(1.4426950408889634f, 1.4426950408889634f)), y0, y1);
  y0 = fminf(fmaxf(y0, -1.0e4f), 1.0e4f);
  y1 = fminf(fmaxf(y1, -1.0e4f), 1.0e4f);
  const uint64_t ym = f2_add(f2_pack(y0, y1), f2_pack(kRintMagic, kRintMagic));
  const uint64_t n = f2_add(ym, f2_pack(-kRintMagic, -kRintMagic));
  float ym0, ym1;
  f2_unpack(ym, ym0, ym1);
  const float f0 = pow2_factor_i(__float_as_int(ym0) - koff0);
  const float f1 = pow2_factor_i(__float_as_int(ym1) - koff1);
  uint64_t r = f2_fma(n, f2_pack(-0.693145751953125f, -0.693145751953125f), t);
  r = f2_fma(n, f2_pack(-1.42860682030941723212e-6f, -1.42860682030941723212e-6f), r);
  uint64_t p = f2_pack(1.3888888888888889e-03f, 1.3888888888888889e-03f);
  p = f2_fma(p, r, f2_pack(8.3333333333333332e-03f, 8.3333333333333332e-03f));
  p = f2_fma(p, r, f2_pack(4.1666666666666664e-02f, 4.1666666666666664e-02f));
  p = f2_fma(p, r, f2_pack(1.6666666666666666e-01f, 1.6666666666666666e-01f));
  p = f2_fma(p, r, f2_pack(0.5f, 0.5f));
  p = f2_fma(p, r, f2_pack(1.0f, 1.0f));
  p = f2_fma(p, r, f2_pack(1.0f, 1.0f));
  return f2_mul(p, f2_pack(f0, f1));
}
// For callers that have checked, for every code they pass, that the clamp is idle (|t log2 e| <= 1e4): the clamp-free
// form, with the 2^-120 cut-off folded into the exponent construction: the polynomial is evaluated
// with every coefficient times 64 (a power-of-two scaling commutes with each rounding of the Horner chain, so
// P64(r) = 64 P(r) bit for bit) and the scale is 2^(n - kt - 6), whose bits are  max(n - kt + 121, 0) << 23:
// for n - kt >= -120 that is the normal number 2^(n - kt - 6) and P64(r) 2^(n - kt - 6) = P(r) 2^(n - kt) exactly
// (both factors and the product are normal), below the cut-off it is +0.  One add-max and one shift per code
// instead of a subtraction, a shift-add, a compare and a select.  kcut = exp_koff(kt) - 121.
__device__ __forceinline__ uint64_t exp_weight2p_cut(uint64_t t, int kcut) {
  const uint64_t y = f2_mul(t, f2_pack(1.4426950408889634f, 1.4426950408889634f));
  const uint64_t ym = f2_add_nofuse(y, f2_pack(kRintMagic, kRintMagic));
  const uint64_t n = f2_add(ym, f2_pack(-kRintMagic, -kRintMagic));
  float ym0, ym1;
  f2_unpack(ym, ym0, ym1);
  const float f0 = __int_as_float(max(__float_as_int(ym0) - kcut, 0) << 23);
  const float f1 = __int_as_float(max(__float_as_int(ym1) - kcut, 0) << 23);
  uint64_t r = f2_fma(n, f2_pack(-0.693145751953125f, -0.693145751953125f), t);
  r = f2_fma(n, f2_pack(-1.42860682030941723212e-6f, -1.42860682030941723212e-6f), r);
  uint64_t p = f2_pack(64.0f * 1.3888888888888889e-03f, 64.0f * 1.3888888888888889e-03f);
  p = f2_fma(p, r, f2_pack(64.0f * 8.3333333333333332e-03f, 64.0f * 8.3333333333333332e-03f));
  p = f2_fma(p, r, f2_pack(64.0f * 4.1666666666666664e-02f, 64.0f * 4.1666666666666664e-02f));
  p = f2_fma(p, r, f2_pack(64.0f * 1.6666666666666666e-01f, 64.0f * 1.6666666666666666e-01f));
  p = f2_fma(p, r, f2_pack(32.0f, 32.0f));
  p = f2_fma(p, r, f2_pack(64.0f, 64.0f));
  p = f2_fma(p, r, f2_pack(64.0f, 64.0f));
  return f2_mul(p, f2_pack(f0, f1));
}
// The same weights again when, in addition, no code of the caller's batch underflows the 2^-120 cut-off
// (n - kt >= -120 for all of them): 2^(n - kt) is then always the plain exponent-field construction, and its bits are
// (ym_bits << 23) + kbias with kbias = 0x3F800000 - (koff << 23) -- one shift-add per code instead of a subtraction,
// a shift-add, a compare and a select.  exp_kbias(koff) is computed once per segment.
__device__ __forceinline__ uint32_t exp_kbias(int koff) { return 0x3F800000u - (static_cast<uint32_t>(koff) << 23); }
__device__ __forceinline__ uint64_t exp_weight2p_nounderflow(uint64_t t, uint32_t kbias) {
  const uint64_t y = f2_mul(t, f2_pack(1.4426950408889634f, 1.4426950408889634f));
  const uint64_t ym = f2_add_nofuse(y, f2_pack(kRintMagic, kRintMagic));
  const uint64_t n = f2_add(ym, f2_pack(-kRintMagic, -kRintMagic));
  float ym0, ym1;
  f2_unpack(ym, ym0, ym1);
  const float f0 = __uint_as_float((__float_as_uint(ym0) << 23) + kbias);
  const float f1 = __uint_as_float((__float_as_uint(ym1) << 23) + kbias);
  uint64_t r = f2_fma(n, f2_pack(-0.693145751953125f, -0.693145751953125f), t);
  r = f2_fma(n, f2_pack(-1.42860682030941723212e-6f, -1.42860682030941723212e-6f), r);
  uint64_t p = f2_pack(1.3888888888888889e-03f, 1.3888888888888889e-03f);
  p = f2_fma(p, r, f2_pack(8.3333333333333332e-03f, 8.3333333333333332e-03f));
  p = f2_fma(p, r, f2_pack(4.1666666666666664e-02f, 4.1666666666666664e-02f));
  p = f2_fma(p, r, f2_pack(1.6666666666666666e-01f, 1.6666666666666666e-01f));
  p = f2_fma(p, r, f2_pack(0.5f, 0.5f));
  p = f2_fma(p, r, f2_pack(1.0f, 1.0f));
  p = f2_fma(p, r, f2_pack(1.0f, 1.0f));
  return f2_mul(p, f2_pack(f0, f1));
}
// NaN-propagating three-input max / min (the range check above must fail when a NaN is present)
__device__ __forceinline__ float max3_nan(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float min3_nan(float a, float b, float c) {
  float d;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void exp_weight2(float t0, float t1, int koff0, int koff1, float& w0, float& w1) {
  f2_unpack(exp_weight2p(f2_pack(t0, t1), koff0, koff1), w0, w1);
}

// merge two adjacent codes at once so the bf16 roundings can use the packed convert (cvt.rn.bf16x2.f32).
// MODE 0 = bf16 rounding after every op, MODE 1 = fp32.  TDIV = false skips the division (T == 1: x / 1 == x).
__device__ __forceinline__ void round2_bf16(float& a, float& b) {
  const uint32_t u = pack_bf16x2(a, b);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xFFFF0000u);
}
template <int MODE, bool TDIV>
__device__ __forceinline__ void cfg_merge_vals(float c0, float c1, float u0, float u1, float w, float T, float& t0,
                                               float& t1) {
  float d0 = __fsub_rn(c0, u0), d1 = __fsub_rn(c1, u1);
  if (MODE == 0) round2_bf16(d0, d1);
  float e0 = __fmul_rn(w, d0), e1 = __fmul_rn(w, d1);
  if (MODE == 0) round2_bf16(e0, e1);
  t0 = __fadd_rn(u0, e0);
  t1 = __fadd_rn(u1, e1);
  if (MODE == 0) round2_bf16(t0, t1);
  if (TDIV) {
    t0 = __fdiv_rn(t0, T);
    t1 = __fdiv_rn(t1, T);
    if (MODE == 0) round2_bf16(t0, t1);
  }
}
template <int MODE, bool TDIV>
__device__ __forceinline__ void cfg_merge2(uint32_t wc, uint32_t wu, float w, float T, float& t0, float& t1) {
  cfg_merge_vals<MODE, TDIV>(__uint_as_float(wc << 16), __uint_as_float(wc & 0xFFFF0000u), __uint_as_float(wu << 16),
                             __uint_as_float(wu & 0xFFFF0000u), w, T, t0, t1);
}

// The same merge on the bf16 pipe (MODE 0 with a cfg_weight that is exactly a bf16 value, e.g. the default 5.0):
// sub / mul / add.rn.bf16x2 round the exact result once; the reference rounds fp32(a op b) to bf16.  The two agree
// bit for bit: for operands with 8-bit significands the fp32 result is exact unless the exponents differ by more
// than 16, and then it lies within 2^-16 of the larger operand, far from any bf16 rounding boundary; a product of
// two 8-bit significands always fits fp32.  Inline PTX, so the compiler cannot contract mul + add into an fma.
__device__ __forceinline__ uint32_t bf16x2_sub(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// is w exactly representable in bf16 (then the packed path may be used)?
__host__ __device__ inline bool bf16_exact(float w) {
  uint32_t u;
  memcpy(&u, &w, 4);
  return (u & 0xFFFFu) == 0u;
}
// wc = two conditional logits, wu = the two unconditional ones (packed bf16), w2 = cfg_weight in both halves
template <bool TDIV>
__device__ __forceinline__ void cfg_merge2_hw(uint32_t wc, uint32_t wu, uint32_t w2, float T, float& t0, float& t1) {
  const uint32_t m = bf16x2_add(wu, bf16x2_mul(w2, bf16x2_sub(wc, wu)));
  t0 = __uint_as_float(m << 16);
  t1 = __uint_as_float(m & 0xFFFF0000u);
  if (TDIV) {
    t0 = __fdiv_rn(t0, T);
    t1 = __fdiv_rn(t1, T);
    round2_bf16(t0, t1);
  }
}

// butterfly tree sum of a segment's 32 weights: x[j] += x[j + 16] (j < 16), then strides 8, 4, 2, 1 -- the order in
// which a warp combines lanes with shfl_xor 16, 8, 4, 2, 1 (IEEE addition is commutative, so which lane adds is
// irrelevant).  The weights arrive as 16 packed pairs (x[i] = codes 2i, 2i + 1): strides 16, 8, 4 and 2 are packed
// additions.
__device__ __forceinline__ float tree_sum32_packed(const uint64_t (&x)[16]) {
  uint64_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = f2_add(x[i], x[i + 8]);
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = f2_add(a[i], a[i + 4]);
#pragma unroll
  for (int i = 0; i < 2; ++i) a[i] = f2_add(a[i], a[i + 2]);
  float lo, hi;
  f2_unpack(f2_add(a[0], a[1]), lo, hi);
  return __fadd_rn(lo, hi);
}

// Transpose-reduce: every lane holds 16 values (one per CFG pair); reduce each of the 16 across the 32 lanes with
// 16 shuffles instead of 80.  Lanes exchange half of their values at every step (strides 16, 8, 4, 2) and combine
// the last pair with stride 1.  On return lane l holds the full reduction of pair  k = (l >> 1) & 15  in the bit
// order  k = b4*8 + b3*4 + b2*2 + b1  (b_i = bit i of l); lanes l and l^1 hold the same pair.
struct OpSum {
  __device__ __forceinline__ float operator()(float a, float b) const { return __fadd_rn(a, b); }
};
struct OpMax {
  __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
template <class Op>
__device__ __forceinline__ float warp_transpose_reduce16(const float (&x)[16], int lane, Op op) {
  float y[8], z[4], w[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = b4 ? x[j] : x[j + 8];
    const float keep = b4 ? x[j + 8] : x[j];
    y[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = b3 ? y[j] : y[j + 4];
    const float keep = b3 ? y[j + 4] : y[j];
    z[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = b2 ? z[j] : z[j + 2];
    const float keep = b2 ? z[j + 2] : z[j];
    w[j] = op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
  const float send = b1 ? w[0] : w[1];
  const float keep = b1 ? w[1] : w[0];
  const float v = op(keep, __shfl_xor_sync(0xffffffffu, send, 2));
  return op(v, __shfl_xor_sync(0xffffffffu, v, 1));
}
__device__ __forceinline__ int transpose_reduce_pair_of_lane(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

struct CfgFusedBuffers {
  float* wbuf;        // [P, V]      weights relative to K_tile
  float* seg_sum;     // [P, V/32]   tree sums relative to K_tile
  float* tile_k;      // [P, V/128]  tile exponents
  float* tile_max;    // [P, V/128]  greedy: max merged logit of the tile
  int* tile_arg;      // [P, V/128]  greedy: its (lowest) index
};

}  // namespace ospo
