// Next row N3: gradient-norm clip + AdamW for the head parameters on the flat fp32 gradient buffer.
//
// reference: Lightning gradient_clip_val = 1.0 -> torch.nn.utils.clip_grad_norm_(params, 1.0)
//            (ospo/utils/train.py:30,50) and torch.optim.AdamW(lr, betas, weight_decay, eps)
//            (ospo/wrapper/train.py:108-115, configs/step5.yaml:37-43), both as PyTorch 2.x defines them:
//   clip:   coef = min(1, max_norm / (sqrt(sum g^2) + 1e-6));  g <- g * coef
//   AdamW:  p <- p * (1 - lr * wd)
//           m <- m + (g - m) * (1 - beta1)                      (lerp)
//           v <- v * beta2 + g * g * (1 - beta2)
//           p <- p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// The head's gradient already is one flat fp32 buffer (dW2 | dW1 | db2 | db1, all-reduced once), so parameters and
// both moments are kept in the same layout and the whole optimizer step is two streaming passes: a deterministic
// squared-norm reduction (336 MB read) and one fused update (g, p, m, v read; p, m, v and the bf16 operand shadow
// written: 30 bytes per element).  Both are HBM-bound; nothing is re-read.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ospo {

constexpr int OPT_THREADS = 256;
constexpr int OPT_VEC = 4;  // float4 per access

// ---- pass 1: sum of squares, fixed summation order (block b sums its grid-stride slice, then a tree) ----
__global__ void __launch_bounds__(OPT_THREADS)
sqnorm_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  const int64_t nvec = n / OPT_VEC;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * OPT_THREADS;
  float acc = 0.0f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(g) + i);
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec * OPT_VEC; i < n; ++i) acc = fmaf(g[i], g[i], acc);  // ragged tail
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += red[w];
    partials[blockIdx.x] = s;
  }
}

// out[0] = sum of the partials in index order (one warp, fixed order: deterministic run to run)
__global__ void __launch_bounds__(32) sqnorm_final_kernel(const float* __restrict__ partials, int count,
                                                           float* __restrict__ out) {
  float acc = 0.0f;
  for (int i = threadIdx.x; i < count; i += 32) acc += partials[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (threadIdx.x == 0) out[0] = acc;
}

// scalars prepared on the host in double precision exactly as torch's Python code does, then rounded to fp32
struct AdamWHyper {
  float decay;       // 1 - lr * weight_decay
  float w1;          // 1 - beta1
  float beta2;
  float w2;          // 1 - beta2
  float bias2_sqrt;  // sqrt(1 - beta2^t)
  float eps;
  float step_size;   // lr / (1 - beta1^t)
  float max_norm;    // <= 0: no clipping
};

__device__ __forceinline__ void adamw_one(float g, float& p, float& m, float& v, float coef, const AdamWHyper& h) {
  g = __fmul_rn(g, coef);
  p = __fmul_rn(p, h.decay);
  m = __fadd_rn(m, __fmul_rn(h.w1, __fsub_rn(g, m)));                                 // lerp_(g, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, h.beta2), __fmul_rn(h.w2, __fmul_rn(g, g)));            // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bias2_sqrt), h.eps);
  p = __fadd_rn(p, __fmul_rn(-h.step_size, __fdiv_rn(m, denom)));                     // addcdiv_(m, denom, -step_size)
}

// ---- pass 2: clip + AdamW, one pass over g / p / m / v; optional bf16 shadow of the first shadow_n elements ----
__global__ void __launch_bounds__(OPT_THREADS)
adamw_kernel(const float* __restrict__ g, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ shadow, int64_t shadow_n, int64_t n, const float* __restrict__ total_sqnorm,
             AdamWHyper h) {
  float coef = 1.0f;
  if (h.max_norm > 0.0f && total_sqnorm != nullptr) {
    const float c = __fdiv_rn(h.max_norm, __fadd_rn(__fsqrt_rn(__ldg(total_sqnorm)), 1e-6f));
    coef = c < 1.0f ? c : 1.0f;
  }
  const int64_t nvec = n / OPT_VEC;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * OPT_THREADS;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += stride) {
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 pv = __ldcs(reinterpret_cast<const float4*>(p) + i);
    float4 mv = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
    adamw_one(gv.x, pv.x, mv.x, vv.x, coef, h);
    adamw_one(gv.y, pv.y, mv.y, vv.y, coef, h);
    adamw_one(gv.z, pv.z, mv.z, vv.z, coef, h);
    adamw_one(gv.w, pv.w, mv.w, vv.w, coef, h);
    __stcs(reinterpret_cast<float4*>(p) + i, pv);
    __stcs(reinterpret_cast<float4*>(m) + i, mv);
    __stcs(reinterpret_cast<float4*>(v) + i, vv);
    if (shadow != nullptr && (i + 1) * OPT_VEC <= shadow_n) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(shadow + i * OPT_VEC) = pk;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec * OPT_VEC; i < n; ++i) {  // ragged tail (and shadow elements that straddle shadow_n)
      float pi = p[i], mi = m[i], vi = v[i];
      adamw_one(g[i], pi, mi, vi, coef, h);
      p[i] = pi;
      m[i] = mi;
      v[i] = vi;
      if (shadow != nullptr && i < shadow_n) shadow[i] = __float2bfloat16_rn(pi);
    }
  }
}

}  // namespace ospo
