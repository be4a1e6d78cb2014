// Epilogue functors for the tcgen05 GEMM engine (gemm_sm100.cuh).
// Each epilogue thread owns ONE accumulator row and receives it 32 columns at a time, so every
// row-wise quantity of the image-token head (online log-sum-exp over the 16384 codes, the
// target-logit gather, the logits row-sum metric) is thread-local: no shuffles, no shared memory.
#pragma once

#include "gemm_sm100.cuh"

namespace ospo {

// exact (erf) GELU and its derivative -- reference: nn.GELU() default approximate='none',
// /root/reference/janus/models/modeling_vlm.py:42
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const float (&v)[32], int valid) {
  if (valid >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      d4[i] = u;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

__device__ __forceinline__ void store_row32_f32(float* dst, const float (&v)[32], int valid) {
  if (valid >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = v[j];
  }
}

__device__ __forceinline__ void load_row32_bf16(const __nv_bfloat16* src, float (&v)[32], int valid) {
  if (valid >= 32 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(s4 + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[8 * i + 2 * k] = __low2float(p);
        v[8 * i + 2 * k + 1] = __high2float(p);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (j < valid) ? __bfloat162float(src[j]) : 0.0f;
  }
}

// ---------------------------------------------------------------------------
// Plain store:  out = acc (+ bias).  TRANSPOSE writes out[col][row] (swap-AB decode GEMMs, where the
// accumulator row is a weight row and the 32 columns are the CFG sample rows); ROW_BIAS indexes the
// bias by accumulator row instead of by column.
// ---------------------------------------------------------------------------
template <typename OutT, bool TRANSPOSE, bool ROW_BIAS>
struct EpiStore {
  struct Params {
    OutT* out;
    int64_t ld;
    const float* bias;  // may be null
  };
  struct State {
    float rb;
  };
  static constexpr int SMEM_BYTES = 0;

  __device__ static void begin(const Params& p, State& st, int row, int, const GemmDims& d, uint8_t*) {
    st.rb = (ROW_BIAS && p.bias != nullptr && row < d.M) ? __ldg(p.bias + row) : 0.0f;
  }
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = d.N - col0;
    if (valid <= 0) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float b = st.rb;
      if (!ROW_BIAS && p.bias != nullptr && j < valid) b = __ldg(p.bias + col0 + j);
      v[j] += b;
    }
    if constexpr (!TRANSPOSE) {
      OutT* dst = p.out + static_cast<int64_t>(row) * p.ld + col0;
      if constexpr (sizeof(OutT) == 4) store_row32_f32(reinterpret_cast<float*>(dst), v, valid);
      else store_row32_bf16(reinterpret_cast<__nv_bfloat16*>(dst), v, valid);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < valid) {
          OutT* dst = p.out + static_cast<int64_t>(col0 + j) * p.ld + row;
          if constexpr (sizeof(OutT) == 4) *reinterpret_cast<float*>(dst) = v[j];
          else *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(v[j]);
        }
      }
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// GEMM1 epilogue:  pre = bf16(acc + b1);  act = bf16(gelu_erf(pre))
// (reference: output_mlp_projector + vision_activation, modeling_vlm.py:47-49; the activation is
//  applied to the bf16-rounded Linear output exactly as the bf16 reference path does.)
// TRANSPOSE: swap-AB decode variant (accumulator row = embed unit, column = sample row).
// ---------------------------------------------------------------------------
template <bool TRANSPOSE, bool STORE_PRE>
struct EpiBiasGelu {
  struct Params {
    const float* bias;     // [E]
    __nv_bfloat16* pre;    // [rows, E] (only if STORE_PRE)
    __nv_bfloat16* act;    // [rows, E]
    int64_t ld;
  };
  struct State {
    float rb;
  };
  static constexpr int SMEM_BYTES = 0;

  __device__ static void begin(const Params& p, State& st, int row, int, const GemmDims& d, uint8_t*) {
    st.rb = (TRANSPOSE && row < d.M) ? __ldg(p.bias + row) : 0.0f;
  }
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = d.N - col0;
    if (valid <= 0) return;
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float b = st.rb;
      if (!TRANSPOSE && j < valid) b = __ldg(p.bias + col0 + j);
      v[j] = bf16_round(v[j] + b);
      a[j] = gelu_erf(v[j]);
    }
    if constexpr (!TRANSPOSE) {
      const int64_t off = static_cast<int64_t>(row) * p.ld + col0;
      if constexpr (STORE_PRE) store_row32_bf16(p.pre + off, v, valid);
      store_row32_bf16(p.act + off, a, valid);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < valid) {
          const int64_t off = static_cast<int64_t>(col0 + j) * p.ld + row;
          if constexpr (STORE_PRE) p.pre[off] = __float2bfloat16_rn(v[j]);
          p.act[off] = __float2bfloat16_rn(a[j]);
        }
      }
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// GEMM2 forward epilogue:  logits = bf16(acc + b2), plus -- without ever materialising fp32 logits or a
// log-softmax tensor -- per (row, N-tile) partial (max, sum-exp), the gathered target logit and the
// partial row-sum of the logits.  The bf16 logits are (optionally) spilled once for the backward pass.
// reference: vision_head Linear (modeling_vlm.py:50) + log_softmax/gather (ospo/wrapper/train.py:391).
// The log-sum-exp is taken over the bf16-rounded logits in fp32, which is what the bf16 reference
// path computes (Linear output bf16, log_softmax autocast to fp32).
// ---------------------------------------------------------------------------
struct EpiLogitsLse {
  struct Params {
    const float* bias;          // [V]
    __nv_bfloat16* logits;      // [rows, V] spill (may be null)
    int64_t ld;
    const int64_t* labels;      // [rows] target code per row
    float2* part;               // [num_n, rows] (max, sumexp) partials
    float* rowsum_part;         // [num_n, rows] partial sums of logits (may be null)
    float* tgt;                 // [rows] gathered target logit
  };
  struct State {
    float m, s, sum, tgt;
    int label;
    bool hit;
  };
  static constexpr int SMEM_BYTES = 0;

  __device__ static void begin(const Params& p, State& st, int row, int, const GemmDims& d, uint8_t*) {
    st.m = -INFINITY;
    st.s = 0.0f;
    st.sum = 0.0f;
    st.tgt = 0.0f;
    st.hit = false;
    st.label = (row < d.M) ? static_cast<int>(__ldg(p.labels + row)) : -1;
  }
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = d.N - col0;
    if (valid <= 0) return;
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float b = (j < valid) ? __ldg(p.bias + col0 + j) : 0.0f;
      v[j] = bf16_round(v[j] + b);
      if (j < valid) cmax = fmaxf(cmax, v[j]);
    }
    if (p.logits != nullptr) store_row32_bf16(p.logits + static_cast<int64_t>(row) * p.ld + col0, v, valid);
    constexpr float LOG2E = 1.4426950408889634f;
    if (cmax > st.m) {
      st.s *= exp2f((st.m - cmax) * LOG2E);  // exp2f(-inf) = 0 on the first chunk
      st.m = cmax;
    }
    const float mneg = -st.m * LOG2E;
    float acc = 0.0f, lsum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < valid) {
        acc += exp2f(fmaf(v[j], LOG2E, mneg));
        lsum += v[j];
      }
    }
    st.s += acc;
    st.sum += lsum;
    const int rel = st.label - col0;
    if (rel >= 0 && rel < valid && rel < 32) {
      st.hit = true;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == rel) st.tgt = v[j];
    }
  }
  __device__ static void end(const Params& p, State& st, int row, int, int n_blk, const GemmDims& d, uint8_t*) {
    if (row >= d.M) return;
    const int64_t idx = static_cast<int64_t>(n_blk) * d.M + row;
    p.part[idx] = make_float2(st.m, st.s);
    if (p.rowsum_part != nullptr) p.rowsum_part[idx] = st.sum;
    if (st.hit) p.tgt[row] = st.tgt;  // only the N-tile that holds the label column writes
  }
};

// ---------------------------------------------------------------------------
// dAct GEMM epilogue:  dpre = bf16(acc * gelu'(pre))      (autograd of GELU)
// ---------------------------------------------------------------------------
struct EpiGeluBwd {
  struct Params {
    const __nv_bfloat16* pre;  // [rows, E]
    __nv_bfloat16* dpre;       // [rows, E]
    int64_t ld;
  };
  struct State {};
  static constexpr int SMEM_BYTES = 0;
  __device__ static void begin(const Params&, State&, int, int, const GemmDims&, uint8_t*) {}
  __device__ static void chunk(const Params& p, State&, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = d.N - col0;
    if (valid <= 0) return;
    const int64_t off = static_cast<int64_t>(row) * p.ld + col0;
    float pre[32];
    load_row32_bf16(p.pre + off, pre, valid);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]) * gelu_erf_grad(pre[j]);
    store_row32_bf16(p.dpre + off, v, valid);
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

}  // namespace ospo
