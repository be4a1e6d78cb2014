// Epilogue functors for the tcgen05 GEMM engine (gemm_sm100.cuh).
// Each epilogue thread owns ONE accumulator row and receives it 32 columns at a time, so every
// row-wise quantity of the image-token head (online log-sum-exp over the 16384 codes, the
// target-logit gather, the logits row-sum metric) is thread-local: no shuffles, no shared memory.
// `chunk<FULL>`: FULL = all 32 columns are inside N -> branch- and predicate-free fast path.
#pragma once

#include "cfg_math.cuh"
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

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// bf16 pair: round two floats to bf16 (RN), return the packed word and the rounded values as floats
__device__ __forceinline__ uint32_t round_pair_bf16(float& a, float& b) {
  const uint32_t u = pack_bf16x2(a, b);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xFFFF0000u);
  return u;
}

// 32 consecutive fp32 values from a 16-byte aligned address (bias slices: same address across the warp)
__device__ __forceinline__ void load32_f32_vec(const float* __restrict__ src, float (&b)[32]) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(s4 + i);
    b[4 * i] = t.x;
    b[4 * i + 1] = t.y;
    b[4 * i + 2] = t.z;
    b[4 * i + 3] = t.w;
  }
}

template <bool FULL>
__device__ __forceinline__ void load_bias32(const float* __restrict__ bias, int col0, int valid, float (&b)[32]) {
  if (bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = 0.0f;
    return;
  }
  if (FULL && ((reinterpret_cast<uintptr_t>(bias + col0) & 15u) == 0)) {
    load32_f32_vec(bias + col0, b);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) b[j] = (j < valid) ? __ldg(bias + col0 + j) : 0.0f;
  }
}

__device__ __forceinline__ void store_packed16(__nv_bfloat16* dst, const uint32_t (&pk)[16]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

template <bool FULL>
__device__ __forceinline__ void store_row32_bf16(__nv_bfloat16* dst, const float (&v)[32], int valid) {
  if (FULL && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    store_packed16(dst, pk);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = __float2bfloat16_rn(v[j]);
  }
}

template <bool FULL>
__device__ __forceinline__ void store_row32_f32(float* dst, const float (&v)[32], int valid) {
  if (FULL && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = v[j];
  }
}

template <bool FULL>
__device__ __forceinline__ void load_row32_bf16(const __nv_bfloat16* src, float (&v)[32], int valid) {
  if (FULL && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = __ldg(s4 + i);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
        v[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u);
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
    RowMap rmap;        // where output row r lands (identity unless the output is row-segmented)
    float scale;        // out = scale * acc (+ bias); 0 means 1 (weight gradients carry 1 / world_size)
  };
  struct State {
    float rb;
    int64_t prow;
  };
  static constexpr int SMEM_BYTES = 0;

  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.rb = (ROW_BIAS && p.bias != nullptr && row < d.M) ? __ldg(p.bias + row) : 0.0f;
    st.prow = p.rmap(row);
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
    if (p.scale != 0.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.scale;
    }
    if constexpr (ROW_BIAS) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += st.rb;
    } else {
      if (p.bias != nullptr) {
        float b[32];
        load_bias32<FULL>(p.bias, col0, valid, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
      }
    }
    if constexpr (!TRANSPOSE) {
      OutT* dst = p.out + st.prow * p.ld + col0;
      if constexpr (sizeof(OutT) == 4) store_row32_f32<FULL>(reinterpret_cast<float*>(dst), v, valid);
      else store_row32_bf16<FULL>(reinterpret_cast<__nv_bfloat16*>(dst), v, valid);
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
// Split-K accumulation of a weight gradient: out += scale * acc with red.global.add (fire-and-forget, 16 bytes per
// instruction) into an output the launcher has zeroed.  With exactly two splits the sum is order-independent.
// ---------------------------------------------------------------------------
struct EpiRedAdd {
  struct Params {
    float* out;
    int64_t ld;
    float scale;  // 0 means 1
  };
  struct State {
    float* dst;
  };
  static constexpr int SMEM_BYTES = 0;
  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.dst = row < d.M ? p.out + static_cast<int64_t>(row) * p.ld : nullptr;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
    if (p.scale != 0.0f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.scale;
    }
    float* dst = st.dst + col0;
    if (FULL && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]),
                     "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) atomicAdd(dst + j, v[j]);
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// Weight-gradient store fused with the reduce-scatter of the data-parallel exchange: the rows of dW are
// partitioned over the ranks, and every rank's epilogue writes each tile straight into the OWNER's inbox over
// NVLink peer memory (plain posted stores: 128 contiguous bytes per thread and chunk) -- slot [source rank] of that
// inbox, so the owner later adds the N contributions in rank order (deterministic) without any rank having run a
// collective kernel.  The exchange rides inside the GEMM that produces the data; no SM is given up to it.
//   dest = inbox[row / rows_per_rank] + slot_off + (row % rows_per_rank) * ld + col
// ---------------------------------------------------------------------------
struct EpiStoreScatter {
  struct Params {
    DpScatter dp;
    int rows_per_rank;    // rows of this gradient matrix owned by each rank (a multiple of the tile's 128 / 256 rows)
    int64_t region_off;   // offset of this matrix's rows inside a shard (dW2 rows first, then dW1 rows, db2, db1)
    int64_t ld;
    float scale;          // 1 / world
  };
  struct State {
    float* dst;           // row base in the owner's inbox
  };
  static constexpr int SMEM_BYTES = 0;
  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.dst = nullptr;
    if (row < d.M) {
      const int owner = row / p.rows_per_rank;
      st.dst = p.dp.inbox[owner] + static_cast<int64_t>(p.dp.rank) * p.dp.shard_elems + p.region_off +
               static_cast<int64_t>(row - owner * p.rows_per_rank) * p.ld;
    }
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.scale;
    float* dst = st.dst + col0;
    if (FULL && ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0)) {
      // 256-bit stores: a thread's 128 bytes leave as four 32-byte sectors.  Peer stores are not merged in the local
      // L2, every store instruction becomes NVLink write packets of its own -- with 16-byte stores the dW2 GEMM of an
      // 8-rank job ran 8.5 instead of 6.7 ms (packet rate, not bytes: 235 MB per GEMM)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * i), "f"(v[8 * i]),
                     "f"(v[8 * i + 1]), "f"(v[8 * i + 2]), "f"(v[8 * i + 3]), "f"(v[8 * i + 4]), "f"(v[8 * i + 5]),
                     "f"(v[8 * i + 6]), "f"(v[8 * i + 7])
                     : "memory");
    } else {
      store_row32_f32<FULL>(dst, v, valid);
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// Split-K partial store for the swap-AB decode GEMMs: part[k_split][col][row] = acc (fp32, transposed so
// that the 32 lanes of a warp write 32 consecutive floats).  The partials are summed in a fixed order by
// the finalize kernel, so the result is deterministic (no atomics).
// ---------------------------------------------------------------------------
struct EpiPartialStoreT {
  struct Params {
    float* part;        // [k_splits, n, ld]
    int64_t ld;         // elements per sample row (= M)
    int64_t split_stride;
  };
  struct State {
    float* base;
  };
  static constexpr int SMEM_BYTES = 0;
  __device__ static void begin(const Params& p, State& st, int row, int, int ks, const GemmDims&, uint8_t*) {
    st.base = p.part + static_cast<int64_t>(ks) * p.split_stride + row;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (FULL || j < valid) st.base[static_cast<int64_t>(col0 + j) * p.ld] = v[j];
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// Swap-AB linear layer with cluster split-K (decode GEMM1, gen_aligner): the partial tile goes to this CTA's shared
// memory as part[col][row]; the CTAs of the cluster then sum the k_splits partials in split order through distributed
// shared memory, add the bias, optionally apply GELU and write out[n][m] (bf16).  Deterministic, no HBM partials, no
// finalize launch.
// ---------------------------------------------------------------------------
template <bool GELU>
struct EpiClusterLinearT {
  struct Params {
    const float* bias;     // b1 [E]
    __nv_bfloat16* act;    // [n, E]
    int64_t ld;            // E
  };
  struct State {};
  static constexpr int SMEM_BYTES = 0;  // uses the drained operand ring: 32 x 128 fp32 = 16 KB
  __device__ static void begin(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
  template <bool FULL>
  __device__ static void chunk(const Params&, State&, int, int col0, float (&v)[32], const GemmDims&, uint8_t* smem) {
    float* part = reinterpret_cast<float*>(smem);
    const int r = ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31);  // row inside the 128-row tile
    const int c0 = col0 & 31;  // BN == 32: one chunk per tile
#pragma unroll
    for (int j = 0; j < 32; ++j) part[(c0 + j) * 128 + r] = v[j];
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
  // Every CTA of the cluster reduces its own slice of the 32 columns (columns [rank*8, rank*8+8) for 4 splits), so the
  // distributed-shared-memory traffic is spread over all the SMs of the cluster and all remote loads of a thread are
  // in flight together.  The partials are always added in split order 0, 1, 2, ... (deterministic).
  __device__ static void cluster_finalize(const Params& p, int row, int n0, int k_splits, int rank, const GemmDims& d,
                                          uint8_t* smem) {
    if (row >= d.M) return;
    const float* part = reinterpret_cast<const float*>(smem);
    const int r = ((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31);
    const int ncols = min(32, d.N - n0);
    const int per = (32 + k_splits - 1) / k_splits;   // columns per CTA (<= 32)
    const int c_lo = rank * per;
    const float b = __ldg(p.bias + row);
    const uint32_t local = smem_u32(part + r);
    float v[8][8];  // [split][column]; k_splits <= 8, per <= 8 whenever k_splits >= 4
    if (per <= 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < k_splits) {
          const uint32_t src = mapa_shared(local, static_cast<uint32_t>(k));
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[k][j] = (j < per && c_lo + j < 32) ? ld_dsmem_f32(src + static_cast<uint32_t>((c_lo + j) * 128 * 4)) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < per && c_lo + j < ncols) {
          float acc = v[0][j];
#pragma unroll
          for (int k = 1; k < 8; ++k)
            if (k < k_splits) acc += v[k][j];
          const float x = bf16_round(acc + b);
          p.act[static_cast<int64_t>(n0 + c_lo + j) * p.ld + row] = __float2bfloat16_rn(GELU ? gelu_erf(x) : x);
        }
      }
    } else {
      // 1 or 2 splits: a CTA owns 32 or 16 columns
      for (int j = 0; j < per; ++j) {
        const int c = c_lo + j;
        if (c >= ncols) break;
        float acc = part[c * 128 + r];
        for (int k = 1; k < k_splits; ++k)
          acc += ld_dsmem_f32(mapa_shared(local, static_cast<uint32_t>(k)) + static_cast<uint32_t>(c * 128 * 4));
        if (rank != 0 && k_splits > 1) {
          // rank 0's partial is remote for the other ranks: rebuild the sum in split order
          acc = ld_dsmem_f32(mapa_shared(local, 0u) + static_cast<uint32_t>(c * 128 * 4));
          for (int k = 1; k < k_splits; ++k)
            acc += ld_dsmem_f32(mapa_shared(local, static_cast<uint32_t>(k)) + static_cast<uint32_t>(c * 128 * 4));
        }
        const float x = bf16_round(acc + b);
        p.act[static_cast<int64_t>(n0 + c) * p.ld + row] = __float2bfloat16_rn(GELU ? gelu_erf(x) : x);
      }
    }
  }
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

  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.rb = (TRANSPOSE && row < d.M) ? __ldg(p.bias + row) : 0.0f;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
    if constexpr (!TRANSPOSE) {
      float b[32];
      load_bias32<FULL>(p.bias, col0, valid, b);
      const int64_t off = static_cast<int64_t>(row) * p.ld + col0;
      if (FULL && ((reinterpret_cast<uintptr_t>(p.act + off) & 15u) == 0)) {
        uint32_t pp[16], pa[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0 = v[2 * i] + b[2 * i], x1 = v[2 * i + 1] + b[2 * i + 1];
          pp[i] = round_pair_bf16(x0, x1);
          pa[i] = pack_bf16x2(gelu_erf(x0), gelu_erf(x1));
        }
        if constexpr (STORE_PRE) store_packed16(p.pre + off, pp);
        store_packed16(p.act + off, pa);
      } else {
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = bf16_round(v[j] + b[j]);
          a[j] = gelu_erf(v[j]);
        }
        if constexpr (STORE_PRE) store_row32_bf16<false>(p.pre + off, v, valid);
        store_row32_bf16<false>(p.act + off, a, valid);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < valid) {
          const float x = bf16_round(v[j] + st.rb);
          const int64_t off = static_cast<int64_t>(col0 + j) * p.ld + row;
          if constexpr (STORE_PRE) p.pre[off] = __float2bfloat16_rn(x);
          p.act[off] = __float2bfloat16_rn(gelu_erf(x));
        }
      }
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// GEMM2 forward epilogue with the softmax-minus-onehot producer fused in (SURVEY §2 K2 + K4).
//   logit l = bf16(acc + b2)                       vision_head Linear, modeling_vlm.py:50
//   e      = exp(l - ref_row)                      the softmax numerator of train.py:391's log_softmax
// Per (row, column sub-tile) it leaves (max l, sum e) partials, the gathered target logit and the partial
// row-sum of the logits; e is (optionally) spilled ONCE as bf16 -- it is the A operand of both backward GEMMs:
// softmax = e * exp(ref - lse) is a per-row scale that commutes with the contraction, so no separate
// softmax-minus-onehot pass ever runs over the [rows, V] tensor (the one-hot term is one element per row,
// written by target_fixup_kernel).  The log-sum-exp is taken over the bf16-rounded logits in fp32, which is what
// the bf16 reference path computes (Linear output bf16, log_softmax autocast to fp32).
// ref_row: 0 in the first pass.  bf16 / fp32 share their exponent range, so e is representable (and the fp32
// accumulations downstream stay in range) while max_row - ref_row lies in [-50, 60]; lse_finalize_kernel checks
// that per row and flags the 128/256-row blocks that violate it, and a second (normally empty) launch of this
// GEMM with blk_mask set recomputes just those blocks against ref_row = max_row.
// ---------------------------------------------------------------------------
struct EpiLogitsExp {
  struct Params {
    const float* bias;          // [V]
    __nv_bfloat16* espill;      // [rows, V] bf16 spill of e (may be null: forward only)
    int64_t ld;
    const int64_t* labels;      // [rows] target code per row
    float2* part;               // [num_sub_tiles, rows] (max logit, sum of e) partials
    float* rowsum_part;         // [num_sub_tiles, rows] partial sums of logits (may be null)
    float* tgt;                 // [rows] gathered target logit
    const float* row_ref;       // [rows] exponent reference (null = 0: first pass)
    const uint8_t* blk_mask;    // [num M-blocks] repair pass: only the flagged blocks are recomputed (null = all)
    const uint8_t* any_flag;    // repair pass: *any_flag == 0 means no block is flagged (the launch returns at once)
  };
  struct State {
    float m, s, sum, tgt, nref;
    int label;
    bool hit;
  };
  static constexpr int SMEM_BYTES = 0;
  static constexpr bool HAS_TILE_MASK = true;
  static constexpr float LOG2E = 1.4426950408889634f;

  __device__ static bool tile_enabled(const Params& p, int m_blk) {
    return p.blk_mask == nullptr || p.blk_mask[m_blk] != 0;
  }
  __device__ static bool launch_enabled(const Params& p) { return p.any_flag == nullptr || *p.any_flag != 0; }
  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.m = -INFINITY;
    st.s = 0.0f;
    st.sum = 0.0f;
    st.tgt = 0.0f;
    st.hit = false;
    st.label = (row < d.M) ? static_cast<int>(__ldg(p.labels + row)) : -1;
    st.nref = (p.row_ref != nullptr && row < d.M) ? -__ldg(p.row_ref + row) * LOG2E : 0.0f;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
    float b[32];
    load_bias32<FULL>(p.bias, col0, valid, b);
    float cmax = -INFINITY;
    float acc0 = 0.0f, acc1 = 0.0f, ls0 = 0.0f, ls1 = 0.0f;
    if constexpr (FULL) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[2 * i] += b[2 * i];
        v[2 * i + 1] += b[2 * i + 1];
        round_pair_bf16(v[2 * i], v[2 * i + 1]);      // the logits exactly as the bf16 Linear emits them
        cmax = fmaxf(cmax, fmaxf(v[2 * i], v[2 * i + 1]));
        const float e0 = ex2_approx(fmaf(v[2 * i], LOG2E, st.nref));
        const float e1 = ex2_approx(fmaf(v[2 * i + 1], LOG2E, st.nref));
        acc0 += e0;
        acc1 += e1;
        ls0 += v[2 * i];
        ls1 += v[2 * i + 1];
        pk[i] = pack_bf16x2(e0, e1);
      }
      if (p.espill != nullptr) {
        __nv_bfloat16* dst = p.espill + static_cast<int64_t>(row) * p.ld + col0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
          store_packed16(dst, pk);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) reinterpret_cast<uint32_t*>(dst)[i] = pk[i];  // rows are 4-byte aligned (V % 8 == 0)
        }
      }
    } else {
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = bf16_round(v[j] + b[j]);
        e[j] = 0.0f;
        if (j < valid) {
          cmax = fmaxf(cmax, v[j]);
          e[j] = ex2_approx(fmaf(v[j], LOG2E, st.nref));
          acc0 += e[j];
          ls0 += v[j];
        }
      }
      if (p.espill != nullptr) store_row32_bf16<false>(p.espill + static_cast<int64_t>(row) * p.ld + col0, e, valid);
    }
    st.m = fmaxf(st.m, cmax);
    st.s += acc0 + acc1;
    st.sum += ls0 + ls1;
    // (in a partial N tile `valid` counts the columns up to the matrix edge, which can be more than this chunk's 32: a
    //  label in a LATER chunk must not count as a hit here -- with the label in the other column half of the tile, two
    //  warps would both store tgt[row], one of them the 0 it started with)
    const int rel = st.label - col0;
    if (rel >= 0 && rel < (valid < 32 ? valid : 32)) {
      st.hit = true;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == rel) st.tgt = v[j];
    }
  }
  __device__ static void end(const Params& p, State& st, int row, int, int sub_tile, const GemmDims& d, uint8_t*) {
    if (row >= d.M) return;
    const int64_t idx = static_cast<int64_t>(sub_tile) * d.M + row;
    p.part[idx] = make_float2(st.m, st.s);
    if (p.rowsum_part != nullptr) p.rowsum_part[idx] = st.sum;
    if (st.hit) p.tgt[row] = st.tgt;  // only the sub-tile that holds the label column writes (lse_finalize ignores
                                      // tgt for labels outside [0, V))
  }
};

// ---------------------------------------------------------------------------
// dAct GEMM epilogue.  The A operand is the forward's unscaled spill g = e - onehot * exp(lse - ref), so
//   dact = w_row * acc                    w_row = -c_row * exp(ref_row - lse_row), c_row = d loss / d logp_row
//   dpre = bf16(bf16(dact) * gelu'(pre))  autograd of GELU, SURVEY §8 a-6
// and, for the weight-gradient GEMM that follows, the row-scaled activations
//   act_w = bf16(w_row * act),  act = bf16(gelu(pre)) recomputed from the pre-activation already in registers
// (dW2 = dlogits^T act = g^T act_w: the row scale rides on the other operand).
// ---------------------------------------------------------------------------
struct EpiDactScale {
  struct Params {
    const __nv_bfloat16* pre;  // [rows, E]
    __nv_bfloat16* dpre;       // [rows, E]
    int64_t ld;
    const float* row_w;        // [rows]
    __nv_bfloat16* act_w;      // [rows, E] out (null: head frozen, no weight gradient)
  };
  struct State {
    float w;
  };
  static constexpr int SMEM_BYTES = 0;
  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.w = (row < d.M) ? __ldg(p.row_w + row) : 0.0f;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t*) {
    if (row >= d.M) return;
    const int valid = FULL ? 32 : d.N - col0;
    if (valid <= 0) return;
    const int64_t off = static_cast<int64_t>(row) * p.ld + col0;
    float pre[32];
    load_row32_bf16<FULL>(p.pre + off, pre, valid);
    if (p.act_w != nullptr) {
      float aw[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = pre[j];
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
        const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
        aw[j] = st.w * bf16_round(gelu_erf(x));
        v[j] = bf16_round(st.w * v[j]) * (cdf + x * pdf);
      }
      store_row32_bf16<FULL>(p.act_w + off, aw, valid);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = bf16_round(st.w * v[j]) * gelu_erf_grad(pre[j]);
    }
    store_row32_bf16<FULL>(p.dpre + off, v, valid);
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

// ---------------------------------------------------------------------------
// Decode GEMM2 (swap-AB) epilogue with the CFG tail fused in.  Accumulator row = code v, the 32 columns are the
// CFG sample rows (2k = conditional, 2k+1 = unconditional).  While the tile is in registers:
//   logit = bf16(acc + b2[v])                         (optionally dumped as logits[row][v])
//   t_k   = merge(logit_2k, logit_2k+1)               image_generation.py:157-161
//   tile exponent K_tile (max over the CTA's 128 codes), weights u = P(r) 2^(n - K_tile) -> wbuf[k][v]
//   segment sums (a warp is exactly one 32-code segment: shfl_xor butterfly = the oracle's pairwise tree)
// so the merged logits never exist in HBM; cfg_finish_kernel completes the draw from 512 sums per pair.
// Needs TILE_M == 128 (one CTA = one tile) and BN == 32 (one chunk = all columns).
// ---------------------------------------------------------------------------
// GREEDY is a template parameter: the arg-max variant is a different (large, fully unrolled) code path, and keeping it
// out of the sampling instantiation takes a quarter off the decode kernel's code size.
template <int MODE, bool TDIV, bool WBF = false, bool GREEDY = false>
struct EpiCfgFused {
  struct Params {
    const float* bias;            // b2 [V]
    float cfg_weight, temperature;
    __nv_bfloat16* logits_dump;   // [2P, V] or null
    int64_t ld;                   // V
    CfgFusedBuffers buf;
    int greedy;
    int vocab;
  };
  struct State {
    float rb;
  };
  static constexpr int SMEM_BYTES = 1024;  // [4 warps][16 pairs] x {float, float, int}

  __device__ static void begin(const Params& p, State& st, int row, int, int, const GemmDims& d, uint8_t*) {
    st.rb = (row < d.M) ? __ldg(p.bias + row) : 0.0f;
  }
  template <bool FULL>
  __device__ static void chunk(const Params& p, State& st, int row, int col0, float (&v)[32], const GemmDims& d,
                               uint8_t* smem) {
    const int lane = threadIdx.x & 31;
    const int q = (threadIdx.x >> 5) & 3;
    const int valid = FULL ? 32 : min(32, d.N - col0);
    const int npairs = valid >> 1;
    const int pair0 = col0 >> 1;
    const int tile = row / SAMPLE_TILE;
    const int ntile = p.vocab / SAMPLE_TILE;
    float* sm_k = reinterpret_cast<float*>(smem);  // [4][16]
    float* sm_v = sm_k + 64;                       // [4][16] greedy value
    int* sm_i = reinterpret_cast<int*>(sm_v + 64);  // [4][16] greedy index
    const int tr = (d.trace_id > 0 && threadIdx.x == 64) ? 5 : 0;  // epilogue-internal timeline (trace row 5)
    trace_stamp(tr, 0);
    // logits exactly as the reference's bf16 Linear output: pr[k] = (cond, uncond) of pair k as packed bf16
    uint32_t pr[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) pr[k] = pack_bf16x2(v[2 * k] + st.rb, v[2 * k + 1] + st.rb);
    if (p.logits_dump != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < valid) {
          const uint16_t h16 = static_cast<uint16_t>((j & 1) ? (pr[j >> 1] >> 16) : (pr[j >> 1] & 0xFFFFu));
          reinterpret_cast<uint16_t*>(p.logits_dump)[static_cast<int64_t>(col0 + j) * p.ld + row] = h16;
        }
      }
    }
    float t[16];
    if constexpr (WBF) {
      // bf16-exact cfg_weight: the op-by-op bf16 merge runs on the bf16x2 pipe, two pairs per instruction
      const uint32_t w2 = pack_bf16x2(p.cfg_weight, p.cfg_weight);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const uint32_t c2 = __byte_perm(pr[k], pr[k + 1], 0x5410);  // (cond k, cond k+1)
        const uint32_t u2 = __byte_perm(pr[k], pr[k + 1], 0x7632);  // (uncond k, uncond k+1)
        cfg_merge2_hw<TDIV>(c2, u2, w2, p.temperature, t[k], t[k + 1]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; k += 2)
        cfg_merge_vals<MODE, TDIV>(__uint_as_float(pr[k] << 16), __uint_as_float(pr[k + 1] << 16),
                                   __uint_as_float(pr[k] & 0xFFFF0000u), __uint_as_float(pr[k + 1] & 0xFFFF0000u),
                                   p.cfg_weight, p.temperature, t[k], t[k + 1]);
    }
    trace_stamp(tr, 1);
    if constexpr (GREEDY) {
      // per-pair arg-max over the tile: (value, code) with the lowest code on ties
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float bv = t[k];
        int bi = row;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        if (lane == k) {
          sm_v[q * 16 + k] = bv;
          sm_i[q * 16 + k] = bi;
        }
      }
      named_bar_sync(1, 128);
      if (q == 0 && lane < npairs) {
        float bv = sm_v[lane];
        int bi = sm_i[lane];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float ov = sm_v[w * 16 + lane];
          const int oi = sm_i[w * 16 + lane];
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        p.buf.tile_max[static_cast<int64_t>(pair0 + lane) * ntile + tile] = bv;
        p.buf.tile_arg[static_cast<int64_t>(pair0 + lane) * ntile + tile] = bi;
      }
      named_bar_sync(1, 128);
      return;
    } else {
    // tile exponent per pair: 16-way transpose-reduce (max is exact, any order)
    const int mypair = transpose_reduce_pair_of_lane(lane);
    {
      // n(t) is non-decreasing in t, so the tile's largest exponent is n(largest t): reduce t, convert once
      const float tw = warp_transpose_reduce16(t, lane, OpMax());
      if ((lane & 1) == 0) sm_k[q * 16 + mypair] = exp_n_only(tw);
    }
    trace_stamp(tr, 2);
    named_bar_sync(1, 128);
    trace_stamp(tr, 3);
    float u[16];
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const float kt0 = fmaxf(fmaxf(sm_k[k], sm_k[16 + k]), fmaxf(sm_k[32 + k], sm_k[48 + k]));
      const float kt1 = fmaxf(fmaxf(sm_k[k + 1], sm_k[17 + k]), fmaxf(sm_k[33 + k], sm_k[49 + k]));
      exp_weight2(t[k], t[k + 1], exp_koff(kt0), exp_koff(kt1), u[k], u[k + 1]);
      if (k < npairs) p.buf.wbuf[static_cast<int64_t>(pair0 + k) * p.vocab + row] = u[k];
      if (k + 1 < npairs) p.buf.wbuf[static_cast<int64_t>(pair0 + k + 1) * p.vocab + row] = u[k + 1];
    }
    // segment sums: the warp is one 32-code segment; the transpose-reduce adds lanes in the oracle's butterfly
    // order (strides 16, 8, 4, 2, 1)
    trace_stamp(tr, 4);
    const float ssum = warp_transpose_reduce16(u, lane, OpSum());
    if ((lane & 1) == 0 && mypair < npairs) {
      p.buf.seg_sum[static_cast<int64_t>(pair0 + mypair) * (p.vocab / SAMPLE_SEG) + row / SAMPLE_SEG] = ssum;
      if (q == 0)
        p.buf.tile_k[static_cast<int64_t>(pair0 + mypair) * ntile + tile] =
            fmaxf(fmaxf(sm_k[mypair], sm_k[16 + mypair]), fmaxf(sm_k[32 + mypair], sm_k[48 + mypair]));
    }
    trace_stamp(tr, 5);
    named_bar_sync(1, 128);  // smem is reused by the next tile of a persistent CTA
    trace_stamp(tr, 6);
    }
  }
  __device__ static void end(const Params&, State&, int, int, int, const GemmDims&, uint8_t*) {}
};

}  // namespace ospo
