// Non-GEMM kernels of the image-token head path: log-sum-exp merge (+ exponent-window check), the one-hot fix-up of
// the forward's softmax-numerator spill, per-sequence log-prob reduction, the SimPO scalar stage, the per-row weights
// of the backward GEMM pair, fixed-order column sums for the bias gradients, the peer-memory gradient exchange, and
// the CFG merge + inverse-CDF sampler.
// All are HBM- or latency-bound; grids are sized from the data, loads are 16-byte vectors.
#pragma once

#include "cfg_math.cuh"
#include "launchers.h"
#include "ptx.cuh"

namespace ospo {

// ---------------------------------------------------------------------------
// Merge the per-(N-tile,row) partials written by EpiLogitsExp into per-row lse and log-prob.
// reference: logits.log_softmax(-1) gathered at labels, ospo/wrapper/train.py:391
//   lse_row = ref_row + log(sum of the partial sums of e = exp(l - ref_row))       (fixed order: deterministic)
// pass 1 (ref = 0 everywhere): also records the row maximum and flags, per GEMM M-block of `tile_m` rows, the
//   blocks in which some row's maximum lies outside the window where e is safely representable ([-50, 60]; a NaN
//   fails the test too).  pass 2 touches only the flagged blocks, whose partials the repair launch of GEMM2 has
//   rewritten against ref_row = max_row.
// A label outside [0, V) (ignore_index inside a promised image span, train.py:387-389) yields log-prob 0 and is
// left out of every count, exactly like a masked position.
// ---------------------------------------------------------------------------
constexpr float E_WINDOW_HI = 60.0f, E_WINDOW_LO = -50.0f;

__global__ void lse_finalize_kernel(const float2* __restrict__ part, const float* __restrict__ rowsum_part,
                                    const float* __restrict__ tgt, const int64_t* __restrict__ labels, int vocab,
                                    int rows, int num_n, int tile_m, int pass, uint8_t* __restrict__ blk_mask,
                                    uint8_t* __restrict__ any_flag, float* __restrict__ row_ref, float* __restrict__ row_max,
                                    float* __restrict__ row_lse, float* __restrict__ row_logp,
                                    float* __restrict__ row_logit_sum) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (pass == 2 && (*any_flag == 0 || blk_mask[r / tile_m] == 0)) return;
  float m = -INFINITY, s = 0.0f, ls = 0.0f;
  int nb = 0;
  for (; nb + 8 <= num_n; nb += 8) {   // eight independent loads in flight, added in the fixed order nb = 0, 1, ...
    float2 p[8];
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      p[i] = part[static_cast<int64_t>(nb + i) * rows + r];
      q[i] = (rowsum_part != nullptr) ? rowsum_part[static_cast<int64_t>(nb + i) * rows + r] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      m = fmaxf(m, p[i].x);
      s += p[i].y;
      ls += q[i];
    }
  }
  for (; nb < num_n; ++nb) {
    const float2 p = part[static_cast<int64_t>(nb) * rows + r];
    m = fmaxf(m, p.x);
    s += p.y;
    if (rowsum_part != nullptr) ls += rowsum_part[static_cast<int64_t>(nb) * rows + r];
  }
  float ref = 0.0f;
  if (pass == 1) {
    row_max[r] = m;
    if (!(m <= E_WINDOW_HI && m >= E_WINDOW_LO)) {   // benign races: every writer stores 1
      blk_mask[r / tile_m] = 1;
      *any_flag = 1;
    }
  } else {
    ref = row_max[r];
  }
  if (row_ref != nullptr) row_ref[r] = ref;
  const float lse = ref + logf(s);
  const int64_t lab = labels[r];
  row_lse[r] = lse;
  row_logp[r] = (lab >= 0 && lab < vocab) ? tgt[r] - lse : 0.0f;
  if (row_logit_sum != nullptr) row_logit_sum[r] = ls;
}

// The one-hot term of softmax-minus-onehot: one element per row of the forward's spill.
//   g[r, label_r] = (p_target - 1) * exp(lse_r - ref_r),   p_target - 1 = expm1(logp_r)  (exact as p_target -> 1)
// so that  dlogits[r, :] = c_r (onehot - softmax) = -c_r exp(ref_r - lse_r) * g[r, :]  for the whole row.
__global__ void target_fixup_kernel(__nv_bfloat16* __restrict__ espill, int64_t ld, const int64_t* __restrict__ labels,
                                    int vocab, int rows, const float* __restrict__ row_logp,
                                    const float* __restrict__ row_lse, const float* __restrict__ row_ref) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int64_t lab = labels[r];
  if (lab < 0 || lab >= vocab) return;
  espill[static_cast<int64_t>(r) * ld + lab] = __float2bfloat16_rn(expm1f(row_logp[r]) * expf(row_lse[r] - row_ref[r]));
}

template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
  // fixed-shape tree: deterministic for a given THREADS
  red[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const float out = red[0];
  __syncthreads();
  return out;
}

// One block per sequence: (per_token_logps * mask).sum(-1) [/ mask.sum(-1)]  (train.py:393-396).
// Rows of sequence s are [seq_off[s], seq_off[s+1]); a row whose label lies outside [0, V) counts as masked
// (its log-prob is already 0), so mask.sum(-1) = seq_count[s] is the number of valid labels.
__global__ void seq_reduce_kernel(const float* __restrict__ row_logp, const float* __restrict__ row_logit_sum,
                                  const int64_t* __restrict__ seq_off, const int64_t* __restrict__ labels, int vocab,
                                  int average, float* __restrict__ seq_logps, float* __restrict__ seq_sum,
                                  float* __restrict__ seq_logit_sum, float* __restrict__ seq_count) {
  __shared__ float red[256];
  const int s = blockIdx.x;
  const int64_t lo = seq_off[s], hi = seq_off[s + 1];
  float a = 0.0f, b = 0.0f, n = 0.0f;
  for (int64_t r = lo + threadIdx.x; r < hi; r += 256) {
    const int64_t lab = labels[r];
    if (lab >= 0 && lab < vocab) {
      n += 1.0f;
      a += row_logp[r];
      if (row_logit_sum != nullptr) b += row_logit_sum[r];
    }
  }
  a = block_sum<256>(a, red);
  b = block_sum<256>(b, red);
  n = block_sum<256>(n, red);
  if (threadIdx.x == 0) {
    seq_sum[s] = a;
    seq_logps[s] = average ? a / n : a;   // 0 / 0 = NaN for a fully masked sequence, as in the reference
    if (seq_logit_sum != nullptr) seq_logit_sum[s] = b;
    seq_count[s] = n;
  }
}

// ---------------------------------------------------------------------------
// SimPO scalar stage (single block).  reference: simpo_loss ospo/wrapper/train.py:317-342,
// losses.mean() :419, optional SFT term :421-430, metrics :432-443; gradient per SURVEY §8 a-6.
// Sequences [0,B) are chosen, [B,2B) rejected (concatenated_forward, train.py:364-365).
// ---------------------------------------------------------------------------
struct SimpoHyper {
  float beta, gamma_beta_ratio, label_smoothing, sft_weight;
  int loss_type;  // 0 = sigmoid, 1 = hinge
  int vocab;
};

enum SimpoScalar {
  SC_LOSS = 0,
  SC_SIMPO_LOSS = 1,
  SC_SFT_LOSS = 2,
  SC_REWARD_CHOSEN = 3,
  SC_REWARD_REJECTED = 4,
  SC_REWARD_ACC = 5,
  SC_REWARD_MARGIN = 6,
  SC_LOGPS_CHOSEN = 7,
  SC_LOGPS_REJECTED = 8,
  SC_LOGITS_CHOSEN = 9,
  SC_LOGITS_REJECTED = 10,
  SC_SFT_ROW_COEF = 11,
  SC_COUNT = 16
};

__device__ __forceinline__ float log_sigmoid(float x) {
  // stable: min(x,0) - log1p(exp(-|x|))
  return fminf(x, 0.0f) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void simpo_scalar_kernel(const float* __restrict__ seq_logps, const float* __restrict__ seq_sum,
                                    const float* __restrict__ seq_logit_sum, const float* __restrict__ seq_count,
                                    int B, SimpoHyper hp, float* __restrict__ losses,
                                    float* __restrict__ chosen_rewards, float* __restrict__ rejected_rewards,
                                    float* __restrict__ grad_seq, float* __restrict__ scalars) {
  __shared__ float red[256];
  float loss_acc = 0.f, rc_acc = 0.f, rr_acc = 0.f, acc_acc = 0.f, lc_acc = 0.f, lr_acc = 0.f;
  float csum_acc = 0.f, ccount_acc = 0.f, rcount_acc = 0.f, clog_acc = 0.f, rlog_acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float c = seq_logps[b], r = seq_logps[B + b];
    const float z = (c - r) - hp.gamma_beta_ratio;
    const float bz = hp.beta * z;
    float loss, dz;  // dz = dloss_b / dz
    if (hp.loss_type == 0) {
      loss = -log_sigmoid(bz) * (1.0f - hp.label_smoothing) - log_sigmoid(-bz) * hp.label_smoothing;
      dz = -hp.beta * ((1.0f - hp.label_smoothing) * sigmoidf(-bz) - hp.label_smoothing * sigmoidf(bz));
    } else {
      const float h = 1.0f - bz;
      loss = fmaxf(h, 0.0f);
      dz = (h > 0.0f) ? -hp.beta : 0.0f;
    }
    losses[b] = loss;
    const float rc = hp.beta * c, rr = hp.beta * r;
    chosen_rewards[b] = rc;
    rejected_rewards[b] = rr;
    grad_seq[b] = dz / static_cast<float>(B);        // d mean(losses) / d chosen_logps[b]
    grad_seq[B + b] = -dz / static_cast<float>(B);   // d mean(losses) / d rejected_logps[b]
    loss_acc += loss;
    rc_acc += rc;
    rr_acc += rr;
    acc_acc += (rc > rr) ? 1.0f : 0.0f;
    lc_acc += c;
    lr_acc += r;
    csum_acc += seq_sum[b];
    ccount_acc += seq_count[b];
    rcount_acc += seq_count[B + b];
    if (seq_logit_sum != nullptr) {
      clog_acc += seq_logit_sum[b];
      rlog_acc += seq_logit_sum[B + b];
    }
  }
  const float loss_sum = block_sum<256>(loss_acc, red);
  const float rc_sum = block_sum<256>(rc_acc, red);
  const float rr_sum = block_sum<256>(rr_acc, red);
  const float acc_sum = block_sum<256>(acc_acc, red);
  const float lc_sum = block_sum<256>(lc_acc, red);
  const float lr_sum = block_sum<256>(lr_acc, red);
  const float csum = block_sum<256>(csum_acc, red);
  const float ccount = block_sum<256>(ccount_acc, red);
  const float rcount = block_sum<256>(rcount_acc, red);
  const float clog = block_sum<256>(clog_acc, red);
  const float rlog = block_sum<256>(rlog_acc, red);
  if (threadIdx.x == 0) {
    const float fB = static_cast<float>(B);
    const float simpo = loss_sum / fB;
    float sft = 0.0f, sft_coef = 0.0f;
    if (hp.sft_weight > 0.0f && ccount > 0.0f) {
      sft = -csum / ccount;  // CrossEntropyLoss(mean) over the unmasked chosen rows
      sft_coef = -hp.sft_weight / ccount;
    }
    scalars[SC_SIMPO_LOSS] = simpo;
    scalars[SC_SFT_LOSS] = sft;
    scalars[SC_LOSS] = hp.sft_weight * sft + simpo;
    scalars[SC_REWARD_CHOSEN] = rc_sum / fB;
    scalars[SC_REWARD_REJECTED] = rr_sum / fB;
    scalars[SC_REWARD_ACC] = acc_sum / fB;
    scalars[SC_REWARD_MARGIN] = (rc_sum - rr_sum) / fB;
    scalars[SC_LOGPS_CHOSEN] = lc_sum / fB;
    scalars[SC_LOGPS_REJECTED] = lr_sum / fB;
    scalars[SC_LOGITS_CHOSEN] = ccount > 0.f ? clog / (ccount * static_cast<float>(hp.vocab)) : 0.f;
    scalars[SC_LOGITS_REJECTED] = rcount > 0.f ? rlog / (rcount * static_cast<float>(hp.vocab)) : 0.f;
    scalars[SC_SFT_ROW_COEF] = sft_coef;
  }
}

// per-row weight of the backward GEMM pair.  With the forward's spill g (EpiLogitsExp + target_fixup_kernel)
//   dlogits[r, :] = c_r (onehot - softmax) = w_r * g[r, :],      w_r = -c_r * exp(ref_r - lse_r)
//   c_r = grad_scale * ( grad_seq[s] * (average ? 1/n_s : 1)  +  [s < num_sft_seqs] * sft_coef )
// n_s counts the valid labels of sequence s; a row with an invalid label carries no gradient (w_r = 0).
__global__ void row_weight_kernel(const float* __restrict__ grad_seq, const int64_t* __restrict__ seq_off,
                                  const int64_t* __restrict__ labels, int vocab, int average,
                                  const float* __restrict__ grad_scale, const float* __restrict__ sft_coef,
                                  int num_sft_seqs, const float* __restrict__ row_lse, const float* __restrict__ row_ref,
                                  float* __restrict__ row_w) {
  __shared__ float red[128];
  const int s = blockIdx.x;
  const int64_t lo = seq_off[s], hi = seq_off[s + 1];
  float n = 0.0f;
  for (int64_t r = lo + threadIdx.x; r < hi; r += 128) {
    const int64_t lab = labels[r];
    n += (lab >= 0 && lab < vocab) ? 1.0f : 0.0f;
  }
  n = block_sum<128>(n, red);
  const float gs = (grad_scale != nullptr) ? *grad_scale : 1.0f;
  float c = grad_seq[s];
  if (average) c /= n;
  if (sft_coef != nullptr && s < num_sft_seqs) c += *sft_coef;
  c *= gs;
  for (int64_t r = lo + threadIdx.x; r < hi; r += 128) {
    const int64_t lab = labels[r];
    row_w[r] = (lab >= 0 && lab < vocab) ? -c * expf(row_ref[r] - row_lse[r]) : 0.0f;
  }
}

// ---------------------------------------------------------------------------
// Bias gradients: (row-weighted) column sums of a bf16 [rows, cols] matrix, in two fixed-order stages so the result
// is bit-reproducible (no atomics):
//   db2[v] = scale * sum_r w_r g[r, v]      db1[e] = scale * sum_r dpre[r, e]
// stage 1: block = 128 threads x 8 columns (16-byte loads) over CS_ROWS_PER_BLOCK rows -> partial[row block][col];
// stage 2: one thread per column adds the row-block partials in order.  HBM-bound: the matrix is read once.
// ---------------------------------------------------------------------------
constexpr int CS_ROWS_PER_BLOCK = 512;
constexpr int CS_UNROLL = 4;  // rows in flight per thread: one 16-byte load each (memory-level parallelism)

template <bool WEIGHTED>
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int rows, int cols,
                      const float* __restrict__ row_w, float* __restrict__ partial) {
  const int col = (blockIdx.x * 128 + threadIdx.x) * 8;
  if (col >= cols) return;
  const int r0 = blockIdx.y * CS_ROWS_PER_BLOCK;
  const int r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int rb = r0; rb < r1; rb += CS_UNROLL) {
    uint4 u[CS_UNROLL];
    float w[CS_UNROLL];
#pragma unroll
    for (int i = 0; i < CS_UNROLL; ++i) {
      const int r = min(rb + i, r1 - 1);  // clamped rows are loaded but carry weight 0
      u[i] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ld + col));
      w[i] = (rb + i < r1) ? (WEIGHTED ? __ldg(row_w + r) : 1.0f) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < CS_UNROLL; ++i) {
      const uint32_t q[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cs[2 * k] = fmaf(w[i], __uint_as_float(q[k] << 16), cs[2 * k]);
        cs[2 * k + 1] = fmaf(w[i], __uint_as_float(q[k] & 0xFFFF0000u), cs[2 * k + 1]);
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.y) * cols + col);
  dst[0] = make_float4(cs[0], cs[1], cs[2], cs[3]);
  dst[1] = make_float4(cs[4], cs[5], cs[6], cs[7]);
}

// dp.world > 0: element c goes to its owner's inbox (slot = this rank) instead of `out` -- the bias gradients take
// the same fused reduce-scatter route as the weight gradients (EpiStoreScatter)
__global__ void colsum_final_kernel(const float* __restrict__ partial, int row_blocks, int cols, float scale,
                                    float* __restrict__ out, DpScatter dp, int64_t region_off) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.0f;
  for (int b = 0; b < row_blocks; ++b) s += partial[static_cast<int64_t>(b) * cols + c];
  if (dp.world > 0) {
    const int per = cols / dp.world;
    const int owner = c / per;
    dp.inbox[owner][static_cast<int64_t>(dp.rank) * dp.shard_elems + region_off + (c - owner * per)] = s * scale;
  } else {
    out[c] = s * scale;
  }
}

// ---------------------------------------------------------------------------
// Second half of the data-parallel exchange.  Every rank's inbox now holds, slot by slot, the N contributions to ITS
// shard of the flat gradient (written by the peers' GEMM epilogues).  The owner adds them in rank order -- the same
// order on every run: bit-reproducible -- and writes each sum into the flat gradient buffer of EVERY rank:
// one multimem.st through the NVSwitch multicast mapping when there is one (the switch replicates the store), else
// one peer store per rank.  All ranks end up with bit-identical averaged gradients (DDP semantics,
// ospo/utils/train.py:26-28) without a collective library call on the data path.
// ---------------------------------------------------------------------------
struct DpGather {
  float* flat[8];         // flat gradient buffer of every rank as mapped into this process
  float* flat_mc;         // multicast mapping of the same buffer (null: use the peer stores)
};

__device__ __forceinline__ void multimem_st_f32x4(float* mc_addr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(256)
dp_reduce_broadcast_kernel(const float* __restrict__ inbox_local, DpGather g, int world, int rank, int64_t shard_elems,
                           int64_t a_elems /* dW2 rows of a shard */, int64_t b_elems /* dW1 rows */, int64_t vb /* db2 */,
                           int64_t VE, int64_t EH, int64_t V, int64_t i_begin, int64_t i_end /* shard elements to do */) {
  for (int64_t i4 = i_begin / 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < i_end / 4;
       i4 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = i4 * 4;
    // all slots are requested before the first addition (memory-level parallelism); the additions keep rank order
    float4 slot[8];
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (s < world) slot[s] = __ldcs(reinterpret_cast<const float4*>(inbox_local + static_cast<int64_t>(s) * shard_elems + i));
    float4 acc = slot[0];
#pragma unroll
    for (int s = 1; s < 8; ++s) {
      if (s < world) {
        acc.x += slot[s].x;
        acc.y += slot[s].y;
        acc.z += slot[s].z;
        acc.w += slot[s].w;
      }
    }
    // shard element -> element of the flat gradient dW2 | dW1 | db2 | db1
    int64_t f;
    if (i < a_elems) f = static_cast<int64_t>(rank) * a_elems + i;
    else if (i < a_elems + b_elems) f = VE + static_cast<int64_t>(rank) * b_elems + (i - a_elems);
    else if (i < a_elems + b_elems + vb) f = VE + EH + static_cast<int64_t>(rank) * vb + (i - a_elems - b_elems);
    else f = VE + EH + V + static_cast<int64_t>(rank) * (shard_elems - a_elems - b_elems - vb) + (i - a_elems - b_elems - vb);
    if (g.flat_mc != nullptr) {
      multimem_st_f32x4(g.flat_mc + f, acc);
    } else {
      for (int s = 0; s < world; ++s) *reinterpret_cast<float4*>(g.flat[s] + f) = acc;
    }
  }
}

// prepare_gen_img_embeds from the memo table: out[row, :] = table[ids[row / id_repeat], :]
__global__ void __launch_bounds__(256)
embed_table_gather_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ table, int codebook,
                          __nv_bfloat16* __restrict__ out, int n, int D, int id_repeat) {
  pdl_launch_dependents();
  pdl_wait();  // the ids come from the sampler
  const int row = blockIdx.x;
  if (row >= n) return;
  int64_t id = ids[row / id_repeat];
  id = id < 0 ? 0 : (id >= codebook ? codebook - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + id * D);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * D);
  for (int i = threadIdx.x; i < D / 8; i += blockDim.x) dst[i] = __ldg(src + i);
}

// ---------------------------------------------------------------------------
// decode GEMM1 finalize: act[n, e] = bf16(gelu_erf(bf16(sum_k part[k][n][e] + b1[e])))  (fixed summation order).
// Small footprint on purpose (128-thread blocks, 4 elements per thread): it sits between the two decode GEMMs of a
// programmatic-dependent-launch chain and must fit beside their resident CTAs.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
decode_act_finalize_kernel(const float* __restrict__ part, int k_splits, int64_t split_stride,
                           const float* __restrict__ b1, __nv_bfloat16* __restrict__ act, int n, int E,
                           int trace) {
  pdl_launch_dependents();  // the next GEMM may become resident and prefetch its weights; it waits for us
  if (threadIdx.x == 0) trace_stamp(trace ? 3 : 0, 0);
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(trace ? 3 : 0, 2);
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= static_cast<int64_t>(n) * E) return;
  float4 s = __ldcg(reinterpret_cast<const float4*>(part + i));
  for (int k = 1; k < k_splits; ++k) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(part + static_cast<int64_t>(k) * split_stride + i));
    s.x += t.x;
    s.y += t.y;
    s.z += t.z;
    s.w += t.w;
  }
  const float4 b = __ldg(reinterpret_cast<const float4*>(b1 + (i % E)));
  float x[4] = {bf16_round(s.x + b.x), bf16_round(s.y + b.y), bf16_round(s.z + b.z), bf16_round(s.w + b.w)};
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = 0.5f * x[j] * (1.0f + erff(x[j] * 0.70710678118654752f));
  uint2 o;
  o.x = pack_bf16x2(x[0], x[1]);
  o.y = pack_bf16x2(x[2], x[3]);
  *reinterpret_cast<uint2*>(act + i) = o;
  if (threadIdx.x == 0) trace_stamp(trace ? 3 : 0, 5);
}

// ---------------------------------------------------------------------------
// prepare_gen_img_embeds, first half (janus/models/modeling_vlm.py:263-264, projector.py:39-45):
//   a[n, d] = bf16(gelu_erf(bf16(sum_k gen_embed[id_n, k] * Wa[d, k] + ba[d])))      (k = 8: the VQ code dimension)
// The second Linear (D x D, 33.5 MB at 7B) is the weight-streaming swap-AB GEMM launch_decode_linear_cluster.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gen_embed_up_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ gen_embed, int codebook,
                    const __nv_bfloat16* __restrict__ wa, const float* __restrict__ ba, __nv_bfloat16* __restrict__ a,
                    int n, int D, int id_repeat) {
  pdl_launch_dependents();
  pdl_wait();  // the ids come from the sampler
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (d >= D || row >= n) return;
  int64_t id = ids[row / id_repeat];
  id = id < 0 ? 0 : (id >= codebook ? codebook - 1 : id);
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(gen_embed + id * 8));
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(wa + static_cast<int64_t>(d) * 8));
  const uint32_t ew[4] = {e.x, e.y, e.z, e.w}, ww[4] = {w.x, w.y, w.z, w.w};
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc = fmaf(__uint_as_float(ew[k] << 16), __uint_as_float(ww[k] << 16), acc);
    acc = fmaf(__uint_as_float(ew[k] & 0xFFFF0000u), __uint_as_float(ww[k] & 0xFFFF0000u), acc);
  }
  const float x = bf16_round(acc + __ldg(ba + d));
  a[static_cast<int64_t>(row) * D + d] = __float2bfloat16_rn(0.5f * x * (1.0f + erff(x * 0.70710678118654752f)));
}

// ---------------------------------------------------------------------------
// Weight pre-packing for the decode step: W [rows, cols] bf16 row-major -> tiles [slab = rows/128][kb = cols/64] of
// 16 KB, each the exact shared-memory image TMA would produce for a {64 cols x 128 rows} box with 128-byte swizzle
// (row r at byte r*128, its 16-byte chunk c stored at chunk position c ^ (r & 7)).  The decode kernel then pulls a
// tile with ONE contiguous bulk copy instead of 128 row requests: better DRAM locality (6.5 vs 6.1 TB/s streamed,
// scripts/gpu_stream_probe.py) and no tensor-map traffic.  One thread moves one 16-byte chunk.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_kernel(const __nv_bfloat16* __restrict__ w, int rows, int cols, uint4* __restrict__ packed) {
  const int num_kb = (cols + 63) / 64;
  const int64_t chunk = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over tiles x 128 rows x 8 chunks
  const int64_t total = static_cast<int64_t>((rows + 127) / 128) * num_kb * 1024;
  if (chunk >= total) return;
  const int c = static_cast<int>(chunk & 7);
  const int r = static_cast<int>((chunk >> 3) & 127);
  const int64_t tile = chunk >> 10;
  const int kb = static_cast<int>(tile % num_kb);
  const int64_t slab = tile / num_kb;
  const int64_t row = slab * 128 + r;
  const int col = kb * 64 + c * 8;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (row < rows && col + 8 <= cols) {
    v = __ldg(reinterpret_cast<const uint4*>(w + row * cols + col));
  } else if (row < rows && col < cols) {
    __nv_bfloat16 tmp[8];
    for (int j = 0; j < 8; ++j) tmp[j] = (col + j < cols) ? w[row * cols + col + j] : __float2bfloat16_rn(0.0f);
    v = *reinterpret_cast<const uint4*>(tmp);
  }
  packed[tile * 1024 + r * 8 + (c ^ (r & 7))] = v;
}

// ---------------------------------------------------------------------------
// CFG merge + temperature + softmax + inverse-CDF sampling on bf16 logits [2P, V]
// (row 2k = conditional, row 2k+1 = unconditional: ospo/wrapper/image_generation.py:135-141,157-158).
//   merge_mode 0 (reference bf16 semantics, op-by-op rounding, image_generation.py:160-161):
//        d = bf16(lc - lu); e = bf16(w * d); m = bf16(lu + e); t = bf16(m / T)
//   merge_mode 1: same four operations in fp32 (no intermediate bf16 rounding).
// Softmax weights are kept relative to a power of two:  e^t = P(r) 2^n  (n = rint(t log2 e), Cody-Waite r,
// degree-6 P);  inside a 128-code TILE  u = P(r) 2^(n - K_tile),  K_tile = max n of the tile;  a 32-code
// SEGMENT sum is a pairwise-adjacent tree;  globally K = max K_tile and S' = S 2^(K_tile - K) (exact
// rescale);  16 GROUPS of 32 segments and then the groups are summed sequentially;
//     id = first code whose running cdf exceeds u * Z   (descent group -> segment -> code).
// The same arithmetic, operation for operation, is in oracle/cfg_sample.c.  The tile-local exponent is what
// lets the decode GEMM's epilogue (EpiCfgFused, epilogues.cuh) produce the weights and segment sums while
// the logits are still in tensor memory.  greedy: argmax_v t_v, lowest index on ties.
// ---------------------------------------------------------------------------
// The descent shared by the stand-alone sampler and the finish kernel of the fused decode step.
// Shared-memory layout of the 512 rescaled segment sums: every 32-segment group starts on a 16-byte boundary and is
// followed by four padding words, so that a group is read as eight 16-byte words and the 16 lanes which each add up
// one group (stride 36 words) cover all banks per quarter-warp.
constexpr int SEG_GRP_PITCH = SAMPLE_GRP + 4;
constexpr int SEG_PAD_WORDS = (SAMPLE_THREADS / SAMPLE_GRP) * SEG_GRP_PITCH;
__device__ __forceinline__ int seg_slot(int seg) { return seg + (seg / SAMPLE_GRP) * (SEG_GRP_PITCH - SAMPLE_GRP); }
constexpr int DESCENT_SCRATCH = 72;  // floats of 16-byte aligned shared memory private to the descending warp

// The descent, executed by ONE WARP.  seg_sum[512] (already rescaled to the global exponent, padded slots) lives in
// shared memory.  The oracle's sequential form is  "base = 0; for each element: nxt = base + x; if (nxt > target)
// stop; base = nxt"  with the last element winning when nothing stops.  Here every lane evaluates the running sum of
// a level redundantly (the same additions in the same order) and leaves every partial sum in the warp's scratch
// (pref[i] = the sum before element i; all lanes store the same value to the same word); afterwards lane i picks up
// pref[i + 1] and one ballot finds the first crossing.  A level costs one addition and one store per element (the
// elements arrive four at a time) -- a compare + two selects per element in the chain made the serial tail of the
// sampler, 950 instructions per draw, the limiter of the whole kernel.
// Group sums first (lane g < 16 adds group g's 32 segment sums in order), then the levels group -> segment.
__device__ __forceinline__ float chain4(float run, const float4 v, float* pref) {
  run = __fadd_rn(run, v.x);
  pref[0] = run;
  run = __fadd_rn(run, v.y);
  pref[1] = run;
  run = __fadd_rn(run, v.z);
  pref[2] = run;
  run = __fadd_rn(run, v.w);
  pref[3] = run;
  return run;
}
__device__ __forceinline__ void warp_descent_segments(const float* seg_sum, float* grp_sum, float* scratch, float u01,
                                                      int lane, int& segi, float& base, float& target) {
  constexpr int NGRP = SAMPLE_THREADS / SAMPLE_GRP;  // 16
  if (lane < NGRP) {
    const float4* gp = reinterpret_cast<const float4*>(seg_sum + lane * SEG_GRP_PITCH);
    float g = 0.0f;
#pragma unroll
    for (int j = 0; j < SAMPLE_GRP / 4; ++j) {
      const float4 v = gp[j];
      g = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(g, v.x), v.y), v.z), v.w);
    }
    grp_sum[lane] = g;
  }
  __syncwarp();
  float* pref = scratch;
  float run = 0.0f;
  pref[0] = 0.0f;
#pragma unroll
  for (int g = 0; g < NGRP / 4; ++g) run = chain4(run, reinterpret_cast<const float4*>(grp_sum)[g], pref + 4 * g + 1);
  __syncwarp();
  target = __fmul_rn(u01, run);  // run = Z
  uint32_t hit = __ballot_sync(0xffffffffu, lane < NGRP - 1 && pref[(lane & (NGRP - 1)) + 1] > target);
  const int g_win = hit ? __ffs(hit) - 1 : NGRP - 1;
  run = pref[g_win];
  __syncwarp();
  pref[0] = run;
  const float4* sp = reinterpret_cast<const float4*>(seg_sum + g_win * SEG_GRP_PITCH);
#pragma unroll
  for (int i = 0; i < SAMPLE_GRP / 4; ++i) run = chain4(run, sp[i], pref + 4 * i + 1);
  __syncwarp();
  hit = __ballot_sync(0xffffffffu, lane < SAMPLE_GRP - 1 && pref[lane + 1] > target);
  const int sg_win = hit ? __ffs(hit) - 1 : SAMPLE_GRP - 1;
  base = pref[sg_win];
  segi = g_win * SAMPLE_GRP + sg_win;
  __syncwarp();  // pref is rewritten by the next level
}
// Last level: lane j holds the (rescaled) weight of code j of the winning segment; returns the code's index in it.
__device__ __forceinline__ int warp_descent_codes(float wl, float base, float target, int lane, float* scratch) {
  float* pref = scratch;
  float* wv = scratch + 36;
  wv[lane] = wl;
  __syncwarp();
  float run = base;
#pragma unroll
  for (int i = 0; i < SAMPLE_SEG / 4; ++i) run = chain4(run, reinterpret_cast<const float4*>(wv)[i], pref + 4 * i + 1);
  __syncwarp();
  const uint32_t hit = __ballot_sync(0xffffffffu, lane < SAMPLE_SEG - 1 && pref[lane + 1] > target);
  __syncwarp();
  return hit ? __ffs(hit) - 1 : SAMPLE_SEG - 1;
}

// Persistent blocks: block b handles pairs b, b + gridDim.x, ...; logits row pitch ld; vocab must be 16384 (= 512 * 32).
// Thread i < 512 owns segment i (codes 32 i .. 32 i + 31), register-resident.
//  * While a pair is being evaluated (ALU-bound: ~14 instructions per code) the block's next pair travels
//    global -> shared as bulk copies (cp.async.bulk, one mbarrier per warp): every WARP fetches exactly the 2 x 2 KB
//    its lanes read back themselves, lane 0 issuing the two copies once the warp has taken the current rows into
//    registers -- no block-wide barrier, 2 instructions per pair instead of 8 16-byte cp.async per thread (whose
//    sector-by-sector arrival cost 16 shared-memory write cycles per instruction).  The rows land linearly; lane l
//    reads its four 16-byte chunks in the order q ^ ((l >> 1) & 3) so a quarter-warp covers all 32 banks.  The
//    thread's codes are therefore held chunk-permuted in registers (slot q = chunk q ^ s): the tile exponent is a
//    maximum and the butterfly tree sum adds chunk 0 + chunk 2 and chunk 1 + chunk 3 element-wise and then the two
//    results -- both invariant under an XOR permutation of the chunks (IEEE addition commutes), so the segment sum
//    has the same bits; the greedy arg-max and the merged-logits store use the true index.
//  * Sampling: the serial part of a draw (global exponent, 16 sequential group sums, descent group -> segment ->
//    code) belongs to two extra warps, one for the block's even pairs and one for the odd.  The 16 evaluating warps
//    only deposit their segment sum and tile exponent in that tail warp's shared buffer (mbarriers full_bar[b] /
//    empty_bar[b]) and go on to the next pair; the tail warp rebuilds the 32 weights of the winning segment from the
//    logits (same operations, same bits) instead of asking its owner.
//    With the CTA-wide barriers the draw used to need, the warps spent 4.7 cycles waiting per instruction issued.
// WBF: MODE 0 with a bf16-exact cfg_weight -> the merge runs on the bf16x2 pipe (cfg_math.cuh)
constexpr int SAMPLE_BLOCK = SAMPLE_THREADS + 64;  // sampling variant: + two tail warps (even / odd pairs of the block)

template <int MODE, bool TDIV, bool WBF>
__device__ __forceinline__ void merge_words(uint32_t wc, uint32_t wu, float cfg_weight, float temperature, float& t0,
                                            float& t1) {
  if constexpr (WBF) cfg_merge2_hw<TDIV>(wc, wu, pack_bf16x2(cfg_weight, cfg_weight), temperature, t0, t1);
  else cfg_merge2<MODE, TDIV>(wc, wu, cfg_weight, temperature, t0, t1);
}

template <int MODE, bool TDIV, bool GREEDY, bool WBF = false>
__global__ void __launch_bounds__(GREEDY ? SAMPLE_THREADS : SAMPLE_BLOCK, 2)
cfg_merge_sample_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, int vocab, float cfg_weight,
                        float temperature, const float* __restrict__ uniforms, int64_t* __restrict__ ids,
                        float* __restrict__ merged_out /* [P, V] optional */, int pairs) {
  extern __shared__ uint4 next_rows[];  // [2 rows][16 warps][2 KB]: the block's next pair, linear
  __shared__ uint64_t row_bar[SAMPLE_THREADS / 32];  // one per evaluating warp: its 2 x 2 KB have landed
  __shared__ uint64_t full_bar[2], empty_bar[2];
  __shared__ __align__(16) float seg_buf[2][SEG_PAD_WORDS];  // sampling: S relative to the tile exponent, rescaled in place by the tail
  __shared__ int kt_buf[2][SAMPLE_THREADS];  // sampling: tile exponent as exp_koff(kt) (an integer + a constant), one copy per segment
  __shared__ __align__(16) float grp_sums[2][SAMPLE_THREADS / SAMPLE_GRP];
  __shared__ __align__(16) float descent_scratch[2][DESCENT_SCRATCH];
  __shared__ float wmax[SAMPLE_THREADS / 32];   // greedy
  __shared__ int warg[SAMPLE_THREADS / 32];     // greedy
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // hand-over of buffer b (pairs k of the block with k & 1 == b) between the 16 evaluating warps and tail warp b:
  // full_bar[b] collects one arrival per evaluating warp, empty_bar[b] the tail warp's.  mbarriers rather than named
  // barriers: an evaluating warp never waits for its 15 siblings, only (two pairs later) for the tail warp, so the
  // warps drift apart and their load / merge / scan phases overlap the others' FMA-bound weight phase.
  if (tid < SAMPLE_THREADS / 32) mbar_init(&row_bar[tid], 1);
  if (tid < 2) {
    mbar_init(&full_bar[tid], SAMPLE_THREADS / 32);
    mbar_init(&empty_bar[tid], 1);
  }
  fence_mbar_init();
  __syncthreads();

  if (!GREEDY && warp >= SAMPLE_THREADS / 32) {
    // ===================== tail warps: one draw per pair; warp 16 takes the block's even pairs, warp 17 the odd =====
    const int b = warp - SAMPLE_THREADS / 32;
    float* grp_sum = grp_sums[b];
    uint32_t use = 0;  // how many times this buffer has been handed over
    for (int p = blockIdx.x + b * gridDim.x; p < pairs; p += 2 * gridDim.x, ++use) {
      float* seg_sum = seg_buf[b];
      const int* kts = kt_buf[b];
      const float u01 = __ldg(uniforms + p);
      mbar_wait(&full_bar[b], use & 1, 0x5B00u + b);
      // global exponent K = max tile exponent; rescale the 512 segment sums (exact: powers of two)
      int kq[SAMPLE_THREADS / 32];
      int K = kts[lane];
      kq[0] = K;
#pragma unroll
      for (int i = 1; i < SAMPLE_THREADS / 32; ++i) {
        kq[i] = kts[i * 32 + lane];
        K = max(K, kq[i]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) K = max(K, __shfl_xor_sync(0xffffffffu, K, off));
#pragma unroll
      for (int i = 0; i < SAMPLE_THREADS / 32; ++i) {
        const int sl = i * SEG_GRP_PITCH + lane;  // = seg_slot(i * 32 + lane)
        seg_sum[sl] = __fmul_rn(seg_sum[sl], pow2_factor_i(kq[i] - K));
      }
      __syncwarp();
      int segi;
      float base, target;
      warp_descent_segments(seg_sum, grp_sum, descent_scratch[b], u01, lane, segi, base, target);
      // the winning segment's 32 weights again, lane j = code j (two codes per packed word, as in the evaluation)
      const int kseg = kts[segi];
      const float kt_seg = __fsub_rn(__int_as_float(kseg), kRintMagic);  // exact: kseg is the bit pattern of kt + 1.5 * 2^23
      const float f = pow2_factor_i(kseg - K);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[b]);  // the buffer is free for the pair after next
      const uint32_t* rc = reinterpret_cast<const uint32_t*>(logits + static_cast<int64_t>(2 * p) * ld) + segi * (SAMPLE_SEG / 2);
      const uint32_t* ru = reinterpret_cast<const uint32_t*>(logits + static_cast<int64_t>(2 * p + 1) * ld) + segi * (SAMPLE_SEG / 2);
      float t0, t1;
      merge_words<MODE, TDIV, WBF>(__ldg(rc + (lane >> 1)), __ldg(ru + (lane >> 1)), cfg_weight, temperature, t0, t1);
      float n;
      const float pr = exp_parts((lane & 1) ? t1 : t0, n);
      const float wl = __fmul_rn(__fmul_rn(pr, pow2_factor(__fsub_rn(n, kt_seg))), f);
      const int j = warp_descent_codes(wl, base, target, lane, descent_scratch[b]);
      if (lane == 0) ids[p] = static_cast<int64_t>(segi) * SAMPLE_SEG + j;
    }
    return;
  }

  // ===================== evaluating warps =====================
  constexpr uint32_t kWarpRowBytes = 32 * SAMPLE_SEG * 2;      // 2 KB: one warp's share of a logits row
  constexpr uint32_t kRowBytes = SAMPLE_THREADS * SAMPLE_SEG * 2;  // 32 KB
  uint8_t* landing = reinterpret_cast<uint8_t*>(next_rows) + warp * kWarpRowBytes;  // this warp's cond share; uncond at + 32 KB
  const int sw = (lane >> 1) & 3;  // chunk permutation of this lane: register slot q holds chunk q ^ sw
  const uint8_t* mine = landing + lane * (SAMPLE_SEG * 2);
  uint64_t* bar = &row_bar[warp];
  const int64_t pair_stride = 2 * ld * static_cast<int64_t>(gridDim.x);  // elements between this block's pairs
  const __nv_bfloat16* nxt = logits + static_cast<int64_t>(2 * blockIdx.x) * ld + warp * (32 * SAMPLE_SEG);  // next pair to fetch
  auto fetch = [&]() {  // lane 0 only
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bar, 2 * kWarpRowBytes);
    bulk_load_evict_first(landing, nxt, kWarpRowBytes, bar);
    bulk_load_evict_first(landing + kRowBytes, nxt + ld, kWarpRowBytes, bar);
  };
  if (static_cast<int>(blockIdx.x) < pairs && lane == 0) fetch();
  nxt += pair_stride;
  int k = 0;
  for (int p = blockIdx.x; p < pairs; p += gridDim.x, ++k) {
    // ---- load + merge: slot j of t = code 8 * ((j >> 3) ^ sw) + (j & 7) of segment tid -----------------
    float t[SAMPLE_SEG];
    {
      uint4 a[4], b[4];
      mbar_wait(bar, k & 1, 0x5A00u + warp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = *reinterpret_cast<const uint4*>(mine + ((i ^ sw) << 4));
        b[i] = *reinterpret_cast<const uint4*>(mine + kRowBytes + ((i ^ sw) << 4));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t wa[4] = {a[i].x, a[i].y, a[i].z, a[i].w}, wb[4] = {b[i].x, b[i].y, b[i].z, b[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          merge_words<MODE, TDIV, WBF>(wa[q], wb[q], cfg_weight, temperature, t[8 * i + 2 * q], t[8 * i + 2 * q + 1]);
      }
    }
    // the rows of this block's next pair stream in while this one is evaluated (the warp's values are in registers)
    __syncwarp();
    if (p + static_cast<int>(gridDim.x) < pairs && lane == 0) fetch();
    nxt += pair_stride;
    if (merged_out != nullptr) {
      float4* mo = reinterpret_cast<float4*>(merged_out + static_cast<int64_t>(p) * vocab + tid * SAMPLE_SEG);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mo[2 * (i ^ sw)] = make_float4(t[8 * i], t[8 * i + 1], t[8 * i + 2], t[8 * i + 3]);
        mo[2 * (i ^ sw) + 1] = make_float4(t[8 * i + 4], t[8 * i + 5], t[8 * i + 6], t[8 * i + 7]);
      }
    }

    if constexpr (GREEDY) {
      // ---- arg-max (exact; lowest index wins ties) ----------------------------------------------
      // per 8-code chunk in index order (strict > keeps the lowest index), then the four chunks by (value, index)
      float lmax = 0.0f;
      int larg = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float cmax = t[8 * i];
        int carg = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          if (t[8 * i + j] > cmax) {
            cmax = t[8 * i + j];
            carg = j;
          }
        }
        carg += (i ^ sw) << 3;
        if (i == 0 || cmax > lmax || (cmax == lmax && carg < larg)) {
          lmax = cmax;
          larg = carg;
        }
      }
      larg += tid * SAMPLE_SEG;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, lmax, off);
        const int oi = __shfl_xor_sync(0xffffffffu, larg, off);
        if (o > lmax || (o == lmax && oi < larg)) {
          lmax = o;
          larg = oi;
        }
      }
      if (lane == 0) {
        wmax[warp] = lmax;
        warg[warp] = larg;
      }
      __syncthreads();
      if (tid == 0) {
        float gmax = wmax[0];
        int garg = warg[0];
        for (int w = 1; w < SAMPLE_THREADS / 32; ++w) {
          if (wmax[w] > gmax || (wmax[w] == gmax && warg[w] < garg)) {
            gmax = wmax[w];
            garg = warg[w];
          }
        }
        ids[p] = garg;
      }
      __syncthreads();  // wmax / warg are reused by the next pair
    } else {
      // ---- tile exponent: K_tile = max n over the 4 segments (= 4 adjacent lanes) of a 128-code tile ------
      // (n is a non-decreasing function of t -- a correctly rounded multiply by a positive constant, two clamps and a
      // round-to-integer -- so max n = n(max t): one fmax per code instead of four operations)
      // Range scan first (NaN-propagating): when the clamp of t log2 e to +-1e4 is idle for the whole warp, the
      // weights come from the form without it (same bits).
      float hi = t[0], lo = t[0];
#pragma unroll
      for (int j = 1; j < SAMPLE_SEG - 1; j += 2) {
        hi = max3_nan(hi, t[j], t[j + 1]);
        lo = min3_nan(lo, t[j], t[j + 1]);
      }
      hi = max3_nan(hi, t[SAMPLE_SEG - 1], t[SAMPLE_SEG - 1]);
      lo = min3_nan(lo, t[SAMPLE_SEG - 1], t[SAMPLE_SEG - 1]);
      const float yhi = __fmul_rn(hi, 1.4426950408889634f), ylo = __fmul_rn(lo, 1.4426950408889634f);
      const bool inrange = __all_sync(0xffffffffu, yhi <= 1.0e4f && ylo >= -1.0e4f);  // false with a NaN or an infinity
      float kt;
      if (inrange) {
        kt = rintf(yhi);  // = exp_n_only(max t): no NaN, clamp idle
      } else {
        float tmax = t[0];
#pragma unroll
        for (int j = 1; j < SAMPLE_SEG; ++j) tmax = fmaxf(tmax, t[j]);
        kt = exp_n_only(tmax);
      }
      kt = fmaxf(kt, __shfl_xor_sync(0xffffffffu, kt, 1));
      kt = fmaxf(kt, __shfl_xor_sync(0xffffffffu, kt, 2));
      // ---- weights relative to K_tile and the segment's tree sum -------------------------------------
      const int koff = exp_koff(kt);
      uint64_t w[SAMPLE_SEG / 2];
      // no code of the warp below the 2^-120 cut-off of its tile (n(min t) - kt >= -120; n is monotone in t)?
      const bool nounderflow =
          inrange && __all_sync(0xffffffffu, exp_koff(rintf(ylo)) - koff >= -120);
      if (nounderflow) {
        const uint32_t kbias = exp_kbias(koff);
#pragma unroll
        for (int j = 0; j < SAMPLE_SEG / 2; ++j) w[j] = exp_weight2p_nounderflow(f2_pack(t[2 * j], t[2 * j + 1]), kbias);
      } else if (inrange) {
        const int kcut = koff - 121;
#pragma unroll
        for (int j = 0; j < SAMPLE_SEG / 2; ++j) w[j] = exp_weight2p_cut(f2_pack(t[2 * j], t[2 * j + 1]), kcut);
      } else {
#pragma unroll
        for (int j = 0; j < SAMPLE_SEG / 2; ++j) w[j] = exp_weight2p(f2_pack(t[2 * j], t[2 * j + 1]), koff, koff);
      }
      const float S = tree_sum32_packed(w);
      // ---- hand the segment over to the tail warp ----------------------------------------------------
      const int b = k & 1;
      if (k >= 2) mbar_wait(&empty_bar[b], ((k >> 1) - 1) & 1, 0x5C00u + b);  // its draw of the pair before last is over
      seg_buf[b][seg_slot(tid)] = S;
      kt_buf[b][tid] = koff;
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[b]);
    }
  }
}

// ---------------------------------------------------------------------------
// Finish kernel of the fused decode step.  The decode GEMM2 epilogue (EpiCfgFused) has already produced, per
// pair: the weights u (relative to each tile's exponent), 512 segment sums, 128 tile exponents and, for the
// greedy mode, per-tile arg-max candidates.  One block per pair finishes the draw.
// ---------------------------------------------------------------------------

// Optional tail (next row N1): once the pair's id is known the block also evaluates the first gen_aligner layer for
// it, a1[2p] = a1[2p+1] = bf16(gelu(bf16(gen_embed[id] . Wa^T + ba)))  (image_generation.py:166-167 up to the D x D
// Linear), so `sample -> next-step embeddings` needs no launch of its own for this stage.
struct EmbedUp {
  const __nv_bfloat16* gen_embed;  // [codebook, 8] or null = off
  const __nv_bfloat16* wa;         // [D, 8]
  const float* ba;                 // [D]
  __nv_bfloat16* a1;               // [2P, D]
  int codebook, D;
  // Optional memo of the whole aligner: table[id, :] = gen_aligner(gen_embed(id)) for every code of the VQ codebook
  // (bf16 [codebook, D]; the module is a pure function of the id, and generation runs with frozen weights).  When set
  // the pair's two output rows are copied from it -- `a1` then IS the next step's input embeddings and no Linear runs.
  const __nv_bfloat16* table;
};

constexpr int EMBED_PER_THREAD = 8;  // D <= 8 * 512 is served from registers loaded before the id is known

struct EmbedRegs {
  uint4 w[EMBED_PER_THREAD];
  float b[EMBED_PER_THREAD];
};

// the aligner's first-layer weights do not depend on anything this step computes: fetch them up front
__device__ __forceinline__ void embed_up_preload(const EmbedUp& eu, EmbedRegs& r) {
#pragma unroll
  for (int i = 0; i < EMBED_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * SAMPLE_THREADS;
    if (d < eu.D) {
      r.w[i] = __ldg(reinterpret_cast<const uint4*>(eu.wa + static_cast<int64_t>(d) * 8));
      r.b[i] = __ldg(eu.ba + d);
    }
  }
}

__device__ __forceinline__ __nv_bfloat16 embed_up_one(const uint32_t (&ew)[4], const uint4& w, float bias) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
  float acc = 0.0f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc = fmaf(__uint_as_float(ew[k] << 16), __uint_as_float(ww[k] << 16), acc);
    acc = fmaf(__uint_as_float(ew[k] & 0xFFFF0000u), __uint_as_float(ww[k] & 0xFFFF0000u), acc);
  }
  const float x = bf16_round(acc + bias);
  return __float2bfloat16_rn(0.5f * x * (1.0f + erff(x * 0.70710678118654752f)));
}

__device__ __forceinline__ void embed_up_rows(const EmbedUp& eu, const EmbedRegs& r, int p, int id) {
  id = id < 0 ? 0 : (id >= eu.codebook ? eu.codebook - 1 : id);
  if (eu.table != nullptr) {
    // 16-byte vectors: D % 8 == 0 and the bases are 16-byte aligned (checked at the ABI)
    const uint4* src = reinterpret_cast<const uint4*>(eu.table + static_cast<int64_t>(id) * eu.D);
    uint4* d0 = reinterpret_cast<uint4*>(eu.a1 + static_cast<int64_t>(2 * p) * eu.D);
    uint4* d1 = reinterpret_cast<uint4*>(eu.a1 + static_cast<int64_t>(2 * p + 1) * eu.D);
    for (int i = threadIdx.x; i < eu.D / 8; i += blockDim.x) {
      const uint4 v = __ldg(src + i);
      d0[i] = v;
      d1[i] = v;
    }
    return;
  }
  const uint4 e = __ldg(reinterpret_cast<const uint4*>(eu.gen_embed + static_cast<int64_t>(id) * 8));
  const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
  __nv_bfloat16* row0 = eu.a1 + static_cast<int64_t>(2 * p) * eu.D;
  __nv_bfloat16* row1 = row0 + eu.D;
#pragma unroll
  for (int i = 0; i < EMBED_PER_THREAD; ++i) {
    const int d = threadIdx.x + i * SAMPLE_THREADS;
    if (d < eu.D) {
      const __nv_bfloat16 y = embed_up_one(ew, r.w[i], r.b[i]);
      row0[d] = y;
      row1[d] = y;
    }
  }
  for (int d = threadIdx.x + EMBED_PER_THREAD * SAMPLE_THREADS; d < eu.D; d += SAMPLE_THREADS) {  // D > 4096
    const __nv_bfloat16 y =
        embed_up_one(ew, __ldg(reinterpret_cast<const uint4*>(eu.wa + static_cast<int64_t>(d) * 8)), __ldg(eu.ba + d));
    row0[d] = y;
    row1[d] = y;
  }
}

// EMBED: with the first aligner layer folded in (a separate instantiation: the plain finish kernel stays small)
template <bool EMBED>
__global__ void __launch_bounds__(SAMPLE_THREADS)
cfg_finish_kernel(CfgFusedBuffers b, int vocab, const float* __restrict__ uniforms, int greedy,
                  int64_t* __restrict__ ids, int trace, EmbedUp eu) {
  __shared__ __align__(16) float seg_sum[SEG_PAD_WORDS];
  __shared__ __align__(16) float grp_sum[SAMPLE_THREADS / SAMPLE_GRP];
  __shared__ __align__(16) float descent_scratch[DESCENT_SCRATCH];
  __shared__ float wmax[SAMPLE_THREADS / 32];
  __shared__ int bc_id;
  pdl_launch_dependents();  // successors may become resident and prefetch; they wait for our completion
  if (threadIdx.x == 0) trace_stamp(trace ? 4 : 0, 0);
  EmbedRegs er;
  if constexpr (EMBED) {
    if (eu.table == nullptr) embed_up_preload(eu, er);
  }
  pdl_wait();
  if (threadIdx.x == 0) trace_stamp(trace ? 4 : 0, 2);
  const int p = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int ntile = vocab / SAMPLE_TILE;  // 128
  if (greedy) {
    if (tid == 0) {
      float gmax = b.tile_max[p * ntile];
      int garg = b.tile_arg[p * ntile];
      for (int t = 1; t < ntile; ++t) {
        const float o = b.tile_max[p * ntile + t];
        const int oi = b.tile_arg[p * ntile + t];
        if (o > gmax || (o == gmax && oi < garg)) {
          gmax = o;
          garg = oi;
        }
      }
      ids[p] = garg;
      bc_id = garg;
    }
    if constexpr (EMBED) {
      __syncthreads();
      embed_up_rows(eu, er, p, bc_id);
    }
    return;
  }
  // the tile exponent, the sum of this thread's segment and the uniform are requested together (one L2 round trip).
  // (Also preloading every segment's 32 weights, so that the winner could walk them from registers, was measured
  // slower: 64 KB of 16-byte-strided loads per block cost more than the one dependent 128-byte load they save.)
  const float kt = b.tile_k[p * ntile + tid / (SAMPLE_TILE / SAMPLE_SEG)];
  const float ss_raw = b.seg_sum[static_cast<int64_t>(p) * SAMPLE_THREADS + tid];
  const float u01 = __ldg(uniforms + p);
  float K = kt;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) K = fmaxf(K, __shfl_xor_sync(0xffffffffu, K, off));
  if (lane == 0) wmax[warp] = K;
  __syncthreads();
  K = wmax[0];
#pragma unroll
  for (int w = 1; w < SAMPLE_THREADS / 32; ++w) K = fmaxf(K, wmax[w]);
  const float f = pow2_factor(__fsub_rn(kt, K));
  seg_sum[seg_slot(tid)] = __fmul_rn(ss_raw, f);
  __syncthreads();
  if (warp == 0) {
    // group sums, descent group -> segment, then the winning segment's 32 weights (one coalesced load) -> code
    int segi;
    float base, target;
    warp_descent_segments(seg_sum, grp_sum, descent_scratch, u01, lane, segi, base, target);
    const float fs = pow2_factor(__fsub_rn(b.tile_k[p * ntile + segi / (SAMPLE_TILE / SAMPLE_SEG)], K));
    const float wl = __fmul_rn(__ldcg(b.wbuf + static_cast<int64_t>(p) * vocab + segi * SAMPLE_SEG + lane), fs);
    const int j = warp_descent_codes(wl, base, target, lane, descent_scratch);
    if (lane == 0) {
      ids[p] = static_cast<int64_t>(segi) * SAMPLE_SEG + j;
      bc_id = segi * SAMPLE_SEG + j;
      trace_stamp(trace ? 4 : 0, 5);
    }
  }
  if constexpr (EMBED) {
    __syncthreads();
    embed_up_rows(eu, er, p, bc_id);
  }
}

}  // namespace ospo
