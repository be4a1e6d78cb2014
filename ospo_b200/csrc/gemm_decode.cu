// Decode-step GEMMs (CFG generation, 2P rows): swap-AB so the weight rows ride the UMMA M axis (128
// TMEM lanes) and the handful of sample rows ride UMMA N.  Weight streaming is the whole cost, so
// every CTA streams a disjoint 128-row slab of W exactly once.
// reference: gen_head(hidden_states[:, -1, :]) ospo/wrapper/image_generation.py:156
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_decode(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }
void set_trace_decode(unsigned long long* dev_ptr) { cudaMemcpyToSymbol(g_trace_buf, &dev_ptr, sizeof(dev_ptr)); }

// decode chain: 4-stage ring (~82 KB), <= 128 registers, weights (A) prefetched before the dependency wait
using CfgS32 = GemmCfg<1, 32, false, false, 0, 5, 2, true>;
using CfgS128 = GemmCfg<1, 128, false, false>;

int decode_gemm1_splits(int num_sms, int H, int E) {
  // enough (tile, k-split) work items to occupy every SM with a disjoint slab of W1; at most 8 partials
  const int num_m = (E + 127) / 128;
  int want = num_sms / (num_m > 0 ? num_m : 1);
  if (want > 8) want = 8;
  int ks, per;
  gemm_split_plan((H + 63) / 64, want, &ks, &per);
  return ks;
}

int launch_decode_gemm1(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, float* part,
                        int64_t split_stride, int n, int H, int E) {
  using Epi = EpiPartialStoreT;
  Epi::Params p{part, E, split_stride};
  const int ks = decode_gemm1_splits(c.num_sms, H, E);
  // D[E, n] = W1[E, H] * h[n, H]^T, as k-split partials part[ks][n][E]
  if (n <= 32)
    return launch_gemm<CfgS32, Epi>(w1, H, h, H, E, n, H, 1 << 20, p, c.num_sms, c.stream, ks, c.pdl, SegOperand(),
                                    SegOperand(), c.trace ? 1 : 0);
  return launch_gemm<CfgS128, Epi>(w1, H, h, H, E, n, H, 1 << 20, p, c.num_sms, c.stream, ks, c.pdl);
}

// decode GEMM1, cluster split-K form: act is final when the kernel ends (no partials in HBM, no finalize launch).
// Returns -100 if the shape does not fit the one-item-per-CTA scheme; the caller then uses the partial + finalize path.
using CfgS32C = GemmCfg<1, 32, false, false, 0, 8, 1, true, true>;
template <bool GELU>
static int run_linear_cluster(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* b,
                              __nv_bfloat16* out, int n, int K, int M, int trace_id) {
  using Epi = EpiClusterLinearT<GELU>;
  if (n > 32) return -100;
  typename Epi::Params p{b, out, M};
  int ks = decode_gemm1_splits(c.num_sms, K, M);
  if (ks == 3) ks = 2;            // cluster sizes: 1, 2, 4, 8
  if (ks > 4 && ks < 8) ks = 4;
  // D[M, n] = W[M, K] * x[n, K]^T, k-splits of a tile = one cluster
  return launch_gemm<CfgS32C, Epi>(w, K, x, K, M, n, K, 1 << 20, p, c.num_sms, c.stream, ks, c.pdl, SegOperand(),
                                   SegOperand(), trace_id);
}

int launch_decode_gemm1_cluster(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1,
                                __nv_bfloat16* act, int n, int H, int E) {
  return run_linear_cluster<true>(c, h, w1, b1, act, n, H, E, c.trace ? 1 : 0);
}

int launch_decode_linear_cluster(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* b,
                                 __nv_bfloat16* out, int n, int in_dim, int out_dim) {
  return run_linear_cluster<false>(c, x, w, b, out, n, in_dim, out_dim, 0);
}

int launch_decode_gemm2(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                        __nv_bfloat16* logits, int n, int E, int V) {
  using Epi = EpiStore<__nv_bfloat16, true, true>;
  Epi::Params p{logits, V, b2};
  // D[V, n] = W2[V, E] * act[n, E]^T, stored transposed as logits[n, V]
  if (n <= 32) return launch_gemm<CfgS32, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream);
  return launch_gemm<CfgS128, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream);
}

// GEMM2: the deepest ring that fits (10 x 20 KB): while it waits for the activations its producer has 20 MB of W2
// in flight across the chip, which fills part of the GEMM1 -> GEMM2 dependency bubble.  (Measured alternative: a
// 3-stage GEMM1 + 7-stage GEMM2 sharing every SM starts the W2 prefetch earlier but slows the W1 stream by more
// than it gains: 48.4 vs 45.9 us per step.)
using CfgF32 = GemmCfg<1, 32, false, false, 1024, 10, 1, true>;

template <int MODE, bool TDIV>
static int run_fused(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                     __nv_bfloat16* logits_dump, int n, int E, int V, float cfg_weight, float temperature, int greedy,
                     const CfgFusedBuffers& buf) {
  // D[V, n] = W2[V, E] * act[n, E]^T; the epilogue consumes the tile in place
  if (greedy) {
    using Epi = EpiCfgFused<MODE, TDIV, false, true>;
    typename Epi::Params p{b2, cfg_weight, temperature, logits_dump, V, buf, greedy, V};
    return launch_gemm<CfgF32, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream, 1, c.pdl, SegOperand(),
                                    SegOperand(), c.trace ? 2 : 0);
  }
  using Epi = EpiCfgFused<MODE, TDIV, false, false>;
  typename Epi::Params p{b2, cfg_weight, temperature, logits_dump, V, buf, greedy, V};
  return launch_gemm<CfgF32, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream, 1, c.pdl, SegOperand(),
                                  SegOperand(), c.trace ? 2 : 0);
}

int launch_decode_gemm2_fused(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                              __nv_bfloat16* logits_dump, int n, int E, int V, float cfg_weight, float temperature,
                              int merge_mode, int greedy, const CfgFusedBuffers& buf) {
  const bool tdiv = (temperature != 1.0f);
  if (merge_mode == 0) {
    if (tdiv) return run_fused<0, true>(c, act, w2, b2, logits_dump, n, E, V, cfg_weight, temperature, greedy, buf);
    return run_fused<0, false>(c, act, w2, b2, logits_dump, n, E, V, cfg_weight, temperature, greedy, buf);
  }
  if (tdiv) return run_fused<1, true>(c, act, w2, b2, logits_dump, n, E, V, cfg_weight, temperature, greedy, buf);
  return run_fused<1, false>(c, act, w2, b2, logits_dump, n, E, V, cfg_weight, temperature, greedy, buf);
}

}  // namespace ospo
