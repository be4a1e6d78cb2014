// Host-side launchers for the tcgen05 GEMM instantiations.  Each translation unit instantiates a
// few (tile, operand-major, epilogue) combinations of gemm_sm100.cuh; abi.cu strings them together.
// cta_group: 1 = one CTA per 128-row tile, 2 = CTA pair per 256-row tile (tcgen05 cta_group::2).
#pragma once

#include <atomic>

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ospo {

struct LaunchCtx {
  int num_sms;
  int cta_group;  // 1 or 2
  int group_m;    // rasterisation group (M-blocks)
  cudaStream_t stream;
  bool pdl;       // launch with programmatic stream serialization (decode chain)
  bool trace = false;  // timeline stamps on (ospo_head_trace); off = the stamps compile to a parameter test
  unsigned long long* trace_buf = nullptr;
  int wgrad_splitk = 0;  // k-splits of the weight-gradient GEMMs: 0 = chosen per shape (gemm_bwd.cu), 1 / 2 = forced
  int a_evict = 0, b_evict = 0;  // L2 eviction hint of the A / B operand loads: 0 normal, 1 evict-first, 2 evict-last
  uint32_t* sync_ctr = nullptr;  // wave lock-step counter for the persistent training GEMMs (null = free-running)  // the installed timeline buffer (merged decode kernel stamps through it)
};

struct CfgFusedBuffers;

// Data-parallel exchange fused into the weight-gradient stores: where a rank's contribution to each
// owner's shard of the flat gradient goes.
struct DpScatter {
  float* inbox[8];        // inbox base of every rank as mapped into THIS process (symmetric memory); [rank] is local
  int world;              // 0 = scatter off
  int rank;
  int64_t shard_elems;    // elements of one rank's shard of the flat gradient (= size of one inbox slot)
};

// cudaFuncSetAttribute is per device: `mask` (one static per kernel instantiation) has one bit per device ordinal.
// Returns true when the calling thread's current device still needs the attributes set; call mark afterwards.
inline bool func_attrs_needed(std::atomic<uint64_t>& mask, int* dev_out) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  *dev_out = dev;
  return !(mask.load(std::memory_order_acquire) & (1ull << dev));
}
inline void func_attrs_mark(std::atomic<uint64_t>& mask, int dev) { mask.fetch_or(1ull << dev, std::memory_order_release); }


// Row-segmented hidden states: x / dx are [segments, seg_pitch, H] tensors of which rows [seg_off, seg_off +
// seg_rows) of every segment are the head's rows (seg_rows == 0: plain contiguous [rows, H]).
struct XLayout {
  int seg_rows = 0;
  int seg_pitch = 0;
  int seg_off = 0;
  int segments = 0;
};

// every TU that contains kernels using mbar_wait owns a copy of the watchdog pointer
void set_watchdog_fwd(uint32_t* dev_ptr);
void set_watchdog_bwd(uint32_t* dev_ptr);
void set_watchdog_decode(uint32_t* dev_ptr);
void set_watchdog_debug(uint32_t* dev_ptr);
void set_trace_decode(unsigned long long* dev_ptr);
void set_watchdog_merged(uint32_t* dev_ptr);
void set_trace_merged(unsigned long long* dev_ptr);

// ---- forward (gemm_fwd.cu) ----
// pre = bf16(x W1^T + b1), act = bf16(gelu(pre));  x [rows,H], w1 [E,H]
int launch_gemm1_bias_gelu(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w1, const float* b1,
                           __nv_bfloat16* pre /*nullable*/, __nv_bfloat16* act, int rows, int H, int E,
                           const XLayout& xl = XLayout());
// forward GEMM2 with the softmax numerator fused in: l = bf16(act W2^T + b2), e = exp(l - row_ref) spilled as bf16
// (+ per sub-tile (max l, sum e) partials, target gather, logits row sums);  act [rows,E], w2 [V,E].
// row_ref null = 0; blk_mask non-null = repair pass over the flagged M-blocks only (gemm2_tile_m rows each), which
// returns at once when *any_flag == 0.
int launch_gemm2_logits_exp(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                            __nv_bfloat16* espill /*nullable*/, const int64_t* labels, float2* part,
                            float* rowsum_part /*nullable*/, float* tgt, const float* row_ref /*nullable*/,
                            const uint8_t* blk_mask /*nullable*/, const uint8_t* any_flag /*nullable*/, int rows,
                            int E, int V);
int gemm2_tile_m(int cta_group);
int gemm2_num_n_tiles(int V);
// plain logits = bf16(act W2^T + b2)
int launch_gemm2_logits(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                        __nv_bfloat16* logits, int64_t ld, int rows, int E, int V);

// ---- backward (gemm_bwd.cu) ----
// dpre = bf16( bf16(row_w * (g W2)) * gelu'(pre) ), act_w = bf16(row_w * gelu(pre));  g [rows,V] (the forward's
// spill), w2 [V,E];  act_w nullable (head frozen)
int launch_dact_gelu_bwd(const LaunchCtx& c, const __nv_bfloat16* g, const __nv_bfloat16* w2, const __nv_bfloat16* pre,
                         const float* row_w, __nv_bfloat16* dpre, __nv_bfloat16* act_w /*nullable*/, int rows, int E,
                         int V);
// dW[out_dim, in_dim] (fp32) = scale * dY^T X;  dY [rows,out_dim], X [rows,in_dim]  (scale 0 = 1)
int launch_wgrad(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* x, float* dw, int rows, int out_dim,
                 int in_dim, float scale = 0.0f, const XLayout& xl = XLayout());
// the weight-gradient GEMM with the data-parallel reduce-scatter fused into its store: row block r of dW goes into
// rank (r / rows_per_rank)'s inbox over NVLink peer memory (slot = this rank).  -100: the partition does not fit.
int launch_wgrad_scatter(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* x, const DpScatter& dp,
                         int64_t region_off, int rows, int out_dim, int in_dim, float scale,
                         const XLayout& xl = XLayout());
// dX[rows, in_dim] (bf16) = dY W;  dY [rows,out_dim], W [out_dim,in_dim]
int launch_dgrad(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* w, __nv_bfloat16* dx, int rows,
                 int out_dim, int in_dim, const XLayout& xl = XLayout());

// ---- decode, swap-AB (gemm_decode.cu) ----
// part[ks][n][E] (fp32) = k-split partials of W1 h^T;  h [n,H] with n = 2P small: W1 is streamed as (128-row tile
// x k-split) work items over all SMs.  decode_act_finalize_kernel sums the partials in split order (deterministic),
// adds b1 and applies GELU.  decode_gemm1_splits() = number of partials written.
int decode_gemm1_splits(int num_sms, int H, int E);
int launch_decode_gemm1(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, float* part,
                        int64_t split_stride, int n, int H, int E);
// cluster split-K form of the decode GEMM1: the k-splits of a tile are one thread-block cluster and are combined
// through distributed shared memory; act[n, e] = bf16(gelu(bf16(W1 h^T + b1))) is final at kernel end.
// Returns -100 when the shape does not fit one work item per SM (use the partial + finalize path then).
int launch_decode_gemm1_cluster(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1,
                                __nv_bfloat16* act, int n, int H, int E);
// out[n, m] = bf16(W x^T + b) for a handful of rows n <= 32 (gen_aligner's second Linear): same swap-AB cluster
// split-K kernel without the activation
int launch_decode_linear_cluster(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* b,
                                 __nv_bfloat16* out, int n, int in_dim, int out_dim);
// logits[n, v] = bf16(W2 act^T + b2)
int launch_decode_gemm2(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                        __nv_bfloat16* logits, int n, int E, int V);
// the same GEMM with the CFG merge / softmax-weight / segment-sum tail fused into the epilogue (EpiCfgFused);
// merge_mode 0 = bf16 op-by-op, 1 = fp32; logits_dump may be null
int launch_decode_gemm2_fused(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                              __nv_bfloat16* logits_dump, int n, int E, int V, float cfg_weight, float temperature,
                              int merge_mode, int greedy, const CfgFusedBuffers& buf);

// ---- decode, one persistent kernel (decode_merged.cu) ----
// act = gelu(h W1^T + b1) (cluster split-K), device-wide flag, then W2 act^T + b2 with the fused CFG epilogue.
// flag: two zeroed words owned by this launch's stream (see abi.cu); the kernel leaves them zero.
// l2_ahead: weight k-blocks (16 KB each, per CTA) requested into L2 ahead of the shared-memory ring.
// next_w [next_rows, next_cols] (optional): the weight of the kernel that follows in the chain (gen_aligner's D x D
// Linear); it is requested into L2 behind this step's last weight tile.
// *grid_ctas receives the number of CTAs launched.
// Returns -100 when the shape / occupancy does not allow every CTA to be resident at once.
int launch_decode_merged(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1,
                         const __nv_bfloat16* w2, const float* b2, __nv_bfloat16* act, uint32_t* flag,
                         __nv_bfloat16* logits_dump, int n, int H, int E, int V, float cfg_weight, float temperature,
                         int merge_mode, int greedy, const CfgFusedBuffers& buf, int l2_ahead,
                         const void* w1_packed = nullptr, const void* w2_packed = nullptr,
                         const __nv_bfloat16* next_w = nullptr, int next_rows = 0, int next_cols = 0,
                         int* grid_ctas = nullptr);

// phase 1 of that kernel on its own: out[n, M] = bf16(act_fn(bf16(x W^T + b))), n <= 32 (gen_aligner's D x D Linear)
int launch_decode_linear(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* b,
                         __nv_bfloat16* out, int n, int K, int M, int gelu);

// ---- debug / validation (gemm_debug.cu) ----
// out[M,N] fp32 = A B^T for one engine variant; see abi.cu for the variant table
int launch_gemm_debug(const LaunchCtx& c, int variant, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                      int64_t ldb, float* out, int64_t ldo, int M, int N, int K);

}  // namespace ospo
