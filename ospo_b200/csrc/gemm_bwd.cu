// Backward GEMMs of the image-token head.  No operand is ever transposed in memory:
//   data gradients   dA = dY W      : A operand K-major (dY [rows,out]), B operand MN-major (W [out,in])
//   weight gradients dW = dY^T X    : both operands MN-major (dY [rows,out], X [rows,in]; K = rows)
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_bwd(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }

using CfgD1 = GemmCfg<1, 256, false, true>;
using CfgD2 = GemmCfg<2, 256, false, true>;
using CfgW1 = GemmCfg<1, 256, true, true>;
using CfgW2 = GemmCfg<2, 256, true, true>;

int launch_dact_gelu_bwd(const LaunchCtx& c, const __nv_bfloat16* g, const __nv_bfloat16* w2, const __nv_bfloat16* pre,
                         const float* row_w, __nv_bfloat16* dpre, __nv_bfloat16* act_w, int rows, int E, int V) {
  using Epi = EpiDactScale;
  Epi::Params p{pre, dpre, E, row_w, act_w};
  // D[rows, E] = g[rows, V] * W2[V, E]:  M = rows, N = E, K = V
  if (c.cta_group == 2) return launch_gemm<CfgD2, Epi>(g, V, w2, E, rows, E, V, c.group_m, p, c.num_sms, c.stream, 1, false,
                                  SegOperand(), SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
  return launch_gemm<CfgD1, Epi>(g, V, w2, E, rows, E, V, c.group_m, p, c.num_sms, c.stream, 1, false,
                                  SegOperand(), SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
}

int launch_dgrad(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* w, __nv_bfloat16* dx, int rows,
                 int out_dim, int in_dim, const XLayout& xl) {
  using Epi = EpiStore<__nv_bfloat16, false, false>;
  Epi::Params p{dx, in_dim, nullptr, RowMap{xl.seg_rows, xl.seg_pitch, xl.seg_off}};
  if (c.cta_group == 2)
    return launch_gemm<CfgD2, Epi>(dy, out_dim, w, in_dim, rows, in_dim, out_dim, c.group_m, p, c.num_sms, c.stream, 1, false,
                                  SegOperand(), SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
  return launch_gemm<CfgD1, Epi>(dy, out_dim, w, in_dim, rows, in_dim, out_dim, c.group_m, p, c.num_sms, c.stream, 1, false,
                                  SegOperand(), SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
}

int launch_wgrad(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* x, float* dw, int rows, int out_dim,
                 int in_dim, float scale, const XLayout& xl) {
  using Epi = EpiStore<float, false, false>;
  Epi::Params p{dw, in_dim, nullptr, RowMap{0, 0, 0}, scale};
  SegOperand sb;
  sb.seg_rows = xl.seg_rows;
  sb.seg_pitch = xl.seg_pitch;
  sb.seg_off = xl.seg_off;
  sb.segments = xl.segments;
  // D[out, in] = dY^T[out, rows] * X[rows, in]:  M = out_dim, N = in_dim, K = rows
  if (c.cta_group == 2)
    return launch_gemm<CfgW2, Epi>(dy, out_dim, x, in_dim, out_dim, in_dim, rows, c.group_m, p, c.num_sms, c.stream, 1,
                                   false, SegOperand(), sb, 0, c.sync_ctr, c.a_evict, c.b_evict);
  return launch_gemm<CfgW1, Epi>(dy, out_dim, x, in_dim, out_dim, in_dim, rows, c.group_m, p, c.num_sms, c.stream, 1,
                                 false, SegOperand(), sb, 0, c.sync_ctr, c.a_evict, c.b_evict);
}

// the same GEMM with the reduce-scatter fused into the store (EpiStoreScatter): rows of dW go to their owner's inbox
int launch_wgrad_scatter(const LaunchCtx& c, const __nv_bfloat16* dy, const __nv_bfloat16* x, const DpScatter& dp,
                         int64_t region_off, int rows, int out_dim, int in_dim, float scale, const XLayout& xl) {
  using Epi = EpiStoreScatter;
  if (dp.world < 1 || dp.world > 8 || (out_dim % dp.world) != 0) return -100;
  const int rows_per_rank = out_dim / dp.world;
  if (rows_per_rank % (c.cta_group == 2 ? CfgW2::TILE_M : CfgW1::TILE_M)) return -100;  // a tile never straddles two owners
  Epi::Params p{dp, rows_per_rank, region_off, in_dim, scale != 0.0f ? scale : 1.0f};
  SegOperand sb;
  sb.seg_rows = xl.seg_rows;
  sb.seg_pitch = xl.seg_pitch;
  sb.seg_off = xl.seg_off;
  sb.segments = xl.segments;
  if (c.cta_group == 2)
    return launch_gemm<CfgW2, Epi>(dy, out_dim, x, in_dim, out_dim, in_dim, rows, c.group_m, p, c.num_sms, c.stream, 1,
                                   false, SegOperand(), sb, 0, c.sync_ctr, c.a_evict, c.b_evict);
  return launch_gemm<CfgW1, Epi>(dy, out_dim, x, in_dim, out_dim, in_dim, rows, c.group_m, p, c.num_sms, c.stream, 1,
                                 false, SegOperand(), sb, 0, c.sync_ctr, c.a_evict, c.b_evict);
}

}  // namespace ospo
