// Forward GEMMs of the image-token head (both operands K-major):
//   GEMM1  pre/act = gelu(x W1^T + b1)      reference: modeling_vlm.py:48-49
//   GEMM2  logits  = act W2^T + b2          reference: modeling_vlm.py:50  (+ fused LSE / target gather)
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_fwd(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }

using Cfg1 = GemmCfg<1, 256, false, false>;
using Cfg2 = GemmCfg<2, 256, false, false>;

// number of (max, sum-exp) partials per row: one per 256-column tile and epilogue column half
int gemm2_num_n_tiles(int V) { return ((V + 255) / 256) * Cfg2::EPI_SPLIT; }

int launch_gemm1_bias_gelu(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w1, const float* b1,
                           __nv_bfloat16* pre, __nv_bfloat16* act, int rows, int H, int E, const XLayout& xl) {
  SegOperand sa;
  sa.seg_rows = xl.seg_rows;
  sa.seg_pitch = xl.seg_pitch;
  sa.seg_off = xl.seg_off;
  sa.segments = xl.segments;
  if (pre != nullptr) {
    using Epi = EpiBiasGelu<false, true>;
    Epi::Params p{b1, pre, act, E};
    if (c.cta_group == 2) return launch_gemm<Cfg2, Epi>(x, H, w1, H, rows, E, H, c.group_m, p, c.num_sms, c.stream, 1, false, sa, SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
    return launch_gemm<Cfg1, Epi>(x, H, w1, H, rows, E, H, c.group_m, p, c.num_sms, c.stream, 1, false, sa, SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
  } else {
    using Epi = EpiBiasGelu<false, false>;
    Epi::Params p{b1, nullptr, act, E};
    if (c.cta_group == 2) return launch_gemm<Cfg2, Epi>(x, H, w1, H, rows, E, H, c.group_m, p, c.num_sms, c.stream, 1, false, sa, SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
    return launch_gemm<Cfg1, Epi>(x, H, w1, H, rows, E, H, c.group_m, p, c.num_sms, c.stream, 1, false, sa, SegOperand(), 0, c.sync_ctr, c.a_evict, c.b_evict);
  }
}

int gemm2_tile_m(int cta_group) { return cta_group == 2 ? Cfg2::TILE_M : Cfg1::TILE_M; }

int launch_gemm2_logits_exp(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                            __nv_bfloat16* espill, const int64_t* labels, float2* part, float* rowsum_part, float* tgt,
                            const float* row_ref, const uint8_t* blk_mask, const uint8_t* any_flag, int rows, int E,
                            int V) {
  using Epi = EpiLogitsExp;
  Epi::Params p{b2, espill, V, labels, part, rowsum_part, tgt, row_ref, blk_mask, any_flag};
  // the repair pass (blk_mask set) computes a handful of tiles at most: no wave lock-step
  uint32_t* sync = blk_mask ? nullptr : c.sync_ctr;
  if (c.cta_group == 2) return launch_gemm<Cfg2, Epi>(act, E, w2, E, rows, V, E, c.group_m, p, c.num_sms, c.stream, 1, false, SegOperand(), SegOperand(), 0,
                                 sync, c.a_evict, c.b_evict);
  return launch_gemm<Cfg1, Epi>(act, E, w2, E, rows, V, E, c.group_m, p, c.num_sms, c.stream, 1, false, SegOperand(), SegOperand(), 0,
                                 sync, c.a_evict, c.b_evict);
}

int launch_gemm2_logits(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                        __nv_bfloat16* logits, int64_t ld, int rows, int E, int V) {
  using Epi = EpiStore<__nv_bfloat16, false, false>;
  Epi::Params p{logits, ld, b2};
  if (c.cta_group == 2) return launch_gemm<Cfg2, Epi>(act, E, w2, E, rows, V, E, c.group_m, p, c.num_sms, c.stream, 1, false, SegOperand(), SegOperand(), 0,
                                 c.sync_ctr, c.a_evict, c.b_evict);
  return launch_gemm<Cfg1, Epi>(act, E, w2, E, rows, V, E, c.group_m, p, c.num_sms, c.stream, 1, false, SegOperand(), SegOperand(), 0,
                                 c.sync_ctr, c.a_evict, c.b_evict);
}

}  // namespace ospo
