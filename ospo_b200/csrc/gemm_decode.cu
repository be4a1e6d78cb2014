// Decode-step GEMMs (CFG generation, 2P rows): swap-AB so the weight rows ride the UMMA M axis (128
// TMEM lanes) and the handful of sample rows ride UMMA N.  Weight streaming is the whole cost, so
// every CTA streams a disjoint 128-row slab of W exactly once.
// reference: gen_head(hidden_states[:, -1, :]) ospo/wrapper/image_generation.py:156
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_decode(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }

using CfgS32 = GemmCfg<1, 32, false, false>;
using CfgS128 = GemmCfg<1, 128, false, false>;

int launch_decode_gemm1(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1,
                        __nv_bfloat16* act, int n, int H, int E) {
  using Epi = EpiBiasGelu<true, false>;
  Epi::Params p{b1, nullptr, act, E};
  // D[E, n] = W1[E, H] * h[n, H]^T
  if (n <= 32) return launch_gemm<CfgS32, Epi>(w1, H, h, H, E, n, H, 1 << 20, p, c.num_sms, c.stream);
  return launch_gemm<CfgS128, Epi>(w1, H, h, H, E, n, H, 1 << 20, p, c.num_sms, c.stream);
}

int launch_decode_gemm2(const LaunchCtx& c, const __nv_bfloat16* act, const __nv_bfloat16* w2, const float* b2,
                        __nv_bfloat16* logits, int n, int E, int V) {
  using Epi = EpiStore<__nv_bfloat16, true, true>;
  Epi::Params p{logits, V, b2};
  // D[V, n] = W2[V, E] * act[n, E]^T, stored transposed as logits[n, V]
  if (n <= 32) return launch_gemm<CfgS32, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream);
  return launch_gemm<CfgS128, Epi>(w2, E, act, E, V, n, E, 1 << 20, p, c.num_sms, c.stream);
}

}  // namespace ospo
