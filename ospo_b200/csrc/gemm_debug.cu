// Validation entry: run one engine variant as a plain fp32-output GEMM so tests can check every
// (cta_group, tile, operand-major) combination against a reference product.
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_debug(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }

using EpiF = EpiStore<float, false, false>;

template <class Cfg>
static int run(const LaunchCtx& c, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b, int64_t ldb, float* out,
               int64_t ldo, int M, int N, int K) {
  EpiF::Params p{out, ldo, nullptr};
  return launch_gemm<Cfg, EpiF>(a, lda, b, ldb, M, N, K, c.group_m, p, c.num_sms, c.stream);
}

// variant = cta_group * 100 + majors * 10 + tile   (majors: 0 = K/K, 1 = K/MN, 2 = MN/MN;  tile: 0 = BN256, 1 = BN32,
// 2 = BN128)
int launch_gemm_debug(const LaunchCtx& c, int variant, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b,
                      int64_t ldb, float* out, int64_t ldo, int M, int N, int K) {
  switch (variant) {
    case 100: return run<GemmCfg<1, 256, false, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 110: return run<GemmCfg<1, 256, false, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 120: return run<GemmCfg<1, 256, true, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 101: return run<GemmCfg<1, 32, false, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 102: return run<GemmCfg<1, 128, false, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    // decode-chain configurations of the BN=32 tile (tuning probes)
    case 103: return run<GemmCfg<1, 32, false, false, 0, 5, 2, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 104: return run<GemmCfg<1, 32, false, false, 0, 5, 1, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 105: return run<GemmCfg<1, 32, false, false, 0, 8, 2, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 106: return run<GemmCfg<1, 32, false, false, 0, 8, 1, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 200: return run<GemmCfg<2, 256, false, false>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 210: return run<GemmCfg<2, 256, false, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    case 220: return run<GemmCfg<2, 256, true, true>>(c, a, lda, b, ldb, out, ldo, M, N, K);
    default: return -100;
  }
}

}  // namespace ospo
