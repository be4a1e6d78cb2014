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
  return launch_gemm<Cfg, EpiF>(a, lda, b, ldb, M, N, K, c.group_m, p, c.num_sms, c.stream, 1, false, SegOperand(),
                                SegOperand(), 0, c.sync_ctr);
}

// ---------------------------------------------------------------------------
// Weight-stream probes (tuning aid, variants 300-303): how fast can the chip pull a [M, K] bf16 matrix into shared
// memory with the decode kernel's ring (10 x 16 KB per CTA), and does the order of the bytes matter?
//   300  2-D tensor map, {64 cols x 128 rows} boxes walked along K (the decode GEMMs' pattern: 128 B per row)
//   301  contiguous 16 KB chunks (1-D bulk copies), one contiguous slab per CTA, M/128 CTAs
//   302  the same with one CTA per SM
//   303  contiguous 16 KB chunks, CTAs interleaved chunk by chunk, one CTA per SM
// Nothing is computed; out[0] receives a token so the launch has a visible effect.
// ---------------------------------------------------------------------------
constexpr int kProbeStages = 10, kProbeChunk = 16384;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64, 1)
stream_probe_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* base, int num_kb, long long chunks,
                    float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kProbeStages * kProbeChunk);
  uint64_t* empty_bar = full_bar + kProbeStages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kProbeStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // work of this CTA: n items starting at `first`, `stride` apart (in chunks)
  long long first, stride, n;
  const long long G = gridDim.x, b = blockIdx.x;
  if (MODE == 0) {
    first = b * num_kb;
    stride = 1;
    n = num_kb;
  } else if (MODE == 1) {
    const long long lo = chunks * b / G, hi = chunks * (b + 1) / G;
    first = lo;
    stride = 1;
    n = hi - lo;
  } else {
    first = b;
    stride = G;
    n = (chunks - b + G - 1) / G;
  }
  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < n; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 21);
        mbar_arrive_expect_tx(&full_bar[s], kProbeChunk);
        if (MODE == 0) tma_load_2d(smem + s * kProbeChunk, &tmap, &full_bar[s], static_cast<int>(i) * 64,
                                   static_cast<int>(b) * 128, kEvictNormal);
        else bulk_load_1d(smem + s * kProbeChunk, base + (first + i * stride) * kProbeChunk, kProbeChunk, &full_bar[s]);
        if (++s == kProbeStages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < n; ++i) {
        mbar_wait(&full_bar[s], ph, 22);
        mbar_arrive(&empty_bar[s]);
        if (++s == kProbeStages) { s = 0; ph ^= 1u; }
      }
      if (blockIdx.x == 0) out[0] = 1.0f;
    }
  }
}

// The same stream with the decode GEMM's second operand beside it (variants 304 / 305): every stage also takes a
// {64 cols x 32 rows} box of a [32, K] activation matrix -- 304: ONE matrix shared by all CTAs (what the decode
// kernels do), 305: a private copy per CTA.  No MMA: what do 128 SMs asking for the same 4 KB at the same time cost?
constexpr int kProbeBBytes = 32 * 64 * 2;
template <bool PRIVATE_B>
__global__ void __launch_bounds__(96, 1)
stream_probe_b_kernel(const __grid_constant__ CUtensorMap tmap_b, const uint8_t* base, int num_kb, long long chunks,
                      float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr int kStage = kProbeChunk + kProbeBBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kProbeStages * kStage);
  uint64_t* empty_bar = full_bar + kProbeStages;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kProbeStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const long long G = gridDim.x, b = blockIdx.x;
  const long long lo = chunks * b / G, hi = chunks * (b + 1) / G;
  const long long n = hi - lo;
  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < n; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 21);
        mbar_arrive_expect_tx(&full_bar[s], kStage);
        bulk_load_1d(smem + s * kStage, base + (lo + i) * kProbeChunk, kProbeChunk, &full_bar[s]);
        if (++s == kProbeStages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < n; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1u, 23);
        const int kb = static_cast<int>((lo + i) % num_kb);
        tma_load_2d(smem + s * kStage + kProbeChunk, &tmap_b, &full_bar[s], kb * 64,
                    PRIVATE_B ? static_cast<int>(b) * 32 : 0, kEvictLast);
        if (++s == kProbeStages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long i = 0; i < n; ++i) {
        mbar_wait(&full_bar[s], ph, 22);
        mbar_arrive(&empty_bar[s]);
        if (++s == kProbeStages) { s = 0; ph ^= 1u; }
      }
      if (blockIdx.x == 0) out[0] = 1.0f;
    }
  }
}

// b: [32 * (PRIVATE_B ? ctas : 1), K] bf16
template <bool PRIVATE_B>
static int run_probe_b(const LaunchCtx& c, const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b, int64_t ldb,
                       float* out, int M, int K, int ctas) {
  if ((M % 128) != 0 || (K % 64) != 0 || lda != K || ldb != K) return -100;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, b, PRIVATE_B ? 32 * ctas : 32, K, ldb, 32);
  if (rc != 0) return rc;
  auto kern = stream_probe_b_kernel<PRIVATE_B>;
  const int smem = 1024 + kProbeStages * (kProbeChunk + kProbeBBytes) + 256;
  static std::atomic<uint64_t> attr_set{0};  // per device ordinal
  int attr_dev = 0;
  if (func_attrs_needed(attr_set, &attr_dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
    func_attrs_mark(attr_set, attr_dev);
  }
  const long long chunks = static_cast<long long>(M) * K * 2 / kProbeChunk;
  kern<<<ctas, 96, smem, c.stream>>>(tm, reinterpret_cast<const uint8_t*>(a), K / 64, chunks, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

template <int MODE>
static int run_probe(const LaunchCtx& c, const __nv_bfloat16* a, int64_t lda, float* out, int M, int K, int ctas) {
  if ((M % 128) != 0 || (K % 64) != 0 || lda != K) return -100;
  CUtensorMap tm;
  int rc = make_tmap_bf16_2d(&tm, a, M, K, lda, 128);
  if (rc != 0) return rc;
  auto kern = stream_probe_kernel<MODE>;
  const int smem = 1024 + kProbeStages * kProbeChunk + 256;
  static std::atomic<uint64_t> attr_set{0};  // per device ordinal
  int attr_dev = 0;
  if (func_attrs_needed(attr_set, &attr_dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -3;
    func_attrs_mark(attr_set, attr_dev);
  }
  const long long chunks = static_cast<long long>(M) * K * 2 / kProbeChunk;
  kern<<<ctas, 64, smem, c.stream>>>(tm, reinterpret_cast<const uint8_t*>(a), K / 64, chunks, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
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
    case 300: return run_probe<0>(c, a, lda, out, M, K, M / 128);
    case 301: return run_probe<1>(c, a, lda, out, M, K, M / 128);
    case 302: return run_probe<1>(c, a, lda, out, M, K, c.num_sms);
    case 303: return run_probe<2>(c, a, lda, out, M, K, c.num_sms);
    case 304: return run_probe_b<false>(c, a, lda, b, ldb, out, M, K, M / 128);
    case 305: return run_probe_b<true>(c, a, lda, b, ldb, out, M, K, M / 128);
    default: return -100;
  }
}

}  // namespace ospo
