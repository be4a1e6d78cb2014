// Persistent, warp-specialised tcgen05 GEMM engine for sm_100a.
//
//   D[M, N] = A[M, K] * B[N, K]^T      (bf16 operands, fp32 accumulation in TMEM)
//
// * operands arrive by TMA (128-byte swizzle) into a multi-stage shared-memory ring,
// * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, or 256 x BN x 16 with cta_group::2),
// * accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1,
// * four epilogue warps read TMEM with tcgen05.ld (one accumulator row per thread) and hand each
//   32-column chunk to an epilogue functor (bias, GELU, online log-sum-exp, target gather, ...).
//
// Either operand may be "K-major" (global layout [rows, K], K contiguous) or "MN-major"
// (global layout [K, rows], rows contiguous); the latter is what the two weight-gradient GEMMs
// and the two data-gradient GEMMs of the head need, so no transposed copies are ever made.
//
// Warp roles:  0 = TMA producer, 1 = MMA issuer (also allocates / frees TMEM),
//              2.. = epilogue (warp w owns TMEM lanes 32*(w%4) .. +32; with 8 epilogue warps, warps 2-5
//              take the left half of the tile's columns and warps 6-9 the right half).
#pragma once

#include <atomic>
#include <cstdlib>

#include "launchers.h"
#include "ptx.cuh"

namespace ospo {

struct GemmDims {
  int M, N, K;
  int group_m;       // rasterisation: tiles are walked M-fastest inside groups of `group_m` M-blocks
  int k_splits;      // split-K: each output tile is produced as k_splits partial tiles (1 = off)
  int kb_per_split;  // k-blocks (of BK) per split; every split is non-empty
  // Row-segmented operand (0 = off): the logical rows of the operand are the rows [seg_off, seg_off + seg_rows)
  // of every segment of a [segments, pitch, cols] tensor (hidden states [S, L+T, H] -> the T image-token rows of
  // each sequence).  seg_rows must be a multiple of 64; the tensor map is then 3-D and rows move in 64-row boxes.
  // a_seg_*: K-major A (rows = M axis);  b_seg_*: MN-major B (rows = K axis).
  int a_seg_rows, a_seg_off;
  int b_seg_rows, b_seg_off;
  int trace_id;  // > 0: CTA timeline stamps into g_trace_buf (tuning aid)
  // Wave lock-step (0 / null = off): the clusters of a persistent grid start their i-th tile together for the first
  // sync_tiles tiles.  Tiles that run at the same time share operand panels through L2 only while they walk K at
  // the same pace; free-running clusters drift apart by more than L2 retains and the shared panels are fetched
  // from HBM several times.  The wait is bounded and falls through (it is a performance hint, never a dependency).
  uint32_t* sync_ctr;
  int sync_tiles;
  int sync_stride;  // lock-step every this many tiles (short tiles need it less often)
  int sync_kb;      // > 0: additional lock-step points inside a tile, every sync_kb k-blocks (very long K: the
                    // clusters of a wave drift apart within one tile by more than L2 retains)
  // L2 eviction hints of the operand loads (kEvictNormal / kEvictFirst / kEvictLast): the operand whose panels are
  // re-used by later waves is kept (evict-last), the one that streams through once per wave goes first
  uint64_t a_hint, b_hint;
};

// where a row-mapped output row lands: logical row r -> physical row of a [segments, pitch, cols] tensor
struct RowMap {
  int seg_rows;   // 0 = identity
  int seg_pitch;
  int seg_off;
  __device__ __forceinline__ int64_t operator()(int r) const {
    if (seg_rows == 0) return r;
    const int s = r / seg_rows;
    return static_cast<int64_t>(s) * seg_pitch + seg_off + (r - s * seg_rows);
  }
};

// watchdog site ids
enum : uint32_t {
  SITE_PRODUCER_EMPTY = 1,
  SITE_MMA_FULL = 2,
  SITE_MMA_TMEM_EMPTY = 3,
  SITE_EPI_TMEM_FULL = 4,
};

// MAX_STAGES_ / MIN_BLOCKS_ / A_INDEPENDENT_ serve the decode chain: a short ring (about 80 KB) and <= 128
// registers let the NEXT kernel of a programmatic-dependent-launch chain become resident beside the running
// one, and A_INDEPENDENT_ (the A operand -- a weight matrix -- does not depend on the predecessor kernel) lets
// its producer start streaming weights before griddepcontrol.wait, so HBM never idles at a kernel boundary.
// CLUSTER_SPLIT_: the k-splits of one output tile are the CTAs of one thread-block cluster (one work item per CTA);
// each CTA parks its fp32 partial tile in its own shared memory and CTA 0 of the cluster combines them through
// distributed shared memory in split order (Epi::cluster_finalize) -- no partials in HBM, no second launch.
template <int CG_, int BN_, bool A_MN_, bool B_MN_, int EPI_SMEM_BYTES_ = 0, int MAX_STAGES_ = 8, int MIN_BLOCKS_ = 1,
          bool A_INDEPENDENT_ = false, bool CLUSTER_SPLIT_ = false>
struct GemmCfg {
  static constexpr bool CLUSTER_SPLIT = CLUSTER_SPLIT_;
  static_assert(!CLUSTER_SPLIT_ || CG_ == 1, "cluster split-K uses single-CTA MMAs");
  static constexpr int MIN_BLOCKS = MIN_BLOCKS_;
  static constexpr bool A_INDEPENDENT = A_INDEPENDENT_;
  static constexpr int CG = CG_;            // CTAs cooperating on one tile (cta_group)
  static constexpr int BN = BN_;            // tile N (UMMA N)
  static constexpr bool A_MN = A_MN_;
  static constexpr bool B_MN = B_MN_;
  static constexpr int BM = 128;            // accumulator rows per CTA (TMEM lanes)
  static constexpr int TILE_M = BM * CG;    // tile M (UMMA M)
  static constexpr int BK = 64;             // K per stage: 64 bf16 = one 128-byte swizzle row
  static constexpr int UMMA_K = 16;
  static constexpr int B_ROWS = BN / CG;    // N extent of B held by each CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_SMEM_BYTES = EPI_SMEM_BYTES_;
  static constexpr int SMEM_BUDGET = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - EPI_SMEM_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > MAX_STAGES_ ? MAX_STAGES_ : STAGES_RAW;
  static_assert(STAGES <= 12, "barrier area holds at most 12 stages");
  static constexpr int ACC_STAGES = (2 * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS_RAW = ACC_STAGES * BN;
  static constexpr int TMEM_COLS =
      TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256 + EPI_SMEM_BYTES;
  // Epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, each taking half of the tile's
  // columns) -- a lone warp per scheduler cannot hide the latency of a math-heavy epilogue.
  static constexpr int EPI_SPLIT = (BN >= 64) ? 2 : 1;
  static constexpr int EPI_WARPS = 4 * EPI_SPLIT;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  // register cap for the co-resident decode chain: bounding on one extra warp leaves room (65536 regs / SM) for
  // two GEMM CTAs plus the small kernel that sits between them
  static constexpr int BOUND_THREADS = (MIN_BLOCKS > 1) ? THREADS + 32 : THREADS;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N must be a multiple of 16 in [16, 256]");
  static_assert(!B_MN || (B_ROWS % 64 == 0), "MN-major B is staged in 64-row swizzle atoms");
  static_assert(B_ROWS % 8 == 0, "K-major B rows come in 8-row swizzle groups");
  static_assert(STAGES >= 2, "need at least a double-buffered ring");
};

struct TileCoord {
  int m_blk, n_blk;
};

__device__ __forceinline__ TileCoord tile_coord(int t, int num_m, int num_n, int group_m, int m_base = 0) {
  const int tiles_per_group = group_m * num_n;
  const int g = t / tiles_per_group;
  const int first_m = g * group_m;
  const int gm = min(num_m - first_m, group_m);
  const int local = t - g * tiles_per_group;
  TileCoord c;
  c.m_blk = m_base + first_m + local % gm;
  c.n_blk = local / gm;
  return c;
}

// Epilogue functor concept:
//   struct Epi {
//     struct Params { ... };                      // POD, passed by value to the kernel
//     struct State  { ... };                      // per-thread registers that live across one tile
//     static constexpr int SMEM_BYTES;
//     __device__ static void begin(const Params&, State&, int row, int n0, int k_split, const GemmDims&, uint8_t* smem);
//     template <bool FULL>
//     __device__ static void chunk(const Params&, State&, int row, int col0, float (&v)[32], const GemmDims&, ...);
//     __device__ static void end(const Params&, State&, int row, int n0, int sub_tile, const GemmDims&, ...);
//   };
// `row` is the global accumulator row owned by the calling thread (may be >= M: the functor must
// guard its global accesses), `col0` the global column of v[0].  FULL promises that all 32 columns of
// the chunk are inside N (the common case: no per-element predicates).  With EPI_SPLIT == 2 a (row, tile)
// is handled by two threads (column halves); `sub_tile` = n_blk * EPI_SPLIT + half identifies the part.

// Optional tile mask: an epilogue functor that declares `static constexpr bool HAS_TILE_MASK = true` and
// `static bool tile_enabled(const Params&, int m_blk)` makes all three roles skip the tiles of disabled M-blocks
// (the repair pass of the forward GEMM2 recomputes only the flagged row blocks; normally that is none).
template <class Epi, class = void>
struct epi_has_tile_mask {
  static constexpr bool value = false;
};
template <class Epi>
struct epi_has_tile_mask<Epi, decltype(void(Epi::HAS_TILE_MASK))> {
  static constexpr bool value = Epi::HAS_TILE_MASK;
};
template <class Epi, class P>
__device__ __forceinline__ bool epi_tile_enabled(const P& ep, int m_blk) {
  if constexpr (epi_has_tile_mask<Epi>::value) return Epi::tile_enabled(ep, m_blk);
  else return true;
}
// ... and `static bool launch_enabled(const Params&)`: false (the same for every thread of the grid) makes the whole
// launch return before it sets anything up.
template <class Epi, class P>
__device__ __forceinline__ bool epi_launch_enabled(const P& ep) {
  if constexpr (epi_has_tile_mask<Epi>::value) return Epi::launch_enabled(ep);
  else return true;
}

template <class Cfg, class Epi>
__global__ void __launch_bounds__(Cfg::BOUND_THREADS, Cfg::MIN_BLOCKS)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, GemmDims dims,
            typename Epi::Params ep) {
  constexpr int CG = Cfg::CG, BN = Cfg::BN, BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;

  pdl_launch_dependents();  // our successor may start its own prologue (and weight prefetch) right away
  if (!epi_launch_enabled<Epi>(ep)) return;  // grid-uniform: nothing to do (e.g. a repair pass with no flagged block)
  if (threadIdx.x == 0) trace_stamp(dims.trace_id, 0);
  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte alignment (in the shared address space)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (cta_rank == 0);

  const int num_m = (dims.M + Cfg::TILE_M - 1) / Cfg::TILE_M;
  const int num_n = (dims.N + BN - 1) / BN;
  const int num_kb = (dims.K + BK - 1) / BK;
  const int ksplits = dims.k_splits;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  // work domain of this cluster: (first tile, stride, tile count, M-block window)
  int dom_first = cluster_id, dom_stride = num_clusters, dom_m0 = 0, dom_nm = num_m;
  const int dom_tiles = dom_nm * num_n * ksplits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader producer's arrive.expect_tx (covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);  // one tcgen05.commit per stage use
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tmem_full_bar[a], 1);        // one tcgen05.commit per tile
      mbar_init(&tmem_empty_bar[a], Cfg::EPI_WARPS * CG);  // one arrive per epilogue warp of every CTA in the group
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<CG>(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // Programmatic dependent launch: everything above overlapped the predecessor; each role executes
  // pdl_wait() before it first touches global memory that the predecessor may still be producing.

  if (warp == 0) {
    // ===================== TMA producer (one elected thread) =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      auto load_a = [&](uint8_t* sa, int s_, int m0, int k0) {
        if constexpr (!Cfg::A_MN) {
          if (dims.a_seg_rows != 0) {
            // 64-row boxes, each inside one segment
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) {
              const int r = m0 + 64 * c;
              const int sg = r / dims.a_seg_rows;
              const int rr = dims.a_seg_off + (r - sg * dims.a_seg_rows);
              if constexpr (CG == 1) tma_load_3d(sa + c * 8192, &tmap_a, &full_bar[s_], k0, rr, sg, dims.a_hint);
              else tma_load_3d_2sm(sa + c * 8192, &tmap_a, &full_bar[s_], k0, rr, sg, dims.a_hint);
            }
          } else if constexpr (CG == 1) tma_load_2d(sa, &tmap_a, &full_bar[s_], k0, m0, dims.a_hint);
          else tma_load_2d_2sm(sa, &tmap_a, &full_bar[s_], k0, m0, dims.a_hint);
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c) {
            if constexpr (CG == 1) tma_load_2d(sa + c * (BK * 128), &tmap_a, &full_bar[s_], m0 + 64 * c, k0, dims.a_hint);
            else tma_load_2d_2sm(sa + c * (BK * 128), &tmap_a, &full_bar[s_], m0 + 64 * c, k0, dims.a_hint);
          }
        }
      };
      auto load_b = [&](uint8_t* sb, int s_, int n0, int k0) {
        if constexpr (!Cfg::B_MN) {
          if constexpr (CG == 1) tma_load_2d(sb, &tmap_b, &full_bar[s_], k0, n0, dims.b_hint);
          else tma_load_2d_2sm(sb, &tmap_b, &full_bar[s_], k0, n0, dims.b_hint);
        } else {
          const bool seg = dims.b_seg_rows != 0;
          int sg = 0, rr = k0;
          if (seg) {
            sg = k0 / dims.b_seg_rows;
            rr = dims.b_seg_off + (k0 - sg * dims.b_seg_rows);
          }
#pragma unroll
          for (int c = 0; c < Cfg::B_ROWS / 64; ++c) {
            if (seg) {
              if constexpr (CG == 1) tma_load_3d(sb + c * (BK * 128), &tmap_b, &full_bar[s_], n0 + 64 * c, rr, sg, dims.b_hint);
              else tma_load_3d_2sm(sb + c * (BK * 128), &tmap_b, &full_bar[s_], n0 + 64 * c, rr, sg, dims.b_hint);
            } else {
              if constexpr (CG == 1) tma_load_2d(sb + c * (BK * 128), &tmap_b, &full_bar[s_], n0 + 64 * c, k0, dims.b_hint);
              else tma_load_2d_2sm(sb + c * (BK * 128), &tmap_b, &full_bar[s_], n0 + 64 * c, k0, dims.b_hint);
            }
          }
        }
      };
      // Weight prefetch ahead of the dependency wait: the first ring-full of A tiles of this CTA's first
      // work item is requested now; their B halves follow after pdl_wait().
      int prefetched = 0;
      if constexpr (Cfg::A_INDEPENDENT) {
        if (dom_first < dom_tiles) {
          const int t = dom_first;
          const TileCoord tc = tile_coord(t / ksplits, dom_nm, num_n, dims.group_m, dom_m0);
          const int m0 = tc.m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BM;
          const int kb0 = (t % ksplits) * dims.kb_per_split;
          const int kb1 = min(num_kb, kb0 + dims.kb_per_split);
          prefetched = min(STAGES, kb1 - kb0);
          for (int i = 0; i < prefetched; ++i) {
            uint8_t* sa = stage_base + i * Cfg::STAGE_BYTES;
            if (is_leader) mbar_arrive_expect_tx(&full_bar[i], Cfg::STAGE_BYTES * CG);
            load_a(sa, i, m0, (kb0 + i) * BK);
          }
        }
      }
      trace_stamp(dims.trace_id, 1);
      pdl_wait();
      trace_stamp(dims.trace_id, 2);
      int tile_no = 0;
      bool sync_wait = true;
      uint32_t sync_idx = 0;  // lock-step points passed so far (the same sequence in every cluster of the full waves)
      for (int t = dom_first; t < dom_tiles; t += dom_stride, ++tile_no) {
        const TileCoord tc = tile_coord(t / ksplits, dom_nm, num_n, dims.group_m, dom_m0);
        if (!epi_tile_enabled<Epi>(ep, tc.m_blk)) continue;
        const int m0 = tc.m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BM;
        const int n0 = tc.n_blk * BN + static_cast<int>(cta_rank) * Cfg::B_ROWS;
        const int kb0 = (t % ksplits) * dims.kb_per_split;
        const int kb1 = min(num_kb, kb0 + dims.kb_per_split);
        // every cluster announces its next lock-step point and waits (at most ~40 us) for the others to get there; a
        // cluster that ever times out (a straggler exists: SMs shared with another kernel) stops waiting for good,
        // so the worst case costs one time-out per cluster and launch
        auto lock_step = [&]() {
          atomicAdd(dims.sync_ctr, 1u);
          ++sync_idx;
          if (sync_wait) {
            const uint32_t want = sync_idx * static_cast<uint32_t>(num_clusters);
            const uint64_t t_start = globaltimer_ns();
            while (*reinterpret_cast<volatile uint32_t*>(dims.sync_ctr) < want) {
              if (globaltimer_ns() - t_start > 40000ull) {
                sync_wait = false;
                break;
              }
            }
          }
        };
        const bool sync_tile = dims.sync_ctr != nullptr && is_leader && tile_no < dims.sync_tiles;
        if (sync_tile && tile_no > 0 && tile_no % dims.sync_stride == 0) lock_step();
        for (int kb = kb0; kb < kb1; ++kb) {
          uint8_t* sa = stage_base + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = kb * BK;
          if (sync_tile && dims.sync_kb > 0 && kb > kb0 && (kb - kb0) % dims.sync_kb == 0) lock_step();
          if (prefetched > 0) {
            // stage already armed and its A tile in flight
            --prefetched;
            load_b(sb, s, n0, k0);
          } else {
            mbar_wait(&empty_bar[s], ph ^ 1u, SITE_PRODUCER_EMPTY);
            if (is_leader) mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES * CG);
            load_a(sa, s, m0, k0);
            load_b(sb, s, n0, k0);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
      trace_stamp(dims.trace_id, 3);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (is_leader) {
      constexpr uint32_t idesc = make_idesc_bf16(Cfg::TILE_M, BN, Cfg::A_MN, Cfg::B_MN);
      // K-major SW128: 8-row groups are 1024 B apart (SBO); LBO unused.  Advance K by 32 B per UMMA_K.
      // MN-major SW128: 64-element MN chunks are BK*128 B apart (LBO), 8-k groups 1024 B apart (SBO);
      //                 advance K by 16 rows * 128 B = 2048 B per UMMA_K.
      constexpr uint32_t A_LBO = Cfg::A_MN ? BK * 128 : 0, A_SBO = 1024, A_KSTEP = Cfg::A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = Cfg::B_MN ? BK * 128 : 0, B_SBO = 1024, B_KSTEP = Cfg::B_MN ? 2048 : 32;
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = dom_first; t < dom_tiles; t += dom_stride) {
        if constexpr (epi_has_tile_mask<Epi>::value) {
          if (!Epi::tile_enabled(ep, tile_coord(t / ksplits, dom_nm, num_n, dims.group_m, dom_m0).m_blk)) continue;
        }
        const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
        const uint32_t aph = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
        ++it;  // counts the tiles this cluster really computes (masked tiles use no accumulator stage)
        mbar_wait(&tmem_empty_bar[as], aph ^ 1u, SITE_MMA_TMEM_EMPTY);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        const int kb0 = (t % ksplits) * dims.kb_per_split;
        const int kb1 = min(num_kb, kb0 + dims.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph, SITE_MMA_FULL);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(stage_base + s * Cfg::STAGE_BYTES);
            const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / Cfg::UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, A_SBO);
              const uint64_t bdesc = make_smem_desc_sw128(sb + k * B_KSTEP, B_LBO, B_SBO);
              umma_bf16<CG>(d_tmem, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            }
            // free the smem slot once these MMAs have drained it; on the last k-block also publish the tile
            if constexpr (CG == 1) {
              umma_commit(&empty_bar[s]);
              if (kb == kb1 - 1) umma_commit(&tmem_full_bar[as]);
            } else {
              umma_commit_2sm(&empty_bar[s], 0x3);
              if (kb == kb1 - 1) umma_commit_2sm(&tmem_full_bar[as], 0x3);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps =====================
    pdl_wait();                        // the epilogue reads / writes global memory
    const int q = warp & 3;            // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;  // which half of the tile's columns (EPI_SPLIT == 2)
    // cluster split-K parks the partial tile in the (by then drained) operand ring
    uint8_t* epi_smem = Cfg::CLUSTER_SPLIT ? stage_base : smem + STAGES * Cfg::STAGE_BYTES + 256;
    uint32_t leader_tmem_empty_addr[ACC_STAGES];
#pragma unroll
    for (int a = 0; a < ACC_STAGES; ++a) {
      const uint32_t local = smem_u32(&tmem_empty_bar[a]);
      leader_tmem_empty_addr[a] = (CG == 2) ? mapa_shared(local, 0) : local;
    }
    int it = 0;
    for (int t = dom_first; t < dom_tiles; t += dom_stride) {
      const TileCoord tc = tile_coord(t / ksplits, dom_nm, num_n, dims.group_m, dom_m0);
      if (!epi_tile_enabled<Epi>(ep, tc.m_blk)) continue;
      const int ks = t % ksplits;
      const int as = (ACC_STAGES == 2) ? (it & 1) : 0;
      const uint32_t aph = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
      ++it;
      const int row = tc.m_blk * Cfg::TILE_M + static_cast<int>(cta_rank) * BM + q * 32 + lane;
      const int n0 = tc.n_blk * BN;
      mbar_wait(&tmem_full_bar[as], aph, SITE_EPI_TMEM_FULL);
      tc_fence_after();
      if (warp == 2 && lane == 0) trace_stamp(dims.trace_id, 4);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      typename Epi::State st;
      Epi::begin(ep, st, row, n0, ks, dims, epi_smem);
      // This warp's share of the tile: NC chunks of 32 columns starting at chunk c0.  TMEM loads are
      // software-pipelined: the load of chunk c+1 is in flight while chunk c is processed.
      constexpr int NC = BN / 32 / Cfg::EPI_SPLIT;
      const int c0 = half * NC;
      const bool full = (n0 + BN <= dims.N);
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 * 32), ra);
      tmem_ld_wait(ra);
#pragma unroll 1
      for (int c = 0; c < NC; c += 2) {
        if (c + 1 < NC) tmem_ld_32x32(taddr + static_cast<uint32_t>((c0 + c + 1) * 32), rb);
        {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
          if (full) Epi::template chunk<true>(ep, st, row, n0 + (c0 + c) * 32, v, dims, epi_smem);
          else Epi::template chunk<false>(ep, st, row, n0 + (c0 + c) * 32, v, dims, epi_smem);
        }
        if (c + 1 < NC) {
          tmem_ld_wait(rb);
          if (c + 2 < NC) tmem_ld_32x32(taddr + static_cast<uint32_t>((c0 + c + 2) * 32), ra);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rb[j]);
          if (full) Epi::template chunk<true>(ep, st, row, n0 + (c0 + c + 1) * 32, v, dims, epi_smem);
          else Epi::template chunk<false>(ep, st, row, n0 + (c0 + c + 1) * 32, v, dims, epi_smem);
          if (c + 2 < NC) tmem_ld_wait(ra);
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tmem_empty_bar[as]);
        else mbar_arrive_cluster(leader_tmem_empty_addr[as]);
      }
      Epi::end(ep, st, row, n0, tc.n_blk * Cfg::EPI_SPLIT + half, dims, epi_smem);
      if (warp == 2 && lane == 0) trace_stamp(dims.trace_id, 5);
    }
  }

  // ===================== teardown =====================
  __syncwarp();  // re-converge the single-thread roles before the block-wide barrier
  tc_fence_before();
  if constexpr (Cfg::CLUSTER_SPLIT) {
    // every CTA's partial tile is in its shared memory: combine them in CTA 0, keep the others alive until done
    cluster_sync();
    if (threadIdx.x == 64) trace_stamp(dims.trace_id, 6);
    if (warp >= 2) {
      const int t = dom_first;  // one work item per CTA
      if (t < dom_tiles) {
        const TileCoord tc = tile_coord(t / ksplits, dom_nm, num_n, dims.group_m, dom_m0);
        const int row = tc.m_blk * Cfg::TILE_M + (warp & 3) * 32 + lane;
        Epi::cluster_finalize(ep, row, tc.n_blk * BN, ksplits, static_cast<int>(cluster_ctarank()), dims, stage_base);
      }
    }
    if (threadIdx.x == 64) trace_stamp(dims.trace_id, 7);
    cluster_sync();
  } else if constexpr (CG == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols] (cols contiguous, row pitch `ld` elements) viewed through a
// {64 cols x box_rows} box with 128-byte swizzle.  Out-of-bounds elements read as zero.
inline int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// 3-D bf16 tensor [segments, pitch, cols] (cols contiguous) viewed through {64 cols x 64 rows x 1 segment} boxes
inline int make_tmap_bf16_seg(CUtensorMap* map, const void* base, uint64_t segments, uint64_t pitch_rows, uint64_t cols,
                              uint64_t ld) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (enc == nullptr) return -1;
  cuuint64_t gdim[3] = {cols, pitch_rows, segments};
  cuuint64_t gstr[2] = {ld * 2, ld * 2 * pitch_rows};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// describes a row-segmented operand for launch_gemm (seg_rows == 0: plain 2-D operand)
struct SegOperand {
  int seg_rows = 0;   // logical rows per segment (multiple of 64)
  int seg_pitch = 0;  // physical rows per segment
  int seg_off = 0;    // first logical row inside a segment
  int segments = 0;
};

// split-K plan: at most `want` splits, every split non-empty
inline void gemm_split_plan(int num_kb, int want, int* k_splits, int* kb_per_split) {
  int ks = want < 1 ? 1 : (want > num_kb ? num_kb : want);
  const int per = (num_kb + ks - 1) / ks;
  ks = (num_kb + per - 1) / per;
  *k_splits = ks;
  *kb_per_split = per;
}

// a / b: global pointers.  K-major operand: [rows, K] with pitch ld; MN-major operand: [K, rows] with pitch ld.
template <class Cfg, class Epi>
int launch_gemm(const void* a, int64_t lda, const void* b, int64_t ldb, int M, int N, int K, int group_m,
                const typename Epi::Params& ep, int num_sms, cudaStream_t stream, int k_splits = 1,
                bool pdl = false, SegOperand a_seg = SegOperand(), SegOperand b_seg = SegOperand(), int trace_id = 0,
                uint32_t* sync_ctr = nullptr, int a_evict = 0, int b_evict = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap ta, tb;
  int rc;
  if (a_seg.seg_rows != 0) {
    // K-major A whose M rows are segmented
    if (Cfg::A_MN || (a_seg.seg_rows % 64) != 0) return -100;
    rc = make_tmap_bf16_seg(&ta, a, a_seg.segments, a_seg.seg_pitch, K, lda);
  } else if constexpr (!Cfg::A_MN) rc = make_tmap_bf16_2d(&ta, a, M, K, lda, Cfg::BM);
  else rc = make_tmap_bf16_2d(&ta, a, K, M, lda, Cfg::BK);
  if (rc != 0) return rc;
  if (b_seg.seg_rows != 0) {
    // MN-major B whose K rows are segmented
    if (!Cfg::B_MN || (b_seg.seg_rows % 64) != 0) return -100;
    rc = make_tmap_bf16_seg(&tb, b, b_seg.segments, b_seg.seg_pitch, N, ldb);
  } else if constexpr (!Cfg::B_MN) rc = make_tmap_bf16_2d(&tb, b, N, K, ldb, Cfg::B_ROWS);
  else rc = make_tmap_bf16_2d(&tb, b, K, N, ldb, Cfg::BK);
  if (rc != 0) return rc;

  auto kern = gemm_kernel<Cfg, Epi>;
  // function attributes are per device: one bit per device ordinal and instantiation
  static std::atomic<uint64_t> attr_set{0};
  int dev = 0;
  if (func_attrs_needed(attr_set, &dev)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return -3;
    // whole L1/shared array as shared memory: lets two kernels of a dependent-launch chain share an SM
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    func_attrs_mark(attr_set, dev);
  }
  GemmDims dims;
  dims.M = M;
  dims.N = N;
  dims.K = K;
  dims.group_m = group_m > 0 ? group_m : 8;
  const int num_kb = (K + Cfg::BK - 1) / Cfg::BK;
  gemm_split_plan(num_kb, k_splits, &dims.k_splits, &dims.kb_per_split);
  dims.a_seg_rows = a_seg.seg_rows;
  dims.a_seg_off = a_seg.seg_off;
  dims.b_seg_rows = b_seg.seg_rows;
  dims.b_seg_off = b_seg.seg_off;
  dims.trace_id = trace_id;
  dims.a_hint = a_evict == 1 ? kEvictFirst : a_evict == 2 ? kEvictLast : kEvictNormal;
  dims.b_hint = b_evict == 1 ? kEvictFirst : b_evict == 2 ? kEvictLast : kEvictNormal;
  const int num_m = (M + Cfg::TILE_M - 1) / Cfg::TILE_M;
  const int num_n = (N + Cfg::BN - 1) / Cfg::BN;
  const int num_tiles = num_m * num_n * dims.k_splits;
  int clusters = num_sms / Cfg::CG;
  if (clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;
  if constexpr (Cfg::CLUSTER_SPLIT) {
    // exactly one (tile, k-split) item per CTA, the splits of a tile forming one cluster
    if (num_tiles > num_sms || dims.k_splits > 8) return -100;
    clusters = num_tiles;
  }

  dims.sync_ctr = nullptr;
  dims.sync_tiles = 0;
  dims.sync_stride = 1;
  dims.sync_kb = 0;
  if (sync_ctr != nullptr && !Cfg::CLUSTER_SPLIT && clusters >= 2 && num_tiles / clusters >= 2) {
    if (cudaMemsetAsync(sync_ctr, 0, sizeof(uint32_t), stream) != cudaSuccess) return -4;
    dims.sync_ctr = sync_ctr;
    dims.sync_tiles = num_tiles / clusters;  // the full waves; the ragged tail runs free
    constexpr int stride_kb = 128;  // short tiles (K = 4096: 64 k-blocks) lock-step every second tile: -1.4 % step time
    const int kb_per_tile = (num_kb + dims.k_splits - 1) / dims.k_splits;
    dims.sync_stride = stride_kb > 0 ? (stride_kb + kb_per_tile - 1) / kb_per_tile : 1;
    if (dims.sync_stride < 1) dims.sync_stride = 1;
    // Tiles much longer than the 1152 k-blocks of configs[1]'s weight-gradient GEMMs (K = 73728 rows) are cut into
    // segments of about that length: at 256 pairs per GPU (K = 294912) wgrad W2 read 122 GB instead of 4 x 12.8 GB
    // with lock-step points at the tile starts only (profiles/r02_launches_256pairs_ncu.csv).
    constexpr int seg_kb = 1152;
    if (seg_kb > 0 && kb_per_tile > seg_kb + seg_kb / 2) {
      const int nseg = (kb_per_tile + seg_kb / 2) / seg_kb;
      dims.sync_kb = (kb_per_tile + nseg - 1) / nseg;
    }
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * Cfg::CG, 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = Cfg::CLUSTER_SPLIT ? dims.k_splits : Cfg::CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (pdl) {
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, dims, ep);
  return e == cudaSuccess ? 0 : -4;
}

}  // namespace ospo
