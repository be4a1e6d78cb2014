// One persistent kernel for a whole CFG decode step of the image-token head:
//
//   phase 1   act[n, E]   = bf16(gelu(bf16(h W1^T + b1)))           (swap-AB, cluster split-K through DSMEM)
//   -------   device-wide flag: every slice of act is in global memory
//   phase 2   D[V, n]     = W2 act^T + b2  -> CFG merge, softmax weights, segment sums (EpiCfgFused)
//
// reference: gen_head(hidden_states[:, -1, :]) + CFG merge + softmax, ospo/wrapper/image_generation.py:156-162.
//
// Why one kernel: the step is a pure weight stream (W1 then W2, 168 MB for the 7B head) with a true dependency in
// the middle.  As two kernels the W2 stream cannot start before the last CTA of the first kernel has left its SM;
// here the TMA producer of every CTA runs straight on from its W1 slab into its W2 slab -- the A (weight) halves of
// up to a ring-full of phase-2 stages are requested while the cluster reduction, the activation and the flag are
// still in progress, and only the tiny B (activation) halves wait for the flag.
//
// Grid: G CTAs in clusters of `ks` (the phase-1 k-splits of one 128-row slab of W1); one CTA per SM, all resident
// (checked on the host with the occupancy API -- the flag wait needs every phase-1 CTA to be running).
#include "epilogues.cuh"
#include "launchers.h"

namespace ospo {

void set_watchdog_merged(uint32_t* dev_ptr) { cudaMemcpyToSymbol(g_watchdog_buf, &dev_ptr, sizeof(dev_ptr)); }
void set_trace_merged(unsigned long long* dev_ptr) { cudaMemcpyToSymbol(g_trace_buf, &dev_ptr, sizeof(dev_ptr)); }

namespace {

constexpr int kBM = 128, kBN = 32, kBK = 64, kUmmaK = 16;
constexpr int kStages = 10;
constexpr int kABytes = kBM * kBK * 2;   // 16 KB weight tile
constexpr int kBBytes = kBN * kBK * 2;   // 4 KB activation tile
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kPartBytes = kBN * kBM * 4;  // gather buffer for the cluster's partials [split][128 rows][32 / splits]
constexpr int kEpiBytes = 1024;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256 + kEpiBytes + kPartBytes;
constexpr int kThreads = 224;  // warp 0 weight producer, 1 MMA issuer, 2-5 epilogue, 6 activation producer
constexpr int kTmemCols = 64;  // two 32-column accumulators

enum : uint32_t {
  SITE_M_PRODUCER_EMPTY = 11,
  SITE_M_MMA_FULL = 12,
  SITE_M_MMA_TMEM_EMPTY = 13,
  SITE_M_EPI_TMEM_FULL = 14,
  SITE_M_PART_READY = 15,
  SITE_M_FLAG = 16,
  SITE_M_FLAG_STATE = 17,
};

struct MergedDims {
  int n, H, E, V;
  int ks;            // cluster size = phase-1 k-splits (1, 2, 4, 8)
  int kb_per_split;  // phase-1 k-blocks per split
  int num_m1;        // 128-row slabs of W1
  int num_m2;        // 128-row slabs of W2
  const float* b1;
  __nv_bfloat16* act;   // [n, E]
  uint32_t* flag;       // [0] phase-1 arrivals, [1] CTA exits; both zero between launches
  const uint8_t* w1p;   // optional pre-packed weights (ospo_head_pack_weight): tile (slab, k-block) = one contiguous
  const uint8_t* w2p;   //   16 KB block that already is the swizzled shared-memory image; null = tensor-map loads
  int linear_only;      // 1: phase 1 alone -- out[n, E] = bf16(act_fn(x W^T + b)), no flag, no phase 2
  int gelu;             // phase-1 activation: 1 = exact-erf GELU (gen_head), 0 = identity (gen_aligner's last Linear)
  uint64_t w_hint;      // L2 cache policy of the weight loads: evict-first -- every weight byte is read exactly once
                        // per step, so it should not displace the activations or the tiles prefetched into L2
                        // (same box: 33.7 -> 31.7 us per step)
  uint64_t pf_hint;     // L2 cache policy of the run-ahead prefetch
  int l2_ahead;         // W2 tiles per CTA requested into L2 while the activation flag is closed (0 = off)
  int next_slabs;       // > 0: the weight of the NEXT kernel of the chain (gen_aligner's D x D Linear, tmap_next) is
  int next_kb;          //   requested into L2 -- next_slabs x next_kb tiles of 16 KB spread over the CTAs -- right behind
                        //   this CTA's last W2 tile, so HBM keeps streaming through the epilogue, the finish kernel and
                        //   the kernel boundary, and the Linear then reads its weight at L2 speed
  unsigned long long* trace;  // timeline buffer [8][160][8] or null (a kernel parameter: stamps cost one store)
};

__device__ __forceinline__ void stamp(unsigned long long* buf, int row, int slot) {
  if (buf != nullptr && blockIdx.x < kTraceCtas)
    buf[(static_cast<size_t>(row) * kTraceCtas + blockIdx.x) * kTraceSlots + slot] = globaltimer_ns();
}

// contiguous 16 KB weight tile -> shared memory (no tensor map: one request instead of 128 row requests)
__device__ __forceinline__ void bulk_load_tile(void* smem_dst, const uint8_t* gsrc, uint64_t* bar, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(kABytes), "r"(smem_u32(bar)), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void bulk_prefetch_tile(const uint8_t* gsrc, uint64_t hint) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(gsrc), "r"(kABytes), "l"(hint)
               : "memory");
}

template <int MODE, bool TDIV, bool WBF, bool GREEDY>
__global__ void __launch_bounds__(kThreads, 1)
decode_merged_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_h,
                     const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_act,
                     const __grid_constant__ CUtensorMap tmap_next, MergedDims d,
                     typename EpiCfgFused<MODE, TDIV, WBF, GREEDY>::Params ep) {
  using Epi = EpiCfgFused<MODE, TDIV, WBF, GREEDY>;
  constexpr int tr2 = 2;
  const int tr1 = d.linear_only ? 6 : 1, trp = d.linear_only ? 7 : 3;  // timeline rows (ospo_head_trace)
  if (d.linear_only) pdl_launch_dependents();  // nothing below needs every CTA resident: the successor may queue up
  if (threadIdx.x == 0) stamp(d.trace, tr1, 0);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* part_ready_bar = tmem_empty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(part_ready_bar + 1);
  uint8_t* epi_smem = smem + kStages * kStageBytes + 256;
  float* part = reinterpret_cast<float*>(epi_smem + kEpiBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x);
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = static_cast<int>(blockIdx.x) / d.ks;

  // phase-1 work of this CTA: k-blocks [kb0, kb1) of W1 slab `cluster_id`
  const int num_kb1 = (d.H + kBK - 1) / kBK;
  const bool has1 = cluster_id < d.num_m1;
  const int kb0 = rank * d.kb_per_split;
  const int n1 = has1 ? max(0, min(num_kb1, kb0 + d.kb_per_split) - kb0) : 0;
  // phase-2 work: W2 slabs blockIdx.x, blockIdx.x + G, ...
  const int num_kb2 = (d.E + kBK - 1) / kBK;
  const int tiles2 = (static_cast<int>(blockIdx.x) < d.num_m2) ? (d.num_m2 - 1 - static_cast<int>(blockIdx.x)) / G + 1 : 0;
  const int total = n1 + tiles2 * num_kb2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_w2);
    tma_prefetch_desc(&tmap_act);
    if (d.next_slabs > 0) tma_prefetch_desc(&tmap_next);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    mbar_init(part_ready_bar, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(part_ready_bar, kPartBytes);  // the cluster pushes one full 32 x 128 fp32 tile into this CTA
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<1>(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  cluster_sync();  // peers' barriers are initialised before anyone arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  if (threadIdx.x == 0) stamp(d.trace, trp, 0);

  // The CTA's k-blocks form one sequence: n1 blocks of W1 (B = h, needs the predecessor kernel), then the blocks
  // of its W2 slabs (B = act, needs the device-wide flag).  The weight (A) halves and the activation (B) halves are
  // issued by different warps: one thread cannot issue two TMA boxes per 16 KB fast enough to keep HBM saturated,
  // and the A stream must not stop while a B dependency is open -- it simply runs up to a ring-full ahead.
  if (warp == 0) {
    // ===================== weight (A) producer =====================
    if (elect_one()) {
      int slot = 0, use = 0, k = kb0 * kBK, left = n1, row = cluster_id * kBM;
      bool p1 = n1 > 0;
      if (!p1) {
        k = 0;
        left = num_kb2;
        row = static_cast<int>(blockIdx.x) * kBM;
      }
      for (int i = 0; i < total; ++i) {
        if (use > 0) mbar_wait(&empty_bar[slot], static_cast<uint32_t>(use - 1) & 1u, SITE_M_PRODUCER_EMPTY);
        mbar_arrive_expect_tx(&full_bar[slot], kStageBytes);  // covers the B half issued by warp 6
        const uint8_t* packed = p1 ? d.w1p : d.w2p;
        if (packed != nullptr) {
          const size_t tile = static_cast<size_t>(row >> 7) * (p1 ? num_kb1 : num_kb2) + (k >> 6);
          bulk_load_tile(stage_base + slot * kStageBytes, packed + tile * kABytes, &full_bar[slot], d.w_hint);
        } else {
          tma_load_2d(stage_base + slot * kStageBytes, p1 ? &tmap_w1 : &tmap_w2, &full_bar[slot], k, row, d.w_hint);
        }
        k += kBK;
        if (--left == 0) {  // next item: a W2 slab
          row = p1 ? static_cast<int>(blockIdx.x) * kBM : row + G * kBM;
          p1 = false;
          k = 0;
          left = num_kb2;
        }
        if (++slot == kStages) {
          slot = 0;
          ++use;
        }
        if (i + 1 == min(total, kStages)) stamp(d.trace, tr1, 1);
        if (i == kStages - 1 && n1 > kStages) {
          // the W1 tiles that do not fit the ring are requested into L2 while the predecessor kernel still runs, so
          // that they arrive at L2 latency once the first MMAs free their slots (-0.25 us per step)
#pragma unroll 1
          for (int j = 0; j < n1 - kStages; ++j) {
            if (d.w1p != nullptr)
              bulk_prefetch_tile(d.w1p + (static_cast<size_t>(row >> 7) * num_kb1 + ((k + j * kBK) >> 6)) * kABytes, d.pf_hint);
            else
              tma_prefetch_l2_2d(&tmap_w1, k + j * kBK, row);
          }
        }
        if (i == n1 + kStages - 1 && d.l2_ahead > 0) {
          // The ring now holds only phase-2 tiles and this thread is about to sleep until the activation flag opens
          // and the MMA warp starts freeing slots.  HBM would idle through that wait: ask for the next l2_ahead
          // tiles of the slab to be brought into L2 meanwhile (cp.async.bulk.prefetch: no destination, no barrier).
          // Measured alternatives, all slower: prefetching every tile ahead of its ring load (the doubled L2
          // traffic costs more than it hides), and prefetching before the predecessor-kernel wait as well.
          const int nb = min(d.l2_ahead, left);
          if (d.w2p != nullptr) {
            const uint8_t* next = d.w2p + (static_cast<size_t>(row >> 7) * num_kb2 + (k >> 6)) * kABytes;
#pragma unroll 1
            for (int j = 0; j < nb; ++j) bulk_prefetch_tile(next + static_cast<size_t>(j) * kABytes, d.pf_hint);
          } else {
#pragma unroll 1
            for (int j = 0; j < nb; ++j) tma_prefetch_l2_2d(&tmap_w2, k + j * kBK, row);
          }
        }
      }
      if (d.next_slabs > 0) {
        // every weight tile of this step has been requested: queue the next kernel's weight behind them
        const int tiles = d.next_slabs * d.next_kb;
#pragma unroll 1
        for (int t = static_cast<int>(blockIdx.x); t < tiles; t += G)
          tma_prefetch_l2_2d(&tmap_next, (t % d.next_kb) * kBK, (t / d.next_kb) * kBM);
      }
      stamp(d.trace, tr2, 3);
    }
  } else if (warp == 6) {
    // ===================== activation (B) producer =====================
    if (elect_one()) {
      int slot = 0, use = 0, k = kb0 * kBK, left = n1;
      bool p1 = n1 > 0;
      if (!p1) {
        k = 0;
        left = num_kb2;
      }
      pdl_wait();  // h comes from the predecessor kernel
      stamp(d.trace, tr1, 2);
      bool flag_seen = false;
      auto wait_flag = [&]() {
        // every phase-1 CTA has published its slice of act (and is therefore done reading its peers' partials)
        const uint32_t want = static_cast<uint32_t>(d.num_m1 * d.ks);
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (ld_acquire_gpu(d.flag) < want) {
          if (t0 == 0) t0 = globaltimer_ns();
          if (((++spins) & 0xFFu) == 0u && globaltimer_ns() - t0 > OSPO_WATCHDOG_NS)
            watchdog_fire(SITE_M_FLAG, ld_acquire_gpu(d.flag), want);
        }
        fence_proxy_async_all();  // act was written by ordinary stores; it is read by TMA
        pdl_launch_dependents();  // every CTA is resident and past its dependency: the finish kernel may queue up
        // This CTA never looks at the flag again.  The last CTA to get here re-arms both words for the next launch
        // (every phase-1 arrival has happened -- the flag is open -- and every other CTA has finished polling).
        if (atomicAdd(d.flag + 1, 1u) == static_cast<uint32_t>(G) - 1u) {
          d.flag[0] = 0u;
          d.flag[1] = 0u;
        }
        stamp(d.trace, tr1, 7);
        stamp(d.trace, tr2, 0);
        flag_seen = true;
      };
      for (int i = 0; i < total; ++i) {
        if (!p1 && !flag_seen) wait_flag();
        // the slot's previous contents must have been consumed (same condition the A producer waits for)
        if (use > 0) mbar_wait(&empty_bar[slot], static_cast<uint32_t>(use - 1) & 1u, SITE_M_PRODUCER_EMPTY);
        uint8_t* sb = stage_base + slot * kStageBytes + kABytes;
        if (p1) tma_load_2d(sb, &tmap_h, &full_bar[slot], k, 0, kEvictNormal);
        else tma_load_2d(sb, &tmap_act, &full_bar[slot], k, 0, kEvictLast);
        k += kBK;
        if (--left == 0) {
          if (p1) stamp(d.trace, tr1, 3);
          p1 = false;
          k = 0;
          left = num_kb2;
        }
        if (++slot == kStages) {
          slot = 0;
          ++use;
        }
      }
      if (!flag_seen && !d.linear_only) wait_flag();  // a CTA without phase-2 work still helps re-arm the flag words
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One elected lane runs the whole sequence (the other lanes wait at the __syncwarp below).  Per 16 KB stage the
    // chain  try_wait -> 4 x tcgen05.mma -> commit  is serial in this thread, so the barrier of stage i + 1 is tested
    // BEFORE the MMAs of stage i are issued: the test's latency (~90 cycles when the phase is already complete) runs
    // under them, and the two operand descriptors of a stage are one OR + one add each.  Measured: 31.2 -> 30.8 us
    // per step (7B), 20.0 -> 19.8 us (1B) -- the pace of phase 2 is set by the memory system, not by this loop.
    constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN, false, false);
    if (elect_one()) {
      int i = 0;  // position in the k-block sequence
      const int items = (has1 ? 1 : 0) + tiles2;
      const uint64_t desc_hi = make_smem_desc_sw128(0, 0, 1024);  // everything but the start address
      bool ready = false;  // full_bar of k-block i already seen complete
      for (int it = 0; it < items; ++it) {
        const int as = it & 1;
        const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[as], aph ^ 1u, SITE_M_MMA_TMEM_EMPTY);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kBN);
        const int nkb = (has1 && it == 0) ? n1 : num_kb2;
        for (int kb = 0; kb < nkb; ++kb, ++i) {
          const int slot = i % kStages;
          if (!ready) mbar_wait(&full_bar[slot], static_cast<uint32_t>(i / kStages) & 1u, SITE_M_MMA_FULL);
          // look ahead (the next k-block of this CTA's sequence, also across an accumulator boundary)
          const int nslot = (i + 1) % kStages;
          ready = (i + 1 < total) && mbar_try_wait(&full_bar[nslot], static_cast<uint32_t>((i + 1) / kStages) & 1u);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + slot * kStageBytes);
          const uint64_t adesc = desc_hi | static_cast<uint64_t>((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = desc_hi | static_cast<uint64_t>(((sa + kABytes) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)  // 32 bytes along K per step: + 2 in the (address >> 4) field
            umma_bf16<1>(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb != 0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[slot]);
          if (kb == nkb - 1) umma_commit(&tmem_full_bar[as]);
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===================== epilogue warps =====================
    pdl_wait();  // the buffers written below are still being read by the previous step's finish kernel
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside a 128-row slab = TMEM lane
    int it = 0;
    if (has1) {
      // ---- phase 1: exchange the partial tiles inside the cluster, add them, publish the activations ----
      const int e = cluster_id * kBM + r;
      const float b = (e < d.E) ? __ldg(d.b1 + e) : 0.0f;  // in flight while the accumulator is still being built
      mbar_wait(&tmem_full_bar[0], 0u, SITE_M_EPI_TMEM_FULL);
      tc_fence_after();
      if (threadIdx.x == 64) stamp(d.trace, tr1, 4);
      uint32_t ra[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16), ra);
      tmem_ld_wait(ra);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[0]);
      // Push model: CTA j of the cluster owns columns [j * per, (j + 1) * per).  Every thread sends its row's values
      // for those columns straight from registers into the owner's gather buffer [split][row][per] with st.async;
      // the bytes complete on the owner's mbarrier, so no fence, no remote arrive and no remote load is needed.
      const int per = kBN / d.ks;                       // 32, 16, 8 or 4
      const int per_shift = 31 - __clz(per);
      const uint32_t gather_local = smem_u32(part) + static_cast<uint32_t>((rank * kBM + r) * per) * 4u;
      const uint32_t bar_local = smem_u32(part_ready_bar);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t owner = static_cast<uint32_t>((4 * g) >> per_shift);
        const uint32_t within = static_cast<uint32_t>((4 * g) & (per - 1));
        st_async_v4(mapa_shared(gather_local + within * 4u, owner), ra[4 * g], ra[4 * g + 1], ra[4 * g + 2], ra[4 * g + 3],
                    mapa_shared(bar_local, owner));
      }
      if (threadIdx.x == 64) stamp(d.trace, tr1, 5);
      mbar_wait(part_ready_bar, 0u, SITE_M_PART_READY);
      if (threadIdx.x == 64) stamp(d.trace, trp, 1);
      // this CTA's share of the 32 columns, partials added in split order (deterministic)
      const int c_lo = rank * per;
      if (e < d.E) {
        for (int c4 = 0; c4 < per; c4 += 4) {
          float4 acc = *reinterpret_cast<const float4*>(part + r * per + c4);
          for (int k = 1; k < d.ks; ++k) {
            const float4 o = *reinterpret_cast<const float4*>(part + (k * kBM + r) * per + c4);
            acc.x += o.x;
            acc.y += o.y;
            acc.z += o.z;
            acc.w += o.w;
          }
          const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = c_lo + c4 + t;
            if (c < d.n) {
              const float x = bf16_round(a4[t] + b);
              d.act[static_cast<int64_t>(c) * d.E + e] = __float2bfloat16_rn(d.gelu ? gelu_erf(x) : x);
            }
          }
        }
      }
      if (threadIdx.x == 64) stamp(d.trace, trp, 2);
      named_bar_sync(1, 128);
      if (threadIdx.x == 64 && !d.linear_only) {
        // release at device scope is cumulative over the stores the barrier above has ordered before it
        fence_proxy_async_all();
        const uint32_t old = atom_add_release_gpu(d.flag, 1u);
        if (old >= static_cast<uint32_t>(d.num_m1 * d.ks)) watchdog_fire(SITE_M_FLAG_STATE, old, 0u);  // dirty flag words
        stamp(d.trace, tr1, 6);
      }
      it = 1;
    }
    // ---- phase 2: fused CFG epilogue on every W2 slab of this CTA ----
    GemmDims gd = {};
    gd.M = d.V;
    gd.N = d.n;
    gd.K = d.E;
    gd.trace_id = 0;
    const bool full = (d.n == kBN);
    for (int t = 0; t < tiles2; ++t, ++it) {
      const int as = it & 1;
      const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
      const int row = (static_cast<int>(blockIdx.x) + t * G) * kBM + r;
      typename Epi::State st;
      Epi::begin(ep, st, row, 0, 0, gd, epi_smem);  // bias load overlaps the wait
      mbar_wait(&tmem_full_bar[as], aph, SITE_M_EPI_TMEM_FULL);
      tc_fence_after();
      if (threadIdx.x == 64) stamp(d.trace, tr2, 4);
      uint32_t ra[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * kBN), ra);
      tmem_ld_wait(ra);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);  // the accumulator is in registers
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
      if (full) Epi::template chunk<true>(ep, st, row, 0, v, gd, epi_smem);
      else Epi::template chunk<false>(ep, st, row, 0, v, gd, epi_smem);
      Epi::end(ep, st, row, 0, 0, gd, epi_smem);
      if (threadIdx.x == 64) stamp(d.trace, tr2, 5);
    }
  }

  // ===================== teardown =====================
  __syncwarp();
  tc_fence_before();
  if (threadIdx.x == 0) stamp(d.trace, trp, 3);
  // No cluster barrier here: the partial tiles pushed into this CTA were complete before it published its
  // activations, and it pushes nothing after that -- no peer touches this CTA's shared memory any more.
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 0) {
    stamp(d.trace, trp, 4);
  }
}

// The phase-1-only launches (gen_aligner's Linear) reuse whichever instantiation the decode step ran last: the two
// alternate inside the generate loop, and running the same code keeps it in the instruction caches (a different
// instantiation for the Linear cost 5 us per step).
int g_last_variant = 1;  // MODE 0, T == 1, bf16-exact cfg_weight: the reference's defaults

template <int MODE, bool TDIV, bool WBF, bool GREEDY>
int run_linear(const LaunchCtx& c, const CUtensorMap& t_w, const CUtensorMap& t_x, const MergedDims& d) {
  using Epi = EpiCfgFused<MODE, TDIV, WBF, GREEDY>;
  typename Epi::Params p{};
  auto kern = decode_merged_kernel<MODE, TDIV, WBF, GREEDY>;
  static std::atomic<uint64_t> attr_set{0};  // per device ordinal
  int attr_dev = 0;
  if (func_attrs_needed(attr_set, &attr_dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return -3;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    func_attrs_mark(attr_set, attr_dev);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(d.num_m1 * d.ks), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = c.stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = static_cast<unsigned>(d.ks);
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (c.pdl) {
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, t_w, t_x, t_w, t_x, t_w, d, p) == cudaSuccess ? 0 : -4;
}

template <int MODE, bool TDIV, bool WBF, bool GREEDY>
int run_merged(const LaunchCtx& c, const CUtensorMap& t_w1, const CUtensorMap& t_h, const CUtensorMap& t_w2,
               const CUtensorMap& t_act, const CUtensorMap& t_next, const MergedDims& d, int G, const float* b2,
               __nv_bfloat16* logits_dump,
               float cfg_weight, float temperature, int greedy, const CfgFusedBuffers& buf) {
  using Epi = EpiCfgFused<MODE, TDIV, WBF, GREEDY>;
  typename Epi::Params p{b2, cfg_weight, temperature, logits_dump, d.V, buf, greedy, d.V};
  auto kern = decode_merged_kernel<MODE, TDIV, WBF, GREEDY>;
  g_last_variant = (GREEDY ? 8 : 0) + MODE * 4 + (TDIV ? 2 : 0) + (WBF ? 1 : 0);
  static std::atomic<uint64_t> attr_set{0};  // per device ordinal
  static int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // by cluster size (the devices of a box are identical)
  int attr_dev = 0;
  if (func_attrs_needed(attr_set, &attr_dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) return -3;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    func_attrs_mark(attr_set, attr_dev);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(G), 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = c.stream;
  cudaLaunchAttribute attrs[3];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = static_cast<unsigned>(d.ks);
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  if (max_clusters[d.ks] == 0) {
    // the flag wait needs all G CTAs running at once: ask the driver how many clusters of this size fit
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      nc = -1;
    }
    max_clusters[d.ks] = nc > 0 ? nc : -1;
  }
  if (max_clusters[d.ks] * d.ks < G) return -100;
  if (c.pdl) {
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  // cooperative launch: the driver either places all G CTAs at once or refuses the launch -- the formal form of the
  // co-residency the flag wait relies on (the occupancy query above cannot see other work on the device)
  attrs[cfg.numAttrs].id = cudaLaunchAttributeCooperative;
  attrs[cfg.numAttrs].val.cooperative = 1;
  ++cfg.numAttrs;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, t_w1, t_h, t_w2, t_act, t_next, d, p);
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    cudaGetLastError();
    return -100;  // the caller uses the two-GEMM chain
  }
  return e == cudaSuccess ? 0 : -4;
}

}  // namespace

// out[n, M] = bf16(act_fn(bf16(x W^T + b))) for n <= 32 rows: phase 1 of the kernel above on its own (cluster split-K,
// st.async exchange, separate weight / activation producers).  Clusters are independent: no flag, no co-residency.
int launch_decode_linear(const LaunchCtx& c, const __nv_bfloat16* x, const __nv_bfloat16* w, const float* b,
                         __nv_bfloat16* out, int n, int K, int M, int gelu) {
  if (n < 1 || n > kBN) return -100;
  MergedDims d;
  d.n = n;
  d.H = K;
  d.E = M;
  d.V = 0;
  d.num_m1 = (M + kBM - 1) / kBM;
  d.num_m2 = 0;
  int ks = decode_gemm1_splits(c.num_sms, K, M);
  ks = ks >= 8 ? 8 : ks >= 4 ? 4 : ks >= 2 ? 2 : 1;
  int per;
  gemm_split_plan((K + kBK - 1) / kBK, ks, &ks, &per);
  if (ks != 1 && ks != 2 && ks != 4 && ks != 8) return -100;
  d.ks = ks;
  d.kb_per_split = per;
  d.b1 = b;
  d.act = out;
  d.flag = nullptr;
  d.w1p = nullptr;
  d.w2p = nullptr;
  d.trace = c.trace ? c.trace_buf : nullptr;
  d.linear_only = 1;
  d.gelu = gelu;
  d.l2_ahead = 0;
  d.next_slabs = 0;
  d.next_kb = 0;
  d.w_hint = kEvictFirst;
  d.pf_hint = kEvictNormal;
  CUtensorMap t_w, t_x;
  int rc;
  if ((rc = make_tmap_bf16_2d(&t_w, w, M, K, K, kBM)) != 0) return rc;
  if ((rc = make_tmap_bf16_2d(&t_x, x, n, K, K, kBN)) != 0) return rc;
#define OSPO_LIN(G, M, T, W) run_linear<M, T, W, G>(c, t_w, t_x, d)
  switch (g_last_variant) {
    case 0: return OSPO_LIN(false, 0, false, false);
    case 1: return OSPO_LIN(false, 0, false, true);
    case 2: return OSPO_LIN(false, 0, true, false);
    case 3: return OSPO_LIN(false, 0, true, true);
    case 4: return OSPO_LIN(false, 1, false, false);
    case 6: return OSPO_LIN(false, 1, true, false);
    case 8: return OSPO_LIN(true, 0, false, false);
    case 9: return OSPO_LIN(true, 0, false, true);
    case 10: return OSPO_LIN(true, 0, true, false);
    case 11: return OSPO_LIN(true, 0, true, true);
    case 12: return OSPO_LIN(true, 1, false, false);
    default: return OSPO_LIN(true, 1, true, false);
  }
#undef OSPO_LIN
}

int launch_decode_merged(const LaunchCtx& c, const __nv_bfloat16* h, const __nv_bfloat16* w1, const float* b1,
                         const __nv_bfloat16* w2, const float* b2, __nv_bfloat16* act, uint32_t* flag,
                         __nv_bfloat16* logits_dump, int n, int H, int E, int V, float cfg_weight, float temperature,
                         int merge_mode, int greedy, const CfgFusedBuffers& buf, int l2_ahead, const void* w1_packed,
                         const void* w2_packed, const __nv_bfloat16* next_w, int next_rows, int next_cols,
                         int* grid_ctas) {
  if (n < 2 || n > kBN || flag == nullptr) return -100;
  MergedDims d;
  d.n = n;
  d.H = H;
  d.E = E;
  d.V = V;
  d.num_m1 = (E + kBM - 1) / kBM;
  d.num_m2 = (V + kBM - 1) / kBM;
  d.b1 = b1;
  d.act = act;
  d.flag = flag;
  d.w1p = static_cast<const uint8_t*>(w1_packed);
  d.w2p = static_cast<const uint8_t*>(w2_packed);
  d.trace = c.trace ? c.trace_buf : nullptr;
  d.l2_ahead = l2_ahead < 0 ? 0 : l2_ahead;
  d.w_hint = kEvictFirst;   // every weight byte is read once per step (evict-first: 33.7 -> 31.7 us per step in round 1)
  d.pf_hint = kEvictNormal;  // evict-first / evict-last on the run-ahead prefetch measured within 0.15 us of this
  d.linear_only = 0;
  d.gelu = 1;
  CUtensorMap t_w1, t_h, t_w2, t_act;
  int rc;
  if ((rc = make_tmap_bf16_2d(&t_w1, w1, E, H, H, kBM)) != 0) return rc;
  if ((rc = make_tmap_bf16_2d(&t_h, h, n, H, H, kBN)) != 0) return rc;
  if ((rc = make_tmap_bf16_2d(&t_w2, w2, V, E, E, kBM)) != 0) return rc;
  if ((rc = make_tmap_bf16_2d(&t_act, act, n, E, E, kBN)) != 0) return rc;
  CUtensorMap t_next = t_w2;
  d.next_slabs = 0;
  d.next_kb = 0;
  if (next_w != nullptr && next_rows > 0 && next_cols > 0 && (next_cols % 8) == 0) {
    if ((rc = make_tmap_bf16_2d(&t_next, next_w, next_rows, next_cols, next_cols, kBM)) != 0) return rc;
    d.next_slabs = (next_rows + kBM - 1) / kBM;
    d.next_kb = (next_cols + kBK - 1) / kBK;
  }
  const bool tdiv = (temperature != 1.0f);
  // phase-1 split: as many k-splits as fill the SMs (cluster sizes 8, 4, 2, 1); a cluster size the device cannot
  // keep resident G / ks times (e.g. 16 clusters of 8 CTAs that each own an SM's shared memory) is halved
  int want = decode_gemm1_splits(c.num_sms, H, E);
  want = want >= 8 ? 8 : want >= 4 ? 4 : want >= 2 ? 2 : 1;
  for (; want >= 1; want >>= 1) {
    int ks, per;
    gemm_split_plan((H + kBK - 1) / kBK, want, &ks, &per);
    if (ks != 1 && ks != 2 && ks != 4 && ks != 8) continue;
    d.ks = ks;
    d.kb_per_split = per;
    const int max_ctas = (c.num_sms / ks) * ks;
    const int need1 = d.num_m1 * ks;
    if (need1 > max_ctas) continue;
    int G = d.num_m2 < max_ctas ? ((d.num_m2 + ks - 1) / ks) * ks : max_ctas;
    if (G < need1) G = need1;
#define OSPO_RUN_MERGED(MODE, TDIV, WBF)                                                                              \
  (greedy ? run_merged<MODE, TDIV, WBF, true>(c, t_w1, t_h, t_w2, t_act, t_next, d, G, b2, logits_dump, cfg_weight,      \
                                              temperature, greedy, buf)                                                  \
          : run_merged<MODE, TDIV, WBF, false>(c, t_w1, t_h, t_w2, t_act, t_next, d, G, b2, logits_dump, cfg_weight,     \
                                               temperature, greedy, buf))
    if (merge_mode == 0 && bf16_exact(cfg_weight)) {
      rc = tdiv ? OSPO_RUN_MERGED(0, true, true) : OSPO_RUN_MERGED(0, false, true);
    } else if (merge_mode == 0) {
      rc = tdiv ? OSPO_RUN_MERGED(0, true, false) : OSPO_RUN_MERGED(0, false, false);
    } else {
      rc = tdiv ? OSPO_RUN_MERGED(1, true, false) : OSPO_RUN_MERGED(1, false, false);
    }
#undef OSPO_RUN_MERGED
    if (rc != -100) {
      if (grid_ctas != nullptr) *grid_ctas = G;
      return rc;
    }
  }
  return -100;
}

}  // namespace ospo
