// C ABI of the image-token head (declared in include/ospo_head.h).  Pure orchestration: argument
// validation, workspace carving and kernel sequencing on the caller's stream.  No allocation, no
// synchronisation, no torch types.
#include "../../include/ospo_head.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "head_kernels.cuh"
#include "optim_kernels.cuh"
#include "launchers.h"

using namespace ospo;

namespace {

// Per-device state.  The library may be called on any number of sm_100 devices from one process: everything that
// lives on (or describes) a device -- SM count, the watchdog / trace symbols of every translation unit, the flag
// pool of the decode kernel and the lock-step counters -- is kept per device ordinal and looked up with
// cudaGetDevice() on every call.
constexpr int kMaxDevices = 64;
struct DeviceState {
  bool ready = false;
  int status = OSPO_OK;
  int num_sms = 0;
  uint32_t* flag_pool = nullptr;
  int flag_next = 0;
  std::unordered_map<uint64_t, int> flag_slot;
};
DeviceState g_dev[kMaxDevices];

struct Runtime {
  bool ready = false;   // knobs parsed, host-side watchdog record allocated
  int cta_group = 2;
  int group_m = 16;
  int decode_fused = 1;  // CFG tail fused into the decode GEMM2 epilogue (0 = separate sampler pass)
  int decode_pdl = 1;    // programmatic dependent launch along the decode kernel chain
  int decode_cluster = 1;  // GEMM1 k-splits combined in-cluster through DSMEM (0 = HBM partials + finalize kernel)
  int decode_merged = 1;   // GEMM1 + GEMM2 + CFG epilogue as one persistent kernel (0 = two GEMM launches)
  int tile_sync = 1;       // wave lock-step of the persistent training GEMMs (OSPO_HEAD_TILE_SYNC)
  // per training GEMM (0 gemm1, 1 gemm2, 2 dact, 3 wgrad W2, 4 wgrad W1, 5 dgrad X): rasterisation group (0 = group_m)
  // and the L2 eviction hints of the A / B operand loads (0 normal, 1 evict-first, 2 evict-last)
  // Measured per kernel under ncu (profiles/r02_tune_sweep_ncu.csv): groups of 8 M-blocks give every GEMM but gemm2 its
  // lowest DRAM traffic (dact 12.2 vs 13.8 GB, wgrad W2 11.7 vs 13.0 GB); gemm2 wants 16 (3.9 vs 5.6 GB); eviction
  // hints of either polarity raise the traffic (a line marked evict-first leaves L2 before the sibling clusters of
  // the same wave have read it), so they stay off.
  int tune[6][3] = {{8, 0, 0}, {16, 0, 0}, {8, 0, 0}, {8, 0, 0}, {8, 0, 0}, {8, 0, 0}};
  int wgrad_splitk = 0;  // k-splits of the weight-gradient GEMMs: 0 = per shape (gemm_bwd.cu: pick_wgrad_splits), 1 / 2 forced
  int decode_l2_ahead = 16;  // merged kernel: W2 k-blocks per CTA requested into L2 while the activation flag is closed
  int decode_next_prefetch = 1;  // merged kernel: the aligner Linear's weight is requested into L2 behind the last W2 tile
  bool trace_on = false;   // ospo_head_trace installed a timeline buffer
  unsigned long long* trace_buf = nullptr;
  uint32_t* wd_host = nullptr;  // mapped, portable host record written by a kernel whose bounded wait expired
};
Runtime g_rt;
std::mutex g_mu;
std::atomic<uint64_t> g_launches{0};

// ---- optional per-kernel timing (CUDA events on the caller's stream) -------------------------------
struct Span {
  int kid;
  cudaEvent_t e0, e1;
};
bool g_profile = false;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// RAII: brackets one kernel launch with events when profiling is on
struct KernelSpan {
  cudaStream_t st;
  int idx = -1;
  KernelSpan(cudaStream_t s, int kid) : st(s) {
    if (!g_profile) return;
    Span sp{kid, take_event(), take_event()};
    cudaEventRecord(sp.e0, st);
    g_spans.push_back(sp);
    idx = static_cast<int>(g_spans.size()) - 1;
  }
  ~KernelSpan() {
    if (idx >= 0) cudaEventRecord(g_spans[idx].e1, st);
  }
};

// the calling thread's current device (its DeviceState is initialised by runtime_init)
int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

// The library reads no environment variables: the defaults below are the measured best settings, and the validation
// hooks at the end of include/ospo_head.h (ospo_head_set_*) are the only way to select another variant.
void parse_knobs_once() {
  if (g_rt.ready) return;
  g_rt.ready = true;
  if (cudaHostAlloc(reinterpret_cast<void**>(&g_rt.wd_host), 64 * sizeof(uint32_t),
                    cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
    for (int i = 0; i < 64; ++i) g_rt.wd_host[i] = 0;
  } else {
    g_rt.wd_host = nullptr;
    cudaGetLastError();
  }
}

int runtime_init() {
  std::lock_guard<std::mutex> lk(g_mu);
  const int dev = current_device();
  if (dev < 0) return OSPO_ERR_CUDA;
  DeviceState& d = g_dev[dev];
  if (d.ready) return d.status;
  d.ready = true;
  parse_knobs_once();
  int major = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10) return d.status = OSPO_ERR_ARCH;
  d.num_sms = sms;
  // this device's copies of the watchdog symbols (one per translation unit) point at the mapped host record
  uint32_t* wd_dev = nullptr;
  if (g_rt.wd_host != nullptr &&
      cudaHostGetDevicePointer(reinterpret_cast<void**>(&wd_dev), g_rt.wd_host, 0) == cudaSuccess) {
    set_watchdog_fwd(wd_dev);
    set_watchdog_bwd(wd_dev);
    set_watchdog_decode(wd_dev);
    set_watchdog_debug(wd_dev);
    set_watchdog_merged(wd_dev);
  } else {
    cudaGetLastError();
  }
  return d.status = OSPO_OK;
}

// Flag words of the merged decode kernel (its device-wide "activations are published" flag).  The only device
// memory the library owns: 16 KB, zeroed once; every kernel leaves its two words zero again.  A slot belongs to a
// stream (launches on one stream are ordered) or, under stream capture, to the graph being captured (a graph
// never overlaps itself), so two launches that may run concurrently never share a slot.  When the pool is
// exhausted the caller falls back to the two-kernel chain.
constexpr int kFlagSlots = 512, kFlagStride = 8;  // 32 bytes per slot

uint32_t* flag_words_for(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int dev = current_device();
  if (dev < 0) return nullptr;
  DeviceState& ds = g_dev[dev];
  uint32_t*& g_flag_pool = ds.flag_pool;
  int& g_flag_next = ds.flag_next;
  std::unordered_map<uint64_t, int>& g_flag_slot = ds.flag_slot;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  unsigned long long cap_id = 0;
  if (cudaStreamGetCaptureInfo(st, &cs, &cap_id) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  const bool capturing = cs == cudaStreamCaptureStatusActive;
  if (g_flag_pool == nullptr) {
    if (capturing) return nullptr;  // cannot zero the pool while a capture is open
    uint32_t* p = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&p), kFlagSlots * kFlagStride * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemset(p, 0, kFlagSlots * kFlagStride * sizeof(uint32_t)) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    g_flag_pool = p;
  }
  const uint64_t key = capturing ? (0x8000000000000000ull | cap_id) : static_cast<uint64_t>(reinterpret_cast<uintptr_t>(st));
  auto it = g_flag_slot.find(key);
  if (it == g_flag_slot.end()) {
    if (g_flag_next >= kFlagSlots) return nullptr;
    it = g_flag_slot.emplace(key, g_flag_next++).first;
  }
  return g_flag_pool + static_cast<size_t>(it->second) * kFlagStride;
}

int current_num_sms() {
  const int dev = current_device();
  return dev >= 0 ? g_dev[dev].num_sms : 0;
}

LaunchCtx make_ctx(cudaStream_t s) {
  LaunchCtx c;
  const int dev = current_device();
  c.num_sms = dev >= 0 ? g_dev[dev].num_sms : 0;
  c.cta_group = g_rt.cta_group;
  c.group_m = g_rt.group_m;
  c.wgrad_splitk = g_rt.wgrad_splitk;
  c.stream = s;
  c.pdl = false;
  c.trace = g_rt.trace_on;
  c.trace_buf = g_rt.trace_buf;
  if (g_rt.tile_sync) {
    uint32_t* w = flag_words_for(s);
    c.sync_ctr = w ? w + 4 : nullptr;  // words 0-1 belong to the decode kernel's flag
  }
  return c;
}

// the launch context of training GEMM k (see Runtime::tune)
LaunchCtx tuned(const LaunchCtx& c, int k) {
  LaunchCtx t = c;
  if (g_rt.tune[k][0] > 0) t.group_m = g_rt.tune[k][0];
  t.a_evict = g_rt.tune[k][1];
  t.b_evict = g_rt.tune[k][2];
  return t;
}

template <typename Kern, typename... Args>
cudaError_t launch_plain(Kern kern, dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// map launcher return codes to ABI status
int map_rc(int rc) {
  if (rc == 0) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return OSPO_OK;
  }
  if (rc == -1 || rc == -2) return OSPO_ERR_TENSORMAP;
  if (rc == -100) return OSPO_ERR_UNSUPPORTED;
  return OSPO_ERR_LAUNCH;
}
int check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? OSPO_OK : OSPO_ERR_LAUNCH;
}

int check_shape(const ospo_head_shape& s, bool need_seqs) {
  if (s.rows <= 0 || s.hidden <= 0 || s.embed <= 0 || s.vocab <= 0) return OSPO_ERR_BAD_SHAPE;
  if (need_seqs && s.num_seqs <= 0) return OSPO_ERR_BAD_SHAPE;
  if ((s.hidden % 8) || (s.embed % 8) || (s.vocab % 8)) return OSPO_ERR_ALIGNMENT;
  return OSPO_OK;
}
int check_weights(const ospo_head_weights& w) {
  if (!w.w1 || !w.b1 || !w.w2 || !w.b2) return OSPO_ERR_NULL;
  if (!aligned16(w.w1) || !aligned16(w.w2)) return OSPO_ERR_ALIGNMENT;
  return OSPO_OK;
}

// workspace carving -------------------------------------------------------------------------
struct Workspace {
  float2* part;          // [num_n, rows]
  float* rowsum_part;    // [num_n, rows]
  float* tgt;            // [rows]
  float* row_logit_sum;  // [rows]
  float* row_coef;       // [rows] backward: row weights w_r of the GEMM pair
  float* row_max;        // [rows] forward: row maxima of the logits (exponent reference of the repair pass)
  uint8_t* blk_mask;     // [rows / 128 + 2] forward: GEMM2 M-blocks whose spill must be recomputed (normally none);
                         // the last byte is the "any block flagged" flag
  float* seq_sum;        // [S]
  float* seq_logit_sum;  // [S]
  float* seq_count;      // [S] valid labels per sequence
  float* colsum_part;    // [rows / 512 + 1, max(V, E)] fixed-order partials of the bias gradients
  __nv_bfloat16* rows_by_e;  // [rows, E]: dpre (backward) / act (plain logits, decode)
  __nv_bfloat16* act_w;      // [rows, E]: row-weighted activations, B operand of the dW2 GEMM (training shapes only)
  float* decode_part;        // [8, rows, E] split-K partials of the decode GEMM1 (decode-sized shapes only)
  CfgFusedBuffers fused;     // outputs of the fused decode-GEMM2 epilogue (decode-sized shapes only)
  __nv_bfloat16* decode_logits;  // [rows, V] scratch logits for the unfused decode variant
  size_t total;
};

constexpr size_t kDecodeMaxRows = 512;  // 2P rows of a CFG decode step

Workspace carve(const ospo_head_shape& s, void* base) {
  Workspace w;
  const size_t rows = static_cast<size_t>(s.rows), S = static_cast<size_t>(s.num_seqs > 0 ? s.num_seqs : 1);
  const size_t num_n = static_cast<size_t>(gemm2_num_n_tiles(s.vocab));
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* q = p ? p + off : nullptr;
    off = align_up(off + bytes, 256);
    return q;
  };
  w.part = reinterpret_cast<float2*>(take(num_n * rows * sizeof(float2)));
  w.rowsum_part = reinterpret_cast<float*>(take(num_n * rows * sizeof(float)));
  w.tgt = reinterpret_cast<float*>(take(rows * sizeof(float)));
  w.row_logit_sum = reinterpret_cast<float*>(take(rows * sizeof(float)));
  w.row_coef = reinterpret_cast<float*>(take(rows * sizeof(float)));
  w.row_max = reinterpret_cast<float*>(take(rows * sizeof(float)));
  w.blk_mask = reinterpret_cast<uint8_t*>(take(rows / 128 + 2));
  w.seq_sum = reinterpret_cast<float*>(take(S * sizeof(float)));
  w.seq_logit_sum = reinterpret_cast<float*>(take(S * sizeof(float)));
  w.seq_count = reinterpret_cast<float*>(take(S * sizeof(float)));
  w.colsum_part = reinterpret_cast<float*>(
      take((rows / CS_ROWS_PER_BLOCK + 1) * static_cast<size_t>(std::max(s.vocab, s.embed)) * sizeof(float)));
  w.rows_by_e = reinterpret_cast<__nv_bfloat16*>(take(rows * static_cast<size_t>(s.embed) * 2));
  w.act_w = (rows > kDecodeMaxRows || s.num_seqs > 1)
                ? reinterpret_cast<__nv_bfloat16*>(take(rows * static_cast<size_t>(s.embed) * 2))
                : nullptr;
  w.decode_part = (rows <= kDecodeMaxRows)
                      ? reinterpret_cast<float*>(take(8 * rows * static_cast<size_t>(s.embed) * sizeof(float)))
                      : nullptr;
  w.fused = CfgFusedBuffers{nullptr, nullptr, nullptr, nullptr, nullptr};
  if (rows <= kDecodeMaxRows) {
    const size_t P = (rows + 1) / 2, V = static_cast<size_t>(s.vocab);
    w.fused.wbuf = reinterpret_cast<float*>(take(P * V * sizeof(float)));
    w.fused.seg_sum = reinterpret_cast<float*>(take(P * (V / SAMPLE_SEG + 1) * sizeof(float)));
    w.fused.tile_k = reinterpret_cast<float*>(take(P * (V / SAMPLE_TILE + 1) * sizeof(float)));
    w.fused.tile_max = reinterpret_cast<float*>(take(P * (V / SAMPLE_TILE + 1) * sizeof(float)));
    w.fused.tile_arg = reinterpret_cast<int*>(take(P * (V / SAMPLE_TILE + 1) * sizeof(int)));
    w.decode_logits = reinterpret_cast<__nv_bfloat16*>(take(rows * V * 2));
  } else {
    w.decode_logits = nullptr;
  }
  w.total = off;
  return w;
}

int check_ws(const ospo_head_shape& s, void* ws, size_t bytes, Workspace* out) {
  if (!ws) return OSPO_ERR_NULL;
  if (!aligned16(ws)) return OSPO_ERR_ALIGNMENT;
  *out = carve(s, ws);
  if (out->total > bytes) return OSPO_ERR_WORKSPACE;
  return OSPO_OK;
}

XLayout x_layout(const ospo_simpo_args* a) {
  XLayout xl;
  xl.seg_rows = a->x_seg_rows;
  xl.seg_pitch = a->x_seg_pitch;
  xl.seg_off = a->x_seg_off;
  xl.segments = a->shape.num_seqs;
  return xl;
}

// shared forward: GEMM1 -> GEMM2 (+ softmax numerator spill, LSE partials) -> merge -> [repair pass] -> one-hot
// fix-up of the spill -> per-sequence reduce
int logps_forward(const ospo_simpo_args* a, const Workspace& w, cudaStream_t st) {
  const ospo_head_shape& s = a->shape;
  const LaunchCtx c = make_ctx(st);
  __nv_bfloat16* espill = static_cast<__nv_bfloat16*>(a->logits);
  if (espill != nullptr && a->row_ref == nullptr) return OSPO_ERR_NULL;
  const int tile_m = gemm2_tile_m(c.cta_group);
  const int num_n = gemm2_num_n_tiles(s.vocab);
  uint8_t* any_flag = w.blk_mask + s.rows / 128 + 1;
  int rc;
  {
    KernelSpan ks(st, OSPO_K_GEMM1);
    if (cudaMemsetAsync(w.blk_mask, 0, static_cast<size_t>(s.rows) / 128 + 2, st) != cudaSuccess) return OSPO_ERR_CUDA;
    rc = map_rc(launch_gemm1_bias_gelu(tuned(c, 0), static_cast<const __nv_bfloat16*>(a->x),
                                       static_cast<const __nv_bfloat16*>(a->w.w1), a->w.b1,
                                       static_cast<__nv_bfloat16*>(a->pre), static_cast<__nv_bfloat16*>(a->act),
                                       s.rows, s.hidden, s.embed, x_layout(a)));
  }
  if (rc) return rc;
  {
    KernelSpan ks(st, OSPO_K_GEMM2_LSE);
    rc = map_rc(launch_gemm2_logits_exp(tuned(c, 1), static_cast<const __nv_bfloat16*>(a->act),
                                        static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2, espill, a->labels, w.part,
                                        w.rowsum_part, w.tgt, nullptr, nullptr, nullptr, s.rows, s.embed, s.vocab));
  }
  if (rc) return rc;
  KernelSpan ks(st, OSPO_K_SCALAR_STAGE);
  const int fin_grid = (s.rows + 255) / 256;
  lse_finalize_kernel<<<fin_grid, 256, 0, st>>>(w.part, w.rowsum_part, w.tgt, a->labels, s.vocab, s.rows, num_n, tile_m,
                                                1, w.blk_mask, any_flag, a->row_ref, w.row_max, a->row_lse,
                                                a->row_logps, w.row_logit_sum);
  if ((rc = check_launch())) return rc;
  // repair pass: M-blocks with a row maximum outside the representable window are recomputed against their own row
  // maxima.  For ordinary logits no block is flagged and both launches return at once.
  rc = map_rc(launch_gemm2_logits_exp(c, static_cast<const __nv_bfloat16*>(a->act),
                                      static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2, espill, a->labels, w.part,
                                      w.rowsum_part, w.tgt, w.row_max, w.blk_mask, any_flag, s.rows, s.embed, s.vocab));
  if (rc) return rc;
  lse_finalize_kernel<<<fin_grid, 256, 0, st>>>(w.part, w.rowsum_part, w.tgt, a->labels, s.vocab, s.rows, num_n, tile_m,
                                                2, w.blk_mask, any_flag, a->row_ref, w.row_max, a->row_lse,
                                                a->row_logps, w.row_logit_sum);
  if ((rc = check_launch())) return rc;
  if (espill != nullptr) {
    target_fixup_kernel<<<fin_grid, 256, 0, st>>>(espill, s.vocab, a->labels, s.vocab, s.rows, a->row_logps,
                                                  a->row_lse, a->row_ref);
    if ((rc = check_launch())) return rc;
  }
  seq_reduce_kernel<<<s.num_seqs, 256, 0, st>>>(a->row_logps, w.row_logit_sum, a->seq_offsets, a->labels, s.vocab,
                                                a->average_log_prob, a->seq_logps, w.seq_sum, w.seq_logit_sum,
                                                w.seq_count);
  return check_launch();
}

int check_simpo_common(const ospo_simpo_args* a, Workspace* w) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if ((rc = check_shape(a->shape, true))) return rc;
  if ((rc = check_weights(a->w))) return rc;
  if (!a->x || !a->labels || !a->seq_offsets) return OSPO_ERR_NULL;
  if (!aligned16(a->x)) return OSPO_ERR_ALIGNMENT;
  if (a->x_seg_rows != 0) {
    if (a->x_seg_rows < 0 || (a->x_seg_rows % 64) || a->x_seg_off < 0 ||
        a->x_seg_off + a->x_seg_rows > a->x_seg_pitch ||
        static_cast<int64_t>(a->x_seg_rows) * a->shape.num_seqs != a->shape.rows)
      return OSPO_ERR_BAD_SHAPE;
  }
  return check_ws(a->shape, a->workspace, a->workspace_bytes, w);
}

// column sums (bias gradients) in two fixed-order stages
template <bool WEIGHTED>
int launch_colsum(const __nv_bfloat16* x, int rows, int cols, const float* row_w, float scale, float* partial,
                  float* out, cudaStream_t st, const DpScatter& dp, int64_t region_off) {
  const int row_blocks = (rows + CS_ROWS_PER_BLOCK - 1) / CS_ROWS_PER_BLOCK;
  dim3 grid((cols + 1023) / 1024, row_blocks);
  colsum_partial_kernel<WEIGHTED><<<grid, 128, 0, st>>>(x, cols, rows, cols, row_w, partial);
  int rc = check_launch();
  if (rc) return rc;
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(partial, row_blocks, cols, scale, out, dp, region_off);
  return check_launch();
}

// partition of the flat gradient over the ranks of a peer-memory exchange (see ospo_dp_exchange)
struct DpPlan {
  int64_t a, b, vb, eb, shard;   // dW2 rows, dW1 rows, db2, db1 elements of one shard; their sum
};
int dp_plan(const ospo_head_shape& s, const ospo_dp_exchange* dp, DpPlan* pl, DpScatter* sc) {
  if (dp->world < 2 || dp->world > 8 || dp->rank < 0 || dp->rank >= dp->world) return OSPO_ERR_BAD_SHAPE;
  if ((s.vocab % (dp->world * 256)) || (s.embed % (dp->world * 256))) return OSPO_ERR_UNSUPPORTED;
  pl->a = static_cast<int64_t>(s.vocab / dp->world) * s.embed;
  pl->b = static_cast<int64_t>(s.embed / dp->world) * s.hidden;
  pl->vb = s.vocab / dp->world;
  pl->eb = s.embed / dp->world;
  pl->shard = pl->a + pl->b + pl->vb + pl->eb;
  sc->world = dp->world;
  sc->rank = dp->rank;
  sc->shard_elems = pl->shard;
  for (int i = 0; i < 8; ++i) {
    sc->inbox[i] = i < dp->world ? dp->inbox[i] : nullptr;
    if (i < dp->world && (!dp->inbox[i] || !aligned16(dp->inbox[i]))) return OSPO_ERR_NULL;
  }
  return OSPO_OK;
}

// shared backward: row weights -> GEMM pair(s) straight on the forward's spill (read-only here)
int head_backward(const ospo_simpo_args* a, const Workspace& w, const float* sft_coef, int num_sft_seqs,
                  cudaStream_t st) {
  const ospo_head_shape& s = a->shape;
  if (!a->dx && !a->flat_grads) return OSPO_OK;
  if (!a->pre || !a->act || !a->logits || !a->row_lse || !a->row_ref || !a->grad_seq) return OSPO_ERR_NULL;
  if (a->flat_grads && !w.act_w) return OSPO_ERR_WORKSPACE;
  LaunchCtx c = make_ctx(st);
  int rc;
  // bwd_stage: 0 = everything; otherwise a bit mask of the parts to run now: 1 = row weights, dpre (+ act_w), db2
  // and dW2;  2 = db1 and dW1;  4 = dX.  A data-parallel caller runs 1, starts the all-reduce of dW2, runs 2,
  // starts the all-reduce of the rest and runs 4.  dpre lives in the workspace: same buffer in every call.
  const int stage = a->bwd_stage == 0 ? 7 : a->bwd_stage;
  if (stage < 1 || stage > 7) return OSPO_ERR_UNSUPPORTED;
  // A staged backward means a collective runs beside dW1 / dX.  There the split-K form of dW1 (gemm_bwd.cu) loses: its
  // red.global.add traffic competes with the all-reduce of dW2 -- 2 GPUs, alternating runs on one box: dW1 2.25 - 2.41
  // ms unsplit against 2.48 - 2.54 ms split, step 28.7 against 28.9 - 29.3 ms
  // (profiles/r02_ab_n2_wgrad_splitk_*.json) -- so it is chosen only when the backward runs in one piece.
  if (a->bwd_stage != 0 && c.wgrad_splitk == 0) c.wgrad_splitk = 1;
  const bool first = (stage & 1) != 0, second = (stage & 2) != 0, third = (stage & 4) != 0;
  if (!first && a->reserve_sms > 0) {
    // leave SMs to a collective kernel running beside the remaining GEMMs (even count: CTA pairs)
    const int keep = c.num_sms - (a->reserve_sms + 1) / 2 * 2;
    if (keep >= c.num_sms / 2) c.num_sms = keep;
  }
  const size_t VE = static_cast<size_t>(s.vocab) * s.embed, EH = static_cast<size_t>(s.embed) * s.hidden;
  float* dW2 = a->flat_grads;
  float* dW1 = a->flat_grads ? a->flat_grads + VE : nullptr;
  float* db2 = a->flat_grads ? a->flat_grads + VE + EH : nullptr;
  float* db1 = a->flat_grads ? db2 + s.vocab : nullptr;
  const float wscale = a->wgrad_scale != 0.0f ? a->wgrad_scale : 1.0f;  // 1 / world_size of a data-parallel caller
  // peer-memory exchange: the gradient stores go to the owners' inboxes instead of flat_grads
  DpScatter sc = {};
  DpPlan pl = {};
  if (a->dp != nullptr && a->flat_grads) {
    if ((rc = dp_plan(s, a->dp, &pl, &sc))) return rc;
  }
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(a->logits);
  __nv_bfloat16* dpre = w.rows_by_e;
  float* row_w = w.row_coef;
  if (first) {
    {
      KernelSpan ks(st, OSPO_K_ROW_WEIGHTS);
      row_weight_kernel<<<s.num_seqs, 128, 0, st>>>(a->grad_seq, a->seq_offsets, a->labels, s.vocab,
                                                    a->average_log_prob, a->grad_loss, sft_coef, num_sft_seqs,
                                                    a->row_lse, a->row_ref, row_w);
      if ((rc = check_launch())) return rc;
    }
    {
      KernelSpan ks(st, OSPO_K_DACT);
      rc = map_rc(launch_dact_gelu_bwd(tuned(c, 2), g, static_cast<const __nv_bfloat16*>(a->w.w2),
                                       static_cast<const __nv_bfloat16*>(a->pre), row_w, dpre,
                                       a->flat_grads ? w.act_w : nullptr, s.rows, s.embed, s.vocab));
      if (rc) return rc;
    }
  }
  if (a->flat_grads && first) {
    {
      KernelSpan ks(st, OSPO_K_COLSUM);
      if ((rc = launch_colsum<true>(g, s.rows, s.vocab, row_w, wscale, w.colsum_part, db2, st, sc, pl.a + pl.b)))
        return rc;
    }
    // dW2 first: it is the largest block of the flat gradient, so a caller that overlaps the
    // all-reduce with the remaining GEMMs can start on it earliest.
    KernelSpan ks(st, OSPO_K_WGRAD2);
    if (sc.world > 0)
      rc = map_rc(launch_wgrad_scatter(tuned(c, 3), g, w.act_w, sc, 0, s.rows, s.vocab, s.embed, wscale));
    else
      rc = map_rc(launch_wgrad(tuned(c, 3), g, w.act_w, dW2, s.rows, s.vocab, s.embed, wscale));
    if (rc) return rc;
  }
  if (a->flat_grads && second) {
    {
      KernelSpan ks(st, OSPO_K_COLSUM);
      if ((rc = launch_colsum<false>(dpre, s.rows, s.embed, nullptr, wscale, w.colsum_part, db1, st, sc,
                                     pl.a + pl.b + pl.vb)))
        return rc;
    }
    {
      KernelSpan ks(st, OSPO_K_WGRAD1);
      if (sc.world > 0)
        rc = map_rc(launch_wgrad_scatter(tuned(c, 4), dpre, static_cast<const __nv_bfloat16*>(a->x), sc, pl.a, s.rows,
                                         s.embed, s.hidden, wscale, x_layout(a)));
      else
        rc = map_rc(launch_wgrad(tuned(c, 4), dpre, static_cast<const __nv_bfloat16*>(a->x), dW1, s.rows, s.embed,
                                 s.hidden, wscale, x_layout(a)));
    }
    if (rc) return rc;
  }
  if (a->dx && third) {
    KernelSpan ks(st, OSPO_K_DGRAD);
    rc = map_rc(launch_dgrad(tuned(c, 5), dpre, static_cast<const __nv_bfloat16*>(a->w.w1), static_cast<__nv_bfloat16*>(a->dx),
                             s.rows, s.embed, s.hidden, x_layout(a)));
    if (rc) return rc;
  }
  return OSPO_OK;
}

template <int MODE, bool TDIV, bool GREEDY, bool WBF>
bool run_sampler(int grid, cudaStream_t st, const __nv_bfloat16* lg, int V, const ospo_cfg_args* a, int pairs) {
  constexpr int kLanding = SAMPLE_THREADS * 8 * 16;  // 64 KB of dynamic shared memory: opt in once per instantiation
  auto kern = cfg_merge_sample_kernel<MODE, TDIV, GREEDY, WBF>;
  static std::atomic<uint64_t> attr_set{0};  // per device ordinal
  int attr_dev = 0;
  if (func_attrs_needed(attr_set, &attr_dev)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kLanding) != cudaSuccess) return false;
    func_attrs_mark(attr_set, attr_dev);
  }
  kern<<<grid, GREEDY ? SAMPLE_THREADS : SAMPLE_BLOCK, kLanding, st>>>(lg, V, V, a->cfg_weight, a->temperature, a->uniforms, a->ids, a->merged,
                                               pairs);
  return true;
}

}  // namespace

extern "C" {

int ospo_head_workspace_bytes(const ospo_head_shape* shape, size_t* out_bytes) {
  if (!shape || !out_bytes) return OSPO_ERR_NULL;
  if (shape->rows <= 0 || shape->embed <= 0 || shape->vocab <= 0) return OSPO_ERR_BAD_SHAPE;
  *out_bytes = carve(*shape, nullptr).total;
  return OSPO_OK;
}

int ospo_head_logits(const ospo_head_args* a, ospo_stream_t stream) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if ((rc = check_shape(a->shape, false))) return rc;
  if ((rc = check_weights(a->w))) return rc;
  if (!a->x || !a->logits) return OSPO_ERR_NULL;
  if (!aligned16(a->x) || !aligned16(a->logits)) return OSPO_ERR_ALIGNMENT;
  Workspace w;
  if ((rc = check_ws(a->shape, a->workspace, a->workspace_bytes, &w))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const LaunchCtx c = make_ctx(st);
  const ospo_head_shape& s = a->shape;
  {
    KernelSpan ks(st, OSPO_K_GEMM1);
    rc = map_rc(launch_gemm1_bias_gelu(c, static_cast<const __nv_bfloat16*>(a->x),
                                       static_cast<const __nv_bfloat16*>(a->w.w1), a->w.b1, nullptr, w.rows_by_e,
                                       s.rows, s.hidden, s.embed));
  }
  if (rc) return rc;
  KernelSpan ks(st, OSPO_K_GEMM2_PLAIN);
  return map_rc(launch_gemm2_logits(c, w.rows_by_e, static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2,
                                    static_cast<__nv_bfloat16*>(a->logits), s.vocab, s.rows, s.embed, s.vocab));
}

int ospo_head_logps_fwd(const ospo_simpo_args* a, ospo_stream_t stream) {
  Workspace w;
  int rc = check_simpo_common(a, &w);
  if (rc) return rc;
  if (!a->act || !a->row_lse || !a->row_logps || !a->seq_logps) return OSPO_ERR_NULL;
  return logps_forward(a, w, reinterpret_cast<cudaStream_t>(stream));
}

int ospo_head_logps_bwd(const ospo_simpo_args* a, ospo_stream_t stream) {
  Workspace w;
  int rc = check_simpo_common(a, &w);
  if (rc) return rc;
  return head_backward(a, w, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

int ospo_head_simpo_fwd(const ospo_simpo_args* a, ospo_stream_t stream) {
  Workspace w;
  int rc = check_simpo_common(a, &w);
  if (rc) return rc;
  if (a->shape.num_seqs % 2) return OSPO_ERR_BAD_SHAPE;
  if (!a->act || !a->row_lse || !a->row_logps || !a->seq_logps || !a->losses || !a->chosen_rewards ||
      !a->rejected_rewards || !a->scalars || !a->grad_seq)
    return OSPO_ERR_NULL;
  if (a->loss_type != OSPO_LOSS_SIGMOID && a->loss_type != OSPO_LOSS_HINGE) return OSPO_ERR_UNSUPPORTED;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((rc = logps_forward(a, w, st))) return rc;
  SimpoHyper hp;
  hp.beta = a->beta;
  hp.gamma_beta_ratio = a->gamma_beta_ratio;
  hp.label_smoothing = a->label_smoothing;
  hp.sft_weight = a->sft_weight;
  hp.loss_type = a->loss_type;
  hp.vocab = a->shape.vocab;
  KernelSpan ks(st, OSPO_K_SCALAR_STAGE);
  simpo_scalar_kernel<<<1, 256, 0, st>>>(a->seq_logps, w.seq_sum, w.seq_logit_sum, w.seq_count,
                                         a->shape.num_seqs / 2, hp, a->losses, a->chosen_rewards, a->rejected_rewards,
                                         a->grad_seq, a->scalars);
  return check_launch();
}

int ospo_head_dp_reduce_broadcast(const ospo_head_shape* shape, const ospo_dp_exchange* dp, int32_t regions,
                                  int32_t max_blocks, ospo_stream_t stream) {
  if (!shape || !dp) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if ((rc = check_shape(*shape, false))) return rc;
  DpScatter sc = {};
  DpPlan pl = {};
  if ((rc = dp_plan(*shape, dp, &pl, &sc))) return rc;
  DpGather g = {};
  for (int i = 0; i < dp->world; ++i) {
    if (!dp->flat[i] || !aligned16(dp->flat[i])) return OSPO_ERR_NULL;
    g.flat[i] = dp->flat[i];
  }
  g.flat_mc = dp->flat_multicast;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  KernelSpan ks(st, OSPO_K_DP_EXCHANGE);
  const int64_t VE = static_cast<int64_t>(shape->vocab) * shape->embed, EH = static_cast<int64_t>(shape->embed) * shape->hidden;
  // regions: 1 = the dW2 rows of the shard, 2 = everything else (dW1 rows, db2, db1), 0 / 3 = all
  const int r = (regions & 3) == 0 ? 3 : (regions & 3);
  const int64_t i_begin = (r & 1) ? 0 : pl.a, i_end = (r & 2) ? pl.shard : pl.a;
  // HBM-bound on the local inbox (world x shard bytes read, shard bytes sent): four resident blocks per SM unless the
  // caller runs it beside GEMMs and asks for a small grid
  int blocks = current_num_sms() * 4;
  if (max_blocks > 0 && max_blocks < blocks) blocks = max_blocks;
  dp_reduce_broadcast_kernel<<<blocks, 256, 0, st>>>(dp->inbox[dp->rank], g, dp->world, dp->rank, pl.shard, pl.a, pl.b,
                                                     pl.vb, VE, EH, shape->vocab, i_begin, i_end);
  return check_launch();
}

int ospo_head_simpo_bwd(const ospo_simpo_args* a, ospo_stream_t stream) {
  Workspace w;
  int rc = check_simpo_common(a, &w);
  if (rc) return rc;
  if (!a->scalars) return OSPO_ERR_NULL;
  const float* sft_coef = (a->sft_weight > 0.0f) ? a->scalars + SC_SFT_ROW_COEF : nullptr;
  return head_backward(a, w, sft_coef, a->shape.num_seqs / 2, reinterpret_cast<cudaStream_t>(stream));
}

static int launch_sampler(const ospo_cfg_args* a, int pairs, cudaStream_t st) {
  const int V = a->shape.vocab;
  if (V != SAMPLE_THREADS * SAMPLE_SEG) return OSPO_ERR_UNSUPPORTED;
  if (!a->greedy && !a->uniforms) return OSPO_ERR_NULL;
  if (!a->ids || !a->logits) return OSPO_ERR_NULL;
  if (!(a->temperature > 0.0f)) return OSPO_ERR_UNSUPPORTED;
  KernelSpan ks(st, OSPO_K_SAMPLER);
  const __nv_bfloat16* lg = static_cast<const __nv_bfloat16*>(a->logits);
  const bool tdiv = (a->temperature != 1.0f);
  const int mode = (a->merge_mode == OSPO_MERGE_FP32) ? 1 : 0;
  // bf16 merge with a cfg_weight that is itself a bf16 value (5.0, 7.5, ...): merge on the bf16x2 pipe
  const bool wbf = (mode == 0) && bf16_exact(a->cfg_weight);
  // persistent blocks (two per SM), each with a 64 KB landing buffer for its next pair's rows
  const int grid = std::min(pairs, 2 * current_num_sms());
  const int variant = (a->greedy ? 8 : 0) | (mode ? 4 : 0) | (tdiv ? 2 : 0) | (wbf ? 1 : 0);
  bool ok = false;
#define OSPO_SAMPLER_CASE(MODE, TDIV, GREEDY, WBF)                                                             \
  case ((GREEDY ? 8 : 0) | (MODE ? 4 : 0) | (TDIV ? 2 : 0) | (WBF ? 1 : 0)):                                   \
    ok = run_sampler<MODE, TDIV, GREEDY, WBF>(grid, st, lg, V, a, pairs);                                      \
    break
  switch (variant) {
    OSPO_SAMPLER_CASE(0, false, false, false);
    OSPO_SAMPLER_CASE(0, false, false, true);
    OSPO_SAMPLER_CASE(0, true, false, false);
    OSPO_SAMPLER_CASE(0, true, false, true);
    OSPO_SAMPLER_CASE(1, false, false, false);
    OSPO_SAMPLER_CASE(1, true, false, false);
    OSPO_SAMPLER_CASE(0, false, true, false);
    OSPO_SAMPLER_CASE(0, false, true, true);
    OSPO_SAMPLER_CASE(0, true, true, false);
    OSPO_SAMPLER_CASE(0, true, true, true);
    OSPO_SAMPLER_CASE(1, false, true, false);
    OSPO_SAMPLER_CASE(1, true, true, false);
    default: break;
  }
#undef OSPO_SAMPLER_CASE
  if (!ok) return OSPO_ERR_LAUNCH;
  return check_launch();
}

int ospo_head_cfg_merge_sample(const ospo_cfg_args* a, ospo_stream_t stream) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if (a->shape.rows <= 0 || (a->shape.rows % 2)) return OSPO_ERR_BAD_SHAPE;
  if (!aligned16(a->logits)) return OSPO_ERR_ALIGNMENT;
  const int steps = a->num_steps > 1 ? a->num_steps : 1;
  return launch_sampler(a, steps * (a->shape.rows / 2), reinterpret_cast<cudaStream_t>(stream));
}

// the decode step proper; `eu` (may be off) is the first aligner layer folded into the finish kernel, *eu_done tells
// the caller whether the variant that ran has done it
static int cfg_sample_step(const ospo_cfg_args* a, ospo_stream_t stream, const EmbedUp& eu, bool* eu_done,
                           const ospo_aligner_args* ne) {
  int rc;
  *eu_done = false;
  if ((rc = check_shape(a->shape, false))) return rc;
  if (a->shape.rows % 2) return OSPO_ERR_BAD_SHAPE;
  if ((rc = check_weights(a->w))) return rc;
  if (!a->h || !a->ids) return OSPO_ERR_NULL;
  if (!a->greedy && !a->uniforms) return OSPO_ERR_NULL;
  if (!aligned16(a->h) || (a->logits && !aligned16(a->logits))) return OSPO_ERR_ALIGNMENT;
  if (a->shape.vocab != SAMPLE_THREADS * SAMPLE_SEG) return OSPO_ERR_UNSUPPORTED;
  if (!(a->temperature > 0.0f)) return OSPO_ERR_UNSUPPORTED;
  if (static_cast<size_t>(a->shape.rows) > kDecodeMaxRows) return OSPO_ERR_UNSUPPORTED;
  Workspace w;
  if ((rc = check_ws(a->shape, a->workspace, a->workspace_bytes, &w))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchCtx c = make_ctx(st);
  c.pdl = g_rt.decode_pdl != 0;
  const ospo_head_shape& s = a->shape;
  const int pairs = s.rows / 2;
  if (g_rt.decode_fused && g_rt.decode_merged && !a->merged) {
    // whole step as one persistent kernel (W1 stream -> cluster reduction -> flag -> W2 stream + CFG epilogue)
    uint32_t* flag = flag_words_for(st);
    int lrc = -100;
    if (flag != nullptr) {
      KernelSpan ks(st, OSPO_K_DECODE_GEMM2);
      lrc = launch_decode_merged(c, static_cast<const __nv_bfloat16*>(a->h), static_cast<const __nv_bfloat16*>(a->w.w1),
                                 a->w.b1, static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2, w.rows_by_e, flag,
                                 static_cast<__nv_bfloat16*>(a->logits), s.rows, s.hidden, s.embed, s.vocab,
                                 a->cfg_weight, a->temperature, a->merge_mode == OSPO_MERGE_FP32 ? 1 : 0, a->greedy,
                                 w.fused, g_rt.decode_l2_ahead, a->w1_packed, a->w2_packed,
                                 (ne && g_rt.decode_next_prefetch) ? static_cast<const __nv_bfloat16*>(ne->wb) : nullptr,
                                 ne ? ne->embed : 0, ne ? ne->embed : 0);
    }
    if (lrc == 0) {
      g_launches.fetch_add(1, std::memory_order_relaxed);
      KernelSpan ks(st, OSPO_K_SAMPLER);
      if (launch_plain(eu.gen_embed ? cfg_finish_kernel<true> : cfg_finish_kernel<false>, dim3(pairs), dim3(SAMPLE_THREADS), st,
                       c.pdl, w.fused, s.vocab, a->uniforms,
                       a->greedy, a->ids, c.trace ? 1 : 0, eu) != cudaSuccess)
        return OSPO_ERR_LAUNCH;
      g_launches.fetch_add(1, std::memory_order_relaxed);
      *eu_done = eu.gen_embed != nullptr;
      return OSPO_OK;
    }
    if (lrc != -100) return map_rc(lrc);
  }
  {
    // W1 slabs over all SMs (split-K partials), then bias + GELU on the summed partials
    KernelSpan ks(st, OSPO_K_DECODE_GEMM1);
    const int64_t split_stride = static_cast<int64_t>(s.rows) * s.embed;
    int lrc = g_rt.decode_cluster
                  ? launch_decode_gemm1_cluster(c, static_cast<const __nv_bfloat16*>(a->h),
                                                static_cast<const __nv_bfloat16*>(a->w.w1), a->w.b1, w.rows_by_e,
                                                s.rows, s.hidden, s.embed)
                  : -100;
    if (lrc == -100) {
      // partial + finalize path (any shape)
      rc = map_rc(launch_decode_gemm1(c, static_cast<const __nv_bfloat16*>(a->h),
                                      static_cast<const __nv_bfloat16*>(a->w.w1), w.decode_part, split_stride, s.rows,
                                      s.hidden, s.embed));
      if (rc) return rc;
      const int64_t n_el = static_cast<int64_t>(s.rows) * s.embed;
      const float* part = w.decode_part;
      if (launch_plain(decode_act_finalize_kernel, dim3(static_cast<unsigned>((n_el / 4 + 127) / 128)), dim3(128), st,
                       c.pdl, part, decode_gemm1_splits(c.num_sms, s.hidden, s.embed), split_stride, a->w.b1,
                       w.rows_by_e, s.rows, s.embed, c.trace ? 1 : 0) != cudaSuccess)
        return OSPO_ERR_LAUNCH;
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if ((rc = map_rc(lrc))) {
      return rc;
    }
  }
  if (g_rt.decode_fused && !a->merged) {
    {
      KernelSpan ks(st, OSPO_K_DECODE_GEMM2);
      rc = map_rc(launch_decode_gemm2_fused(c, w.rows_by_e, static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2,
                                            static_cast<__nv_bfloat16*>(a->logits), s.rows, s.embed, s.vocab,
                                            a->cfg_weight, a->temperature, a->merge_mode == OSPO_MERGE_FP32 ? 1 : 0,
                                            a->greedy, w.fused));
    }
    if (rc) return rc;
    KernelSpan ks(st, OSPO_K_SAMPLER);
    if (launch_plain(eu.gen_embed ? cfg_finish_kernel<true> : cfg_finish_kernel<false>, dim3(pairs), dim3(SAMPLE_THREADS), st,
                       c.pdl, w.fused, s.vocab, a->uniforms,
                     a->greedy, a->ids, c.trace ? 1 : 0, eu) != cudaSuccess)
      return OSPO_ERR_LAUNCH;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    *eu_done = eu.gen_embed != nullptr;
    return OSPO_OK;
  }
  // unfused variant: materialise the bf16 logits, then the stand-alone merge + sample pass
  ospo_cfg_args b = *a;
  if (!b.logits) b.logits = w.decode_logits;
  {
    KernelSpan ks(st, OSPO_K_DECODE_GEMM2);
    rc = map_rc(launch_decode_gemm2(c, w.rows_by_e, static_cast<const __nv_bfloat16*>(a->w.w2), a->w.b2,
                                    static_cast<__nv_bfloat16*>(b.logits), s.rows, s.embed, s.vocab));
  }
  if (rc) return rc;
  return launch_sampler(&b, pairs, st);
}

static int check_aligner(const ospo_aligner_args* a, bool need_ids) {
  if (a->rows <= 0 || a->embed <= 0 || a->codebook <= 0) return OSPO_ERR_BAD_SHAPE;
  if (a->code_dim != 8 || a->rows > 32) return OSPO_ERR_UNSUPPORTED;
  if (a->embed % 8) return OSPO_ERR_ALIGNMENT;
  if (a->table != nullptr) {
    // memo-table form: only ids, table and out are used
    if ((need_ids && !a->ids) || !a->out) return OSPO_ERR_NULL;
    if (!aligned16(a->table) || !aligned16(a->out)) return OSPO_ERR_ALIGNMENT;
    const int rep_t = a->id_repeat > 1 ? a->id_repeat : 1;
    return (a->rows % rep_t) ? OSPO_ERR_BAD_SHAPE : OSPO_OK;
  }
  if ((need_ids && !a->ids) || !a->gen_embed || !a->wa || !a->ba || !a->wb || !a->bb || !a->out || !a->workspace)
    return OSPO_ERR_NULL;
  if (!aligned16(a->gen_embed) || !aligned16(a->wa) || !aligned16(a->wb) || !aligned16(a->out) ||
      !aligned16(a->workspace))
    return OSPO_ERR_ALIGNMENT;
  if (a->workspace_bytes < static_cast<size_t>(a->rows) * a->embed * 2) return OSPO_ERR_WORKSPACE;
  const int rep = a->id_repeat > 1 ? a->id_repeat : 1;
  if (a->rows % rep) return OSPO_ERR_BAD_SHAPE;
  return OSPO_OK;
}

// the D x D Linear of gen_aligner on a1 [rows, D] (already in the aligner workspace)
static int aligner_second_linear(const ospo_aligner_args* a, const LaunchCtx& c) {
  const __nv_bfloat16* a1 = static_cast<const __nv_bfloat16*>(a->workspace);
  int lrc = g_rt.decode_merged ? launch_decode_linear(c, a1, static_cast<const __nv_bfloat16*>(a->wb), a->bb,
                                                      static_cast<__nv_bfloat16*>(a->out), a->rows, a->embed, a->embed, 0)
                               : -100;
  if (lrc == -100)
    lrc = launch_decode_linear_cluster(c, a1, static_cast<const __nv_bfloat16*>(a->wb), a->bb,
                                       static_cast<__nv_bfloat16*>(a->out), a->rows, a->embed, a->embed);
  return map_rc(lrc);
}

int ospo_head_cfg_sample(const ospo_cfg_args* a, ospo_stream_t stream) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  EmbedUp eu = {nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr};
  const ospo_aligner_args* ne = a->next_embeds;
  if (ne != nullptr) {
    if ((rc = check_aligner(ne, false))) return rc;
    if (ne->rows != a->shape.rows || ne->id_repeat != 2) return OSPO_ERR_BAD_SHAPE;
    if (ne->table != nullptr) {
      // the finish kernel copies the pair's rows of the memo table straight into `out`: nothing else runs
      eu = EmbedUp{static_cast<const __nv_bfloat16*>(ne->table), nullptr, nullptr, static_cast<__nv_bfloat16*>(ne->out),
                   ne->codebook, ne->embed, static_cast<const __nv_bfloat16*>(ne->table)};
    } else {
      eu = EmbedUp{static_cast<const __nv_bfloat16*>(ne->gen_embed), static_cast<const __nv_bfloat16*>(ne->wa), ne->ba,
                   static_cast<__nv_bfloat16*>(ne->workspace), ne->codebook, ne->embed, nullptr};
    }
  }
  bool eu_done = false;
  if ((rc = cfg_sample_step(a, stream, eu, &eu_done, (ne && ne->table) ? nullptr : ne))) return rc;
  if (ne == nullptr) return OSPO_OK;
  if (eu_done && ne->table != nullptr) return OSPO_OK;
  if (!eu_done) {
    // this decode variant has no finish kernel: run the stand-alone embedding path on the sampled ids
    ospo_aligner_args full = *ne;
    full.ids = a->ids;
    return ospo_head_gen_img_embeds(&full, stream);
  }
  LaunchCtx c = make_ctx(reinterpret_cast<cudaStream_t>(stream));
  c.pdl = g_rt.decode_pdl != 0;
  KernelSpan ks(c.stream, OSPO_K_ALIGNER);
  return aligner_second_linear(ne, c);
}

int ospo_head_gen_img_embeds(const ospo_aligner_args* a, ospo_stream_t stream) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if ((rc = check_aligner(a, true))) return rc;
  const int rep = a->id_repeat > 1 ? a->id_repeat : 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LaunchCtx c = make_ctx(st);
  c.pdl = g_rt.decode_pdl != 0;
  KernelSpan ks(st, OSPO_K_ALIGNER);
  if (a->table != nullptr) {
    if (launch_plain(embed_table_gather_kernel, dim3(a->rows), dim3(256), st, c.pdl, a->ids,
                     static_cast<const __nv_bfloat16*>(a->table), a->codebook, static_cast<__nv_bfloat16*>(a->out),
                     a->rows, a->embed, rep) != cudaSuccess)
      return OSPO_ERR_LAUNCH;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return OSPO_OK;
  }
  __nv_bfloat16* a1 = static_cast<__nv_bfloat16*>(a->workspace);
  if (launch_plain(gen_embed_up_kernel, dim3((a->embed + 255) / 256, a->rows), dim3(256), st, c.pdl, a->ids,
                   static_cast<const __nv_bfloat16*>(a->gen_embed), a->codebook,
                   static_cast<const __nv_bfloat16*>(a->wa), a->ba, a1, a->rows, a->embed, rep) != cudaSuccess)
    return OSPO_ERR_LAUNCH;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return aligner_second_linear(a, c);
}

int ospo_head_packed_weight_bytes(int32_t rows, int32_t cols, size_t* out_bytes) {
  if (!out_bytes) return OSPO_ERR_NULL;
  if (rows <= 0 || cols <= 0) return OSPO_ERR_BAD_SHAPE;
  *out_bytes = static_cast<size_t>((rows + 127) / 128) * ((cols + 63) / 64) * 16384;
  return OSPO_OK;
}

int ospo_head_pack_weight(const void* w, int32_t rows, int32_t cols, void* packed, ospo_stream_t stream) {
  int rc = runtime_init();
  if (rc) return rc;
  if (!w || !packed) return OSPO_ERR_NULL;
  if (rows <= 0 || cols <= 0) return OSPO_ERR_BAD_SHAPE;
  if (!aligned16(w) || !aligned16(packed) || (cols % 8)) return OSPO_ERR_ALIGNMENT;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = static_cast<int64_t>((rows + 127) / 128) * ((cols + 63) / 64) * 1024;
  pack_weight_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(w), rows, cols, static_cast<uint4*>(packed));
  return check_launch();
}

int ospo_head_grad_sqnorm(const float* grads, int64_t numel, float* out_sq, void* workspace, size_t workspace_bytes,
                          ospo_stream_t stream) {
  int rc = runtime_init();
  if (rc) return rc;
  if (!grads || !out_sq || !workspace) return OSPO_ERR_NULL;
  if (numel <= 0) return OSPO_ERR_BAD_SHAPE;
  if (!aligned16(grads)) return OSPO_ERR_ALIGNMENT;
  const int blocks = current_num_sms() * 4;  // four resident blocks per SM keep every HBM channel busy
  if (workspace_bytes < static_cast<size_t>(blocks) * sizeof(float)) return OSPO_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  KernelSpan ks(st, OSPO_K_OPTIMIZER);
  float* partials = static_cast<float*>(workspace);
  sqnorm_partial_kernel<<<blocks, OPT_THREADS, 0, st>>>(grads, numel, partials);
  if ((rc = check_launch())) return rc;
  sqnorm_final_kernel<<<1, 32, 0, st>>>(partials, blocks, out_sq);
  return check_launch();
}

int ospo_head_adamw_step(const ospo_adamw_args* a, ospo_stream_t stream) {
  if (!a) return OSPO_ERR_NULL;
  int rc = runtime_init();
  if (rc) return rc;
  if (a->numel <= 0 || a->step < 1) return OSPO_ERR_BAD_SHAPE;
  if (!a->grads || !a->params || !a->exp_avg || !a->exp_avg_sq) return OSPO_ERR_NULL;
  if (a->max_norm > 0.0f && !a->total_sqnorm) return OSPO_ERR_NULL;
  if (!aligned16(a->grads) || !aligned16(a->params) || !aligned16(a->exp_avg) || !aligned16(a->exp_avg_sq))
    return OSPO_ERR_ALIGNMENT;
  if (a->params_bf16 && (!aligned16(a->params_bf16) || (a->shadow_numel % 4) || a->shadow_numel < 0 ||
                         a->shadow_numel > a->numel))
    return OSPO_ERR_BAD_SHAPE;
  if (!(a->beta1 >= 0.0 && a->beta1 < 1.0 && a->beta2 >= 0.0 && a->beta2 < 1.0)) return OSPO_ERR_UNSUPPORTED;
  // the scalars torch derives in Python (double precision), rounded to fp32 where torch hands them to its kernels
  AdamWHyper h;
  const double lr = a->lr, b1 = a->beta1, b2 = a->beta2;
  h.decay = static_cast<float>(1.0 - lr * a->weight_decay);
  h.w1 = static_cast<float>(1.0 - b1);
  h.beta2 = static_cast<float>(b2);
  h.w2 = static_cast<float>(1.0 - b2);
  h.bias2_sqrt = static_cast<float>(std::sqrt(1.0 - std::pow(b2, static_cast<double>(a->step))));
  h.eps = static_cast<float>(a->eps);
  h.step_size = static_cast<float>(lr / (1.0 - std::pow(b1, static_cast<double>(a->step))));
  h.max_norm = a->max_norm;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  KernelSpan ks(st, OSPO_K_OPTIMIZER);
  const int blocks = current_num_sms() * 4;
  adamw_kernel<<<blocks, OPT_THREADS, 0, st>>>(a->grads, a->params, a->exp_avg, a->exp_avg_sq,
                                               static_cast<__nv_bfloat16*>(a->params_bf16),
                                               a->params_bf16 ? a->shadow_numel : 0, a->numel, a->total_sqnorm, h);
  return check_launch();
}

const char* ospo_head_strerror(int status) {
  switch (status) {
    case OSPO_OK: return "ok";
    case OSPO_ERR_BAD_SHAPE: return "bad shape";
    case OSPO_ERR_ALIGNMENT: return "pointer must be 16-byte aligned and H/E/V multiples of 8";
    case OSPO_ERR_NULL: return "required pointer is NULL";
    case OSPO_ERR_WORKSPACE: return "workspace too small (see ospo_head_workspace_bytes)";
    case OSPO_ERR_ARCH: return "device is not sm_100 (B200); this library has no other code path";
    case OSPO_ERR_TENSORMAP: return "cuTensorMapEncodeTiled failed or is unavailable";
    case OSPO_ERR_LAUNCH: return "kernel launch failed";
    case OSPO_ERR_CUDA: return "CUDA runtime error";
    case OSPO_ERR_UNSUPPORTED: return "unsupported argument combination";
    default: return "unknown status";
  }
}

int ospo_head_set_cta_group(int cta_group) {
  runtime_init();
  if (cta_group == 1 || cta_group == 2) g_rt.cta_group = cta_group;
  return g_rt.cta_group;
}

int ospo_head_set_decode_mode(int fused, int pdl) {
  runtime_init();
  if (fused == 0 || fused == 1) g_rt.decode_fused = fused;
  if (pdl == 0 || pdl == 1) g_rt.decode_pdl = pdl;
  return g_rt.decode_fused | (g_rt.decode_pdl << 1);
}

int ospo_head_set_decode_merged(int merged) {
  runtime_init();
  if (merged == 0 || merged == 1) g_rt.decode_merged = merged;
  return g_rt.decode_merged;
}

int ospo_head_set_decode_l2_ahead(int kblocks) {
  runtime_init();
  if (kblocks >= 0) g_rt.decode_l2_ahead = kblocks;
  return g_rt.decode_l2_ahead;
}

int ospo_head_set_group_m(int group_m) {
  runtime_init();
  if (group_m > 0) g_rt.group_m = group_m;
  return g_rt.group_m;
}

int ospo_head_set_wgrad_splitk(int splits) {
  runtime_init();
  if (splits >= 0 && splits <= 2) g_rt.wgrad_splitk = splits;
  return g_rt.wgrad_splitk;
}

int ospo_head_set_kernel_tune(int kernel, int group_m, int a_evict, int b_evict) {
  runtime_init();
  if (kernel < 0 || kernel >= 6) return OSPO_ERR_UNSUPPORTED;
  if (group_m >= 0) g_rt.tune[kernel][0] = group_m;
  if (a_evict >= 0 && a_evict <= 2) g_rt.tune[kernel][1] = a_evict;
  if (b_evict >= 0 && b_evict <= 2) g_rt.tune[kernel][2] = b_evict;
  return OSPO_OK;
}

int ospo_head_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_profile = enable != 0;
  if (g_profile) {
    // events are created up front so that no cudaEventCreate lands inside a timed region
    g_spans.reserve(4096);
    while (g_event_pool.size() < 1024) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      g_event_pool.push_back(e);
    }
  }
  return g_profile ? 1 : 0;
}

int ospo_head_profile_read(float* total_ms, int32_t* counts, int32_t n) {
  // synchronises on the recorded events; accumulates per kernel id, then recycles the events
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < n; ++i) {
    if (total_ms) total_ms[i] = 0.0f;
    if (counts) counts[i] = 0;
  }
  int rc = OSPO_OK;
  for (const Span& sp : g_spans) {
    float ms = 0.0f;
    if (cudaEventSynchronize(sp.e1) != cudaSuccess || cudaEventElapsedTime(&ms, sp.e0, sp.e1) != cudaSuccess) {
      rc = OSPO_ERR_CUDA;
      cudaGetLastError();
    } else if (sp.kid >= 0 && sp.kid < n) {
      if (total_ms) total_ms[sp.kid] += ms;
      if (counts) counts[sp.kid] += 1;
    }
    g_event_pool.push_back(sp.e0);
    g_event_pool.push_back(sp.e1);
  }
  g_spans.clear();
  return rc;
}

int ospo_head_trace(void* device_buf) {
  // tuning aid: [8][160][8] u64 globaltimer stamps of the decode chain's CTAs (NULL switches it off)
  runtime_init();
  unsigned long long* p = static_cast<unsigned long long*>(device_buf);
  set_trace_decode(p);
  set_trace_merged(p);
  g_rt.trace_on = p != nullptr;
  g_rt.trace_buf = p;
  return cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)) == cudaSuccess ? OSPO_OK : OSPO_ERR_CUDA;
}

uint64_t ospo_head_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const uint32_t* ospo_head_watchdog_record_host(void) { return g_rt.wd_host; }

int ospo_head_gemm_debug(int variant, const void* a, int64_t lda, const void* b, int64_t ldb, float* out, int64_t ldo,
                         int32_t M, int32_t N, int32_t K, ospo_stream_t stream) {
  int rc = runtime_init();
  if (rc) return rc;
  if (!a || !b || !out) return OSPO_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0) return OSPO_ERR_BAD_SHAPE;
  if (!aligned16(a) || !aligned16(b) || (lda % 8) || (ldb % 8)) return OSPO_ERR_ALIGNMENT;
  const LaunchCtx c = make_ctx(reinterpret_cast<cudaStream_t>(stream));
  return map_rc(launch_gemm_debug(c, variant, static_cast<const __nv_bfloat16*>(a), lda,
                                  static_cast<const __nv_bfloat16*>(b), ldb, out, ldo, M, N, K));
}

}  // extern "C"
