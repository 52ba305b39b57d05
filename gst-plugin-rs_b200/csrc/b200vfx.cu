// b200vfx.cu -- C ABI (include/b200vfx.h) over the sm_100a kernels in kernels.cuh.
//
// Host-pointer calls run a row-chunked H2D -> kernel -> D2H pipeline on three streams so the two
// PCIe directions and the kernel overlap inside ONE synchronous transform_frame call.
// Device-pointer calls enqueue the kernel on the context stream and return.
// There is no CPU fallback anywhere in this file.
#include "../../include/b200vfx.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "stream_map.cuh"
#include "blockhash_tma.cuh"
#include "tile_gather.cuh"
#include "colordetect.cuh"
#include "memo_tile.cuh"
#include "hash_kernels.cuh"
#include "convert.cuh"
#include "hash_host.h"

using namespace b200vfx;

namespace {

thread_local std::string g_last_error;

struct DevBuf {
  uint8_t *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

// A lookup table is built by a kernel on whatever stream the first frame came in on (context / user stream for device
// frames, the internal pipeline stream for host frames); a later call on ANOTHER stream must not read it before that build
// has finished: the build records an event, launches on other streams wait for it.
struct TableSync {
  cudaEvent_t ev = nullptr;
  cudaStream_t st = nullptr;
  bool valid = false;
};

struct b200vfx_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;   // default stream for device-pointer calls
  cudaStream_t user_stream = nullptr;
  bool use_user_stream = false;
  cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;  // host-pointer pipeline
  std::vector<cudaEvent_t> ev_in, ev_k;
  cudaEvent_t ev_order = nullptr;      // orders the internal pipeline stream after the context stream (mixed host/device calls)
  cudaEvent_t ev_order_back = nullptr; // ... and the context stream after the internal one (asynchronous mode)
  DevBuf stage_in, stage_out, stage_sums;
  // asynchronous host-frame mode (b200vfx_ctx_set_host_async): calls return once their copies and kernels are enqueued, so
  // frame i+1's upload overlaps frame i's download.  Two staging slots of their own; a slot is reused only after the
  // download of its previous frame (slot_done, recorded on s_d2h resp. s_k) -- the slot's upload and kernels wait for it.
  bool host_async = false;
  int async_slot = 0;
  DevBuf astage_in[2], astage_out[2];
  cudaEvent_t slot_done[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false};
  uint8_t *result_pinned = nullptr;   // small results (block sums, resized luma, histogram) come back through pinned memory
  DevBuf reduce_scratch;               // single-launch reductions: [0,64) two grid counters (kept zero between launches), then partials
  int chunk_rows = 0;
  int sm_count = 148;
  int l2_persist = 0;        // 1 = mark the 2^24-entry answer tables as L2-persisting (access-policy window per launch +
                             // device-wide set-aside).  OFF: measured harmful on B200 -- the set-aside halves the L2 left for
                             // the frames (frame A 11.5 -> 24.5 us) and does not help frame B, which is L1-gather-bound
                             // (profiles/r01_l2_persist_experiment.jsonl)
  size_t l2_persist_max = 0, l2_window_max = 0, l2_set_aside = 0;
  bool blockhash_tma = false; // videocompare block sums through the TMA-fed kernel (measured equal or slightly slower than the register-staged LDG kernel: profiles/r01_kernel_matrix.md)
  int blockhash_rows = 2;      // videocompare block sums: whole-row streaming kernel with 2 * SMs / value CTAs (2 = one CTA per SM, which
                               // leaves room for the next launch's CTAs to stream while this one drains: 6.4 vs 8.3 us per 4K frame);
                               // 0 = one CTA per hash block, the round-1 decomposition
  int zero_copy = 2;       // pinned host frames: TMA kernel reads/writes host memory directly; 0 never, 1 always, 2 auto-probe
  int zc_calls = 0, zc_bad_streak = 0; double zc_best_ms[2] = {1e30, 1e30};   // auto-probe state: [0] zero-copy, [1] staged
  int zc_cfg = 2, zc_ctas = 1, zc_grid = 96;
  int zc_hybrid = 0;       // experiment: staged path with one direction zero-copy (1: kernel stores to host, 2: kernel loads from host)  // stream-kernel variant / CTAs per SM / absolute grid cap for the zero-copy path
  int stream_grid = 0;     // absolute cap on the persistent grid of the stream kernels (0 = none)
  bool pdl = true;       // programmatic dependent launch for out-of-place frame kernels
  bool pdl_now = false;  // per launch: false when this call (re)built a table the kernel reads
  int stream_cfg = 0, stream_ctas = 0, stream_hint = 1, memo_px = 8;  // tuning knobs (env overrides, see ctx_create)
  uint64_t launches = 0;
  int tg_path = 0, tg_cfg = 0, tg_ctas = 8;   // fused tile gather: 0 register path (LDG/STG), 1 TMA; variant; CTAs per SM
  int memo_ctas = 4;     // CTAs per SM of the persistent table-lookup kernels: 4 x 256 threads = half the thread slots, so the next
                         // frame's kernel (PDL) is resident beside this one (profiles/r01_memo_ctas_experiment.jsonl)
  int rgba64_x4 = 1;     // RGBA64 + 3D LUT: 4 consecutive pixels per thread with the LUT cell cached in registers (0: one pixel per thread)
  int memo_tile = 0;     // 4-byte-pixel table lookups through memo_tile_kernel (per-tile shared-memory copy of the colour sub-cube)
  int cd_split = 0;      // colordetect, asynchronous calls: a launch takes 1/cd_split of the SMs so that consecutive launches overlap (0 = auto: 4 for quality <= 2, else 2; 1 = all SMs)
  int cd_cluster = 2;    // colordetect: CTAs per cluster merging their shared-memory histograms (1, 2, 4, 8)
  int peer_timeout_ms = 2000;  // deadline of the cross-GPU waits in the tile-gather kernel
  std::string err;
  // videocompare: Lanczos3 tap tables per (source length, output length), computed on the host once and kept on the device
  struct TapsDev { float *taps = nullptr; int2 *meta = nullptr; int max_taps = 0; };
  std::map<std::pair<int, int>, TapsDev> taps_cache;

  // colorlut state (State{lut}, colorlut/imp.rs:50-53)
  bool have_lut = false;
  int lut_kind = 0, lut_size = 0;
  int mode = 0;  // 0 auto (memo for u8), 1 direct
  int stream_path = -1;  // TMA-pipelined streaming kernels: -1 auto (1D memo: on; 3D memo: off, the gathers dominate and the
                         // plain LDG kernel measured faster, profiles/r01_sweep_memo.md), 0 off, 1 on
  float scale[3] = {1, 1, 1}, offset[3] = {0, 0, 0};
  LutPair *d_pair = nullptr;     // x-pair table (3D)
  float *d_lut1d = nullptr;
  uint4 *d_axis8 = nullptr;      // [3][256] axis table for RGBA (RGBA64 evaluates its axis entries per pixel)
  uint32_t *d_memo = nullptr;   // 2^24 x u32 (3D)
  uint8_t *d_memo1d = nullptr;  // 768 bytes (1D)
  bool memo_ready = false;
  TableSync ts_colorlut, ts_hf, ts_hd;   // axis + memo tables / hsvfilter table / hsvdetector bitmap

  // hsvfilter / hsvdetector memoisation (settings-keyed; built after 2^24 pixels with unchanged settings)
  int hsv_memo = -1;  // -1 auto (rent-or-buy), 0 never, 1 build on first use
  HsvFilterSettings hf_key{}; bool hf_key_valid = false, hf_ready = false; uint64_t hf_px_seen = 0; uint32_t *d_hf_memo = nullptr;
  HsvDetectSettings hd_key{}; bool hd_key_valid = false, hd_ready = false; uint64_t hd_px_seen = 0; uint32_t *d_hd_bitmap = nullptr;

  cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }
};

namespace {

int fail(b200vfx_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(ctx, call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ctx, B200VFX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));     \
  } while (0)

int table_built(b200vfx_ctx *c, TableSync &t, cudaStream_t st) {
  if (!t.ev) CU(c, cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming));
  CU(c, cudaEventRecord(t.ev, st));
  t.st = st; t.valid = true;
  return 0;
}
int table_wait(b200vfx_ctx *c, const TableSync &t, cudaStream_t st) {
  if (t.valid && t.st != st) CU(c, cudaStreamWaitEvent(st, t.ev, 0));
  return 0;
}

// A call that mixes host and device planes runs on the internal pipeline stream s_k, but its device planes belong to the
// context (or user) stream: whatever was enqueued there before this call -- e.g. the element upstream that produces the
// device frame -- must have finished before s_k touches them.
int order_after_ctx_stream(b200vfx_ctx *c, cudaStream_t st) {
  if (st == c->stream()) return 0;
  if (!c->ev_order) CU(c, cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
  CU(c, cudaEventRecord(c->ev_order, c->stream()));
  CU(c, cudaStreamWaitEvent(st, c->ev_order, 0));
  return 0;
}

// the reverse: what follows on the context stream is ordered after what `st` has been given so far (asynchronous host-frame
// mode with one plane in HBM: nothing synchronises the internal stream before the call returns)
int order_ctx_stream_after(b200vfx_ctx *c, cudaStream_t st) {
  if (st == c->stream()) return 0;
  if (!c->ev_order_back) CU(c, cudaEventCreateWithFlags(&c->ev_order_back, cudaEventDisableTiming));
  CU(c, cudaEventRecord(c->ev_order_back, st));
  CU(c, cudaStreamWaitEvent(c->stream(), c->ev_order_back, 0));
  return 0;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// page-locked host memory that the device can address directly (UVA): returns its device alias
bool pinned_device_ptr(const void *p, void **dev) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
  *dev = a.devicePointer;
  return true;
}

struct FmtInfo { int bpp, coff; bool bgr, alpha; };
bool fmt8_info(int fmt, FmtInfo *o) {
  switch (fmt) {
    case B200VFX_FORMAT_RGBX: *o = {4, 0, false, false}; return true;
    case B200VFX_FORMAT_RGBA: *o = {4, 0, false, true}; return true;
    case B200VFX_FORMAT_XRGB: *o = {4, 1, false, false}; return true;
    case B200VFX_FORMAT_ARGB: *o = {4, 1, false, true}; return true;
    case B200VFX_FORMAT_BGRX: *o = {4, 0, true, false}; return true;
    case B200VFX_FORMAT_BGRA: *o = {4, 0, true, true}; return true;
    case B200VFX_FORMAT_XBGR: *o = {4, 1, true, false}; return true;
    case B200VFX_FORMAT_ABGR: *o = {4, 1, true, true}; return true;
    case B200VFX_FORMAT_RGB: *o = {3, 0, false, false}; return true;
    case B200VFX_FORMAT_BGR: *o = {3, 0, true, false}; return true;
    default: return false;
  }
}

inline bool aligned(const void *p, long stride, int a) {
  return ((uintptr_t)p % (uintptr_t)a) == 0 && (stride % a) == 0;
}
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline unsigned grid_rows(int height) { return (unsigned)std::min(height, 65535); }
// compute-heavy 1-pixel-per-thread kernels loop over rows: size grid.y so that ~12 CTAs per SM exist and every
// thread amortises its per-CTA prologue (shared-memory tables) over many pixels
inline unsigned grid_rows_persistent(int grid_x, int height, int sm_count) {
  const long target = (long)sm_count * 12;
  long gy = std::max<long>(1, target / std::max(grid_x, 1));
  return (unsigned)std::min<long>(std::min<long>(gy, height), 65535);
}

LutDev lut_dev(const b200vfx_ctx *c) {
  LutDev L;
  L.pair = c->d_pair; L.lut1d = c->d_lut1d;
  L.axis = c->d_axis8;
  L.axis_len = 256;
  L.size = c->lut_size; L.kind = c->lut_kind;
  L.ident_domain = 1;
  L.neg_zero = -0.0f;
  for (int i = 0; i < 3; i++) {
    L.scale[i] = c->scale[i]; L.offset[i] = c->offset[i];
    if (!(c->scale[i] == 1.0f && c->offset[i] == 0.0f)) L.ident_domain = 0;   // -0.0 == 0.0: the default domain yields offset -0.0
  }
  return L;
}

int build_axis(b200vfx_ctx *c, cudaStream_t st) {
  uint4 **slot = &c->d_axis8;
  if (*slot) return 0;
  AxisBuildParams p;
  p.size = c->lut_size; p.kind = c->lut_kind; p.axis_len = 256;
  p.denom = 255.0f;   // norm_comp (imp.rs:471-474)
  for (int i = 0; i < 3; i++) { p.scale[i] = c->scale[i]; p.offset[i] = c->offset[i]; }
  CU(c, cudaMalloc(slot, (size_t)3 * p.axis_len * sizeof(uint4)));
  colorlut_axis_table_kernel<<<ceil_div(3 * p.axis_len, 256), 256, 0, st>>>(*slot, p);
  c->launches++;
  CU(c, cudaGetLastError());
  return 0;
}

// L2 residency control for the big answer tables: the launch carries an access-policy window that marks table lines
// "persisting" (they live in the L2 set-aside reserved with cudaLimitPersistingL2CacheSize) while everything else of
// the launch (the frame) is "streaming".  thread_local: set around one launch by with_l2_window().
thread_local cudaAccessPolicyWindow g_l2_window = {};
struct with_l2_window {
  with_l2_window(const b200vfx_ctx *c, const void *table, size_t bytes);
  ~with_l2_window() { g_l2_window = cudaAccessPolicyWindow{}; }
};

with_l2_window::with_l2_window(const b200vfx_ctx *c, const void *table, size_t bytes) {
  g_l2_window = cudaAccessPolicyWindow{};
  if (!c->l2_persist || !table || c->l2_window_max == 0 || c->l2_set_aside == 0) return;
  g_l2_window.base_ptr = const_cast<void *>(table);
  g_l2_window.num_bytes = std::min(bytes, c->l2_window_max);
  g_l2_window.hitRatio = (float)std::min(1.0, (double)c->l2_set_aside / (double)g_l2_window.num_bytes);
  g_l2_window.hitProp = cudaAccessPropertyPersisting;
  g_l2_window.missProp = cudaAccessPropertyStreaming;
}

// reserve L2 set-aside for persisting lines (device-wide limit; only raised, never lowered, by this library)
int ensure_l2_set_aside(b200vfx_ctx *c, size_t want) {
  if (!c->l2_persist || c->l2_persist_max == 0) return 0;
  const size_t target = std::min(c->l2_persist_max, want);
  size_t cur = 0;
  CU(c, cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize));
  if (cur < target) CU(c, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, target));
  c->l2_set_aside = std::max(cur, target);
  return 0;
}

// Programmatic dependent launch (PDL): our out-of-place frame kernels do not depend on the previous frame's kernel,
// so they are launched with programmaticStreamSerialization and trigger `griddepcontrol.launch_dependents` at
// entry: the next frame's CTAs fill the SMs while this frame's tail drains (same stream, no extra streams).
// Foreign kernels (which never trigger early) and plain launches/copies still wait for full completion.
// PDL overlap is only used when the new kernel cannot conflict with any of OUR recent (possibly still running)
// kernels on the same stream: its output must not touch their inputs/outputs and its input must not be one of
// their outputs.  The tracker is process-wide (elements chained on one stream use different contexts).
// Which of OUR earlier launches on a stream may still be running when a new PDL launch starts?
//  * Persistent kernels with a capped grid (large frames) end on griddepcontrol.wait ("linger"): they complete in launch
//    order and their CTAs keep their SM slots until the previous kernel has completed.  Kernel i starts only after i-1 is
//    fully resident; if an older kernel j is still incomplete, ALL CTAs of j+1 .. i-1 are still resident, so their thread
//    counts sum to less than what the device holds.  Walking back from the newest launch, the point where that sum reaches
//    the device's thread capacity ends the list of possibly-running launches (2 launches for 4 CTAs per SM).
//  * Anything else (small frames, non-persistent kernels) does not linger -- the wait costs ~3 us of completion latency per
//    kernel, which small frames cannot hide.  Nothing bounds how many of those can be co-resident (they trigger at entry and
//    complete out of order), so EVERY such launch since the last plain launch (= full barrier) counts as possibly running;
//    the record holds kPdlQueueMax launches and a full record turns the next launch into a plain one.
constexpr size_t kPdlQueueMax = 16;
struct Span { uintptr_t lo, hi; };
struct RecentLaunch { Span src, dst; long long threads; bool lingers; };
std::mutex g_recent_mu;
std::map<cudaStream_t, std::deque<RecentLaunch>> g_recent;
long long g_capacity_threads = 148LL * 2048;   // largest (SM count x resident threads per SM) of the devices in use

inline Span span_of(const void *p, long stride, size_t row_bytes, int rows) {
  if (!p || rows <= 0) return Span{0, 0};
  const uintptr_t lo = (uintptr_t)p;
  return Span{lo, lo + (size_t)(rows - 1) * (size_t)stride + row_bytes};
}
inline bool overlap(Span a, Span b) { return a.lo < b.hi && b.lo < a.hi; }

// returns whether the launch may use PDL, and records it
// dst_after_wait: the kernel writes dst only after its griddepcontrol.wait, i.e. after every earlier launch has completed --
// its writes cannot race with them, only its early reads of src can
bool pdl_admit(bool want, cudaStream_t st, Span src, Span dst, bool dst_after_wait = false) {
  std::lock_guard<std::mutex> g(g_recent_mu);
  std::deque<RecentLaunch> &q = g_recent[st];
  bool ok = want;
  if (ok) {
    // Walking back from the newest launch.  Launch r has provably completed when every launch newer than r lingers (those
    // complete in launch order and hold their slots until their predecessor has completed) and together they fill the
    // device: were r still running, none of them could have retired a CTA.  A launch that does not linger gives no
    // such bound for anything older than itself: those stay "possibly running" until the next plain launch (= barrier).
    long long newer = 0;       // threads of the launches newer than the one being examined
    bool newer_linger = true;  // ... and whether all of them linger
    size_t oldest_unknown = q.size();   // index of the oldest launch that could not be proven complete
    for (size_t i = q.size(); i-- > 0;) {
      const RecentLaunch &r = q[i];
      if (!(newer_linger && newer >= g_capacity_threads)) {
        if ((!dst_after_wait && (overlap(dst, r.src) || overlap(dst, r.dst))) || overlap(src, r.dst)) { ok = false; break; }
        oldest_unknown = i;
      }
      newer += r.threads;
      newer_linger = newer_linger && r.lingers;
    }
    if (ok && oldest_unknown > 0 && oldest_unknown <= q.size()) q.erase(q.begin(), q.begin() + (long)std::min(oldest_unknown, q.size()));
    // the record is bounded: when it is full the new launch becomes a plain one (a barrier), which empties it
    if (ok && q.size() >= kPdlQueueMax) ok = false;
  }
  if (!ok) q.clear();  // a normal launch starts only after everything before it has completed
  q.push_back(RecentLaunch{src, dst, 0, false});
  return ok;
}
void pdl_note_linger(cudaStream_t st) {   // the launch just admitted ends on griddepcontrol.wait
  std::lock_guard<std::mutex> g(g_recent_mu);
  auto it = g_recent.find(st);
  if (it != g_recent.end() && !it->second.empty()) it->second.back().lingers = true;
}
// size of the launch just admitted (called by launch_k)
void pdl_note_threads(cudaStream_t st, long long threads) {
  std::lock_guard<std::mutex> g(g_recent_mu);
  auto it = g_recent.find(st);
  if (it != g_recent.end() && !it->second.empty()) it->second.back().threads = threads;
}
void pdl_forget(cudaStream_t st) {  // after a stream synchronisation nothing of ours is in flight
  std::lock_guard<std::mutex> g(g_recent_mu);
  g_recent.erase(st);
}

template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  pdl_note_threads(st, (long long)grid.x * grid.y * grid.z * block.x * block.y * block.z);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (g_l2_window.num_bytes) {  // lookup table of this launch: L2 persisting access window (set by with_l2_window)
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow = g_l2_window;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- per-element kernel launchers on DEVICE frames ------------------------------------------
struct Frame { const uint8_t *src; long sstride; uint8_t *dst; long dstride; int width, height; };

// TMA streaming kernel variants (tile bytes, stages, threads, gathers in flight per thread); the
// default was picked from the sweep in profiles/ (B200VFX_STREAM_CFG / _CTAS / _HINT override it for A/B runs)
template <int TILE, int STAGES, int THREADS, int B>
int launch_memo_stream_t(b200vfx_ctx *c, const uint8_t *src, long ss, uint8_t *dst, long ds, int row_bytes, int h,
                         cudaStream_t st) {
  constexpr int smem = stream_smem_bytes<TILE, STAGES>();
  auto k3 = colorlut_memo_stream_kernel<TILE, STAGES, THREADS, B>;
  auto k1 = colorlut_memo1d_stream_kernel<TILE, STAGES, THREADS, B>;
  // function attributes are per device: remember per (kernel variant, device) whether they are set and what the
  // occupancy query returned (both are driver calls we do not want on every frame)
  static std::mutex mu;
  static int per_sm_cache[64] = {0};
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> g(mu);
    const int dev = (c->device >= 0 && c->device < 64) ? c->device : 0;
    if (per_sm_cache[dev] == 0) {
      CU(c, cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CU(c, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int n = 0;
      CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k3, THREADS, smem));
      per_sm_cache[dev] = n < 1 ? 1 : n;
    }
    per_sm = per_sm_cache[dev];
  }
  if (c->stream_ctas > 0) per_sm = std::min(per_sm, c->stream_ctas);
  const long long ntiles = (long long)ceil_div(row_bytes, TILE) * h;
  long long gmax = (long long)c->sm_count * per_sm;
  if (c->stream_grid > 0) gmax = std::min<long long>(gmax, c->stream_grid);
  const unsigned grid = (unsigned)std::min<long long>(ntiles, gmax);
  if (c->lut_kind == 3) CU(c, launch_k(c->pdl_now, k3, dim3(grid), dim3(THREADS), smem, st, c->d_memo, c->stream_hint, src, ss, dst, ds, row_bytes, h));
  else CU(c, launch_k(c->pdl_now, k1, dim3(grid), dim3(THREADS), smem, st, c->d_memo1d, src, ss, dst, ds, row_bytes, h));
  return 0;
}

int launch_memo_stream(b200vfx_ctx *c, const uint8_t *src, long ss, uint8_t *dst, long ds, int row_bytes, int h,
                       cudaStream_t st) {
  switch (c->stream_cfg) {
    case 1: return launch_memo_stream_t<8192, 4, 256, 8>(c, src, ss, dst, ds, row_bytes, h, st);
    case 2: return launch_memo_stream_t<4096, 4, 256, 4>(c, src, ss, dst, ds, row_bytes, h, st);
    case 3: return launch_memo_stream_t<4096, 8, 128, 8>(c, src, ss, dst, ds, row_bytes, h, st);
    case 4: return launch_memo_stream_t<2048, 8, 128, 4>(c, src, ss, dst, ds, row_bytes, h, st);
    case 5: return launch_memo_stream_t<8192, 6, 512, 4>(c, src, ss, dst, ds, row_bytes, h, st);
    case 6: return launch_memo_stream_t<4096, 6, 512, 2>(c, src, ss, dst, ds, row_bytes, h, st);
    case 7: return launch_memo_stream_t<2048, 6, 256, 2>(c, src, ss, dst, ds, row_bytes, h, st);
    default: return launch_memo_stream_t<16384, 4, 256, 8>(c, src, ss, dst, ds, row_bytes, h, st);
  }
}

// once per LUT: evaluate all 2^24 colours (3D) / 3x256 channel values (1D) with the exact direct evaluator
int ensure_colorlut_memo(b200vfx_ctx *c, const LutDev &p, cudaStream_t st) {
  if (c->memo_ready) return table_wait(c, c->ts_colorlut, st);
  if (c->lut_kind == 3) {
    if (int rc = ensure_l2_set_aside(c, (size_t)72 << 20)) return rc;
    if (!c->d_memo) CU(c, cudaMalloc(&c->d_memo, sizeof(uint32_t) << 24));
    colorlut_memo_build_kernel<<<(1u << 24) / 256, 256, 0, st>>>(p, c->d_memo);
  } else {
    if (!c->d_memo1d) CU(c, cudaMalloc(&c->d_memo1d, 768));
    colorlut_memo1d_build_kernel<<<3, 256, 0, st>>>(p, c->d_memo1d);
  }
  c->launches++;
  CU(c, cudaGetLastError());
  c->memo_ready = true;
  return table_built(c, c->ts_colorlut, st);
}

// fused tile gather, TMA variant: per-(variant, device) attribute / occupancy cache like launch_memo_stream_t
template <int TILE, int STAGES, int THREADS, int B>
int launch_tile_gather_tma_t(b200vfx_ctx *c, bool l1d, const uint8_t *src, long ss, const PeerSet &ps, long ds, long doff,
                             int row_bytes, int h, cudaStream_t st) {
  constexpr int smem = stream_smem_bytes<TILE, STAGES>();
  auto k3 = colorlut_tile_gather_tma_kernel<TILE, STAGES, THREADS, B, false>;
  auto k1 = colorlut_tile_gather_tma_kernel<TILE, STAGES, THREADS, B, true>;
  static std::mutex mu;
  static int per_sm_cache[64] = {0};
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> g(mu);
    const int dev = (c->device >= 0 && c->device < 64) ? c->device : 0;
    if (per_sm_cache[dev] == 0) {
      CU(c, cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      CU(c, cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      int n = 0;
      CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k3, THREADS, smem));
      per_sm_cache[dev] = n < 1 ? 1 : n;
    }
    per_sm = per_sm_cache[dev];
  }
  if (c->tg_ctas > 0) per_sm = std::min(per_sm, c->tg_ctas);
  const long long ntiles = (long long)std::max(1, ceil_div(row_bytes, TILE)) * h;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ntiles, (long long)c->sm_count * per_sm));
  if (l1d) CU(c, launch_k(false, k1, dim3(grid), dim3(THREADS), smem, st, c->d_memo, c->d_memo1d, src, ss, ps, ds, doff, row_bytes, h));
  else CU(c, launch_k(false, k3, dim3(grid), dim3(THREADS), smem, st, c->d_memo, c->d_memo1d, src, ss, ps, ds, doff, row_bytes, h));
  return 0;
}
int launch_tile_gather_tma(b200vfx_ctx *c, bool l1d, const uint8_t *src, long ss, const PeerSet &ps, long ds, long doff,
                           int row_bytes, int h, cudaStream_t st) {
  switch (c->tg_cfg) {
    case 1: return launch_tile_gather_tma_t<8192, 4, 256, 8>(c, l1d, src, ss, ps, ds, doff, row_bytes, h, st);
    case 2: return launch_tile_gather_tma_t<4096, 4, 256, 4>(c, l1d, src, ss, ps, ds, doff, row_bytes, h, st);
    case 3: return launch_tile_gather_tma_t<32768, 3, 512, 8>(c, l1d, src, ss, ps, ds, doff, row_bytes, h, st);
    default: return launch_tile_gather_tma_t<16384, 4, 256, 8>(c, l1d, src, ss, ps, ds, doff, row_bytes, h, st);
  }
}

int launch_colorlut(b200vfx_ctx *c, int fmt, const Frame &f, cudaStream_t st) {
  if (f.width == 0 || f.height == 0) return 0;
  const bool wide = fmt != B200VFX_FORMAT_RGBA;
  {
    const size_t rb = (size_t)f.width * (wide ? 8 : 4);
    const bool tables_ready = c->d_axis8 != nullptr && (fmt != B200VFX_FORMAT_RGBA || c->mode != 0 || c->memo_ready);
    c->pdl_now = pdl_admit(c->pdl && tables_ready, st, span_of(f.src, f.sstride, rb, f.height), span_of(f.dst, f.dstride, rb, f.height));
  }
  if (int rc = build_axis(c, st)) return rc;
  const LutDev p = lut_dev(c);
  if (fmt == B200VFX_FORMAT_RGBA && c->mode == 0) {
    if (int rc = ensure_colorlut_memo(c, p, st)) return rc;
    const bool al = aligned(f.src, f.sstride, 4) && aligned(f.dst, f.dstride, 4);
    with_l2_window l2w(c, c->lut_kind == 3 ? c->d_memo : nullptr, sizeof(uint32_t) << 24);
    int w = f.width, h = f.height;
    long ss = f.sstride, ds = f.dstride;
    if (c->memo_tile && c->lut_kind == 3 && (w % 4) == 0 && aligned(f.src, ss, 16) && aligned(f.dst, ds, 16)) {
      dim3 grid((unsigned)ceil_div(w, kTileW), (unsigned)ceil_div(h, kTileH));
      CU(c, launch_k(c->pdl_now, memo_tile_kernel<0, false>, grid, dim3(256), 0, st, c->d_memo, f.src, ss, f.dst, ds, w, h));
      c->launches++;
      CU(c, cudaGetLastError());
      return 0;
    }
    if (al && ss == 4L * w && ds == 4L * w && (long long)w * h < (1LL << 28)) { w = w * h; h = 1; }  // packed: 1-D
    const bool al16 = aligned(f.src, ss, 16) && aligned(f.dst, ds, 16) && (w % 4) == 0;
    const bool use_stream = c->stream_path == 1 || (c->stream_path < 0 && c->lut_kind == 1);
    if (al16 && use_stream) {  // TMA-pipelined streaming kernel (the normal case for GStreamer buffers)
      if (int rc = launch_memo_stream(c, f.src, ss, f.dst, ds, 4 * w, h, st)) return rc;
    } else if (al) {
      const int PX = c->memo_px;
      const long long items = (long long)ceil_div(w, 8 * 32 * PX) * h, cap = (long long)c->sm_count * c->memo_ctas;
      dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(items, cap)));
      const int linger = (c->pdl_now && items > cap) ? 1 : 0;   // capped grid: consecutive frames overlap deeply
      if (linger) pdl_note_linger(st);
#define LAUNCH_PLAIN(P)                                                                                               \
  do {                                                                                                                \
    if (c->lut_kind == 3) CU(c, launch_k(c->pdl_now, colorlut_memo_apply_kernel<P>, grid, dim3(256), 0, st, c->d_memo, f.src, ss, f.dst, ds, w, h, linger));  \
    else CU(c, launch_k(c->pdl_now, colorlut_memo1d_apply_kernel<P>, grid, dim3(256), 0, st, c->d_memo1d, f.src, ss, f.dst, ds, w, h, linger));      \
  } while (0)
      if (PX == 16) LAUNCH_PLAIN(16); else if (PX == 8) LAUNCH_PLAIN(8); else LAUNCH_PLAIN(4);
#undef LAUNCH_PLAIN
    } else {
      dim3 grid((unsigned)ceil_div(w, 256), grid_rows(h));
      colorlut_memo_apply_bytes_kernel<<<grid, 256, 0, st>>>(c->lut_kind == 3 ? c->d_memo : nullptr, c->d_memo1d,
                                                             f.src, ss, f.dst, ds, w, h);
    }
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  }
  dim3 grid((unsigned)ceil_div(f.width, 256), grid_rows_persistent(ceil_div(f.width, 256), f.height, c->sm_count));
  const bool al4 = aligned(f.src, f.sstride, 4) && aligned(f.dst, f.dstride, 4);
  const bool al8 = aligned(f.src, f.sstride, 8) && aligned(f.dst, f.dstride, 8);
  if (((uintptr_t)f.src | (uintptr_t)f.dst) & 1u && fmt != B200VFX_FORMAT_RGBA)
    return fail(c, B200VFX_ERR_INVALID, "RGBA64 planes must be 2-byte aligned (as_slice_of::<u16>, imp.rs:323-324)");
#define LAUNCH_DIRECT(F, A) CU(c, launch_k(c->pdl_now, colorlut_direct_kernel<F, A>, grid, dim3(256), 0, st, p, f.src, f.sstride, f.dst, f.dstride, f.width, f.height))
  if (wide && c->lut_kind == 3 && c->rgba64_x4 && (f.width % 4) == 0 && aligned(f.src, f.sstride, 32) && aligned(f.dst, f.dstride, 32)) {
    // four consecutive pixels per thread, LUT cell entries cached in registers (colorlut_direct64x4_kernel)
    const int gx = ceil_div(f.width / 4, 256);
    dim3 g4((unsigned)gx, grid_rows_persistent(gx, f.height, c->sm_count));
    if (fmt == B200VFX_FORMAT_RGBA64_LE) CU(c, launch_k(c->pdl_now, colorlut_direct64x4_kernel<1>, g4, dim3(256), 0, st, p, f.src, f.sstride, f.dst, f.dstride, f.width, f.height));
    else CU(c, launch_k(c->pdl_now, colorlut_direct64x4_kernel<2>, g4, dim3(256), 0, st, p, f.src, f.sstride, f.dst, f.dstride, f.width, f.height));
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  }
  switch (fmt) {
    case B200VFX_FORMAT_RGBA: if (al4) LAUNCH_DIRECT(0, true); else LAUNCH_DIRECT(0, false); break;
    case B200VFX_FORMAT_RGBA64_LE: if (al8) LAUNCH_DIRECT(1, true); else LAUNCH_DIRECT(1, false); break;
    case B200VFX_FORMAT_RGBA64_BE: if (al8) LAUNCH_DIRECT(2, true); else LAUNCH_DIRECT(2, false); break;
    default: return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut: unsupported format %d", fmt);
  }
#undef LAUNCH_DIRECT
  c->launches++;
  CU(c, cudaGetLastError());
  return 0;
}

#ifndef B200VFX_MAP_PX
#define B200VFX_MAP_PX 8
#endif
constexpr int kMapPx = B200VFX_MAP_PX;   // pixels (independent gathers) per thread of map_u32_kernel

// rent-or-buy: build the answer table once the current settings have processed as many pixels as the table has entries
bool hsv_memo_decide(int option, bool key_same, uint64_t &px_seen, uint64_t npx, bool &ready) {
  if (!key_same) { ready = false; px_seen = 0; }
  px_seen += npx;
  if (option == 0) return false;
  return ready || option == 1 || px_seen >= (1ull << 24);
}

constexpr int kHsvDirectPx = 4;   // independent pixels per thread of hsv_direct_map_kernel

template <int BPP, int COFF, bool BGR>
void launch_hsvfilter_t(const HsvFilterSettings &s, const uint32_t *memo, uint8_t *data, long stride, int w, int h,
                        cudaStream_t st, int sm_count) {
  const bool al = BPP == 4 && aligned(data, stride, 4);
  const int cls = hsvf_shift_class(s.hue_shift);
  if (memo) {
    dim3 grid((unsigned)ceil_div(w, 256), grid_rows(h));
    if (al) hsvfilter_kernel<BPP, COFF, BGR, true, true><<<grid, 256, 0, st>>>(s, cls, memo, data, stride, w, h);
    else hsvfilter_kernel<BPP, COFF, BGR, false, true><<<grid, 256, 0, st>>>(s, cls, memo, data, stride, w, h);
  } else if (al) {   // 4-byte pixels: lane-consecutive map kernel, kHsvDirectPx independent pixels per thread
    int ww = w, hh = h;
    if (stride == 4L * w && (long long)w * h < (1LL << 28)) { ww = w * h; hh = 1; }
    const long long items = (long long)ceil_div(ww, 8 * 32 * kHsvDirectPx) * hh;
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(items, (long long)sm_count * 8)));
    if (cls) hsv_direct_map_kernel<HsvFilterDirectOp<BPP == 4 ? COFF : 0, BGR, 1>, kHsvDirectPx><<<grid, 256, 0, st>>>(
        HsvFilterDirectOp<BPP == 4 ? COFF : 0, BGR, 1>{s, 1.0f, -0.0f}, data, stride, data, stride, ww, hh);
    else hsv_direct_map_kernel<HsvFilterDirectOp<BPP == 4 ? COFF : 0, BGR, 0>, kHsvDirectPx><<<grid, 256, 0, st>>>(
        HsvFilterDirectOp<BPP == 4 ? COFF : 0, BGR, 0>{s, 1.0f, -0.0f}, data, stride, data, stride, ww, hh);
  } else {
    dim3 grid((unsigned)ceil_div(w, 256), grid_rows_persistent(ceil_div(w, 256), h, sm_count));
    hsvfilter_kernel<BPP, COFF, BGR, false, false><<<grid, 256, 0, st>>>(s, cls, nullptr, data, stride, w, h);
  }
}

int launch_hsvfilter(b200vfx_ctx *c, const FmtInfo &fi, const HsvFilterSettings &s, uint8_t *data, long stride,
                     int w, int h, cudaStream_t st) {
  if (w == 0 || h == 0) return 0;
  const bool same = c->hf_key_valid && std::memcmp(&c->hf_key, &s, sizeof s) == 0;
  const bool use_memo = hsv_memo_decide(c->hsv_memo, same, c->hf_px_seen, (uint64_t)w * h, c->hf_ready);
  c->hf_key = s; c->hf_key_valid = true;
  const uint32_t *memo = nullptr;
  bool built_now = false;
  if (use_memo) {
    if (!c->hf_ready) {
      if (int rc = ensure_l2_set_aside(c, (size_t)72 << 20)) return rc;
      if (!c->d_hf_memo) CU(c, cudaMalloc(&c->d_hf_memo, sizeof(uint32_t) << 24));
      pdl_admit(false, st, Span{0, 0}, Span{0, 0});  // plain launch
      hsvfilter_memo_build_kernel<<<(unsigned)c->sm_count * 8, 256, 0, st>>>(s, hsvf_shift_class(s.hue_shift), c->d_hf_memo);
      c->launches++;
      CU(c, cudaGetLastError());
      c->hf_ready = true;
      built_now = true;
      if (int rc = table_built(c, c->ts_hf, st)) return rc;
    } else if (int rc = table_wait(c, c->ts_hf, st)) return rc;
    memo = c->d_hf_memo;
  }
  if (memo && fi.bpp == 4 && aligned(data, stride, 4)) {  // table-lookup map kernel, PDL-overlapped when frames are disjoint
    int ww = w, hh = h;
    long ss = stride;
    if (ss == 4L * w && (long long)w * h < (1LL << 28)) { ww = w * h; hh = 1; }
    const Span sp = span_of(data, stride, (size_t)w * 4, h);
    const bool pdl = pdl_admit(c->pdl && !built_now, st, sp, sp);
    if (c->memo_tile && (w % 4) == 0 && aligned(data, stride, 16)) {
      dim3 tgrid((unsigned)ceil_div(w, kTileW), (unsigned)ceil_div(h, kTileH));
#define LT(CO, BG) CU(c, launch_k(pdl, memo_tile_kernel<CO, BG>, tgrid, dim3(256), 0, st, memo, (const uint8_t *)data, stride, data, stride, w, h))
      if (fi.coff == 0) { if (fi.bgr) LT(0, true); else LT(0, false); }
      else { if (fi.bgr) LT(1, true); else LT(1, false); }
#undef LT
      c->launches++;
      CU(c, cudaGetLastError());
      return 0;
    }
    with_l2_window l2w(c, memo, sizeof(uint32_t) << 24);
    const long long items = (long long)ceil_div(ww, 8 * 32 * kMapPx) * hh, cap = (long long)c->sm_count * c->memo_ctas;
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(items, cap)));
    const int linger = (pdl && items > cap) ? 1 : 0;
    if (linger) pdl_note_linger(st);
#define LM(CO, BG) CU(c, launch_k(pdl, map_u32_kernel<HsvFilterMemoOp<CO, BG>, kMapPx>, grid, dim3(256), 0, st, HsvFilterMemoOp<CO, BG>{memo}, (const uint8_t *)data, ss, data, ss, ww, hh, linger))
    if (fi.coff == 0) { if (fi.bgr) LM(0, true); else LM(0, false); }
    else { if (fi.bgr) LM(1, true); else LM(1, false); }
#undef LM
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  }
  if (memo && fi.bpp == 3 && aligned(data, stride, 4)) {  // RGB / BGR through the same table: map_rgb24_kernel
    int ww = w, hh = h;
    long ss = stride;
    if (ss == 3L * w && (3LL * w * h) % 4 == 0 && (long long)w * h < (1LL << 28)) { ww = w * h; hh = 1; }
    const Span sp = span_of(data, stride, (size_t)w * 3, h);
    const bool pdl = pdl_admit(c->pdl && !built_now, st, sp, sp);
    constexpr int G = 2;
    dim3 grid((unsigned)ceil_div(ceil_div(ww, 4), 32 * G * 8), grid_rows(hh));
    if (fi.bgr) CU(c, launch_k(pdl, map_rgb24_kernel<HsvFilterMemoOp<0, true>, G, false, false>, grid, dim3(256), 0, st, HsvFilterMemoOp<0, true>{memo}, (const uint8_t *)data, ss, data, ss, ww, hh));
    else CU(c, launch_k(pdl, map_rgb24_kernel<HsvFilterMemoOp<0, false>, G, false, false>, grid, dim3(256), 0, st, HsvFilterMemoOp<0, false>{memo}, (const uint8_t *)data, ss, data, ss, ww, hh));
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  }
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});  // plain launch: waits for, and is waited on by, everything around it
  if (fi.bpp == 3) { if (fi.bgr) launch_hsvfilter_t<3, 0, true>(s, memo, data, stride, w, h, st, c->sm_count); else launch_hsvfilter_t<3, 0, false>(s, memo, data, stride, w, h, st, c->sm_count); }
  else if (fi.coff == 0) { if (fi.bgr) launch_hsvfilter_t<4, 0, true>(s, memo, data, stride, w, h, st, c->sm_count); else launch_hsvfilter_t<4, 0, false>(s, memo, data, stride, w, h, st, c->sm_count); }
  else { if (fi.bgr) launch_hsvfilter_t<4, 1, true>(s, memo, data, stride, w, h, st, c->sm_count); else launch_hsvfilter_t<4, 1, false>(s, memo, data, stride, w, h, st, c->sm_count); }
  c->launches++;
  CU(c, cudaGetLastError());
  return 0;
}

template <int IBPP, int ICOFF, bool IBGR>
int launch_hsvdetector_t(b200vfx_ctx *c, const FmtInfo &fo, const HsvDetectSettings &s, const uint32_t *bitmap, bool pdl,
                         const Frame &f, cudaStream_t st) {
  const bool al = aligned(f.dst, f.dstride, 4) && (IBPP == 3 || aligned(f.src, f.sstride, 4));
  const int cls = hsvf_shift_class(180.0f - s.hue_ref);
  const int gx = ceil_div(f.width, 256);
  dim3 grid((unsigned)gx, bitmap ? grid_rows(f.height) : grid_rows_persistent(gx, f.height, c->sm_count));
#define L(OC, OB)                                                                                                          \
  do {                                                                                                                     \
    if (bitmap) {                                                                                                          \
      if (al) CU(c, launch_k(pdl, hsvdetector_kernel<IBPP, ICOFF, IBGR, OC, OB, true, true>, grid, dim3(256), 0, st, s, cls, bitmap, f.src, f.sstride, f.dst, f.dstride, f.width, f.height)); \
      else CU(c, launch_k(pdl, hsvdetector_kernel<IBPP, ICOFF, IBGR, OC, OB, false, true>, grid, dim3(256), 0, st, s, cls, bitmap, f.src, f.sstride, f.dst, f.dstride, f.width, f.height)); \
    } else {                                                                                                               \
      if (al && IBPP == 4) {                                                                                               \
        int ww = f.width, hh = f.height;                                                                                   \
        if (f.sstride == 4L * ww && f.dstride == 4L * ww && (long long)ww * hh < (1LL << 28)) { ww = ww * hh; hh = 1; }    \
        const long long items = (long long)ceil_div(ww, 8 * 32 * kHsvDirectPx) * hh;                                       \
        dim3 g2((unsigned)std::max<long long>(1, std::min<long long>(items, (long long)c->sm_count * 8)));                 \
        if (cls) hsv_direct_map_kernel<HsvDetectDirectOp<IBPP == 4 ? ICOFF : 0, IBGR, OC, OB, 1>, kHsvDirectPx><<<g2, 256, 0, st>>>( \
            HsvDetectDirectOp<IBPP == 4 ? ICOFF : 0, IBGR, OC, OB, 1>{s, 1.0f}, f.src, f.sstride, f.dst, f.dstride, ww, hh);     \
        else hsv_direct_map_kernel<HsvDetectDirectOp<IBPP == 4 ? ICOFF : 0, IBGR, OC, OB, 0>, kHsvDirectPx><<<g2, 256, 0, st>>>( \
            HsvDetectDirectOp<IBPP == 4 ? ICOFF : 0, IBGR, OC, OB, 0>{s, 1.0f}, f.src, f.sstride, f.dst, f.dstride, ww, hh);     \
      } else if (al) hsvdetector_kernel<IBPP, ICOFF, IBGR, OC, OB, true, false><<<grid, 256, 0, st>>>(s, cls, nullptr, f.src, f.sstride, f.dst, f.dstride, f.width, f.height); \
      else hsvdetector_kernel<IBPP, ICOFF, IBGR, OC, OB, false, false><<<grid, 256, 0, st>>>(s, cls, nullptr, f.src, f.sstride, f.dst, f.dstride, f.width, f.height);  \
    }                                                                                                                      \
  } while (0)
  if (fo.coff == 0) { if (fo.bgr) L(0, true); else L(0, false); }
  else { if (fo.bgr) L(1, true); else L(1, false); }
#undef L
  return 0;
}

int launch_hsvdetector(b200vfx_ctx *c, const FmtInfo &fi, const FmtInfo &fo, const HsvDetectSettings &s,
                       const Frame &f, cudaStream_t st) {
  if (f.width == 0 || f.height == 0) return 0;
  const bool same = c->hd_key_valid && std::memcmp(&c->hd_key, &s, sizeof s) == 0;
  const bool use_memo = hsv_memo_decide(c->hsv_memo, same, c->hd_px_seen, (uint64_t)f.width * f.height, c->hd_ready);
  c->hd_key = s; c->hd_key_valid = true;
  const uint32_t *bitmap = nullptr;
  bool pdl = false;
  if (use_memo) {
    bool built_now = false;
    if (!c->hd_ready) {
      if (!c->d_hd_bitmap) CU(c, cudaMalloc(&c->d_hd_bitmap, (1u << 24) / 8));
      pdl_admit(false, st, Span{0, 0}, Span{0, 0});
      hsvdetector_bitmap_build_kernel<<<1024, 256, 0, st>>>(s, hsvf_shift_class(180.0f - s.hue_ref), c->d_hd_bitmap);
      c->launches++;
      CU(c, cudaGetLastError());
      c->hd_ready = true;
      built_now = true;
      if (int rc = table_built(c, c->ts_hd, st)) return rc;
    } else if (int rc = table_wait(c, c->ts_hd, st)) return rc;
    bitmap = c->d_hd_bitmap;
    pdl = pdl_admit(c->pdl && !built_now, st, span_of(f.src, f.sstride, (size_t)f.width * fi.bpp, f.height),
                    span_of(f.dst, f.dstride, (size_t)f.width * 4, f.height));
    if (fi.bpp == 4 && aligned(f.src, f.sstride, 4) && aligned(f.dst, f.dstride, 4)) {  // table-lookup map kernel
      int ww = f.width, hh = f.height;
      if (f.sstride == 4L * ww && f.dstride == 4L * ww && (long long)ww * hh < (1LL << 28)) { ww = ww * hh; hh = 1; }
      const long long items = (long long)ceil_div(ww, 8 * 32 * kMapPx) * hh, cap = (long long)c->sm_count * c->memo_ctas;
      dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(items, cap)));
      const int linger = (pdl && items > cap) ? 1 : 0;
      if (linger) pdl_note_linger(st);
#define LD(IC, IB, OC, OB) CU(c, launch_k(pdl, map_u32_kernel<HsvDetectBitmapOp<IC, IB, OC, OB>, kMapPx>, grid, dim3(256), 0, st, HsvDetectBitmapOp<IC, IB, OC, OB>{bitmap}, f.src, f.sstride, f.dst, f.dstride, ww, hh, linger))
#define LD2(IC, IB) do { if (fo.coff == 0) { if (fo.bgr) LD(IC, IB, 0, true); else LD(IC, IB, 0, false); } else { if (fo.bgr) LD(IC, IB, 1, true); else LD(IC, IB, 1, false); } } while (0)
      if (fi.coff == 0) { if (fi.bgr) LD2(0, true); else LD2(0, false); }
      else { if (fi.bgr) LD2(1, true); else LD2(1, false); }
#undef LD2
#undef LD
      c->launches++;
      CU(c, cudaGetLastError());
      return 0;
    }
    if (fi.bpp == 3 && aligned(f.src, f.sstride, 4) && aligned(f.dst, f.dstride, 4)) {  // RGB / BGR in: map_rgb24_kernel
      int ww = f.width, hh = f.height;
      if (f.sstride == 3L * ww && f.dstride == 4L * ww && (3LL * ww * hh) % 4 == 0 && (long long)ww * hh < (1LL << 28)) { ww = ww * hh; hh = 1; }
      constexpr int G = 2;
      dim3 grid((unsigned)ceil_div(ceil_div(ww, 4), 32 * G * 8), grid_rows(hh));
      const bool d16 = aligned(f.dst, f.dstride, 16);
#define LD3(IB, OC, OB) do { using OpT = HsvDetectBitmapOp<0, IB, OC, OB>; \
        if (d16) CU(c, launch_k(pdl, map_rgb24_kernel<OpT, G, true, true>, grid, dim3(256), 0, st, OpT{bitmap}, f.src, f.sstride, f.dst, f.dstride, ww, hh)); \
        else CU(c, launch_k(pdl, map_rgb24_kernel<OpT, G, true, false>, grid, dim3(256), 0, st, OpT{bitmap}, f.src, f.sstride, f.dst, f.dstride, ww, hh)); } while (0)
#define LD32(IB) do { if (fo.coff == 0) { if (fo.bgr) LD3(IB, 0, true); else LD3(IB, 0, false); } else { if (fo.bgr) LD3(IB, 1, true); else LD3(IB, 1, false); } } while (0)
      if (fi.bgr) LD32(true); else LD32(false);
#undef LD32
#undef LD3
      c->launches++;
      CU(c, cudaGetLastError());
      return 0;
    }
  } else {
    pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  }
  int rc;
  if (fi.bpp == 3) rc = fi.bgr ? launch_hsvdetector_t<3, 0, true>(c, fo, s, bitmap, pdl, f, st) : launch_hsvdetector_t<3, 0, false>(c, fo, s, bitmap, pdl, f, st);
  else if (fi.coff == 0) rc = fi.bgr ? launch_hsvdetector_t<4, 0, true>(c, fo, s, bitmap, pdl, f, st) : launch_hsvdetector_t<4, 0, false>(c, fo, s, bitmap, pdl, f, st);
  else rc = fi.bgr ? launch_hsvdetector_t<4, 1, true>(c, fo, s, bitmap, pdl, f, st) : launch_hsvdetector_t<4, 1, false>(c, fo, s, bitmap, pdl, f, st);
  if (rc) return rc;
  c->launches++;
  CU(c, cudaGetLastError());
  return 0;
}

// zero-copy host path of the memoised hsv elements: one persistent TMA streaming kernel reads the pinned host frame over
// PCIe and writes the result straight back (in place for hsvfilter), like the colorlut zero-copy path
template <typename Op>
int launch_map_zero_copy(b200vfx_ctx *c, Op op, const uint8_t *dsrc, long ss, uint8_t *ddst, long ds, int row_bytes, int h) {
  constexpr int TILE = 4096, STAGES = 4, THREADS = 256, B = 4;
  constexpr int smem = stream_smem_bytes<TILE, STAGES>();
  auto k = map_stream_kernel<TILE, STAGES, THREADS, B, Op>;
  static std::mutex mu;
  static bool attr_set[64] = {false};
  {
    std::lock_guard<std::mutex> g(mu);
    const int dev = (c->device >= 0 && c->device < 64) ? c->device : 0;
    if (!attr_set[dev]) { CU(c, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set[dev] = true; }
  }
  const long long ntiles = (long long)ceil_div(row_bytes, TILE) * h;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(ntiles, c->zc_grid > 0 ? c->zc_grid : 96));
  pdl_admit(false, c->s_k, Span{0, 0}, Span{0, 0});
  k<<<grid, THREADS, smem, c->s_k>>>(op, dsrc, ss, ddst, ds, row_bytes, h);
  c->launches++;
  CU(c, cudaGetLastError());
  if (c->host_async) return 0;   // asynchronous host-frame mode: the caller holds a fence
  CU(c, cudaStreamSynchronize(c->s_k));
  pdl_forget(c->s_k);
  return 0;
}

// ---- host-pointer pipeline -------------------------------------------------------------------
// Copies rows [r0,r1) of a host/device plane pair through device staging in chunks; `launch`
// enqueues the element kernel for one chunk of DEVICE rows on the given stream.
struct Staged {
  const uint8_t *src; long sstride; size_t in_row_bytes;   // src == nullptr: no input plane
  uint8_t *dst; long dstride; size_t out_row_bytes;        // in place: dst == src plane
  int height;
  bool in_place;
  int device_addressable = 0;  // bit 0: src, bit 1: dst are pinned host planes the kernel may address directly (zero-copy)
};

template <typename LaunchFn>
int run_staged(b200vfx_ctx *c, const Staged &s, LaunchFn launch) {
  if (s.height == 0) return 0;
  const bool src_dev = !s.src || (s.device_addressable & 1) || is_device_ptr(s.src);  // no input plane counts as "nothing to upload"
  const bool dst_dev = (s.device_addressable & 2) || is_device_ptr(s.dst);
  if (s.in_place ? dst_dev : (src_dev && dst_dev)) {  // everything already in HBM: just enqueue
    return launch(s.src, s.sstride, s.dst, s.dstride, 0, s.height, c->stream());
  }
  // device staging strides: tightly packed rows rounded up to 16 B so vector paths apply
  const long in_ds = (long)((s.in_row_bytes + 15) & ~(size_t)15);
  const long out_ds = s.in_place ? in_ds : (long)((s.out_row_bytes + 15) & ~(size_t)15);
  const uint8_t *d_in = nullptr;
  uint8_t *d_out = nullptr;
  long d_in_stride = 0, d_out_stride = 0;
  const bool need_h2d = s.in_place ? true : !src_dev;
  const bool need_d2h = s.in_place ? true : !dst_dev;
  const bool async = c->host_async;
  const int slot = c->async_slot;
  DevBuf &st_in = async ? c->astage_in[slot] : c->stage_in, &st_out = async ? c->astage_out[slot] : c->stage_out;
  if (async) {
    c->async_slot ^= 1;
    if (!c->slot_done[slot]) CU(c, cudaEventCreateWithFlags(&c->slot_done[slot], cudaEventDisableTiming));
    if (c->slot_used[slot]) {   // the frame that used this slot two calls ago must have left it
      CU(c, cudaStreamWaitEvent(c->s_h2d, c->slot_done[slot], 0));
      CU(c, cudaStreamWaitEvent(c->s_k, c->slot_done[slot], 0));
    }
    // growing a slot frees memory that earlier frames may still use: drain first (happens on a caps change only)
    const size_t want_in = (s.in_place || !src_dev) ? (size_t)in_ds * s.height : 0, want_out = (!s.in_place && !dst_dev) ? (size_t)out_ds * s.height : 0;
    if (want_in > st_in.cap || want_out > st_out.cap) CU(c, cudaDeviceSynchronize());
  }
  if (s.in_place) {
    CU(c, st_in.reserve((size_t)in_ds * s.height));
    d_in = d_out = st_in.p; d_in_stride = d_out_stride = in_ds;
  } else {
    if (src_dev) { d_in = s.src; d_in_stride = s.sstride; }
    else { CU(c, st_in.reserve((size_t)in_ds * s.height)); d_in = st_in.p; d_in_stride = in_ds; }
    if (dst_dev) { d_out = s.dst; d_out_stride = s.dstride; }
    else { CU(c, st_out.reserve((size_t)out_ds * s.height)); d_out = st_out.p; d_out_stride = out_ds; }
  }
  // one plane in HBM, the other on the host: the device plane is ordered on the context stream (header contract)
  if (!s.in_place && ((src_dev && s.src && !(s.device_addressable & 1)) || (dst_dev && !(s.device_addressable & 2))))
    if (int rc = order_after_ctx_stream(c, c->s_k)) return rc;
  int rows = c->chunk_rows;
  if (rows <= 0 && async) rows = s.height;   // frames overlap each other: chunking one frame only adds events (1379 vs 1084 frames/s)
  if (rows <= 0) {
    // measured on B200/PCIe5 (profiles/r01_e2e_chunks.md): ~8 MB chunks win for 4K frames (4 chunks);
    // never fewer than 4 chunks (so the two PCIe directions overlap) and never more than 16
    const size_t per_row = std::max<size_t>(std::max(s.in_row_bytes, s.out_row_bytes), 1);
    const size_t total = per_row * (size_t)s.height;
    size_t nch = std::min<size_t>(16, std::max<size_t>(4, total / ((size_t)8 << 20)));
    if (total / nch < ((size_t)256 << 10)) nch = std::max<size_t>(1, total / ((size_t)256 << 10));
    rows = (int)std::max<size_t>(1, ((size_t)s.height + nch - 1) / nch);
  }
  const int nchunks = ceil_div(s.height, rows);
  while ((int)c->ev_in.size() < nchunks) {
    cudaEvent_t a, b;
    CU(c, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    CU(c, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    c->ev_in.push_back(a); c->ev_k.push_back(b);
  }
  for (int k = 0; k < nchunks; k++) {
    const int r0 = k * rows, r1 = std::min(s.height, r0 + rows);
    if (need_h2d) {
      const uint8_t *hp = s.src + (size_t)r0 * s.sstride;
      uint8_t *dp = const_cast<uint8_t *>(d_in) + (size_t)r0 * d_in_stride;
      if ((size_t)s.sstride == s.in_row_bytes && (size_t)d_in_stride == s.in_row_bytes)
        CU(c, cudaMemcpyAsync(dp, hp, s.in_row_bytes * (size_t)(r1 - r0), cudaMemcpyHostToDevice, c->s_h2d));
      else
        CU(c, cudaMemcpy2DAsync(dp, (size_t)d_in_stride, hp, (size_t)s.sstride, s.in_row_bytes, (size_t)(r1 - r0),
                                cudaMemcpyHostToDevice, c->s_h2d));
      CU(c, cudaEventRecord(c->ev_in[k], c->s_h2d));
      CU(c, cudaStreamWaitEvent(c->s_k, c->ev_in[k], 0));
    }
    int rc = launch(d_in ? d_in + (size_t)r0 * d_in_stride : nullptr, d_in_stride, d_out + (size_t)r0 * d_out_stride,
                    d_out_stride, r0, r1 - r0, c->s_k);
    if (rc) return rc;
    if (need_d2h) {
      CU(c, cudaEventRecord(c->ev_k[k], c->s_k));
      CU(c, cudaStreamWaitEvent(c->s_d2h, c->ev_k[k], 0));
      uint8_t *hp = s.dst + (size_t)r0 * s.dstride;
      const uint8_t *dp = d_out + (size_t)r0 * d_out_stride;
      if ((size_t)s.dstride == s.out_row_bytes && (size_t)d_out_stride == s.out_row_bytes)
        CU(c, cudaMemcpyAsync(hp, dp, s.out_row_bytes * (size_t)(r1 - r0), cudaMemcpyDeviceToHost, c->s_d2h));
      else
        CU(c, cudaMemcpy2DAsync(hp, (size_t)s.dstride, dp, (size_t)d_out_stride, s.out_row_bytes, (size_t)(r1 - r0),
                                cudaMemcpyDeviceToHost, c->s_d2h));
    }
  }
  if (async) {   // the call returns here; b200vfx_fence / b200vfx_ctx_synchronize tell when the frame is on the host
    CU(c, cudaEventRecord(c->slot_done[slot], need_d2h ? c->s_d2h : c->s_k));
    c->slot_used[slot] = true;
    // a plane in HBM was read or written on the internal stream: later device-pointer calls (context stream) come after it
    if (!s.in_place && ((src_dev && s.src && !(s.device_addressable & 1)) || (dst_dev && !(s.device_addressable & 2))))
      if (int rc = order_ctx_stream_after(c, c->s_k)) return rc;
    return 0;
  }
  if (need_d2h) CU(c, cudaStreamSynchronize(c->s_d2h));
  else CU(c, cudaStreamSynchronize(c->s_k));
  pdl_forget(c->s_k);
  return 0;
}

// scratch of the single-launch reductions: 64 bytes of grid counters (zero between launches: the kernels re-arm them) followed
// by `bytes` of partials.  Growing the buffer synchronises first (a previous launch may still use the old one).
int reduce_scratch(b200vfx_ctx *c, size_t bytes, unsigned **counters, uint32_t **partials) {
  const size_t need = 64 + bytes;
  if (need > c->reduce_scratch.cap) {
    CU(c, cudaDeviceSynchronize());
    CU(c, c->reduce_scratch.reserve(std::max<size_t>(need, (size_t)1 << 20)));
    CU(c, cudaMemset(c->reduce_scratch.p, 0, 64));
  }
  *counters = (unsigned *)c->reduce_scratch.p;
  if (partials) *partials = (uint32_t *)(c->reduce_scratch.p + 64);
  return 0;
}

int check_frame(b200vfx_ctx *c, int width, int height, const void *a, long astride, size_t a_row, const void *b,
                long bstride, size_t b_row) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (width < 0 || height < 0) return fail(c, B200VFX_ERR_INVALID, "negative frame size");
  if (height > 0 && width > 0) {
    if (!a || (b_row && !b)) return fail(c, B200VFX_ERR_INVALID, "null plane pointer");
    if (astride < 0 || (size_t)astride < a_row) return fail(c, B200VFX_ERR_INVALID, "stride %ld < row bytes %zu", astride, a_row);
    if (b_row && (bstride < 0 || (size_t)bstride < b_row)) return fail(c, B200VFX_ERR_INVALID, "stride %ld < row bytes %zu", bstride, b_row);
  }
  return 0;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

// =================================================================================================
extern "C" {

int b200vfx_abi_version(void) { return B200VFX_ABI_VERSION; }

int b200vfx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char *b200vfx_last_error(const b200vfx_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int b200vfx_ctx_create(b200vfx_ctx **out, int device) {
  if (!out) return fail(nullptr, B200VFX_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(nullptr, B200VFX_ERR_CUDA, "no CUDA device available (%s); libb200vfx has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0) CU(nullptr, cudaGetDevice(&device));
  if (device >= n) return fail(nullptr, B200VFX_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  DeviceGuard g(device);
  b200vfx_ctx *c = new b200vfx_ctx();
  c->device = device;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, device) == cudaSuccess && v > 0) c->l2_persist_max = (size_t)v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxAccessPolicyWindowSize, device) == cudaSuccess && v > 0) c->l2_window_max = (size_t)v;
    if (const char *e = getenv("B200VFX_L2_PERSIST")) c->l2_persist = atoi(e);
  }
  if (c->sm_count <= 0) c->sm_count = 148;
  {
    int tps = 0;
    if (cudaDeviceGetAttribute(&tps, cudaDevAttrMaxThreadsPerMultiProcessor, device) != cudaSuccess || tps <= 0) tps = 2048;
    std::lock_guard<std::mutex> g(g_recent_mu);
    g_capacity_threads = std::max(g_capacity_threads, (long long)c->sm_count * tps);   // larger = more conservative
  }
  if (const char *e = getenv("B200VFX_STREAM_PATH")) c->stream_path = atoi(e);
  if (const char *e = getenv("B200VFX_STREAM_CFG")) c->stream_cfg = atoi(e);
  if (const char *e = getenv("B200VFX_STREAM_CTAS")) c->stream_ctas = atoi(e);
  if (const char *e = getenv("B200VFX_STREAM_HINT")) c->stream_hint = atoi(e);
  if (const char *e = getenv("B200VFX_MEMO_PX")) c->memo_px = atoi(e);
  cudaError_t err = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking);
  if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking);
  if (err != cudaSuccess) {
    int rc = fail(nullptr, B200VFX_ERR_CUDA, "stream creation failed: %s", cudaGetErrorString(err));
    b200vfx_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

void b200vfx_ctx_destroy(b200vfx_ctx *c) {
  if (!c) return;
  DeviceGuard g(c->device);
  cudaDeviceSynchronize();
  pdl_forget(c->own_stream); pdl_forget(c->s_k);
  if (c->use_user_stream) pdl_forget(c->user_stream);   // everything of ours on it has completed (device synchronised)
  b200vfx_colorlut_clear(c);
  if (c->d_hf_memo) cudaFree(c->d_hf_memo);
  if (c->d_hd_bitmap) cudaFree(c->d_hd_bitmap);
  for (TableSync *t : {&c->ts_colorlut, &c->ts_hf, &c->ts_hd}) if (t->ev) cudaEventDestroy(t->ev);
  for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_k) cudaEventDestroy(e);
  if (c->ev_order) cudaEventDestroy(c->ev_order);
  if (c->ev_order_back) cudaEventDestroy(c->ev_order_back);
  for (auto &kv : c->taps_cache) { cudaFree(kv.second.taps); cudaFree(kv.second.meta); }
  c->stage_in.release(); c->stage_out.release(); c->stage_sums.release(); c->reduce_scratch.release();
  if (c->result_pinned) { cudaFreeHost(c->result_pinned); c->result_pinned = nullptr; }
  for (int i = 0; i < 2; i++) { c->astage_in[i].release(); c->astage_out[i].release(); if (c->slot_done[i]) { cudaEventDestroy(c->slot_done[i]); c->slot_done[i] = nullptr; } }
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_k) cudaStreamDestroy(c->s_k);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
}

int b200vfx_ctx_set_stream(b200vfx_ctx *c, void *cuda_stream) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  c->user_stream = (cudaStream_t)cuda_stream;
  c->use_user_stream = true;
  return 0;
}

int b200vfx_ctx_synchronize(b200vfx_ctx *c) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  DeviceGuard g(c->device);
  CU(c, cudaStreamSynchronize(c->stream()));
  if (c->host_async) {   // frames submitted asynchronously from host memory
    CU(c, cudaStreamSynchronize(c->s_k));
    CU(c, cudaStreamSynchronize(c->s_d2h));
    pdl_forget(c->s_k);
  }
  return 0;
}

int b200vfx_ctx_set_host_async(b200vfx_ctx *c, int enable) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  DeviceGuard g(c->device);
  if (c->host_async && !enable) {   // leaving the mode: nothing may be in flight when synchronous calls resume
    CU(c, cudaStreamSynchronize(c->s_h2d));
    CU(c, cudaStreamSynchronize(c->s_k));
    CU(c, cudaStreamSynchronize(c->s_d2h));
    pdl_forget(c->s_k);
  }
  c->host_async = enable != 0;
  return 0;
}

struct b200vfx_fence { cudaEvent_t ev; int device; };

int b200vfx_fence_create(b200vfx_ctx *c, b200vfx_fence **out) {
  if (!c || !out) return fail(c, B200VFX_ERR_INVALID, "fence: null argument");
  DeviceGuard g(c->device);
  cudaEvent_t ev = nullptr, k = nullptr;
  CU(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  // everything the host-frame pipeline has been given so far: the download stream follows the kernel stream, which follows
  // the upload stream, chunk by chunk; a call without a download ends on the kernel stream
  if (cudaEventCreateWithFlags(&k, cudaEventDisableTiming) != cudaSuccess) { cudaEventDestroy(ev); return fail(c, B200VFX_ERR_CUDA, "fence: event"); }
  cudaError_t e = cudaEventRecord(k, c->s_k);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_d2h, k, 0);
  if (e == cudaSuccess) e = cudaEventRecord(ev, c->s_d2h);
  cudaEventDestroy(k);   // destruction is deferred until the event has completed
  if (e != cudaSuccess) { cudaEventDestroy(ev); return fail(c, B200VFX_ERR_CUDA, "fence: %s", cudaGetErrorString(e)); }
  *out = new b200vfx_fence{ev, c->device};
  return 0;
}
int b200vfx_fence_wait(b200vfx_fence *f) {
  if (!f) return B200VFX_ERR_INVALID;
  return cudaEventSynchronize(f->ev) == cudaSuccess ? 0 : B200VFX_ERR_CUDA;
}
int b200vfx_fence_query(b200vfx_fence *f) {
  if (!f) return B200VFX_ERR_INVALID;
  const cudaError_t e = cudaEventQuery(f->ev);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  return B200VFX_ERR_CUDA;
}
void b200vfx_fence_destroy(b200vfx_fence *f) {
  if (!f) return;
  cudaEventDestroy(f->ev);
  delete f;
}

int b200vfx_ctx_set_chunk_rows(b200vfx_ctx *c, int rows) {
  if (!c || rows < 0) return fail(c, B200VFX_ERR_INVALID, "bad chunk rows");
  c->chunk_rows = rows;
  return 0;
}

uint64_t b200vfx_ctx_kernel_launches(const b200vfx_ctx *c) { return c ? c->launches : 0; }

int b200vfx_ctx_set_option(b200vfx_ctx *c, const char *name, int value) {
  if (!c || !name) return fail(c, B200VFX_ERR_INVALID, "null argument");
  const std::string n(name);
  if (n == "stream_path") c->stream_path = value;
  else if (n == "stream_cfg") c->stream_cfg = value;
  else if (n == "stream_ctas") c->stream_ctas = value;
  else if (n == "stream_hint") c->stream_hint = value;
  else if (n == "memo_px") c->memo_px = value;
  else if (n == "pdl") c->pdl = value != 0;
  else if (n == "cd_split") c->cd_split = std::max(0, std::min(value, 8));
  else if (n == "blockhash_rows") c->blockhash_rows = std::max(0, std::min(value, 8));
  else if (n == "tile_gather_path") c->tg_path = value;
  else if (n == "tile_gather_cfg") c->tg_cfg = value;
  else if (n == "tile_gather_ctas") c->tg_ctas = value;
  else if (n == "peer_timeout_ms") c->peer_timeout_ms = value > 0 ? value : 1;
  else if (n == "zero_copy") { c->zero_copy = value; c->zc_calls = 0; c->zc_best_ms[0] = c->zc_best_ms[1] = 1e30; }
  else if (n == "blockhash_tma") c->blockhash_tma = value != 0;
  else if (n == "memo_tile") c->memo_tile = value != 0;
  else if (n == "rgba64_x4") c->rgba64_x4 = value != 0;
  else if (n == "memo_ctas") c->memo_ctas = std::max(2, std::min(8, value));
  else if (n == "cd_cluster") c->cd_cluster = (value == 1 || value == 2 || value == 4 || value == 8) ? value : 2;
  else if (n == "l2_persist") c->l2_persist = value;
  else if (n == "zc_cfg") c->zc_cfg = value;
  else if (n == "zc_hybrid") c->zc_hybrid = value;
  else if (n == "zc_ctas") c->zc_ctas = value;
  else if (n == "zc_grid") c->zc_grid = value;
  else if (n == "stream_grid") c->stream_grid = value;
  else if (n == "hsv_memo") { c->hsv_memo = value; c->hf_ready = c->hd_ready = false; c->hf_px_seen = c->hd_px_seen = 0; }
  else return fail(c, B200VFX_ERR_INVALID, "unknown option '%s'", name);
  return 0;
}

void *b200vfx_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void b200vfx_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- test hooks ------------------------------------------------------------------------------------
int b200vfx_debug_pdl_admit(void *stream_key, uintptr_t src_lo, uintptr_t src_hi, uintptr_t dst_lo, uintptr_t dst_hi,
                            int want_pdl, long long threads, int lingers) {
  cudaStream_t st = (cudaStream_t)stream_key;
  const bool ok = pdl_admit(want_pdl != 0, st, Span{src_lo, src_hi}, Span{dst_lo, dst_hi}, (lingers & 2) != 0);
  pdl_note_threads(st, threads);
  if (lingers) pdl_note_linger(st);
  return ok ? 1 : 0;
}
void b200vfx_debug_pdl_reset(void *stream_key) { pdl_forget((cudaStream_t)stream_key); }

// ---- device-resident frames ---------------------------------------------------------------------
void *b200vfx_device_alloc(b200vfx_ctx *c, size_t bytes) {
  if (!c) { fail(nullptr, B200VFX_ERR_INVALID, "null context"); return nullptr; }
  DeviceGuard g(c->device);
  void *p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { fail(c, B200VFX_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError())); return nullptr; }
  return p;
}
void b200vfx_device_free(b200vfx_ctx *c, void *p) {
  if (!c || !p) return;
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream());   // nothing of ours may still use it
  pdl_forget(c->stream());
  cudaFree(p);
}
static int copy_rows(b200vfx_ctx *c, void *dst, int dstride, const void *src, int sstride, size_t row_bytes, int rows, cudaMemcpyKind kind) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (rows < 0 || dstride < 0 || sstride < 0 || (rows > 0 && row_bytes > 0 && (!dst || !src || (size_t)dstride < row_bytes || (size_t)sstride < row_bytes)))
    return fail(c, B200VFX_ERR_INVALID, "bad plane description");
  if (rows == 0 || row_bytes == 0) return 0;
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream();
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});   // a copy is an ordinary stream operation: everything before it completes first
  if ((size_t)dstride == row_bytes && (size_t)sstride == row_bytes) CU(c, cudaMemcpyAsync(dst, src, row_bytes * (size_t)rows, kind, st));
  else CU(c, cudaMemcpy2DAsync(dst, (size_t)dstride, src, (size_t)sstride, row_bytes, (size_t)rows, kind, st));
  return 0;
}
int b200vfx_upload(b200vfx_ctx *c, void *dev_dst, int dst_stride, const void *host_src, int src_stride, size_t row_bytes, int rows) {
  return copy_rows(c, dev_dst, dst_stride, host_src, src_stride, row_bytes, rows, cudaMemcpyHostToDevice);
}
int b200vfx_download(b200vfx_ctx *c, void *host_dst, int dst_stride, const void *dev_src, int src_stride, size_t row_bytes, int rows) {
  return copy_rows(c, host_dst, dst_stride, dev_src, src_stride, row_bytes, rows, cudaMemcpyDeviceToHost);
}

int b200vfx_copy_plane(b200vfx_ctx *c, void *dst, int dst_stride, const void *src, int src_stride, size_t row_bytes, int rows) {
  return copy_rows(c, dst, dst_stride, src, src_stride, row_bytes, rows, cudaMemcpyDefault);
}
int b200vfx_pointer_is_device(const void *p) { return p && is_device_ptr(p) ? 1 : 0; }

// ---- colorlut --------------------------------------------------------------------------------
int b200vfx_colorlut_clear(b200vfx_ctx *c) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  DeviceGuard g(c->device);
  if (c->d_pair) cudaFree(c->d_pair);
  if (c->d_lut1d) cudaFree(c->d_lut1d);
  if (c->d_axis8) cudaFree(c->d_axis8);
  if (c->d_memo) cudaFree(c->d_memo);
  if (c->d_memo1d) cudaFree(c->d_memo1d);
  c->d_pair = nullptr; c->d_lut1d = nullptr; c->d_axis8 = nullptr; c->d_memo = nullptr; c->d_memo1d = nullptr;
  c->have_lut = false; c->memo_ready = false; c->lut_kind = 0; c->lut_size = 0;
  return 0;
}

int b200vfx_colorlut_set_mode(b200vfx_ctx *c, int mode) {
  if (!c || (mode != 0 && mode != 1)) return fail(c, B200VFX_ERR_INVALID, "colorlut mode must be 0 (auto) or 1 (direct)");
  c->mode = mode;
  return 0;
}

int b200vfx_colorlut_set_lut(b200vfx_ctx *c, int kind, int size, const float *values, const float scale[3],
                             const float offset[3]) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!values || !scale || !offset) return fail(c, B200VFX_ERR_INVALID, "null LUT argument");
  if (kind == 1) { if (size < 2 || size > 65536) return fail(c, B200VFX_ERR_INVALID, "Invalid LUT size %d, expected 2..=65536", size); }
  else if (kind == 3) { if (size < 2 || size > 256) return fail(c, B200VFX_ERR_INVALID, "Invalid LUT size %d, expected 2..=256", size); }
  else return fail(c, B200VFX_ERR_INVALID, "LUT kind must be 1 or 3");
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream();
  CU(c, cudaStreamSynchronize(st));
  b200vfx_colorlut_clear(c);
  const size_t n = kind == 1 ? (size_t)size : (size_t)size * size * size;
  if (kind == 3) {
    // x-pair layout: {lut[x], lut[min(x+1,max)] - lut[x]} per entry; the difference is the same single f32
    // rounding lerp4's `b - a` performs per pixel (imp.rs:528-535).  The alpha lane (1.0, parser.rs:253-256) is
    // never consumed (imp.rs:444-448) and is not stored.
    std::vector<LutPair> h(n);
    const size_t N = (size_t)size;
    for (size_t i = 0; i < n; i++) {
      const size_t x = i % N, i1 = (x + 1 < N) ? i + 1 : i;
      float av[3], dv[3];
      for (int k = 0; k < 3; k++) {
        const volatile float a = values[3 * i + k], b = values[3 * i1 + k];
        const volatile float d = b - a;  // volatile: one IEEE binary32 subtraction, no extended precision / fusion
        av[k] = a; dv[k] = d;
      }
      // register-pair friendly order for the packed f32x2 evaluator: {a.r,a.g | d.r,d.g | a.b,d.b | pad}
      h[i].a_rg[0] = av[0]; h[i].a_rg[1] = av[1]; h[i].d_rg[0] = dv[0]; h[i].d_rg[1] = dv[1]; h[i].a_b = av[2]; h[i].d_b = dv[2];
      h[i].pad[0] = h[i].pad[1] = 0.0f;
    }
    CU(c, cudaMalloc(&c->d_pair, n * sizeof(LutPair)));
    CU(c, cudaMemcpyAsync(c->d_pair, h.data(), n * sizeof(LutPair), cudaMemcpyHostToDevice, st));
    CU(c, cudaStreamSynchronize(st));
  } else {
    std::vector<float> h(3 * n);
    for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) h[(size_t)k * n + i] = values[3 * i + k];  // parser.rs:232-236
    CU(c, cudaMalloc(&c->d_lut1d, 3 * n * sizeof(float)));
    CU(c, cudaMemcpyAsync(c->d_lut1d, h.data(), 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(c, cudaStreamSynchronize(st));
  }
  for (int i = 0; i < 3; i++) { c->scale[i] = scale[i]; c->offset[i] = offset[i]; }
  c->lut_kind = kind; c->lut_size = size;
  if (int rc = build_axis(c, st)) return rc;
  CU(c, cudaStreamSynchronize(st));
  c->have_lut = true;
  return 0;
}

int b200vfx_colorlut_load_file(b200vfx_ctx *c, const char *location) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!location) return fail(c, B200VFX_ERR_INVALID, "LUT file location is not configured");  // imp.rs:175-180
  int kind = 0, size = 0;
  float *values = nullptr, scale[3], offset[3];
  char err[400];
  int rc = b200vfx_cube_parse_file(location, &kind, &size, &values, scale, offset, err, sizeof err);
  if (rc) return fail(c, rc, "Failed to parse LUT file %s: %s", location, err);  // imp.rs:182-187
  rc = b200vfx_colorlut_set_lut(c, kind, size, values, scale, offset);
  b200vfx_cube_free(values);
  return rc;
}

int b200vfx_colorlut_process(b200vfx_ctx *c, int fmt, int width, int height, const void *src, int src_stride,
                             void *dst, int dst_stride) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!c->have_lut) return fail(c, B200VFX_ERR_NOT_NEGOTIATED, "No LUT configured");  // imp.rs:210-213
  int bpp;
  if (fmt == B200VFX_FORMAT_RGBA) bpp = 4;
  else if (fmt == B200VFX_FORMAT_RGBA64_LE || fmt == B200VFX_FORMAT_RGBA64_BE) bpp = 8;
  else return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut: format %d is not RGBA / RGBA64_LE / RGBA64_BE", fmt);
  const size_t row = (size_t)width * bpp;
  if (int rc = check_frame(c, width, height, src, src_stride, row, dst, dst_stride, row)) return rc;
  if (width == 0 || height == 0) return 0;
  DeviceGuard g(c->device);
  // Zero-copy variant for PINNED host frames: the TMA streaming kernel bulk-loads tiles straight from host memory
  // over PCIe and bulk-stores the results straight back -- no staging buffers, both PCIe directions busy for the
  // whole frame, no chunk pipeline to fill and drain (RGBA memo path).  Option "zero_copy": 0 never, 1 always,
  // 2 auto (default): the first eligible calls alternate between this kernel and the staged copy-engine pipeline,
  // timed with the host clock (both are synchronous), and the faster one is kept -- how well kernel-issued bulk
  // reads of host memory perform depends on the host's PCIe/IOMMU set-up (profiles/r01_e2e.md).
  void *dsrc = nullptr, *ddst = nullptr;
  const bool eligible = c->zero_copy != 0 && fmt == B200VFX_FORMAT_RGBA && c->mode == 0 && (width % 4) == 0 &&
                        aligned(src, src_stride, 16) && aligned(dst, dst_stride, 16) && pinned_device_ptr(src, &dsrc) &&
                        pinned_device_ptr(dst, &ddst);
  // asynchronous host-frame mode: consecutive frames overlap on the copy engines (upload of frame i+1 beside the download of
  // frame i), which beats kernel-issued PCIe traffic (1379 vs 1180 frames/s, profiles/r02_e2e_async.jsonl): "auto" means
  // the staged pipeline there, and nothing is probed
  const bool probing = eligible && c->zero_copy == 2 && c->zc_calls < 6 && !c->host_async;
  const bool use_zc = eligible && c->zc_hybrid == 0 &&
                      (c->zero_copy == 1 || (!c->host_async && (probing ? (c->zc_calls % 2 == 0) : (c->zc_best_ms[0] <= c->zc_best_ms[1]))));
  const auto t_begin = std::chrono::steady_clock::now();
  auto probe_done = [&](int which) {
    if (!eligible || c->zero_copy != 2) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    if (probing) {
      if (c->zc_calls >= 2) c->zc_best_ms[which] = std::min(c->zc_best_ms[which], ms);   // first call of each kind is a warm-up
      c->zc_calls++;
      return;
    }
    // watchdog: the kernel-issued PCIe reads occasionally fall into a slow mode (measured 4x, profiles/r01_e2e.md);
    // three consecutive calls 1.5x slower than the other path's best trigger a new probe
    if (ms > 1.5 * c->zc_best_ms[1 - which]) { if (++c->zc_bad_streak >= 3) { c->zc_calls = 0; c->zc_bad_streak = 0; c->zc_best_ms[0] = c->zc_best_ms[1] = 1e30; } }
    else c->zc_bad_streak = 0;
  };
  if (use_zc) {
    // PCIe needs far fewer bytes in flight than HBM: a small grid of small tiles measured best (profiles/)
    const int saved_path = c->stream_path, saved_cfg = c->stream_cfg, saved_ctas = c->stream_ctas, saved_grid = c->stream_grid;
    c->stream_path = 1; c->stream_cfg = c->zc_cfg; c->stream_ctas = c->zc_ctas; c->stream_grid = c->zc_grid;
    int rc = launch_colorlut(c, fmt, Frame{(const uint8_t *)dsrc, src_stride, (uint8_t *)ddst, dst_stride, width, height}, c->s_k);
    c->stream_path = saved_path; c->stream_cfg = saved_cfg; c->stream_ctas = saved_ctas; c->stream_grid = saved_grid;
    if (rc) return rc;
    if (c->host_async) return 0;   // asynchronous host-frame mode: the caller holds a fence
    CU(c, cudaStreamSynchronize(c->s_k));
    pdl_forget(c->s_k);
    probe_done(0);
    return 0;
  }
  Staged s{(const uint8_t *)src, src_stride, row, (uint8_t *)dst, dst_stride, row, height, false};
  if (eligible && c->zc_hybrid == 1) { s.dst = (uint8_t *)ddst; s.device_addressable = 2; }        // copy engine in, kernel stores out
  else if (eligible && c->zc_hybrid == 2) { s.src = (const uint8_t *)dsrc; s.device_addressable = 1; }  // kernel loads in, copy engine out
  const int rc = run_staged(c, s, [&](const uint8_t *ds, long dss, uint8_t *dd, long dds, int, int rows, cudaStream_t st) {
    return launch_colorlut(c, fmt, Frame{ds, dss, dd, dds, width, rows}, st);
  });
  if (rc == 0 && !c->host_async) probe_done(1);
  return rc;
}

// ---- hsvfilter / hsvdetector ------------------------------------------------------------------
int b200vfx_hsvfilter_process(b200vfx_ctx *c, int fmt, int width, int height, void *data, int stride,
                              float hue_shift, float saturation_mul, float saturation_off, float value_mul,
                              float value_off) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  FmtInfo fi;
  if (!fmt8_info(fmt, &fi)) return fail(c, B200VFX_ERR_UNSUPPORTED, "hsvfilter: unsupported format %d", fmt);
  const size_t row = (size_t)width * fi.bpp;
  if (int rc = check_frame(c, width, height, data, stride, row, nullptr, 0, 0)) return rc;
  if (width == 0 || height == 0) return 0;
  DeviceGuard g(c->device);
  const HsvFilterSettings hs{hue_shift, saturation_mul, saturation_off, value_mul, value_off};
  {  // pinned host frame + answer table already built for exactly these settings: zero-copy, in place over PCIe
    void *dptr = nullptr;
    if (c->zero_copy != 0 && fi.bpp == 4 && (width % 4) == 0 && aligned(data, stride, 16) && c->hf_ready && c->hf_key_valid &&
        c->hsv_memo != 0 && std::memcmp(&c->hf_key, &hs, sizeof hs) == 0 && pinned_device_ptr(data, &dptr)) {
      uint8_t *d = (uint8_t *)dptr;
      if (int rc = table_wait(c, c->ts_hf, c->s_k)) return rc;
#define ZF(CO, BG) return launch_map_zero_copy(c, HsvFilterMemoOp<CO, BG>{c->d_hf_memo}, d, stride, d, stride, 4 * width, height)
      if (fi.coff == 0) { if (fi.bgr) ZF(0, true); else ZF(0, false); }
      else { if (fi.bgr) ZF(1, true); else ZF(1, false); }
#undef ZF
    }
  }
  Staged s{(const uint8_t *)data, stride, row, (uint8_t *)data, stride, row, height, true};
  return run_staged(c, s, [&](const uint8_t *, long, uint8_t *dd, long dds, int, int rows, cudaStream_t st) {
    return launch_hsvfilter(c, fi, hs, dd, dds, width, rows, st);
  });
}

int b200vfx_hsvdetector_process(b200vfx_ctx *c, int in_fmt, int out_fmt, int width, int height, const void *src,
                                int src_stride, void *dst, int dst_stride, float hue_ref, float hue_var,
                                float saturation_ref, float saturation_var, float value_ref, float value_var) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  FmtInfo fi, fo;
  if (!fmt8_info(in_fmt, &fi) || fi.alpha)  // sink caps hsvdetector/imp.rs:78-87
    return fail(c, B200VFX_ERR_UNSUPPORTED, "hsvdetector: input format %d not in RGBx,xRGB,BGRx,xBGR,RGB,BGR", in_fmt);
  if (!fmt8_info(out_fmt, &fo) || !fo.alpha)  // src caps :89-96
    return fail(c, B200VFX_ERR_UNSUPPORTED, "hsvdetector: output format %d not in RGBA,ARGB,BGRA,ABGR", out_fmt);
  const size_t irow = (size_t)width * fi.bpp, orow = (size_t)width * 4;
  if (int rc = check_frame(c, width, height, src, src_stride, irow, dst, dst_stride, orow)) return rc;
  if (width == 0 || height == 0) return 0;
  DeviceGuard g(c->device);
  const HsvDetectSettings hs{hue_ref, hue_var, saturation_ref, saturation_var, value_ref, value_var};
  {  // pinned host frames + hit bitmap already built for exactly these settings: zero-copy over PCIe
    void *dsrc = nullptr, *ddst = nullptr;
    if (c->zero_copy != 0 && fi.bpp == 4 && (width % 4) == 0 && aligned(src, src_stride, 16) && aligned(dst, dst_stride, 16) &&
        c->hd_ready && c->hd_key_valid && c->hsv_memo != 0 && std::memcmp(&c->hd_key, &hs, sizeof hs) == 0 &&
        pinned_device_ptr(src, &dsrc) && pinned_device_ptr(dst, &ddst)) {
      const uint8_t *ps = (const uint8_t *)dsrc;
      uint8_t *pd = (uint8_t *)ddst;
      if (int rc = table_wait(c, c->ts_hd, c->s_k)) return rc;
#define ZD(IC, IB, OC, OB) return launch_map_zero_copy(c, HsvDetectBitmapOp<IC, IB, OC, OB>{c->d_hd_bitmap}, ps, src_stride, pd, dst_stride, 4 * width, height)
#define ZD2(IC, IB) do { if (fo.coff == 0) { if (fo.bgr) ZD(IC, IB, 0, true); else ZD(IC, IB, 0, false); } else { if (fo.bgr) ZD(IC, IB, 1, true); else ZD(IC, IB, 1, false); } } while (0)
      if (fi.coff == 0) { if (fi.bgr) ZD2(0, true); else ZD2(0, false); }
      else { if (fi.bgr) ZD2(1, true); else ZD2(1, false); }
#undef ZD2
#undef ZD
    }
  }
  Staged s{(const uint8_t *)src, src_stride, irow, (uint8_t *)dst, dst_stride, orow, height, false};
  return run_staged(c, s, [&](const uint8_t *ds, long dss, uint8_t *dd, long dds, int, int rows, cudaStream_t st) {
    return launch_hsvdetector(c, fi, fo, hs, Frame{ds, dss, dd, dds, width, rows}, st);
  });
}

// ---- roundedcorners ---------------------------------------------------------------------------
int b200vfx_roundmask_generate(b200vfx_ctx *c, int width, int height, int stride, unsigned border_radius_px,
                               void *a8_out) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (width <= 0 || height <= 0 || stride < width || !a8_out) return fail(c, B200VFX_ERR_INVALID, "roundmask: bad geometry");
  DeviceGuard g(c->device);
  const int total_rows = (height + 1) & ~1;  // border/imp.rs:469-470
  int r = -1;
  if (border_radius_px != 0) {
    const long lim = std::min(width, height) / 2;
    r = (int)std::min<long>((long)std::min<unsigned>(border_radius_px, 0x7FFFFFFFu), lim);
  }
  Staged s{nullptr, 0, 0, (uint8_t *)a8_out, stride, (size_t)stride, total_rows, false};
  return run_staged(c, s, [&](const uint8_t *, long, uint8_t *dd, long dds, int r0, int rows, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(stride, 256), (unsigned)rows);
    roundmask_kernel<<<grid, 256, 0, st>>>(dd, dds, width, height, stride, r0, rows, r);
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  });
}

// ---- videocompare / blockhash -----------------------------------------------------------------
// device -> host of a small result + stream synchronisation.  A copy into pageable memory makes the driver stage and
// synchronise internally (~25 us per call); through the context's pinned bounce buffer it is one DMA and one spin.
constexpr size_t kResultPinnedBytes = 256 << 10;
static int read_back_small(b200vfx_ctx *c, void *host_dst, const void *dev_src, size_t n, cudaStream_t st) {
  if (n > kResultPinnedBytes) {
    CU(c, cudaMemcpyAsync(host_dst, dev_src, n, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    return 0;
  }
  if (!c->result_pinned) CU(c, cudaHostAlloc((void **)&c->result_pinned, kResultPinnedBytes, cudaHostAllocDefault));
  CU(c, cudaMemcpyAsync(c->result_pinned, dev_src, n, cudaMemcpyDeviceToHost, st));
  CU(c, cudaStreamSynchronize(st));
  std::memcpy(host_dst, c->result_pinned, n);
  return 0;
}

int b200vfx_blockhash_sums_batch(b200vfx_ctx *c, int fmt, int width, int height, int n_frames, const void *const *srcs,
                                 const int *strides, int hw, int hh, uint32_t *sums) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (fmt != B200VFX_FORMAT_RGB && fmt != B200VFX_FORMAT_RGBA)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "videocompare: format %d is not RGB / RGBA", fmt);
  if (hw <= 0 || hh <= 0 || hw > 4096 || hh > 4096 || !sums) return fail(c, B200VFX_ERR_INVALID, "blockhash: bad hash size");
  if (n_frames < 1 || n_frames > kBlockhashMaxFrames || !srcs || !strides)
    return fail(c, B200VFX_ERR_INVALID, "blockhash: 1..%d frames per call", kBlockhashMaxFrames);
  const int bpp = fmt == B200VFX_FORMAT_RGB ? 3 : 4;
  const size_t row = (size_t)width * bpp;
  for (int f = 0; f < n_frames; f++)
    if (int rc = check_frame(c, width, height, srcs[f], strides[f], row, nullptr, 0, 0)) return rc;
  if (width <= 0 || height <= 0 || width % hw || height % hh)
    return fail(c, B200VFX_ERR_UNSUPPORTED,
                "blockhash: %dx%d is not a multiple of the %dx%d hash grid (image_hasher's fractional-weight path is not implemented)",
                width, height, hw, hh);
  DeviceGuard g(c->device);
  const int bw = width / hw, bh = height / hh;
  bool all_dev = true, on_dev[kBlockhashMaxFrames];
  for (int f = 0; f < n_frames; f++) { on_dev[f] = is_device_ptr(srcs[f]); all_dev = all_dev && on_dev[f]; }
  const bool sums_dev = is_device_ptr(sums);
  cudaStream_t st = all_dev ? c->stream() : c->s_k;
  {
    bool any_dev = sums_dev;
    for (int f = 0; f < n_frames; f++) any_dev = any_dev || on_dev[f];
    if (!all_dev && any_dev) if (int rc = order_after_ctx_stream(c, st)) return rc;   // mixed batch: device frames follow the context stream
  }
  BlockhashFrames fr{};
  const long staged_stride = (long)((row + 15) & ~(size_t)15);
  size_t staged = 0;
  for (int f = 0; f < n_frames; f++) if (!on_dev[f]) staged++;
  if (staged) CU(c, c->stage_in.reserve((size_t)staged_stride * height * staged));
  size_t slot = 0;
  for (int f = 0; f < n_frames; f++) {
    if (on_dev[f]) { fr.src[f] = (const uint8_t *)srcs[f]; fr.stride[f] = strides[f]; continue; }
    uint8_t *dp = c->stage_in.p + (size_t)staged_stride * height * slot++;   // single pass, copy-bound: no chunking
    if ((size_t)strides[f] == row && (size_t)staged_stride == row)
      CU(c, cudaMemcpyAsync(dp, srcs[f], row * (size_t)height, cudaMemcpyHostToDevice, st));
    else
      CU(c, cudaMemcpy2DAsync(dp, (size_t)staged_stride, srcs[f], (size_t)strides[f], row, (size_t)height, cudaMemcpyHostToDevice, st));
    fr.src[f] = dp; fr.stride[f] = staged_stride;
  }
  for (int f = n_frames; f < kBlockhashMaxFrames; f++) { fr.src[f] = fr.src[0]; fr.stride[f] = fr.stride[0]; }
  uint32_t *d_sums = sums;
  const int nbins = hw * hh * n_frames;
  const size_t nb = sizeof(uint32_t) * (size_t)nbins;
  if (!sums_dev) { CU(c, c->stage_sums.reserve(nb)); d_sums = (uint32_t *)c->stage_sums.p; }
  // rows per CTA: aim for >= ~8 CTAs per SM over the whole batch
  int rows_per_cta = bh;
  const long ctas_target = (long)c->sm_count * 8;
  if ((long)hw * hh * n_frames < ctas_target) rows_per_cta = std::max(1, (int)((long)bh * hw * hh * n_frames / ctas_target));
  const int zchunks = ceil_div(bh, rows_per_cta);
  dim3 grid((unsigned)hw, (unsigned)hh, (unsigned)(zchunks * n_frames));
  bool vec = bpp == 4 && (bw % 4) == 0;
  for (int f = 0; f < n_frames; f++) vec = vec && aligned(fr.src[f], fr.stride[f], 16);
  if (vec && c->blockhash_tma && n_frames == 1 && bw * 4 <= kBlockhashTileBytes) {
    // TMA-fed variant (option "blockhash_tma"): whole tile in flight through cp.async.bulk, tile <= 32 KB
    pdl_admit(false, st, Span{0, 0}, Span{0, 0});
    CU(c, cudaMemsetAsync(d_sums, 0, nb, st));
    const int rows_tma = std::max(1, std::min(rows_per_cta, kBlockhashTileBytes / (bw * 4)));
    dim3 g2((unsigned)hw, (unsigned)hh, (unsigned)ceil_div(bh, rows_tma));
    blockhash_sums_tma_kernel<<<g2, 128, (size_t)rows_tma * bw * 4, st>>>(fr.src[0], fr.stride[0], bw, bh, hw, rows_tma, d_sums);
  } else {
    uint32_t *partials = nullptr; unsigned *ticket = nullptr;
    const bool rows_kernel = vec && hw <= kBlockhashRowsMaxHW && c->blockhash_rows;
    int zc = zchunks, rpc = rows_per_cta;
    if (rows_kernel) {   // whole rows per CTA, two CTAs per SM in ONE wave
      const int want = std::max(1, std::min(bh, (2 * c->sm_count / c->blockhash_rows) / std::max(1, hh * n_frames)));
      rpc = ceil_div(bh, want);
      zc = ceil_div(bh, rpc);
    }
    if (int rc = reduce_scratch(c, (size_t)nbins * zc * 4, &ticket, &partials)) return rc;
    if (st != c->stream()) if (int rc = order_after_ctx_stream(c, st)) return rc;   // one launch of this context at a time uses the scratch
    // consecutive frames overlap (programmatic dependent launch): the kernel reads its frames while the previous launch
    // drains and touches scratch / sums only after griddepcontrol.wait.  Frames staged from the host follow a copy: plain.
    Span src_hull{0, 0};
    for (int f = 0; f < n_frames; f++) {
      const Span sp = span_of(fr.src[f], fr.stride[f], row, height);
      src_hull = f == 0 ? sp : Span{std::min(src_hull.lo, sp.lo), std::max(src_hull.hi, sp.hi)};
    }
    const bool pdl = pdl_admit(c->pdl && all_dev && st == c->stream(), st, src_hull, span_of(d_sums, (long)nb, nb, 1), true);
    if (rows_kernel) {
      const int n4row = (bw >> 2) * hw;
      dim3 gr((unsigned)zc, (unsigned)hh, (unsigned)n_frames);
      if (n4row <= 256) CU(c, launch_k(pdl, blockhash_rows_kernel<1>, gr, dim3(256), 0, st, fr, zc, bw, bh, hw, hh, rpc, d_sums, partials, ticket));
      else if (n4row <= 512) CU(c, launch_k(pdl, blockhash_rows_kernel<2>, gr, dim3(256), 0, st, fr, zc, bw, bh, hw, hh, rpc, d_sums, partials, ticket));
      else CU(c, launch_k(pdl, blockhash_rows_kernel<4>, gr, dim3(256), 0, st, fr, zc, bw, bh, hw, hh, rpc, d_sums, partials, ticket));
    } else if (vec) CU(c, launch_k(pdl, blockhash_sums_kernel<4, true>, grid, dim3(128), 0, st, fr, zchunks, bw, bh, hw, hh, rows_per_cta, d_sums, partials, ticket));
    else if (bpp == 4) CU(c, launch_k(pdl, blockhash_sums_kernel<4, false>, grid, dim3(128), 0, st, fr, zchunks, bw, bh, hw, hh, rows_per_cta, d_sums, partials, ticket));
    else CU(c, launch_k(pdl, blockhash_sums_kernel<3, false>, grid, dim3(128), 0, st, fr, zchunks, bw, bh, hw, hh, rows_per_cta, d_sums, partials, ticket));
    pdl_note_linger(st);   // every CTA passes griddepcontrol.wait before it retires
  }
  c->launches++;
  CU(c, cudaGetLastError());
  if (!sums_dev) {
    if (int rc = read_back_small(c, sums, d_sums, nb, st)) return rc;
  } else if (!all_dev) {
    CU(c, cudaStreamSynchronize(st));
  }
  return 0;
}

int b200vfx_blockhash_sums(b200vfx_ctx *c, int fmt, int width, int height, const void *src, int stride, int hw,
                           int hh, uint32_t *sums) {
  return b200vfx_blockhash_sums_batch(c, fmt, width, height, 1, &src, &stride, hw, hh, sums);
}

void b200vfx_blockhash_bits(const uint32_t *sums, int hw, int hh, int width, int height, uint8_t *bits_out) {
  // image_hasher 3.1.1 alg/blockhash.rs gen_hash! (recalled; parity unpinned -- SURVEY A.7): groups of `hash width * 4`
  // blocks (4 rows of the hash grid: two groups for the 8x8 hash), group median = element len/2 of the sorted group,
  // bit = v > m || (v == m && m > 255 * 3 * block_area / 2).
  const int n = hw * hh, group = hw * 4;
  if (group <= 0) return;
  const uint32_t half = (uint32_t)(((uint64_t)765 * (uint64_t)(width / hw) * (uint64_t)(height / hh)) / 2);
  std::vector<uint32_t> tmp;
  for (int g0 = 0; g0 < n; g0 += group) {
    const int len = std::min(group, n - g0);
    tmp.assign(sums + g0, sums + g0 + len);
    std::nth_element(tmp.begin(), tmp.begin() + len / 2, tmp.end());
    const uint32_t m = tmp[(size_t)len / 2];
    for (int i = 0; i < len; i++) {
      const uint32_t v = sums[g0 + i];
      bits_out[g0 + i] = (uint8_t)(v > m || (v == m && m > half));
    }
  }
}

int b200vfx_hash_distance(const uint8_t *a, const uint8_t *b, int nbits) {
  int d = 0;
  for (int i = 0; i < nbits; i++) d += (a[i] != 0) != (b[i] != 0);
  return d;
}

// ---- format conversion: what `videoconvert` does either side of these elements, on the device ------
namespace {
bool packed_fmt(int fmt, PackedFmt *o) {
  switch (fmt) {
    case B200VFX_FORMAT_RGBX: *o = {4, 0, 1, 2, -1}; return true;
    case B200VFX_FORMAT_RGBA: *o = {4, 0, 1, 2, 3}; return true;
    case B200VFX_FORMAT_XRGB: *o = {4, 1, 2, 3, -1}; return true;
    case B200VFX_FORMAT_ARGB: *o = {4, 1, 2, 3, 0}; return true;
    case B200VFX_FORMAT_BGRX: *o = {4, 2, 1, 0, -1}; return true;
    case B200VFX_FORMAT_BGRA: *o = {4, 2, 1, 0, 3}; return true;
    case B200VFX_FORMAT_XBGR: *o = {4, 3, 2, 1, -1}; return true;
    case B200VFX_FORMAT_ABGR: *o = {4, 3, 2, 1, 0}; return true;
    case B200VFX_FORMAT_RGB: *o = {3, 0, 1, 2, -1}; return true;
    case B200VFX_FORMAT_BGR: *o = {3, 2, 1, 0, -1}; return true;
    default: return false;
  }
}
// PRMT selector turning a 4-byte `sf` pixel into a 4-byte `df` pixel; source bytes 4-7 read 0xFF
uint32_t swizzle_selector(const PackedFmt &sf, const PackedFmt &df) {
  uint32_t nib[4] = {4, 4, 4, 4};
  nib[df.r] = (uint32_t)sf.r; nib[df.g] = (uint32_t)sf.g; nib[df.b] = (uint32_t)sf.b;
  const int fourth = 6 - df.r - df.g - df.b;
  nib[fourth] = (df.a >= 0 && sf.a >= 0) ? (uint32_t)sf.a : 4u;   // alpha only when both sides have one; else 255
  return nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
}
YuvMatrix yuv_matrix(int matrix, int height) {
  // limited-range BT.601 (SD) / BT.709 (HD), 8-bit fixed point x 256; matrix: 0 = by height like GStreamer's default
  // colorimetry (<= 576 lines: bt601), 601, 709
  const bool hd = matrix == 709 || (matrix == 0 && height > 576);
  const double kr = hd ? 0.2126 : 0.299, kb = hd ? 0.0722 : 0.114, kg = 1.0 - kr - kb;
  const double sy = 219.0 / 255.0, sc = 224.0 / 255.0;
  auto q = [](double v) { return (int)std::lrint(v * 256.0); };
  YuvMatrix m;
  m.yr = q(kr * sy); m.yg = q(kg * sy); m.yb = q(kb * sy);
  m.ur = q(-kr / (2.0 * (1.0 - kb)) * sc); m.ug = q(-kg / (2.0 * (1.0 - kb)) * sc); m.ub = q(0.5 * sc);
  m.vr = q(0.5 * sc); m.vg = q(-kg / (2.0 * (1.0 - kr)) * sc); m.vb = q(-kb / (2.0 * (1.0 - kr)) * sc);
  m.ry = m.gy = m.by = q(1.0 / sy);
  m.rv = q(2.0 * (1.0 - kr) / sc);
  m.gu = q(-2.0 * (1.0 - kb) * kb / kg / sc); m.gv = q(-2.0 * (1.0 - kr) * kr / kg / sc);
  m.bu = q(2.0 * (1.0 - kb) / sc);
  return m;
}
struct PlaneSet { const uint8_t *p[4]; long stride[4]; size_t row[4]; int rows[4]; int n; };
// I420 / A420 plane geometry
void planar_geometry(int fmt, int w, int h, PlaneSet *ps) {
  ps->n = fmt == B200VFX_FORMAT_A420 ? 4 : 3;
  ps->row[0] = (size_t)w; ps->rows[0] = h;
  ps->row[1] = ps->row[2] = (size_t)((w + 1) / 2); ps->rows[1] = ps->rows[2] = (h + 1) / 2;
  ps->row[3] = (size_t)w; ps->rows[3] = h;
}
}  // namespace

// ColorLut::transform_frame with the neighbouring videoconverts folded in: any 4-byte 8-bit RGB format in, any out
int b200vfx_colorlut_process_fmt(b200vfx_ctx *c, int in_fmt, int out_fmt, int width, int height, const void *src, int src_stride,
                                 void *dst, int dst_stride) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (in_fmt == B200VFX_FORMAT_RGBA && out_fmt == B200VFX_FORMAT_RGBA)
    return b200vfx_colorlut_process(c, in_fmt, width, height, src, src_stride, dst, dst_stride);
  if (!c->have_lut) return fail(c, B200VFX_ERR_NOT_NEGOTIATED, "No LUT configured");  // imp.rs:210-213
  PackedFmt sf, df;
  if (!packed_fmt(in_fmt, &sf) || !packed_fmt(out_fmt, &df) || sf.bpp != 4 || df.bpp != 4)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut (fused convert): formats %d -> %d are not both 4-byte 8-bit RGB formats", in_fmt, out_fmt);
  const size_t row = (size_t)width * 4;
  if (int rc = check_frame(c, width, height, src, src_stride, row, dst, dst_stride, row)) return rc;
  if (width == 0 || height == 0) return 0;
  DeviceGuard g(c->device);
  ColorLutFmtOp<true> op;
  op.in_sel = (uint32_t)sf.r | ((uint32_t)sf.g << 4) | ((uint32_t)sf.b << 8) | (4u << 12);
  {
    uint32_t nib[4];
    nib[df.r] = 0; nib[df.g] = 1; nib[df.b] = 2;
    const int fourth = 6 - df.r - df.g - df.b, src_fourth = 6 - sf.r - sf.g - sf.b;
    nib[fourth] = 4u + (uint32_t)src_fourth;                       // the source pixel's 4th byte ...
    op.out_sel = nib[0] | (nib[1] << 4) | (nib[2] << 8) | (nib[3] << 12);
    op.src_or = (sf.a >= 0 && df.a >= 0) ? 0u : (0xFFu << (8 * src_fourth));   // ... forced to 255 unless alpha -> alpha
  }
  Staged s{(const uint8_t *)src, src_stride, row, (uint8_t *)dst, dst_stride, row, height, false};
  return run_staged(c, s, [&](const uint8_t *ds, long dss, uint8_t *dd, long dds, int, int rows, cudaStream_t st) {
    if (!(aligned(ds, dss, 4) && aligned(dd, dds, 4))) return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut (fused convert): rows must be 4-byte aligned");
    const bool tables_ready = c->d_axis8 != nullptr && c->memo_ready;
    const bool pdl = pdl_admit(c->pdl && tables_ready, st, span_of(ds, dss, row, rows), span_of(dd, dds, row, rows));
    if (int rc = build_axis(c, st)) return rc;
    if (int rc = ensure_colorlut_memo(c, lut_dev(c), st)) return rc;
    op.memo = c->lut_kind == 3 ? c->d_memo : nullptr;
    op.memo1d = c->d_memo1d;
    int ww = width, hh = rows;
    if (dss == 4L * width && dds == 4L * width && (long long)width * rows < (1LL << 28)) { ww = width * rows; hh = 1; }
    const long long items = (long long)ceil_div(ww, 8 * 32 * kMapPx) * hh, cap = (long long)c->sm_count * c->memo_ctas;
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(items, cap)));
    const int linger = (pdl && items > cap) ? 1 : 0;
    if (linger) pdl_note_linger(st);
    if (c->lut_kind == 3) CU(c, launch_k(pdl, map_u32_kernel<ColorLutFmtOp<true>, kMapPx>, grid, dim3(256), 0, st, op, ds, dss, dd, dds, ww, hh, linger));
    else {
      ColorLutFmtOp<false> op1{op.memo, op.memo1d, op.in_sel, op.out_sel, op.src_or};
      CU(c, launch_k(pdl, map_u32_kernel<ColorLutFmtOp<false>, kMapPx>, grid, dim3(256), 0, st, op1, ds, dss, dd, dds, ww, hh, linger));
    }
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  });
}

int b200vfx_convert_packed(b200vfx_ctx *c, int src_fmt, int dst_fmt, int width, int height, const void *src, int src_stride,
                           void *dst, int dst_stride) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  PackedFmt sf, df;
  if (!packed_fmt(src_fmt, &sf) || !packed_fmt(dst_fmt, &df))
    return fail(c, B200VFX_ERR_UNSUPPORTED, "convert: formats %d -> %d are not both packed 8-bit RGB formats", src_fmt, dst_fmt);
  const size_t irow = (size_t)width * sf.bpp, orow = (size_t)width * df.bpp;
  if (int rc = check_frame(c, width, height, src, src_stride, irow, dst, dst_stride, orow)) return rc;
  if (width == 0 || height == 0) return 0;
  DeviceGuard g(c->device);
  Staged s{(const uint8_t *)src, src_stride, irow, (uint8_t *)dst, dst_stride, orow, height, false};
  return run_staged(c, s, [&](const uint8_t *ds, long dss, uint8_t *dd, long dds, int, int rows, cudaStream_t st) {
    const bool pdl = pdl_admit(c->pdl && sf.bpp == 4 && df.bpp == 4, st, span_of(ds, dss, irow, rows), span_of(dd, dds, orow, rows));
    if (sf.bpp == 4 && df.bpp == 4) {
      const int vec = (width % 4) == 0 && aligned(ds, dss, 16) && aligned(dd, dds, 16);
      const int gx = ceil_div(vec ? width / 4 : width, 256);
      dim3 grid((unsigned)gx, grid_rows_persistent(gx, rows, c->sm_count));
      CU(c, launch_k(pdl, swizzle44_kernel, grid, dim3(256), 0, st, ds, dss, dd, dds, width, rows, swizzle_selector(sf, df), vec));
    } else {
      const int gx = ceil_div(width, 256);
      dim3 grid((unsigned)gx, grid_rows_persistent(gx, rows, c->sm_count));
      swizzle_generic_kernel<<<grid, 256, 0, st>>>(ds, dss, sf, dd, dds, df, width, rows);
    }
    c->launches++;
    CU(c, cudaGetLastError());
    return 0;
  });
}

namespace {
// planar frames: every plane in HBM (kernel enqueued on the context stream) or every plane on the host (staged, synchronous)
int planes_location(b200vfx_ctx *c, const void *const *planes, int n, const void *packed, bool *on_device) {
  int dev = 0;
  for (int i = 0; i < n; i++) dev += is_device_ptr(planes[i]) ? 1 : 0;
  const bool pd = is_device_ptr(packed);
  if (!((dev == n && pd) || (dev == 0 && !pd))) return fail(c, B200VFX_ERR_INVALID, "convert: planes must be all device or all host memory");
  *on_device = pd;
  return 0;
}
}  // namespace

int b200vfx_convert_to_planar(b200vfx_ctx *c, int src_fmt, int dst_fmt, int width, int height, const void *src, int src_stride,
                              void *const *planes, const int *strides, int matrix) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  PackedFmt sf;
  if (!packed_fmt(src_fmt, &sf) || (dst_fmt != B200VFX_FORMAT_I420 && dst_fmt != B200VFX_FORMAT_A420))
    return fail(c, B200VFX_ERR_UNSUPPORTED, "convert: %d -> %d is not packed RGB -> I420 / A420", src_fmt, dst_fmt);
  if (!planes || !strides || (matrix != 0 && matrix != 601 && matrix != 709)) return fail(c, B200VFX_ERR_INVALID, "convert: bad argument");
  if (width <= 0 || height <= 0) return fail(c, B200VFX_ERR_INVALID, "convert: empty frame");
  PlaneSet ps;
  planar_geometry(dst_fmt, width, height, &ps);
  const size_t irow = (size_t)width * sf.bpp;
  if (!src || src_stride < 0 || (size_t)src_stride < irow) return fail(c, B200VFX_ERR_INVALID, "convert: bad source plane");
  for (int i = 0; i < ps.n; i++)
    if (!planes[i] || strides[i] < 0 || (size_t)strides[i] < ps.row[i]) return fail(c, B200VFX_ERR_INVALID, "convert: bad plane %d", i);
  DeviceGuard g(c->device);
  bool dev;
  if (int rc = planes_location(c, (const void *const *)planes, ps.n, src, &dev)) return rc;
  const YuvMatrix m = yuv_matrix(matrix, height);
  cudaStream_t st = dev ? c->stream() : c->s_k;
  const uint8_t *d_src = (const uint8_t *)src;
  long d_ss = src_stride;
  uint8_t *dp[4] = {nullptr, nullptr, nullptr, nullptr};
  long dstr[4] = {0, 0, 0, 0};
  if (dev) { for (int i = 0; i < ps.n; i++) { dp[i] = (uint8_t *)planes[i]; dstr[i] = strides[i]; } }
  else {
    CU(c, c->stage_in.reserve((size_t)src_stride * height));
    CU(c, cudaMemcpyAsync(c->stage_in.p, src, (size_t)src_stride * (height - 1) + irow, cudaMemcpyHostToDevice, st));
    d_src = c->stage_in.p;
    size_t total = 0;
    for (int i = 0; i < ps.n; i++) total += (size_t)strides[i] * ps.rows[i];
    CU(c, c->stage_out.reserve(total));
    size_t off = 0;
    for (int i = 0; i < ps.n; i++) { dp[i] = c->stage_out.p + off; dstr[i] = strides[i]; off += (size_t)strides[i] * ps.rows[i]; }
  }
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  dim3 block(32, 8), grid((unsigned)ceil_div((width + 1) / 2, 32), (unsigned)ceil_div((height + 1) / 2, 8));
  rgb_to_i420_kernel<<<grid, block, 0, st>>>(d_src, d_ss, sf, width, height, m, dp[0], dstr[0], dp[1], dstr[1], dp[2], dstr[2],
                                             ps.n == 4 ? dp[3] : nullptr, dstr[3]);
  c->launches++;
  CU(c, cudaGetLastError());
  if (!dev) {
    for (int i = 0; i < ps.n; i++)
      CU(c, cudaMemcpy2DAsync(planes[i], (size_t)strides[i], dp[i], (size_t)dstr[i], ps.row[i], (size_t)ps.rows[i], cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    pdl_forget(st);
  }
  return 0;
}

int b200vfx_convert_from_planar(b200vfx_ctx *c, int src_fmt, int dst_fmt, int width, int height, const void *const *planes,
                                const int *strides, void *dst, int dst_stride, int matrix) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  PackedFmt df;
  if (!packed_fmt(dst_fmt, &df) || (src_fmt != B200VFX_FORMAT_I420 && src_fmt != B200VFX_FORMAT_A420))
    return fail(c, B200VFX_ERR_UNSUPPORTED, "convert: %d -> %d is not I420 / A420 -> packed RGB", src_fmt, dst_fmt);
  if (!planes || !strides || (matrix != 0 && matrix != 601 && matrix != 709)) return fail(c, B200VFX_ERR_INVALID, "convert: bad argument");
  if (width <= 0 || height <= 0) return fail(c, B200VFX_ERR_INVALID, "convert: empty frame");
  PlaneSet ps;
  planar_geometry(src_fmt, width, height, &ps);
  const size_t orow = (size_t)width * df.bpp;
  if (!dst || dst_stride < 0 || (size_t)dst_stride < orow) return fail(c, B200VFX_ERR_INVALID, "convert: bad destination plane");
  for (int i = 0; i < ps.n; i++)
    if (!planes[i] || strides[i] < 0 || (size_t)strides[i] < ps.row[i]) return fail(c, B200VFX_ERR_INVALID, "convert: bad plane %d", i);
  DeviceGuard g(c->device);
  bool dev;
  if (int rc = planes_location(c, planes, ps.n, dst, &dev)) return rc;
  const YuvMatrix m = yuv_matrix(matrix, height);
  cudaStream_t st = dev ? c->stream() : c->s_k;
  const uint8_t *sp[4] = {nullptr, nullptr, nullptr, nullptr};
  long sstr[4] = {0, 0, 0, 0};
  uint8_t *d_dst = (uint8_t *)dst;
  if (dev) { for (int i = 0; i < ps.n; i++) { sp[i] = (const uint8_t *)planes[i]; sstr[i] = strides[i]; } }
  else {
    size_t total = 0;
    for (int i = 0; i < ps.n; i++) total += (size_t)strides[i] * ps.rows[i];
    CU(c, c->stage_in.reserve(total));
    size_t off = 0;
    for (int i = 0; i < ps.n; i++) {
      CU(c, cudaMemcpy2DAsync(c->stage_in.p + off, (size_t)strides[i], planes[i], (size_t)strides[i], ps.row[i], (size_t)ps.rows[i], cudaMemcpyHostToDevice, st));
      sp[i] = c->stage_in.p + off; sstr[i] = strides[i]; off += (size_t)strides[i] * ps.rows[i];
    }
    CU(c, c->stage_out.reserve((size_t)dst_stride * height));
    d_dst = c->stage_out.p;
  }
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  const int gx = ceil_div(width, 256);
  dim3 grid((unsigned)gx, grid_rows_persistent(gx, height, c->sm_count));
  i420_to_rgb_kernel<<<grid, 256, 0, st>>>(sp[0], sstr[0], sp[1], sstr[1], sp[2], sstr[2], ps.n == 4 ? sp[3] : nullptr, sstr[3], width, height, m,
                                           d_dst, dst_stride, df);
  c->launches++;
  CU(c, cudaGetLastError());
  if (!dev) {
    CU(c, cudaMemcpy2DAsync(dst, (size_t)dst_stride, d_dst, (size_t)dst_stride, orow, (size_t)height, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    pdl_forget(st);
  }
  return 0;
}

// ColorLut::transform_frame on a planar YUV frame: I420 -> RGB, the LUT, RGB -> I420 in one kernel (convert.cuh).  The A plane
// of an A420 frame is copied (colorlut copies alpha, imp.rs:262,291).
int b200vfx_colorlut_process_planar(b200vfx_ctx *c, int fmt, int width, int height, const void *const *src_planes,
                                    const int *src_strides, void *const *dst_planes, const int *dst_strides, int matrix) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!c->have_lut) return fail(c, B200VFX_ERR_NOT_NEGOTIATED, "No LUT configured");  // imp.rs:210-213
  if (fmt != B200VFX_FORMAT_I420 && fmt != B200VFX_FORMAT_A420)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut (planar): format %d is not I420 / A420", fmt);
  if (!src_planes || !src_strides || !dst_planes || !dst_strides || (matrix != 0 && matrix != 601 && matrix != 709))
    return fail(c, B200VFX_ERR_INVALID, "colorlut (planar): bad argument");
  if (width <= 0 || height <= 0) return fail(c, B200VFX_ERR_INVALID, "colorlut (planar): empty frame");
  PlaneSet ps;
  planar_geometry(fmt, width, height, &ps);
  for (int i = 0; i < ps.n; i++) {
    if (!src_planes[i] || src_strides[i] < 0 || (size_t)src_strides[i] < ps.row[i] || !dst_planes[i] || dst_strides[i] < 0 ||
        (size_t)dst_strides[i] < ps.row[i])
      return fail(c, B200VFX_ERR_INVALID, "colorlut (planar): bad plane %d", i);
    if (src_planes[i] == dst_planes[i]) return fail(c, B200VFX_ERR_INVALID, "colorlut (planar): in-place is not supported (the element is never in place)");
  }
  DeviceGuard g(c->device);
  bool sdev, ddev;
  if (int rc = planes_location(c, src_planes, ps.n, src_planes[0], &sdev)) return rc;
  if (int rc = planes_location(c, (const void *const *)dst_planes, ps.n, dst_planes[0], &ddev)) return rc;
  if (sdev != ddev) return fail(c, B200VFX_ERR_INVALID, "colorlut (planar): source and destination planes must both be device or both be host memory");
  const bool dev = sdev;
  const YuvMatrix m = yuv_matrix(matrix, height);
  cudaStream_t st = dev ? c->stream() : c->s_k;
  const uint8_t *sp[4] = {nullptr, nullptr, nullptr, nullptr};
  uint8_t *dp[4] = {nullptr, nullptr, nullptr, nullptr};
  long sstr[4] = {0, 0, 0, 0}, dstr[4] = {0, 0, 0, 0};
  if (dev) {
    for (int i = 0; i < ps.n; i++) { sp[i] = (const uint8_t *)src_planes[i]; sstr[i] = src_strides[i]; dp[i] = (uint8_t *)dst_planes[i]; dstr[i] = dst_strides[i]; }
  } else {   // staged copies keep the caller's strides, rounded up so that the vector kernel stays usable
    size_t tin = 0, tout = 0;
    for (int i = 0; i < ps.n; i++) { tin += (((size_t)src_strides[i] + 15) & ~(size_t)15) * ps.rows[i]; tout += (((size_t)dst_strides[i] + 15) & ~(size_t)15) * ps.rows[i]; }
    CU(c, c->stage_in.reserve(tin));
    CU(c, c->stage_out.reserve(tout));
    size_t oin = 0, oout = 0;
    for (int i = 0; i < ps.n; i++) {
      const size_t si = ((size_t)src_strides[i] + 15) & ~(size_t)15, so = ((size_t)dst_strides[i] + 15) & ~(size_t)15;
      CU(c, cudaMemcpy2DAsync(c->stage_in.p + oin, si, src_planes[i], (size_t)src_strides[i], ps.row[i], (size_t)ps.rows[i], cudaMemcpyHostToDevice, st));
      sp[i] = c->stage_in.p + oin; sstr[i] = (long)si; oin += si * ps.rows[i];
      dp[i] = c->stage_out.p + oout; dstr[i] = (long)so; oout += so * ps.rows[i];
    }
  }
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  if (int rc = build_axis(c, st)) return rc;
  if (int rc = ensure_colorlut_memo(c, lut_dev(c), st)) return rc;
  // forward matrix as dp4a byte weights (Y: unsigned, chroma: signed): true for BT.601 / BT.709 at 8 fractional bits
  auto ub = [](int v) { return v >= 0 && v <= 255; };
  auto sb = [](int v) { return v >= -128 && v <= 127; };
  if (!(ub(m.yr) && ub(m.yg) && ub(m.yb) && sb(m.ur) && sb(m.ug) && sb(m.ub) && sb(m.vr) && sb(m.vg) && sb(m.vb)) || m.ry != m.gy || m.ry != m.by)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "colorlut (planar): matrix coefficients do not fit the packed form");
  auto pack = [](int a, int b, int d) { return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(d & 255) << 16); };
  const YuvPacked w{pack(m.yr, m.yg, m.yb), pack(m.ur, m.ug, m.ub), pack(m.vr, m.vg, m.vb)};
  const bool l1d = c->lut_kind != 3;
  const bool x8 = (width % 8) == 0 && (height % 2) == 0 && aligned(sp[0], sstr[0], 8) && aligned(dp[0], dstr[0], 8) &&
                  aligned(sp[1], sstr[1], 4) && aligned(sp[2], sstr[2], 4) && aligned(dp[1], dstr[1], 4) && aligned(dp[2], dstr[2], 4);
  dim3 block(32, 8);
  if (x8) {
    dim3 grid((unsigned)ceil_div(width / 8, 32), (unsigned)ceil_div(height / 2, 8));
    if (l1d) colorlut_i420_x8_kernel<true><<<grid, block, 0, st>>>(c->d_memo, c->d_memo1d, sp[0], sstr[0], sp[1], sstr[1], sp[2], sstr[2], width, height, m, w, dp[0], dstr[0], dp[1], dstr[1], dp[2], dstr[2]);
    else colorlut_i420_x8_kernel<false><<<grid, block, 0, st>>>(c->d_memo, c->d_memo1d, sp[0], sstr[0], sp[1], sstr[1], sp[2], sstr[2], width, height, m, w, dp[0], dstr[0], dp[1], dstr[1], dp[2], dstr[2]);
  } else {
    dim3 grid((unsigned)ceil_div((width + 1) / 2, 32), (unsigned)ceil_div((height + 1) / 2, 8));
    if (l1d) colorlut_i420_kernel<true><<<grid, block, 0, st>>>(c->d_memo, c->d_memo1d, sp[0], sstr[0], sp[1], sstr[1], sp[2], sstr[2], width, height, m, w, dp[0], dstr[0], dp[1], dstr[1], dp[2], dstr[2]);
    else colorlut_i420_kernel<false><<<grid, block, 0, st>>>(c->d_memo, c->d_memo1d, sp[0], sstr[0], sp[1], sstr[1], sp[2], sstr[2], width, height, m, w, dp[0], dstr[0], dp[1], dstr[1], dp[2], dstr[2]);
  }
  c->launches++;
  CU(c, cudaGetLastError());
  if (ps.n == 4) CU(c, cudaMemcpy2DAsync(dp[3], (size_t)dstr[3], sp[3], (size_t)sstr[3], ps.row[3], (size_t)ps.rows[3], cudaMemcpyDeviceToDevice, st));
  if (!dev) {
    for (int i = 0; i < ps.n; i++)
      CU(c, cudaMemcpy2DAsync(dst_planes[i], (size_t)dst_strides[i], dp[i], (size_t)dstr[i], ps.row[i], (size_t)ps.rows[i], cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    pdl_forget(st);
  }
  return 0;
}

// RoundedCorners::prepare_output_buffer (border/imp.rs:482-559) for a device-resident pipeline: the reference appends the
// shared alpha GstMemory to the I420 buffer; device frames are plain plane pointers, so A420 = the three I420 planes + the
// mask plane copied into the output frame (device-to-device, asynchronous on the context stream)
int b200vfx_a420_append(b200vfx_ctx *c, int width, int height, const void *const *i420_planes, const int *i420_strides, const void *a8,
                        int a8_stride, void *const *out_planes, const int *out_strides) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!i420_planes || !i420_strides || !a8 || !out_planes || !out_strides || width <= 0 || height <= 0)
    return fail(c, B200VFX_ERR_INVALID, "a420_append: bad argument");
  PlaneSet ps;
  planar_geometry(B200VFX_FORMAT_A420, width, height, &ps);
  for (int i = 0; i < 4; i++) {
    const void *sp = i < 3 ? i420_planes[i] : a8;
    const int ss = i < 3 ? i420_strides[i] : a8_stride;
    if (out_planes[i] == sp) continue;   // plane shared with the input frame: nothing to copy
    if (int rc = b200vfx_copy_plane(c, out_planes[i], out_strides[i], sp, ss, ps.row[i], ps.rows[i])) return rc;
  }
  return 0;
}

// ---- videocompare: the other hash algorithms and blockhash on sizes that are not multiples of the hash grid -------------
namespace {
// stages a host plane on the device (whole plane, one copy) or passes a device plane through; returns the stream to use
int stage_plane(b200vfx_ctx *c, const void *src, int stride, size_t row_bytes, int height, const uint8_t **d_src, long *d_stride,
                cudaStream_t *st) {
  if (is_device_ptr(src)) { *d_src = (const uint8_t *)src; *d_stride = stride; *st = c->stream(); return 0; }
  const long ds = (long)((row_bytes + 15) & ~(size_t)15);
  CU(c, c->stage_in.reserve((size_t)ds * height));
  *st = c->s_k;
  if ((size_t)stride == row_bytes && (size_t)ds == row_bytes) CU(c, cudaMemcpyAsync(c->stage_in.p, src, row_bytes * (size_t)height, cudaMemcpyHostToDevice, *st));
  else CU(c, cudaMemcpy2DAsync(c->stage_in.p, (size_t)ds, src, (size_t)stride, row_bytes, (size_t)height, cudaMemcpyHostToDevice, *st));
  *d_src = c->stage_in.p; *d_stride = ds;
  return 0;
}
}  // namespace

int b200vfx_luma_resize(b200vfx_ctx *c, int fmt, int width, int height, const void *src, int stride, int nw, int nh, uint8_t *out) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (fmt != B200VFX_FORMAT_RGB && fmt != B200VFX_FORMAT_RGBA) return fail(c, B200VFX_ERR_UNSUPPORTED, "videocompare: format %d is not RGB / RGBA", fmt);
  if (!out || nw <= 0 || nh <= 0 || nw * nh > 4096) return fail(c, B200VFX_ERR_INVALID, "luma_resize: bad output size");
  const int bpp = fmt == B200VFX_FORMAT_RGB ? 3 : 4;
  const size_t row = (size_t)width * bpp;
  if (int rc = check_frame(c, width, height, src, stride, row, nullptr, 0, 0)) return rc;
  if (width <= 0 || height <= 0) return fail(c, B200VFX_ERR_INVALID, "luma_resize: empty frame");
  DeviceGuard g(c->device);
  const uint8_t *d_src; long d_stride; cudaStream_t st;
  if (int rc = stage_plane(c, src, stride, row, height, &d_src, &d_stride, &st)) return rc;
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  const bool same = nw == width && nh == height;
  // scratch: [tmp f32 width*nh][out nw*nh]
  const size_t n_tmp = same ? 0 : (size_t)width * nh;
  const size_t off_out = (n_tmp * 4 + 15) & ~(size_t)15, total = off_out + (size_t)nw * nh;
  CU(c, c->stage_sums.reserve(total + 64));
  uint8_t *base = c->stage_sums.p;
  uint8_t *d_out = base + off_out;
  if (same) {
    const int n = width * height;
    if (bpp == 4) luma_copy_kernel<4><<<ceil_div(n, 256), 256, 0, st>>>(d_src, d_stride, width, height, d_out);
    else luma_copy_kernel<3><<<ceil_div(n, 256), 256, 0, st>>>(d_src, d_stride, width, height, d_out);
    c->launches++;
  } else {
    auto taps_for = [&](int in_len, int out_len, b200vfx_ctx::TapsDev *out_t) -> int {
      auto it = c->taps_cache.find({in_len, out_len});
      if (it == c->taps_cache.end()) {
        if (c->taps_cache.size() >= 16) {   // bounded: a caps change brings new sizes
          CU(c, cudaDeviceSynchronize());
          for (auto &kv : c->taps_cache) { cudaFree(kv.second.taps); cudaFree(kv.second.meta); }
          c->taps_cache.clear();
        }
        const ResizeTaps t = make_resize_taps(in_len, out_len);
        std::vector<int2> meta((size_t)out_len);
        for (int i = 0; i < out_len; i++) meta[(size_t)i] = make_int2(t.left[(size_t)i], t.count[(size_t)i]);
        b200vfx_ctx::TapsDev d;
        d.max_taps = t.max_taps;
        CU(c, cudaMalloc(&d.taps, t.taps.size() * sizeof(float)));
        CU(c, cudaMalloc(&d.meta, meta.size() * sizeof(int2)));
        CU(c, cudaMemcpy(d.taps, t.taps.data(), t.taps.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(c, cudaMemcpy(d.meta, meta.data(), meta.size() * sizeof(int2), cudaMemcpyHostToDevice));
        it = c->taps_cache.emplace(std::make_pair(in_len, out_len), d).first;
      }
      *out_t = it->second;
      return 0;
    };
    b200vfx_ctx::TapsDev tv, th;
    if (int rc = taps_for(height, nh, &tv)) return rc;
    if (int rc = taps_for(width, nw, &th)) return rc;
    float *d_tmp = (float *)base;
    dim3 grid((unsigned)ceil_div(width, kVresCols), (unsigned)nh);
    const bool al = bpp == 4 && aligned(d_src, d_stride, 4);
    if (al) luma_vresize_kernel<4, true><<<grid, kVresThreads, 0, st>>>(d_src, d_stride, width, tv.taps, tv.meta, tv.max_taps, d_tmp);
    else if (bpp == 4) luma_vresize_kernel<4, false><<<grid, kVresThreads, 0, st>>>(d_src, d_stride, width, tv.taps, tv.meta, tv.max_taps, d_tmp);
    else luma_vresize_kernel<3, false><<<grid, kVresThreads, 0, st>>>(d_src, d_stride, width, tv.taps, tv.meta, tv.max_taps, d_tmp);
    luma_hresize_kernel<<<(unsigned)(nw * nh), 256, 0, st>>>(d_tmp, width, nw, nh, th.taps, th.meta, th.max_taps, d_out);
    c->launches += 2;
  }
  CU(c, cudaGetLastError());
  if (is_device_ptr(out)) {
    CU(c, cudaMemcpyAsync(out, d_out, (size_t)nw * nh, cudaMemcpyDeviceToDevice, st));
    CU(c, cudaStreamSynchronize(st));
  } else if (int rc = read_back_small(c, out, d_out, (size_t)nw * nh, st)) return rc;
  pdl_forget(st);
  return 0;
}

int b200vfx_blockhash_sums_f32(b200vfx_ctx *c, int fmt, int width, int height, const void *src, int stride, int hw, int hh, float *sums) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (fmt != B200VFX_FORMAT_RGB && fmt != B200VFX_FORMAT_RGBA) return fail(c, B200VFX_ERR_UNSUPPORTED, "videocompare: format %d is not RGB / RGBA", fmt);
  if (hw <= 0 || hh <= 0 || hw * hh > 4096 || !sums) return fail(c, B200VFX_ERR_INVALID, "blockhash: bad hash size");
  const int bpp = fmt == B200VFX_FORMAT_RGB ? 3 : 4;
  const size_t row = (size_t)width * bpp;
  if (int rc = check_frame(c, width, height, src, stride, row, nullptr, 0, 0)) return rc;
  // block sizes of at most one pixel make the reference's `x + 1. % block_width` weights genuinely fractional: such
  // frames (narrower or lower than the hash grid) are not served
  if (width <= hw || height <= hh)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "blockhash: a %dx%d frame is not larger than the %dx%d hash grid", width, height, hw, hh);
  DeviceGuard g(c->device);
  const uint8_t *d_src; long d_stride; cudaStream_t st;
  if (int rc = stage_plane(c, src, stride, row, height, &d_src, &d_stride, &st)) return rc;
  pdl_admit(false, st, Span{0, 0}, Span{0, 0});
  const float bwf = (float)width / (float)hw, bhf = (float)height / (float)hh;   // one IEEE division each, as the reference
  const int n = hw * hh;
  CU(c, c->stage_sums.reserve((size_t)n * 4));
  std::vector<float> host((size_t)n);
  const double max_block = (std::ceil((double)bwf) + 1.0) * (std::ceil((double)bhf) + 1.0) * 765.0;
  if (max_block < 16777216.0) {   // every partial sum is an integer below 2^24: f32 accumulation is exact, any order
    uint32_t *d_u = (uint32_t *)c->stage_sums.p;
    CU(c, cudaMemsetAsync(d_u, 0, (size_t)n * 4, st));
    const int rows_per_cta = std::max(1, ceil_div(height, c->sm_count * 4));
    const unsigned grid = (unsigned)ceil_div(height, rows_per_cta);
    if (bpp == 4) blockhash_frac_kernel<4><<<grid, 256, (size_t)n * 4, st>>>(d_src, d_stride, width, height, hw, hh, bwf, bhf, rows_per_cta, d_u);
    else blockhash_frac_kernel<3><<<grid, 256, (size_t)n * 4, st>>>(d_src, d_stride, width, height, hw, hh, bwf, bhf, rows_per_cta, d_u);
    c->launches++;
    CU(c, cudaGetLastError());
    std::vector<uint32_t> hu((size_t)n);
    CU(c, cudaMemcpyAsync(hu.data(), d_u, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) host[(size_t)i] = (float)hu[(size_t)i];
  } else {                        // sums leave the exact range: replay the raster-order chain per block
    float *d_f = (float *)c->stage_sums.p;
    if (bpp == 4) blockhash_seq_kernel<4><<<(unsigned)n, 32, 0, st>>>(d_src, d_stride, width, height, hw, hh, bwf, bhf, d_f);
    else blockhash_seq_kernel<3><<<(unsigned)n, 32, 0, st>>>(d_src, d_stride, width, height, hw, hh, bwf, bhf, d_f);
    c->launches++;
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(host.data(), d_f, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
  }
  pdl_forget(st);
  if (is_device_ptr(sums)) CU(c, cudaMemcpy(sums, host.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  else std::memcpy(sums, host.data(), (size_t)n * 4);
  return 0;
}

int b200vfx_hash_image(b200vfx_ctx *c, int algo, int fmt, int width, int height, const void *src, int stride, uint8_t *bits_out,
                       int *n_bits) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!bits_out || !n_bits) return fail(c, B200VFX_ERR_INVALID, "hash_image: null output");
  if (algo == B200VFX_HASH_BLOCKHASH) {
    if (width > 0 && height > 0 && width % 8 == 0 && height % 8 == 0) {   // integer fast path
      uint32_t sums[64];
      if (int rc = b200vfx_blockhash_sums(c, fmt, width, height, src, stride, 8, 8, sums)) return rc;
      b200vfx_blockhash_bits(sums, 8, 8, width, height, bits_out);
    } else {
      float sums[64];
      if (int rc = b200vfx_blockhash_sums_f32(c, fmt, width, height, src, stride, 8, 8, sums)) return rc;
      b200vfx_blockhash_bits_f32(sums, 8, 8, width, height, bits_out);
    }
    *n_bits = 64;
    return 0;
  }
  int nw = 0, nh = 0;
  if (b200vfx_hash_resize_dims(algo, &nw, &nh)) return fail(c, B200VFX_ERR_INVALID, "unknown hash algorithm %d", algo);
  uint8_t luma[96];
  if (int rc = b200vfx_luma_resize(c, fmt, width, height, src, stride, nw, nh, luma)) return rc;
  *n_bits = b200vfx_hash_bits_from_luma(algo, luma, nw, nh, bits_out);
  return 0;
}

// ---- colordetect ----------------------------------------------------------------------------------
int b200vfx_colordetect_histogram(b200vfx_ctx *c, int fmt, int width, int height, const void *src, int stride,
                                  int quality, uint32_t *hist) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  int bpp, fi;
  switch (fmt) {  // color_parts() of color-thief per ColorFormat (set_info, colordetect/imp.rs:268-275)
    case B200VFX_FORMAT_RGB: fi = 0; bpp = 3; break;
    case B200VFX_FORMAT_RGBA: fi = 1; bpp = 4; break;
    case B200VFX_FORMAT_ARGB: fi = 2; bpp = 4; break;
    case B200VFX_FORMAT_BGR: fi = 3; bpp = 3; break;
    case B200VFX_FORMAT_BGRA: fi = 4; bpp = 4; break;
    default: return fail(c, B200VFX_ERR_UNSUPPORTED, "colordetect: format %d is not RGB / RGBA / ARGB / BGR / BGRA", fmt);
  }
  if (!hist) return fail(c, B200VFX_ERR_INVALID, "colordetect: null histogram");
  if (quality < 1 || quality > 10) return fail(c, B200VFX_ERR_INVALID, "colordetect: quality %d outside 1..10 (color-thief range check)", quality);
  const size_t row = (size_t)width * bpp;
  if (int rc = check_frame(c, width, height, src, stride, row, nullptr, 0, 0)) return rc;
  DeviceGuard g(c->device);
  // the reference samples the FLAT plane slice: stride * height bytes, pixel i at byte i * bpp
  const size_t plane_bytes = (width > 0 && height > 0) ? (size_t)stride * (size_t)height : 0;
  const long long npix = (long long)(plane_bytes / (size_t)bpp);
  const long long nsamples = (npix + quality - 1) / quality;
  const bool src_dev = plane_bytes == 0 || is_device_ptr(src), hist_dev = is_device_ptr(hist);
  cudaStream_t st = src_dev ? c->stream() : c->s_k;
  if (!src_dev && hist_dev) if (int rc = order_after_ctx_stream(c, st)) return rc;   // device histogram, host frame
  const uint8_t *d_src = (const uint8_t *)src;
  if (!src_dev) {
    CU(c, c->stage_in.reserve(plane_bytes));
    CU(c, cudaMemcpyAsync(c->stage_in.p, src, plane_bytes, cudaMemcpyHostToDevice, st));
    d_src = c->stage_in.p;
  }
  uint32_t *d_hist = hist;
  const size_t nb = sizeof(uint32_t) * kColorDetectBins;
  if (!hist_dev) { CU(c, c->stage_sums.reserve(nb)); d_hist = (uint32_t *)c->stage_sums.p; }
  if (((uintptr_t)d_hist % 4) != 0) return fail(c, B200VFX_ERR_INVALID, "colordetect: histogram pointer is not 4-byte aligned");
  // consecutive frames overlap: the kernel reads the plane while the previous launch drains, histogram and counters after
  // its griddepcontrol.wait.  A frame staged from the host follows a copy: plain launch.
  const bool pdl = pdl_admit(c->pdl && src_dev && nsamples > 0 && st == c->stream(), st, Span{(uintptr_t)d_src, (uintptr_t)d_src + plane_bytes},
                             span_of(d_hist, (long)nb, nb, 1), true);
  unsigned *gsync = nullptr;
  if (int rc = reduce_scratch(c, 0, &gsync, nullptr)) return rc;
  gsync += 4;                                                     // words 4,5: colordetect's counters (word 0: blockhash ticket)
  if (st != c->stream()) if (int rc = order_after_ctx_stream(c, st)) return rc;   // one launch of this context at a time uses the counters
  if (nsamples <= 0) CU(c, cudaMemsetAsync(d_hist, 0, nb, st));   // nothing to count: the kernel (which zeroes the bins) does not run
  if (nsamples > 0) {
    const int mode = (bpp == 4 && quality == 1 && ((uintptr_t)d_src % 16) == 0) ? 2 : ((bpp == 4 && ((uintptr_t)d_src % 4) == 0) ? 1 : 0);
    using KernelT = void (*)(const uint8_t *, long long, int, uint32_t *, unsigned *);
    static const KernelT table[5][3] = {
        {colordetect_hist_kernel<0, 0>, nullptr, nullptr},
        {colordetect_hist_kernel<1, 0>, colordetect_hist_kernel<1, 1>, colordetect_hist_kernel<1, 2>},
        {colordetect_hist_kernel<2, 0>, colordetect_hist_kernel<2, 1>, colordetect_hist_kernel<2, 2>},
        {colordetect_hist_kernel<3, 0>, nullptr, nullptr},
        {colordetect_hist_kernel<4, 0>, colordetect_hist_kernel<4, 1>, colordetect_hist_kernel<4, 2>}};
    const KernelT k = table[fi][mode];
    const int cluster = (c->cd_cluster == 1 || c->cd_cluster == 2 || c->cd_cluster == 4 || c->cd_cluster == 8) ? c->cd_cluster : 2;
    // per (device, cluster size): opt in to 128 KB dynamic shared memory once, and ask how many clusters are co-resident
    static std::mutex mu;
    static bool attr_set[64] = {false};
    static int max_ctas[64][9] = {{0}};
    const int dev = (c->device >= 0 && c->device < 64) ? c->device : 0;
    int resident = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!attr_set[dev]) {
        for (int a = 0; a < 5; a++)
          for (int b = 0; b < 3; b++)
            if (table[a][b]) CU(c, cudaFuncSetAttribute(table[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb));
        attr_set[dev] = true;
      }
      if (max_ctas[dev][cluster] == 0) {
        int n = c->sm_count / cluster;
        if (cluster > 1) {
          cudaLaunchConfig_t q = {};
          q.gridDim = dim3((unsigned)(c->sm_count - c->sm_count % cluster)); q.blockDim = dim3(kColorDetectThreads); q.dynamicSmemBytes = nb;
          cudaLaunchAttribute qa[1];
          qa[0].id = cudaLaunchAttributeClusterDimension;
          qa[0].val.clusterDim.x = (unsigned)cluster; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
          q.attrs = qa; q.numAttrs = 1;
          int nc = 0;
          if (cudaOccupancyMaxActiveClusters(&nc, k, &q) == cudaSuccess && nc > 0) n = std::min(n, nc); else cudaGetLastError();
        }
        max_ctas[dev][cluster] = std::max(1, n) * cluster;
      }
      resident = max_ctas[dev][cluster];
    }
    const long long units = mode == 2 ? (nsamples >> 2) : nsamples;
    const long long per_cta = (long long)kColorDetectThreads * 4;
    // An asynchronous call (device histogram) is one of a train: taking only a fraction of the SMs leaves the rest to the
    // next launch (PDL), which then runs beside this one, and fewer CTAs flush fewer partial histograms with global atomics:
    // 4K, quality 10: 11.9 -> 7.8 us per frame (14.1 -> 10.3 on noise), quality 1 on noise 32.2 -> 15.8 us
    // (profiles/r02_colordetect_split.jsonl).  A call that returns the histogram to the host is alone: all SMs.
    const int split = c->cd_split > 0 ? c->cd_split : (quality <= 2 ? 4 : 2);
    const long long cap = (pdl && hist_dev && split > 1) ? std::max<long long>(cluster, resident / split) : resident;
    unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(cap, (units + per_cta - 1) / per_cta));
    const unsigned cl = grid >= (unsigned)cluster ? (unsigned)cluster : 1u;
    grid -= grid % cl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kColorDetectThreads); cfg.dynamicSmemBytes = nb; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    pdl_note_threads(st, (long long)grid * kColorDetectThreads);
    pdl_note_linger(st);   // every CTA passes griddepcontrol.wait before it retires
    CU(c, cudaLaunchKernelEx(&cfg, k, d_src, nsamples, quality, d_hist, gsync));
    c->launches++;
    CU(c, cudaGetLastError());
  }
  if (!hist_dev) {
    if (int rc = read_back_small(c, hist, d_hist, nb, st)) return rc;
  } else if (!src_dev) {
    CU(c, cudaStreamSynchronize(st));
  }
  return 0;
}

// ---- multi-GPU tile gather (tile_gather.cuh) ----------------------------------------------------
int b200vfx_peer_alloc(b200vfx_ctx *c, size_t bytes, void **dev_ptr, unsigned char handle_out[B200VFX_IPC_HANDLE_BYTES]) {
  if (!c || !dev_ptr) return fail(c, B200VFX_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == B200VFX_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(c->device);
  void *p = nullptr;
  CU(c, cudaMalloc(&p, bytes ? bytes : 1));
  CU(c, cudaMemset(p, 0, bytes ? bytes : 1));
  CU(c, cudaDeviceSynchronize());
  if (handle_out) {
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(c, B200VFX_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, sizeof h);
  }
  *dev_ptr = p;
  return 0;
}
int b200vfx_peer_free(b200vfx_ctx *c, void *dev_ptr) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  DeviceGuard g(c->device);
  if (dev_ptr) CU(c, cudaFree(dev_ptr));
  return 0;
}
int b200vfx_peer_open(b200vfx_ctx *c, const unsigned char handle[B200VFX_IPC_HANDLE_BYTES], void **dev_ptr) {
  if (!c || !handle || !dev_ptr) return fail(c, B200VFX_ERR_INVALID, "null argument");
  DeviceGuard g(c->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CU(c, cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int b200vfx_peer_close(b200vfx_ctx *c, void *dev_ptr) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  DeviceGuard g(c->device);
  if (dev_ptr) CU(c, cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}
int b200vfx_peer_enable_access(b200vfx_ctx *c, int peer_device) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (peer_device == c->device) return 0;
  DeviceGuard g(c->device);
  int can = 0;
  CU(c, cudaDeviceCanAccessPeer(&can, c->device, peer_device));
  if (!can) return fail(c, B200VFX_ERR_UNSUPPORTED, "device %d cannot access device %d", c->device, peer_device);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  CU(c, e);
  return 0;
}
int b200vfx_peer_status(b200vfx_ctx *c, const void *flags, uint32_t *error_epoch) {
  if (!c || !flags || !error_epoch) return fail(c, B200VFX_ERR_INVALID, "null argument");
  DeviceGuard g(c->device);
  CU(c, cudaStreamSynchronize(c->stream()));
  pdl_forget(c->stream());
  CU(c, cudaMemcpy(error_epoch, (const uint32_t *)flags + PF_ERR, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}

int b200vfx_colorlut_process_tile_gather(b200vfx_ctx *c, int fmt, int width, int tile_rows, const void *src,
                                         int src_stride, int world, int rank, void *const *frames, int frame_stride,
                                         int frame_row0, void *const *flags, uint32_t epoch) {
  return b200vfx_colorlut_process_tile_gather_mc(c, fmt, width, tile_rows, src, src_stride, world, rank, frames, nullptr,
                                                 frame_stride, frame_row0, flags, epoch);
}

int b200vfx_colorlut_process_tile_gather_mc(b200vfx_ctx *c, int fmt, int width, int tile_rows, const void *src,
                                            int src_stride, int world, int rank, void *const *frames, void *multicast_frame,
                                            int frame_stride, int frame_row0, void *const *flags, uint32_t epoch) {
  if (!c) return fail(nullptr, B200VFX_ERR_INVALID, "null context");
  if (!c->have_lut) return fail(c, B200VFX_ERR_NOT_NEGOTIATED, "No LUT configured");  // imp.rs:210-213
  if (fmt != B200VFX_FORMAT_RGBA || c->mode != 0)
    return fail(c, B200VFX_ERR_UNSUPPORTED, "tile gather: RGBA in memo mode only (format %d, mode %d)", fmt, c->mode);
  if (world < 1 || world > B200VFX_MAX_PEERS || rank < 0 || rank >= world)
    return fail(c, B200VFX_ERR_INVALID, "tile gather: bad world/rank %d/%d (max %d ranks)", world, rank, B200VFX_MAX_PEERS);
  if (!frames || !flags || epoch == 0 || width < 0 || tile_rows < 0 || frame_row0 < 0)
    return fail(c, B200VFX_ERR_INVALID, "tile gather: null pointer table, zero epoch or negative size");
  const size_t row = (size_t)width * 4;
  if (tile_rows > 0 && width > 0 && (!src || (size_t)std::abs(src_stride) < row || (size_t)frame_stride < row || frame_stride < 0))
    return fail(c, B200VFX_ERR_INVALID, "tile gather: stride smaller than a row (or negative frame stride)");
  if (tile_rows > 0 && width > 0 && !is_device_ptr(src)) return fail(c, B200VFX_ERR_INVALID, "tile gather: src must be device memory");
  PeerSet ps = {};
  for (int p = 0; p < world; p++) {
    if (!frames[p] || !flags[p]) return fail(c, B200VFX_ERR_INVALID, "tile gather: null frame/flag pointer for rank %d", p);
    ps.frame[p] = (uint8_t *)frames[p];
    ps.flags[p] = (uint32_t *)flags[p];
  }
  ps.world = world; ps.rank = rank; ps.epoch = epoch; ps.timeout_ms = (uint32_t)c->peer_timeout_ms;
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream();
  const long doff = (long)frame_row0 * frame_stride;
  // never PDL: the kernel's entry handshake tells the peers that everything before it on this stream has finished
  pdl_admit(false, st, span_of(src, src_stride, row, tile_rows), span_of(ps.frame[rank] + doff, frame_stride, row, tile_rows));
  if (int rc = build_axis(c, st)) return rc;
  if (int rc = ensure_colorlut_memo(c, lut_dev(c), st)) return rc;
  int w = width, h = tile_rows;
  long ss = src_stride, ds = frame_stride;
  if (w == 0 || h == 0) { w = 0; h = 1; }   // an empty tile still takes part in the handshake
  else if (ss == 4L * w && ds == 4L * w && (long long)w * h < (1LL << 28)) { w = w * h; h = 1; }  // packed: 1-D
  const bool al4 = w == 0 || (aligned(src, ss, 4) && (doff % 4) == 0 && (ds % 4) == 0);
  if (!al4) return fail(c, B200VFX_ERR_UNSUPPORTED, "tile gather: planes must be 4-byte aligned");
  bool vec = (w % 4) == 0 && (ds % 16) == 0 && (doff % 16) == 0;
  for (int p = 0; p < world; p++) {
    if ((uintptr_t)ps.frame[p] % 4) return fail(c, B200VFX_ERR_UNSUPPORTED, "tile gather: planes must be 4-byte aligned");
    if ((uintptr_t)ps.frame[p] % 16) vec = false;
  }
  const bool l1d = c->lut_kind != 3;
  const bool al16 = vec && (w == 0 || (aligned(src, ss, 16)));
  if (multicast_frame) {   // multimem.st moves 16 bytes: the vector path only
    if (!vec || (uintptr_t)multicast_frame % 16)
      return fail(c, B200VFX_ERR_UNSUPPORTED, "tile gather: the multicast path needs width %% 4 == 0 and 16-byte aligned frames");
    ps.mc = (uint8_t *)multicast_frame;
  }
  if (c->tg_path == 1 && al16 && !ps.mc) {   // TMA: one bulk store per destination out of the shared-memory tile
    if (int rc = launch_tile_gather_tma(c, l1d, (const uint8_t *)src, ss, ps, ds, doff, 4 * w, h, st)) return rc;
  } else {
    constexpr int PX = 8;
    const long long nwork = (long long)std::max(1, ceil_div(w, 8 * 32 * PX)) * h;
    dim3 grid((unsigned)std::min<long long>(nwork, (long long)c->sm_count * std::max(1, c->tg_ctas)));
#define LAUNCH_TG(V, L)                                                                                              \
  CU(c, launch_k(false, colorlut_tile_gather_kernel<PX, V, L>, grid, dim3(256), 0, st, c->d_memo, c->d_memo1d,       \
                 (const uint8_t *)src, ss, ps, ds, doff, w, h))
    bool v32 = vec && c->tg_cfg == 1 && !ps.mc && (w % 8) == 0 && (ds % 32) == 0 && (doff % 32) == 0;
    for (int p = 0; p < world; p++) v32 = v32 && (uintptr_t)ps.frame[p] % 32 == 0;
    if (v32) { if (l1d) LAUNCH_TG(2, true); else LAUNCH_TG(2, false); }
    else if (vec) { if (l1d) LAUNCH_TG(1, true); else LAUNCH_TG(1, false); }
    else { if (l1d) LAUNCH_TG(0, true); else LAUNCH_TG(0, false); }
#undef LAUNCH_TG
  }
  c->launches++;
  CU(c, cudaGetLastError());
  return 0;
}

}  // extern "C"
