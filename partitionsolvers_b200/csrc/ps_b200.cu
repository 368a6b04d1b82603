// ps_b200.cu -- C ABI (include/ps_b200.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side orchestration of one DPSolver construction (reference DP_impl.hpp:69-230):
//   sort_by_priority  -> sort_keys_kernel + stable LSD radix sort (radix_sort.cuh) + gather_kernel
//   createContext     -> build_rec_* (no n x n tables)
//   create()          -> level-1 kernel, then per level j=2..T: level_tile_kernel + combine_kernel
//   optimize()        -> backtrace_bounds_kernel + backtrace_scores_kernel
// Everything runs on the context's own non-blocking stream; the call blocks on that stream only.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: the functions are resolved at run time (nccl_api below)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/ps_b200.h"
#include "kernels.cuh"
#include "radix_sort.cuh"
#include "screen.cuh"
#include "screen_run.cuh"
#include "screen_select.cuh"

namespace {

constexpr int BR = 128;  // rows per tile == threads per CTA == records per shared-memory stage
constexpr int PS_MAX_BATCH_SLICE = 65535;  // instances per launch (grid.y)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

// Output of a solve staged in pinned memory: which ps_result field, where in the staging buffer, how many bytes.
struct StagedOut {
  int field;  // 0 perm, 1 bounds, 2 subset_scores, 3 max_score, 4 next_start
  size_t off, bytes;
};
// A small solve recorded as a CUDA graph (one cudaGraphLaunch instead of ~45 runtime calls), keyed by everything that shapes
// the sequence of launches.
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  unsigned long long epoch = 0;  // ps_ctx::alloc_epoch at capture time
  std::vector<StagedOut> staged;
  long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  int poisson = 0;  // the recorded solve assumed a > 0 as well as the safe range
  int seen = 0;     // eager solves seen with this key (the second one is recorded)
  int failed = 0;   // capture failed once: stay eager
};

struct ps_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host -> device copies of the later parts of a large batch (solve_pipelined)
  std::string err;
  ps_timings tm;
  // grow-only device workspaces
  DevBuf a_in, b_in, key_in, key_out, idx_in, perm, a_s, b_s, rec, SA, SB, part_val, part_arg, ms, ns,
      bounds, scores, sort_hist, scalar, recf, lb, scr_stats, scr_counters, tile_map, scr_bmax, fxinfo, pfx, cpA, cpB, ipos,
      live, rb_info, live_hdr, kbest, rb_slots, rb_cnt,  // live-tile protocol
      shard_send, shard_recv;  // row-sharded solves: packed level chunk, gathered chunks
  int map_key[4] = {-1, -1, -1, -1};  // (NCH, G, NRB, ntiles) the tile map was built for
  std::vector<cudaEvent_t> events;
  size_t ev_used = 0;
  long long launches = 0;
  int* pinned_flags = nullptr;  // [0] range/exactness flags, [2..5] screen counters (as 2 x u64)
  unsigned long long* pinned_levels = nullptr;  // per level: cumulative exact-cell counter (screen give-up rule)
  unsigned char* pinned_out = nullptr;          // staging for small results (PINNED_OUT_BYTES): see copy_out in solve_impl
  size_t pinned_levels_cap = 0;
  int last_fast = 0;
  int dbg_level = 0;  // development: record per-tile times of this level (PS_DEBUG_LEVEL)
  DevBuf dbg;
  int sm_count = 148;
  cudaEvent_t wait_event = nullptr;  // ps_stream_wait
  // small solves replayed from CUDA graphs (solve_small_graph below)
  unsigned long long alloc_epoch = 0;  // bumped whenever a workspace is reallocated: graphs hold the old pointers
  bool capturing = false;              // solve_impl is being recorded into a graph: no events, no syncs, no host reads
  std::vector<StagedOut> cap_staged;   // outputs of the recorded solve inside pinned_out
  unsigned char* pinned_in = nullptr;  // inputs of a replayed solve (PINNED_IN_BYTES)
  std::map<std::array<long long, 12>, GraphEntry> graphs;
  int graphs_enabled = 1;              // PS_NO_GRAPH=1 switches the replay off
};

// NCCL is bound at run time: a host that already carries an NCCL (PyTorch bundles its own libnccl.so.2) keeps exactly one
// copy in the process, and a host without multi-GPU needs never loads it.  Order: the copy already loaded (RTLD_NOLOAD),
// $PS_NCCL_LIB, the system's libnccl.so.2.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string err;
  bool ok = false;
};

struct ps_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

struct ps_shard {
  ps_ctx* ctx = nullptr;
  int dtype = 0;  // 4 or 8
  ps_problem p;
  int rank = 0, world = 1;
  int L = 0, nch_alloc = 0, owned_blocks = 0, chunk_rows = 0;
  void* rec = nullptr;   // [2][n+1]
  void* SA = nullptr;
  void* SB = nullptr;
  void* part_val = nullptr;
  int* part_arg = nullptr;
  int* ns = nullptr;     // [T+1][n] (owned rows filled)
  int cur = 0;           // record buffer holding the previous level in its p field
  int fast = 0;          // inputs passed the safe-range precheck (math_fast.cuh)
  int screen = 0;        // levels run the screened kernel (screen.cuh): prefix records, float screening records
  float4* recf = nullptr;   // double only: [2][n+1] screening records
  void* lb = nullptr;       // [n] row lower bounds of the level in flight
  float* bmax = nullptr;    // [n/128 + 2] block maxima of pm
  int* dflags = nullptr;    // [2] device flags: range/exactness, sortedness
  int2* tile_map = nullptr;
  int map_key[4] = {-1, -1, -1, -1};
  unsigned long long* counters = nullptr;
  int4* live = nullptr;        // live-tile protocol of the screened levels (screen.cuh)
  unsigned int* rb_info = nullptr;  // per-tile stage masks
  ps::LiveHdr* live_hdr = nullptr;  // one header per level
  int* kbest = nullptr;             // per row: position of the best lower-bound sample (cost hint)
  int* rb_slots = nullptr;          // per row block: slots that hold a partial, ascending
  int* rb_cnt = nullptr;
};

namespace {

#define PS_CUDA(ctx, expr)                                                                     \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
      return (e__ == cudaErrorMemoryAllocation) ? PS_ERR_NOMEM : PS_ERR_CUDA;                  \
    }                                                                                          \
  } while (0)

#define PS_TRY(expr)                 \
  do {                               \
    ps_status s__ = (expr);          \
    if (s__ != PS_OK) return s__;    \
  } while (0)

ps_status ensure(ps_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return PS_OK;
  if (ctx->capturing) {  // a recorded solve must find every workspace in place (the eager solve before it sized them)
    ctx->err = "graph capture: a workspace would have to grow";
    return PS_ERR_CUDA;
  }
  ctx->alloc_epoch++;
  if (b.p) {
    PS_CUDA(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + (bytes >> 3) + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    ctx->err = "cudaMalloc(" + std::to_string(bytes) + " B): " + cudaGetErrorString(e);
    return PS_ERR_NOMEM;
  }
  b.cap = want;
  return PS_OK;
}

ps_status mark(ps_ctx* ctx, cudaEvent_t* out) {
  if (ctx->capturing) {  // timings of a replayed solve come from two events around the graph launch
    *out = nullptr;
    return PS_OK;
  }
  if (ctx->ev_used == ctx->events.size()) {
    cudaEvent_t e;
    PS_CUDA(ctx, cudaEventCreate(&e));
    ctx->events.push_back(e);
  }
  cudaEvent_t e = ctx->events[ctx->ev_used++];
  PS_CUDA(ctx, cudaEventRecord(e, ctx->stream));
  *out = e;
  return PS_OK;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

void scatter_staged(const ps_ctx* ctx, const std::vector<StagedOut>& staged, const ps_result* res) {
  void* dst[5] = {res->perm, res->bounds, res->subset_scores, res->max_score, res->next_start};
  for (const StagedOut& sg : staged)
    if (dst[sg.field]) memcpy(dst[sg.field], ctx->pinned_out + sg.off, sg.bytes);
}

NcclApi& nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    const char* names[3] = {"libnccl.so.2", getenv("PS_NCCL_LIB"), "libnccl.so.2"};
    const int flags[3] = {RTLD_NOW | RTLD_NOLOAD, RTLD_NOW, RTLD_NOW};
    for (int q = 0; q < 3 && !a.handle; ++q)
      if (names[q]) a.handle = dlopen(names[q], flags[q]);
    if (!a.handle) {
      a.err = "libnccl.so.2 not found (set PS_NCCL_LIB)";
      return a;
    }
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(a.handle, "ncclGetVersion");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
    if (!a.ok) a.err = "libnccl.so.2 lacks a required symbol";
    return a;
  }();
  return api;
}

#define PS_NCCL(ctx, expr)                                                                     \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) {                                                                  \
      (ctx)->err = std::string(#expr) + ": " + nccl_api().GetErrorString(r__);                 \
      return PS_ERR_COMM;                                                                      \
    }                                                                                          \
  } while (0)

// k-chunk length L: a multiple of the tile height Q (128, or 256 for the two-rows-per-thread kernel) giving enough
// equal tiles to balance 148 SMs (SURVEY.md hard part 5).  Measured on C3: 1024..1280-wide chunks beat 4096 by
// 5 % (finer tail); small instances get the smallest legal chunk so that a level still spreads over many SMs.
int choose_chunk(int n, int batch, int user, int Q) {
  if (user > 0) return std::max(Q, (user / Q) * Q);
  const long long want_tiles = 148LL * 256;
  const int nrb = cdiv(n, Q);
  long long nch = (2 * want_tiles + (long long)batch * nrb - 1) / ((long long)batch * nrb);
  nch = std::max(1LL, std::min(128LL, nch));
  long long L = ((cdiv(n, nch) + Q - 1) / Q) * (long long)Q;
  L = std::min<long long>(L, 16384);
  L = std::min<long long>(L, ((n + Q - 1) / Q) * (long long)Q);
  return (int)std::max<long long>(L, Q);
}

// The screened kernels' error budget rests on two properties of the MUFU unit (screen.cuh header: |x rcp(x) - 1| <= 4u,
// |lg2 x - log2 x| <= 4u max(1, |log2 x|), u = 2^-24; measured 1.65u / 3.85u on B200).  They are verified on the device at
// hand -- every float of the safe range, ~20 ms -- the first time a context on that device wants a screened level; the verdict
// is cached per device for the life of the process.  A device that fails keeps the exhaustive kernel (same tables, slower).
int g_mufu_state[64] = {0};  // 0 unknown, 1 claims hold, -1 claims fail
double g_mufu_err[64][2];
std::mutex g_mufu_mutex;
bool mufu_claims_hold(ps_ctx* ctx) {
  const int d = ctx->device;
  if (d < 0 || d >= 64) return false;
  std::lock_guard<std::mutex> lk(g_mufu_mutex);
  if (g_mufu_state[d] == 0) {
    double out[2] = {1e9, 1e9};
    bool ran = ensure(ctx, ctx->scalar, 64) == PS_OK;
    ran = ran && cudaMemsetAsync(ctx->scalar.p, 0, 2 * sizeof(double), ctx->stream) == cudaSuccess;
    if (ran) {
      ps::selftest_mufu_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((double*)ctx->scalar.p);
      ran = cudaMemcpyAsync(out, ctx->scalar.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
            cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    g_mufu_err[d][0] = out[0];
    g_mufu_err[d][1] = out[1];
    g_mufu_state[d] = (ran && out[0] > 0.0 && out[0] <= 4.0 && out[1] > 0.0 && out[1] <= 4.0) ? 1 : -1;
  }
  return g_mufu_state[d] == 1;
}

// Small instances: with 128-row tiles a level of n = 1000 is 36 work items whose threads each walk 128 cells one after the
// other (23 us per level, 148 SMs mostly idle).  32-row tiles of 32 cells give ~500 one-warp work items and a chain a quarter
// as long; same kernels, same order of evaluation inside a row, same partials folded in ascending k.
constexpr int SMALL_ROWS = 32;
constexpr size_t PINNED_OUT_BYTES = 256 * 1024;  // results up to this size are staged through pinned memory (solve_impl)

template <typename D>
struct LevelLaunch {
  using Fn = void (*)(ps::Geom, const ps::Rec4<D>*, const D*, const D*, D*, int*);
  static Fn pick(int dist, int rp, int running, int fast) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                              \
  return fast ? (running ? (Fn)level_tile_kernel<D, DI, RPV, true, true, BR>                          \
                         : (Fn)level_tile_kernel<D, DI, RPV, false, true, BR>)                        \
              : (running ? (Fn)level_tile_kernel<D, DI, RPV, true, false, BR>                         \
                         : (Fn)level_tile_kernel<D, DI, RPV, false, false, BR>);
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
  // small instances (SMALL_ROWS-row tiles, chunk length SMALL_ROWS; range-prechecked math only)
  static Fn pick_small(int dist, int rp, int running) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                  \
  return running ? (Fn)level_tile_kernel<D, DI, RPV, true, true, SMALL_ROWS>              \
                 : (Fn)level_tile_kernel<D, DI, RPV, false, true, SMALL_ROWS>;
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
};

template <typename D>
struct LastLevelLaunch {
  using Fn = void (*)(ps::Geom, const ps::Rec4<D>*, const D*, const D*, D*, int*);
  static Fn pick(int dist, int rp, int running, int fast) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                            \
  return fast ? (running ? (Fn)last_level_kernel<D, DI, RPV, true, true, 128>                       \
                         : (Fn)last_level_kernel<D, DI, RPV, false, true, 128>)                     \
              : (running ? (Fn)last_level_kernel<D, DI, RPV, true, false, 128>                      \
                         : (Fn)last_level_kernel<D, DI, RPV, false, false, 128>);
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
};

template <typename D>
struct ScreenLaunch {
  // screened level, exactly summable inputs (screen.cuh): select kernel (range tiers -> compacted list of live tiles) and
  // the persistent live-tile kernel (group / cell tiers + the fold of a finished row block)
  using SelFn = void (*)(ps::Geom, const ps::Rec4<D>*, const D*, const float*, int, const int*, int4*, long long,
                         unsigned int*, int, ps::LiveHdr*, unsigned long long*, const int*, int*, int*, ps::SelSrc);
  using Fn = void (*)(ps::Geom, const ps::Rec4<D>*, const float4*, const D*, D*, int*, unsigned long long*, const int*,
                      const int4*, long long, ps::LiveHdr*, unsigned long long*);
  using LbFn = void (*)(ps::Geom, const ps::Rec4<D>*, float4*, D*, float*, int, int*);
  static Fn pick(int dist, int rp, int audit) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                             \
  return audit ? (Fn)level_live_screen_kernel<D, DI, RPV, true, BR>                  \
               : (Fn)level_live_screen_kernel<D, DI, RPV, false, BR>;
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
  static SelFn pick_select(int dist, int rp, int audit, int tpr) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                               \
  return audit ? (tpr == 4 ? (SelFn)screen_select_kernel<D, DI, RPV, true, 4, SEL_PREFIX>              \
                           : (SelFn)screen_select_kernel<D, DI, RPV, true, 1, SEL_PREFIX>)             \
               : (tpr == 4 ? (SelFn)screen_select_kernel<D, DI, RPV, false, 4, SEL_PREFIX>             \
                           : (SelFn)screen_select_kernel<D, DI, RPV, false, 1, SEL_PREFIX>);
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
  static LbFn pick_lb(int dist, int rp) {
    using namespace ps;
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: return (LbFn)screen_lb_kernel<D, DIST_GAUSSIAN, false>;
      case 1: return (LbFn)screen_lb_kernel<D, DIST_GAUSSIAN, true>;
      case 2: return (LbFn)screen_lb_kernel<D, DIST_POISSON, false>;
      case 3: return (LbFn)screen_lb_kernel<D, DIST_POISSON, true>;
      default: return (LbFn)screen_lb_kernel<D, DIST_RATIONAL, false>;
    }
  }
};

struct ScreenRunLaunch {
  // double inputs whose sums round (screen_run.cuh): live-tile protocol with the fixed-point select kernel
  using Fn = void (*)(ps::Geom, const ps::Rec4<double>*, const float4*, const longlong2*, const ps::FxInfo*,
                      const double*, const double*, const double*, double*, int*, unsigned long long*, const int*,
                      const int4*, long long, ps::LiveHdr*);
  using LbFn = void (*)(ps::Geom, const ps::Rec4<double>*, float4*, const double*, const double*, double*, float*, int,
                        int*);
  using SelFn = ScreenLaunch<double>::SelFn;
  static SelFn pick_select(int dist, int rp, int audit, int tpr) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                                      \
  return audit ? (tpr == 4 ? (SelFn)screen_select_kernel<double, DI, RPV, true, 4, SEL_FIXED>                 \
                           : (SelFn)screen_select_kernel<double, DI, RPV, true, 1, SEL_FIXED>)                \
               : (tpr == 4 ? (SelFn)screen_select_kernel<double, DI, RPV, false, 4, SEL_FIXED>                \
                           : (SelFn)screen_select_kernel<double, DI, RPV, false, 1, SEL_FIXED>);
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
  static Fn pick(int dist, int rp, int audit) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                          \
  return audit ? (Fn)level_live_screen_run_kernel<DI, RPV, true, BR>              \
               : (Fn)level_live_screen_run_kernel<DI, RPV, false, BR>;
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: PS_PICK(DIST_GAUSSIAN, false)
      case 1: PS_PICK(DIST_GAUSSIAN, true)
      case 2: PS_PICK(DIST_POISSON, false)
      case 3: PS_PICK(DIST_POISSON, true)
      default: PS_PICK(DIST_RATIONAL, false)
    }
#undef PS_PICK
  }
  static LbFn pick_lb(int dist, int rp) {
    using namespace ps;
    switch (dist * 2 + (rp ? 1 : 0)) {
      case 0: return (LbFn)screen_lb_run_kernel<DIST_GAUSSIAN, false>;
      case 1: return (LbFn)screen_lb_run_kernel<DIST_GAUSSIAN, true>;
      case 2: return (LbFn)screen_lb_run_kernel<DIST_POISSON, false>;
      case 3: return (LbFn)screen_lb_run_kernel<DIST_POISSON, true>;
      default: return (LbFn)screen_lb_run_kernel<DIST_RATIONAL, false>;
    }
  }
};

struct ScreenRun32Launch {
  // float inputs whose sums round (screen_run.cuh): live-tile protocol, bounds from the running-sum checkpoints
  using Fn = void (*)(ps::Geom, const ps::Rec4<float>*, const float*, const float*, long long, const int*, const float*,
                      float*, int*, unsigned long long*, const int*, const int4*, long long, ps::LiveHdr*);
  using LbFn = void (*)(ps::Geom, const ps::Rec4<float>*, float4*, const float*, const float*, long long, float*, float*,
                        int, int*);
  using SelFn = ScreenLaunch<float>::SelFn;
  static bool supported(int, int) { return true; }
  static SelFn pick_select(int dist, int rp, int audit, int tpr) {
    using namespace ps;
#define PS_PICK(DI, RPV)                                                                                      \
  return audit ? (tpr == 4 ? (SelFn)screen_select_kernel<float, DI, RPV, true, 4, SEL_RUN32>                  \
                           : (SelFn)screen_select_kernel<float, DI, RPV, true, 1, SEL_RUN32>)                 \
               : (tpr == 4 ? (SelFn)screen_select_kernel<float, DI, RPV, false, 4, SEL_RUN32>                 \
                           : (SelFn)screen_select_kernel<float, DI, RPV, false, 1, SEL_RUN32>);
    if (dist == PS_DIST_GAUSSIAN && !rp) PS_PICK(DIST_GAUSSIAN, false)
    if (dist == PS_DIST_GAUSSIAN && rp) PS_PICK(DIST_GAUSSIAN, true)
    if (dist == PS_DIST_POISSON && !rp) PS_PICK(DIST_POISSON, false)
    if (dist == PS_DIST_POISSON && rp) PS_PICK(DIST_POISSON, true)
    PS_PICK(DIST_RATIONAL, false)
#undef PS_PICK
  }
  static Fn pick(int dist, int rp, int audit) {
    using namespace ps;
    if (dist == PS_DIST_GAUSSIAN && !rp)
      return audit ? (Fn)level_live_screen_run32_kernel<DIST_GAUSSIAN, false, true, BR>
                   : (Fn)level_live_screen_run32_kernel<DIST_GAUSSIAN, false, false, BR>;
    if (dist == PS_DIST_GAUSSIAN && rp)
      return audit ? (Fn)level_live_screen_run32_kernel<DIST_GAUSSIAN, true, true, BR>
                   : (Fn)level_live_screen_run32_kernel<DIST_GAUSSIAN, true, false, BR>;
    if (dist == PS_DIST_POISSON && !rp)
      return audit ? (Fn)level_live_screen_run32_kernel<DIST_POISSON, false, true, BR>
                   : (Fn)level_live_screen_run32_kernel<DIST_POISSON, false, false, BR>;
    if (dist == PS_DIST_POISSON && rp)
      return audit ? (Fn)level_live_screen_run32_kernel<DIST_POISSON, true, true, BR>
                   : (Fn)level_live_screen_run32_kernel<DIST_POISSON, true, false, BR>;
    return audit ? (Fn)level_live_screen_run32_kernel<DIST_RATIONAL, false, true, BR>
                 : (Fn)level_live_screen_run32_kernel<DIST_RATIONAL, false, false, BR>;
  }
  static LbFn pick_lb(int dist, int rp) {
    using namespace ps;
    if (dist == PS_DIST_GAUSSIAN && !rp) return (LbFn)screen_lb_run32_kernel<DIST_GAUSSIAN, false>;
    if (dist == PS_DIST_GAUSSIAN && rp) return (LbFn)screen_lb_run32_kernel<DIST_GAUSSIAN, true>;
    if (dist == PS_DIST_POISSON && !rp) return (LbFn)screen_lb_run32_kernel<DIST_POISSON, false>;
    if (dist == PS_DIST_POISSON && rp) return (LbFn)screen_lb_run32_kernel<DIST_POISSON, true>;
    return (LbFn)screen_lb_run32_kernel<DIST_RATIONAL, false>;
  }
};

ps::Geom make_geom(int n, int j, int Tc, int L, int nch_alloc, int rb_first, int rb_step, int tile_rows = BR) {
  ps::Geom g;
  g.n = n;
  g.kmax = n - j + 1;
  g.nrows = (j == Tc) ? 1 : (n - j + 1);  // DP_impl.hpp:114-116: last level needs row 0 only
  g.L = L;
  g.NCH = cdiv(g.kmax, L);
  g.G = L / tile_rows;
  const int nrb_total = cdiv(g.nrows, tile_rows);
  g.NRB = nrb_total > rb_first ? (nrb_total - rb_first + rb_step - 1) / rb_step : 0;
  g.nch_alloc = nch_alloc;
  g.rec_stride = (long long)n + 1;
  g.rb_first = rb_first;
  g.rb_step = rb_step;
  g.ntiles = ps::geom_count_tiles(g.NCH, g.G, g.NRB, rb_first, rb_step);
  g.tile_map = nullptr;
  g.sa_mul = 1;
  g.sa_nch = nch_alloc;
  g.nsub = 1;
  return g;
}

long long level_cells(int n, int j, int Tc) {
  long long m = n - j + 1;
  return (j == Tc) ? m : m * (m + 1) / 2;
}

template <typename D>
ps_status sort_stage(ps_ctx* ctx, const ps_problem& p, const D* d_a, const D* d_b, int R, int n) {
  const long long total = (long long)R * n;
  cudaStream_t st = ctx->stream;
  PS_TRY(ensure(ctx, ctx->perm, total * sizeof(int)));
  PS_TRY(ensure(ctx, ctx->a_s, total * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->b_s, total * sizeof(D)));
  int* d_perm = (int*)ctx->perm.p;
  if (p.sort_mode == PS_SORT_GIVEN) {
    PS_CUDA(ctx, cudaMemcpyAsync(d_perm, p.perm, total * sizeof(int),
                                 p.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!p.on_device) ctx->tm.h2d_bytes += total * sizeof(int);
  } else if (n <= ps::SORT_SMALL_MAX) {
    // small instances: keys + bitonic sort of (key, index) + gather fused in one kernel, one CTA per instance
    using U = typename ps::SortKey<D>::U;
    const size_t smem = (size_t)ps::sort_small_pow2(n) * (sizeof(U) + sizeof(int));
    auto kfn = ps::sort_small_kernel<D>;
    if (smem > 48 * 1024)
      PS_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<R, ps::SORT_SMALL_THREADS, smem, st>>>(n, d_a, d_b, d_perm, (D*)ctx->a_s.p, (D*)ctx->b_s.p);
    ctx->launches++;
    PS_CUDA(ctx, cudaGetLastError());
    return PS_OK;
  } else {
    // stable LSD radix sort of (a/b, index), all instances at once (radix_sort.cuh)
    using U = typename ps::SortKey<D>::U;
    const int tiles = cdiv(n, ps::SORT_TILE);
    PS_TRY(ensure(ctx, ctx->key_in, total * sizeof(U)));
    PS_TRY(ensure(ctx, ctx->key_out, total * sizeof(U)));
    PS_TRY(ensure(ctx, ctx->idx_in, total * sizeof(int)));
    PS_TRY(ensure(ctx, ctx->sort_hist, (size_t)R * 256 * tiles * sizeof(unsigned int)));
    U* kbuf[2] = {(U*)ctx->key_in.p, (U*)ctx->key_out.p};
    int* ibuf[2] = {d_perm, (int*)ctx->idx_in.p};  // an even number of passes ends in d_perm
    unsigned int* hist = (unsigned int*)ctx->sort_hist.p;
    ps::sort_keys_kernel<D><<<cdiv(total, 256), 256, 0, st>>>(total, n, d_a, d_b, kbuf[0], ibuf[0]);
    ctx->launches++;
    static_assert(ps::SortKey<D>::PASSES % 2 == 0, "result must land in the first buffer");
    for (int pass = 0; pass < ps::SortKey<D>::PASSES; ++pass) {
      const int src = pass & 1, dst = src ^ 1;
      ps::sort_histogram_kernel<U><<<dim3(tiles, R), ps::SORT_THREADS, 0, st>>>(n, tiles, 8 * pass, kbuf[src], hist);
      ps::sort_scan_kernel<<<R, ps::SORT_THREADS, 0, st>>>(tiles, hist);
      ps::sort_scatter_kernel<U><<<dim3(tiles, R), ps::SORT_THREADS, 0, st>>>(n, tiles, 8 * pass, kbuf[src], ibuf[src],
                                                                             kbuf[dst], ibuf[dst], hist);
      ctx->launches += 3;
    }
  }
  ps::gather_kernel<D><<<cdiv(total, 256), 256, 0, st>>>(total, n, d_perm, d_a, d_b, (D*)ctx->a_s.p,
                                                         (D*)ctx->b_s.p);
  ctx->launches++;
  PS_CUDA(ctx, cudaGetLastError());
  return PS_OK;
}

template <typename D>
ps_status solve_impl(ps_ctx* ctx, const ps_problem* pp, const D* a, const D* b, ps_result* res,
                     const D* staged_a = nullptr, const D* staged_b = nullptr, const ps_comm* comm = nullptr);

// timings of a call that was executed as several sub-calls
void add_timings(ps_timings& agg, const ps_timings& t) {
  agg.total_ms += t.total_ms;
  agg.h2d_ms += t.h2d_ms;
  agg.sort_scan_ms += t.sort_scan_ms;
  agg.prepass_ms += t.prepass_ms;
  agg.levels_ms += t.levels_ms;
  agg.backtrace_ms += t.backtrace_ms;
  agg.d2h_ms += t.d2h_ms;
  agg.n_levels = t.n_levels;
  for (int q = 0; q < t.n_levels; ++q) {
    agg.level_ms[q] += t.level_ms[q];
    agg.level_cells[q] += t.level_cells[q];
  }
  agg.kernel_launches += t.kernel_launches;
  agg.h2d_bytes += t.h2d_bytes;
  agg.d2h_bytes += t.d2h_bytes;
  agg.screen_used = t.screen_used;
  agg.screen_exact_cells += t.screen_exact_cells;
  agg.screen_violations += t.screen_violations;
  agg.screen_fallback_level = std::max(agg.screen_fallback_level, t.screen_fallback_level);
  for (int q = 0; q < 6; ++q) agg.screen_tiers[q] += t.screen_tiers[q];
}

// sub-problem / sub-result of instances [r0, r0 + rk) of a batch (plain pointer arithmetic: host or device buffers alike)
template <typename D>
void slice_batch(const ps_problem& p, const ps_result& res, int r0, int rk, ps_problem& pk, ps_result& rk_res) {
  const int n = p.n, Tc = std::min(p.T, n), nS = p.sweep ? Tc + 1 : 1;
  const long long off = (long long)r0 * n;
  pk = p;
  pk.batch = rk;
  if (p.perm) pk.perm = p.perm + off;
  rk_res = res;
  if (res.perm) rk_res.perm = res.perm + off;
  if (res.bounds) rk_res.bounds = res.bounds + (long long)r0 * nS * (Tc + 1);
  if (res.subset_scores) rk_res.subset_scores = (D*)res.subset_scores + (long long)r0 * nS * Tc;
  if (res.max_score) rk_res.max_score = (D*)res.max_score + (long long)r0 * (Tc + 1) * n;
  if (res.next_start) rk_res.next_start = res.next_start + (long long)r0 * (Tc + 1) * n;
}

// Large batches of independent instances from HOST buffers (randomization replicas): the batch is cut into parts; all
// host -> device copies are queued on a second stream at once, and each part is solved as soon as its inputs have arrived,
// so the copies of the later parts run under the kernels of the earlier ones.  Results are those of one call.  The overlap
// needs PINNED host buffers (cudaHostAlloc / cudaHostRegister): from pageable memory cudaMemcpyAsync stages synchronously
// and the copies simply run first (same results, no overlap).
template <typename D>
ps_status solve_pipelined(ps_ctx* ctx, const ps_problem& p, const D* a, const D* b, ps_result* res, int parts) {
  const int n = p.n, R = p.batch;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) PS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  const long long total = (long long)R * n;
  PS_TRY(ensure(ctx, ctx->a_in, total * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->b_in, total * sizeof(D)));
  D* d_a = (D*)ctx->a_in.p;
  D* d_b = (D*)ctx->b_in.p;
  const int per = cdiv(R, parts);
  std::vector<cudaEvent_t> arrived;
  cudaEvent_t e_c0 = nullptr, e_c1 = nullptr;  // copy-stream time of all parts -> ps_timings.h2d_ms
  struct EventGuard {  // every exit path destroys the events created here
    std::vector<cudaEvent_t>& v;
    cudaEvent_t &a, &b;
    ~EventGuard() {
      for (cudaEvent_t e : v) cudaEventDestroy(e);
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } event_guard{arrived, e_c0, e_c1};
  // buffers of an earlier call may still be read by kernels on the compute stream
  PS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  PS_CUDA(ctx, cudaEventCreate(&e_c0));
  PS_CUDA(ctx, cudaEventCreate(&e_c1));
  PS_CUDA(ctx, cudaEventRecord(e_c0, ctx->copy_stream));
  for (int r0 = 0; r0 < R; r0 += per) {
    const long long off = (long long)r0 * n, cnt = (long long)std::min(per, R - r0) * n;
    PS_CUDA(ctx, cudaMemcpyAsync(d_a + off, a + off, cnt * sizeof(D), cudaMemcpyHostToDevice, ctx->copy_stream));
    PS_CUDA(ctx, cudaMemcpyAsync(d_b + off, b + off, cnt * sizeof(D), cudaMemcpyHostToDevice, ctx->copy_stream));
    cudaEvent_t e;
    PS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    PS_CUDA(ctx, cudaEventRecord(e, ctx->copy_stream));
    arrived.push_back(e);
  }
  PS_CUDA(ctx, cudaEventRecord(e_c1, ctx->copy_stream));
  ps_timings agg;
  memset(&agg, 0, sizeof(agg));
  ps_status st = PS_OK;
  int part = 0;
  for (int r0 = 0; r0 < R && st == PS_OK; r0 += per, ++part) {
    const int rk = std::min(per, R - r0);
    const long long off = (long long)r0 * n;
    ps_problem pk;
    ps_result rk_res;
    slice_batch<D>(p, *res, r0, rk, pk, rk_res);
    cudaError_t ce = cudaStreamWaitEvent(ctx->stream, arrived[part], 0);
    if (ce != cudaSuccess) {
      ctx->err = std::string("cudaStreamWaitEvent: ") + cudaGetErrorString(ce);
      st = PS_ERR_CUDA;
      break;
    }
    st = solve_impl<D>(ctx, &pk, a + off, b + off, &rk_res, d_a + off, d_b + off);
    if (st != PS_OK) break;
    add_timings(agg, ctx->tm);
  }
  cudaStreamSynchronize(ctx->copy_stream);
  if (st != PS_OK) return st;
  cudaEventElapsedTime(&agg.h2d_ms, e_c0, e_c1);
  agg.h2d_bytes = 2 * total * (long long)sizeof(D) + (p.perm ? total * (long long)sizeof(int) : 0);
  ctx->tm = agg;
  return PS_OK;
}

template <typename D>
ps_status solve_impl(ps_ctx* ctx, const ps_problem* pp, const D* a, const D* b, ps_result* res, const D* staged_a,
                     const D* staged_b, const ps_comm* comm) {
  if (!ctx) return PS_ERR_INVALID;
  ctx->err.clear();
  if (!pp || !a || !b || !res) {
    ctx->err = "null argument";
    return PS_ERR_INVALID;
  }
  const ps_problem p = *pp;
  if (p.n < 1 || p.T < 1 || p.batch < 1) {
    ctx->err = "n, T and batch must be >= 1";
    return PS_ERR_INVALID;
  }
  if (p.batch > PS_MAX_BATCH_SLICE && !comm) {
    // the instance index is a grid dimension (<= 65535): larger batches run as consecutive slices, results as one call
    ps_timings agg;
    memset(&agg, 0, sizeof(agg));
    for (int r0 = 0; r0 < p.batch; r0 += PS_MAX_BATCH_SLICE) {
      const int rk = std::min(PS_MAX_BATCH_SLICE, p.batch - r0);
      ps_problem pk;
      ps_result rk_res;
      slice_batch<D>(p, *res, r0, rk, pk, rk_res);
      const long long off = (long long)r0 * p.n;
      PS_TRY(solve_impl<D>(ctx, &pk, a + off, b + off, &rk_res));
      add_timings(agg, ctx->tm);
    }
    ctx->tm = agg;
    return PS_OK;
  }
  if (p.dist < 0 || p.dist > 2) {
    ctx->err = "Bad distributional assignment";
    return PS_ERR_DIST;
  }
  if (p.sort_mode == PS_SORT_GIVEN && !p.perm) {
    ctx->err = "PS_SORT_GIVEN without perm";
    return PS_ERR_INVALID;
  }
  if (p.sum_mode != PS_SUM_RUNNING && p.sum_mode != PS_SUM_PREFIX) {
    ctx->err = "bad sum_mode";
    return PS_ERR_INVALID;
  }
  const int n = p.n, R = p.batch;
  if (comm && (R != 1 || comm->device != ctx->device)) {
    ctx->err = "sharded solve: one instance per call, communicator and context on the same device";
    return PS_ERR_INVALID;
  }
  if (!staged_a && !p.on_device && R >= 256 && (long long)R * n * sizeof(D) >= (16LL << 20))
    return solve_pipelined<D>(ctx, p, a, b, res, 4);
  // Row sharding (one very large instance over the GPUs of `comm`): rank r owns the 128-row blocks rb with rb % world == r
  // (equal rows AND equal cells of the triangle).  Everything up to the level records is replicated (sort, scan: O(n));
  // a level is computed for the owned rows only, then ONE ncclAllGather on this stream moves every rank's packed
  // {value, split} chunk to all ranks and a scatter kernel puts them back in row order -- no host synchronisation between
  // levels.  The backtrace is replicated: every rank ends with the full tables.
  const int rb_first = comm ? comm->rank : 0, rb_step = comm ? comm->world : 1;
  const int world = comm ? comm->world : 1;
  const int chunk_rows = cdiv(cdiv(n, BR), world) * BR;  // owned rows, padded: the same on every rank
  const int owned_blocks = cdiv(n, BR) > rb_first ? (cdiv(n, BR) - rb_first + rb_step - 1) / rb_step : 0;
  const size_t chunk_bytes = (size_t)chunk_rows * (sizeof(D) + sizeof(int));
  const int Tc = std::min(p.T, n);  // DP_impl.hpp:37-38
  int running = (p.sum_mode == PS_SUM_RUNNING);  // cleared below when the screened path takes over (exact sums)
  const int rp = p.risk_part ? 1 : 0;
  if (p.screen < 0 || p.screen > 2) {
    ctx->err = "bad screen option";
    return PS_ERR_INVALID;
  }
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  memset(&ctx->tm, 0, sizeof(ctx->tm));
  ctx->ev_used = 0;
  ctx->launches = 0;
  const long long total = (long long)R * n;
  const cudaMemcpyKind in_kind = p.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = p.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;

  cudaEvent_t e_begin, e_h2d, e_sort, e_pre, e_lev, e_bt, e_end;
  PS_TRY(mark(ctx, &e_begin));

  // ---- inputs
  const D* d_a = a;
  const D* d_b = b;
  if (staged_a) {  // inputs already on the device (a part of a pipelined batch)
    d_a = staged_a;
    d_b = staged_b;
  } else if (!p.on_device) {
    PS_TRY(ensure(ctx, ctx->a_in, total * sizeof(D)));
    PS_TRY(ensure(ctx, ctx->b_in, total * sizeof(D)));
    PS_CUDA(ctx, cudaMemcpyAsync(ctx->a_in.p, a, total * sizeof(D), in_kind, st));
    PS_CUDA(ctx, cudaMemcpyAsync(ctx->b_in.p, b, total * sizeof(D), in_kind, st));
    ctx->tm.h2d_bytes += 2 * total * sizeof(D);
    d_a = (const D*)ctx->a_in.p;
    d_b = (const D*)ctx->b_in.p;
  }
  PS_TRY(mark(ctx, &e_h2d));

  // ---- safe-range precheck for the branch-free math path (math_fast.cuh); result read before the levels
  int host_flags = 1;
  cudaEvent_t e_flags = nullptr;
  const bool want_fast = running && p.no_fast_math != 1;
  // screened levels (screen.cuh): worth it once a level has enough cells to amortise the extra launch
  bool want_screen = want_fast && p.screen != 1 && p.no_fast_math == 0 && std::min(p.T, n) >= 3 && n >= 512 &&
                     n <= 4000000 &&            // tile geometry of the live-tile kernels (<= 32 stages, <= 1000 chunks)
                     (long long)n * R >= 6000;  // measured crossover (tools/screen_threshold.py)
  if (want_screen && !mufu_claims_hold(ctx)) want_screen = false;  // hardware claims of the error budget, once per device
  if (want_fast) {
    PS_TRY(ensure(ctx, ctx->scalar, 64));
    PS_CUDA(ctx, cudaMemsetAsync(ctx->scalar.p, 0, 2 * sizeof(int), st));  // [0] range/exactness, [1] sortedness
    ps::range_check_kernel<D><<<cdiv(total, 256), 256, 0, st>>>(total, d_a, d_b, (int*)ctx->scalar.p);
    ctx->launches++;
    if (want_screen) {  // exact summability of every instance -> flag bit 2
      PS_TRY(ensure(ctx, ctx->scr_stats, (size_t)R * sizeof(ps::ScreenStats)));
      ps::ScreenStats* d_st = (ps::ScreenStats*)ctx->scr_stats.p;
      ps::screen_stats_init_kernel<<<cdiv(R, 256), 256, 0, st>>>(R, d_st);
      const int bx = std::max(1, std::min(cdiv(n, 256 * 4), std::max(1, 2048 / R)));
      ps::screen_stats_kernel<D><<<dim3(bx, R), 256, 0, st>>>(n, d_a, d_b, d_st);
      ps::screen_decide_kernel<D><<<cdiv(R, 256), 256, 0, st>>>(R, d_st, (int*)ctx->scalar.p, n);
      ctx->launches += 3;
      if (sizeof(D) == 8) {  // fixed-point shifts and conditions of the running-sum screen (screen_run.cuh) -> flag bit 3
        PS_TRY(ensure(ctx, ctx->fxinfo, (size_t)R * sizeof(ps::FxInfo)));
        ps::fx_setup_kernel<<<cdiv(R, 256), 256, 0, st>>>(R, n, (p.dist == PS_DIST_POISSON && rp) ? 1 : 0, d_st,
                                                          (ps::FxInfo*)ctx->fxinfo.p, (int*)ctx->scalar.p);
        ctx->launches++;
      }
    }
    if (!ctx->pinned_flags) PS_CUDA(ctx, cudaMallocHost((void**)&ctx->pinned_flags, 128));
    PS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_flags, ctx->scalar.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    PS_TRY(mark(ctx, &e_flags));
  }

  // ---- stage 1: sort by priority, gather, records
  PS_TRY(sort_stage<D>(ctx, p, d_a, d_b, R, n));
  const long long rec_stride = (long long)n + 1;
  PS_TRY(ensure(ctx, ctx->rec, 2 * R * rec_stride * sizeof(ps::Rec4<D>)));
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)ctx->rec.p;
  ps::Rec4<D>* rec1 = rec0 + R * rec_stride;
  ps::Rec4<D>* recs[2] = {rec0, rec1};
  const D* d_as = (const D*)ctx->a_s.p;
  const D* d_bs = (const D*)ctx->b_s.p;
  int fast = 0, screen = 0, screen_run = 0, screen_run32 = 0;
  const int cp_rows = n / ps::SCR_BLOCK + 1;  // float running-sum screen: one checkpoint row per 128-block boundary
  if (want_fast) {  // the precheck ran right after the inputs arrived; its flags are on the host by now
    if (ctx->capturing) {
      host_flags = 0;  // recorded for inputs that pass the precheck; every replay verifies the flags it produced
    } else {
      PS_CUDA(ctx, cudaEventSynchronize(e_flags));
      host_flags = *ctx->pinned_flags;
    }
    fast = ((host_flags & 1) == 0) && (p.dist != PS_DIST_POISSON || (host_flags & 2) == 0);
    // screening needs positive sums for every score (relative error of the float range sums) and exact summability
    screen = want_screen && fast && (host_flags & 2) == 0 && (host_flags & 4) == 0;
    // double inputs whose sums round: float bounds from fixed-point prefix sums, survivors from the reference's running sums
    // (measured crossover against the exhaustive kernel at T = 8: n ~ 14000 for double, ~ 10000 for float, tools/run_c3r.py threshold; audit runs take it at any size)
    screen_run = want_screen && fast && !screen && sizeof(D) == 8 && (host_flags & 8) == 0 &&
                 (n >= 14336 || p.screen == 2);
    // float inputs whose sums round, A^2/B scores: bounds from the running sums themselves (per-block checkpoints)
    screen_run32 = want_screen && fast && !screen && sizeof(D) == 4 && ScreenRun32Launch::supported(p.dist, rp) &&
                   (host_flags & 16) == 0 && (n >= 10240 || p.screen == 2) &&
                   (double)R * cp_rows * n * 8.0 <= 8e9;
  }
  if (screen) running = 0;  // exactly summable: prefix differences ARE the reference's running sums, bit for bit
  if (running) {
    ps::build_rec_running_kernel<D><<<dim3(cdiv(n + 1, 256), R), 256, 0, st>>>(n, rec_stride, d_as, d_bs,
                                                                              rec0, rec1);
  } else if (screen && n >= 4 * ps::PFX_BLOCK) {
    // exactly summable inputs: the order of additions does not matter, the scan spreads over the machine
    const int nb = cdiv(n, ps::PFX_BLOCK);
    PS_TRY(ensure(ctx, ctx->key_in, std::max<size_t>(ctx->key_in.cap, (size_t)2 * R * nb * sizeof(D))));
    D* d_pa = (D*)ctx->key_in.p;  // the sort is done with its key buffers
    D* d_pb = d_pa + (size_t)R * nb;
    ps::prefix_block_sums_kernel<D><<<dim3(nb, R), 256, 0, st>>>(n, nb, d_as, d_bs, d_pa, d_pb);
    ps::prefix_scan_partials_kernel<D><<<R, 32, 0, st>>>(nb, d_pa, d_pb);
    ps::prefix_write_kernel<D><<<dim3(nb, R), 256, 0, st>>>(n, nb, rec_stride, d_as, d_bs, d_pa, d_pb, rec0, rec1);
    ctx->launches += 2;
  } else {
    ps::build_rec_prefix_kernel<D, 1024><<<R, 1024, 0, st>>>(n, rec_stride, d_as, d_bs, rec0, rec1);
  }
  ctx->launches++;

  // ---- DP tables
  const long long table_stride = (long long)(Tc + 1) * n;
  PS_TRY(ensure(ctx, ctx->ms, R * table_stride * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->ns, R * table_stride * sizeof(int)));
  D* d_ms = (D*)ctx->ms.p;
  int* d_ns = (int*)ctx->ns.p;
  ps::init_tables_kernel<D><<<cdiv(R * table_stride, 256), 256, 0, st>>>(R * table_stride, n, d_ms, d_ns);
  ps::init_row0_kernel<D><<<dim3(cdiv(n, 256), R), 256, 0, st>>>(n, table_stride, d_ms);
  ctx->launches += 2;
  PS_TRY(mark(ctx, &e_sort));

  // ---- level 1
  // float Poisson on the fast path runs the two-rows-per-thread packed kernel (256-row tiles) once there are enough
  // rows to fill the machine with such tiles; smaller instances keep 128-row tiles (more, finer work items)
  const bool use_x2 = (sizeof(D) == 4) && fast && running && p.dist == PS_DIST_POISSON && p.no_fast_math != 2 && !comm &&
                      !screen_run32 &&  // the screened levels have their own 128-row geometry
                      (long long)n * R >= 16384;
  // small geometry (SMALL_ROWS above) while 128-row tiles would leave most SMs without work
  const long long nrb128 = cdiv(n, BR);
  const bool small = !comm && !screen && !screen_run && !screen_run32 && !use_x2 && fast && p.chunk_len <= 0 && n > SMALL_ROWS &&
                     (long long)R * nrb128 * (nrb128 + 1) / 2 < 4LL * ctx->sm_count;
  int L = small ? SMALL_ROWS : choose_chunk(n, R, p.chunk_len, use_x2 ? 2 * BR : BR);
  // screened levels: most tiles are dropped whole by the tile tier, so finer tiles pay (measured: 512 at n = 1e5);
  // at most ~200 chunks so that the per-chunk partials stay small
  if ((screen || screen_run || screen_run32) && p.chunk_len <= 0) L = std::min(L, std::max(512, cdiv(cdiv(n, 200), BR) * BR));
  // double running-sum screen: 256-cell chunks (shorter lazy walks from the chunk's checkpoint, lower-bound probes twice as
  // dense; measured at the C3 shape: 13.6 -> 13.1 ms, 512 and more are slower) while the partial arrays stay below ~1 GB
  if (screen_run && p.chunk_len <= 0 && (long long)n * R <= 200000) L = std::min(L, 2 * BR);
  if (screen || screen_run || screen_run32) {
    L = std::min(L, ps::SCR_MAX_G * BR);                   // the live-stage mask of a tile is one 32-bit word
    L = std::max(L, cdiv(cdiv(n, 1000), BR) * BR);         // <= 1000 chunks: per-chunk staging of the select kernel
  }
  const int nch_alloc = std::max(1, cdiv(std::max(1, n - 1), L));
  // live-tile kernels: partial slots per k-chunk.  One big instance: the tiles around a row block's maximum are split into
  // two work items (otherwise a single tile runs for most of a level; measured: 2 beats 1 and 4 at L = 512); batches have enough tiles to level the SMs
  int nsub = ((screen || screen_run || screen_run32) && R < 64 && p.screen != 2) ? std::min(2, L / BR) : 1;
  const long long pstride = (long long)nch_alloc * nsub * n;
  PS_TRY(ensure(ctx, ctx->part_val, R * pstride * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->part_arg, R * pstride * sizeof(int)));
  D* d_pv = (D*)ctx->part_val.p;
  int* d_pa = (int*)ctx->part_arg.p;
  D* d_SA = nullptr;
  D* d_SB = nullptr;
  float4* recf[2] = {nullptr, nullptr};
  // sortedness flag, row lower bounds, block maxima, counters and live-tile lists of the screened level kernels
  auto screen_buffers = [&]() -> ps_status {
    ps::screen_sorted_kernel<D><<<dim3(cdiv(n, 256), R), 256, 0, st>>>(n, d_as, d_bs, (int*)ctx->scalar.p + 1);
    ctx->launches++;
    PS_TRY(ensure(ctx, ctx->lb, (size_t)R * n * sizeof(D)));
    PS_TRY(ensure(ctx, ctx->scr_bmax, (size_t)R * (n / ps::SCR_BLOCK + 2) * sizeof(float)));
    PS_TRY(ensure(ctx, ctx->scr_counters, 8 * sizeof(unsigned long long)));
    PS_CUDA(ctx, cudaMemsetAsync(ctx->scr_counters.p, 0, 8 * sizeof(unsigned long long), st));
    // live-tile protocol: the compacted tile list (every tile could be live), one segment record per row block, one header
    // (tiles appended, cursor) per level
    PS_TRY(ensure(ctx, ctx->live, (size_t)ps::SCR_CLASSES * R * nch_alloc * nsub * cdiv(n, BR) * sizeof(int4)));
    PS_TRY(ensure(ctx, ctx->rb_info, (size_t)R * nch_alloc * nsub * cdiv(n, BR) * sizeof(unsigned int)));  // per-slot stage masks
    PS_TRY(ensure(ctx, ctx->live_hdr, (size_t)(Tc + 2) * sizeof(ps::LiveHdr)));
    PS_TRY(ensure(ctx, ctx->kbest, (size_t)R * n * sizeof(int)));
    PS_TRY(ensure(ctx, ctx->rb_slots, (size_t)R * nch_alloc * nsub * cdiv(n, BR) * sizeof(int)));
    PS_TRY(ensure(ctx, ctx->rb_cnt, (size_t)R * cdiv(n, BR) * sizeof(int)));
    PS_CUDA(ctx, cudaMemsetAsync(ctx->live_hdr.p, 0, (size_t)(Tc + 2) * sizeof(ps::LiveHdr), st));
    if (ctx->dbg_level) {
      PS_TRY(ensure(ctx, ctx->dbg, (size_t)R * nch_alloc * nsub * cdiv(n, BR) * 32));
      PS_CUDA(ctx, cudaMemsetAsync(ctx->dbg.p, 0, (size_t)R * nch_alloc * nsub * cdiv(n, BR) * 32, st));
    }
    return PS_OK;
  };
  // checkpoints of the running sums: at the k-chunk boundaries, or (float running-sum screen) at every 128-block boundary
  // row-sharded solves: packed send chunk {D value[chunk_rows]; int split[chunk_rows]} and the gathered [world] chunks
  D* send_val = nullptr;
  int* send_arg = nullptr;
  unsigned char* recv_all = nullptr;
  auto shard_buffers = [&]() -> ps_status {
    if (send_val) return PS_OK;
    PS_TRY(ensure(ctx, ctx->shard_send, chunk_bytes));
    PS_TRY(ensure(ctx, ctx->shard_recv, chunk_bytes * (size_t)world));
    send_val = (D*)ctx->shard_send.p;
    send_arg = (int*)((unsigned char*)ctx->shard_send.p + (size_t)chunk_rows * sizeof(D));
    recv_all = (unsigned char*)ctx->shard_recv.p;
    PS_CUDA(ctx, cudaMemsetAsync(ctx->shard_send.p, 0, chunk_bytes, st));
    return PS_OK;
  };
  // all-gather the level's chunks and scatter them to row order: maxScore_[lev], nextStart_[lev], next records' p / pm
  auto exchange = [&](int lev, int nrows_lev, ps::Rec4<D>* rec_next, float* pm_next) -> ps_status {
    NcclApi& nc = nccl_api();
    PS_NCCL(ctx, nc.AllGather(ctx->shard_send.p, recv_all, chunk_bytes, ncclInt8, comm->comm, st));
    ps::shard_scatter_kernel<D, BR><<<cdiv((long long)world * chunk_rows, 256), 256, 0, st>>>(
        n, nrows_lev, world, chunk_rows, recv_all, (D*)ctx->ms.p + (long long)lev * n, (int*)ctx->ns.p + (long long)lev * n,
        rec_next, pm_next);
    ctx->launches++;
    ctx->tm.comm_bytes += (long long)chunk_bytes * world;
    return PS_OK;
  };
  const int sa_L = screen_run32 ? ps::SCR_BLOCK : L;
  const int sa_nch = screen_run32 ? cp_rows : nch_alloc;
  if (running) {
    if (sa_nch > 1) {
      DevBuf& bufA = screen_run32 ? ctx->cpA : ctx->SA;
      DevBuf& bufB = screen_run32 ? ctx->cpB : ctx->SB;
      PS_TRY(ensure(ctx, bufA, (size_t)R * sa_nch * n * sizeof(D)));
      PS_TRY(ensure(ctx, bufB, (size_t)R * sa_nch * n * sizeof(D)));
      d_SA = (D*)bufA.p;
      d_SB = (D*)bufB.p;
    }
    if (small) {
      static_assert(SMALL_ROWS == 32, "level1_running_small_kernel: one warp per row block, chunk length 32");
      ps::level1_running_small_kernel<D><<<dim3((cdiv(n, SMALL_ROWS) + 1) / 2, R), SMALL_ROWS, 0, st>>>(
          n, sa_nch, rec_stride, p.dist, rp, rec0, d_SA, d_SB, d_ms + n, d_ns + n, table_stride, rec1,
          cdiv(n, SMALL_ROWS));
    } else if (!comm) {
      ps::level1_running_kernel<D, BR><<<dim3((cdiv(n, BR) + 1) / 2, R), BR, 0, st>>>(
          n, sa_L, sa_nch, rec_stride, p.dist, rp, rec0, d_SA, d_SB, d_ms + n, d_ns + n, table_stride, rec1,
          0, 1, nullptr, cdiv(n, BR));
    } else {
      // owned rows only (the O(n^2) walk is what is worth sharding); values travel like every other level's, splits = n
      PS_TRY(shard_buffers());
      if (owned_blocks > 0)
        ps::level1_running_kernel<D, BR><<<dim3((owned_blocks + 1) / 2, 1), BR, 0, st>>>(
            n, sa_L, sa_nch, rec_stride, p.dist, rp, rec0, d_SA, d_SB, (D*)nullptr, (int*)nullptr, table_stride,
            (ps::Rec4<D>*)nullptr, rb_first, rb_step, send_val, owned_blocks);
      ps::fill_int_kernel<<<cdiv(chunk_rows, 256), 256, 0, st>>>(chunk_rows, n, send_arg);
      ctx->launches++;
      PS_TRY(exchange(1, n, rec1, nullptr));
    }
    if constexpr (sizeof(D) == 4) {
      if (screen_run32) {
        PS_TRY(ensure(ctx, ctx->ipos, (size_t)R * sizeof(int)));
        PS_CUDA(ctx, cudaMemsetAsync(ctx->ipos.p, 0, (size_t)R * sizeof(int), st));
        ps::ipos_kernel<<<dim3(cdiv(n, 256), R), 256, 0, st>>>(n, d_as, (int*)ctx->ipos.p);
        ctx->launches++;
        PS_TRY(screen_buffers());
      }
    }
    if constexpr (sizeof(D) == 8) {
      if (screen_run) {
        PS_TRY(ensure(ctx, ctx->recf, 2 * R * rec_stride * sizeof(float4)));
        PS_TRY(ensure(ctx, ctx->pfx, R * rec_stride * sizeof(longlong2)));
        recf[0] = (float4*)ctx->recf.p;
        recf[1] = recf[0] + R * rec_stride;
        ps::fx_prefix_kernel<1024><<<R, 1024, 0, st>>>(n, rec_stride, d_as, d_bs, (ps::FxInfo*)ctx->fxinfo.p,
                                                      (longlong2*)ctx->pfx.p);
        ps::build_recf_fx_kernel<<<dim3(cdiv(n + 1, 256), R), 256, 0, st>>>(
            n, rec_stride, (const longlong2*)ctx->pfx.p, (const ps::FxInfo*)ctx->fxinfo.p, recf[0], recf[1]);
        ps::pm_from_rec_kernel<<<dim3(cdiv(n + 1, 256), R), 256, 0, st>>>(n, rec_stride, rec1, recf[1]);
        ctx->launches += 3;
        PS_TRY(screen_buffers());
      }
    }
  } else {
    // screened levels read float records: double -> separate {lx, ly, -, pm} arrays, float -> the records themselves
    if (screen) {
      if constexpr (sizeof(D) == 8) {
        PS_TRY(ensure(ctx, ctx->recf, 2 * R * rec_stride * sizeof(float4)));
        recf[0] = (float4*)ctx->recf.p;
        recf[1] = recf[0] + R * rec_stride;
        ps::build_recf_kernel<<<dim3(cdiv(n + 1, 256), R), 256, 0, st>>>(n, rec_stride, rec0, recf[0], recf[1]);
        ctx->launches++;
      } else {
        recf[0] = (float4*)rec0;
        recf[1] = (float4*)rec1;
      }
      PS_TRY(screen_buffers());
    }
    ps::level1_prefix_kernel<D><<<dim3(cdiv(n, 256), R), 256, 0, st>>>(n, rec_stride, p.dist, rp, rec0,
                                                                      d_ms + n, d_ns + n, table_stride, rec1,
                                                                      (screen && sizeof(D) == 8) ? (float*)recf[1] : nullptr);
  }
  ctx->launches++;
  PS_TRY(mark(ctx, &e_pre));

  // ---- levels 2..Tc  (DP_impl.hpp:101-118)
  ctx->last_fast = fast;
  if (comm) ctx->map_key[0] = -1;  // the tile map depends on the ownership pattern
  auto fn = LevelLaunch<D>::pick(p.dist, rp, running, fast);
  // float Poisson on the fast path: two rows per thread, packed f32x2 pipeline, 256-row tiles
  int tile_rows = BR;
  int tile_threads = BR;
  if (small) {
    tile_rows = tile_threads = SMALL_ROWS;
    fn = LevelLaunch<D>::pick_small(p.dist, rp, running);
  }
  if constexpr (sizeof(D) == 4) {
    if (use_x2) {
      tile_rows = 2 * BR;
      fn = rp ? ps::level_tile_kernel_pois_f32x2<true> : ps::level_tile_kernel_pois_f32x2<false>;
    }
  }
  auto fn_last = LastLevelLaunch<D>::pick(p.dist, rp, running, fast);
  auto fn_scr = ScreenLaunch<D>::pick(p.dist, rp, p.screen == 2);
  // select kernel: four threads per row once a row block has enough tiles to share out
  const int sel_tpr = (nch_alloc >= 32) ? 4 : 1;
  auto fn_sel = ScreenLaunch<D>::pick_select(p.dist, rp, p.screen == 2, sel_tpr);
  auto fn_lb = ScreenLaunch<D>::pick_lb(p.dist, rp);
  D* d_lb = (D*)ctx->lb.p;
  unsigned long long* d_cnt = (unsigned long long*)ctx->scr_counters.p;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> lev_ev;
  // Exact branch-and-bound has a worst case: on inputs whose rows have no pronounced maximum (constant a/b: every cell of a
  // row ties) the bounds dismiss nothing and a screened level costs more than an exhaustive one.  The count of cells evaluated
  // with the reference's arithmetic is read back per level (a few bytes, no stall: the host runs two levels ahead of the
  // device, except for one wait after the first screened level); once a level evaluates more than SCREEN_GIVE_UP of its
  // cells exactly, the remaining levels run the exhaustive kernel on the same (prefix) records -- same tables either way.
  constexpr double SCREEN_GIVE_UP = 0.30;
  bool gave_up = false;
  std::vector<cudaEvent_t> cnt_ev(Tc + 1, nullptr);
  unsigned long long* cnt_pin = nullptr;
  if (screen && p.screen == 0) {
    if ((size_t)(Tc + 2) > ctx->pinned_levels_cap) {
      if (ctx->pinned_levels) cudaFreeHost(ctx->pinned_levels);
      ctx->pinned_levels = nullptr;
      ctx->pinned_levels_cap = 0;
      PS_CUDA(ctx, cudaMallocHost((void**)&ctx->pinned_levels, (size_t)(Tc + 2 + 64) * sizeof(unsigned long long)));
      ctx->pinned_levels_cap = (size_t)Tc + 2 + 64;
    }
    cnt_pin = ctx->pinned_levels;
    cnt_pin[1] = 0;
  }
  auto exact_share = [&](int lev) {  // share of level lev's cells evaluated exactly (its counter copy has completed)
    const unsigned long long d = cnt_pin[lev] - cnt_pin[lev - 1];
    return 32.0 * (double)d / (double)(level_cells(n, lev, Tc) * R);
  };
  for (int j = 2; j <= Tc; ++j) {
    if (cnt_pin && !gave_up && j < Tc) {
      const int look = (j == 3) ? 2 : j - 2;  // one blocking look at level 2, then two levels behind the launches
      if (look >= 2 && cnt_ev[look]) {
        PS_CUDA(ctx, cudaEventSynchronize(cnt_ev[look]));
        if (exact_share(look) > SCREEN_GIVE_UP) {
          gave_up = true;
          ctx->tm.screen_fallback_level = j;
        }
      }
    }
    const bool screened_level = screen && !gave_up;
    ps::Geom g = make_geom(n, j, Tc, L, nch_alloc, rb_first, rb_step, tile_rows);
    const ps::Rec4<D>* rin = recs[(j - 1) & 1];
    ps::Rec4<D>* rout = recs[j & 1];
    const bool timed = (j - 2) < PS_MAX_TIMED_LEVELS;
    bool fused_fold = false;  // the level kernel wrote maxScore_ / nextStart_ itself (live-tile kernels)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) PS_TRY(mark(ctx, &e0));
    if (j == Tc && Tc >= 2) {
      // last level: row 0 only -> k-parallel kernel, one CTA per chunk (128-row geometry)
      // sharded: row 0 (and its running-sum checkpoints) belongs to rank 0; its value travels like every other level's
      g = make_geom(n, j, Tc, L, nch_alloc, 0, 1, small ? SMALL_ROWS : BR);
      g.sa_mul = L / sa_L;
      g.sa_nch = sa_nch;
      if (!comm || rb_first == 0) fn_last<<<dim3(g.NCH, R), 128, 0, st>>>(g, rin, d_SA, d_SB, d_pv, d_pa);
      if (comm) {
        const int nch = g.NCH;
        g = make_geom(n, j, Tc, L, nch_alloc, rb_first, rb_step, BR);  // owned blocks of the one-row level: {0} on rank 0
        g.NCH = nch;
      }
    } else {
      if (!(screened_level || screen_run || screen_run32) && (long long)g.ntiles * R >= 2048 && !ctx->capturing) {
        // tile -> (row block, chunk) table (one per geometry, shared by the instances of a batch), rebuilt only when the
        // geometry changes (kmax shrinks by one per level)
        if (ctx->map_key[0] != g.NCH || ctx->map_key[1] != g.G || ctx->map_key[2] != g.NRB ||
            ctx->map_key[3] != g.ntiles) {
          PS_TRY(ensure(ctx, ctx->tile_map, (size_t)g.ntiles * sizeof(int2)));
          ps::build_tile_map_kernel<<<cdiv(g.ntiles, 128), 128, 0, st>>>(g, (int2*)ctx->tile_map.p);
          ctx->launches++;
          ctx->map_key[0] = g.NCH, ctx->map_key[1] = g.G, ctx->map_key[2] = g.NRB, ctx->map_key[3] = g.ntiles;
        }
        g.tile_map = (const int2*)ctx->tile_map.p;
      }
      if (screen_run32 || screen_run || screened_level) {
        // Screened level (screen.cuh, "four launches"): row lower bounds from a few exactly evaluated cells; range tiers for
        // every tile -> compacted, cost-ordered lists of live tiles; persistent CTAs work the lists off; fold over the slots
        // that hold a partial (no combine launch).  The three families differ in where a bound's range sums come from.
        ps::LiveHdr* hdr = (ps::LiveHdr*)ctx->live_hdr.p + j;
        const long long live_cap = (long long)R * nch_alloc * nsub * cdiv(n, BR);
        const int nblk = n / ps::SCR_BLOCK + 2;
        const int audit = p.screen == 2;
        g.nsub = nsub;
        ps::SelSrc src;
        memset(&src, 0, sizeof(src));
        typename ScreenLaunch<D>::SelFn fn_select = fn_sel;
        if (screen_run32) {
          if constexpr (sizeof(D) == 4) {
            const long long cp_stride = (long long)cp_rows * n;
            ScreenRun32Launch::pick_lb(p.dist, rp)<<<dim3(cdiv(g.nrows, 128), R), 128, 0, st>>>(
                g, rin, (float4*)rin, d_SA, d_SB, cp_stride, d_lb, (float*)ctx->scr_bmax.p, nblk, (int*)ctx->kbest.p);
            src.cpA = d_SA;
            src.cpB = d_SB;
            src.cp_stride = cp_stride;
            src.ipos = (const int*)ctx->ipos.p;
            fn_select = ScreenRun32Launch::pick_select(p.dist, rp, audit, sel_tpr);
          }
        } else if (screen_run) {
          if constexpr (sizeof(D) == 8) {
            ScreenRunLaunch::pick_lb(p.dist, rp)<<<dim3(cdiv(g.nrows, 128), R), 128, 0, st>>>(
                g, rin, recf[(j - 1) & 1], d_SA, d_SB, d_lb, (float*)ctx->scr_bmax.p, nblk, (int*)ctx->kbest.p);
            src.pfx = (const longlong2*)ctx->pfx.p;
            src.fx = (const ps::FxInfo*)ctx->fxinfo.p;
            fn_select = ScreenRunLaunch::pick_select(p.dist, rp, audit, sel_tpr);
          }
        } else {
          fn_lb<<<dim3(cdiv(g.nrows, 128), R), 128, 0, st>>>(g, rin, recf[(j - 1) & 1], d_lb, (float*)ctx->scr_bmax.p, nblk,
                                                             (int*)ctx->kbest.p);
        }
        if (g.NRB > 0)
          fn_select<<<dim3(g.NRB, R), 128 * sel_tpr, (size_t)g.NCH * (4 * 8 + 12 + nsub), st>>>(
              g, rin, d_lb, (const float*)ctx->scr_bmax.p, nblk, (const int*)ctx->scalar.p, (int4*)ctx->live.p, live_cap,
              (unsigned int*)ctx->rb_info.p, cdiv(n, BR), hdr, d_cnt, (const int*)ctx->kbest.p, (int*)ctx->rb_slots.p,
              (int*)ctx->rb_cnt.p, src);
        ps::LevelOut<D> lo;
        lo.ms_row = d_ms + (long long)j * n;
        lo.ns_row = d_ns + (long long)j * n;
        lo.table_stride = table_stride;
        lo.rec_next = rout;
        lo.pm_next = ((screen || screen_run) && sizeof(D) == 8) ? (float*)recf[j & 1] : nullptr;
        lo.shard_chunk = nullptr;
        lo.shard_arg = nullptr;
        if (comm) {  // owned rows go to the send chunk; the scatter after the all-gather writes the tables and records
          PS_TRY(shard_buffers());
          lo.ms_row = nullptr;
          lo.ns_row = nullptr;
          lo.rec_next = nullptr;
          lo.pm_next = nullptr;
          lo.shard_chunk = send_val;
          lo.shard_arg = send_arg;
        }
        // persistent CTAs, 8 per SM (64 registers): they pull tiles until the lists are empty
        const int grid_live = (int)std::min<long long>((long long)g.ntiles * R * nsub, (long long)ctx->sm_count * 8);
        if (grid_live > 0) {
          if (screen_run32) {
            if constexpr (sizeof(D) == 4)
              ScreenRun32Launch::pick(p.dist, rp, audit)<<<grid_live, BR, 0, st>>>(
                  g, rin, d_SA, d_SB, (long long)cp_rows * n, (const int*)ctx->ipos.p, d_lb, d_pv, d_pa, d_cnt,
                  (const int*)ctx->scalar.p, (const int4*)ctx->live.p, live_cap, hdr);
          } else if (screen_run) {
            if constexpr (sizeof(D) == 8)
              ScreenRunLaunch::pick(p.dist, rp, audit)<<<grid_live, BR, 0, st>>>(
                  g, rin, recf[(j - 1) & 1], (const longlong2*)ctx->pfx.p, (const ps::FxInfo*)ctx->fxinfo.p, d_SA, d_SB, d_lb,
                  d_pv, d_pa, d_cnt, (const int*)ctx->scalar.p, (const int4*)ctx->live.p, live_cap, hdr);
          } else {
            fn_scr<<<grid_live, BR, 0, st>>>(g, rin, recf[(j - 1) & 1], d_lb, d_pv, d_pa, d_cnt, (const int*)ctx->scalar.p,
                                             (const int4*)ctx->live.p, live_cap, hdr,
                                             (ctx->dbg_level == j) ? (unsigned long long*)ctx->dbg.p : nullptr);
          }
        }
        if (g.NRB > 0) {
          if (g.NCH >= 32)  // four lanes per row once a row has enough chunks to fold
            ps::screen_fold_kernel<D, BR, 4><<<dim3(g.NRB, R), BR * 4, 0, st>>>(
                g, (const int*)ctx->rb_slots.p, (const int*)ctx->rb_cnt.p, cdiv(n, BR), d_pv, d_pa, lo);
          else
            ps::screen_fold_kernel<D, BR, 1><<<dim3(g.NRB, R), BR, 0, st>>>(
                g, (const int*)ctx->rb_slots.p, (const int*)ctx->rb_cnt.p, cdiv(n, BR), d_pv, d_pa, lo);
        }
        ctx->launches += 3;
        fused_fold = true;
        if (cnt_pin && screened_level) {  // cumulative count of exactly evaluated warp-cells after this level -> pinned host slot
          PS_CUDA(ctx, cudaMemcpyAsync(&cnt_pin[j], d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
          PS_TRY(mark(ctx, &cnt_ev[j]));
        }
      } else if (g.ntiles > 0) {
        fn<<<dim3(g.ntiles, R), tile_threads, 0, st>>>(g, rin, d_SA, d_SB, d_pv, d_pa);
      }
    }
    const bool sharded_level = comm != nullptr;
    if (!fused_fold && g.NRB > 0) {
      const bool wide_combine = g.NCH >= 32;  // four lanes per row once a row has enough chunks to fold
      auto fn_combine = wide_combine ? ps::combine_kernel<D, BR, 4> : ps::combine_kernel<D, BR, 1>;
      float* pm_out = ((screen || screen_run) && sizeof(D) == 8) ? (float*)recf[j & 1] : nullptr;
      if (sharded_level) {
        PS_TRY(shard_buffers());
        fn_combine<<<dim3(cdiv((wide_combine ? 4LL : 1LL) * g.NRB * tile_rows, 256), R), 256, 0, st>>>(
            g, d_pv, d_pa, (D*)nullptr, (int*)nullptr, table_stride, (ps::Rec4<D>*)nullptr, send_val, (float*)nullptr,
            (const int*)nullptr, 0, 0, send_arg);
      } else {
        fn_combine<<<dim3(cdiv((wide_combine ? 4LL : 1LL) * g.NRB * tile_rows, 256), R), 256, 0, st>>>(
            g, d_pv, d_pa, d_ms + (long long)j * n, d_ns + (long long)j * n, table_stride, rout, nullptr, pm_out,
            (const int*)nullptr, 0, 0, (int*)nullptr);
      }
      ctx->launches++;
    }
    if (sharded_level)
      PS_TRY(exchange(j, g.nrows, rout, ((screen || screen_run) && sizeof(D) == 8) ? (float*)recf[j & 1] : nullptr));
    if (timed) PS_TRY(mark(ctx, &e1));
    ctx->launches++;
    if (timed) {
      lev_ev.push_back({e0, e1});
      ctx->tm.level_cells[j - 2] = level_cells(n, j, Tc) * R;
    }
  }
  PS_CUDA(ctx, cudaGetLastError());
  PS_TRY(mark(ctx, &e_lev));

  // ---- backtrace
  const int nS = p.sweep ? Tc + 1 : 1;
  PS_TRY(ensure(ctx, ctx->bounds, (size_t)R * nS * (Tc + 1) * sizeof(int)));
  PS_TRY(ensure(ctx, ctx->scores, (size_t)R * nS * Tc * sizeof(D)));
  int* d_bounds = (int*)ctx->bounds.p;
  D* d_scores = (D*)ctx->scores.p;
  ps::backtrace_bounds_kernel<<<cdiv((long long)R * nS, 128), 128, 0, st>>>(R, n, Tc, p.sweep ? 1 : 0, d_ns,
                                                                           d_bounds);
  ps::backtrace_scores_kernel<D><<<cdiv((long long)R * nS * Tc, 128), 128, 0, st>>>(
      R, n, Tc, p.sweep ? 1 : 0, p.dist, rp, running, d_bounds, d_as, d_bs, rec0, rec_stride,
      comm ? (const D*)nullptr : d_SA, comm ? (const D*)nullptr : d_SB,  // sharded: checkpoints exist for owned rows only
      sa_L, sa_nch, d_scores);
  if (comm) ctx->map_key[0] = -1;
  ctx->launches += 2;
  PS_CUDA(ctx, cudaGetLastError());
  PS_TRY(mark(ctx, &e_bt));

  // ---- outputs
  // Small results for host callers travel through one pinned staging buffer: a device-to-pageable copy blocks the host for
  // ~12 us each (three of them were 39 of the 245 us of a C1 solve), copies into pinned memory queue behind the kernels and
  // cost one synchronisation together; the host then moves the bytes to the caller's arrays.
  std::vector<StagedOut> staged;
  size_t staged_bytes = 0;
  {
    auto want = [&](const void* dst, size_t bytes) { return dst ? ((bytes + 15) & ~(size_t)15) : 0; };
    const size_t need = want(res->perm, total * sizeof(int)) + want(res->bounds, (size_t)R * nS * (Tc + 1) * sizeof(int)) +
                        want(res->subset_scores, (size_t)R * nS * Tc * sizeof(D)) +
                        want(res->max_score, R * table_stride * sizeof(D)) +
                        want(res->next_start, R * table_stride * sizeof(int));
    if (!p.on_device && need > 0 && need <= PINNED_OUT_BYTES) {
      if (!ctx->pinned_out) PS_CUDA(ctx, cudaMallocHost((void**)&ctx->pinned_out, PINNED_OUT_BYTES));
      staged.reserve(5);
    }
  }
  const bool stage_out = staged.capacity() > 0;
  if (ctx->capturing && !stage_out) {
    ctx->err = "graph capture: results do not fit the staging buffer";
    return PS_ERR_CUDA;
  }
  auto copy_out = [&](int field, void* dst, const void* src, size_t bytes) -> ps_status {
    if (!dst) return PS_OK;
    if (stage_out) {
      PS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_out + staged_bytes, src, bytes, cudaMemcpyDeviceToHost, st));
      staged.push_back({field, staged_bytes, bytes});
      staged_bytes += (bytes + 15) & ~(size_t)15;
    } else {
      PS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, out_kind, st));
    }
    if (!p.on_device) ctx->tm.d2h_bytes += bytes;
    return PS_OK;
  };
  PS_TRY(copy_out(0, res->perm, ctx->perm.p, total * sizeof(int)));
  PS_TRY(copy_out(1, res->bounds, d_bounds, (size_t)R * nS * (Tc + 1) * sizeof(int)));
  PS_TRY(copy_out(2, res->subset_scores, d_scores, (size_t)R * nS * Tc * sizeof(D)));
  PS_TRY(copy_out(3, res->max_score, d_ms, R * table_stride * sizeof(D)));
  PS_TRY(copy_out(4, res->next_start, d_ns, R * table_stride * sizeof(int)));
  if (screen || screen_run || screen_run32)
    PS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_flags + 2, d_cnt, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  PS_TRY(mark(ctx, &e_end));
  if (ctx->capturing) {  // recorded, not run: the replay copies the results out and fills the timings
    ctx->cap_staged = staged;
    ctx->tm.kernel_launches = ctx->launches;
    return PS_OK;
  }
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  scatter_staged(ctx, staged, res);
  if (ctx->dbg_level && screen && ctx->dbg.p) {
    const size_t cnt = (size_t)R * nch_alloc * nsub * cdiv(n, BR) * 4;
    std::vector<unsigned long long> h(cnt);
    cudaMemcpy(h.data(), ctx->dbg.p, cnt * 8, cudaMemcpyDeviceToHost);
    const char* path = getenv("PS_DEBUG_FILE");
    if (FILE* f = fopen(path ? path : "tile_times.bin", "wb")) {
      fwrite(h.data(), 8, cnt, f);
      fclose(f);
    }
  }
  ctx->tm.screen_used = screen ? 1 : ((screen_run || screen_run32) ? 2 : 0);
  if (screen || screen_run || screen_run32) {
    unsigned long long cnt[8];
    memcpy(cnt, ctx->pinned_flags + 2, sizeof(cnt));
    ctx->tm.screen_exact_cells = (long long)cnt[0] * 32;
    ctx->tm.screen_violations = (long long)cnt[1];
    for (int q = 0; q < 6; ++q) ctx->tm.screen_tiers[q] = (long long)cnt[2 + q];
  }

  auto ms_between = [&](cudaEvent_t x, cudaEvent_t y) {
    float v = 0.f;
    cudaEventElapsedTime(&v, x, y);
    return v;
  };
  ctx->tm.total_ms = ms_between(e_begin, e_end);
  ctx->tm.h2d_ms = ms_between(e_begin, e_h2d);
  ctx->tm.sort_scan_ms = ms_between(e_h2d, e_sort);
  ctx->tm.prepass_ms = ms_between(e_sort, e_pre);
  ctx->tm.levels_ms = ms_between(e_pre, e_lev);
  ctx->tm.backtrace_ms = ms_between(e_lev, e_bt);
  ctx->tm.d2h_ms = ms_between(e_bt, e_end);
  ctx->tm.n_levels = (int)lev_ev.size();
  for (size_t q = 0; q < lev_ev.size(); ++q)
    ctx->tm.level_ms[q] = ms_between(lev_ev[q].first, lev_ev[q].second);
  ctx->tm.kernel_launches = ctx->launches;
  return PS_OK;
}

template <typename D>
ps_status range_score_impl(ps_ctx* ctx, int n, const D* a, const D* b, int dist, int rp, D* out) {
  if (!ctx || !a || !b || !out || n < 1) return PS_ERR_INVALID;
  if (dist < 0 || dist > 2) return PS_ERR_DIST;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  PS_TRY(ensure(ctx, ctx->a_in, (size_t)n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->b_in, (size_t)n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->scalar, 64));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->a_in.p, a, (size_t)n * sizeof(D), cudaMemcpyHostToDevice, st));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->b_in.p, b, (size_t)n * sizeof(D), cudaMemcpyHostToDevice, st));
  ps::range_score_kernel<D><<<1, 32, 0, st>>>(n, dist, rp, (const D*)ctx->a_in.p, (const D*)ctx->b_in.p,
                                             (D*)ctx->scalar.p);
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaMemcpyAsync(out, ctx->scalar.p, sizeof(D), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  return PS_OK;
}

ps_status ltss_impl(ps_ctx* ctx, int n, const float* a, const float* b, int dist, int* perm_out, int* lo, int* hi,
                    float* score) {
  using D = float;
  if (!ctx) return PS_ERR_INVALID;
  ctx->err.clear();
  if (!a || !b || !lo || !hi || !score || n < 1) {
    ctx->err = "ps_ltss_f32: bad argument";
    return PS_ERR_INVALID;
  }
  if (dist < 0 || dist > 2) {
    ctx->err = "Bad distributional assignment";
    return PS_ERR_DIST;
  }
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  ctx->tm.h2d_bytes = 0;
  PS_TRY(ensure(ctx, ctx->a_in, (size_t)n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->b_in, (size_t)n * sizeof(D)));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->a_in.p, a, (size_t)n * sizeof(D), cudaMemcpyHostToDevice, st));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->b_in.p, b, (size_t)n * sizeof(D), cudaMemcpyHostToDevice, st));
  ps_problem p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  p.sort_mode = PS_SORT_DEVICE;  // stable, like LTSS's std::stable_sort (LTSS_impl.hpp:19)
  PS_TRY(sort_stage<D>(ctx, p, (const D*)ctx->a_in.p, (const D*)ctx->b_in.p, 1, n));
  const long long rec_stride = (long long)n + 1;
  PS_TRY(ensure(ctx, ctx->rec, 2 * rec_stride * sizeof(ps::Rec4<D>)));
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)ctx->rec.p;
  ps::Rec4<D>* rec1 = rec0 + rec_stride;
  const D* d_as = (const D*)ctx->a_s.p;
  const D* d_bs = (const D*)ctx->b_s.p;
  ps::build_rec_running_kernel<D><<<dim3(cdiv(n + 1, 256), 1), 256, 0, st>>>(n, rec_stride, d_as, d_bs, rec0, rec1);
  // suffix scores S(i,n) under the mult-clust objective == level 1 of the DP
  PS_TRY(ensure(ctx, ctx->ms, (size_t)n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->ns, (size_t)n * sizeof(int)));
  ps::level1_running_kernel<D, BR><<<dim3((cdiv(n, BR) + 1) / 2, 1), BR, 0, st>>>(
      n, n + BR, 1, rec_stride, dist, 0, rec0, (D*)nullptr, (D*)nullptr, (D*)ctx->ms.p, (int*)ctx->ns.p, 0,
      (ps::Rec4<D>*)nullptr, 0, 1, (D*)nullptr, cdiv(n, BR));
  PS_TRY(ensure(ctx, ctx->part_val, 2 * (size_t)n * sizeof(D)));
  D* PA = (D*)ctx->part_val.p;
  D* PB = PA + n;
  ps::ltss_prefix_sums_kernel<D><<<1, 32, 0, st>>>(n, d_as, d_bs, PA, PB);
  PS_TRY(ensure(ctx, ctx->scalar, 64));
  PS_CUDA(ctx, cudaMemsetAsync(ctx->scalar.p, 0, sizeof(unsigned long long), st));
  ps::ltss_argmax_kernel<D><<<cdiv(n, 256), 256, 0, st>>>(n, dist, PA, PB, (const D*)ctx->ms.p,
                                                          (unsigned long long*)ctx->scalar.p);
  PS_CUDA(ctx, cudaGetLastError());
  unsigned long long key = 0;
  PS_CUDA(ctx, cudaMemcpyAsync(&key, ctx->scalar.p, sizeof(key), cudaMemcpyDeviceToHost, st));
  if (perm_out) PS_CUDA(ctx, cudaMemcpyAsync(perm_out, ctx->perm.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  if (key == 0) {
    *lo = *hi = 0;
    *score = -3.402823466e+38f;
    return PS_OK;
  }
  const unsigned int order = 0xffffffffu - (unsigned int)(key & 0xffffffffu);
  unsigned int u = (unsigned int)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  memcpy(score, &u, sizeof(float));
  if (order < (unsigned)n) {
    *lo = 0;
    *hi = (int)order + 1;
  } else {
    *lo = n - 1 - (int)(order - (unsigned)n);
    *hi = n;
  }
  return PS_OK;
}

template <typename D>
ps_status eval_log_impl(ps_ctx* ctx, long long n, const D* x, D* yg, D* yf) {
  if (!ctx || !x || !yg || !yf || n < 1) return PS_ERR_INVALID;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  PS_TRY(ensure(ctx, ctx->a_in, n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->a_s, n * sizeof(D)));
  PS_TRY(ensure(ctx, ctx->b_s, n * sizeof(D)));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->a_in.p, x, n * sizeof(D), cudaMemcpyHostToDevice, st));
  ps::eval_log_kernel<D><<<cdiv(n, 256), 256, 0, st>>>(n, (const D*)ctx->a_in.p, (D*)ctx->a_s.p, (D*)ctx->b_s.p);
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaMemcpyAsync(yg, ctx->a_s.p, n * sizeof(D), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaMemcpyAsync(yf, ctx->b_s.p, n * sizeof(D), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  return PS_OK;
}

// ------------------------------------------------------------------------------- row sharding
template <typename D>
ps_status shard_create_impl(ps_ctx* ctx, const ps_problem* pp, const D* a_s, const D* b_s, int rank,
                            int world, ps_shard** out) {
  if (!ctx || !pp || !a_s || !b_s || !out || world < 1 || rank < 0 || rank >= world) return PS_ERR_INVALID;
  if (pp->n < 2 || pp->T < 1 || pp->batch > 1) return PS_ERR_INVALID;
  if (pp->dist < 0 || pp->dist > 2) return PS_ERR_DIST;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  ps_shard* s = new (std::nothrow) ps_shard();
  if (!s) return PS_ERR_NOMEM;
  struct Guard {  // any early return below releases what has been allocated so far (cudaFree(nullptr) is a no-op)
    ps_shard* s;
    ~Guard() {
      if (s) ps_shard_destroy(s);
    }
  } guard{s};
  s->ctx = ctx;
  s->dtype = (int)sizeof(D);
  s->p = *pp;
  s->rank = rank;
  s->world = world;
  const int n = pp->n;
  const int Tc = std::min(pp->T, n);
  s->L = choose_chunk(n, 1, pp->chunk_len, BR);
  s->nch_alloc = std::max(1, cdiv(std::max(1, n - 1), s->L));
  const int nrb_total = cdiv(n, BR);
  s->owned_blocks = nrb_total > rank ? (nrb_total - rank + world - 1) / world : 0;
  s->chunk_rows = cdiv(nrb_total, world) * BR;  // same on every rank
  const long long rec_stride = (long long)n + 1;
  cudaStream_t st = ctx->stream;
  PS_CUDA(ctx, cudaMalloc(&s->rec, 2 * rec_stride * sizeof(ps::Rec4<D>)));
  PS_CUDA(ctx, cudaMalloc((void**)&s->ns, (size_t)(Tc + 1) * n * sizeof(int)));
  PS_CUDA(ctx, cudaMemsetAsync(s->ns, 0xff, (size_t)(Tc + 1) * n * sizeof(int), st));
  int running = (pp->sum_mode == PS_SUM_RUNNING);
  if (running && pp->no_fast_math != 1) {
    // safe range, exact summability and sortedness of the (already sorted) inputs -> fast math / screened levels
    PS_CUDA(ctx, cudaMalloc((void**)&s->dflags, 2 * sizeof(int)));
    PS_CUDA(ctx, cudaMemsetAsync(s->dflags, 0, 2 * sizeof(int), st));
    ps::range_check_kernel<D><<<cdiv(n, 256), 256, 0, st>>>(n, a_s, b_s, s->dflags);
    const bool want_screen = pp->screen != 1 && pp->screen != 2 && pp->no_fast_math == 0 && Tc >= 3 && n >= 6000 && n <= 4000000 &&
                             mufu_claims_hold(ctx);
    if (want_screen) {
      PS_TRY(ensure(ctx, ctx->scr_stats, sizeof(ps::ScreenStats)));
      ps::ScreenStats* d_st = (ps::ScreenStats*)ctx->scr_stats.p;
      ps::screen_stats_init_kernel<<<1, 32, 0, st>>>(1, d_st);
      ps::screen_stats_kernel<D><<<dim3(std::max(1, std::min(cdiv(n, 1024), 2048)), 1), 256, 0, st>>>(n, a_s, b_s, d_st);
      ps::screen_decide_kernel<D><<<1, 32, 0, st>>>(1, d_st, s->dflags);
      ps::screen_sorted_kernel<D><<<dim3(cdiv(n, 256), 1), 256, 0, st>>>(n, a_s, b_s, s->dflags + 1);
    }
    int flags[2] = {1, 1};
    PS_CUDA(ctx, cudaMemcpyAsync(flags, s->dflags, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    PS_CUDA(ctx, cudaStreamSynchronize(st));
    s->fast = ((flags[0] & 1) == 0) && (pp->dist != PS_DIST_POISSON || (flags[0] & 2) == 0);
    s->screen = want_screen && s->fast && (flags[0] & 6) == 0;
  }
  if (s->screen) {
    running = 0;  // exactly summable inputs: prefix differences are the reference's running sums, bit for bit
    s->L = std::min(s->L, std::max(512, cdiv(cdiv(n, 200), BR) * BR));
    s->L = std::min(s->L, ps::SCR_MAX_G * BR);
    s->L = std::max(s->L, cdiv(cdiv(n, 1000), BR) * BR);
    s->nch_alloc = std::max(1, cdiv(std::max(1, n - 1), s->L));
  }
  const long long pstride = (long long)s->nch_alloc * n;
  PS_CUDA(ctx, cudaMalloc(&s->part_val, pstride * sizeof(D)));
  PS_CUDA(ctx, cudaMalloc((void**)&s->part_arg, pstride * sizeof(int)));
  if (running && s->nch_alloc > 1) {
    PS_CUDA(ctx, cudaMalloc(&s->SA, pstride * sizeof(D)));
    PS_CUDA(ctx, cudaMalloc(&s->SB, pstride * sizeof(D)));
  }
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)s->rec;
  ps::Rec4<D>* rec1 = rec0 + rec_stride;
  if (s->screen) {
    PS_CUDA(ctx, cudaMalloc(&s->lb, (size_t)n * sizeof(D)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->bmax, (size_t)(n / ps::SCR_BLOCK + 2) * sizeof(float)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->counters, 8 * sizeof(unsigned long long)));
    PS_CUDA(ctx, cudaMemsetAsync(s->counters, 0, 8 * sizeof(unsigned long long), st));
    PS_CUDA(ctx, cudaMalloc((void**)&s->live, (size_t)ps::SCR_CLASSES * s->nch_alloc * (s->owned_blocks + 1) * sizeof(int4)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->rb_info, (size_t)s->nch_alloc * cdiv(n, BR) * sizeof(unsigned int)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->kbest, (size_t)n * sizeof(int)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->rb_slots, (size_t)s->nch_alloc * cdiv(n, BR) * sizeof(int)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->rb_cnt, (size_t)cdiv(n, BR) * sizeof(int)));
    PS_CUDA(ctx, cudaMalloc((void**)&s->live_hdr, (size_t)(Tc + 2) * sizeof(ps::LiveHdr)));
    PS_CUDA(ctx, cudaMemsetAsync(s->live_hdr, 0, (size_t)(Tc + 2) * sizeof(ps::LiveHdr), st));
  }
  if (running)
    ps::build_rec_running_kernel<D><<<dim3(cdiv(n + 1, 256), 1), 256, 0, st>>>(n, rec_stride, a_s, b_s, rec0,
                                                                              rec1);
  else
    ps::build_rec_prefix_kernel<D, 1024><<<1, 1024, 0, st>>>(n, rec_stride, a_s, b_s, rec0, rec1);
  if (s->screen) {
    if constexpr (sizeof(D) == 8) {
      PS_CUDA(ctx, cudaMalloc((void**)&s->recf, 2 * rec_stride * sizeof(float4)));
      ps::build_recf_kernel<<<dim3(cdiv(n + 1, 256), 1), 256, 0, st>>>(n, rec_stride, rec0, s->recf, s->recf + rec_stride);
    }
  }
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  s->cur = 0;
  guard.s = nullptr;
  *out = s;
  return PS_OK;
}

template <typename D>
ps_status shard_level1_impl(ps_shard* s, D* my_chunk) {
  ps_ctx* ctx = s->ctx;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = s->p.n;
  const long long rec_stride = (long long)n + 1;
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)s->rec;
  cudaStream_t st = ctx->stream;
  PS_CUDA(ctx, cudaMemsetAsync(my_chunk, 0, (size_t)s->chunk_rows * sizeof(D), st));
  // level 1 always walks the row (RUNNING) so that SA/SB get filled; in PREFIX mode the same kernel is
  // run on the prefix records' increments?  No: PREFIX shards use the closed form below.
  if (s->screen) {
    if (s->owned_blocks > 0)
      ps::level1_prefix_shard_kernel<D, BR><<<s->owned_blocks, BR, 0, st>>>(
          n, s->p.dist, s->p.risk_part ? 1 : 0, rec0, s->ns + n, s->rank, s->world, my_chunk);
  } else if (s->p.sum_mode == PS_SUM_RUNNING) {
    if (s->owned_blocks > 0)
      ps::level1_running_kernel<D, BR><<<dim3((s->owned_blocks + 1) / 2, 1), BR, 0, st>>>(
          n, s->L, s->nch_alloc, rec_stride, s->p.dist, s->p.risk_part ? 1 : 0, rec0, (D*)s->SA, (D*)s->SB,
          (D*)nullptr, s->ns + n, 0, (ps::Rec4<D>*)nullptr, s->rank, s->world, my_chunk, s->owned_blocks);
  } else {
    return PS_ERR_UNSUPPORTED;
  }
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  return PS_OK;
}

template <typename D>
ps_status shard_set_prev_impl(ps_shard* s, const D* gathered) {
  ps_ctx* ctx = s->ctx;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = s->p.n;
  const long long rec_stride = (long long)n + 1;
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)s->rec;
  s->cur ^= 1;
  ps::Rec4<D>* dst = rec0 + (long long)s->cur * rec_stride;
  const long long tot = (long long)s->world * s->chunk_rows;
  float* pm_next = (s->screen && sizeof(D) == 8) ? (float*)(s->recf + (long long)s->cur * rec_stride) : nullptr;
  ps::shard_unpermute_kernel<D, BR><<<cdiv(tot, 256), 256, 0, ctx->stream>>>(n, s->world, s->chunk_rows,
                                                                            gathered, dst, (D*)nullptr, pm_next);
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PS_OK;
}

template <typename D>
ps_status shard_level_impl(ps_shard* s, int j, D* my_chunk) {
  ps_ctx* ctx = s->ctx;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = s->p.n;
  const int Tc = std::min(s->p.T, n);
  if (j < 2 || j > Tc) return PS_ERR_INVALID;
  const long long rec_stride = (long long)n + 1;
  ps::Rec4<D>* rec0 = (ps::Rec4<D>*)s->rec;
  const ps::Rec4<D>* rin = rec0 + (long long)s->cur * rec_stride;
  ps::Geom g = make_geom(n, j, Tc, s->L, s->nch_alloc, s->rank, s->world);
  cudaStream_t st = ctx->stream;
  PS_CUDA(ctx, cudaMemsetAsync(my_chunk, 0, (size_t)s->chunk_rows * sizeof(D), st));
  if (g.ntiles > 0 && s->screen && j < Tc) {
    // lower bounds and pm maxima for every row / record (cheap; each rank needs them for all k); range tiers, live tiles and
    // the fold for owned row blocks only
    float4* rf = (sizeof(D) == 8) ? s->recf + (long long)s->cur * rec_stride : (float4*)rin;
    const int rp = s->p.risk_part ? 1 : 0;
    auto fn_lb = ScreenLaunch<D>::pick_lb(s->p.dist, rp);
    const int sel_tpr = (s->nch_alloc >= 32) ? 4 : 1;
    auto fn_sel = ScreenLaunch<D>::pick_select(s->p.dist, rp, 0, sel_tpr);
    auto fn_scr = ScreenLaunch<D>::pick(s->p.dist, rp, 0);
    fn_lb<<<dim3(cdiv(g.nrows, 128), 1), 128, 0, st>>>(g, rin, rf, (D*)s->lb, s->bmax, n / ps::SCR_BLOCK + 2, s->kbest);
    const long long live_cap = (long long)s->nch_alloc * (s->owned_blocks + 1);
    fn_sel<<<dim3(g.NRB, 1), 128 * sel_tpr, (size_t)g.NCH * (4 * sizeof(D) + 12 + 1), st>>>(
        g, rin, (const D*)s->lb, s->bmax, n / ps::SCR_BLOCK + 2, s->dflags, s->live, live_cap, s->rb_info, cdiv(n, BR),
        s->live_hdr + j, s->counters, s->kbest, s->rb_slots, s->rb_cnt, ps::SelSrc{});
    ps::LevelOut<D> lo;
    lo.ms_row = nullptr;
    lo.ns_row = s->ns + (long long)j * n;
    lo.table_stride = 0;
    lo.rec_next = nullptr;
    lo.pm_next = nullptr;
    lo.shard_chunk = my_chunk;
    fn_scr<<<std::min(g.ntiles, ctx->sm_count * 8), BR, 0, st>>>(g, rin, rf, (const D*)s->lb, (D*)s->part_val, s->part_arg,
                                                        s->counters, s->dflags, s->live, live_cap, s->live_hdr + j, nullptr);
    ps::screen_fold_kernel<D, BR, 4><<<dim3(g.NRB, 1), BR * 4, 0, st>>>(g, s->rb_slots, s->rb_cnt, cdiv(n, BR),
                                                                       (const D*)s->part_val, s->part_arg, lo);
  } else if (g.ntiles > 0) {
    const int running = (s->p.sum_mode == PS_SUM_RUNNING) && !s->screen;
    auto fn = LevelLaunch<D>::pick(s->p.dist, s->p.risk_part ? 1 : 0, running, s->fast);
    fn<<<dim3(g.ntiles, 1), BR, 0, st>>>(g, rin, (const D*)s->SA, (const D*)s->SB, (D*)s->part_val,
                                         s->part_arg);
    ps::combine_kernel<D, BR, 4><<<dim3(cdiv(4LL * g.NRB * BR, 256), 1), 256, 0, st>>>(
        g, (const D*)s->part_val, s->part_arg, (D*)nullptr, s->ns + (long long)j * n, 0,
        (ps::Rec4<D>*)nullptr, my_chunk, nullptr);
  }
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  return PS_OK;
}

template <typename D>
ps_status shard_subset_scores_impl(ps_shard* s, int count, const int* bounds, D* scores) {
  ps_ctx* ctx = s->ctx;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = s->p.n;
  if (count < 1 || count > n) return PS_ERR_INVALID;
  cudaStream_t st = ctx->stream;
  PS_TRY(ensure(ctx, ctx->bounds, (size_t)(count + 1) * sizeof(int)));
  PS_TRY(ensure(ctx, ctx->scores, (size_t)count * sizeof(D)));
  PS_CUDA(ctx, cudaMemcpyAsync(ctx->bounds.p, bounds, (size_t)(count + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  const int prefix = s->screen || s->p.sum_mode == PS_SUM_PREFIX;
  ps::shard_subset_scores_kernel<D><<<cdiv(count, 32), 32, 0, st>>>(count, n, s->p.dist, s->p.risk_part ? 1 : 0, prefix,
                                                                   (const int*)ctx->bounds.p,
                                                                   (const ps::Rec4<D>*)s->rec, (D*)ctx->scores.p);
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaMemcpyAsync(scores, ctx->scores.p, (size_t)count * sizeof(D), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  return PS_OK;
}

}  // namespace

// =================================================================================== C ABI
namespace {

constexpr size_t PINNED_IN_BYTES = 256 * 1024;

// Small solves from host buffers, replayed from a CUDA graph.  Such a solve is ~16 kernels of a few microseconds each and ~45
// runtime calls (launches, event records, copies); the host's call rate, not the device, sets its latency.  The first solve
// of a shape runs eagerly (it sizes the workspaces), the second is recorded by running solve_impl under stream capture
// (no events, no synchronisation, the device flags assumed to pass), and from then on a solve is: copy a, b into pinned
// memory, ONE cudaGraphLaunch, one synchronisation, copy the results out of pinned memory.  The only data-dependent host
// decision of these shapes is the safe-range / positivity precheck (math_fast.cuh); the recorded graph still computes the
// flags, and a replay whose flags would have sent the eager solve down another path is discarded and redone eagerly.
// A graph holds workspace pointers: any reallocation (ps_ctx::alloc_epoch) retires it.
template <typename D>
ps_status solve_small_graph(ps_ctx* ctx, const ps_problem* pp, const D* a, const D* b, ps_result* res) {
  if (!ctx || !pp || !a || !b || !res || !ctx->graphs_enabled) return solve_impl<D>(ctx, pp, a, b, res);
  const ps_problem& p = *pp;
  const long long n = p.n, R = p.batch;
  const long long nrb128 = (n + BR - 1) / BR;
  const size_t in_bytes = (size_t)(R * n) * sizeof(D);
  const bool eligible = n > SMALL_ROWS && n <= ps::SORT_SMALL_MAX && R >= 1 && p.T >= 1 && !p.on_device && !p.perm &&
                        p.sort_mode != PS_SORT_GIVEN && p.sum_mode == PS_SUM_RUNNING && p.no_fast_math == 0 &&
                        p.chunk_len <= 0 && p.screen >= 0 && p.screen <= 2 && p.dist >= 0 && p.dist <= 2 &&
                        !ctx->dbg_level && 2 * in_bytes <= PINNED_IN_BYTES &&
                        R * nrb128 * (nrb128 + 1) / 2 < 4LL * ctx->sm_count &&
                        // below the screening threshold of solve_impl (want_screen)
                        (p.screen == 1 || std::min<long long>(p.T, n) < 3 || n < 512 || n * R < 6000);
  if (!eligible) return solve_impl<D>(ctx, pp, a, b, res);
  const std::array<long long, 12> key = {(long long)sizeof(D), n, p.T, p.dist, p.risk_part ? 1 : 0, p.sweep ? 1 : 0, R,
                                         (res->perm ? 1 : 0) | (res->bounds ? 2 : 0) | (res->subset_scores ? 4 : 0) |
                                             (res->max_score ? 8 : 0) | (res->next_start ? 16 : 0),
                                         p.screen, 0, 0, 0};
  if (ctx->graphs.size() >= 256 && !ctx->graphs.count(key)) {  // a caller that sweeps over many shapes: start over
    for (auto& kv : ctx->graphs)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    ctx->graphs.clear();
  }
  GraphEntry& ge = ctx->graphs[key];
  if (ge.exec && ge.epoch != ctx->alloc_epoch) {  // a workspace moved since the recording
    cudaGraphExecDestroy(ge.exec);
    ge.exec = nullptr;
    ge.seen = 1;
  }
  if (!ge.exec) {
    if (ge.failed || ge.seen == 0) {
      ge.seen++;
      return solve_impl<D>(ctx, pp, a, b, res);
    }
    // second solve of this shape: record it
    if (cudaSetDevice(ctx->device) != cudaSuccess) return solve_impl<D>(ctx, pp, a, b, res);
    if (!ctx->pinned_in && cudaMallocHost((void**)&ctx->pinned_in, PINNED_IN_BYTES) != cudaSuccess) {
      cudaGetLastError();
      ge.failed = 1;
      return solve_impl<D>(ctx, pp, a, b, res);
    }
    const D* pa = reinterpret_cast<const D*>(ctx->pinned_in);
    const D* pb = reinterpret_cast<const D*>(ctx->pinned_in + in_bytes);
    cudaGraph_t graph = nullptr;
    ps_status st = PS_ERR_CUDA;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      ctx->capturing = true;
      st = solve_impl<D>(ctx, pp, pa, pb, res);
      ctx->capturing = false;
      if (cudaStreamEndCapture(ctx->stream, &graph) != cudaSuccess) st = PS_ERR_CUDA;
    }
    if (st == PS_OK && graph && cudaGraphInstantiate(&ge.exec, graph, 0) != cudaSuccess) {
      ge.exec = nullptr;
      st = PS_ERR_CUDA;
    }
    if (graph) cudaGraphDestroy(graph);
    if (st != PS_OK || !ge.exec) {
      cudaGetLastError();
      ge.failed = 1;
      ge.exec = nullptr;
      return solve_impl<D>(ctx, pp, a, b, res);
    }
    ge.epoch = ctx->alloc_epoch;
    ge.staged = ctx->cap_staged;
    ge.launches = ctx->tm.kernel_launches;
    ge.h2d_bytes = ctx->tm.h2d_bytes;
    ge.d2h_bytes = ctx->tm.d2h_bytes;
    ge.poisson = p.dist == PS_DIST_POISSON;
  }
  // replay
  ctx->err.clear();
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  memcpy(ctx->pinned_in, a, in_bytes);
  memcpy(ctx->pinned_in + in_bytes, b, in_bytes);
  ctx->ev_used = 0;
  cudaEvent_t e0, e1;
  PS_TRY(mark(ctx, &e0));
  PS_CUDA(ctx, cudaGraphLaunch(ge.exec, ctx->stream));
  PS_TRY(mark(ctx, &e1));
  PS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int flags = *ctx->pinned_flags;
  if ((flags & 1) || (ge.poisson && (flags & 2)))  // these inputs take the general-math kernels: solve them eagerly
    return solve_impl<D>(ctx, pp, a, b, res);
  scatter_staged(ctx, ge.staged, res);
  memset(&ctx->tm, 0, sizeof(ctx->tm));
  cudaEventElapsedTime(&ctx->tm.total_ms, e0, e1);
  ctx->tm.kernel_launches = ge.launches;
  ctx->tm.h2d_bytes = ge.h2d_bytes;
  ctx->tm.d2h_bytes = ge.d2h_bytes;
  ctx->last_fast = 1;
  return PS_OK;
}

}  // namespace

extern "C" {

int ps_abi_version(void) { return PS_B200_ABI_VERSION; }

const char* ps_status_string(ps_status s) {
  switch (s) {
    case PS_OK: return "ok";
    case PS_ERR_INVALID: return "invalid argument";
    case PS_ERR_DIST: return "Bad distributional assignment";
    case PS_ERR_CUDA: return "CUDA error";
    case PS_ERR_NOMEM: return "out of memory";
    case PS_ERR_NODEVICE: return "no CUDA device (this library has no CPU path)";
    case PS_ERR_UNSUPPORTED: return "unsupported";
    case PS_ERR_COMM: return "NCCL error";
  }
  return "unknown";
}

ps_status ps_create(ps_ctx** out, int device) {
  if (!out) return PS_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return PS_ERR_NODEVICE;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return PS_ERR_NODEVICE;
  }
  if (device >= count) return PS_ERR_INVALID;
  ps_ctx* ctx = new (std::nothrow) ps_ctx();
  if (!ctx) return PS_ERR_NOMEM;
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return PS_ERR_CUDA;
  }
  memset(&ctx->tm, 0, sizeof(ctx->tm));
  if (const char* e = getenv("PS_DEBUG_LEVEL")) ctx->dbg_level = atoi(e);
  if (const char* e = getenv("PS_NO_GRAPH")) ctx->graphs_enabled = (atoi(e) == 0);
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->sm_count = sms;
  }
  *out = ctx;
  return PS_OK;
}

void ps_destroy(ps_ctx* ctx) {
  if (!ctx) return;
  if (cudaSetDevice(ctx->device) == cudaErrorCudartUnloading) {  // process exit after the runtime went away: nothing to free
    delete ctx;
    return;
  }
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->a_in,  &ctx->b_in,     &ctx->key_in,   &ctx->key_out, &ctx->idx_in, &ctx->perm,
                    &ctx->a_s,   &ctx->b_s,      &ctx->rec,      &ctx->SA,      &ctx->SB,     &ctx->part_val,
                    &ctx->part_arg, &ctx->ms,    &ctx->ns,       &ctx->bounds,  &ctx->scores, &ctx->sort_hist,
                    &ctx->scalar, &ctx->recf,    &ctx->lb,       &ctx->scr_stats, &ctx->scr_counters, &ctx->tile_map,
                    &ctx->scr_bmax, &ctx->fxinfo, &ctx->pfx, &ctx->cpA, &ctx->cpB, &ctx->ipos,
                    &ctx->live, &ctx->rb_info, &ctx->live_hdr, &ctx->kbest, &ctx->rb_slots, &ctx->rb_cnt,
                    &ctx->shard_send, &ctx->shard_recv};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
  if (ctx->pinned_flags) cudaFreeHost(ctx->pinned_flags);
  if (ctx->pinned_out) cudaFreeHost(ctx->pinned_out);
  if (ctx->pinned_in) cudaFreeHost(ctx->pinned_in);
  for (auto& kv : ctx->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
  if (ctx->pinned_levels) cudaFreeHost(ctx->pinned_levels);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* ps_last_error(const ps_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int ps_device(const ps_ctx* ctx) { return ctx ? ctx->device : -1; }
int ps_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
uintptr_t ps_stream(const ps_ctx* ctx) { return ctx ? (uintptr_t)ctx->stream : 0; }

ps_status ps_stream_wait(ps_ctx* ctx, uintptr_t other_stream) {
  if (!ctx) return PS_ERR_INVALID;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->wait_event) PS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->wait_event, cudaEventDisableTiming));
  PS_CUDA(ctx, cudaEventRecord(ctx->wait_event, (cudaStream_t)other_stream));
  PS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->wait_event, 0));
  return PS_OK;
}

ps_status ps_solve_f32(ps_ctx* ctx, const ps_problem* p, const float* a, const float* b, ps_result* r) {
  return solve_small_graph<float>(ctx, p, a, b, r);
}
ps_status ps_solve_f64(ps_ctx* ctx, const ps_problem* p, const double* a, const double* b, ps_result* r) {
  return solve_small_graph<double>(ctx, p, a, b, r);
}

// ---- multi-GPU: the library drives NCCL itself, on the context's stream
ps_status ps_comm_unique_id(void* id_out) {
  if (!id_out) return PS_ERR_INVALID;
  NcclApi& nc = nccl_api();
  if (!nc.ok) return PS_ERR_COMM;
  ncclUniqueId id;
  if (nc.GetUniqueId(&id) != ncclSuccess) return PS_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == PS_COMM_ID_BYTES, "ps_b200.h: PS_COMM_ID_BYTES");
  memcpy(id_out, &id, sizeof(id));
  return PS_OK;
}

ps_status ps_comm_create(ps_ctx* ctx, const void* id_bytes, int rank, int world, ps_comm** out) {
  if (!ctx || !id_bytes || !out || world < 1 || rank < 0 || rank >= world) return PS_ERR_INVALID;
  *out = nullptr;
  NcclApi& nc = nccl_api();
  if (!nc.ok) {
    ctx->err = nc.err;
    return PS_ERR_COMM;
  }
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ps_comm* c = new (std::nothrow) ps_comm();
  if (!c) return PS_ERR_NOMEM;
  c->rank = rank;
  c->world = world;
  c->device = ctx->device;
  ncclResult_t r = nc.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    ctx->err = std::string("ncclCommInitRank: ") + nc.GetErrorString(r);
    delete c;
    return PS_ERR_COMM;
  }
  *out = c;
  return PS_OK;
}

void ps_comm_destroy(ps_comm* c) {
  if (!c) return;
  if (c->comm && nccl_api().ok) {
    cudaSetDevice(c->device);
    nccl_api().CommDestroy(c->comm);
  }
  delete c;
}
int ps_comm_rank(const ps_comm* c) { return c ? c->rank : -1; }
int ps_comm_world(const ps_comm* c) { return c ? c->world : 0; }
int ps_nccl_version(void) {
  int v = 0;
  NcclApi& nc = nccl_api();
  if (nc.ok && nc.GetVersion) nc.GetVersion(&v);
  return v;
}

ps_status ps_solve_sharded_f32(ps_ctx* ctx, ps_comm* comm, const ps_problem* p, const float* a, const float* b,
                               ps_result* r) {
  if (!comm) return PS_ERR_INVALID;
  return solve_impl<float>(ctx, p, a, b, r, nullptr, nullptr, comm);
}
ps_status ps_solve_sharded_f64(ps_ctx* ctx, ps_comm* comm, const ps_problem* p, const double* a, const double* b,
                               ps_result* r) {
  if (!comm) return PS_ERR_INVALID;
  return solve_impl<double>(ctx, p, a, b, r, nullptr, nullptr, comm);
}

ps_status ps_get_timings(const ps_ctx* ctx, ps_timings* out) {
  if (!ctx || !out) return PS_ERR_INVALID;
  *out = ctx->tm;
  return PS_OK;
}

ps_status ps_range_score_f32(ps_ctx* ctx, int n, const float* a, const float* b, int dist, int risk_part,
                             float* out) {
  return range_score_impl<float>(ctx, n, a, b, dist, risk_part ? 1 : 0, out);
}
ps_status ps_range_score_f64(ps_ctx* ctx, int n, const double* a, const double* b, int dist, int risk_part,
                             double* out) {
  return range_score_impl<double>(ctx, n, a, b, dist, risk_part ? 1 : 0, out);
}

ps_status ps_selftest_math(ps_ctx* ctx, int which, long long samples, unsigned long long seed,
                           long long* mismatches) {
  if (!ctx || !mismatches || which < 0 || which > 3 || samples < 1) return PS_ERR_INVALID;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  PS_TRY(ensure(ctx, ctx->scalar, 64));
  cudaStream_t st = ctx->stream;
  PS_CUDA(ctx, cudaMemsetAsync(ctx->scalar.p, 0, sizeof(unsigned long long), st));
  const int blocks = 148 * 8, threads = 256;
  const long long per_thread = (samples + (long long)blocks * threads - 1) / ((long long)blocks * threads);
  ps::selftest_math_kernel<<<blocks, threads, 0, st>>>(which, per_thread, seed, (unsigned long long*)ctx->scalar.p);
  PS_CUDA(ctx, cudaGetLastError());
  unsigned long long bad = 0;
  PS_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->scalar.p, sizeof(bad), cudaMemcpyDeviceToHost, st));
  PS_CUDA(ctx, cudaStreamSynchronize(st));
  *mismatches = (long long)bad;
  return PS_OK;
}

ps_status ps_selftest_mufu(ps_ctx* ctx, double out[2]) {
  if (!ctx || !out) return PS_ERR_INVALID;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  PS_TRY(ensure(ctx, ctx->scalar, 64));
  PS_CUDA(ctx, cudaMemsetAsync(ctx->scalar.p, 0, 2 * sizeof(double), ctx->stream));
  ps::selftest_mufu_kernel<<<148 * 8, 256, 0, ctx->stream>>>((double*)ctx->scalar.p);
  PS_CUDA(ctx, cudaGetLastError());
  PS_CUDA(ctx, cudaMemcpyAsync(out, ctx->scalar.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PS_OK;
}

ps_status ps_ltss_f32(ps_ctx* ctx, int n, const float* a, const float* b, int dist, int* perm, int* lo, int* hi,
                      float* score) {
  return ltss_impl(ctx, n, a, b, dist, perm, lo, hi, score);
}

ps_status ps_eval_log_f64(ps_ctx* ctx, long long n, const double* x, double* y_general, double* y_fast) {
  return eval_log_impl<double>(ctx, n, x, y_general, y_fast);
}
ps_status ps_eval_log_f32(ps_ctx* ctx, long long n, const float* x, float* y_general, float* y_fast) {
  return eval_log_impl<float>(ctx, n, x, y_general, y_fast);
}

ps_status ps_shard_create_f64(ps_ctx* ctx, const ps_problem* p, const double* a, const double* b, int rank,
                              int world, ps_shard** out) {
  return shard_create_impl<double>(ctx, p, a, b, rank, world, out);
}
ps_status ps_shard_create_f32(ps_ctx* ctx, const ps_problem* p, const float* a, const float* b, int rank,
                              int world, ps_shard** out) {
  return shard_create_impl<float>(ctx, p, a, b, rank, world, out);
}
int ps_shard_chunk_rows(const ps_shard* s) { return s ? s->chunk_rows : 0; }

ps_status ps_shard_level1(ps_shard* s, void* my_chunk_dev) {
  if (!s || !my_chunk_dev) return PS_ERR_INVALID;
  return s->dtype == 4 ? shard_level1_impl<float>(s, (float*)my_chunk_dev)
                       : shard_level1_impl<double>(s, (double*)my_chunk_dev);
}
ps_status ps_shard_set_prev(ps_shard* s, const void* gathered_dev) {
  if (!s || !gathered_dev) return PS_ERR_INVALID;
  return s->dtype == 4 ? shard_set_prev_impl<float>(s, (const float*)gathered_dev)
                       : shard_set_prev_impl<double>(s, (const double*)gathered_dev);
}
ps_status ps_shard_level(ps_shard* s, int j, void* my_chunk_dev) {
  if (!s || !my_chunk_dev) return PS_ERR_INVALID;
  return s->dtype == 4 ? shard_level_impl<float>(s, j, (float*)my_chunk_dev)
                       : shard_level_impl<double>(s, j, (double*)my_chunk_dev);
}
int ps_shard_owner(const ps_shard* s, int i) { return s ? (i / BR) % s->world : -1; }

ps_status ps_shard_next_start(ps_shard* s, int j, int i, int* out) {
  if (!s || !out) return PS_ERR_INVALID;
  const int n = s->p.n;
  const int Tc = std::min(s->p.T, n);
  if (j < 1 || j > Tc || i < 0 || i >= n) return PS_ERR_INVALID;
  if (ps_shard_owner(s, i) != s->rank) {
    *out = -2;
    return PS_OK;
  }
  ps_ctx* ctx = s->ctx;
  PS_CUDA(ctx, cudaSetDevice(ctx->device));
  PS_CUDA(ctx, cudaMemcpyAsync(out, s->ns + (long long)j * n + i, sizeof(int), cudaMemcpyDeviceToHost,
                               ctx->stream));
  PS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return PS_OK;
}

ps_status ps_shard_subset_scores(ps_shard* s, int count, const int* bounds, void* scores) {
  if (!s || !bounds || !scores) return PS_ERR_INVALID;
  return s->dtype == 4 ? shard_subset_scores_impl<float>(s, count, bounds, (float*)scores)
                       : shard_subset_scores_impl<double>(s, count, bounds, (double*)scores);
}

int ps_shard_screened(const ps_shard* s) { return s ? s->screen : 0; }
void ps_shard_destroy(ps_shard* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaFree(s->rec);
  cudaFree(s->SA);
  cudaFree(s->SB);
  cudaFree(s->part_val);
  cudaFree(s->part_arg);
  cudaFree(s->ns);
  cudaFree(s->recf);
  cudaFree(s->lb);
  cudaFree(s->bmax);
  cudaFree(s->dflags);
  cudaFree(s->tile_map);
  cudaFree(s->counters);
  cudaFree(s->live);
  cudaFree(s->rb_info);
  cudaFree(s->live_hdr);
  cudaFree(s->kbest);
  cudaFree(s->rb_slots);
  cudaFree(s->rb_cnt);
  delete s;
}

}  // extern "C"
