/* ps_b200.h -- C ABI of the B200-native consecutive-partitions dynamic-programming engine.
 *
 * This is the drop-in boundary for ONE path of pehlivanian/PartitionSolvers: everything the
 * constructor of `DPSolver<DataType>` does (reference: src/cpp/include/DP.hpp:43-149,194-197 ->
 * DP_impl.hpp:69-230, score_impl.hpp:233-285).  The reference has no FFI of its own; its hot path
 * is a C++ class template.  The host facade in include/partitionsolvers/DP.hpp keeps that class
 * (same constructors and getters) and calls ONLY the functions declared here.  INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No C++ or torch types cross this boundary.
 *   - the caller owns every buffer it passes; the library owns its device workspaces (inside ps_ctx).
 *   - errors are return codes (ps_status); nothing throws across the ABI.
 *   - a ps_ctx may be used by one host thread at a time; any number of contexts may run concurrently
 *     (the reference constructs solvers from pool threads: python_dpsolver.cpp:275-301).  CUDA is
 *     initialised lazily inside ps_create, never at load time, so fork()ed workers are safe
 *     (solverSWIG_DP.py:152-167).
 *   - there is no CPU fallback: without a usable CUDA device ps_create fails with PS_ERR_NODEVICE.
 */
#ifndef PS_B200_H
#define PS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_B200_ABI_VERSION 3

typedef enum ps_status {
  PS_OK = 0,
  PS_ERR_INVALID = 1,     /* bad argument (n<1, T<1, NULL input, unknown enum value ...)            */
  PS_ERR_DIST = 2,        /* parametric_dist outside {0,1,2}: the facade maps this to                */
                          /* distributionException (reference LTSS.hpp:21-25, DP_impl.hpp:62-64)      */
  PS_ERR_CUDA = 3,        /* a CUDA runtime call failed; ps_last_error(ctx) has the text              */
  PS_ERR_NOMEM = 4,       /* device or host allocation failed                                         */
  PS_ERR_NODEVICE = 5,    /* no CUDA device / driver: there is deliberately no CPU path               */
  PS_ERR_UNSUPPORTED = 6,
  PS_ERR_COMM = 7         /* NCCL missing or an NCCL call failed; ps_last_error(ctx) has the text              */
} ps_status;

/* objective_fn of the reference (score.hpp:20-22) */
enum { PS_DIST_GAUSSIAN = 0, PS_DIST_POISSON = 1, PS_DIST_RATIONAL = 2 };

/* How range sums Sigma a, Sigma b over [i,k) are formed.
 * RUNNING reproduces the reference's rounding exactly: sums are accumulated left to right starting
 * at i (score_impl.hpp:248-256).  PREFIX forms them as differences of a device prefix scan
 * (bit-identical to RUNNING whenever all partial sums are exactly representable, e.g. counts). */
enum { PS_SUM_RUNNING = 0, PS_SUM_PREFIX = 1 };

/* Priority sort (DP_impl.hpp:7-28).  DEVICE = stable LSD radix sort of a[i]/b[i] on the GPU (ties
 * keep input order).  GIVEN = the caller supplies the permutation (e.g. std::sort's, whose order on
 * tied priorities is implementation-defined and cannot be reproduced by any other algorithm). */
enum { PS_SORT_DEVICE = 0, PS_SORT_GIVEN = 1 };

typedef struct ps_ctx ps_ctx; /* opaque: device, stream, workspaces, timing events */

typedef struct ps_problem {
  int n;           /* ground-set size; a and b hold n values per instance                            */
  int T;           /* subsets requested; capped at n like DP_impl.hpp:37-38                          */
  int dist;        /* PS_DIST_*                                                                       */
  int risk_part;   /* risk_partitioning_objective (0 = multiple clustering)                           */
  int sum_mode;    /* PS_SUM_*                                                                        */
  int sort_mode;   /* PS_SORT_*                                                                       */
  const int* perm; /* PS_SORT_GIVEN: [batch][n] permutation, position -> original index              */
  int sweep;       /* 0: backtrace S=T only (DP_impl.hpp:226-228); 1: every S=T..1 (:203-214)         */
  int batch;       /* independent instances R>=1 (replicas); a,b are [R][n] row-major.  Host-buffer    */
                   /* batches of >= 256 instances and >= 16 MB are solved in four parts whose copies   */
                   /* overlap the kernels of the parts before them (same outputs as one call; the       */
                   /* overlap needs PINNED host buffers, pageable ones are staged synchronously).       */
                   /* More than 65535 instances are solved in consecutive slices of that size.          */
  int on_device;   /* 1: a, b, perm AND every ps_result pointer are device pointers on ctx's device   */
  int no_fast_math;/* debug: 1 = never take the range-prechecked branch-free math path (csrc/math_fast.cuh);  */
                   /*        2 = take it, but not the packed two-rows-per-thread float kernel               */
  int chunk_len;   /* tuning, 0 = default                                                             */
  int screen;      /* levels 2..T-1: 0 = auto: the screened kernel (csrc/screen.cuh: a float upper bound per  */
                   /* cell, the reference's arithmetic only for cells that can still be the row's maximum;    */
                   /* same tables bit for bit) when the instance qualifies (PS_SUM_RUNNING, safe range, a > 0,*/
                   /* exactly summable inputs such as counts), else the exhaustive kernel;                    */
                   /* 1 = always the exhaustive kernel; 2 = screened + audit (every cell is ALSO evaluated    */
                   /* exactly and ps_timings.screen_violations counts cells whose bound was too low)          */
} ps_problem;

/* Caller-allocated outputs; any pointer may be NULL (= not wanted).  D = float or double.
 * Let Tc = min(T, n).
 *   perm          [R][n]          priority_sortind_: sorted position -> original index
 *   bounds        sweep=0: [R][Tc+1]          0 = b_0 < b_1 < ... < b_Tc = n  (subset t = positions
 *                 sweep=1: [R][Tc+1][Tc+1]    [b_t, b_{t+1}) of the sorted order); row S of the sweep
 *                                             table holds the S+1 boundaries of the S-partition
 *   subset_scores sweep=0: [R][Tc]            S(b_t, b_{t+1}) in backtrace order, BEFORE the
 *                 sweep=1: [R][Tc+1][Tc]      mult-clust reorder (the facade applies DP_impl.hpp:232-259)
 *   max_score     [R][Tc+1][n]    maxScore_  rows 0..Tc  (unwritten cells = lowest(), like the reference)
 *   next_start    [R][Tc+1][n]    nextStart_ rows 0..Tc  (unwritten cells = -1)
 */
typedef struct ps_result {
  int* perm;
  int* bounds;
  void* subset_scores;
  void* max_score;
  int* next_start;
} ps_result;

#define PS_MAX_TIMED_LEVELS 130
typedef struct ps_timings {
  float total_ms;       /* whole call on the context's stream (CUDA events)                           */
  float h2d_ms;
  float sort_scan_ms;   /* keys + radix sort + gather + scan / record build                           */
  float prepass_ms;     /* level 1 (+ chunk-start sums in RUNNING mode)                               */
  float levels_ms;      /* sum over levels 2..T of tile kernel + combine                              */
  float backtrace_ms;
  float d2h_ms;
  int n_levels;                           /* number of level launches timed                          */
  float level_ms[PS_MAX_TIMED_LEVELS];    /* tile kernel only, per level j = 2.. (index j-2)         */
  long long level_cells[PS_MAX_TIMED_LEVELS]; /* (i,k) cells evaluated by that launch, all instances  */
  long long kernel_launches;              /* kernels launched by the call                            */
  long long h2d_bytes, d2h_bytes;
  int screen_used;                  /* 1: screened levels (exact sums); 2: screened levels, running sums; 0: exhaustive */
  long long screen_exact_cells;     /* cells (32 per warp vote) evaluated with the reference's arithmetic      */
  long long screen_violations;      /* audit mode only; must be 0                                              */
  long long screen_tiers[6];        /* audit mode only: tiles kept, dropped; warp-stages kept, dropped;        */
                                    /* warp-groups kept, dropped (a stage = 128 k, a group = 8 k, x 32 rows)   */
  long long comm_bytes;             /* row-sharded solves: bytes this rank received through ncclAllGather          */
  int screen_fallback_level;        /* 0, or the first level handed to the exhaustive kernel because a screened  */
                                    /* level had to evaluate > 30 % of its cells exactly (inputs whose rows have */
                                    /* no pronounced maximum, e.g. constant a/b); tables are identical either way */
} ps_timings;

int ps_abi_version(void);
const char* ps_status_string(ps_status s);

/* device < 0 selects the current CUDA device.  Creates a non-blocking stream. */
ps_status ps_create(ps_ctx** out, int device);
void ps_destroy(ps_ctx* ctx);
const char* ps_last_error(const ps_ctx* ctx);
int ps_device(const ps_ctx* ctx);
/* CUDA devices visible to this process (0 if the runtime reports none); for callers that spread independent instances over
 * several devices, one context and one host thread per device (include/partitionsolvers/replicas.hpp does exactly that) */
int ps_device_count(void);
/* cudaStream_t of the context, as an integer (for callers that order their own work after ours) */
uintptr_t ps_stream(const ps_ctx* ctx);
/* The engine works on its OWN non-blocking stream, which has no implicit ordering with any other stream.  Device buffers
 * handed to the library (ps_problem.on_device = 1, ps_shard_*) must therefore be complete when the call is made -- or the
 * caller orders them explicitly: ps_stream_wait makes everything the engine enqueues from now on wait for the work
 * `other_stream` (a cudaStream_t as an integer; 0 = the legacy default stream) has been given so far.  Results are complete
 * on return (every ps_solve_* blocks on the engine's stream). */
ps_status ps_stream_wait(ps_ctx* ctx, uintptr_t other_stream);

/* One call = the whole constructor of DPSolver<D>: sort_by_priority, range scores, the T-level DP
 * fill and the backtrace(s) (DP_impl.hpp:69-153,200-230).  Blocking. */
/* Small solves from host buffers (n <= 4096, below the screening threshold, device sort, RUNNING sums, inputs <= 256 KB):
 * the second solve of a shape on a context is recorded as a CUDA graph and later ones replay it -- one graph launch instead of
 * ~45 runtime calls (n = 1000, T = 5: 0.23 -> 0.145 ms per call).  Results are identical; ps_timings of a replayed solve
 * carries total_ms, kernel_launches and the byte counts only.  A replay whose inputs fail the safe-range / positivity
 * precheck is redone on the general-math kernels.  PS_NO_GRAPH=1 in the environment at ps_create time switches this off. */
ps_status ps_solve_f32(ps_ctx* ctx, const ps_problem* p, const float* a, const float* b, ps_result* r);
ps_status ps_solve_f64(ps_ctx* ctx, const ps_problem* p, const double* a, const double* b, ps_result* r);

/* Timings of the most recent ps_solve_* on this context. */
ps_status ps_get_timings(const ps_ctx* ctx, ps_timings* out);

/* S(0,n) of the inputs as given, no sort: `compute_score` (python_dpsolver.cpp:74-111). */
ps_status ps_range_score_f32(ps_ctx* ctx, int n, const float* a, const float* b, int dist, int risk_part,
                             float* out);
ps_status ps_range_score_f64(ps_ctx* ctx, int n, const double* a, const double* b, int dist, int risk_part,
                             double* out);

/* LTSSSolver<float> (reference LTSS_impl.hpp:14-35,73-111): stable priority sort, then the best single subset
 * among all prefixes and suffixes of the sorted order under the multiple-clustering score.  The subset is
 * perm[lo:hi]; score is float.  perm may be NULL.  If no candidate beats -FLT_MAX (all scores NaN), lo=hi=0. */
ps_status ps_ltss_f32(ps_ctx* ctx, int n, const float* a, const float* b, int dist, int* perm, int* lo, int* hi,
                      float* score);

/* Self-test of the branch-free routines used when a per-instance precheck proves the inputs well-scaled
 * (csrc/math_fast.cuh, csrc/glibc_log.cuh): draws `samples` random operands in the safe range on the device and
 * counts those where the fast routine differs from the general one.  which: 0 float division vs div.rn,
 * 1 logf fast vs general entry, 2 double division vs div.rn, 3 log fast vs general entry. */
ps_status ps_selftest_math(ps_ctx* ctx, int which, long long samples, unsigned long long seed,
                           long long* mismatches);
/* The two hardware claims the screened kernel's error budget rests on, checked over EVERY float of the safe range:
 * out[0] = max |x*rcp.approx(x) - 1| / 2^-24,  out[1] = max |lg2.approx(x) - log2 x| / (2^-24 max(1,|log2 x|)).
 * csrc/screen.cuh assumes both <= 4. */
ps_status ps_selftest_mufu(ps_ctx* ctx, double out[2]);
/* Device log()/logf() (the restatement of glibc's routines the Poisson scores use) evaluated on n host
 * arguments; y_general takes the entry point that handles every input, y_fast the safe-range one. */
ps_status ps_eval_log_f64(ps_ctx* ctx, long long n, const double* x, double* y_general, double* y_fast);
ps_status ps_eval_log_f32(ps_ctx* ctx, long long n, const float* x, float* y_general, float* y_fast);

/* ---- row-sharded single instance inside the library -----------------------------------------------------
 * One very large instance (BASELINE configs[4]: n = 1e6, T = 32) solved by `world` GPUs, one host process OR thread per
 * GPU, each with its own ps_ctx.  The reference has no counterpart (its solver is one process; the nearest thing is the pool
 * dispatch of independent solves, python_dpsolver.cpp:263-316).  Rank r owns the 128-row blocks rb with rb % world == r
 * (equal rows and equal cells of the triangle per rank); per level each rank computes its rows and the library issues ONE
 * ncclAllGather of the packed {maxScore_, nextStart_} chunk on the context's stream, then scatters it to row order -- no
 * host synchronisation between levels.  Sort, scan and the backtrace are replicated (level T, row 0 only, is rank 0's and
 * travels like the others), so EVERY rank returns the full result (same bytes as ps_solve_* on one GPU: sort index,
 * boundaries, subset scores, and the maxScore_ / nextStart_ tables if asked for; tests/test_gpu_shard.py).  All level
 * kernels take part: the exact-sum and running-sum screens and the exhaustive kernel.  NCCL is bound at run
 * time (dlopen: the copy already in the process, $PS_NCCL_LIB, the system's libnccl.so.2) and only by these calls.
 *   rank 0:     ps_comm_unique_id(id)            -> send the PS_COMM_ID_BYTES to every rank by any means (MPI, a socket,
 *                                                   torch.distributed, a shared variable between threads)
 *   every rank: ps_comm_create(ctx, id, rank, world, &comm)         (collective: ncclCommInitRank)
 *               ps_solve_sharded_f64(ctx, comm, &p, a, b, &r)        (collective; p.batch must be 1; same a, b on every rank)
 *               ps_comm_destroy(comm)                                                                                    */
#define PS_COMM_ID_BYTES 128
typedef struct ps_comm ps_comm;
ps_status ps_comm_unique_id(void* id_out /* PS_COMM_ID_BYTES */);
ps_status ps_comm_create(ps_ctx* ctx, const void* id_bytes, int rank, int world, ps_comm** out);
void ps_comm_destroy(ps_comm* comm);
int ps_comm_rank(const ps_comm* comm);
int ps_comm_world(const ps_comm* comm);
int ps_nccl_version(void); /* 0 if NCCL could not be loaded */
ps_status ps_solve_sharded_f32(ps_ctx* ctx, ps_comm* comm, const ps_problem* p, const float* a, const float* b,
                               ps_result* r);
ps_status ps_solve_sharded_f64(ps_ctx* ctx, ps_comm* comm, const ps_problem* p, const double* a, const double* b,
                               ps_result* r);

/* ---- row-sharded single instance: the same thing as building blocks driven by a multi-process host (kept for hosts that
 * bring their own collective; partitionsolvers_b200/multigpu.ShardedSolver, tests/test_dist_gloo.py) ----
 * Rank `rank` of `world` owns row blocks rb with rb % world == rank.  The caller (one process per GPU)
 * holds the level vector, calls ps_shard_level_* for its rows, and exchanges the level vector with an
 * all-gather of its own (torch.distributed / NCCL).  All pointers are DEVICE pointers.            */
typedef struct ps_shard ps_shard; /* opaque per-rank state */
ps_status ps_shard_create_f64(ps_ctx* ctx, const ps_problem* p, const double* a_sorted_dev,
                              const double* b_sorted_dev, int rank, int world, ps_shard** out);
ps_status ps_shard_create_f32(ps_ctx* ctx, const ps_problem* p, const float* a_sorted_dev,
                              const float* b_sorted_dev, int rank, int world, ps_shard** out);
/* rows owned by this rank at any level, padded: the all-gather chunk length (same on every rank) */
int ps_shard_chunk_rows(const ps_shard* s);
/* level 1: writes this rank's chunk of maxScore_[1] (owned rows, block-cyclic order) */
ps_status ps_shard_level1(ps_shard* s, void* my_chunk_dev);
/* gathered [world][chunk] -> natural order level vector inside the shard state */
ps_status ps_shard_set_prev(ps_shard* s, const void* gathered_dev);
/* level j>=2: computes owned rows from the previous level vector; writes value chunk and keeps splits */
ps_status ps_shard_level(ps_shard* s, int j, void* my_chunk_dev);
/* split index next_start[j][i] (host lookup of one element; valid on the owning rank, else -2) */
ps_status ps_shard_next_start(ps_shard* s, int j, int i, int* out);
int ps_shard_owner(const ps_shard* s, int i);
/* scores of the `count` subsets [bounds[t], bounds[t+1]) of the sorted sequence (DP_impl.hpp:131-139: score_by_subset);
 * HOST pointers; every rank holds the sorted inputs, any rank can answer */
ps_status ps_shard_subset_scores(ps_shard* s, int count, const int* bounds, void* scores);
/* 1 if this shard's levels run the screened kernel (ps_problem.screen, csrc/screen.cuh) */
int ps_shard_screened(const ps_shard* s);
void ps_shard_destroy(ps_shard* s);

#ifdef __cplusplus
}
#endif
#endif /* PS_B200_H */
