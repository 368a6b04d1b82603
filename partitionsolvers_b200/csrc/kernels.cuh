// kernels.cuh -- device side of the consecutive-partitions DP for sm_100a.
//
// Data layout in HBM (per instance; instances of a batch are strided copies):
//   rec[2][n+1]   level records {x, y, p, -}: one 16 B (float) / 32 B (double) record per k so that a
//                 k-step of the inner loop is ONE broadcast shared-memory load.  x,y are constant
//                 across levels (RUNNING: a[k-1], b[k-1];  PREFIX: scan values Pa[k], Pb[k]);
//                 p = maxScore_[j-1][k] is rewritten by every level into the other buffer.
//   SA,SB[NCH][n] RUNNING mode only: sums over [i, c*L) for every chunk boundary c*L > i, produced by
//                 the level-1 pass; lets any (row block, k-chunk) tile start mid-row while keeping the
//                 reference's left-to-right summation order (score_impl.hpp:248-256).
//   part[NCH][n]  per-chunk (best value, argmax k) of the level in flight.
//   ms, ns[T+1][n]  maxScore_ / nextStart_ (DP.hpp:162-163); ns drives the backtrace.
//
// Work decomposition of one level (DP_impl.hpp:101-118, the O(n^2) part): the upper-triangular
// (i,k) space is cut into tiles of BR rows x L columns.  One CTA = one tile, one thread = one row;
// the k loop streams records through a double-buffered shared-memory stage, so the three per-k
// operands are loaded from L2 once per CTA and broadcast to all rows.  Tiles are launched highest rows
// first (decode_tile).  A second, tiny
// kernel folds the per-chunk partials in ascending k, which preserves the reference's
// "first maximum wins" rule (strict >, DP_impl.hpp:107).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "score.cuh"

namespace ps {

template <typename D>
struct Rec4;
template <>
struct __align__(16) Rec4<float> {
  float x, y, p, w;
};
template <>
struct __align__(32) Rec4<double> {
  double x, y, p, w;
};

struct Geom {
  int n;        // ground-set size
  int nrows;    // rows computed at this level
  int kmax;     // largest k considered (n - j + 1)
  int L;        // k-chunk length, multiple of BR
  int NCH;      // chunks covering 1..kmax
  int G;        // L / BR
  int NRB;      // row blocks covering nrows
  int ntiles;
  int nch_alloc;          // chunk stride of part/SA arrays
  long long rec_stride;   // records per instance (n+1)
  // row sharding (single huge instance over several GPUs): this rank owns row blocks
  // rb_first + rb_step * m, m = 0..NRB-1 (NRB counts OWNED blocks).  Unsharded: 0, 1.
  int rb_first, rb_step;
  const int2* tile_map;   // optional table tile -> (rb, c) in launch order; nullptr = decode arithmetically
  // SA/SB checkpoints may be kept at a finer spacing than the k-chunks (screen_run.cuh, float inputs): the sums at chunk
  // boundary c L live in row c * sa_mul of arrays with sa_nch rows per instance.  Default: 1, nch_alloc.
  int sa_mul, sa_nch;
  // live-tile kernels (screen.cuh): partial slots per k-chunk (an expensive tile is split into up to nsub parts); default 1
  int nsub;
};

__host__ __device__ inline int geom_count_tiles(int NCH, int G, int NRB, int rb_first, int rb_step) {
  // every owned row block rb has one tile per chunk from its diagonal chunk rb/G to NCH-1
  long long t = 0;
  for (int m = 0; m < NRB; ++m) t += NCH - (rb_first + rb_step * m) / G;
  return (int)t;
}

// Launch order of the tiles of a level: HIGHEST rows first.  Rows are sorted by priority a/b, so the high rows
// are the ones whose range sums are past the A = B frontier -- the cells that actually evaluate division and log
// in the multiple-clustering scores -- and a 128 x L tile of such cells runs for a fifth of the whole launch when
// it shares an SM with eight others.  Starting the expensive tiles first and letting the cheap low rows fill the
// tail keeps the SMs level (first-fit in decreasing cost); with uniform cost (risk partitioning) the order is
// harmless.  Tile t -> (row block rb, chunk c), c from the row block's diagonal chunk rb/G up to NCH-1.
__device__ __forceinline__ void decode_tile(const Geom& g, int t, int& rb, int& c) {
  if (g.tile_map) {  // unsharded runs: the (rb, c) of every tile, tabulated once per geometry (build_tile_map_kernel)
    const int2 m = g.tile_map[t];
    rb = m.x;
    c = m.y;
    return;
  }
  if (g.rb_step == 1 && g.rb_first == 0) {
    int s = (g.NRB - 1) / g.G;  // super-block = G row blocks sharing a diagonal chunk
    for (;;) {
      const int rbs = min(g.G, g.NRB - s * g.G);
      const int cnt = rbs * (g.NCH - s);
      if (t < cnt) {
        rb = s * g.G + t % rbs;
        c = s + t / rbs;
        return;
      }
      t -= cnt;
      --s;
    }
  }
  for (int m = g.NRB - 1;; --m) {  // row-sharded run: owned blocks rb_first + rb_step * m
    rb = g.rb_first + g.rb_step * m;
    const int cnt = g.NCH - rb / g.G;
    if (t < cnt) {
      c = rb / g.G + t;
      return;
    }
    t -= cnt;
  }
}

// The decode loop walks up to NRB/G super-blocks; with 30k short tiles per level that is a few per cent of a level,
// so unsharded runs tabulate it once per geometry.
__global__ void build_tile_map_kernel(Geom g, int2* __restrict__ map) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ntiles) return;
  g.tile_map = nullptr;
  int rb, c;
  decode_tile(g, t, rb, c);
  map[t] = make_int2(rb, c);
}

// ---------------------------------------------------------------------------------------------
// One (row block, k-chunk) tile of one level.
// ---------------------------------------------------------------------------------------------
template <typename D, int DIST, bool RP, bool RUNNING, bool FAST, int BR>
__global__ void __launch_bounds__(BR)
level_tile_kernel(Geom g, const Rec4<D>* __restrict__ rec, const D* __restrict__ SA,
                  const D* __restrict__ SB, D* __restrict__ part_val, int* __restrict__ part_arg) {
  constexpr int BK = BR;
  __shared__ Rec4<D> sbuf[2][BK];
  // double Poisson on the fast path: a shared-memory copy of glibc's 128-entry log table (2 KB), so the
  // per-cell lookup is one LDS at a 32-bit offset instead of a 64-bit address computation + global load
  constexpr bool kLogTab = (sizeof(D) == 8) && (DIST == DIST_POISSON) && FAST;
  __shared__ GlogEntry s_log[kLogTab ? 128 : 1];
  if (kLogTab) {
    for (int t = threadIdx.x; t < 128; t += BR) s_log[t] = g_glog_tab[t];
  }

  const int inst = blockIdx.y;
  rec += (long long)inst * g.rec_stride;
  const long long pstride = (long long)g.nch_alloc * g.n;
  part_val += inst * pstride;
  part_arg += inst * pstride;
  if (RUNNING) {
    SA += inst * pstride;
    SB += inst * pstride;
  }

  // ---- tile decode
  int rb, c;
  decode_tile(g, blockIdx.x, rb, c);

  const int tid = threadIdx.x;
  const int i0 = rb * BR;
  const int i = i0 + tid;
  const int k_lo = max(c * g.L + 1, i0 + 1);
  const int k_hi = min((c + 1) * g.L, g.kmax);
  const bool diag = (c == rb / g.G);

  D A = 0, B = 0, Xi = 0, Yi = 0;
  {
    const int il = min(i, g.n - 1);
    if (RUNNING) {
      if (!diag) {
        A = SA[(long long)c * g.n + il];
        B = SB[(long long)c * g.n + il];
      }
    } else {
      Xi = rec[il].x;
      Yi = rec[il].y;
    }
  }
  D best = Lim<D>::lowest();
  int arg = -1;
  auto score_of = [&](D a_sum, D b_sum) -> D {
    // The multiple-clustering Poisson and Gaussian scores are 0 unless A > B (score_impl.hpp:385-388, 429-431).
    // The reference's ternary skips the division and the log there, and so does this branch: lanes are adjacent
    // rows at the same k, so a warp is uniformly on one side of the A = B frontier except along the frontier itself.
    if constexpr (!RP && DIST != DIST_RATIONAL) {
      if (!gt_ab<FAST>(a_sum, b_sum)) return (D)0;
      if constexpr (kLogTab)
        return ScoreFn<D, DIST, RP, FAST>::eval_gt(a_sum, b_sum, s_log);
      else
        return ScoreFn<D, DIST, RP, FAST>::eval_gt(a_sum, b_sum);
    } else if constexpr (kLogTab) {
      return ScoreFn<D, DIST, RP, FAST>::eval(a_sum, b_sum, s_log);
    } else {
      return ScoreFn<D, DIST, RP, FAST>::eval(a_sum, b_sum);
    }
  };

  const int nk = k_hi - k_lo + 1;
  const int nstages = (nk + BK - 1) / BK;
  {
    int k = k_lo + tid;
    Rec4<D> r;
    r.x = r.y = r.p = r.w = 0;
    if (k <= k_hi) r = rec[k];
    sbuf[0][tid] = r;
  }
  for (int s = 0; s < nstages; ++s) {
    __syncthreads();
    const int kb = k_lo + s * BK;
    const int kn = kb + BK + tid;
    Rec4<D> nx;
    nx.x = nx.y = nx.p = nx.w = 0;
    const bool has_next = (s + 1 < nstages);
    if (has_next && kn <= k_hi) nx = rec[kn];
    const int cnt = min(BK, k_hi - kb + 1);
    const Rec4<D>* __restrict__ sb = sbuf[s & 1];
    if (kb > i0 + BR - 1) {
      // interior: every row of the block sees every k of the stage
      if (cnt == BK) {
#pragma unroll 4
        for (int kk = 0; kk < BK; ++kk) {
          const Rec4<D> r = sb[kk];
          if (RUNNING) {
            A = add_rn<D>(A, r.x);
            B = add_rn<D>(B, r.y);
          } else {
            A = sub_rn<D>(r.x, Xi);
            B = sub_rn<D>(r.y, Yi);
          }
          const D v = add_rn<D>(score_of(A, B), r.p);
          if (v > best) {
            best = v;
            arg = kb + kk;
          }
        }
      } else {
        for (int kk = 0; kk < cnt; ++kk) {
          const Rec4<D> r = sb[kk];
          if (RUNNING) {
            A = add_rn<D>(A, r.x);
            B = add_rn<D>(B, r.y);
          } else {
            A = sub_rn<D>(r.x, Xi);
            B = sub_rn<D>(r.y, Yi);
          }
          const D v = add_rn<D>(score_of(A, B), r.p);
          if (v > best) {
            best = v;
            arg = kb + kk;
          }
        }
      }
    } else {
      // diagonal stage: row i only sees k > i
      for (int kk = 0; kk < cnt; ++kk) {
        const int k = kb + kk;
        const Rec4<D> r = sb[kk];
        if (k > i) {
          if (RUNNING) {
            A = add_rn<D>(A, r.x);
            B = add_rn<D>(B, r.y);
          } else {
            A = sub_rn<D>(r.x, Xi);
            B = sub_rn<D>(r.y, Yi);
          }
          const D v = add_rn<D>(score_of(A, B), r.p);
          if (v > best) {
            best = v;
            arg = k;
          }
        }
      }
    }
    if (has_next) sbuf[(s + 1) & 1][tid] = nx;
  }
  if (i < g.nrows) {
    part_val[(long long)c * g.n + i] = best;
    part_arg[(long long)c * g.n + i] = arg;
  }
}

// ---------------------------------------------------------------------------------------------
// Poisson, float, RUNNING sums, safe-range inputs: TWO rows per thread (i and i+128) so that the whole
// cell pipeline runs on packed f32x2 instructions (FADD2/FMUL2/FFMA2) -- the kernel is issue-bound, and
// a packed instruction retires two IEEE operations per issue slot (operand negation, immediates and
// scalar broadcast are free modifiers of the packed forms).  Lane-wise the arithmetic is the scalar
// kernel's, bit for bit.  Tile = 256 rows x L columns.
// ---------------------------------------------------------------------------------------------
template <bool RP>
__global__ void __launch_bounds__(128)
level_tile_kernel_pois_f32x2(Geom g, const Rec4<float>* __restrict__ rec, const float* __restrict__ SA,
                             const float* __restrict__ SB, float* __restrict__ part_val,
                             int* __restrict__ part_arg) {
  constexpr int BK = 128, BRT = 256;
  __shared__ Rec4<float> sbuf[2][BK];
  __shared__ GlogEntry s_logf[16];
  if (threadIdx.x < 16) s_logf[threadIdx.x] = g_glogf_tab[threadIdx.x];

  const int inst = blockIdx.y;
  rec += (long long)inst * g.rec_stride;
  const long long pstride = (long long)g.nch_alloc * g.n;
  part_val += inst * pstride;
  part_arg += inst * pstride;
  SA += inst * pstride;
  SB += inst * pstride;

  int rb, c;
  decode_tile(g, blockIdx.x, rb, c);
  const int tid = threadIdx.x;
  const int i0 = rb * BRT;
  const int ia = i0 + tid, ib = i0 + 128 + tid;
  const int k_lo = max(c * g.L + 1, i0 + 1);
  const int k_hi = min((c + 1) * g.L, g.kmax);
  const bool diag = (c == rb / g.G);

  float2 A = make_float2(0.f, 0.f), B = A;
  if (!diag) {
    const int la = min(ia, g.n - 1), lb = min(ib, g.n - 1);
    A = make_float2(SA[(long long)c * g.n + la], SA[(long long)c * g.n + lb]);
    B = make_float2(SB[(long long)c * g.n + la], SB[(long long)c * g.n + lb]);
  }
  float best0 = Lim<float>::lowest(), best1 = best0;
  int arg0 = -1, arg1 = -1;

  const int nk = k_hi - k_lo + 1;
  const int nstages = (nk + BK - 1) / BK;
  {
    Rec4<float> r;
    r.x = r.y = r.p = r.w = 0;
    if (k_lo + tid <= k_hi) r = rec[k_lo + tid];
    sbuf[0][tid] = r;
  }
  for (int s = 0; s < nstages; ++s) {
    __syncthreads();
    const int kb = k_lo + s * BK;
    const int kn = kb + BK + tid;
    Rec4<float> nx;
    nx.x = nx.y = nx.p = nx.w = 0;
    const bool has_next = (s + 1 < nstages);
    if (has_next && kn <= k_hi) nx = rec[kn];
    const int cnt = min(BK, k_hi - kb + 1);
    const Rec4<float>* __restrict__ sb = sbuf[s & 1];
    if (kb > i0 + BRT - 1 && cnt == BK) {
#pragma unroll 2
      for (int kk = 0; kk < BK; ++kk) {
        const Rec4<float> r = sb[kk];
        A = __fadd2_rn(A, make_float2(r.x, r.x));
        B = __fadd2_rn(B, make_float2(r.y, r.y));
        float2 sc = make_float2(0.f, 0.f);
        // mult-clust: the score is 0 unless A > B (score_impl.hpp:385-388); like the reference's ternary, skip the
        // division and the log when neither of the thread's two rows is past the A = B frontier
        if (RP || (A.x > B.x) || (A.y > B.y)) {
          const float2 q = fast_fdiv2(A, B);
          const float2 l = make_float2(glibc_logf_fast(q.x, s_logf), glibc_logf_fast(q.y, s_logf));
          // A*l with two scalar multiplies: ptxas (12.9) contracts mul.rn.f32x2 (and fma.rn.f32x2 with a -0
          // addend) followed by add.rn.f32x2 into one FFMA2 despite the explicit rounding modifiers and
          // --fmad=false, which would skip the reference's intermediate rounding of the product
          sc = make_float2(__fmul_rn(A.x, l.x), __fmul_rn(A.y, l.y));
          if (!RP) {
            sc = __fadd2_rn(sc, B);
            sc = __fadd2_rn(sc, make_float2(-A.x, -A.y));
            sc.x = (A.x > B.x) ? sc.x : 0.f;
            sc.y = (A.y > B.y) ? sc.y : 0.f;
          }
        }
        const float2 v = __fadd2_rn(sc, make_float2(r.p, r.p));
        if (v.x > best0) {
          best0 = v.x;
          arg0 = kb + kk;
        }
        if (v.y > best1) {
          best1 = v.y;
          arg1 = kb + kk;
        }
      }
    } else {
      // diagonal or ragged stage: scalar, masked (row i only sees k > i)
      for (int kk = 0; kk < cnt; ++kk) {
        const int k = kb + kk;
        const Rec4<float> r = sb[kk];
        if (k > ia) {
          A.x = __fadd_rn(A.x, r.x);
          B.x = __fadd_rn(B.x, r.y);
          const float v = __fadd_rn(ScoreFn<float, DIST_POISSON, RP, true>::eval(A.x, B.x), r.p);
          if (v > best0) {
            best0 = v;
            arg0 = k;
          }
        }
        if (k > ib) {
          A.y = __fadd_rn(A.y, r.x);
          B.y = __fadd_rn(B.y, r.y);
          const float v = __fadd_rn(ScoreFn<float, DIST_POISSON, RP, true>::eval(A.y, B.y), r.p);
          if (v > best1) {
            best1 = v;
            arg1 = k;
          }
        }
      }
    }
    if (has_next) sbuf[(s + 1) & 1][tid] = nx;
  }
  if (ia < g.nrows) {
    part_val[(long long)c * g.n + ia] = best0;
    part_arg[(long long)c * g.n + ia] = arg0;
  }
  if (ib < g.nrows) {
    part_val[(long long)c * g.n + ib] = best1;
    part_arg[(long long)c * g.n + ib] = arg1;
  }
}

// ---------------------------------------------------------------------------------------------
// Level T needs row 0 only (DP_impl.hpp:114-116).  One thread per row would leave a single lane walking a
// whole chunk with nothing to hide its latency, so this kernel turns the k axis into the parallel one: per
// sub-block of 1024 k's, thread 0 forms the running sums left to right (the reference's order; two dependent adds
// per k), all threads then evaluate their cells, and a shuffle/shared-memory reduction keeps the largest value and,
// among equals, the smallest k -- the reference's first-maximum rule.  One CTA per k-chunk; the result goes through
// the same per-chunk partials and combine_kernel as every other level.
// ---------------------------------------------------------------------------------------------
template <typename D, int DIST, bool RP, bool RUNNING, bool FAST, int NT>
__global__ void __launch_bounds__(NT)
last_level_kernel(Geom g, const Rec4<D>* __restrict__ rec, const D* __restrict__ SA, const D* __restrict__ SB,
                  D* __restrict__ part_val, int* __restrict__ part_arg) {
  constexpr int SUB = 1024;
  __shared__ D sA[SUB], sB[SUB];
  __shared__ D red_v[NT / 32];
  __shared__ int red_k[NT / 32];
  __shared__ D carry[2];
  const int inst = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  rec += (long long)inst * g.rec_stride;
  const long long pstride = (long long)g.nch_alloc * g.n;
  const int k_lo = c * g.L + 1, k_hi = min((c + 1) * g.L, g.kmax);
  const D X0 = rec[0].x, Y0 = rec[0].y;  // PREFIX: scan values at row 0 (= 0)
  if (tid == 0) {
    const long long so = inst * (long long)g.sa_nch * g.n + (long long)c * g.sa_mul * g.n;
    carry[0] = (RUNNING && c > 0) ? SA[so] : (D)0;
    carry[1] = (RUNNING && c > 0) ? SB[so] : (D)0;
  }
  D best = Lim<D>::lowest();
  int arg = -1;
  for (int kb = k_lo; kb <= k_hi; kb += SUB) {
    const int cnt = min(SUB, k_hi - kb + 1);
    __syncthreads();
    if (RUNNING) {
      if (tid == 0) {
        D A = carry[0], B = carry[1];
        for (int q = 0; q < cnt; ++q) {
          const Rec4<D> r = rec[kb + q];
          A = add_rn<D>(A, r.x);
          B = add_rn<D>(B, r.y);
          sA[q] = A;
          sB[q] = B;
        }
        carry[0] = A;
        carry[1] = B;
      }
      __syncthreads();
    }
    for (int q = tid; q < cnt; q += NT) {  // ascending k per thread
      const int k = kb + q;
      D A, B;
      if (RUNNING) {
        A = sA[q];
        B = sB[q];
      } else {
        A = sub_rn<D>(rec[k].x, X0);
        B = sub_rn<D>(rec[k].y, Y0);
      }
      const D v = add_rn<D>(ScoreFn<D, DIST, RP, FAST>::eval(A, B), rec[k].p);
      if (v > best) {
        best = v;
        arg = k;
      }
    }
  }
  // largest value wins; among equal values the smallest k (arg = -1 means "nothing won" and never beats a k)
  auto better = [](D v1, int k1, D v2, int k2) {
    if (k1 < 0) return false;
    if (k2 < 0) return true;
    return v1 > v2 || (v1 == v2 && k1 < k2);
  };
  for (int o = 16; o > 0; o >>= 1) {
    const D ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, arg, o);
    if (better(ov, ok, best, arg)) {
      best = ov;
      arg = ok;
    }
  }
  if ((tid & 31) == 0) {
    red_v[tid >> 5] = best;
    red_k[tid >> 5] = arg;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NT / 32; ++w)
      if (better(red_v[w], red_k[w], best, arg)) {
        best = red_v[w];
        arg = red_k[w];
      }
    part_val[inst * pstride + (long long)c * g.n] = (arg < 0) ? Lim<D>::lowest() : best;
    part_arg[inst * pstride + (long long)c * g.n] = arg;
  }
}

// Screened levels (screen.cuh): pm = a float that dominates p (the level value the next level reads) together with
// every rounding that involves it in the float bound, and the round-down conversion used for thresholds.
template <typename D>
__device__ __forceinline__ float screen_pm(D p) {
  const float f = (sizeof(D) == 8) ? __double2float_ru((double)p) : (float)p;
  return __fmaf_ru(fabsf(f), 9.5367431640625e-07f, f);  // + 2^-20 |p|
}
template <typename D>
__device__ __forceinline__ float screen_rd(D v) {
  return (sizeof(D) == 8) ? __double2float_rd((double)v) : (float)v;
}

// Fold the per-chunk partials of row i in ascending chunk order (= ascending k): strict > keeps the
// first maximum, exactly like the reference's scan over k.  Writes maxScore_[j][i], nextStart_[j][i]
// and the p field of the NEXT level's records.  Sharded runs also emit the value in block-cyclic
// "chunk" order for the all-gather.
template <typename D, int BR, int CW = 1>
__global__ void combine_kernel(Geom g, const D* __restrict__ part_val, const int* __restrict__ part_arg,
                               D* __restrict__ ms_row, int* __restrict__ ns_row, long long table_stride,
                               Rec4<D>* __restrict__ rec_next, D* __restrict__ shard_chunk,
                               float* __restrict__ pm_next, const int* __restrict__ tile_stamp = nullptr,
                               int stamp = 0, int stamp_stride = 0, int* __restrict__ shard_arg = nullptr) {
  // CW lanes share a row: lane q folds the q-th part of the row's chunks (contiguous, ascending), then the parts are merged
  // in ascending order with the same strict > -- the first maximum still wins.  A row's fold is a chain of dependent L2
  // round trips; with many chunks (large n) splitting it four ways shortens the chain and quadruples the loads in flight;
  // instances with few chunks (replica batches) take CW = 1.
  const int inst = blockIdx.y;
  const long long pstride = (long long)g.nch_alloc * g.n;
  part_val += inst * pstride;
  part_arg += inst * pstride;
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = gt / CW, q = gt % CW;  // index among owned rows, quarter
  int i;
  bool row_ok = true;
  if (g.rb_step == 1 && g.rb_first == 0) {
    i = m;  // unsharded: rows are contiguous whatever the tile height of the level kernel
    row_ok = i < g.nrows;
  } else {
    row_ok = m < g.NRB * BR;
    const int rb = g.rb_first + g.rb_step * (m / BR);
    i = rb * BR + (m % BR);
  }
  D best = Lim<D>::lowest();
  int cbest = -1;
  if (row_ok && i < g.nrows) {
    const int c_first = i / g.L;
    const int per = (g.NCH - c_first + CW - 1) / CW;
    const int c0 = c_first + q * per, c1 = min(g.NCH, c0 + per);
    const int* __restrict__ stamps =
        tile_stamp ? tile_stamp + (long long)inst * g.nch_alloc * stamp_stride + i / BR : nullptr;
    for (int c = c0; c < c1; c += 4) {
      bool live[4];
      D v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // screened levels: a tile that was dropped whole wrote nothing and left its stamp alone
        live[u] = (c + u < c1) && (!stamps || stamps[(long long)(c + u) * stamp_stride] == stamp);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = live[u] ? part_val[(long long)(c + u) * g.n + i] : Lim<D>::lowest();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v[u] > best) {
          best = v[u];
          cbest = c + u;
        }
    }
  }
  // ordered merge over the CW lanes of the row (lanes of a row are adjacent; all 32 lanes take part in the shuffles)
  const unsigned int lane = threadIdx.x & 31u, base = lane & ~(unsigned)(CW - 1);
  D mb = __shfl_sync(0xffffffffu, best, base);
  int mc = __shfl_sync(0xffffffffu, cbest, base);
#pragma unroll
  for (int u = 1; u < CW; ++u) {
    const D ob = __shfl_sync(0xffffffffu, best, base + u);
    const int oc = __shfl_sync(0xffffffffu, cbest, base + u);
    if (ob > mb) {
      mb = ob;
      mc = oc;
    }
  }
  if (q != 0 || !row_ok) return;
  int arg = -1;
  if (i < g.nrows) {
    // the split index is read once, for the winning chunk only (a partial equal to lowest() never wins: split -1)
    arg = (mc >= 0) ? part_arg[(long long)mc * g.n + i] : -1;
    if (ms_row) ms_row[inst * table_stride + i] = mb;
    if (ns_row) ns_row[inst * table_stride + i] = arg;
    if (rec_next) rec_next[inst * g.rec_stride + i].p = mb;
    if (pm_next) pm_next[(inst * g.rec_stride + i) * 4 + 3] = screen_pm<D>(mb);  // .w of the float record
  }
  if (shard_chunk) shard_chunk[m] = (i < g.nrows) ? mb : Lim<D>::lowest();
  if (shard_arg) shard_arg[m] = arg;
}

// ---------------------------------------------------------------------------------------------
// Level 1 (DP_impl.hpp:90-95): maxScore_[1][i] = S(i,n), nextStart_[1][i] = n.
// RUNNING: one thread per row walks k = i+1..n accumulating the sums the way the reference does and
// drops the running sums at every chunk boundary into SA/SB.  2 adds per cell, once per solve.
// ---------------------------------------------------------------------------------------------
template <typename D, int BR>
__global__ void __launch_bounds__(BR)
level1_running_kernel(int n, int L, int nch_alloc, long long rec_stride, int dist, int rp,
                      const Rec4<D>* __restrict__ rec, D* __restrict__ SA, D* __restrict__ SB,
                      D* __restrict__ ms1, int* __restrict__ ns1, long long table_stride,
                      Rec4<D>* __restrict__ rec_next, int rb_first, int rb_step, D* __restrict__ shard_chunk,
                      int nblocks) {
  constexpr int BK = BR;
  __shared__ Rec4<D> sbuf[2][BK];
  const int inst = blockIdx.y;
  rec += (long long)inst * rec_stride;
  const long long pstride = (long long)nch_alloc * n;
  const int tid = threadIdx.x;
  // The rows form a triangle: block m walks n - 128 m cells per row.  A CTA takes block m and then block nblocks-1-m, so
  // every CTA walks the same number of cells and the SMs finish together (grid = ceil(nblocks / 2)).
  for (int pass = 0; pass < 2; ++pass) {
  const int m = pass ? nblocks - 1 - (int)blockIdx.x : (int)blockIdx.x;
  if (pass && m <= (int)blockIdx.x) break;
  if (pass) __syncthreads();
  const int rb = rb_first + rb_step * m;
  const int i0 = rb * BR;
  const int i = i0 + tid;
  D A = 0, B = 0;
  if (i0 < n) {
    const int k_lo = i0 + 1, k_hi = n;
    const int nstages = (k_hi - k_lo + 1 + BK - 1) / BK;
    {
      int k = k_lo + tid;
      Rec4<D> r;
      r.x = r.y = r.p = r.w = 0;
      if (k <= k_hi) r = rec[k];
      sbuf[0][tid] = r;
    }
    for (int s = 0; s < nstages; ++s) {
      __syncthreads();
      const int kb = k_lo + s * BK;
      const int kn = kb + BK + tid;
      Rec4<D> nx;
      nx.x = nx.y = nx.p = nx.w = 0;
      const bool has_next = (s + 1 < nstages);
      if (has_next && kn <= k_hi) nx = rec[kn];
      const int cnt = min(BK, k_hi - kb + 1);
      const Rec4<D>* __restrict__ sb = sbuf[s & 1];
      if (s == 0) {
        for (int kk = 0; kk < cnt; ++kk) {
          if (kb + kk > i) {
            A = add_rn<D>(A, sb[kk].x);
            B = add_rn<D>(B, sb[kk].y);
          }
        }
      } else if (cnt == BK) {
        // full stage: constant trip count.  The walk runs at ~2.7 warps per scheduler (one thread per row), too few to hide
        // the shared-memory latency behind other warps, so the records are read eight ahead of the two dependent add chains.
        constexpr int PF = 8;
        D xa[PF], xb[PF];
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          xa[q] = sb[q].x;
          xb[q] = sb[q].y;
        }
#pragma unroll
        for (int kk = 0; kk < BK; kk += PF) {
          D na[PF], nb[PF];
          if (kk + PF < BK) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
              na[q] = sb[kk + PF + q].x;
              nb[q] = sb[kk + PF + q].y;
            }
          }
#pragma unroll
          for (int q = 0; q < PF; ++q) {
            A = add_rn<D>(A, xa[q]);
            B = add_rn<D>(B, xb[q]);
          }
          if (kk + PF < BK) {
#pragma unroll
            for (int q = 0; q < PF; ++q) {
              xa[q] = na[q];
              xb[q] = nb[q];
            }
          }
        }
      } else {
        for (int kk = 0; kk < cnt; ++kk) {
          A = add_rn<D>(A, sb[kk].x);
          B = add_rn<D>(B, sb[kk].y);
        }
      }
      const int kend = kb + cnt - 1;  // last k of the stage
      if (cnt == BK && (kend % L) == 0 && i < n && SA) {
        const int c = kend / L;
        if (c < nch_alloc) {
          SA[inst * pstride + (long long)c * n + i] = A;
          SB[inst * pstride + (long long)c * n + i] = B;
        }
      }
      if (has_next) sbuf[(s + 1) & 1][tid] = nx;
    }
  }
  D v = Lim<D>::lowest();
  if (i < n) {
    v = score_dyn<D>(dist, rp, A, B);
    if (ms1) ms1[inst * table_stride + i] = v;
    if (ns1) ns1[inst * table_stride + i] = n;
    if (rec_next) rec_next[inst * rec_stride + i].p = v;
  }
  if (shard_chunk) shard_chunk[m * BR + tid] = v;
  }
}

// global -> shared copies that bypass the register file (cp.async, 16 bytes each)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned int sa = (unsigned int)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Level 1, RUNNING sums, small instances (ps_b200.cu: SMALL_ROWS = 32-row tiles, chunk length 32): one WARP per 32 rows.  The
// walk is one dependent add chain per row whatever the geometry; with 32-cell stages the 128-row kernel's scheme (next stage
// fetched into registers while the current one is walked) leaves ~300 cycles of work to hide a global load, so the records
// come through a four-deep cp.async ring instead (three stages = 96 cells ahead) and the only synchronisation is __syncwarp.
// Checkpoints at every multiple of 32 (L = 32).  Same sums, same order: every row adds its cells left to right.
template <typename D>
__global__ void __launch_bounds__(32)
level1_running_small_kernel(int n, int nch_alloc, long long rec_stride, int dist, int rp, const Rec4<D>* __restrict__ rec,
                            D* __restrict__ SA, D* __restrict__ SB, D* __restrict__ ms1, int* __restrict__ ns1,
                            long long table_stride, Rec4<D>* __restrict__ rec_next, int nblocks) {
  constexpr int BK = 32, NBUF = 4;
  __shared__ Rec4<D> sbuf[NBUF][BK];
  const int inst = blockIdx.y;
  rec += (long long)inst * rec_stride;
  const long long pstride = (long long)nch_alloc * n;
  const int tid = threadIdx.x;
  for (int pass = 0; pass < 2; ++pass) {  // block m, then block nblocks-1-m: every warp walks the same number of cells
    const int m = pass ? nblocks - 1 - (int)blockIdx.x : (int)blockIdx.x;
    if (pass && m <= (int)blockIdx.x) break;
    __syncwarp();
    const int i0 = m * BK;
    const int i = i0 + tid;
    D A = 0, B = 0;
    if (i0 < n) {
      const int k_lo = i0 + 1, k_hi = n;
      const int nstages = (k_hi - k_lo + 1 + BK - 1) / BK;
      auto issue = [&](const int s) {
        if (s < nstages) {
          const int k = min(k_lo + s * BK + tid, k_hi);  // cells past k_hi are never consumed
          cp_async16(&sbuf[s % NBUF][tid], &rec[k]);
          if (sizeof(D) == 8)
            cp_async16(reinterpret_cast<char*>(&sbuf[s % NBUF][tid]) + 16, reinterpret_cast<const char*>(&rec[k]) + 16);
        }
        cp_async_commit();  // an empty group keeps the count uniform
      };
      for (int s = 0; s < NBUF - 1; ++s) issue(s);
      for (int s = 0; s < nstages; ++s) {
        issue(s + NBUF - 1);
        cp_async_wait_group<NBUF - 1>();
        __syncwarp();
        const int kb = k_lo + s * BK;
        const int cnt = min(BK, k_hi - kb + 1);
        const Rec4<D>* __restrict__ sb = sbuf[s % NBUF];
        if (s == 0) {
          for (int kk = 0; kk < cnt; ++kk) {
            if (kb + kk > i) {
              A = add_rn<D>(A, sb[kk].x);
              B = add_rn<D>(B, sb[kk].y);
            }
          }
        } else if (cnt == BK) {
#pragma unroll
          for (int kk = 0; kk < BK; ++kk) {
            A = add_rn<D>(A, sb[kk].x);
            B = add_rn<D>(B, sb[kk].y);
          }
        } else {
          for (int kk = 0; kk < cnt; ++kk) {
            A = add_rn<D>(A, sb[kk].x);
            B = add_rn<D>(B, sb[kk].y);
          }
        }
        if (cnt == BK && i < n && SA) {  // kend = i0 + 32 (s + 1): a chunk boundary
          const int c = (kb + BK - 1) / BK;
          if (c < nch_alloc) {
            SA[inst * pstride + (long long)c * n + i] = A;
            SB[inst * pstride + (long long)c * n + i] = B;
          }
        }
        __syncwarp();  // the buffer of this stage is refilled by the next iteration's issue
      }
      cp_async_wait_group<0>();
    }
    if (i < n) {
      const D v = score_dyn<D>(dist, rp, A, B);
      ms1[inst * table_stride + i] = v;
      ns1[inst * table_stride + i] = n;
      rec_next[inst * rec_stride + i].p = v;
    }
  }
}

// Sharded runs, prefix records: level 1 of the owned rows, in block-cyclic chunk order.
template <typename D, int BR>
__global__ void level1_prefix_shard_kernel(int n, int dist, int rp, const Rec4<D>* __restrict__ rec,
                                           int* __restrict__ ns1, int rb_first, int rb_step,
                                           D* __restrict__ shard_chunk) {
  const int rb = rb_first + rb_step * blockIdx.x;
  const int i = rb * BR + threadIdx.x;
  D v = Lim<D>::lowest();
  if (i < n) {
    v = score_dyn<D>(dist, rp, sub_rn<D>(rec[n].x, rec[i].x), sub_rn<D>(rec[n].y, rec[i].y));
    ns1[i] = n;
  }
  shard_chunk[blockIdx.x * BR + threadIdx.x] = v;
}

template <typename D>
__global__ void level1_prefix_kernel(int n, long long rec_stride, int dist, int rp,
                                     const Rec4<D>* __restrict__ rec, D* __restrict__ ms1,
                                     int* __restrict__ ns1, long long table_stride,
                                     Rec4<D>* __restrict__ rec_next, float* __restrict__ pm_next) {
  const int inst = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rec += (long long)inst * rec_stride;
  const D A = sub_rn<D>(rec[n].x, rec[i].x);
  const D B = sub_rn<D>(rec[n].y, rec[i].y);
  const D v = score_dyn<D>(dist, rp, A, B);
  ms1[inst * table_stride + i] = v;
  ns1[inst * table_stride + i] = n;
  rec_next[inst * rec_stride + i].p = v;
  if (pm_next) pm_next[(inst * rec_stride + i) * 4 + 3] = screen_pm<D>(v);
}

// ---------------------------------------------------------------------------------------------
// Stage 1 (after the radix sort in radix_sort.cuh): gather, level records.
// ---------------------------------------------------------------------------------------------
template <typename D>
__global__ void gather_kernel(long long total, int n, const int* __restrict__ perm,
                              const D* __restrict__ a, const D* __restrict__ b, D* __restrict__ as,
                              D* __restrict__ bs) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const long long base = (g / n) * n;
  const int p = perm[g];
  as[g] = a[base + p];
  bs[g] = b[base + p];
}

// Safe-range precheck for the FAST kernels (math_fast.cuh): flags[0] |= 1 if some element falls outside
// "b in [2^-30, 2^30] and |a| in {0} U [2^-30, 2^30]"; flags[0] |= 2 if some a <= 0 (Poisson additionally
// needs strictly positive sums).  NaNs fail every comparison and set bit 0.
template <typename D>
__global__ void range_check_kernel(long long total, const D* __restrict__ a, const D* __restrict__ b,
                                   int* __restrict__ flags) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int f = 0;
  if (g < total) {
    const D lo = (D)9.313225746154785e-10, hi = (D)1073741824.0;  // 2^-30, 2^30
    const D av = a[g], bv = b[g];
    const D aa = av < 0 ? -av : av;
    if (!(bv >= lo && bv <= hi)) f |= 1;
    if (!(aa == 0 || (aa >= lo && aa <= hi))) f |= 1;
    if (!(av > 0)) f |= 2;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// RUNNING records: rec[k] = {a[k-1], b[k-1]}; both level buffers.
template <typename D>
__global__ void build_rec_running_kernel(int n, long long rec_stride, const D* __restrict__ as,
                                         const D* __restrict__ bs, Rec4<D>* __restrict__ rec0,
                                         Rec4<D>* __restrict__ rec1) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  Rec4<D> r;
  r.x = k ? as[(long long)inst * n + k - 1] : (D)0;
  r.y = k ? bs[(long long)inst * n + k - 1] : (D)0;
  r.p = 0;
  r.w = 0;
  rec0[inst * rec_stride + k] = r;
  rec1[inst * rec_stride + k] = r;
}

// PREFIX records: rec[k] = {sum a[0..k), sum b[0..k)}.  One CTA per instance: every thread owns a
// contiguous slice, slice totals are scanned in shared memory, slices are re-walked with their offset.
template <typename D, int NT>
__global__ void __launch_bounds__(NT)
build_rec_prefix_kernel(int n, long long rec_stride, const D* __restrict__ as, const D* __restrict__ bs,
                        Rec4<D>* __restrict__ rec0, Rec4<D>* __restrict__ rec1) {
  __shared__ D sa[NT], sb[NT];
  const int inst = blockIdx.x;
  as += (long long)inst * n;
  bs += (long long)inst * n;
  rec0 += inst * rec_stride;
  rec1 += inst * rec_stride;
  const int tid = threadIdx.x;
  const int per = (n + NT - 1) / NT;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  D ta = 0, tb = 0;
#pragma unroll 8
  for (int k = lo; k < hi; ++k) {
    ta = add_rn<D>(ta, as[k]);
    tb = add_rn<D>(tb, bs[k]);
  }
  sa[tid] = ta;
  sb[tid] = tb;
  __syncthreads();
  if (tid == 0) {  // NT partials: serial exclusive scan, left to right
    D ra = 0, rb = 0;
    for (int q = 0; q < NT; ++q) {
      D xa = sa[q], xb = sb[q];
      sa[q] = ra;
      sb[q] = rb;
      ra = add_rn<D>(ra, xa);
      rb = add_rn<D>(rb, xb);
    }
  }
  __syncthreads();
  D ra = sa[tid], rb = sb[tid];
  if (tid == 0) {
    Rec4<D> z;
    z.x = z.y = z.p = z.w = 0;
    rec0[0] = z;
    rec1[0] = z;
  }
#pragma unroll 4
  for (int k = lo; k < hi; ++k) {
    ra = add_rn<D>(ra, as[k]);
    rb = add_rn<D>(rb, bs[k]);
    Rec4<D> r;
    r.x = ra;
    r.y = rb;
    r.p = 0;
    r.w = 0;
    rec0[k + 1] = r;
    rec1[k + 1] = r;
  }
}

// PREFIX records of EXACTLY SUMMABLE inputs (the screened path, screen.cuh P2): every partial sum is representable, so any
// order of additions gives the reference's values and the scan can spread over the machine.  Three small kernels: sums of
// 1024-element blocks, exclusive scan of the block sums (one CTA per instance), block-local scan + offset.
constexpr int PFX_BLOCK = 1024;  // elements per CTA (256 threads x 4)
template <typename D>
__global__ void __launch_bounds__(256)
prefix_block_sums_kernel(int n, int nb, const D* __restrict__ as, const D* __restrict__ bs, D* __restrict__ pa,
                         D* __restrict__ pb) {
  __shared__ D wa[8], wb[8];
  const int inst = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x;
  as += (long long)inst * n;
  bs += (long long)inst * n;
  D ta = 0, tb = 0;
  const int k0 = blk * PFX_BLOCK + tid * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (k0 + q < n) {
      ta = add_rn<D>(ta, as[k0 + q]);
      tb = add_rn<D>(tb, bs[k0 + q]);
    }
  for (int o = 16; o > 0; o >>= 1) {
    ta = add_rn<D>(ta, __shfl_xor_sync(0xffffffffu, ta, o));
    tb = add_rn<D>(tb, __shfl_xor_sync(0xffffffffu, tb, o));
  }
  if ((tid & 31) == 0) {
    wa[tid >> 5] = ta;
    wb[tid >> 5] = tb;
  }
  __syncthreads();
  if (tid == 0) {
    D ra = 0, rb = 0;
    for (int w = 0; w < 8; ++w) {
      ra = add_rn<D>(ra, wa[w]);
      rb = add_rn<D>(rb, wb[w]);
    }
    pa[(long long)inst * nb + blk] = ra;
    pb[(long long)inst * nb + blk] = rb;
  }
}
template <typename D>
__global__ void prefix_scan_partials_kernel(int nb, D* __restrict__ pa, D* __restrict__ pb) {
  // exclusive scan of the block sums of one instance; nb is small (n / 1024): thread 0 walks a, thread 1 walks b
  if (threadIdx.x > 1) return;
  D* p = (threadIdx.x ? pb : pa) + (long long)blockIdx.x * nb;
  D run = 0;
  for (int q = 0; q < nb; ++q) {
    const D v = p[q];
    p[q] = run;
    run = add_rn<D>(run, v);
  }
}
template <typename D>
__global__ void __launch_bounds__(256)
prefix_write_kernel(int n, int nb, long long rec_stride, const D* __restrict__ as, const D* __restrict__ bs,
                    const D* __restrict__ pa, const D* __restrict__ pb, Rec4<D>* __restrict__ rec0,
                    Rec4<D>* __restrict__ rec1) {
  __shared__ D wa[8], wb[8];
  const int inst = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  as += (long long)inst * n;
  bs += (long long)inst * n;
  rec0 += inst * rec_stride;
  rec1 += inst * rec_stride;
  const int k0 = blk * PFX_BLOCK + tid * 4;
  D xa[4], xb[4];
  D ta = 0, tb = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    xa[q] = (k0 + q < n) ? as[k0 + q] : (D)0;
    xb[q] = (k0 + q < n) ? bs[k0 + q] : (D)0;
    ta = add_rn<D>(ta, xa[q]);
    tb = add_rn<D>(tb, xb[q]);
  }
  // inclusive scan of the thread totals inside the warp, then of the warp totals
  D sa = ta, sb = tb;
  for (int o = 1; o < 32; o <<= 1) {
    const D ua = __shfl_up_sync(0xffffffffu, sa, o), ub = __shfl_up_sync(0xffffffffu, sb, o);
    if (lane >= o) {
      sa = add_rn<D>(sa, ua);
      sb = add_rn<D>(sb, ub);
    }
  }
  if (lane == 31) {
    wa[wid] = sa;
    wb[wid] = sb;
  }
  __syncthreads();
  D oa = pa[(long long)inst * nb + blk], ob = pb[(long long)inst * nb + blk];
  for (int w = 0; w < wid; ++w) {
    oa = add_rn<D>(oa, wa[w]);
    ob = add_rn<D>(ob, wb[w]);
  }
  D ra = add_rn<D>(oa, sub_rn<D>(sa, ta)), rb = add_rn<D>(ob, sub_rn<D>(sb, tb));  // sums before this thread's elements
  if (blk == 0 && tid == 0) {
    Rec4<D> z;
    z.x = z.y = z.p = z.w = 0;
    rec0[0] = z;
    rec1[0] = z;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (k0 + q < n) {
      ra = add_rn<D>(ra, xa[q]);
      rb = add_rn<D>(rb, xb[q]);
      Rec4<D> r;
      r.x = ra;
      r.y = rb;
      r.p = 0;
      r.w = 0;
      rec0[k0 + q + 1] = r;
      rec1[k0 + q + 1] = r;
    }
  }
}

template <typename D>
__global__ void init_tables_kernel(long long total, int n, D* __restrict__ ms, int* __restrict__ ns) {
  // row 0 of every instance: 0 / -1; all other rows: lowest() / -1  (DP_impl.hpp:83-95)
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  (void)n;
  ms[g] = Lim<D>::lowest();
  ns[g] = -1;
}
template <typename D>
__global__ void init_row0_kernel(int n, long long inst_stride, D* __restrict__ ms) {
  const int inst = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ms[inst * inst_stride + i] = 0;
}

// ---------------------------------------------------------------------------------------------
// Backtrace (DP_impl.hpp:122-139): boundaries by walking nextStart_, then one range score per subset.
// ---------------------------------------------------------------------------------------------
// One thread per (instance, S).  bounds row layout: stride (Tc+1) per S.
__global__ void backtrace_bounds_kernel(int R, int n, int Tc, int sweep, const int* __restrict__ ns,
                                        int* __restrict__ bounds) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int nS = sweep ? Tc + 1 : 1;
  if (g >= R * nS) return;
  const int inst = g / nS;
  const int S = sweep ? (g % nS) : Tc;
  int* out = bounds + ((long long)inst * nS + (sweep ? S : 0)) * (Tc + 1);
  const int* nsi = ns + (long long)inst * (Tc + 1) * n;
  int cur = 0;
  out[0] = 0;
  for (int t = S; t > 0; --t) {
    int nxt = (cur >= 0 && cur < n) ? nsi[(long long)t * n + cur] : -1;
    out[S - t + 1] = nxt;
    cur = nxt;
  }
  for (int q = S + 1; q <= Tc; ++q) out[q] = -1;
}

template <typename D>
__global__ void backtrace_scores_kernel(int R, int n, int Tc, int sweep, int dist, int rp, int running,
                                        const int* __restrict__ bounds, const D* __restrict__ as,
                                        const D* __restrict__ bs, const Rec4<D>* __restrict__ rec,
                                        long long rec_stride, const D* __restrict__ SA,
                                        const D* __restrict__ SB, int L, int nch_alloc,
                                        D* __restrict__ scores) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int nS = sweep ? Tc + 1 : 1;
  const long long total = (long long)R * nS * Tc;
  if (g >= total) return;
  const int t = g % Tc;
  const int sidx = (g / Tc) % nS;
  const int inst = g / (Tc * nS);
  const int S = sweep ? sidx : Tc;
  D out = 0;
  if (t < S) {
    const int* bd = bounds + ((long long)inst * nS + sidx) * (Tc + 1);
    const int lo = bd[t], hi = bd[t + 1];
    if (lo >= 0 && hi >= lo && hi <= n) {
      D A = 0, B = 0;
      if (running) {
        const D* ap = as + (long long)inst * n;
        const D* bp = bs + (long long)inst * n;
        // left-to-right sum from lo, like the reference; the level-1 pass already holds that running sum at
        // every chunk boundary c*L > lo (SA/SB), so only the tail after the last boundary is walked here
        int k = lo;
        if (SA) {
          const int c = min(hi / L, nch_alloc - 1);
          if (c >= 1 && (long long)c * L > lo) {
            const long long pstride = (long long)nch_alloc * n;
            A = SA[inst * pstride + (long long)c * n + lo];
            B = SB[inst * pstride + (long long)c * n + lo];
            k = c * L;
          }
        }
#pragma unroll 8
        for (; k < hi; ++k) {
          A = add_rn<D>(A, ap[k]);
          B = add_rn<D>(B, bp[k]);
        }
      } else {
        const Rec4<D>* r = rec + inst * rec_stride;
        A = sub_rn<D>(r[hi].x, r[lo].x);
        B = sub_rn<D>(r[hi].y, r[lo].y);
      }
      out = score_dyn<D>(dist, rp, A, B);
    } else {
      out = nan("");
    }
  }
  scores[g] = out;
}

// Row-sharded runs: the per-subset scores of a backtraced partition (DP_impl.hpp:131-139), one thread per subset.
// prefix != 0: the records hold scan values of exactly summable inputs, so the difference IS the reference's running sum;
// otherwise the subset is walked left to right like the reference does.
template <typename D>
__global__ void shard_subset_scores_kernel(int count, int n, int dist, int rp, int prefix,
                                           const int* __restrict__ bounds, const Rec4<D>* __restrict__ rec,
                                           D* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int lo = bounds[t], hi = bounds[t + 1];
  if (!(lo >= 0 && hi >= lo && hi <= n)) {
    out[t] = (D)nan("");
    return;
  }
  D A = 0, B = 0;
  if (prefix) {
    A = sub_rn<D>(rec[hi].x, rec[lo].x);
    B = sub_rn<D>(rec[hi].y, rec[lo].y);
  } else {
#pragma unroll 8
    for (int k = lo + 1; k <= hi; ++k) {
      A = add_rn<D>(A, rec[k].x);
      B = add_rn<D>(B, rec[k].y);
    }
  }
  out[t] = score_dyn<D>(dist, rp, A, B);
}

// compute_score (python_dpsolver.cpp:74-111): S(0,n) of the unsorted inputs, sequential sums.
template <typename D>
__global__ void range_score_kernel(int n, int dist, int rp, const D* __restrict__ a,
                                   const D* __restrict__ b, D* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  D A = 0, B = 0;
#pragma unroll 8
  for (int k = 0; k < n; ++k) {
    A = add_rn<D>(A, a[k]);
    B = add_rn<D>(B, b[k]);
  }
  *out = score_dyn<D>(dist, rp, A, B);
}

// Self-test of math_fast.cuh: counts operands (uniform in log-space over the safe range) on which the
// branch-free routine differs from the CUDA math library.  which: 0 fdiv, 1 logf, 2 ddiv, 3 log.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& s) {
  unsigned long long z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float rand_safe_f32(unsigned long long& s, int emin, int emax) {
  const unsigned long long r = splitmix64(s);
  const int e = emin + (int)((r >> 40) % (unsigned)(emax - emin + 1));
  return __int_as_float(((e + 127) << 23) | (int)(r & 0x7fffff));
}
__device__ __forceinline__ double rand_safe_f64(unsigned long long& s, int emin, int emax) {
  const unsigned long long r = splitmix64(s);
  const unsigned long long r2 = splitmix64(s);
  const long long e = emin + (long long)((r >> 40) % (unsigned)(emax - emin + 1));
  return __longlong_as_double(((e + 1023) << 52) | (long long)(r2 & 0xfffffffffffffULL));
}
__global__ void selftest_math_kernel(int which, long long per_thread, unsigned long long seed,
                                     unsigned long long* __restrict__ mismatches) {
  unsigned long long s = seed + 0x632be59bd9b4e019ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
  unsigned long long bad = 0;
  for (long long it = 0; it < per_thread; ++it) {
    if (which == 0) {
      const float a = rand_safe_f32(s, -30, 60), b = rand_safe_f32(s, -30, 60);
      bad += __float_as_int(fast_fdiv(a, b)) != __float_as_int(__fdiv_rn(a, b));
    } else if (which == 1) {
      // fast and general entry points of the glibc logf restatement agree on the safe range
      const float x = (it & 1) ? rand_safe_f32(s, -1, 0) : rand_safe_f32(s, -90, 90);
      bad += __float_as_int(glibc_logf_fast(x)) != __float_as_int(glibc_logf(x));
    } else if (which == 2) {
      const double a = rand_safe_f64(s, -30, 60), b = rand_safe_f64(s, -30, 60);
      bad += __double_as_longlong(fast_ddiv(a, b)) != __double_as_longlong(__ddiv_rn(a, b));
    } else {
      const double x = (it & 1) ? rand_safe_f64(s, -1, 0) : rand_safe_f64(s, -90, 90);
      bad += __double_as_longlong(glibc_log_fast(x)) != __double_as_longlong(glibc_log(x));
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

// ---------------------------------------------------------------------------------------------
// LTSS (reference LTSS_impl.hpp:73-111): best single subset among the prefixes S(0,i), i = 1..n, then the
// suffixes S(i,n), i = n-1..0, of the priority-sorted sequence under the multiple-clustering score; the first
// strictly greater candidate in that visiting order wins.  Suffix scores are exactly level 1 of the DP
// (level1_running_kernel); prefix sums are one sequential walk (two dependent adds per element).
// ---------------------------------------------------------------------------------------------
template <typename D>
__global__ void ltss_prefix_sums_kernel(int n, const D* __restrict__ as, const D* __restrict__ bs,
                                        D* __restrict__ PA, D* __restrict__ PB) {
  // thread 0 walks a, thread 1 walks b: left to right from 0, like a_sums_[0][j] (score_impl.hpp:248-256)
  if (blockIdx.x || threadIdx.x > 1) return;
  const D* src = threadIdx.x ? bs : as;
  D* dst = threadIdx.x ? PB : PA;
  D acc = 0;
#pragma unroll 8
  for (int k = 0; k < n; ++k) {
    acc = add_rn<D>(acc, src[k]);
    dst[k] = acc;
  }
}

__device__ __forceinline__ unsigned int float_order_key(float v) {
  const unsigned int u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// key = (orderable score << 32) | (0xffffffff - visiting order): atomicMax picks the largest score and, among
// equal scores, the candidate visited first.  Scores are float whatever D is (LTSS.hpp:53,59).
template <typename D>
__global__ void ltss_argmax_kernel(int n, int dist, const D* __restrict__ PA, const D* __restrict__ PB,
                                   const D* __restrict__ suffix_score, unsigned long long* __restrict__ best) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long key = 0;
  if (g < n) {
    // prefix [0, g+1): visiting order g
    const float sp = (float)score_dyn<D>(dist, 0, PA[g], PB[g]);
    if (sp > -3.402823466e+38f)
      key = ((unsigned long long)float_order_key(sp) << 32) | (0xffffffffu - (unsigned)g);
    // suffix [g, n): visited after every prefix, from i = n-1 down to 0
    const float ss = (float)suffix_score[g];
    if (ss > -3.402823466e+38f) {
      const unsigned long long k2 =
          ((unsigned long long)float_order_key(ss) << 32) | (0xffffffffu - (unsigned)(n + (n - 1 - g)));
      key = k2 > key ? k2 : key;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}

// Evaluate the device log routines on caller-supplied arguments (tests compare the bits with the host's libm).
template <typename D>
__global__ void eval_log_kernel(long long n, const D* __restrict__ x, D* __restrict__ y_general,
                                D* __restrict__ y_fast) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const D v = x[g];
  if (sizeof(D) == 4) {
    y_general[g] = (D)glibc_logf((float)v);
    y_fast[g] = (D)glibc_logf_fast((float)v);
  } else {
    y_general[g] = (D)glibc_log((double)v);
    y_fast[g] = (D)glibc_log_fast((double)v);
  }
}

// Sharded runs: scatter the all-gathered [world][chunk] level values back to natural row order
// and into the p field of the next level's records.
template <typename D, int BR>
__global__ void shard_unpermute_kernel(int n, int world, int chunk, const D* __restrict__ gathered,
                                       Rec4<D>* __restrict__ rec_next, D* __restrict__ level_nat,
                                       float* __restrict__ pm_next) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)world * chunk) return;
  const int r = (int)(g / chunk);
  const int m = (int)(g % chunk);
  const int rb = r + world * (m / BR);
  const int i = rb * BR + (m % BR);
  if (i < n) {
    const D v = gathered[g];
    rec_next[i].p = v;
    if (pm_next) pm_next[(long long)i * 4 + 3] = screen_pm<D>(v);
    if (level_nat) level_nat[i] = v;
  }
}

// Row-sharded solves inside the library (ps_solve_sharded_*): after a level every rank all-gathers one packed chunk
// { D value[chunk]; int split[chunk]; } (owned rows in block-cyclic order).  This kernel scatters the gathered chunks back to
// natural row order: maxScore_[j], nextStart_[j], the p field of the next level's records and its float image pm.
template <typename D, int BR>
__global__ void shard_scatter_kernel(int n, int nrows, int world, int chunk, const unsigned char* __restrict__ gathered,
                                     D* __restrict__ ms_row, int* __restrict__ ns_row, Rec4<D>* __restrict__ rec_next,
                                     float* __restrict__ pm_next) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)world * chunk) return;
  const int r = (int)(g / chunk);
  const int m = (int)(g % chunk);
  const int rb = r + world * (m / BR);
  const int i = rb * BR + (m % BR);
  if (i >= nrows) return;
  const size_t stride = (size_t)chunk * (sizeof(D) + sizeof(int));
  const D v = reinterpret_cast<const D*>(gathered + (size_t)r * stride)[m];
  const int a = reinterpret_cast<const int*>(gathered + (size_t)r * stride + (size_t)chunk * sizeof(D))[m];
  if (ms_row) ms_row[i] = v;
  if (ns_row) ns_row[i] = a;
  if (rec_next) rec_next[i].p = v;
  if (pm_next) pm_next[(long long)i * 4 + 3] = screen_pm<D>(v);
}

__global__ void fill_int_kernel(long long count, int value, int* __restrict__ dst) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < count) dst[g] = value;
}

}  // namespace ps
