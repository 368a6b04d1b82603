// screen.cuh -- the SCREENED level kernel: every (i,k) cell is bounded in float, only cells that can still
// win are evaluated with the reference's arithmetic.
//
// One DP level (DP_impl.hpp:101-118) is  out[i] = max_k S(i,k) + prev[k], first maximum wins.  The exhaustive
// kernel (kernels.cuh) spends 60-100 issue cycles per cell on IEEE division and a bit-exact glibc log although
// all but a few cells of a row lose by a wide margin.  This kernel visits the same cells, but per cell it forms a
// cheap float UPPER BOUND  ub >= S_ref(i,k) + prev[k]  (MUFU.RCP / MUFU.LG2, ~14 instructions) and compares it with a
// LOWER BOUND on the row's maximum (an exactly evaluated cell of the same row); a cell whose upper bound is below
// that cannot be the row's maximum and is dropped.  Surviving cells (a few per cent: the neighbourhood of the
// maximum) are evaluated by the same ScoreFn functors as the exhaustive kernel, in ascending k with a strict >,
// so value AND first-argmax are the reference's, bit for bit:
//   * the first maximiser k* has v(k*) = max >= LB, hence ub(k*) >= LB: it is always evaluated exactly;
//   * every evaluated cell before k* is strictly smaller (k* is the FIRST maximiser), cells after it never replace it;
//   * a NaN bound compares false with "ub < thr" and is evaluated exactly; a NaN value never wins, as in the reference.
//
// Preconditions (checked per call on the device; otherwise the exhaustive kernel runs):
//   (P1) safe range (range_check_kernel) and a[i] > 0, b[i] > 0: all range sums are positive normal numbers;
//   (P2) exact summability (exact_check_kernel): every a[i] (b[i]) is a multiple of one power of two 2^e and
//        sum|a| <= 2^(e+p) (p = 24/53), e.g. counts.  Then EVERY partial sum is representable, every addition the
//        reference performs is exact, and range sums formed as differences of a prefix scan are bit-identical to
//        the reference's running sums (score_impl.hpp:248-256) -- cells can be evaluated in any order.
//
// Error budget of the bound (u = 2^-24).  Hardware claims, verified over EVERY float of the safe range by
// ps_selftest_mufu (measured on B200: 1.65u and 3.85u):
//   (R1) |x * rcp.approx(x) - 1| <= 4u;   (R2) |lg2.approx(x) - log2 x| <= 4u * max(1, |log2 x|).
// Poisson mult-clust, g(A,B) = A ln(A/B) + B - A, q = A/B > 1, t = fl(At * lg2(fl(At * rcp(Bt)))):
//                                                                        relative to  t ln2 |  A  | |p|
//   inputs  double: At = fl(lx[k] + bx), lx = fl32(P[k] - P[o]), bx = fl32(P[o] - P[i]) >= 0,
//           |At - A| <= 2u A; float: At = P[k] - P[i] exactly (P2).
//           |g(At,Bt) - g(A,B)| <= 2u (A |ln q| + |A - B|)                      2u    | 2u  |
//   MUFU    ln(q) from rcp, one multiply, lg2: |error| <= 5u + 4u ln2 + 4u |ln q|   4u+1u | 7.8u|
//   chain   dm, dm + pm, the final fma: each rounds by u relative to its result        1u    | 3u  | 2u
//   ref     the reference's own rounding, D = float (negligible for double):
//           fl(A/B), logf (< 1 ulp), the product, + B, - A, + p                         6u    | 2u  | 1u
//                                                                             total    14u    |14.8u| 3u
//   => ub = t ln2 (1 + 16u) + (B - A) + 20u A + pm,  pm = ru(p) + 16u |p|   dominates S_ref(A,B) + p.
//   Poisson risk-part (A ln q, d/dA = ln q + 1, d/dB = -q) needs 15.8u on A.  The A^2/B family is bounded
//   relative to g itself: inputs 6u, rcp 4u, roundings 4u < 32u.  Every dropped cell is re-checked in audit mode
//   (ps_problem.screen = 2); the GPU tests require 0 violations on ~10^10 cells.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "kernels.cuh"

namespace ps {

__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr float SCR_U = 5.9604644775390625e-08f;                                        // u = 2^-24
constexpr float SCR_LN2_UP = 0.693147182464599609375f * (1.0f + 16.0f * SCR_U);         // ln2 (1 + 16u)
constexpr float SCR_REL_UP = 1.0f + 32.0f * SCR_U;                                      // 1 + 32u
constexpr float SCR_CA = 20.0f * SCR_U;                                                 // 20 u
constexpr float SCR_TINY = 9.31322574615478515625e-10f;                                 // 2^-30

// Upper bound of S(A,B) + p from float sums At, Bt (> 0) and pm = screen_pm(p).
template <int DIST, bool RP>
struct ScreenUB;

template <>
struct ScreenUB<DIST_POISSON, false> {  // (A > B) ? A ln(A/B) + B - A : 0      score_impl.hpp:385-388
  static __device__ __forceinline__ float eval(float A, float B, float pm) {
    if (A > B) {
      const float q = A * rcp_approx(B);
      const float t = A * lg2_approx(q);
      const float dm = fmaf(A, -(1.0f - SCR_CA), B);  // B - A + 48u A
      return fmaf(t, SCR_LN2_UP, dm + pm);
    }
    // At <= Bt although A > B is only possible within the input error: then S <= (6uA)^2 / B << 2^-30 B
    return fmaf(B, SCR_TINY, pm);
  }
};
template <>
struct ScreenUB<DIST_POISSON, true> {  // A ln(A/B), either sign       score_impl.hpp:377-378
  // d/dA = ln q + 1, d/dB = -q: input error 3u (A |ln q| + 2A); the remaining terms as for the mult-clust score
  static __device__ __forceinline__ float eval(float A, float B, float pm) {
    const float q = A * rcp_approx(B);
    const float t = A * lg2_approx(q);
    const float e = fmaf(fabsf(t), 0.693147182464599609375f * 16.0f * SCR_U, fmaf(A, SCR_CA, pm));
    return fmaf(t, 0.693147182464599609375f, e);
  }
  // One bound for every cell of a k-range of the row: A in [Af, Al], B >= Bf (the sums grow with k; needs a, b > 0 only, no
  // sortedness).  f(A,B) = A ln(A/B) falls in B and is convex in A, so f <= max(f(Af,Bf), f(Al,Bf)) on the box.  Margins:
  // the log term relative to |f| -- if the cell's f is negative and below the corner maximum m, then f (1 - 16u) <= m -/+ 16u |m|
  // whatever the sign of m -- and 20u A with A <= Al.
  static __device__ __forceinline__ float box(float Af, float Bf, float Al, float pm) {
    const float rb = rcp_approx(Bf);
    const float t1 = Af * lg2_approx(Af * rb);
    const float t2 = Al * lg2_approx(Al * rb);
    const float e1 = fmaf(fabsf(t1), 0.693147182464599609375f * 16.0f * SCR_U, t1 * 0.693147182464599609375f);
    const float e2 = fmaf(fabsf(t2), 0.693147182464599609375f * 16.0f * SCR_U, t2 * 0.693147182464599609375f);
    return fmaxf(e1, e2) + fmaf(Al, SCR_CA, pm);
  }
  // Ascending priorities: along a row f is QUASI-CONVEX in k, so every cell of a k-range is bounded by the larger of the two
  // end cells.  Adding element k (ratio r = a_k / b_k >= q = A/B, the b-weighted mean of the ratios before it) changes f by
  // b_k (r (1 + ln q) - q): negative while q <= 1/e; for q > 1/e it is >= 0 iff r >= h(q) = q / (1 + ln q), and once that holds
  // it holds for every later element -- q' in [q, r], r' >= r, h falls on (1/e, 1) so h(q') <= h(q) <= r <= r', and
  // h(q') <= q' <= r' for q' >= 1.  So f falls, then rises.  The box above decouples A from B and is useless where q ~ 1
  // (null data: 93 % of the 8-cell groups kept); this bound is as tight as the cell bound at the better end.  Margins as for
  // the box: the log term relative to the corner maximum, 20u A with the range's largest A.
  static __device__ __forceinline__ float hull(float Af, float Bf, float Al, float Bl, float pm) {
    const float t1 = Af * lg2_approx(Af * rcp_approx(Bf));
    const float t2 = Al * lg2_approx(Al * rcp_approx(Bl));
    const float e1 = fmaf(fabsf(t1), 0.693147182464599609375f * 16.0f * SCR_U, t1 * 0.693147182464599609375f);
    const float e2 = fmaf(fabsf(t2), 0.693147182464599609375f * 16.0f * SCR_U, t2 * 0.693147182464599609375f);
    return fmaxf(e1, e2) + fmaf(Al, SCR_CA, pm);
  }
};
template <>
struct ScreenUB<DIST_GAUSSIAN, false> {  // (A > B) ? .5 (A^2/B + B) - A : 0 = (A-B)^2 / (2B)   score_impl.hpp:429-431
  // g(At,Bt) is evaluated in its cancellation-free form, relative error 2^-21; input error
  // |dg| <= (q-1) 3uA + .5 (q^2-1) 3uB <= 6u A (q-1); reference rounding <= 9 u_D A q (+ u g when D = float).
  static __device__ __forceinline__ float eval(float A, float B, float pm) {
    if (A > B) {
      const float dif = A - B;
      const float h = dif * rcp_approx(B);  // q - 1
      const float gt = (0.5f * dif) * h;
      const float ah = A * h;
      float e = fmaf(ah, 8.0f * 5.9604644775390625e-08f, pm);
      e = fmaf(A, SCR_TINY, e);
      return fmaf(gt, SCR_REL_UP, e);
    }
    return fmaf(B, SCR_TINY, pm);
  }
};
template <>
struct ScreenUB<DIST_GAUSSIAN, true> {  // A^2 / (2B)       score_impl.hpp:438-439
  // everything is relative to g: inputs 9u, rcp 4u, roundings 3u, reference 3u  < 32u
  static __device__ __forceinline__ float eval(float A, float B, float pm) {
    const float gt = (0.5f * A) * (A * rcp_approx(B));
    return fmaf(gt, SCR_REL_UP, pm);
  }
};
template <bool RP>
struct ScreenUB<DIST_RATIONAL, RP> {  // A^2 / B          score_impl.hpp:464-465,471-472
  static __device__ __forceinline__ float eval(float A, float B, float pm) {
    const float gt = A * (A * rcp_approx(B));
    return fmaf(gt, SCR_REL_UP, pm);
  }
};

// Bound for a whole k-range of a row from its first (Af, Bf) and last (Al, Bl) cell and the largest pm of the range.
// Scores that never decrease when the range grows (ascending priorities): the bound at the last cell.  Poisson risk-part:
// the larger of the two end cells when the priorities ascend (hull), else the box bound.  `mono` is the range-tier mode of the
// level: 0 none, 1 on, 2 on and the priorities are known to ascend.
template <int DIST, bool RP>
__device__ __forceinline__ float screen_range_ub(float Af, float Bf, float Al, float Bl, float pm, const int mono = 1) {
  if constexpr (DIST == DIST_POISSON && RP)
    return (mono == 2) ? ScreenUB<DIST_POISSON, true>::hull(Af, Bf, Al, Bl, pm)
                       : ScreenUB<DIST_POISSON, true>::box(Af, Bf, Al, pm);
  else
    return ScreenUB<DIST, RP>::eval(Al, Bl, pm);
}

// ---------------------------------------------------------------------------------------------
// (P2) exact summability.  stats[inst] = {min over elements of the exponent of the lowest set bit, sum |.|}
// for a and for b; screen_decide_kernel raises flag bit 2 (value 4) if some instance fails.
// ---------------------------------------------------------------------------------------------
struct ScreenStats {
  int ea, eb;
  double sa, sb;
  // for the fixed-point bounds of screen_run.cuh: bit patterns of the smallest positive a and the smallest b
  // (positive doubles order like their bit patterns), number of elements with a <= 0
  unsigned long long min_a_pos, min_b;
  int n_nonpos, pad;
};
__device__ __forceinline__ int lowbit_exp(double v) {
  const long long bits = __double_as_longlong(v);
  const int e = (int)((bits >> 52) & 0x7ff) - 1075;
  const long long m = (bits & 0xfffffffffffffLL) | (1LL << 52);
  return e + __ffsll(m) - 1;
}
__device__ __forceinline__ int lowbit_exp(float v) {
  const int bits = __float_as_int(v);
  const int e = ((bits >> 23) & 0xff) - 150;
  const int m = (bits & 0x7fffff) | 0x800000;
  return e + __ffs(m) - 1;
}
__global__ void screen_stats_init_kernel(int R, ScreenStats* __restrict__ st) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) {
    st[r].ea = st[r].eb = 0x7fffffff;
    st[r].sa = st[r].sb = 0.0;
    st[r].min_a_pos = st[r].min_b = 0x7ff0000000000000ULL;
    st[r].n_nonpos = st[r].pad = 0;
  }
}
template <typename D>
__global__ void __launch_bounds__(256)
screen_stats_kernel(int n, const D* __restrict__ a, const D* __restrict__ b, ScreenStats* __restrict__ st) {
  const int inst = blockIdx.y;
  a += (long long)inst * n;
  b += (long long)inst * n;
  int ea = 0x7fffffff, eb = 0x7fffffff;
  double sa = 0.0, sb = 0.0;
  unsigned long long ma = 0x7ff0000000000000ULL, mb = 0x7ff0000000000000ULL;
  int np = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const D av = a[k], bv = b[k];
    if (av != (D)0) ea = min(ea, lowbit_exp(av));
    if (bv != (D)0) eb = min(eb, lowbit_exp(bv));
    sa += fabs((double)av);
    sb += fabs((double)bv);
    if (av > (D)0)
      ma = min(ma, (unsigned long long)__double_as_longlong((double)av));
    else
      ++np;
    if (bv > (D)0) mb = min(mb, (unsigned long long)__double_as_longlong((double)bv));
  }
  for (int o = 16; o > 0; o >>= 1) {
    ea = min(ea, __shfl_xor_sync(0xffffffffu, ea, o));
    eb = min(eb, __shfl_xor_sync(0xffffffffu, eb, o));
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    ma = min(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    mb = min(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    np += __shfl_xor_sync(0xffffffffu, np, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&st[inst].ea, ea);
    atomicMin(&st[inst].eb, eb);
    atomicAdd(&st[inst].sa, sa);
    atomicAdd(&st[inst].sb, sb);
    atomicMin(&st[inst].min_a_pos, ma);
    atomicMin(&st[inst].min_b, mb);
    if (np) atomicAdd(&st[inst].n_nonpos, np);
  }
}
template <typename D>
__global__ void screen_decide_kernel(int R, const ScreenStats* __restrict__ st, int* __restrict__ flags, int n = 0) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (st[r].n_nonpos > n / 8) atomicOr(flags, 16);  // too many rows without bounds for the running-sum screens
  constexpr int P = (sizeof(D) == 8) ? 53 : 24;
  // the double sums are exact whenever the instance qualifies; the 2^-20 slack covers their rounding when it does not
  const double slack = 1.0 - 9.5367431640625e-07;
  bool ok = true;
  if (st[r].ea != 0x7fffffff) ok = ok && (st[r].sa <= ldexp(slack, st[r].ea + P));
  if (st[r].eb != 0x7fffffff) ok = ok && (st[r].sb <= ldexp(slack, st[r].eb + P));
  if (!ok) atomicOr(flags, 4);
}

// Priorities ascending?  (always true after the device sort; a caller-supplied permutation is only trusted after this
// check.)  flag[0] |= 1 if some a[k]/b[k] > a[k+1]/b[k+1] or a key is NaN.
template <typename D>
__global__ void screen_sorted_kernel(int n, const D* __restrict__ as, const D* __restrict__ bs, int* __restrict__ flag) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= n) return;
  as += (long long)inst * n;
  bs += (long long)inst * n;
  const D k0 = as[k] / bs[k], k1 = as[k + 1] / bs[k + 1];
  if (!(k0 <= k1)) atomicOr(flag, 1);
}

// Float screening records, one float4 per k, staged by the level kernel:
//   double: a separate array {lx, ly, gmax, pm}: prefix sums RELATIVE to the start of the 128-block k falls in
//           (origin o(k) = 128 floor((k-1)/128); a stage of the level kernel is exactly one block), so that float range
//           sums carry a relative error (header); pm = screen_pm(prev[k]) is written by level1_prefix_kernel /
//           combine_kernel.
//   float : the level records themselves {Px, Py, p, gmax}; pm = p + 2^-20 |p| is formed on the fly.
// gmax (screen_lb_kernel, once per level) is kept in the LAST record of every aligned group of 8 (k = 8m+8): the
// largest pm of k = 8m+1 .. 8m+8.
constexpr int SCR_BLOCK = 128;
__global__ void build_recf_kernel(int n, long long rec_stride, const Rec4<double>* __restrict__ rec,
                                  float4* __restrict__ recf0, float4* __restrict__ recf1) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  rec += inst * rec_stride;
  const int o = (k > 0) ? ((k - 1) / SCR_BLOCK) * SCR_BLOCK : 0;
  const Rec4<double> r = rec[k], q = rec[o];
  const float4 f = make_float4(__double2float_rn(r.x - q.x), __double2float_rn(r.y - q.y), 0.f, 0.f);
  recf0[inst * rec_stride + k] = f;
  recf1[inst * rec_stride + k] = f;
}

template <typename D>
__device__ __forceinline__ float screen_pm_of(const float4 r) {
  if (sizeof(D) == 8) return r.w;
  return fmaf(fabsf(r.z), 9.5367431640625e-07f, r.z);
}
template <typename D>
__device__ __forceinline__ float screen_gmax_of(const float4 r) {
  return (sizeof(D) == 8) ? r.z : r.w;
}

template <typename D, int DIST, bool RP, bool LOGTAB>
__device__ __forceinline__ D screen_exact_score(D A, D B, const GlogEntry* s_log) {
  if constexpr (!RP && DIST != DIST_RATIONAL) {
    if (!gt_ab<true>(A, B)) return (D)0;
    if constexpr (LOGTAB)
      return ScoreFn<D, DIST, RP, true>::eval_gt(A, B, s_log);
    else
      return ScoreFn<D, DIST, RP, true>::eval_gt(A, B);
  } else if constexpr (LOGTAB) {
    return ScoreFn<D, DIST, RP, true>::eval(A, B, s_log);
  } else {
    return ScoreFn<D, DIST, RP, true>::eval(A, B);
  }
}

// ---------------------------------------------------------------------------------------------
// Per level, before the tiles: (1) a lower bound of every row's maximum = the best of <= 2*32 EXACTLY evaluated cells
// of the row (a coarse sweep over the row's k range, then a finer one around the best coarse sample).  Any evaluated
// cell is a valid bound; the sampling only decides how tight it is (on C3 the best sample is within 1 of the maximum).
// (2) the group maxima of pm (one thread per aligned group of 8 records).
// ---------------------------------------------------------------------------------------------
template <typename D, int DIST, bool RP>
__global__ void __launch_bounds__(128)
screen_lb_kernel(Geom g, const Rec4<D>* __restrict__ rec, float4* __restrict__ recf, D* __restrict__ LB,
                 float* __restrict__ bmax, int nblk, int* __restrict__ kbest = nullptr) {
  const int inst = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nrows) return;
  rec += (long long)inst * g.rec_stride;
  recf += (long long)inst * g.rec_stride;
  if ((i & (SCR_BLOCK - 1)) == 0) {  // largest pm of the 128-block k = i+1 .. i+128 (tile tier)
    float m = -3.402823466e+38f;
    const int ke = min(i + SCR_BLOCK, g.n);
    for (int k = i + 1; k <= ke; ++k) m = fmaxf(m, screen_pm_of<D>(recf[k]));
    bmax[(long long)inst * nblk + i / SCR_BLOCK] = m;
  }
  if ((i & 7) == 0 && i + 8 <= g.n) {
    float m = screen_pm_of<D>(recf[i + 1]);
#pragma unroll
    for (int q = 2; q <= 8; ++q) m = fmaxf(m, screen_pm_of<D>(recf[i + q]));
    if (sizeof(D) == 8)
      recf[i + 8].z = m;
    else
      recf[i + 8].w = m;
  }
  {  // row-sharded runs: the group / block maxima above are needed for every k, the lower bound only for owned rows
    const int rbi = i / SCR_BLOCK;
    if (rbi < g.rb_first || (rbi - g.rb_first) % g.rb_step != 0) return;
  }
  const D Xi = rec[i].x, Yi = rec[i].y;
  const int k0 = i + 1, k1 = g.kmax;
  const int len = k1 - k0 + 1;
  D best = Lim<D>::lowest();
  int kb = k1;
  auto probe = [&](int k) {
    const Rec4<D> r = rec[k];
    const D v = add_rn<D>(screen_exact_score<D, DIST, RP, false>(sub_rn<D>(r.x, Xi), sub_rn<D>(r.y, Yi), nullptr), r.p);
    if (v > best) {
      best = v;
      kb = k;
    }
  };
  // a sweep k = lo, lo + step, ... <= hi, four samples at a time: the record loads of a batch are independent and issued
  // together (one thread walks one row, so nothing else hides their latency); ties keep the earlier sample
  auto sweep = [&](const int lo, const int hi, const int step) {
    for (int k = lo; k <= hi; k += 4 * step) {
      Rec4<D> r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = rec[min(k + u * step, hi)];
      D v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = add_rn<D>(screen_exact_score<D, DIST, RP, false>(sub_rn<D>(r[u].x, Xi), sub_rn<D>(r[u].y, Yi), nullptr), r[u].p);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k + u * step <= hi && v[u] > best) {
          best = v[u];
          kb = k + u * step;
        }
    }
  };
  const int ns = max(4, min(32, len >> 8));
  const int sc = max(1, len / ns);
  sweep(k0, k1, sc);
  probe(k1);
  if (sc > 1) {
    const int lo = max(k0, kb - sc), hi = min(k1, kb + sc);
    sweep(lo, hi, max(1, (hi - lo) / ns));
  }
  LB[(long long)inst * g.n + i] = best;
  if (kbest) kbest[(long long)inst * g.n + i] = kb;  // where the row's best sample sits: the select kernel's cost hint
}

// ---------------------------------------------------------------------------------------------
// Screened level = four launches:
//   screen_lb_kernel      row lower bounds + group / block maxima of pm (above);
//   screen_select_kernel  the RANGE tiers for every (row block, k-chunk) tile, before anything heavy is launched: one bound
//                         per (row, tile) and, where that does not already lose, one per (row, 128-cell stage) -- S at the
//                         stage's last cell + the stage's largest pm.  A tile none of whose stages can hold a row maximum is
//                         dropped here; the others are appended, with their mask of live stages, to a compacted list
//                         (one list per cost class); every tile's mask is also kept for the fold;
//   level_live_screen_kernel   persistent CTAs pull live tiles from that list (atomic cursor) and run the group and cell
//                         tiers on the live stages only.  Dropped tiles cost nothing;
//   screen_fold_kernel    folds every row block's partials over its live tiles only, in ascending k (strict >: first maximum
//                         wins, as in combine_kernel) and writes maxScore_[j], nextStart_[j] and the next level's p / pm.
// Tiers inside a live tile, cheapest first:
//   stage  (scores monotone in k) re-checked with the threshold the tile has raised so far;
//   group  one bound per 8 cells, with the group's largest pm;
//   cell   one bound per cell, 8 independent chains, one vote; a group with a survivor is evaluated exactly.
// Monotonicity needs ascending priorities (dev_flags[1] == 0, screen_sorted_kernel).
// counters[0] += warp-cells evaluated exactly, counters[1] += audit violations (a dropped cell that does not lose);
// audit runs list EVERY tile (so that the dropped ones can be re-evaluated cell by cell) and also count tiles kept/dropped
// [2,3], warp-stages kept/dropped [4,5], warp-groups kept/dropped [6,7].
// ---------------------------------------------------------------------------------------------
// Live tiles are appended to one of SCR_CLASSES lists by expected cost (share of the tile's (row, stage) pairs the range
// tiers left alive): the tiles that hold a row block's maximum are evaluated almost entirely with the reference's arithmetic
// and take 10-100x longer than a tile that merely touches the neighbourhood, so the persistent CTAs take the expensive
// classes first and the cheap tiles fill the tail.
constexpr int SCR_CLASSES = 4;
struct LiveHdr {
  int nlive[SCR_CLASSES];  // tiles appended to each class list by the select kernel of this level
  int cursor;              // next tile a persistent CTA takes (classes in order)
  int pad[3];
};
constexpr int SCR_MAX_G = 32;  // stages per tile: the live-stage mask of a tile is one 32-bit word

// One live tile: group and cell tiers on the live stages, partials to part_val / part_arg.
template <typename D, int DIST, bool RP, bool AUDIT, int BR>
__device__ __forceinline__ void screen_live_tile(const Geom& g, const int inst, const int rb, const int slot,
                                                 const unsigned int stage_mask, const Rec4<D>* __restrict__ rec,
                                                 const float4* __restrict__ recf, const D* __restrict__ LB,
                                                 D* __restrict__ part_val, int* __restrict__ part_arg,
                                                 unsigned long long* __restrict__ counters, const int mono,
                                                 float4 (*sbuf)[BR], float (*s_pmax)[BR / 32],
                                                 Rec4<D>* sdbl, const GlogEntry* __restrict__ s_log) {
  constexpr int BK = BR;
  constexpr int U = 8;   // cells per group
  constexpr int GQ = 4;  // groups whose bounds are formed together
  constexpr bool DBL = (sizeof(D) == 8);
  constexpr bool FRONTIER = !RP && DIST != DIST_RATIONAL;  // scores that are 0 unless A > B
  // Scores that never decrease when a row's range is extended to the right (priorities ascending, a > 0): with
  // r = a_k/b_k >= q = A/B,  dS/db = r ln q + 1 - q >= q ln q + 1 - q >= 0 (Poisson mult-clust, q > 1; S = 0 below),
  // (q-1) r - (q^2-1)/2 >= (q-1)^2/2 (Gaussian mult-clust), 2 q r - q^2 >= q^2 (A^2/B, A^2/2B).  Poisson risk-part
  // (A ln(A/B), decreasing while q < 1/e) is not monotone but quasi-convex: its range tiers take the larger end cell
  // (ScreenUB::hull), or the box bound when the priorities do not ascend (mono == 1).
  constexpr bool kLogTab = DBL && (DIST == DIST_POISSON);
  const int tid = threadIdx.x;
  const int i0 = rb * BR;
  const int i = i0 + tid;
  const int c = slot / g.nsub;  // the k-chunk; the partial goes to the entry's own slot
  const int k_lo = max(c * g.L + 1, i0 + 1);
  const int k_hi = min((c + 1) * g.L, g.kmax);

  const int il = min(i, g.n - 1);
  const D Xi = rec[il].x, Yi = rec[il].y;
  float bx = -(float)Xi, by = -(float)Yi;  // float: At = Px[k] - Px[i], exact.  double: set per stage
  D best = Lim<D>::lowest();
  int arg = -1;
  const float thr0 = screen_rd<D>(LB[min(il, max(g.nrows - 1, 0))]);  // the threshold the select kernel dropped with
  float thr = thr0;
  unsigned int nexact = 0, nviol = 0;

  // The records the reference's arithmetic reads are staged with the float ones: float data -- the screening record IS the
  // level record {Px, Py, p, -}; double data -- a second shared buffer filled by cp.async.  k lies in the stage in flight.
  int kbase = 0, buf = 0;
  auto exact_at = [&](const int k) -> D {
    D Rx, Ry, Rp;
    if constexpr (DBL) {
      const Rec4<D>& R = sdbl[buf * BR + (k - kbase)];
      Rx = R.x, Ry = R.y, Rp = R.p;
    } else {
      const float4 r = sbuf[buf][k - kbase];
      Rx = r.x, Ry = r.y, Rp = r.z;
    }
    return add_rn<D>(screen_exact_score<D, DIST, RP, kLogTab>(sub_rn<D>(Rx, Xi), sub_rn<D>(Ry, Yi), s_log), Rp);
  };
  auto stage_dbl = [&](const int kb, const int b) {  // this thread's record of the stage starting at kb -> sdbl[b]
    if constexpr (DBL) {
      if (kb + tid <= k_hi) {
        cp_async16(&sdbl[b * BR + tid], &rec[kb + tid]);
        cp_async16(reinterpret_cast<char*>(&sdbl[b * BR + tid]) + 16, reinterpret_cast<const char*>(&rec[kb + tid]) + 16);
      }
      cp_async_commit();
    }
  };
  // audit: a dropped cell must lose against the threshold it was dropped with
  auto audit_dropped = [&](const int k, const bool valid, const float t) {
    if (valid && !((double)exact_at(k) < (double)t)) ++nviol;
  };
  // one cell on its own: bound, vote, exact evaluation of the warp if any lane survives (diagonal and ragged stages)
  auto cell = [&](const int k, const float4 r, const bool valid, auto from_double) {
    float A, B;
    if constexpr (decltype(from_double)::value) {
      const Rec4<D>& R = sdbl[buf * BR + (k - kbase)];
      A = __double2float_rn((double)R.x - (double)Xi);
      B = __double2float_rn((double)R.y - (double)Yi);
    } else {
      A = __fadd_rn(r.x, bx);
      B = __fadd_rn(r.y, by);
    }
    const float ub = ScreenUB<DIST, RP>::eval(A, B, screen_pm_of<D>(r));
    const bool pass = valid && !(ub < thr);
    if (__any_sync(0xffffffffu, pass)) {
      const D v = exact_at(k);
      if (AUDIT && valid && (double)ub < (double)v) ++nviol;
      ++nexact;
      if (valid && v > best) {
        best = v;
        arg = k;
        thr = fmaxf(thr, screen_rd<D>(best));
      }
    } else if (AUDIT) {
      audit_dropped(k, valid, thr);
    }
  };
  using T_ = std::integral_constant<bool, true>;
  using F_ = std::integral_constant<bool, false>;
  // the largest pm of a stage from the group maxima held by the records with (k-1) % 8 == 7, i.e. lanes 7, 15, 23, 31
  auto stage_max = [&](const float4 r, const int buf) {
    float w = ((tid & 7) == 7) ? screen_gmax_of<D>(r) : -3.402823466e+38f;
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 8));
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 16));
    if ((tid & 31) == 31) s_pmax[buf][tid >> 5] = w;
  };

  const int nk = k_hi - k_lo + 1;
  const int nstages = (nk + BK - 1) / BK;
  if (AUDIT && counters && tid == 0) atomicAdd(&counters[stage_mask ? 2 : 3], 1ULL);
  // live stages in ascending k; audit runs walk every stage and re-evaluate the cells of the dropped ones
  unsigned int todo = AUDIT ? (nstages >= 32 ? 0xffffffffu : ((1u << nstages) - 1u)) : stage_mask;
  int s = todo ? (__ffs(todo) - 1) : -1;
  Rec4<D> onx;  // double: record at the origin of the staged stage (k = kb - 1)
  onx.x = onx.y = onx.p = onx.w = 0;
  if (s >= 0) {
    const int kb = k_lo + s * BK;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kb + tid <= k_hi) r = recf[kb + tid];
    sbuf[0][tid] = r;
    stage_max(r, 0);
    if (DBL) onx = rec[kb - 1];
    stage_dbl(kb, 0);
    if (DBL) cp_async_wait_all();
  }
  while (s >= 0) {
    __syncthreads();
    todo &= todo - 1u;
    const int sn = todo ? (__ffs(todo) - 1) : -1;  // next live stage: its records are fetched while this one is processed
    const int kb = k_lo + s * BK;
    kbase = kb;
    if (sn >= 0) stage_dbl(k_lo + sn * BK, buf ^ 1);
    float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sn >= 0 && k_lo + sn * BK + tid <= k_hi) nx = recf[k_lo + sn * BK + tid];
    if (DBL) {
      // At = lx[k] + bx with bx = fl32(P[kb-1] - P[i]) >= 0 on every stage right of the diagonal
      bx = __double2float_rn((double)onx.x - (double)Xi);
      by = __double2float_rn((double)onx.y - (double)Yi);
      if (sn >= 0) onx = rec[k_lo + sn * BK - 1];
    }
    const int cnt = min(BK, k_hi - kb + 1);
    const float4* __restrict__ sb = sbuf[buf];
    const bool sel_dropped = AUDIT && !((stage_mask >> s) & 1u);
    if (sel_dropped) {
      // (audit) a stage the select kernel dropped: every cell must lose against the row's initial threshold
      if (counters && (tid & 31) == 0) atomicAdd(&counters[5], 1ULL);
      for (int kk = 0; kk < cnt; ++kk) audit_dropped(kb + kk, (kb + kk > i) && (i < g.nrows), thr0);
    } else if (kb <= i0 + BR - 1) {
      // diagonal stage: row i only sees k > i; the double kernel converts exact differences (kb - 1 = i0 <= i)
      if (DBL) {
        for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], kb + kk > i, T_{});
      } else if (AUDIT || cnt < BK) {
        for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], kb + kk > i, F_{});
      } else {
        // float data: Px[k] - Px[i] is exact whatever the sign, so the diagonal stage runs the group scheme of the interior
        // stages with a per-lane mask
#pragma unroll 1
        for (int kk = 0; kk < BK; kk += U) {
          const int kg = kb + kk;
          if (mono) {
            const float4 re = sb[kk + U - 1], rf = sb[kk];
            const float ubg = screen_range_ub<DIST, RP>(__fadd_rn(rf.x, bx), __fadd_rn(rf.y, by), __fadd_rn(re.x, bx),
                                                        __fadd_rn(re.y, by), screen_gmax_of<D>(re), mono);
            // a lane with no cell in the group drops it, a lane that starts inside it keeps it, the others bound it
            const bool keep = (kg + U - 1 > i) && (!(kg > i) || !(ubg < thr));
            if (!__any_sync(0xffffffffu, keep)) continue;
          }
          float A[U], B[U], W[U];
          bool gt = false;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float4 r = sb[kk + u];
            A[u] = __fadd_rn(r.x, bx);
            B[u] = __fadd_rn(r.y, by);
            W[u] = screen_pm_of<D>(r);
            if (FRONTIER) gt |= (kg + u > i) && (A[u] > B[u]);
          }
          bool alive = false;
          if (FRONTIER && !__any_sync(0xffffffffu, gt)) {
#pragma unroll
            for (int u = 0; u < U; ++u) alive |= (kg + u > i) && !(fmaf(B[u], SCR_TINY, W[u]) < thr);
          } else {
#pragma unroll
            for (int u = 0; u < U; ++u) alive |= (kg + u > i) && !(ScreenUB<DIST, RP>::eval(A[u], B[u], W[u]) < thr);
          }
          if (__any_sync(0xffffffffu, alive)) {
            D v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = exact_at(kg + u);
#pragma unroll
            for (int u = 0; u < U; ++u)
              if ((kg + u > i) && v[u] > best) {
                best = v[u];
                arg = kg + u;
              }
            thr = fmaxf(thr, screen_rd<D>(best));
            nexact += U;
          }
        }
      }
    } else if (cnt < BK) {
      for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], true, F_{});
    } else {
      bool skip = false;
      if (mono) {
        // stage tier again, now against the threshold this tile has raised: the bound at the stage's LAST cell with the
        // stage's largest pm covers all 128 cells
        float pmx = s_pmax[buf][0];
#pragma unroll
        for (int w = 1; w < BR / 32; ++w) pmx = fmaxf(pmx, s_pmax[buf][w]);
        const float4 re = sb[BK - 1], rf = sb[0];
        const float ubs = screen_range_ub<DIST, RP>(__fadd_rn(rf.x, bx), __fadd_rn(rf.y, by), __fadd_rn(re.x, bx),
                                                    __fadd_rn(re.y, by), pmx, mono);
        skip = !__any_sync(0xffffffffu, !(ubs < thr));
      }
      if (AUDIT && counters && (tid & 31) == 0) atomicAdd(&counters[skip ? 5 : 4], 1ULL);
      if (skip) {
        if (AUDIT)
          for (int kk = 0; kk < BK; ++kk) audit_dropped(kb + kk, true, thr);
      } else {
#pragma unroll 1
        for (int kq = 0; kq < BK; kq += GQ * U) {
          // group tier, GQ groups at a time: their bound chains (LDS, 2 adds, MUFU.RCP, MUFU.LG2, FMA) are independent and
          // overlap; a group kept on a threshold that a group before it raises afterwards is merely kept once too often
          unsigned int keepm = (1u << GQ) - 1u;
          if (mono) {
            float ubg[GQ];
#pragma unroll
            for (int q = 0; q < GQ; ++q) {
              const float4 re = sb[kq + q * U + U - 1], rf = sb[kq + q * U];
              ubg[q] = screen_range_ub<DIST, RP>(__fadd_rn(rf.x, bx), __fadd_rn(rf.y, by), __fadd_rn(re.x, bx),
                                                 __fadd_rn(re.y, by), screen_gmax_of<D>(re), mono);
            }
            keepm = 0;
#pragma unroll
            for (int q = 0; q < GQ; ++q)
              if (__any_sync(0xffffffffu, !(ubg[q] < thr))) keepm |= 1u << q;
            if (AUDIT) {
              for (int q = 0; q < GQ; ++q) {
                const bool gdrop = !((keepm >> q) & 1u);
                if (counters && (tid & 31) == 0) atomicAdd(&counters[gdrop ? 7 : 6], 1ULL);
                if (gdrop)
                  for (int u = 0; u < U; ++u) audit_dropped(kb + kq + q * U + u, true, thr);
              }
            }
            if (!keepm) continue;
          }
#pragma unroll 1
          for (int q = 0; q < GQ; ++q) {
            if (!((keepm >> q) & 1u)) continue;
            const int kk = kq + q * U;
            // U independent chains (LDS, 2 adds, MUFU.RCP, MUFU.LG2, 4-5 FMA-pipe operations): their latencies overlap,
            // one vote decides for the group
            float A[U], B[U], W[U];
            bool gt = false;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const float4 r = sb[kk + u];
              A[u] = __fadd_rn(r.x, bx);
              B[u] = __fadd_rn(r.y, by);
              W[u] = screen_pm_of<D>(r);
              if (FRONTIER) gt |= A[u] > B[u];
            }
            bool alive = false;
            if (FRONTIER && !__any_sync(0xffffffffu, gt)) {
              // the whole 32 x U patch is on the S = 0 side of the A = B frontier
#pragma unroll
              for (int u = 0; u < U; ++u) alive |= !(fmaf(B[u], SCR_TINY, W[u]) < thr);
            } else {
#pragma unroll
              for (int u = 0; u < U; ++u) alive |= !(ScreenUB<DIST, RP>::eval(A[u], B[u], W[u]) < thr);
            }
            if (AUDIT) {
#pragma unroll 1
              for (int u = 0; u < U; ++u) cell(kb + kk + u, sb[kk + u], true, F_{});
            } else if (__any_sync(0xffffffffu, alive)) {
              // a survivor: the group lies in the neighbourhood of the row maximum, where most cells survive --
              // evaluate all U exactly (independent chains, loads issued together), then take them in ascending k
              D v[U];
#pragma unroll
              for (int u = 0; u < U; ++u) v[u] = exact_at(kb + kk + u);
#pragma unroll
              for (int u = 0; u < U; ++u)
                if (v[u] > best) {
                  best = v[u];
                  arg = kb + kk + u;
                }
              thr = fmaxf(thr, screen_rd<D>(best));
              nexact += U;
            }
          }
        }
      }
    }
    if (sn >= 0) {
      sbuf[buf ^ 1][tid] = nx;
      stage_max(nx, buf ^ 1);
      if (DBL) cp_async_wait_all();
    }
    buf ^= 1;
    s = sn;
  }
  if (i < g.nrows) {
    part_val[(long long)slot * g.n + i] = best;
    part_arg[(long long)slot * g.n + i] = arg;
  }
  if (counters) {
    if ((tid & 31) == 0 && nexact) atomicAdd(&counters[0], (unsigned long long)nexact);
    if (AUDIT) {
      for (int o = 16; o > 0; o >>= 1) nviol += __shfl_xor_sync(0xffffffffu, nviol, o);
      if ((tid & 31) == 0 && nviol) atomicAdd(&counters[1], (unsigned long long)nviol);
    }
  }
}

// Outputs of a level, written by screen_fold_kernel (what combine_kernel writes on the exhaustive path): maxScore_[j], nextStart_[j], the p field of the next level's records, its float image pm (double
// data), and for row-sharded runs the value in block-cyclic chunk order.
template <typename D>
struct LevelOut {
  D* ms_row;               // maxScore_[j] of instance 0 (nullptr: not kept)
  int* ns_row;             // nextStart_[j] of instance 0
  long long table_stride;  // elements between instances in ms / ns
  Rec4<D>* rec_next;       // next level's records (p field), instance 0
  float* pm_next;          // float4 screening records of the next level viewed as floats (.w), or nullptr
  D* shard_chunk;          // row-sharded runs: owned rows in block-cyclic order, or nullptr
  int* shard_arg;          // row-sharded runs: the split indices in the same order, or nullptr
};

// Fold the partials of every row block over its live tiles (ascending chunks = ascending k; strict > keeps the first
// maximum, exactly like the reference's scan over k) and write the level's outputs.  CW lanes share a row: lane q folds the
// q-th part of the segment, the parts are merged in ascending order with the same strict >.
template <typename D, int BR, int CW>
__global__ void __launch_bounds__(BR * CW)
screen_fold_kernel(Geom g, const int* __restrict__ rb_slots, const int* __restrict__ rb_cnt, int rb_stride,
                   const D* __restrict__ part_val, const int* __restrict__ part_arg, LevelOut<D> out) {
  const int inst = blockIdx.y;
  const int rb = g.rb_first + g.rb_step * (int)blockIdx.x;
  const int r = threadIdx.x / CW, q = threadIdx.x % CW;
  const int i = rb * BR + r;
  const long long pstride = (long long)g.nch_alloc * g.nsub * g.n;
  part_val += inst * pstride;
  part_arg += inst * pstride;
  rb_slots += ((long long)inst * rb_stride + rb) * g.nch_alloc * g.nsub;
  const int cnt = rb_cnt[(long long)inst * rb_stride + rb];
  D best = Lim<D>::lowest();
  int cbest = -1;
  if (i < g.nrows) {
    const int per = (cnt + CW - 1) / CW;
    const int e0 = q * per, e1 = min(cnt, e0 + per);
    for (int e = e0; e < e1; e += 4) {
      int sl[4];
      D v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) sl[u] = (e + u < e1) ? rb_slots[e + u] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (sl[u] >= 0) ? part_val[(long long)sl[u] * g.n + i] : Lim<D>::lowest();
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v[u] > best) {
          best = v[u];
          cbest = sl[u];
        }
    }
  }
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
  if (q != 0) return;
  int arg = -1;
  if (i < g.nrows) {
    // the split index is read once, for the winning chunk only (a partial equal to lowest() never wins: split -1)
    arg = (mc >= 0) ? part_arg[(long long)mc * g.n + i] : -1;
    if (out.ms_row) out.ms_row[inst * out.table_stride + i] = mb;
    if (out.ns_row) out.ns_row[inst * out.table_stride + i] = arg;
    if (out.rec_next) out.rec_next[inst * g.rec_stride + i].p = mb;
    if (out.pm_next) out.pm_next[(inst * g.rec_stride + i) * 4 + 3] = screen_pm<D>(mb);
  }
  if (out.shard_chunk) {
    const int m = (rb - g.rb_first) / g.rb_step;
    out.shard_chunk[(long long)m * BR + r] = (i < g.nrows) ? mb : Lim<D>::lowest();
    if (out.shard_arg) out.shard_arg[(long long)m * BR + r] = arg;
  }
}

template <typename D, int DIST, bool RP, bool AUDIT, int BR>
__global__ void __launch_bounds__(BR, 8)
level_live_screen_kernel(Geom g, const Rec4<D>* __restrict__ rec, const float4* __restrict__ recf,
                         const D* __restrict__ LB, D* __restrict__ part_val, int* __restrict__ part_arg,
                         unsigned long long* __restrict__ counters, const int* __restrict__ dev_flags,
                         const int4* __restrict__ live, long long live_cap, LiveHdr* __restrict__ hdr,
                         unsigned long long* __restrict__ dbg) {
  static_assert(BR == SCR_BLOCK, "a stage is one 128-block of k");
  constexpr bool DBL = (sizeof(D) == 8);
  constexpr bool FRONTIER = !RP && DIST != DIST_RATIONAL;
  constexpr bool GROUPB = !(DIST == DIST_POISSON && RP);
  constexpr bool BOX = !GROUPB;
  __shared__ float4 sbuf[2][BR];
  __shared__ float s_pmax[2][BR / 32];  // per stage and warp: largest pm among the records that warp staged
  __shared__ Rec4<D> sdbl[DBL ? 2 * BR : 1];  // double data: the level records of the stage in flight / the next one
  constexpr bool kLogTab = DBL && (DIST == DIST_POISSON);
  __shared__ GlogEntry s_log[kLogTab ? 128 : 1];
  __shared__ int s_next;
  if (kLogTab) {
    if (threadIdx.x < 128) s_log[threadIdx.x] = g_glog_tab[threadIdx.x];
  }
  // range-tier mode (screen_range_ub): the box bound needs no sortedness, everything else ascending priorities
  const bool sorted = dev_flags && (dev_flags[1] == 0);
  const int mono = BOX ? (sorted ? 2 : 1) : (((FRONTIER || GROUPB) && sorted) ? 1 : 0);
  const long long pstride = (long long)g.nch_alloc * g.nsub * g.n;
  const int tid = threadIdx.x;
  // class boundaries in the order the tiles are taken (final: the select kernel of this level has completed)
  int cum[SCR_CLASSES];
  {
    int acc = 0;
#pragma unroll
    for (int q = 0; q < SCR_CLASSES; ++q) {
      acc += hdr->nlive[q];
      cum[q] = acc;
    }
  }
  const int nlive = cum[SCR_CLASSES - 1];
  for (;;) {
    __syncthreads();  // the previous tile is done with the shared buffers and with s_next
    if (tid == 0) s_next = atomicAdd(&hdr->cursor, 1);  // persistent CTAs pull the next tile
    __syncthreads();
    const int t = s_next;
    if (t >= nlive) break;
    int cls = 0;
#pragma unroll
    for (int q = 0; q < SCR_CLASSES - 1; ++q) cls += (t >= cum[q]) ? 1 : 0;
    const int4 e = live[(long long)cls * live_cap + (t - (cls ? cum[cls - 1] : 0))];
    const int inst = e.x, rb = e.y, slot = e.z;
    unsigned long long t0 = 0;  // dbg: development aid (PS_DEBUG_LEVEL), per-tile start / end / SM / class
    if (dbg && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    screen_live_tile<D, DIST, RP, AUDIT, BR>(g, inst, rb, slot, (unsigned int)e.w, rec + (long long)inst * g.rec_stride,
                                             recf + (long long)inst * g.rec_stride, LB + (long long)inst * g.n,
                                             part_val + inst * pstride, part_arg + inst * pstride, counters, mono, sbuf,
                                             s_pmax, sdbl, s_log);
    if (dbg && tid == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      dbg[4 * (long long)t + 0] = t0;
      dbg[4 * (long long)t + 1] = t1;
      dbg[4 * (long long)t + 2] = ((unsigned long long)smid << 32) | (unsigned int)cls;
      dbg[4 * (long long)t + 3] = ((unsigned long long)(unsigned int)rb << 32) | (unsigned int)slot;
    }
  }
}

// Exhaustive check of the two hardware claims the error budget rests on, over every float in the safe range:
// out[0] = max |x rcp(x) - 1| / u  over x in [2^-62, 2^62];  out[1] = max |lg2(x) - log2 x| / (u max(1, |log2 x|))
// over x in [2^-92, 2^92]  (u = 2^-24; references computed in double).
__global__ void selftest_mufu_kernel(double* __restrict__ out) {
  const unsigned int stride = gridDim.x * blockDim.x;
  double e_rcp = 0.0, e_lg2 = 0.0;
  const unsigned int lo = (127u - 92u) << 23, hi = (127u + 92u) << 23;
  for (unsigned int bits = lo + blockIdx.x * blockDim.x + threadIdx.x; bits < hi; bits += stride) {
    const float x = __uint_as_float(bits);
    const int ex = (int)(bits >> 23) - 127;
    if (ex >= -62 && ex < 62) e_rcp = fmax(e_rcp, fabs((double)x * (double)rcp_approx(x) - 1.0));
    const double l = log2((double)x);
    e_lg2 = fmax(e_lg2, fabs((double)lg2_approx(x) - l) / fmax(1.0, fabs(l)));
  }
  for (int o = 16; o > 0; o >>= 1) {
    e_rcp = fmax(e_rcp, __shfl_xor_sync(0xffffffffu, e_rcp, o));
    e_lg2 = fmax(e_lg2, __shfl_xor_sync(0xffffffffu, e_lg2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // non-negative doubles order like their bit patterns
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_rcp * 16777216.0));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_lg2 * 16777216.0));
  }
}

}  // namespace ps
