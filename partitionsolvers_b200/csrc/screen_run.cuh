// screen_run.cuh -- screened levels for double inputs whose partial sums are NOT exactly representable
// (real-valued a, b: the Gaussian / Rational workloads, tract-like expected counts).
//
// screen.cuh needs exact summability because it forms a cell's range sums as prefix differences, which equal the
// reference's left-to-right running sums (score_impl.hpp:248-256) only when no addition rounds.  Here the two jobs
// are separated:
//   * the BOUND of a cell still comes from float range sums, now derived from 64-bit FIXED-POINT prefix sums
//     (integer differences are exact, so there is no cancellation however long the prefix is);
//   * a cell that survives is evaluated from the reference's own RUNNING sums, re-created by walking the records from
//     the nearest checkpoint: the level-1 pass (level1_running_kernel) leaves every row's running sums at every
//     k-chunk boundary (SA/SB, the same arrays the exhaustive RUNNING kernel starts its tiles from).  The walk is
//     lazy and warp-uniform (lanes = adjacent rows at the same k): a warp only walks the tiles in which it has a
//     survivor, so the 2 DADD per walked cell are paid on a few per cent of the cells.
// Values and first-argmax are therefore the reference's, bit for bit, exactly as argued in screen.cuh: the first
// maximiser's bound survives, every evaluated cell uses reference arithmetic on reference sums, ascending k, strict >.
//
// Fixed point.  Per instance and array: shift s with sum|x| * 2^s < 2^61 (fx_setup_kernel);
//   a_fx = ceil(a 2^s)  (range sums of a are bounded from ABOVE),  b_fx = floor(b 2^s)  (from BELOW).
// Every bound of screen.cuh is non-decreasing in A and non-increasing in B on the region where it is used, except
// Poisson risk-part (A ln(A/B) decreases in A while A/B < 1/e), which therefore also needs the two-sided condition below.
// Conditions checked on the device (else the exhaustive kernel runs):
//   (F1) min b * 2^sb >= 2^28: |B_fx 2^-sb - B| <= 2^-28 B for every range  (0.06 u, u = 2^-24);
//   (F2) Poisson risk-part only: the same for the positive a;
//   (F3) n <= 2^22: the reference's running sum of m <= n positive terms differs from the true sum by at most
//        (m-1) 2^-53 <= 2^-31 relatively  (0.01 u);
//   (F4) at most n/8 elements with a <= 0.  Rows i < ipos = 1 + (last index with a <= 0) own range sums that need not be
//        positive: the float bounds do not apply and their tiles are evaluated exhaustively inside this kernel.
// Float range sums: At = fl(lx[k] + bx), lx = fl(2^-s (P[k] - P[o])), bx = fl(2^-s (P[o] - P[i])), o = origin of k's
// 128-block: three roundings on non-negative terms, |At - A_fx| <= 2u A_fx as in screen.cuh; with (F1)-(F3) the input
// line of the error budget becomes 2.07u (of the 2u + 2u/5u spare the margins carry).  The monotone tiers (tile, stage,
// group) are argued on the TRUE sums; the reference's values differ from the true-sum values by the (F3) term.
// Audit mode (ps_problem.screen = 2) re-evaluates every dropped cell with the reference's arithmetic, as in screen.cuh.
#pragma once
#include <cuda_runtime.h>

#include "screen.cuh"

namespace ps {

struct FxInfo {
  int sa, sb;      // shifts
  float sca, scb;  // 2^-sa, 2^-sb
  int ipos;        // rows >= ipos have only positive a to their right (written by fx_prefix_kernel)
  int ok;          // instance qualifies
  int pad[2];
};

// Per instance: shifts and the conditions (F1)-(F4).  flags[0] |= 8 if some instance does not qualify.
__global__ void fx_setup_kernel(int R, int n, int need_tight_a, const ScreenStats* __restrict__ st,
                                FxInfo* __restrict__ fx, int* __restrict__ flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const ScreenStats s = st[r];
  int ea = 0, eb = 0;
  frexp(fmax(s.sa, 1e-300), &ea);  // sa < 2^ea
  frexp(fmax(s.sb, 1e-300), &eb);
  FxInfo f;
  f.sa = 60 - ea;
  f.sb = 60 - eb;
  f.sca = (float)ldexp(1.0, -f.sa);
  f.scb = (float)ldexp(1.0, -f.sb);
  f.ipos = 0;
  const double min_a = __longlong_as_double((long long)s.min_a_pos);
  const double min_b = __longlong_as_double((long long)s.min_b);
  const double tight = 268435456.0;  // 2^28
  bool ok = (n <= (1 << 22)) && (f.sa >= -90 && f.sa <= 90) && (f.sb >= -90 && f.sb <= 90);
  ok = ok && s.min_b != 0x7ff0000000000000ULL && ldexp(min_b, f.sb) >= tight;
  if (need_tight_a) ok = ok && s.min_a_pos != 0x7ff0000000000000ULL && ldexp(min_a, f.sa) >= tight;
  ok = ok && (s.n_nonpos <= n / 8);
  f.ok = ok ? 1 : 0;
  f.pad[0] = f.pad[1] = 0;
  fx[r] = f;
  if (!ok) atomicOr(flags, 8);
}

// Fixed-point prefix sums of the SORTED arrays: pfx[k] = {sum_{j<k} ceil(a_j 2^sa), sum_{j<k} floor(b_j 2^sb)}.
// Integer addition is associative: any scan order gives the same (exact) result.  One CTA per instance, every thread owns
// a contiguous slice.  Also ipos.
template <int NT>
__global__ void __launch_bounds__(NT)
fx_prefix_kernel(int n, long long rec_stride, const double* __restrict__ as, const double* __restrict__ bs,
                 FxInfo* __restrict__ fx, longlong2* __restrict__ pfx) {
  __shared__ long long sa[NT], sb[NT];
  __shared__ int s_ipos;
  const int inst = blockIdx.x;
  as += (long long)inst * n;
  bs += (long long)inst * n;
  pfx += inst * rec_stride;
  const int tid = threadIdx.x;
  if (tid == 0) s_ipos = 0;
  const int sha = fx[inst].sa, shb = fx[inst].sb;
  const int per = (n + NT - 1) / NT;
  const int lo = min(n, tid * per), hi = min(n, lo + per);
  long long ta = 0, tb = 0;
  int ip = 0;
  for (int k = lo; k < hi; ++k) {
    const double av = as[k];
    ta += __double2ll_ru(scalbn(av, sha));
    tb += __double2ll_rd(scalbn(bs[k], shb));
    if (!(av > 0.0)) ip = k + 1;
  }
  sa[tid] = ta;
  sb[tid] = tb;
  __syncthreads();
  if (ip) atomicMax(&s_ipos, ip);
  if (tid == 0) {
    long long ra = 0, rb = 0;
    for (int q = 0; q < NT; ++q) {
      const long long xa = sa[q], xb = sb[q];
      sa[q] = ra;
      sb[q] = rb;
      ra += xa;
      rb += xb;
    }
  }
  __syncthreads();
  long long ra = sa[tid], rb = sb[tid];
  if (tid == 0) {
    pfx[0] = make_longlong2(0, 0);
    fx[inst].ipos = s_ipos;
  }
  for (int k = lo; k < hi; ++k) {
    ra += __double2ll_ru(scalbn(as[k], sha));
    rb += __double2ll_rd(scalbn(bs[k], shb));
    pfx[k + 1] = make_longlong2(ra, rb);
  }
}

// Float screening records {lx, ly, gmax, pm} from the fixed-point prefix sums (both level buffers); lx, ly relative to the
// origin of k's 128-block as in build_recf_kernel.
__global__ void build_recf_fx_kernel(int n, long long rec_stride, const longlong2* __restrict__ pfx,
                                     const FxInfo* __restrict__ fx, float4* __restrict__ recf0,
                                     float4* __restrict__ recf1) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  pfx += inst * rec_stride;
  const FxInfo F = fx[inst];
  const int o = (k > 0) ? ((k - 1) / SCR_BLOCK) * SCR_BLOCK : 0;
  const longlong2 r = pfx[k], q = pfx[o];
  const float4 f = make_float4(__ll2float_rn(r.x - q.x) * F.sca, __ll2float_rn(r.y - q.y) * F.scb, 0.f, 0.f);
  recf0[inst * rec_stride + k] = f;
  recf1[inst * rec_stride + k] = f;
}

// pm of level 1 (level1_running_kernel wrote the level values into the p field of the records)
__global__ void pm_from_rec_kernel(int n, long long rec_stride, const Rec4<double>* __restrict__ rec,
                                   float4* __restrict__ recf) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  recf[inst * rec_stride + k].w = screen_pm<double>(rec[inst * rec_stride + k].p);
}

// Row lower bounds for the running-sum path: the cells whose running sums are at hand are the chunk boundaries
// (i, c L), c L > i (at most ~200 per row) -- a coarse sweep over them, then every boundary around the best coarse sample.  Also the group / block maxima of pm, as in
// screen_lb_kernel.
template <int DIST, bool RP>
__global__ void __launch_bounds__(128)
screen_lb_run_kernel(Geom g, const Rec4<double>* __restrict__ rec, float4* __restrict__ recf,
                     const double* __restrict__ SA, const double* __restrict__ SB, double* __restrict__ LB,
                     float* __restrict__ bmax, int nblk, int* __restrict__ kbest = nullptr) {
  using D = double;
  const int inst = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nrows) return;
  rec += (long long)inst * g.rec_stride;
  recf += (long long)inst * g.rec_stride;
  const long long pstride = (long long)g.nch_alloc * g.n;
  if ((i & (SCR_BLOCK - 1)) == 0) {
    float m = -3.402823466e+38f;
    const int ke = min(i + SCR_BLOCK, g.n);
    for (int k = i + 1; k <= ke; ++k) m = fmaxf(m, recf[k].w);
    bmax[(long long)inst * nblk + i / SCR_BLOCK] = m;
  }
  if ((i & 7) == 0 && i + 8 <= g.n) {
    float m = recf[i + 1].w;
#pragma unroll
    for (int q = 2; q <= 8; ++q) m = fmaxf(m, recf[i + q].w);
    recf[i + 8].z = m;
  }
  {  // row-sharded runs: the maxima above are needed for every k, lower bounds (and checkpoints) exist for owned rows only
    const int rbi = i / SCR_BLOCK;
    if (rbi < g.rb_first || (rbi - g.rb_first) % g.rb_step != 0) return;
  }
  D best = Lim<D>::lowest();
  int cb = i / g.L;  // chunk of the best sample (the row's own chunk if no boundary lies right of it): the cost hint
  if (SA) {
    const int c0 = i / g.L + 1, c1 = min(g.kmax / g.L, g.nch_alloc - 1);
    cb = c0;
    auto probe = [&](const int c) {
      const D A = SA[inst * pstride + (long long)c * g.n + i];
      const D B = SB[inst * pstride + (long long)c * g.n + i];
      const D v = add_rn<D>(screen_exact_score<D, DIST, RP, false>(A, B, nullptr), rec[c * g.L].p);
      if (v > best) {
        best = v;
        cb = c;
      }
    };
    if (c1 >= c0) {  // a coarse sweep over the row's boundaries, then every boundary around the best coarse sample
      const int sc = max(1, (c1 - c0 + 1) / 24);
      for (int c = c0; c <= c1; c += sc) probe(c);
      probe(c1);
      if (sc > 1) {
        const int lo = max(c0, cb - sc + 1), hi = min(c1, cb + sc - 1);
        for (int c = lo; c <= hi; ++c) probe(c);
      }
    }
  }
  LB[(long long)inst * g.n + i] = best;
  if (kbest) kbest[(long long)inst * g.n + i] = max(1, min(cb * g.L, g.kmax));
}

// =====================================================================================================================
// FLOAT inputs whose sums round.  The reference's float running sums carry up to m u of rounding after m terms, so no
// bound derived from the TRUE sums can be tight.  Here every sum a bound is formed from IS a reference running sum:
//   * the level-1 pass leaves every row's running sums at EVERY 128-block boundary (checkpoints, 8 n^2/128 bytes);
//   * the tile and stage tiers take the row's sums at the end of the tile / stage straight from the checkpoints;
//   * a stage that survives is walked from its starting checkpoint (one packed FADD2 per cell), group by group; the group
//     tier bounds at the group's last cell, the cell tier at every cell, survivors are evaluated on the spot.
// Every score is supported (run32_range_ub).  Monotonicity on the running sums themselves, A^2/B family: adding an
// element with priority r to sums (A, B), q = A/B, changes A^2/B by b (2 q r - q^2) >= 0 whenever r >= q/2; the running
// ratio exceeds the true ratio of the range (<= r, priorities ascending) by at most a factor 1 + 2 m u < 2.  The m' additions
// between a cell and the cell the bound is formed at round: S(k') >= S(k) (1 - 3 m' u), carried as 1 + 64u (group, m' <= 7),
// 1 + 512u (stage) and 1 + (3 L + 64) u (tile).  The cell tier needs no monotonicity: its inputs are exact, the margins of
// screen.cuh cover MUFU.RCP, the bound's own roundings and the reference's.  Rows i < ipos (some a <= 0 to their right) take
// no bounds.  Audit mode re-evaluates every dropped cell.
// =====================================================================================================================
// Bound for every cell of a k-range of row i from the row's RUNNING sums (A, B) at the range's last cell k', the largest pm
// of the range, span = number of additions between the range's first and last cell, len = k' - i.
// Poisson risk-partitioning, (Af, Bf) = the row's running sums at or before the range's first cell:
//   * priorities not known to ascend: the box (ScreenUB<DIST_POISSON, true>::box).  Adding a >= 0 / b >= 0 to a float never
//     lowers it (rounding is monotone), so every cell of the range has A in [Af, A] and B >= Bf exactly -- no span term;
//   * ascending priorities (`sorted`): the larger END cell (screen.cuh: hull).  Continue the running sums from (Af, Bf) with the
//     elements added EXACTLY: along that path f is quasi-convex whatever the starting ratio (the persistence step of the proof
//     only needs r' >= r and q' between q and r), so it is bounded by its two ends.  The float path strays from it by at most
//     m u relatively after m additions, and |df| <= m u A (|1 + ln q| + 1); the same at the far end.  Carried:
//     2 span u A (max |ln q| + 2) on top of the cell bound's margins.  At span = 8 that is 2.4e-6 A where the box is loose by
//     A dB/B ~ 1e-4 A.
template <int DIST, bool RP>
__device__ __forceinline__ float run32_range_ub(float Af, float Bf, float A, float B, float pm, float span, float len,
                                                const bool sorted = false) {
  if constexpr (DIST == DIST_POISSON && RP) {
    if (!sorted) return ScreenUB<DIST_POISSON, true>::box(Af, Bf, A, pm);
    const float l1 = lg2_approx(Af * rcp_approx(Bf)), l2 = lg2_approx(A * rcp_approx(B));
    const float t1 = Af * l1, t2 = A * l2;
    const float e1 = fmaf(fabsf(t1), 0.693147182464599609375f * 16.0f * SCR_U, t1 * 0.693147182464599609375f);
    const float e2 = fmaf(fabsf(t2), 0.693147182464599609375f * 16.0f * SCR_U, t2 * 0.693147182464599609375f);
    const float lq = fmaf(fmaxf(fabsf(l1), fabsf(l2)), 0.6931472f, 2.05f);  // max |ln q| + 2, a little more
    const float stray = (2.02f * SCR_U * span) * (A * lq);
    return fmaxf(e1, e2) + fmaf(A, SCR_CA, pm) + stray;
  } else if constexpr (DIST == DIST_GAUSSIAN && !RP) {
    // g = (A - B)^2 / (2B) for A > B, else 0.  Adding an element of priority r raises g whenever r >= (q + 1) / 2.  Along the
    // exact continuation of the running sums q <= r (1 + 2 gam), gam = 1.01 len u, so a decrease needs q < 1 + 4.1 gam, where
    // g <= 8.4 gam^2 B: every cell of the range has g <= g(end) + 8.4 gam^2 B(end).  The span roundings move (A, B) by span u
    // relatively: |dg| <= 2 span u A (q - 1), and at most 2 (span u)^2 B where the sign of A - B flips.  Carried: 12 gam^2 B
    // and (8 + 2 span + 8) u on A (q - 1).
    const float gam = len * (1.01f * SCR_U);
    const float slack = 12.0f * gam * gam;
    if (A > B) {
      const float dif = A - B;
      const float h = dif * rcp_approx(B);
      const float gt = (0.5f * dif) * h;
      float e = fmaf(A * h, (2.0f * span + 16.0f) * SCR_U, pm);
      e = fmaf(A, SCR_TINY, e);
      e = fmaf(B, slack, e);
      return fmaf(gt, SCR_REL_UP, e);
    }
    return fmaf(B, SCR_TINY + slack, pm);
  } else if constexpr (DIST == DIST_POISSON && !RP) {
    // g = A ln(A/B) + B - A for A > B, else 0: convex in (A, B) with gradient (ln q, 1 - q), so adding an element (a, b) of
    // priority r changes it by at least b (r ln q + 1 - q), which is >= 0 whenever r >= (q - 1) / ln q (< q).  Along the exact
    // continuation of the running sums q <= r (1 + 2 gam), gam = 1.01 len u; with q = 1 + h, q ln q / (1 + 2 gam) + 1 - q ~
    // h^2/2 - 2 gam h, so a decrease needs h < 4.1 gam, where g <= (A - B)^2 / 2B <= 8.4 gam^2 B (the Gaussian score dominates
    // the Poisson one): every cell of the range has g <= g(end) + 8.4 gam^2 B(end) -- the argument of the Gaussian branch above.
    // The span roundings move (A, B) by span u relatively: |dg| <= span u (A ln q + A - B) <= 2 span u A (q - 1).  The bound at
    // the range's last cell is the cell bound itself (screen.cuh: it already covers MUFU, its own roundings and the reference's
    // at that cell; the reference's rounding at an earlier cell, 2u A + 6u A ln q with A, q no larger than at the end up to
    // 2 gam, sits in the cell bound's 5u of spare margin on A).  Carried on top: 12 gam^2 B and (2 span + 16) u on A (q - 1).
    const float gam = len * (1.01f * SCR_U);
    const float slack = 12.0f * gam * gam;
    if (A > B) {
      const float h = (A - B) * rcp_approx(B);
      float e = (A * h) * ((2.0f * span + 16.0f) * SCR_U);
      e = fmaf(B, slack, e);
      return ScreenUB<DIST_POISSON, false>::eval(A, B, pm) + e;
    }
    return fmaf(B, SCR_TINY + slack, pm);
  } else {
    // A^2/B family: S(k') >= S(k) (1 - 3 span u) on the running sums (header above)
    return ScreenUB<DIST, RP>::eval(A, B, pm) * (1.0f + (3.0f * span + 40.0f) * SCR_U);
  }
}

__global__ void ipos_kernel(int n, const float* __restrict__ as, int* __restrict__ ipos) {
  const int inst = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int v = 0;
  if (k < n && !(as[(long long)inst * n + k] > 0.f)) v = k + 1;
  v = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && v) atomicMax(&ipos[inst], v);
}

// Row lower bounds from checkpoint cells (i, 128 m): a coarse sweep over the row's blocks, then every block around the best.
template <int DIST, bool RP>
__global__ void __launch_bounds__(128)
screen_lb_run32_kernel(Geom g, const Rec4<float>* __restrict__ rec, float4* __restrict__ recf,
                       const float* __restrict__ SA, const float* __restrict__ SB, long long cp_stride,
                       float* __restrict__ LB, float* __restrict__ bmax, int nblk, int* __restrict__ kbest = nullptr) {
  using D = float;
  const int inst = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nrows) return;
  rec += (long long)inst * g.rec_stride;
  recf += (long long)inst * g.rec_stride;
  SA += inst * cp_stride;
  SB += inst * cp_stride;
  if ((i & (SCR_BLOCK - 1)) == 0) {
    float m = -3.402823466e+38f;
    const int ke = min(i + SCR_BLOCK, g.n);
    for (int k = i + 1; k <= ke; ++k) m = fmaxf(m, screen_pm_of<D>(recf[k]));
    bmax[(long long)inst * nblk + i / SCR_BLOCK] = m;
  }
  if ((i & 7) == 0 && i + 8 <= g.n) {
    float m = screen_pm_of<D>(recf[i + 1]);
#pragma unroll
    for (int q = 2; q <= 8; ++q) m = fmaxf(m, screen_pm_of<D>(recf[i + q]));
    recf[i + 8].w = m;
  }
  {  // row-sharded runs: lower bounds (and checkpoints) for owned rows only
    const int rbi = i / SCR_BLOCK;
    if (rbi < g.rb_first || (rbi - g.rb_first) % g.rb_step != 0) return;
  }
  const int m0 = i / SCR_BLOCK + 1, m1 = g.kmax / SCR_BLOCK;  // boundaries 128 m in (i, kmax]
  D best = Lim<D>::lowest();
  int mb = m0;
  auto probe = [&](int m) {
    const D A = SA[(long long)m * g.n + i], B = SB[(long long)m * g.n + i];
    const D v = add_rn<D>(screen_exact_score<D, DIST, RP, false>(A, B, nullptr), rec[m * SCR_BLOCK].p);
    if (v > best) {
      best = v;
      mb = m;
    }
  };
  if (m1 >= m0) {
    const int len = m1 - m0 + 1;
    const int sc = max(1, len / 32);
    for (int m = m0; m <= m1; m += sc) probe(m);
    probe(m1);
    if (sc > 1) {
      const int lo = max(m0, mb - sc), hi = min(m1, mb + sc);
      for (int m = lo; m <= hi; ++m) probe(m);
    }
  }
  LB[(long long)inst * g.n + i] = best;
  if (kbest) kbest[(long long)inst * g.n + i] = max(1, min(mb * SCR_BLOCK, g.kmax));
}

// =====================================================================================================================
// Live-tile protocol (screen.cuh, "Screened level = four launches") for the running-sum screens: the range tiers have run in
// screen_select_kernel<.., SEL_FIXED / SEL_RUN32>; persistent CTAs pull live tiles from the cost-ordered lists and run the
// group and cell tiers on the live stages; screen_fold_kernel folds the partial slots.
// =====================================================================================================================
// One live tile, double data whose sums round: bounds from the fixed-point prefix sums, survivors from the reference's running
// sums (lazy warp-uniform walk from the chunk-boundary checkpoint SA/SB, which also passes through stages that are not live).
template <int DIST, bool RP, bool AUDIT, int BR>
__device__ __forceinline__ void screen_run_live_tile(const Geom& g, const int rb, const int slot,
                                                     const unsigned int stage_mask, const Rec4<double>* __restrict__ rec,
                                                     const float4* __restrict__ recf, const longlong2* __restrict__ pfx,
                                                     const FxInfo& F, const double* __restrict__ SA,
                                                     const double* __restrict__ SB, const double* __restrict__ LB,
                                                     double* __restrict__ part_val, int* __restrict__ part_arg,
                                                     unsigned long long* __restrict__ counters, const int mono_in,
                                                     float4 (*sbuf)[BR], float (*s_pmax)[BR / 32],
                                                     const GlogEntry* __restrict__ s_log) {
  using D = double;
  constexpr int BK = BR;
  constexpr int U = 8;
  constexpr int GQ = AUDIT ? 1 : 4;  // groups whose bounds are formed together (audit runs keep ascending k for the walk)
  constexpr bool FRONTIER = !RP && DIST != DIST_RATIONAL;
  constexpr bool kLogTab = (DIST == DIST_POISSON);
  const int tid = threadIdx.x;
  const int i0 = rb * BR;
  const int i = i0 + tid;
  const int c = slot / g.nsub;
  const int k_lo = max(c * g.L + 1, i0 + 1);
  const int k_hi = min((c + 1) * g.L, g.kmax);
  const bool diag_tile = (c == rb / g.G);
  const int il = min(i, g.n - 1);
  const longlong2 Pi = pfx[il];
  float bx = 0.f, by = 0.f;
  D best = Lim<D>::lowest();
  int arg = -1;
  const float thr0 = screen_rd<D>(LB[min(il, max(g.nrows - 1, 0))]);  // the threshold the select kernel dropped with
  float thr = thr0;
  unsigned int nexact = 0, nviol = 0;
  const bool forced = i0 < F.ipos;  // (F4) a row of this block may own non-positive sums: no bounds, every cell exactly
  const int mono = forced ? 0 : mono_in;  // range-tier mode (screen_range_ub)

  // the reference's running sums of row i through cell kw (warp-uniform kw); a tile right of the diagonal starts from the
  // checkpoint the level-1 pass left at its chunk boundary (loaded on first use)
  D Ar = 0, Br = 0;
  int kw = diag_tile ? i0 : c * g.L;
  bool loaded = diag_tile;
  auto advance_to = [&](const int k) {
    if (!loaded) {
      Ar = SA[(long long)c * g.n + il];
      Br = SB[(long long)c * g.n + il];
      loaded = true;
    }
    int q = kw + 1;
    if (q <= i0 + BR - 1) {  // diagonal stage: row i starts at k = i + 1
      const int qe = min(k, i0 + BR - 1);
      for (; q <= qe; ++q) {
        const double2 ab = *reinterpret_cast<const double2*>(&rec[q]);
        if (q > i) {
          Ar = add_rn<D>(Ar, ab.x);
          Br = add_rn<D>(Br, ab.y);
        }
      }
    }
#pragma unroll 4
    for (; q <= k; ++q) {
      const double2 ab = *reinterpret_cast<const double2*>(&rec[q]);
      Ar = add_rn<D>(Ar, ab.x);
      Br = add_rn<D>(Br, ab.y);
    }
    kw = max(kw, k);
  };
  // calls are warp-uniform and ascending in k
  auto exact_at = [&](const int k) -> D {
    advance_to(k);
    return add_rn<D>(screen_exact_score<D, DIST, RP, kLogTab>(Ar, Br, s_log), rec[k].p);
  };
  auto audit_dropped = [&](const int k, const bool valid, const float t) {
    const D v = exact_at(k);
    if (valid && !(v < (double)t)) ++nviol;
  };
  auto fsum = [&](const longlong2 P, float& A, float& B) {
    A = __ll2float_rn(P.x - Pi.x) * F.sca;
    B = __ll2float_rn(P.y - Pi.y) * F.scb;
  };
  auto cell = [&](const int k, const float4 r, const bool valid, auto from_fixed) {
    float A, B;
    if constexpr (decltype(from_fixed)::value) {
      fsum(pfx[k], A, B);
    } else {
      A = __fadd_rn(r.x, bx);
      B = __fadd_rn(r.y, by);
    }
    const float ub = ScreenUB<DIST, RP>::eval(A, B, r.w);
    const bool pass = valid && (forced || !(ub < thr));
    if (__any_sync(0xffffffffu, pass)) {
      const D v = exact_at(k);
      if (AUDIT && valid && !forced && (double)ub < v) ++nviol;
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
  auto stage_max = [&](const float4 r, const int buf) {
    float w = ((tid & 7) == 7) ? r.z : -3.402823466e+38f;
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 8));
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 16));
    if ((tid & 31) == 31) s_pmax[buf][tid >> 5] = w;
  };

  const int nk = k_hi - k_lo + 1;
  const int nstages = (nk + BK - 1) / BK;
  if (AUDIT && counters && tid == 0) atomicAdd(&counters[stage_mask ? 2 : 3], 1ULL);
  unsigned int todo = AUDIT ? (nstages >= 32 ? 0xffffffffu : ((1u << nstages) - 1u)) : stage_mask;
  int s = todo ? (__ffs(todo) - 1) : -1;
  longlong2 onx = make_longlong2(0, 0);  // fixed-point prefix at the origin of the staged stage (k = kb - 1)
  if (s >= 0) {
    const int kb = k_lo + s * BK;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kb + tid <= k_hi) r = recf[kb + tid];
    sbuf[0][tid] = r;
    stage_max(r, 0);
    onx = pfx[kb - 1];
  }
  int buf = 0;
  while (s >= 0) {
    __syncthreads();
    todo &= todo - 1u;
    const int sn = todo ? (__ffs(todo) - 1) : -1;
    const int kb = k_lo + s * BK;
    float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sn >= 0 && k_lo + sn * BK + tid <= k_hi) nx = recf[k_lo + sn * BK + tid];
    fsum(onx, bx, by);  // >= 0 on every stage right of the diagonal
    if (sn >= 0) onx = pfx[k_lo + sn * BK - 1];
    const int cnt = min(BK, k_hi - kb + 1);
    const float4* __restrict__ sb = sbuf[buf];
    const bool sel_dropped = AUDIT && !((stage_mask >> s) & 1u);
    if (sel_dropped) {
      if (counters && (tid & 31) == 0) atomicAdd(&counters[5], 1ULL);
      for (int kk = 0; kk < cnt; ++kk) audit_dropped(kb + kk, (kb + kk > i) && (i < g.nrows), thr0);
    } else if (forced) {
      for (int kk = 0; kk < cnt; ++kk) {
        const D v = exact_at(kb + kk);
        ++nexact;
        if (kb + kk > i && v > best) {
          best = v;
          arg = kb + kk;
        }
      }
    } else if (kb <= i0 + BR - 1) {
      for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], kb + kk > i, T_{});
    } else if (cnt < BK) {
      for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], true, F_{});
    } else {
      bool skip = false;
      if (mono) {
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
          unsigned int keepm = (1u << GQ) - 1u;
          if (mono) {
            float ubg[GQ];
#pragma unroll
            for (int q = 0; q < GQ; ++q) {
              const float4 re = sb[kq + q * U + U - 1], rf = sb[kq + q * U];
              ubg[q] = screen_range_ub<DIST, RP>(__fadd_rn(rf.x, bx), __fadd_rn(rf.y, by), __fadd_rn(re.x, bx),
                                                 __fadd_rn(re.y, by), re.z, mono);
            }
            keepm = 0;
#pragma unroll
            for (int q = 0; q < GQ; ++q)
              if (__any_sync(0xffffffffu, !(ubg[q] < thr))) keepm |= 1u << q;
            if (AUDIT) {  // GQ == 1
              const bool gdrop = !(keepm & 1u);
              if (counters && (tid & 31) == 0) atomicAdd(&counters[gdrop ? 7 : 6], 1ULL);
              if (gdrop)
                for (int u = 0; u < U; ++u) audit_dropped(kb + kq + u, true, thr);
            }
            if (!keepm) continue;
          }
#pragma unroll 1
          for (int q = 0; q < GQ; ++q) {
            if (!((keepm >> q) & 1u)) continue;
            const int kk = kq + q * U;
            float A[U], B[U], W[U];
            bool gt = false;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const float4 r = sb[kk + u];
              A[u] = __fadd_rn(r.x, bx);
              B[u] = __fadd_rn(r.y, by);
              W[u] = r.w;
              if (FRONTIER) gt |= A[u] > B[u];
            }
            bool alive = false;
            if (FRONTIER && !__any_sync(0xffffffffu, gt)) {
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
              // a survivor: walk to the group, then its U cells -- the sums are sequential (2 DADD per cell), the score
              // evaluations that hang off them are independent chains
              advance_to(kb + kk - 1);
              D v[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const Rec4<D> R = rec[kb + kk + u];
                Ar = add_rn<D>(Ar, R.x);
                Br = add_rn<D>(Br, R.y);
                v[u] = add_rn<D>(screen_exact_score<D, DIST, RP, kLogTab>(Ar, Br, s_log), R.p);
              }
              kw = kb + kk + U - 1;
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

template <int DIST, bool RP, bool AUDIT, int BR>
__global__ void __launch_bounds__(BR, 8)
level_live_screen_run_kernel(Geom g, const Rec4<double>* __restrict__ rec, const float4* __restrict__ recf,
                             const longlong2* __restrict__ pfx, const FxInfo* __restrict__ fx,
                             const double* __restrict__ SA, const double* __restrict__ SB,
                             const double* __restrict__ LB, double* __restrict__ part_val, int* __restrict__ part_arg,
                             unsigned long long* __restrict__ counters, const int* __restrict__ dev_flags,
                             const int4* __restrict__ live, long long live_cap, LiveHdr* __restrict__ hdr) {
  static_assert(BR == SCR_BLOCK, "a stage is one 128-block of k");
  constexpr bool FRONTIER = !RP && DIST != DIST_RATIONAL;
  constexpr bool GROUPB = !(DIST == DIST_POISSON && RP);
  constexpr bool BOX = !GROUPB;
  __shared__ float4 sbuf[2][BR];
  __shared__ float s_pmax[2][BR / 32];
  constexpr bool kLogTab = (DIST == DIST_POISSON);
  __shared__ GlogEntry s_log[kLogTab ? 128 : 1];
  __shared__ int s_next;
  if (kLogTab) {
    if (threadIdx.x < 128) s_log[threadIdx.x] = g_glog_tab[threadIdx.x];
  }
  const bool sorted = dev_flags && (dev_flags[1] == 0);
  const int mono = BOX ? (sorted ? 2 : 1) : (((FRONTIER || GROUPB) && sorted) ? 1 : 0);  // range-tier mode (screen_range_ub)
  const long long pstride = (long long)g.nch_alloc * g.nsub * g.n;  // partial slots
  const long long sstride = (long long)g.nch_alloc * g.n;           // SA / SB checkpoints
  const int tid = threadIdx.x;
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
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(&hdr->cursor, 1);
    __syncthreads();
    const int t = s_next;
    if (t >= nlive) break;
    int cls = 0;
#pragma unroll
    for (int q = 0; q < SCR_CLASSES - 1; ++q) cls += (t >= cum[q]) ? 1 : 0;
    const int4 e = live[(long long)cls * live_cap + (t - (cls ? cum[cls - 1] : 0))];
    const int inst = e.x;
    const FxInfo F = fx[inst];
    screen_run_live_tile<DIST, RP, AUDIT, BR>(g, e.y, e.z, (unsigned int)e.w, rec + (long long)inst * g.rec_stride,
                                              recf + (long long)inst * g.rec_stride, pfx + (long long)inst * g.rec_stride,
                                              F, SA + inst * sstride, SB + inst * sstride, LB + (long long)inst * g.n,
                                              part_val + inst * pstride, part_arg + inst * pstride, counters, mono, sbuf,
                                              s_pmax, s_log);
  }
}

// One live tile, float data whose sums round: every sum a bound is formed from IS a reference running sum.  A live stage
// starts from the row's checkpoint at its first boundary (the level-1 pass left one at every 128-block boundary), so the
// stages of a tile are independent of each other -- stages that are not live are never touched.
template <int DIST, bool RP, bool AUDIT, int BR>
__device__ __forceinline__ void screen_run32_live_tile(const Geom& g, const int rb, const int slot,
                                                       const unsigned int stage_mask, const Rec4<float>* __restrict__ rec,
                                                       const float* __restrict__ SA, const float* __restrict__ SB,
                                                       const bool forced, const float* __restrict__ LB,
                                                       float* __restrict__ part_val, int* __restrict__ part_arg,
                                                       unsigned long long* __restrict__ counters, const int mono,
                                                       float4 (*sbuf)[BR], float (*s_pmax)[BR / 32]) {
  using D = float;
  constexpr int BK = BR;
  constexpr int U = 8;
  const float4* __restrict__ recf = reinterpret_cast<const float4*>(rec);
  const int tid = threadIdx.x;
  const int i0 = rb * BR;
  const int i = i0 + tid;
  const int c = slot / g.nsub;
  const int k_lo = max(c * g.L + 1, i0 + 1);
  const int k_hi = min((c + 1) * g.L, g.kmax);
  const int il = min(i, g.n - 1);
  D best = Lim<D>::lowest();
  int arg = -1;
  const float thr0 = LB[min(il, max(g.nrows - 1, 0))];
  float thr = thr0;
  unsigned int nexact = 0, nviol = 0;
  float2 S = make_float2(0.f, 0.f);  // running sums through cell kb - 1 at the top of a stage
  auto checkpoint = [&](const int kbnd) -> float2 {  // kbnd: multiple of 128, > i for every row of the block
    const long long o = (long long)(kbnd / SCR_BLOCK) * g.n + il;
    return make_float2(SA[o], SB[o]);
  };
  auto step = [&](const float4 r) { S = __fadd2_rn(S, make_float2(r.x, r.y)); };
  auto exact_here = [&](const float4 r) -> D {
    return add_rn<D>(screen_exact_score<D, DIST, RP, false>(S.x, S.y, nullptr), r.z);
  };
  // one cell on its own (diagonal, ragged, forced and audit paths): walk, bound, vote, evaluate
  auto cell = [&](const int k, const float4 r, const bool valid) {
    if (valid) step(r);
    const float ub = ScreenUB<DIST, RP>::eval(S.x, S.y, screen_pm_of<D>(r));
    const bool pass = valid && (forced || !(ub < thr));
    if (__any_sync(0xffffffffu, pass)) {
      const D v = exact_here(r);
      if (AUDIT && valid && !forced && (double)ub < (double)v) ++nviol;
      ++nexact;
      if (valid && v > best) {
        best = v;
        arg = k;
        thr = fmaxf(thr, best);
      }
    } else if (AUDIT && valid) {
      if (!(exact_here(r) < thr)) ++nviol;
    }
  };
  auto stage_max = [&](const float4 r, const int buf) {
    float w = ((tid & 7) == 7) ? r.w : -3.402823466e+38f;
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 8));
    w = fmaxf(w, __shfl_xor_sync(0xffffffffu, w, 16));
    if ((tid & 31) == 31) s_pmax[buf][tid >> 5] = w;
  };

  const int nk = k_hi - k_lo + 1;
  const int nstages = (nk + BK - 1) / BK;
  if (AUDIT && counters && tid == 0) atomicAdd(&counters[stage_mask ? 2 : 3], 1ULL);
  unsigned int todo = AUDIT ? (nstages >= 32 ? 0xffffffffu : ((1u << nstages) - 1u)) : stage_mask;
  int s = todo ? (__ffs(todo) - 1) : -1;
  if (s >= 0) {
    const int kb = k_lo + s * BK;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kb + tid <= k_hi) r = recf[kb + tid];
    sbuf[0][tid] = r;
    stage_max(r, 0);
  }
  int buf = 0;
  while (s >= 0) {
    __syncthreads();
    todo &= todo - 1u;
    const int sn = todo ? (__ffs(todo) - 1) : -1;
    const int kb = k_lo + s * BK;
    float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sn >= 0 && k_lo + sn * BK + tid <= k_hi) nx = recf[k_lo + sn * BK + tid];
    const int cnt = min(BK, k_hi - kb + 1);
    const float4* __restrict__ sb = sbuf[buf];
    const bool diag_stage = kb <= i0 + BR - 1;
    S = diag_stage ? make_float2(0.f, 0.f) : checkpoint(kb - 1);
    const bool sel_dropped = AUDIT && !((stage_mask >> s) & 1u);
    if (sel_dropped) {
      // (audit) a stage the select kernel dropped: walk it, every cell must lose against the row's initial threshold
      if (counters && (tid & 31) == 0) atomicAdd(&counters[5], 1ULL);
      for (int kk = 0; kk < cnt; ++kk) {
        if (kb + kk > i) {
          step(sb[kk]);
          if (i < g.nrows && !(exact_here(sb[kk]) < thr0)) ++nviol;
        }
      }
    } else if (diag_stage || cnt < BK || forced) {
      // diagonal stage (row i only sees k > i), ragged last stage, rows that take no bounds: cell by cell
      for (int kk = 0; kk < cnt; ++kk) cell(kb + kk, sb[kk], kb + kk > i);
    } else {
      const float2 E = checkpoint(kb + BK - 1);  // the row's sums at the stage's last cell
      bool skip = false;
      if (mono) {
        float pmx = s_pmax[buf][0];
#pragma unroll
        for (int w = 1; w < BR / 32; ++w) pmx = fmaxf(pmx, s_pmax[buf][w]);
        const float ubs = run32_range_ub<DIST, RP>(S.x, S.y, E.x, E.y, pmx, (float)(BK + 8), (float)(kb + BK - 1 - i), mono == 2);
        skip = !__any_sync(0xffffffffu, !(ubs < thr));
      }
      if (AUDIT && counters && (tid & 31) == 0) atomicAdd(&counters[skip ? 5 : 4], 1ULL);
      if (skip) {
        if (AUDIT) {
          for (int kk = 0; kk < BK; ++kk) {
            step(sb[kk]);
            if (!(exact_here(sb[kk]) < thr)) ++nviol;
          }
        }
      } else {
#pragma unroll 1
        for (int kk = 0; kk < BK; kk += U) {
          float2 W[U];  // the row's running sums at the U cells of the group
#pragma unroll
          for (int u = 0; u < U; ++u) {
            step(sb[kk + u]);
            W[u] = S;
          }
          if (mono) {
            const float ubg = run32_range_ub<DIST, RP>(W[0].x, W[0].y, W[U - 1].x, W[U - 1].y, sb[kk + U - 1].w, (float)U,
                                                       (float)(kb + kk + U - 1 - i), mono == 2);
            const bool gdrop = !__any_sync(0xffffffffu, !(ubg < thr));
            if (AUDIT && counters && (tid & 31) == 0) atomicAdd(&counters[gdrop ? 7 : 6], 1ULL);
            if (gdrop) {
              if (AUDIT) {
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                  const D v = add_rn<D>(screen_exact_score<D, DIST, RP, false>(W[u].x, W[u].y, nullptr), sb[kk + u].z);
                  if (!(v < thr)) ++nviol;
                }
              }
              continue;
            }
          }
          bool alive = false;
          float ub[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            ub[u] = ScreenUB<DIST, RP>::eval(W[u].x, W[u].y, screen_pm_of<D>(sb[kk + u]));
            alive |= !(ub[u] < thr);
          }
          const bool ev = __any_sync(0xffffffffu, alive);
          if (AUDIT || ev) {
            D v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
              v[u] = add_rn<D>(screen_exact_score<D, DIST, RP, false>(W[u].x, W[u].y, nullptr), sb[kk + u].z);
            if (AUDIT) {
#pragma unroll
              for (int u = 0; u < U; ++u) {
                if ((double)ub[u] < (double)v[u]) ++nviol;  // a cell bound that was too low
                if (!ev && !(v[u] < thr)) ++nviol;           // a dropped cell that does not lose
              }
            }
            if (ev) {
#pragma unroll
              for (int u = 0; u < U; ++u)
                if (v[u] > best) {
                  best = v[u];
                  arg = kb + kk + u;
                }
              thr = fmaxf(thr, best);
              nexact += U;
            }
          }
        }
      }
    }
    if (sn >= 0) {
      sbuf[buf ^ 1][tid] = nx;
      stage_max(nx, buf ^ 1);
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

template <int DIST, bool RP, bool AUDIT, int BR>
__global__ void __launch_bounds__(BR, 8)
level_live_screen_run32_kernel(Geom g, const Rec4<float>* __restrict__ rec, const float* __restrict__ SA,
                               const float* __restrict__ SB, long long cp_stride, const int* __restrict__ ipos,
                               const float* __restrict__ LB, float* __restrict__ part_val, int* __restrict__ part_arg,
                               unsigned long long* __restrict__ counters, const int* __restrict__ dev_flags,
                               const int4* __restrict__ live, long long live_cap, LiveHdr* __restrict__ hdr) {
  static_assert(BR == SCR_BLOCK, "a stage is one 128-block of k");
  __shared__ float4 sbuf[2][BR];
  __shared__ float s_pmax[2][BR / 32];
  __shared__ int s_next;
  // range-tier mode: 0 none, 1 on, 2 on and the priorities ascend (Poisson risk-part: end-cell bound instead of the box, which
  // holds in any element order); the other scores need the priorities ascending
  const bool sorted = dev_flags && (dev_flags[1] == 0);
  const int mono = (DIST == DIST_POISSON && RP) ? (sorted ? 2 : 1) : (sorted ? 1 : 0);
  const long long pstride = (long long)g.nch_alloc * g.nsub * g.n;
  const int tid = threadIdx.x;
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
    __syncthreads();
    if (tid == 0) s_next = atomicAdd(&hdr->cursor, 1);
    __syncthreads();
    const int t = s_next;
    if (t >= nlive) break;
    int cls = 0;
#pragma unroll
    for (int q = 0; q < SCR_CLASSES - 1; ++q) cls += (t >= cum[q]) ? 1 : 0;
    const int4 e = live[(long long)cls * live_cap + (t - (cls ? cum[cls - 1] : 0))];
    const int inst = e.x;
    const bool forced = (e.y * BR) < ipos[inst];
    screen_run32_live_tile<DIST, RP, AUDIT, BR>(g, e.y, e.z, (unsigned int)e.w, rec + (long long)inst * g.rec_stride,
                                                SA + inst * cp_stride, SB + inst * cp_stride, forced,
                                                LB + (long long)inst * g.n, part_val + inst * pstride,
                                                part_arg + inst * pstride, counters, forced ? 0 : mono, sbuf, s_pmax);
  }
}

}  // namespace ps
