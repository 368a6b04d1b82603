// screen_select.cuh -- the RANGE tiers of the screened levels, run before anything heavy is launched, for all three families:
// exact-sum screen (screen.cuh; SEL_PREFIX), double data whose sums round (screen_run.cuh; SEL_FIXED: bounds from 64-bit
// fixed-point prefix sums), float data whose sums round (screen_run.cuh; SEL_RUN32: bounds from the row's own running sums at
// the 128-block boundaries).  One CTA per row block; the live tiles go, with the mask of their live stages, to compacted
// lists ordered by expected cost; see the header of the section "Screened level = four launches" in screen.cuh.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "screen.cuh"
#include "screen_run.cuh"

namespace ps {

enum { SEL_PREFIX = 0, SEL_FIXED = 1, SEL_RUN32 = 2 };
struct SelSrc {
  const longlong2* pfx;   // SEL_FIXED: fixed-point prefix sums [R][n+1]
  const FxInfo* fx;       // SEL_FIXED: per-instance shifts and ipos
  const float* cpA;       // SEL_RUN32: running sums of a at the 128-block boundaries, [R][n/128+1][n]
  const float* cpB;
  long long cp_stride;    // SEL_RUN32: elements per instance
  const int* ipos;        // SEL_RUN32: per instance, rows >= ipos have only positive a to their right
};

// SCR_SEL_TPR threads share a row (each takes every SCR_SEL_TPR-th batch of four tiles): one thread per row leaves ~5 warps
// per scheduler on a 782-block level and the bound chains (LDS, DADD, F2F, MUFU.RCP, MUFU.LG2) run at their latency.
// Instances with few chunks (replica batches) take one thread per row.
template <typename D, int DIST, bool RP, bool AUDIT, int SCR_SEL_TPR, int SRC>
__global__ void __launch_bounds__(SCR_BLOCK * SCR_SEL_TPR)
screen_select_kernel(Geom g, const Rec4<D>* __restrict__ rec, const D* __restrict__ LB, const float* __restrict__ bmax,
                     int nblk, const int* __restrict__ dev_flags, int4* __restrict__ live, long long live_cap,
                     unsigned int* __restrict__ tile_mask, int rb_stride, LiveHdr* __restrict__ hdr,
                     unsigned long long* __restrict__ counters, const int* __restrict__ kbest,
                     int* __restrict__ rb_slots, int* __restrict__ rb_cnt, SelSrc src) {
  static_assert(SRC == SEL_PREFIX || (SRC == SEL_FIXED && sizeof(D) == 8) || (SRC == SEL_RUN32 && sizeof(D) == 4), "");
  constexpr int BR = SCR_BLOCK;
  constexpr int NT = BR * SCR_SEL_TPR;  // threads per CTA
  constexpr bool DBL = (sizeof(D) == 8);
  constexpr bool FRONTIER = !RP && DIST != DIST_RATIONAL;
  constexpr bool GROUPB = !(DIST == DIST_POISSON && RP);
  constexpr bool BOX = !GROUPB;
  // per chunk of this row block: live-stage mask, the record at the tile's last cell (and, box bound, at the first cell of a
  // tile right of the diagonal), the largest pm of the tile -- staged once so that the per-row loop below reads shared memory
  extern __shared__ __align__(16) unsigned char s_raw[];
  using Raw = typename std::conditional<SRC == SEL_FIXED, long long, D>::type;  // what a staged sum is before conversion
  Raw* s_ex = reinterpret_cast<Raw*>(s_raw);              // [NCH] x at k_hi
  Raw* s_ey = s_ex + g.NCH;                               // [NCH] y at k_hi
  Raw* s_fx = s_ey + g.NCH;                               // [NCH] x at k_lo (BOX)
  Raw* s_fy = s_fx + g.NCH;                               // [NCH]
  float* s_tmax = reinterpret_cast<float*>(s_fy + g.NCH);  // [NCH]
  unsigned int* s_mask = reinterpret_cast<unsigned int*>(s_tmax + g.NCH);  // [NCH]
  unsigned int* s_cost = s_mask + g.NCH;  // [NCH] live (row, stage) pairs of the tile
  unsigned char* s_emit = reinterpret_cast<unsigned char*>(s_cost + g.NCH);  // [NCH * nsub] slot has a list entry
  __shared__ int s_cmin, s_cmax;          // chunks that hold the best samples of this block's rows: the "ridge" 
  const int inst = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  // highest row blocks first: their segments lead the list, so the persistent CTAs start on the expensive tiles (rows past
  // the A = B frontier) and the cheap low rows fill the tail
  const int rb = g.rb_first + g.rb_step * (g.NRB - 1 - (int)blockIdx.x);
  const int part = tid / BR;  // warp-uniform: which share of the tiles this thread takes for its row
  const int i0 = rb * BR, i = i0 + (tid & (BR - 1));
  rec += (long long)inst * g.rec_stride;
  LB += (long long)inst * g.n;
  bmax += (long long)inst * nblk;
  const int c_diag = rb / g.G;
  if (tid == 0) {
    s_cmin = 0x7fffffff;
    s_cmax = -1;
  }
  for (int c = c_diag + tid; c < g.NCH; c += NT) {
    const int k_lo = max(c * g.L + 1, i0 + 1);
    const int k_hi = min((c + 1) * g.L, g.kmax);
    s_mask[c] = 0u;
    s_cost[c] = 0u;
    float tmax = -3.402823466e+38f;
    for (int b = (k_lo - 1) / BR; b <= (k_hi - 1) / BR; ++b) tmax = fmaxf(tmax, bmax[b]);
    s_tmax[c] = tmax;
    if constexpr (SRC == SEL_PREFIX) {
      const Rec4<D> E = rec[max(k_hi, 0)];
      s_ex[c] = E.x;
      s_ey[c] = E.y;
      if (BOX) {
        const Rec4<D> F = rec[min(k_lo, g.n)];
        s_fx[c] = F.x;
        s_fy[c] = F.y;
      }
    } else if constexpr (SRC == SEL_FIXED) {
      const longlong2 E = src.pfx[(long long)inst * g.rec_stride + max(k_hi, 0)];
      s_ex[c] = E.x;
      s_ey[c] = E.y;
      if (BOX) {
        const longlong2 F = src.pfx[(long long)inst * g.rec_stride + min(k_lo, g.n)];
        s_fx[c] = F.x;
        s_fy[c] = F.y;
      }
    }
  }
  __syncthreads();
  const int il = min(i, g.n - 1);
  // what a range sum of this thread's row is formed from: SEL_PREFIX level records (exact sums), SEL_FIXED 64-bit fixed-point
  // prefix sums (screen_run.cuh, double data), SEL_RUN32 the row's own float running sums at the 128-block boundaries
  D Xi = 0, Yi = 0;
  float bx = 0.f, by = 0.f;
  longlong2 Pi = make_longlong2(0, 0);
  float sca = 1.f, scb = 1.f;
  bool no_bounds = false;  // rows that own non-positive sums take no bounds (screen_run.cuh F4): every stage is live
  if constexpr (SRC == SEL_PREFIX) {
    Xi = rec[il].x, Yi = rec[il].y;
    bx = -(float)Xi, by = -(float)Yi;  // float data: Px[k] - Px[i] is exact
  } else if constexpr (SRC == SEL_FIXED) {
    Pi = src.pfx[(long long)inst * g.rec_stride + il];
    const FxInfo F = src.fx[inst];
    sca = F.sca, scb = F.scb;
    no_bounds = i0 < F.ipos;
  } else {
    no_bounds = i0 < src.ipos[inst];
  }
  const float* cpA = (SRC == SEL_RUN32) ? src.cpA + inst * src.cp_stride : nullptr;
  const float* cpB = (SRC == SEL_RUN32) ? src.cpB + inst * src.cp_stride : nullptr;
  const float thr = screen_rd<D>(LB[min(il, max(g.nrows - 1, 0))]);
  const bool mono = !no_bounds && (BOX || ((FRONTIER || GROUPB) && dev_flags && (dev_flags[1] == 0)));
  const int rmode = (BOX && dev_flags && (dev_flags[1] == 0)) ? 2 : 1;  // Poisson risk-part: hull bound on ascending priorities
  const bool row_ok = i < g.nrows;
  if (kbest) {
    // the chunk of the row's best sample; the row maximum and the cells that survive the bounds lie around it
    const int cb = row_ok ? (max(kbest[(long long)inst * g.n + i], 1) - 1) / g.L : -1;
    const int lo = __reduce_min_sync(0xffffffffu, row_ok ? cb : 0x7fffffff);
    const int hi = __reduce_max_sync(0xffffffffu, cb);
    if (lane == 0) {
      atomicMin(&s_cmin, lo);
      atomicMax(&s_cmax, hi);
    }
  }
  auto fsum2 = [&](const Raw x, const Raw y, float& A, float& B) {
    if constexpr (SRC == SEL_FIXED) {
      A = __ll2float_rn((long long)x - Pi.x) * sca;
      B = __ll2float_rn((long long)y - Pi.y) * scb;
    } else {
      A = DBL ? __double2float_rn((double)x - (double)Xi) : __fadd_rn((float)x, bx);
      B = DBL ? __double2float_rn((double)y - (double)Yi) : __fadd_rn((float)y, by);
    }
  };
  auto fsums = [&](const int k, float& A, float& B) {
    if constexpr (SRC == SEL_FIXED) {
      const longlong2 P = src.pfx[(long long)inst * g.rec_stride + k];
      fsum2(P.x, P.y, A, B);
    } else if constexpr (SRC == SEL_RUN32) {  // k is a multiple of 128 beyond the row: the row's running sums through cell k
      const long long o = (long long)(k / SCR_BLOCK) * g.n + il;
      A = cpA[o];
      B = cpB[o];
    } else {
      const Rec4<D> R = rec[k];
      fsum2(R.x, R.y, A, B);
    }
  };
  // bound for the cells (k_first .. k_last] of this row; SEL_RUN32 forms it from the running sums at k_last (a 128-block
  // boundary right of the whole row block) and carries the roundings of the additions in between (run32_range_ub)
  auto range_ub = [&](const float Af, const float Bf, const float Al, const float Bl, const float pm, const int span,
                      const int k_last) -> float {
    if constexpr (SRC == SEL_RUN32)
      return run32_range_ub<DIST, RP>(Af, Bf, Al, Bl, pm, (float)(span + 8), (float)(k_last - i), rmode == 2);
    else
      return screen_range_ub<DIST, RP>(Af, Bf, Al, Bl, pm, rmode);
  };
  auto bound_at = [&](const int k_last) -> bool {  // can a bound be formed at k_last for every row of the block?
    return SRC != SEL_RUN32 || ((k_last % SCR_BLOCK) == 0 && k_last >= i0 + BR);
  };
  // stage tier of one tile for this thread's row: the tile bound per 128-block -- S at the stage's last cell, the stage's own
  // largest pm -- then the warp's mask / live-pair count go to shared memory
  auto stage_tier = [&](const int c, const bool lose) {
    const int k_lo = max(c * g.L + 1, i0 + 1);
    const int k_hi = min((c + 1) * g.L, g.kmax);
    const int nst = (k_hi - k_lo + BR) / BR;
    unsigned int m = 0u;
#pragma unroll 4
    for (int s = 0; s < nst; ++s) {
      const int kb = k_lo + s * BR, ke = min(kb + BR - 1, k_hi);
      const int k0 = max(kb, i + 1);
      bool lose_s = lose || (k0 > ke);
      if (!lose_s && bound_at(ke) && (SRC != SEL_RUN32 || kb > i0 + BR - 1)) {
        float Ae, Be;
        fsums(ke, Ae, Be);
        float A0 = Ae, B0 = Be;
        if (BOX) fsums(SRC == SEL_RUN32 ? kb - 1 : min(k0, ke), A0, B0);  // running sums: the boundary before the stage
        lose_s = range_ub(A0, B0, Ae, Be, bmax[(kb - 1) / BR], BR, ke) < thr;
      }
      if (!lose_s) m |= 1u << s;
    }
    const unsigned int pairs = __reduce_add_sync(0xffffffffu, (unsigned int)__popc(m));
    m = __reduce_or_sync(0xffffffffu, m);
    if (lane == 0 && m) {
      atomicOr(&s_mask[c], m);
      atomicAdd(&s_cost[c], pairs);
    }
  };
  if (!mono) {
    // no range tiers (caller-supplied order that is not sorted by priority; rows that take no bounds): every stage is live
    for (int c = c_diag + tid; c < g.NCH; c += NT) {
      const int k_lo = max(c * g.L + 1, i0 + 1);
      const int k_hi = min((c + 1) * g.L, g.kmax);
      const int nst = (k_hi - k_lo + BR) / BR;
      s_mask[c] = (k_hi >= k_lo) ? (nst >= 32 ? 0xffffffffu : ((1u << nst) - 1u)) : 0u;
      s_cost[c] = (unsigned int)BR * (unsigned int)max(nst, 0);
    }
  } else {
    // tile tier: the bound at the tile's last cell with the largest pm of the tile's 128-blocks covers the whole tile.  Four
    // tiles per step: their bound chains are independent, and one vote dismisses all four (most tiles lose for every row)
    for (int c4 = c_diag + 4 * part; c4 < g.NCH; c4 += 4 * SCR_SEL_TPR) {
      if (c4 + 3 < g.NCH && c4 > c_diag) {
        // all four at once: S at the last cell of the fourth tile with the largest pm of the four covers them all; far from
        // the row maxima this loses for every row and the four separate bounds are never formed
        const int ce = c4 + 3;
        const int k_end = min((ce + 1) * g.L, g.kmax);
        if (bound_at(k_end)) {
          float At, Bt;
          if constexpr (SRC == SEL_RUN32)
            fsums(k_end, At, Bt);
          else
            fsum2(s_ex[ce], s_ey[ce], At, Bt);
          float Af = At, Bf = Bt;
          if constexpr (BOX && SRC == SEL_RUN32)
            fsums(c4 * g.L, Af, Bf);  // the boundary before the first of the four tiles (c4 > c_diag: right of the block)
          else if (BOX)
            fsum2(s_fx[c4], s_fy[c4], Af, Bf);
          const float pm4 = fmaxf(fmaxf(s_tmax[c4], s_tmax[c4 + 1]), fmaxf(s_tmax[c4 + 2], s_tmax[ce]));
          const bool lose4 = !row_ok || (range_ub(Af, Bf, At, Bt, pm4, 4 * g.L, k_end) < thr);
          if (!__any_sync(0xffffffffu, !lose4)) continue;
        }
      }
      bool lose[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = min(c4 + u, g.NCH - 1);
        const int k_lo = max(c * g.L + 1, i0 + 1);
        const int k_hi = min((c + 1) * g.L, g.kmax);
        const int kf = max(k_lo, i + 1);  // the row's first cell in this tile
        lose[u] = !row_ok || (c4 + u >= g.NCH) || (kf > k_hi);
        if (!lose[u] && bound_at(k_hi) && !(BOX && SRC == SEL_RUN32 && c == c_diag)) {
          float At, Bt;
          if constexpr (SRC == SEL_RUN32)
            fsums(k_hi, At, Bt);
          else
            fsum2(s_ex[c], s_ey[c], At, Bt);
          float Af = At, Bf = Bt;
          if constexpr (BOX && SRC == SEL_RUN32) {
            fsums(c * g.L, Af, Bf);  // c > c_diag: the tile's left boundary lies right of the whole row block
          } else if (BOX) {
            if (c == c_diag)
              fsums(min(kf, k_hi), Af, Bf);
            else
              fsum2(s_fx[c], s_fy[c], Af, Bf);
          }
          lose[u] = range_ub(Af, Bf, At, Bt, s_tmax[c], g.L, k_hi) < thr;
        }
      }
      if (!__any_sync(0xffffffffu, !(lose[0] && lose[1] && lose[2] && lose[3]))) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (__any_sync(0xffffffffu, !lose[u])) stage_tier(c4 + u, lose[u]);
    }
  }
  __syncthreads();
  // Every partial slot's mask goes to tile_mask (the fold reads it); live tiles are appended to the list of their cost
  // class.  A tile has g.nsub partial slots: the expensive classes are split into up to nsub list entries (consecutive runs of
  // stages, one slot each) so that no single work item dominates the level; the others use their first slot.
  const int nsub = g.nsub;
  tile_mask += ((long long)inst * rb_stride + rb) * g.nch_alloc * nsub;
  for (int cb = c_diag + (tid & ~31); cb < g.NCH; cb += NT) {  // warp-uniform trip count: all 32 lanes vote
    const int c = cb + lane;
    const bool in = c < g.NCH;
    const int k_lo = max(c * g.L + 1, i0 + 1);
    const int k_hi = min((c + 1) * g.L, g.kmax);
    const int nst = (k_hi - k_lo + BR) / BR;
    const unsigned int mk = in ? s_mask[c] : 0u;
    const bool take = in && (k_hi >= k_lo) && (AUDIT || mk != 0u);
    int cls = SCR_CLASSES - 1;
    if (take) {
      // cost class: 0 = the tile holds the best sample of some row of the block (most of its cells will be evaluated with
      // the reference's arithmetic), 1 = next to such a tile, 2 / 3 = the rest by the share of live (row, stage) pairs
      const unsigned int full = (unsigned int)BR * (unsigned int)nst;
      const unsigned int cost = s_cost[c];
      const int cmin = s_cmin, cmax = s_cmax;
      if (kbest && c >= cmin && c <= cmax)
        cls = 0;
      else if (kbest && c >= cmin - 1 && c <= cmax + 1)
        cls = 1;
      else
        cls = (2u * cost >= full) ? 2 : 3;
    }
    const bool split = take && !AUDIT && nsub > 1 && cls <= 1;
    const int per = max(1, (nst + nsub - 1) / nsub);  // stages per part of a split tile
    for (int h = 0; h < nsub; ++h) {
      unsigned int sub = (h == 0) ? mk : 0u;
      if (split) sub = (h * per < 32) ? (mk & (((per >= 32) ? 0xffffffffu : ((1u << per) - 1u)) << (h * per))) : 0u;
      if (in) tile_mask[c * nsub + h] = sub;
      const bool emit = take && (split ? (sub != 0u) : (h == 0));
      if (in) s_emit[c * nsub + h] = emit ? 1 : 0;
      // warp-aggregated append: one atomic per class present in the warp
      const unsigned int peers = __match_any_sync(0xffffffffu, emit ? cls : -1);
      if (emit) {
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&hdr->nlive[cls], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        live[(long long)cls * live_cap + base + __popc(peers & ((1u << lane) - 1u))] =
            make_int4(inst, rb, c * nsub + h, (int)sub);
      }
    }
  }
  // the row block's slots that hold a partial, ascending (= ascending k): the fold reads exactly these
  __syncthreads();
  if (tid < 32) {
    rb_slots += ((long long)inst * rb_stride + rb) * g.nch_alloc * nsub;
    int pos = 0;
    for (int sb = c_diag * nsub; sb < g.NCH * nsub; sb += 32) {
      const int sl = sb + lane;
      const bool on = (sl < g.NCH * nsub) && s_emit[sl];
      const unsigned int bal = __ballot_sync(0xffffffffu, on);
      if (on) rb_slots[pos + __popc(bal & ((1u << lane) - 1u))] = sl;
      pos += __popc(bal);
    }
    if (lane == 0) rb_cnt[(long long)inst * rb_stride + rb] = pos;
  }
}


}  // namespace ps
