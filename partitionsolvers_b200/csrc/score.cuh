// score.cuh -- range-score functors evaluated in registers (the n x n score table of the reference,
// score.hpp:47 `partialSums_`, is never materialised).
//
// Each functor reproduces the ROUNDING RECIPE of the reference formula it replaces
// (score_impl.hpp:374-390 Poisson, :426-441 Gaussian, :461-474 RationalScore):
//   * DataType=float, Poisson: everything in float (std::log(float) is logf).
//   * DataType=float, Gaussian mult-clust and RationalScore: std::pow(float,int) returns double, so the
//     expression is evaluated in double from the float sums and rounded to float once.
//   * DataType=float, Gaussian risk-part: A*A in float, the two divisions in double, one rounding.
//   * DataType=double: everything in double.
//   * no fused multiply-add (reference build: -O3 -mavx2, no -mfma): every product/sum below is an
//     explicit round-to-nearest intrinsic so nvcc cannot contract it; divisions are IEEE (div.rn).
// log/logf are restatements of glibc's routines (glibc_log.cuh), so the Poisson scores carry the same bits
// as the reference's std::log calls.
#pragma once
#include <cuda_runtime.h>

#include "glibc_log.cuh"
#include "math_fast.cuh"

namespace ps {

enum { DIST_GAUSSIAN = 0, DIST_POISSON = 1, DIST_RATIONAL = 2 };

// FAST = operands proven to be in the safe range of math_fast.cuh (per-instance precheck): the same
// arithmetic, bit for bit, without the range checks and special-value handling.
template <typename D, int DIST, bool RP, bool FAST = false>
struct ScoreFn;

template <bool FAST> __device__ __forceinline__ float div_f(float a, float b) { return FAST ? fast_fdiv(a, b) : __fdiv_rn(a, b); }
template <bool FAST> __device__ __forceinline__ double div_d(double a, double b) { return FAST ? fast_ddiv(a, b) : __ddiv_rn(a, b); }
// log = glibc's algorithm (glibc_log.cuh): bit-identical to the std::log the reference calls
template <bool FAST> __device__ __forceinline__ float log_f(float x) { return FAST ? glibc_logf_fast(x) : glibc_logf(x); }
template <bool FAST> __device__ __forceinline__ double log_d(double x, const GlogEntry* tab = g_glog_tab) {
  return FAST ? glibc_log_fast(x, tab) : glibc_log(x);
}

// A > B.  On the fast path both are positive normal numbers, for which the order of the bit patterns is the
// numeric order: the double compare then runs on the integer pipe instead of the FP64 pipe.
template <bool FAST> __device__ __forceinline__ bool gt_ab(float a, float b) { return a > b; }
template <bool FAST> __device__ __forceinline__ bool gt_ab(double a, double b) {
  return FAST ? (__double_as_longlong(a) > __double_as_longlong(b)) : (a > b);
}

// ------------------------------------------------------------------ float
template <bool FAST>
struct ScoreFn<float, DIST_POISSON, false, FAST> {
  // the expression of the A > B branch alone (callers that have already tested A > B)
  static __device__ __forceinline__ float eval_gt(float A, float B) {
    float q = div_f<FAST>(A, B);
    return __fsub_rn(__fadd_rn(__fmul_rn(A, log_f<FAST>(q)), B), A);
  }
  static __device__ __forceinline__ float eval(float A, float B) { return (A > B) ? eval_gt(A, B) : 0.f; }
};
template <bool FAST>
struct ScoreFn<float, DIST_POISSON, true, FAST> {
  static __device__ __forceinline__ float eval(float A, float B) {
    return __fmul_rn(A, log_f<FAST>(div_f<FAST>(A, B)));
  }
};
template <bool FAST>
struct ScoreFn<float, DIST_GAUSSIAN, false, FAST> {
  static __device__ __forceinline__ float eval_gt(float A, float B) {
    double a = (double)A, b = (double)B;
    double s = __dsub_rn(__dmul_rn(.5, __dadd_rn(div_d<FAST>(__dmul_rn(a, a), b), b)), a);
    return __double2float_rn(s);
  }
  static __device__ __forceinline__ float eval(float A, float B) { return (A > B) ? eval_gt(A, B) : 0.f; }
};
template <bool FAST>
struct ScoreFn<float, DIST_GAUSSIAN, true, FAST> {
  static __device__ __forceinline__ float eval(float A, float B) {
    float sq = __fmul_rn(A, A);
    return __double2float_rn(div_d<FAST>(__ddiv_rn((double)sq, 2.), (double)B));
  }
};
template <bool RP, bool FAST>
struct ScoreFn<float, DIST_RATIONAL, RP, FAST> {
  static __device__ __forceinline__ float eval(float A, float B) {
    double a = (double)A;
    return __double2float_rn(div_d<FAST>(__dmul_rn(a, a), (double)B));
  }
};

// ------------------------------------------------------------------ double
template <bool FAST>
struct ScoreFn<double, DIST_POISSON, false, FAST> {
  static __device__ __forceinline__ double eval_gt(double A, double B, const GlogEntry* tab = g_glog_tab) {
    double q = div_d<FAST>(A, B);
    return __dsub_rn(__dadd_rn(__dmul_rn(A, log_d<FAST>(q, tab)), B), A);
  }
  static __device__ __forceinline__ double eval(double A, double B, const GlogEntry* tab = g_glog_tab) {
    return gt_ab<FAST>(A, B) ? eval_gt(A, B, tab) : 0.;
  }
};
template <bool FAST>
struct ScoreFn<double, DIST_POISSON, true, FAST> {
  static __device__ __forceinline__ double eval(double A, double B, const GlogEntry* tab = g_glog_tab) {
    return __dmul_rn(A, log_d<FAST>(div_d<FAST>(A, B), tab));
  }
};
template <bool FAST>
struct ScoreFn<double, DIST_GAUSSIAN, false, FAST> {
  static __device__ __forceinline__ double eval_gt(double A, double B) {
    return __dsub_rn(__dmul_rn(.5, __dadd_rn(div_d<FAST>(__dmul_rn(A, A), B), B)), A);
  }
  static __device__ __forceinline__ double eval(double A, double B) { return (A > B) ? eval_gt(A, B) : 0.; }
};
template <bool FAST>
struct ScoreFn<double, DIST_GAUSSIAN, true, FAST> {
  static __device__ __forceinline__ double eval(double A, double B) {
    return div_d<FAST>(__ddiv_rn(__dmul_rn(A, A), 2.), B);
  }
};
template <bool RP, bool FAST>
struct ScoreFn<double, DIST_RATIONAL, RP, FAST> {
  static __device__ __forceinline__ double eval(double A, double B) {
    return div_d<FAST>(__dmul_rn(A, A), B);
  }
};

// Runtime-dispatched evaluation for the cold paths (level 1, backtrace, compute_score).
template <typename D>
__device__ __forceinline__ D score_dyn(int dist, int rp, D A, D B) {
  switch (dist * 2 + (rp ? 1 : 0)) {
    case 0: return ScoreFn<D, DIST_GAUSSIAN, false>::eval(A, B);
    case 1: return ScoreFn<D, DIST_GAUSSIAN, true>::eval(A, B);
    case 2: return ScoreFn<D, DIST_POISSON, false>::eval(A, B);
    case 3: return ScoreFn<D, DIST_POISSON, true>::eval(A, B);
    default: return ScoreFn<D, DIST_RATIONAL, false>::eval(A, B);
  }
}

template <typename D>
struct Lim;
template <>
struct Lim<float> {
  static __device__ __host__ __forceinline__ float lowest() { return -3.402823466e+38f; }
};
template <>
struct Lim<double> {
  static __device__ __host__ __forceinline__ double lowest() { return -1.7976931348623157e+308; }
};

template <typename D>
__device__ __forceinline__ D add_rn(D x, D y);
template <>
__device__ __forceinline__ float add_rn<float>(float x, float y) { return __fadd_rn(x, y); }
template <>
__device__ __forceinline__ double add_rn<double>(double x, double y) { return __dadd_rn(x, y); }
template <typename D>
__device__ __forceinline__ D sub_rn(D x, D y);
template <>
__device__ __forceinline__ float sub_rn<float>(float x, float y) { return __fsub_rn(x, y); }
template <>
__device__ __forceinline__ double sub_rn<double>(double x, double y) { return __dsub_rn(x, y); }

}  // namespace ps
