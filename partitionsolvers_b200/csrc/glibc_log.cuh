// glibc_log.cuh -- device restatement of glibc's log() and logf(), operation for operation.
//
// The reference's Poisson scores call std::log (score_impl.hpp:374-390), which on the reference's platform is
// glibc's table-driven log/logf (2.28 and later; sysdeps/ieee754/dbl-64/e_log.c, sysdeps/ieee754/flt-32/e_logf.c,
// the x86-64 multiarch variant built with FMA that every AVX2-class CPU selects).  IEEE add/mul/fma are
// deterministic, so executing the SAME sequence of operations on the SAME coefficient tables gives the SAME
// bits on the GPU as on the host: with these routines the Poisson path is bit-identical to the reference
// instead of "within 1 ulp of log".  The operation order below (in particular which products are fused) was
// read off the disassembly of the library's *_fma variants and checked on the host against libm itself on
// 2*10^8 arguments before being ported (tools/check_glibc_log.c keeps that check runnable).
//
//   log  : N=128 table {1/c, log c}; r = fma(z, 1/c, -1); degree-5 polynomial in r; explicit hi/lo sums.
//          |x-1| < ~2^-4 takes a separate degree-11 polynomial with a split of r for the r^2/2 term.
//   logf : N=16 table, evaluated in double: y = fma(r^2, fma(r^2, A0, fma(r, A1, A2)), fma(k, ln2, log c) + r).
#pragma once
#include <cuda_runtime.h>

#include "glibc_log_data.inc"

namespace ps {

struct __align__(16) GlogEntry {
  double invc, logc;
};
__device__ const GlogEntry g_glog_tab[128] = PS_GLOG_TAB_INIT;
__device__ const GlogEntry g_glogf_tab[16] = PS_GLOGF_TAB_INIT;

// Polynomial coefficients live in the constant bank: an FP64 instruction can take a c[bank][offset] operand
// directly, whereas a 64-bit literal costs two uniform-register moves every time it is rematerialised
// (12 of the 78 instructions per cell before this change; FP64 instructions hold the issue port for two
// cycles each, so every other instruction is on the critical path of this kernel).
struct GlogCoef {
  double ln2hi, ln2lo, a[5], b[11], one27, neg_one27;
};
__constant__ GlogCoef c_glog = {
    PS_GLOG_LN2HI, PS_GLOG_LN2LO,
    {PS_GLOG_A0, PS_GLOG_A1, PS_GLOG_A2, PS_GLOG_A3, PS_GLOG_A4},
    {PS_GLOG_B0, PS_GLOG_B1, PS_GLOG_B2, PS_GLOG_B3, PS_GLOG_B4, PS_GLOG_B5, PS_GLOG_B6, PS_GLOG_B7, PS_GLOG_B8,
     PS_GLOG_B9, PS_GLOG_B10},
    0x1p27, -0x1p27};

// x positive, normal, finite (hi/lo are its two words; the subnormal path passes rescaled words).
// tab = the 128-entry table: g_glog_tab in global memory, or a shared-memory copy (2 KB) in the level kernels.
__device__ __forceinline__ double glibc_log_core(double x, int hi, int lo,
                                                 const GlogEntry* __restrict__ tab = g_glog_tab) {
  if ((unsigned)(hi - 0x3FEE0000) < 0x30900u) {
    // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    const double r = __dadd_rn(x, -1.0);
    double t2 = __fma_rn(r, c_glog.b[2], c_glog.b[1]);
    double t3 = __fma_rn(r, c_glog.b[5], c_glog.b[4]);
    const double r2 = __dmul_rn(r, r);
    const double t5 = __fma_rn(r, c_glog.b[8], c_glog.b[7]);
    t2 = __fma_rn(r2, c_glog.b[3], t2);
    t3 = __fma_rn(r2, c_glog.b[6], t3);
    const double r3 = __dmul_rn(r, r2);
    double t1 = __fma_rn(r2, c_glog.b[9], t5);
    t1 = __fma_rn(r3, c_glog.b[10], t1);
    t1 = __fma_rn(t1, r3, t3);
    t1 = __fma_rn(t1, r3, t2);
    const double w = __fma_rn(r, c_glog.one27, r);
    const double rhi = __fma_rn(c_glog.neg_one27, r, w);
    const double rhi2 = __dmul_rn(rhi, rhi);
    const double rlo = __dadd_rn(r, -rhi);
    const double h = __fma_rn(rhi2, c_glog.b[0], r);
    const double d = __dadd_rn(r, -h);
    const double s = __dadd_rn(r, rhi);
    double l = __fma_rn(rhi2, c_glog.b[0], d);
    const double m = __dmul_rn(c_glog.b[0], rlo);
    l = __fma_rn(m, s, l);
    const double y = __fma_rn(t1, r3, l);
    return __dadd_rn(h, y);
  }
  const int thi = hi - 0x3fe60000;
  const double z = __hiloint2double(hi - (thi & 0xfff00000), lo);
  const GlogEntry e = *reinterpret_cast<const GlogEntry*>(reinterpret_cast<const char*>(tab) + ((thi >> 9) & 0x7f0));
  const double kd = __int2double_rn(thi >> 20);  // conversion pipe (idle in this kernel) instead of an FP64 add
  const double w = __fma_rn(kd, c_glog.ln2hi, e.logc);
  const double r = __fma_rn(z, e.invc, -1.0);
  const double p12 = __fma_rn(r, c_glog.a[2], c_glog.a[1]);
  const double h = __dadd_rn(r, w);
  const double r2 = __dmul_rn(r, r);
  double l = __dadd_rn(__dadd_rn(w, -h), r);
  l = __fma_rn(kd, c_glog.ln2lo, l);
  const double r3 = __dmul_rn(r, r2);
  const double p34 = __fma_rn(r, c_glog.a[4], c_glog.a[3]);
  l = __fma_rn(r2, c_glog.a[0], l);
  const double p = __fma_rn(p34, r2, p12);
  const double y = __fma_rn(r3, p, l);
  return __dadd_rn(y, h);
}

// operands proven positive and normal by the safe-range precheck
__device__ __forceinline__ double glibc_log_fast(double x, const GlogEntry* __restrict__ tab = g_glog_tab) {
  return glibc_log_core(x, __double2hiint(x), __double2loint(x), tab);
}

// every double: zero -> -inf, negative / NaN -> NaN, +inf -> +inf, subnormals rescaled by 2^52
__device__ __forceinline__ double glibc_log(double x) {
  int hi = __double2hiint(x), lo = __double2loint(x);
  if ((unsigned)(hi - 0x3FEE0000) >= 0x30900u && (unsigned)(((unsigned)hi >> 16) - 0x0010u) >= 0x7ff0u - 0x0010u) {
    if (((hi & 0x7fffffff) | lo) == 0) return -__longlong_as_double(0x7ff0000000000000LL);
    if (hi == 0x7ff00000 && lo == 0) return x;
    if (hi < 0 || (hi & 0x7ff00000) == 0x7ff00000) return __longlong_as_double(0x7ff8000000000000LL);
    const double xs = __dmul_rn(x, 0x1p52);
    hi = __double2hiint(xs) - (52 << 20);
    lo = __double2loint(xs);
    x = __hiloint2double(hi, lo);
  }
  return glibc_log_core(x, hi, lo);
}

// ---- logf -------------------------------------------------------------------------------------------
// tab = the 16-entry table, in global memory (g_glogf_tab) or a shared-memory copy of it
__device__ __forceinline__ float glibc_logf_core(int ix, const GlogEntry* __restrict__ tab = g_glogf_tab) {
  const int tmp = ix - 0x3f330000;
  // k = tmp >> 23 (arithmetic), written as a signed bit-field extract: with the plain shift feeding an
  // integer->double bit trick, ptxas 12.9 was seen to fold "(ix + c) >> 23, then ^ 0x80000000" into a single
  // add for one of two inlined copies, dropping the shift (caught by the bit-exactness tests)
  int k;
  asm("bfe.s32 %0, %1, 23, 9;" : "=r"(k) : "r"(tmp));
  const double kd = __int2double_rn(k);  // conversion pipe; keeps an FP64 add (two issue cycles) out of the cell
  const int iz = ix - (tmp & 0xff800000);
  const GlogEntry e = *reinterpret_cast<const GlogEntry*>(reinterpret_cast<const char*>(tab) + ((tmp >> 15) & 0xf0));
  const double z = (double)__int_as_float(iz);  // exact (cvt.f64.f32)
  const double r = __fma_rn(z, e.invc, -1.0);
  const double y0 = __fma_rn(kd, PS_GLOGF_LN2, e.logc);
  double y = __fma_rn(r, PS_GLOGF_A1, PS_GLOGF_A2);
  const double r2 = __dmul_rn(r, r);
  const double q = __dadd_rn(r, y0);
  y = __fma_rn(r2, PS_GLOGF_A0, y);
  y = __fma_rn(r2, y, q);
  return __double2float_rn(y);
}

__device__ __forceinline__ float glibc_logf_fast(float x, const GlogEntry* __restrict__ tab = g_glogf_tab) {
  // x == 1 needs no special case here: k = 0, the table entry around 1 is {1, 0}, r = 0 and the sum is +0
  return glibc_logf_core(__float_as_int(x), tab);
}

__device__ __forceinline__ float glibc_logf(float x) {
  int ix = __float_as_int(x);
  if (ix == 0x3f800000) return 0.0f;
  if ((unsigned)(ix - 0x00800000) >= 0x7f800000u - 0x00800000u) {
    if ((ix & 0x7fffffff) == 0) return -__int_as_float(0x7f800000);
    if (ix == 0x7f800000) return x;
    if (ix < 0 || (ix & 0x7f800000) == 0x7f800000) return __int_as_float(0x7fc00000);
    ix = __float_as_int(__fmul_rn(x, 0x1p23f)) - (23 << 23);
  }
  return glibc_logf_core(ix);
}

}  // namespace ps
