#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
/* tools/check_glibc_log.c -- host check that the operation sequence restated in
 * partitionsolvers_b200/csrc/glibc_log.cuh IS glibc's log(): the same sequence in C (explicit fma, contraction
 * off) against libm on 2*10^8 arguments, near 1, over the ratio range and over 2^-90..2^90.
 *   gcc -O2 -mfma -ffp-contract=off -I partitionsolvers_b200/csrc -o /tmp/chk tools/check_glibc_log.c -lm && /tmp/chk
 * Expected output: "mismatches 0 of 200000000" on any host whose glibc selects the FMA variant of log. */
#include "glibc_log_data.inc"
static const double LN2HI = PS_GLOG_LN2HI, LN2LO = PS_GLOG_LN2LO;
static const double A_[5] = {PS_GLOG_A0, PS_GLOG_A1, PS_GLOG_A2, PS_GLOG_A3, PS_GLOG_A4};
static const double B_[11] = {PS_GLOG_B0, PS_GLOG_B1, PS_GLOG_B2, PS_GLOG_B3, PS_GLOG_B4, PS_GLOG_B5,
                              PS_GLOG_B6, PS_GLOG_B7, PS_GLOG_B8, PS_GLOG_B9, PS_GLOG_B10};
static const struct { double invc, logc; } TABS[128] = PS_GLOG_TAB_INIT;
static inline uint64_t asu(double x){uint64_t u; memcpy(&u,&x,8); return u;}
static inline double asd(uint64_t u){double x; memcpy(&x,&u,8); return x;}
static double mylog(double x){
  uint64_t ix=asu(x);
  const uint64_t LO=asu(1.0-0x1p-4), HI=asu(1.0+0x1.09p-4);
  if (ix-LO < HI-LO){
    if (ix==asu(1.0)) return 0;
    double r=x-1.0;
    double t2=__builtin_fma(r,B_[2],B_[1]);
    double t3=__builtin_fma(r,B_[5],B_[4]);
    double r2=r*r;
    double t5=__builtin_fma(r,B_[8],B_[7]);
    t2=__builtin_fma(r2,B_[3],t2);
    t3=__builtin_fma(r2,B_[6],t3);
    double r3=r*r2;
    double t1=__builtin_fma(r2,B_[9],t5);
    t1=__builtin_fma(r3,B_[10],t1);
    t1=__builtin_fma(t1,r3,t3);
    t1=__builtin_fma(t1,r3,t2);
    double w=__builtin_fma(r,0x1p27,r);
    double rhi=__builtin_fma(-0x1p27,r,w);
    double rhi2=rhi*rhi;
    double rlo=r-rhi;
    double hi=__builtin_fma(rhi2,B_[0],r);
    double d=r-hi;
    double s=r+rhi;
    double lo=__builtin_fma(rhi2,B_[0],d);
    double m=B_[0]*rlo;
    lo=__builtin_fma(m,s,lo);
    double y=__builtin_fma(t1,r3,lo);
    return hi+y;
  }
  uint64_t tmp=ix-0x3fe6000000000000ULL;
  int i=(tmp>>45)&127;
  int k=(int64_t)tmp>>52;
  uint64_t iz=ix-(tmp&(0xfffULL<<52));
  double invc=TABS[i].invc, logc=TABS[i].logc;
  double z=asd(iz);
  double kd=(double)k;
  double w=__builtin_fma(kd,LN2HI,logc);
  double r=__builtin_fma(z,invc,-1.0);
  double p12=__builtin_fma(r,A_[2],A_[1]);
  double hi=r+w;
  double r2=r*r;
  double lo=(w-hi)+r;
  lo=__builtin_fma(kd,LN2LO,lo);
  double r3=r*r2;
  double p34=__builtin_fma(r,A_[4],A_[3]);
  lo=__builtin_fma(r2,A_[0],lo);
  double p=__builtin_fma(p34,r2,p12);
  double y=__builtin_fma(r3,p,lo);
  return y+hi;
}
int main(){
  uint64_t s=88172645463325252ULL; long bad=0; long N=200000000;
  for(long it=0;it<N;++it){
    s^=s<<13; s^=s>>7; s^=s<<17;
    double x;
    int mode=it&3;
    if(mode==0){ // near 1
      x=1.0+((double)(int64_t)(s>>11)/9007199254740992.0-0.5)*0.14;
    } else if(mode==1){ // ratio range
      x=0.1+((double)(s>>11)/9007199254740992.0)*20.0;
    } else { // wide exponent range 2^-90..2^90
      int e=(int)((s>>52)%181)-90; x=ldexp(1.0+(double)((s>>11)&0xfffffffffffffULL)/4503599627370496.0,e);
    }
    double a=mylog(x), b=log(x);
    if(asu(a)!=asu(b)){ if(bad<5) printf("x=%a mine=%a glibc=%a\n",x,a,b); ++bad; }
  }
  printf("mismatches %ld of %ld\n",bad,N);
  return 0;
}
