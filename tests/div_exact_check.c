/* div_exact_check.c - CPU check of the division k_eval's DDA set-up uses (fcvm_device.cuh: div_exact): a / b from the correctly
 * rounded reciprocal y = RN(1 / b) with two fused multiply-adds, q0 = RN(a y), r = fma(-b, q0, a), q = fma(r, y, q0), against the
 * hardware divider, for every integer divisor of the reciprocal table (1 .. 1023) and for the grid resolutions.  The only misses
 * allowed are subnormal numerators, which the marchers never produce (1 - s >= 2^-53).  Prints the counts; exit code 1 on a miss
 * among normal numbers.   gcc -O2 -mfma -ffp-contract=off tests/div_exact_check.c -lm */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
int main(){
  long bad=0, n=0;
  for (int d=1; d<1024; d++){
    double y = 1.0/(double)d;
    for (int k=0;k<40000;k++){
      double a;
      switch(k&3){ case 0: a=(double)(rnd()>>11)/9007199254740992.0; break;
                   case 1: a=1.0-ldexp((double)(rnd()>>11)/9007199254740992.0,-(int)(rnd()%52)); break;
                   case 2: a=ldexp((double)(rnd()>>11)/9007199254740992.0,-(int)(rnd()%60)); break;
                   default: a=(double)(rnd()%1000)/1000.0; }
      if (k==0) a=1.0; if (k==1) a=0.0; if (k==2) a=4.9e-324; if(k==3) a=1e-310;
      double q0=a*y; double r=__builtin_fma(-(double)d,q0,a); double q1=__builtin_fma(r,y,q0);
      double t=a/(double)d; n++;
      if (q1!=t && fabs(a) >= 2.3e-308){ bad++; if(bad<10) printf("mismatch a=%a d=%d q1=%a true=%a\n",a,d,q1,t);}    
    }
  }
  printf("tested %ld, mismatches %ld\n", n, bad);
  // division by res
  double ress[]={0.1,0.2,0.4,0.25,0.05,0.3,1.0,0.123456789};
  for(int i=0;i<8;i++){ double res=ress[i], y=1.0/res; long b=0; for(long k=0;k<3000000;k++){ double a=((double)(int64_t)(rnd()>>11)/9007199254740992.0-0.5)*ldexp(1.0,(int)(rnd()%12)); if(k&1) a=(double)(float)a; double q0=a*y; double r=__builtin_fma(-res,q0,a); double q1=__builtin_fma(r,y,q0); if(q1!=a/res) b++; } printf("res %g mismatches %ld\n",res,b); bad+=b;}  
  return bad ? 1 : 0; }
