/* Host-side brute-force check of the FMA-corrected division by C_LIGHT used by the pass kernel (div_by_c in
 * mcrat_b200/csrc/pass_kernels.cuh): 4e9 random significands over 41 binades, structured significands, and the neighbours
 * of exact multiples of C_LIGHT, each compared with the IEEE division.
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp tools/div_by_c_check.c -o /tmp/div_by_c_check -lm && /tmp/div_by_c_check
 * prints "mismatches 0 of 4000000000" and "structured mismatches 0"; an argument sets the number of random samples
 * (tests/test_mathlib.py runs 2e8). */
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <omp.h>
#include <stdlib.h>
static const double C = 2.99792458e10;
static inline uint64_t sm64(uint64_t *s){uint64_t z=(*s+=0x9e3779b97f4a7c15ULL);z=(z^(z>>30))*0xbf58476d1ce4e5b9ULL;z=(z^(z>>27))*0x94d049bb133111ebULL;return z^(z>>31);}
int main(int argc, char **argv){
  const double rc = 1.0/C;
  long bad=0; long N = argc > 1 ? atol(argv[1]) : 4000000000L;
  #pragma omp parallel reduction(+:bad)
  {
    uint64_t s = 12345u + 977u*omp_get_thread_num();
    #pragma omp for
    for(long k=0;k<N;k++){
      uint64_t m = sm64(&s);
      /* random significand, exponent 0 (the result's significand depends only on x's significand up to a power of 2) and a few other exponents */
      uint64_t bits = (m & 0x000fffffffffffffULL) | ((uint64_t)(1023 + (int)((m>>52)%41) - 20) << 52);
      double x; memcpy(&x,&bits,8);
      double q = x*rc;
      double r = fma(-q, C, x);
      double q2 = fma(r, rc, q);
      double t = x/C;
      if (q2!=t) bad++;
    }
  }
  printf("mismatches %ld of %ld\n", bad, N);
  /* structured: significands near the all-ones / powers of two, and multiples of C's significand */
  long bad2=0;
  for (uint64_t m=0;m<(1u<<24);m++) for(int hi=0;hi<4;hi++){
    uint64_t sig = hi==0? m : hi==1? (0x000fffffffffffffULL - m) : hi==2 ? (m<<28) : ((m<<28)|0xfffffff);
    uint64_t bits = sig | (1023ULL<<52); double x; memcpy(&x,&bits,8);
    double q=x*rc, r=fma(-q,C,x), q2=fma(r,rc,q); if(q2!=x/C) bad2++;
  }
  /* exact multiples and half-way neighbours: x = RN(k*C) +- ulp */
  for (uint64_t k=1;k<(1u<<24);k++){ double x0=(double)k*C; for(int dlt=-2;dlt<=2;dlt++){ double x=x0; for(int j=0;j<abs(dlt);j++) x=nextafter(x, dlt>0?INFINITY:0);
    double q=x*rc, r=fma(-q,C,x), q2=fma(r,rc,q); if(q2!=x/C) bad2++; }}
  printf("structured mismatches %ld\n", bad2);
  return 0;}
