// exhaustive-ish check: q = RN(a/b) from y = RN(1/b) via fma corrections
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <omp.h>
static inline uint64_t rotl(uint64_t x,int k){return (x<<k)|(x>>(64-k));}
typedef struct {uint64_t s[4];} rng_t;
static inline uint64_t next(rng_t* r){uint64_t* s=r->s; uint64_t res=rotl(s[1]*5,7)*9,t=s[1]<<17; s[2]^=s[0]; s[3]^=s[1]; s[1]^=s[2]; s[0]^=s[3]; s[2]^=t; s[3]=rotl(s[3],45); return res;}
static inline double mk(uint64_t bits,int e){ uint64_t u=(bits&0x000FFFFFFFFFFFFFull)|((uint64_t)(e+1023)<<52)|(bits&0x8000000000000000ull); double d; memcpy(&d,&u,8); return d;}
static inline double fd1(double a,double b,double y){ double q0=a*y; double r0=fma(-b,q0,a); return fma(r0,y,q0);}
static inline double fd2(double a,double b,double y){ double q0=a*y; double r0=fma(-b,q0,a); double q1=fma(r0,y,q0); double r1=fma(-b,q1,a); return fma(r1,y,q1);}
int main(){
  long long N=4000000000LL; long long bad1=0,bad2=0;
  #pragma omp parallel reduction(+:bad1,bad2)
  { rng_t r; int t=omp_get_thread_num(); r.s[0]=0x9E3779B97F4A7C15ull*(t+1); r.s[1]=0xBF58476D1CE4E5B9ull^t; r.s[2]=0x94D049BB133111EBull+t; r.s[3]=12345+t*777; for(int i=0;i<20;i++) next(&r);
    #pragma omp for schedule(static)
    for(long long i=0;i<N;i++){
      uint64_t x=next(&r), z=next(&r), m=next(&r);
      int ea=(int)(m%601)-300, eb=(int)((m>>10)%601)-300; if (ea-eb > 900 || eb-ea > 900) { ea = 0; }
      // adversarial mantissas part of the time: all-ones, near-power-of-two, few bits
      int kind=(m>>20)&7;
      if(kind==0){ z|=0x000FFFFFFFFFF000ull; } else if(kind==1){ z&=0xFFF0000000000FFFull; } else if(kind==2){ x|=0x000FFFFFFFFFFF00ull; } else if(kind==3){ x&=0xFFF00000000000FFull; z&=0xFFF00000000000FFull; }
      double a=mk(x,ea), b=mk(z,eb);
      double y=1.0/b; double q=a/b;
      if(fd1(a,b,y)!=q) bad1++;
      if(fd2(a,b,y)!=q) bad2++;
    }
  }
  printf("N=%lld  one-correction mismatches=%lld  two-correction mismatches=%lld\n",N,bad1,bad2);
  // products of a few structured cases: a = b*k +- ulp
  long long bad=0,cnt=0;
  for(int k=1;k<2000;k++) for(int j=0;j<2000;j++){ double b=1.0+j*2.220446049250313e-16*7919; double a=nextafter(b*k, (j&1)?1e300:-1e300); double y=1.0/b; if(fd2(a,b,y)!=a/b) bad++; cnt++; }
  printf("structured %lld cases, two-correction mismatches=%lld\n",cnt,bad);
  return 0; }
