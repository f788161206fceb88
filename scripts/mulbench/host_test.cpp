#include "mul29.cuh"
extern "C" void mul29_fq(const uint32_t* a, const uint32_t* b, uint32_t* r, int variant) {
  L29<Fq29> x, y; for (int i=0;i<9;++i){x.l[i]=a[i];y.l[i]=b[i];}
  L29<Fq29> z = variant ? mul29<Fq29,1>(x,y) : mul29<Fq29,0>(x,y);
  for (int i=0;i<9;++i) r[i]=z.l[i];
}
extern "C" void mul29_fr(const uint32_t* a, const uint32_t* b, uint32_t* r, int variant) {
  L29<Fr29> x, y; for (int i=0;i<9;++i){x.l[i]=a[i];y.l[i]=b[i];}
  L29<Fr29> z = variant ? mul29<Fr29,1>(x,y) : mul29<Fr29,0>(x,y);
  for (int i=0;i<9;++i) r[i]=z.l[i];
}
