#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "distmesh_host.h"
int main(int argc,char**argv){
  int64_t n=atol(argv[1]); int nth=atoi(argv[2]);
  std::mt19937_64 g(1); std::uniform_real_distribution<double> u(-1,1);
  std::vector<double> p; p.reserve(3*n);
  while((int64_t)p.size()<3*n){double x=u(g),y=u(g),z=u(g); if(x*x+y*y+z*z<1){p.push_back(x);p.push_back(y);p.push_back(z);}}
  int64_t cap=dmh_delaunay3d_max_cells(n),T,d,l; std::vector<int32_t> c(4*cap), c1(4*cap);
  int64_t T1;
  dmh_delaunay3d(p.data(),n,c1.data(),cap,&T1,&d,&l);
  for(int r=0;r<3;r++){
  auto t0=std::chrono::steady_clock::now();
  int rc=dmh_delaunay3d_mt(p.data(),n,c.data(),cap,&T,&d,&l,nth);
  double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("rc=%d T=%ld dups=%ld lost=%ld %.3fs %.2f us/pt\n",rc,T,d,l,s,s/n*1e6);}
  // same cell set (and order up to column order)?
  bool same = T==T1;
  for(int64_t i=0;i<T && same;i++){ std::array<int,4> a{c[4*i],c[4*i+1],c[4*i+2],c[4*i+3]}, b{c1[4*i],c1[4*i+1],c1[4*i+2],c1[4*i+3]}; std::sort(a.begin(),a.end()); std::sort(b.begin(),b.end()); same = a==b; }
  printf("same as serial: %d\n",(int)same);
  // the triangulation that stays around: 3/4 of the points built, the last quarter (shifted beyond a side of the
  // hull: a ghost layer) inserted as one batch
  { int64_t n0=3*n/4, m=n-n0; std::vector<double> q(p.begin()+3*n0,p.end()); for(int64_t i=0;i<m;i++) q[3*i+1]+=2.0;
    int rc=0; void*h=dmh_dt3_build(p.data(),n0,nth,&rc); int rc2=dmh_dt3_insert(h,q.data(),m); int64_t T2=0;
    int rc3=dmh_dt3_cells(h,c.data(),cap,&T2,&d,&l); dmh_dt3_free(h);
    printf("build %ld + insert %ld: rc=%d/%d/%d T=%ld dups=%ld lost=%ld\n",n0,m,rc,rc2,rc3,T2,d,l); }
}
