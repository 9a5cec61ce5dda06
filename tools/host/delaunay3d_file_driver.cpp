#include <algorithm>
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "distmesh_host.h"
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); fseek(f,0,SEEK_SET); int64_t n=sz/24; std::vector<double> p(3*n); if(fread(p.data(),8,3*n,f)!=(size_t)(3*n)) return 1; fclose(f);
  int nth=atoi(argv[2]);
  int64_t cap=dmh_delaunay3d_max_cells(n),T,d,l; std::vector<int32_t> c(4*cap);
  for(int r=0;r<2;r++){
  auto t0=std::chrono::steady_clock::now();
  int rc=dmh_delaunay3d_mt(p.data(),n,c.data(),cap,&T,&d,&l,nth);
  double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
  printf("rc=%d N=%ld T=%ld dups=%ld lost=%ld %.3fs %.2f us/pt\n",rc,n,T,d,l,s,s/n*1e6);}
}
