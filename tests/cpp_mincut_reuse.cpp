// The context keeps ONE SinkForestCut and reuses it (32-bit / 64-bit flow storage, time stamps that are never cleared):
// a reused object must give the flow value and the labels of a fresh one, for changing sizes and capacities.
#include "../slowflow_b200/csrc/sf_gridcut.hpp"
#include <cstdio>
#include <random>
using namespace sf;
int main(){
  std::mt19937 g(11);
  SinkForestCut reused;
  int w=57,h=43; size_t n=(size_t)w*h; long bad=0;
  for(int it=0;it<300;it++){
    if(it%50==49){w=20+g()%60;h=20+g()%60;n=(size_t)w*h;}
    std::vector<int64_t> a(n),b;
    for(size_t p=0;p<n;p++) a[p]=(g()%15==0)? -(int64_t)(g()%200000) : (int64_t)(g()%4000);
    b=a; int64_t pair=(int64_t)(g()%900);
    if(it%7==0) pair=((int64_t)1<<31)+g()%1000; // wide path
    SinkForestCut fresh;
    int64_t f1=reused.solve(w,h,a.data(),pair), f2=fresh.solve(w,h,b.data(),pair);
    if(f1!=f2) bad++;
    for(size_t p=0;p<n;p++) if(reused.label(p)!=fresh.label(p)) bad++;
  }
  printf("mismatches %ld\n",bad);
  return bad!=0;
}
