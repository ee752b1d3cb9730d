#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include "qr_kscene.h"
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET);
  void*b=malloc(n); if (fread(b,1,n,f) != (size_t)n) return 1; fclose(f);
  qr_kpacker pk; void*out=aligned_alloc(64,(size_t)1<<30);
  for(int r=0;r<5;r++){
    auto t0=std::chrono::steady_clock::now();
    int rc=pk.plan(b);
    auto t1=std::chrono::steady_clock::now();
    pk.write(out);
    auto t2=std::chrono::steady_clock::now();
    printf("rc %d plan %.1f us write %.1f us bytes %zu\n",rc,std::chrono::duration<double,std::micro>(t1-t0).count(),std::chrono::duration<double,std::micro>(t2-t1).count(),pk.bytes());
  }
}
