/* headbench.c -- cost of reading the head of a BLOW5 record on one core: sf_s5_parse_head(), and its two decoders
 * side by side (sfinflate.c: the table decoder stopped after 384 bytes, the table-free prefix decoder).
 * gcc -O2 -std=c99 -D_GNU_SOURCE -Iinclude -Isigfish_b200/host -o /tmp/headbench tools/hostfeed/headbench.c \
 *     sigfish_b200/host/s5read.c sigfish_b200/host/sfinflate.c -lz -lm
 * /tmp/headbench reads.blow5 100000   (tools/hostfeed/run.py writes such files under /tmp/sigfish_b200/hostfeed) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "s5read.h"
#include "sfinflate.h"
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){
  char err[256]; sf_s5file_t*f=sf_s5_open(argv[1],err,256); if(!f){puts(err);return 1;}
  int N=atoi(argv[2]); const char**ptr=malloc(sizeof(*ptr)*N); int64_t*sz=malloc(8*N); int n=0;
  while(n<N){ sz[n]=sf_s5_get_next_view(f,&ptr[n]); if(sz[n]<=0)break; n++; }
  sf_rec_t r; memset(&r,0,sizeof r); char*scr=NULL; size_t cap=0; int32_t pos; int64_t sb; long bad=0; unsigned long acc=0;
  for(int rep=0;rep<3;rep++){ double t0=now();
   for(int i=0;i<n;i++){ if(sf_s5_parse_head(f,ptr[i],sz[i],&r,&pos,&sb,&scr,&cap))bad++; acc+=r.len_raw_signal+pos; }
   double t=now()-t0; printf("%d records, %.2f us per head (bad %ld, acc %lu)\n",n,t/n*1e6,bad,acc);}
  /* how many were answered by the prefix decoder */
  sf_inflater*d=calloc(1,sizeof *d); unsigned char out[400]; size_t got; long ans=0;
  for(int i=0;i<n;i++) ans+=sf_zlib_inflate_prefix(d,(const uint8_t*)ptr[i],sz[i],out,384,48,&got);
  printf("prefix answered %ld of %d\n",ans,n);
  for(int rep=0;rep<2;rep++){double t0=now(); long k=0; for(int i=0;i<n;i++){ k+=sf_zlib_inflate(d,(const uint8_t*)ptr[i],sz[i],out,384,&got);} printf("table decoder, 384 bytes: %.2f us (%ld)\n",(now()-t0)/n*1e6,k);}
  for(int rep=0;rep<2;rep++){double t0=now(); long k=0; for(int i=0;i<n;i++){ k+=sf_zlib_inflate_prefix(d,(const uint8_t*)ptr[i],sz[i],out,384,48,&got);} printf("prefix decoder: %.2f us (%ld)\n",(now()-t0)/n*1e6,k);}
  return 0;}
