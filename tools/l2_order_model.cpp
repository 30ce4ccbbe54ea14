// tools/l2_order_model.cpp — sector-level model of K1's DRAM read traffic on an nb^3-block periodic box for a given block order
// (development tool: g++ -O2 -o l2sim tools/l2_order_model.cpp && ./l2sim 64 xslab 12 1.0 48e6).  A sector hits iff fewer than C bytes were
// inserted into the cache since its last access (writes count with weight wfac); calibrated on the Morton order (C = 24-48 MB gives
// the +10.6 % over-read ncu measured at 512^3), it gave the direction
// (x-slab orders read less than the Morton curve) and the capacity cliff before the GPU runs, and under-estimated the cost of the tile faces:
// the measured numbers are in DESIGN.md section 4.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <string>
#include <cstring>
static uint64_t s3(uint32_t v){uint64_t x=v&0x1fffff;x=(x|x<<32)&0x1f00000000ffffULL;x=(x|x<<16)&0x1f0000ff0000ffULL;x=(x|x<<8)&0x100f00f00f00f00fULL;x=(x|x<<4)&0x10c30c30c30c30c3ULL;x=(x|x<<2)&0x1249249249249249ULL;return x;}
static uint64_t m3(uint32_t x,uint32_t y,uint32_t z){return s3(x)|(s3(y)<<1)|(s3(z)<<2);}
static uint64_t m2(uint32_t x,uint32_t y){return s3(x)|(s3(y)<<1);} // fine as an order
int main(int argc,char**argv){
  int nbx=atoi(argv[1]), nby=nbx, nbz=nbx; std::string order=argv[2]; int T=argc>3?atoi(argv[3]):8; double wfac=argc>4?atof(argv[4]):1.0;
  int N=nbx*nby*nbz;
  std::vector<int> ord(N); std::vector<uint64_t> key(N);
  for(int z=0;z<nbz;z++)for(int y=0;y<nby;y++)for(int x=0;x<nbx;x++){int i=x+nbx*(y+nby*z);
    if(order=="morton") key[i]=m3(x,y,z);
    else if(order=="xslab") key[i]=((m2(y/T,z/T)*(nbx+1)+x)*T+(y%T))*T+(z%T);
    else if(order=="raster") key[i]=i;
    else if(order=="xslabm") key[i]=((m2(y/T,z/T)*(nbx+1)+x)*64*64)+m2(y%T,z%T);
    else if(order=="xpair") key[i]=(m3(x/2,y,z)<<1)|(x&1);   // Morton over x-pairs
    else if(order=="x4") key[i]=(m3(x/4,y,z)<<2)|(x&3);
    else if(order=="x8") key[i]=(m3(x/8,y,z)<<3)|(x&7);
    else {fprintf(stderr,"order?\n");return 1;}
    ord[i]=i;}
  std::sort(ord.begin(),ord.end(),[&](int a,int b){return key[a]<key[b];});
  // sector timestamps: f: block*27*64 + k*64 + row ; vel: block*3*64 + c*64 + row
  std::vector<uint32_t> tf((size_t)N*27*64,0), tv((size_t)N*3*64,0);
  std::vector<double> Cs={16e6,24e6,32e6,48e6,64e6,96e6,128e6};
  int nc=Cs.size(); std::vector<uint64_t> miss(nc,0); uint64_t acc=0;
  // one clock per capacity (inserted sectors differ) -> approximate with ONE clock driven by the middle capacity? use separate runs instead
  double C=argc>5?atof(argv[5]):48e6; uint32_t Csec=(uint32_t)(C/32);
  uint32_t clk=Csec+1; uint64_t m=0;
  auto touch=[&](uint32_t&t){acc++; if(clk-t>=Csec){m++;clk++;} t=clk;};
  auto blk=[&](int x,int y,int z)->long{ x=(x+nbx)%nbx; y=(y+nby)%nby; z=(z+nbz)%nbz; return x+nbx*(y+(long)nby*z);};
  uint32_t wsec=(uint32_t)((27*64+3*64+64)*wfac);
  for(int oi=0;oi<N;oi++){int b=ord[oi]; int bx=b%nbx, by=(b/nbx)%nby, bz=b/(nbx*nby);
    for(int k=0;k<27;k++){int cx=k%3-1, cy=(k/3)%3-1, cz=k/9-1;
      for(int z=0;z<8;z++)for(int y=0;y<8;y++){int sy=y-cy, sz=z-cz; int oy=sy<0?-1:sy>7?1:0, oz=sz<0?-1:sz>7?1:0; int row=(sz&7)*8+(sy&7);
        long sb=blk(bx,by+oy,bz+oz); touch(tf[((size_t)sb*27+k)*64+row]);
        if(cx!=0){long xb=blk(bx-cx,by+oy,bz+oz); touch(tf[((size_t)xb*27+k)*64+row]);}
      }}
    for(int c=0;c<3;c++){
      for(int r=0;r<64;r++){touch(tv[((size_t)b*3+c)*64+r]); touch(tv[((size_t)blk(bx-1,by,bz)*3+c)*64+r]); touch(tv[((size_t)blk(bx+1,by,bz)*3+c)*64+r]);}
      for(int z=0;z<8;z++){touch(tv[((size_t)blk(bx,by-1,bz)*3+c)*64+z*8+7]); touch(tv[((size_t)blk(bx,by+1,bz)*3+c)*64+z*8+0]);}
      for(int y=0;y<8;y++){touch(tv[((size_t)blk(bx,by,bz-1)*3+c)*64+56+y]); touch(tv[((size_t)blk(bx,by,bz+1)*3+c)*64+y]);}
    }
    clk+=wsec;
  }
  double minb=(double)N*512*120;
  printf("%s T=%d C=%.0fMB wfac=%.2f: accesses %.1fM, read %.3f GB, min %.3f GB, over %.2f%%, B/LU read %.1f\n",order.c_str(),T,C/1e6,wfac,acc/1e6,m*32/1e9,minb/1e9,(m*32/minb-1)*100,m*32.0/(N*512.0));
}
