// emul_build.h — host-side copy of the device scene build (Morton LBVH, 64-byte nodes, triangle records) for the offline experiments.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <numeric>
#include <vector_types.h>
#include <vector_functions.h>
#include "../../nlos_surface_optimization_b200/csrc/nlos_core.cuh"
using namespace nlos;
struct B6 { float lo[3], hi[3]; };
static inline B6 uni(const B6& a, const B6& b) { B6 r; for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); } return r; }
static inline f3 ldv(const float* v, int i) { return mk3(v[3*i], v[3*i+1], v[3*i+2]); }
struct EmulScene { std::vector<float4> ttris, stris; std::vector<BvhNode> nodes; std::vector<int> order; float pad = 0; float vlo[3], vhi[3]; };
static inline void emul_build(const float* origin, int L, const float* verts, int V, const int* faces, int F, int leafmax, float padscale, EmulScene& S) {
  std::vector<int>& order = S.order; std::vector<BvhNode>& nodes = S.nodes; std::vector<float4>& ttris = S.ttris; std::vector<float4>& stris = S.stris;
  float absmax = 0; for (int i = 0; i < 3*V; ++i) absmax = std::max(absmax, fabsf(verts[i])); for (int i = 0; i < 3*L; ++i) absmax = std::max(absmax, fabsf(origin[i]));
  const float pad = absmax * padscale;
  std::vector<B6> leaf(F); std::vector<f3> cen(F);
  for (int f = 0; f < F; ++f) { f3 a = ldv(verts, faces[3*f]), b = ldv(verts, faces[3*f+1]), c = ldv(verts, faces[3*f+2]);
    B6& bx = leaf[f]; bx.lo[0]=fminf(a.x,fminf(b.x,c.x))-pad; bx.lo[1]=fminf(a.y,fminf(b.y,c.y))-pad; bx.lo[2]=fminf(a.z,fminf(b.z,c.z))-pad; bx.hi[0]=fmaxf(a.x,fmaxf(b.x,c.x))+pad; bx.hi[1]=fmaxf(a.y,fmaxf(b.y,c.y))+pad; bx.hi[2]=fmaxf(a.z,fmaxf(b.z,c.z))+pad;
    cen[f] = mk3(0.5f*(bx.lo[0]+bx.hi[0]), 0.5f*(bx.lo[1]+bx.hi[1]), 0.5f*(bx.lo[2]+bx.hi[2])); }
  
  {
    float lo[3]={3e38f,3e38f,3e38f}, hi[3]={-3e38f,-3e38f,-3e38f};
    for (int f=0;f<F;++f){ float cc[3]={cen[f].x,cen[f].y,cen[f].z}; for(int k=0;k<3;++k){lo[k]=std::min(lo[k],cc[k]);hi[k]=std::max(hi[k],cc[k]);} }
    std::vector<uint64_t> keys(F);
    for (int f=0;f<F;++f) keys[f] = ((uint64_t)morton30((cen[f].x-lo[0])/(hi[0]-lo[0]),(cen[f].y-lo[1])/(hi[1]-lo[1]),(cen[f].z-lo[2])/(hi[2]-lo[2]))<<32)|(uint32_t)f;
    std::sort(keys.begin(), keys.end()); order.resize(F); for (int p=0;p<F;++p) order[p]=(int)(uint32_t)keys[p];
    const int NI=F-1; std::vector<int> first(NI),last(NI),cl(NI),cr(NI);
    for (int i=0;i<NI;++i){int a,b,s; lbvh_range(keys.data(),F,i,a,b,s); first[i]=a;last[i]=b;cl[i]=(a==s)?~s:s;cr[i]=(b==s+1)?~(s+1):s+1;}
    std::vector<B6> nb(NI); std::vector<int> ord2; std::vector<int> st={0}; while(!st.empty()){int n=st.back();st.pop_back();ord2.push_back(n);if(cl[n]>=0)st.push_back(cl[n]);if(cr[n]>=0)st.push_back(cr[n]);}
    for (int k=(int)ord2.size()-1;k>=0;--k){int n=ord2[k]; B6 a=cl[n]<0?leaf[order[~cl[n]]]:nb[cl[n]], b=cr[n]<0?leaf[order[~cr[n]]]:nb[cr[n]]; nb[n]=uni(a,b);}
    nodes.resize(NI);
    for (int i=0;i<NI;++i){int link[2],cnt[2];B6 bx[2];int cc[2]={cl[i],cr[i]};
      for(int k=0;k<2;++k){int c=cc[k]; if(c<0){link[k]=~c;cnt[k]=1;bx[k]=leaf[order[~c]];} else {int size=last[c]-first[c]+1;bx[k]=nb[c]; if(size<=leafmax){link[k]=first[c];cnt[k]=size;} else {link[k]=c;cnt[k]=0;}}}
      BvhNode n; n.a=make_float4(bx[0].lo[0],bx[0].lo[1],bx[0].lo[2],bx[0].hi[0]); n.b=make_float4(bx[0].hi[1],bx[0].hi[2],bx[1].lo[0],bx[1].lo[1]); n.c=make_float4(bx[1].lo[2],bx[1].hi[0],bx[1].hi[1],bx[1].hi[2]); n.d=make_int4(cnt[0]>0?leaf_ref(link[0],cnt[0]):link[0], cnt[1]>0?leaf_ref(link[1],cnt[1]):link[1],0,0); nodes[i]=n;}
  }
  ttris.assign(4*(size_t)F, make_float4(0,0,0,0)); stris.assign(4*(size_t)F, make_float4(0,0,0,0));
  for (int p=0;p<F;++p){int f=order[p]; f3 v1=ldv(verts,faces[3*f]),v2=ldv(verts,faces[3*f+1]),v3=ldv(verts,faces[3*f+2]); TriRec tr=make_tri(v1,v2,v3);
    ttris[4*p]=make_float4(tr.v0.x,tr.v0.y,tr.v0.z,i2f(f)); ttris[4*p+1]=make_float4(tr.e1.x,tr.e1.y,tr.e1.z,0); ttris[4*p+2]=make_float4(tr.e2.x,tr.e2.y,tr.e2.z,0); ttris[4*p+3]=make_float4(tr.Ng.x,tr.Ng.y,tr.Ng.z,0);
    f3 N=cross3(v2-v1,v3-v1); float A=len3(N)/2; f3 nf=N/(2*A); stris[4*p]=make_float4(v1.x,v1.y,v1.z,A); stris[4*p+1]=make_float4(v2.x,v2.y,v2.z,nf.x); stris[4*p+2]=make_float4(v3.x,v3.y,v3.z,nf.y); stris[4*p+3]=make_float4(nf.z,0,0,0);}

  S.pad = pad;
  for (int k = 0; k < 3; ++k) { S.vlo[k] = 3e38f; S.vhi[k] = -3e38f; }
  for (int i = 0; i < V; ++i) for (int k = 0; k < 3; ++k) { S.vlo[k] = std::min(S.vlo[k], verts[3*i+k]); S.vhi[k] = std::max(S.vhi[k], verts[3*i+k]); }
}
