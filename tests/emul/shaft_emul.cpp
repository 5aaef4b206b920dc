// shaft_emul.cpp — offline experiment (not a test, not product): cost model of the round-2 forward design.
//
// Idea under test: all rays of one (wall point, triangle tile) pair leave the same origin and end inside the tile's bounding
// box, so the set of BVH leaves ANY of them can touch is bounded by the shaft {o + t (p - o) : p in tile box, t in [0,1]} —
// a "fat ray" whose box test is an interval slab test.  One fat traversal per (tile, source) yields a short candidate-leaf
// list; the tile's rays then test only those leaves (leaf box, then triangles), no per-ray tree walk.
// The harness counts, on the real mesh: fat-traversal node visits, candidates per list, per-ray leaf-box hits and triangle
// tests, and checks that the visibility answer equals occluded() for every ray.
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
static B6 uni(const B6& a, const B6& b) { B6 r; for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); } return r; }
static f3 ldv(const float* v, int i) { return mk3(v[3*i], v[3*i+1], v[3*i+2]); }

struct Fat { float o[3], qlo[3], qhi[3]; };
// exact shaft-vs-box test (interval slab): exists t in [0, tmax] with o + t*[qlo,qhi] meeting [blo,bhi] on every axis
static bool fat_box(const Fat& f, const float* blo, const float* bhi, float tmaxv) {
  float lower = 0.f, upper = tmaxv;
  for (int a = 0; a < 3; ++a) {
    const float H = bhi[a] - f.o[a], Lw = blo[a] - f.o[a];
    const float ql = f.qlo[a], qh = f.qhi[a];
    if (ql > 1e-20f) upper = std::min(upper, H / ql); else if (ql < -1e-20f) lower = std::max(lower, H / ql); else if (H < 0 && ql == 0.f) { /* keep */ }
    if (qh > 1e-20f) lower = std::max(lower, Lw / qh); else if (qh < -1e-20f) upper = std::min(upper, Lw / qh);
  }
  return lower <= upper * 1.00001f + 1e-6f;
}

extern "C" {
// out: [0] active pairs, [1] rays, [2] fat node visits / pair, [3] candidates / pair, [4] leaf-box hits / ray, [5] tri tests / ray,
//      [6] mismatches, [7] lane efficiency of the fat traversal (mean/max visits over 32 consecutive sources), [8] max candidates,
//      [9] rays per active pair, [10] per-ray trip efficiency (mean/max leaf hits within a pair), [11] fraction of pairs active
int shaft_emul(const float* origin, int L, const float* verts, int V, const int* faces, int F, int tile, int leafmax, float padscale,
               int src_stride, double* out, int* hist /*128 bins: candidates per pair*/) {
  float absmax = 0; for (int i = 0; i < 3*V; ++i) absmax = std::max(absmax, fabsf(verts[i])); for (int i = 0; i < 3*L; ++i) absmax = std::max(absmax, fabsf(origin[i]));
  const float pad = absmax * padscale;
  std::vector<B6> leaf(F); std::vector<f3> cen(F);
  for (int f = 0; f < F; ++f) { f3 a = ldv(verts, faces[3*f]), b = ldv(verts, faces[3*f+1]), c = ldv(verts, faces[3*f+2]);
    B6& bx = leaf[f]; bx.lo[0]=fminf(a.x,fminf(b.x,c.x))-pad; bx.lo[1]=fminf(a.y,fminf(b.y,c.y))-pad; bx.lo[2]=fminf(a.z,fminf(b.z,c.z))-pad; bx.hi[0]=fmaxf(a.x,fmaxf(b.x,c.x))+pad; bx.hi[1]=fmaxf(a.y,fmaxf(b.y,c.y))+pad; bx.hi[2]=fmaxf(a.z,fmaxf(b.z,c.z))+pad;
    cen[f] = mk3(0.5f*(bx.lo[0]+bx.hi[0]), 0.5f*(bx.lo[1]+bx.hi[1]), 0.5f*(bx.lo[2]+bx.hi[2])); }
  std::vector<int> order; std::vector<BvhNode> nodes;
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
  std::vector<float4> ttris(4*(size_t)F), stris(4*(size_t)F);
  for (int p=0;p<F;++p){int f=order[p]; f3 v1=ldv(verts,faces[3*f]),v2=ldv(verts,faces[3*f+1]),v3=ldv(verts,faces[3*f+2]); TriRec tr=make_tri(v1,v2,v3);
    ttris[4*p]=make_float4(tr.v0.x,tr.v0.y,tr.v0.z,i2f(f)); ttris[4*p+1]=make_float4(tr.e1.x,tr.e1.y,tr.e1.z,0); ttris[4*p+2]=make_float4(tr.e2.x,tr.e2.y,tr.e2.z,0); ttris[4*p+3]=make_float4(tr.Ng.x,tr.Ng.y,tr.Ng.z,0);
    f3 N=cross3(v2-v1,v3-v1); float A=len3(N)/2; f3 nf=N/(2*A); stris[4*p]=make_float4(v1.x,v1.y,v1.z,A); stris[4*p+1]=make_float4(v2.x,v2.y,v2.z,nf.x); stris[4*p+2]=make_float4(v3.x,v3.y,v3.z,nf.y); stris[4*p+3]=make_float4(nf.z,0,0,0);}

  for (int i = 0; i < 128; ++i) hist[i] = 0;
  double n_pairs = 0, n_all_pairs = 0, n_rays = 0, n_visits = 0, n_cand = 0, n_leafhit = 0, n_tri = 0, n_mis = 0, eff_num = 0, eff_den = 0, maxc = 0, trip_num = 0, trip_den = 0;
  const int ntiles = (F + tile - 1) / tile;
#pragma omp parallel for schedule(dynamic, 8) reduction(+:n_pairs,n_all_pairs,n_rays,n_visits,n_cand,n_leafhit,n_tri,n_mis,eff_num,eff_den,trip_num,trip_den) reduction(max:maxc)
  for (int tl = 0; tl < ntiles; ++tl) {
    const int p0 = tl * tile, p1 = std::min(F, p0 + tile);
    B6 tb; for (int k=0;k<3;++k){tb.lo[k]=3e38f;tb.hi[k]=-3e38f;}
    for (int p = p0; p < p1; ++p) for (int j = 0; j < 3; ++j) { const float4 v = stris[4*p+j]; const float c[3]={v.x,v.y,v.z}; for (int k=0;k<3;++k){tb.lo[k]=std::min(tb.lo[k],c[k]-pad);tb.hi[k]=std::max(tb.hi[k],c[k]+pad);} }
    std::vector<int> cand; std::vector<int> local_hist(128, 0);
    int grp_n = 0; double grp_sum = 0, grp_max = 0;
    for (int s = 0; s < L; s += 1) {
      if ((s / 64) % src_stride != 0) continue;          // every src_stride-th wall row, whole rows
      const f3 o = ldv(origin, s);
      n_all_pairs += 1;
      // rays of this pair
      struct R { Ray ray; float ts; int prim; };
      std::vector<R> rays;
      for (int p = p0; p < p1; ++p) { ShadeTri st; TriRec tr; st.v1=xyz(stris[4*p]);st.A=stris[4*p].w;st.v2=xyz(stris[4*p+1]);st.v3=xyz(stris[4*p+2]);st.nf=mk3(stris[4*p+1].w,stris[4*p+2].w,stris[4*p+3].x);
        tr.v0=xyz(ttris[4*p]);tr.e1=xyz(ttris[4*p+1]);tr.e2=xyz(ttris[4*p+2]);tr.Ng=xyz(ttris[4*p+3]); int prim=f2i(ttris[4*p].w);
        SampleGeom g; if(!sample_self_hit(5489,s,prim,0,o,st,tr,g)) continue;
        float ff=-dot3(st.nf,g.d)*g.d.z; if(!(ff>0)) continue;
        R r; r.ray = make_ray(o, g.d); r.ts = g.t; r.prim = prim; rays.push_back(r); }
      if (rays.empty()) { if (++grp_n == 32) { if (grp_max > 0) { eff_num += grp_sum; eff_den += grp_max * 32; } grp_n = 0; grp_sum = 0; grp_max = 0; } continue; }
      n_pairs += 1; n_rays += rays.size();
      Fat f; f.o[0]=o.x; f.o[1]=o.y; f.o[2]=o.z; for (int k=0;k<3;++k){ f.qlo[k]=tb.lo[k]-f.o[k]; f.qhi[k]=tb.hi[k]-f.o[k]; }
      const float tmaxv = 1.00001f;
      cand.clear();
      int visits = 0; int stack[256]; int sp = 0; int cur = 0;
      while (true) {
        const BvhNode& nd = nodes[cur]; ++visits;
        const float b0lo[3]={nd.a.x,nd.a.y,nd.a.z}, b0hi[3]={nd.a.w,nd.b.x,nd.b.y}, b1lo[3]={nd.b.z,nd.b.w,nd.c.x}, b1hi[3]={nd.c.y,nd.c.z,nd.c.w};
        bool h0 = fat_box(f, b0lo, b0hi, tmaxv), h1 = fat_box(f, b1lo, b1hi, tmaxv);
        if (h0 && nd.d.x < 0) { cand.push_back(cur * 2); h0 = false; }
        if (h1 && nd.d.y < 0) { cand.push_back(cur * 2 + 1); h1 = false; }
        if (h0 && h1) { stack[sp++] = nd.d.y; cur = nd.d.x; } else if (h0) cur = nd.d.x; else if (h1) cur = nd.d.y; else { if (!sp) break; cur = stack[--sp]; }
      }
      n_visits += visits; n_cand += cand.size(); maxc = std::max(maxc, (double)cand.size()); local_hist[std::min<size_t>(127, cand.size())]++;
      grp_sum += visits; grp_max = std::max(grp_max, (double)visits);
      if (++grp_n == 32) { eff_num += grp_sum; eff_den += grp_max * 32; grp_n = 0; grp_sum = 0; grp_max = 0; }
      double pmax = 0, psum = 0;
      for (const R& r : rays) {
        const float tlim = r.ts * 1.000001f; bool occ = false; int hits = 0;
        for (int c : cand) {
          const BvhNode& nd = nodes[c >> 1]; float tn; bool h; int ref;
          if (c & 1) { h = slab(r.ray, nd.b.z, nd.b.w, nd.c.x, nd.c.y, nd.c.z, nd.c.w, tlim, tn); ref = nd.d.y; }
          else { h = slab(r.ray, nd.a.x, nd.a.y, nd.a.z, nd.a.w, nd.b.x, nd.b.y, tlim, tn); ref = nd.d.x; }
          if (!h) continue;
          ++hits; n_leafhit += 1;
          const int f0 = leaf_first(ref), c0 = leaf_count(ref);
          for (int j = 0; j < c0 && !occ; ++j) { n_tri += 1; occ = tri_occludes(ttris.data(), f0 + j, r.ray, r.ts, r.prim); }
          if (occ) break;
        }
        psum += hits; pmax = std::max(pmax, (double)hits);
        const bool ref_occ = occluded(nodes.data(), ttris.data(), 0, r.ray, r.ts, r.prim);
        if (ref_occ != occ) n_mis += 1;
      }
      if (pmax > 0) { trip_num += psum; trip_den += pmax * rays.size(); }
    }
#pragma omp critical
    for (int i = 0; i < 128; ++i) hist[i] += local_hist[i];
  }
  out[0]=n_pairs; out[1]=n_rays; out[2]=n_visits/n_pairs; out[3]=n_cand/n_pairs; out[4]=n_leafhit/n_rays; out[5]=n_tri/n_rays; out[6]=n_mis;
  out[7]=eff_num/std::max(1.0,eff_den); out[8]=maxc; out[9]=n_rays/n_pairs; out[10]=trip_num/std::max(1.0,trip_den); out[11]=n_pairs/n_all_pairs;
  return 0;
}
}
