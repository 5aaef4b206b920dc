// bvh_quality.cpp — offline experiment (not a test, not product): traversal cost of the any-hit query under different
// BVH builders, using the device traversal code compiled for the host.
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
static float area(const B6& b) { float x = b.hi[0]-b.lo[0], y = b.hi[1]-b.lo[1], z = b.hi[2]-b.lo[2]; return 2*(x*y+y*z+z*x); }
static f3 ldv(const float* v, int i) { return mk3(v[3*i], v[3*i+1], v[3*i+2]); }

struct Built { std::vector<float4> ttris, stris; std::vector<BvhNode> nodes; int root_count = 0; };

// top-down binned SAH (16 bins), leaves <= leafmax, emits the device node format; order = sorted triangle order
struct SahBuilder {
  const std::vector<B6>& leaf; const std::vector<f3>& cen; std::vector<int> order; std::vector<BvhNode> nodes; int leafmax;
  SahBuilder(const std::vector<B6>& l, const std::vector<f3>& c, int lm) : leaf(l), cen(c), leafmax(lm) { order.resize(l.size()); std::iota(order.begin(), order.end(), 0); }
  B6 bounds(int a, int b) { B6 r = leaf[order[a]]; for (int i = a+1; i < b; ++i) r = uni(r, leaf[order[i]]); return r; }
  // returns (link,count,box)
  void build(int a, int b, int& link, int& cnt, B6& box) {
    box = bounds(a, b);
    if (b - a <= leafmax) { link = a; cnt = b - a; return; }
    float clo[3] = {3e38f,3e38f,3e38f}, chi[3] = {-3e38f,-3e38f,-3e38f};
    for (int i = a; i < b; ++i) { const f3& c = cen[order[i]]; float cc[3] = {c.x,c.y,c.z}; for (int k=0;k<3;++k){clo[k]=std::min(clo[k],cc[k]);chi[k]=std::max(chi[k],cc[k]);} }
    int bestAx = -1, bestSplit = -1; float bestCost = 3e38f; const int NB = 16;
    for (int ax = 0; ax < 3; ++ax) {
      float ext = chi[ax]-clo[ax]; if (ext <= 0) continue;
      B6 bb[NB]; int bc[NB] = {0}; bool init[NB] = {false};
      for (int i = a; i < b; ++i) { const f3& c = cen[order[i]]; float v = ax==0?c.x:(ax==1?c.y:c.z); int bi = std::min(NB-1, (int)((v-clo[ax])/ext*NB)); if (!init[bi]) { bb[bi] = leaf[order[i]]; init[bi]=true; } else bb[bi] = uni(bb[bi], leaf[order[i]]); bc[bi]++; }
      float la[NB], ra[NB]; int lc[NB], rc[NB]; B6 acc; bool has=false; int c=0;
      for (int i=0;i<NB;++i){ if(init[i]){acc = has?uni(acc,bb[i]):bb[i];has=true;} c+=bc[i]; la[i]=has?area(acc):0; lc[i]=c; }
      has=false;c=0; for (int i=NB-1;i>=0;--i){ if(init[i]){acc = has?uni(acc,bb[i]):bb[i];has=true;} c+=bc[i]; ra[i]=has?area(acc):0; rc[i]=c; }
      for (int i=0;i<NB-1;++i){ if(lc[i]==0||rc[i+1]==0) continue; float cost = la[i]*lc[i]+ra[i+1]*rc[i+1]; if (cost<bestCost){bestCost=cost;bestAx=ax;bestSplit=i;} }
    }
    int mid;
    if (bestAx < 0) mid = (a+b)/2;
    else { float ext = chi[bestAx]-clo[bestAx]; auto it = std::partition(order.begin()+a, order.begin()+b, [&](int t){ const f3& c = cen[t]; float v = bestAx==0?c.x:(bestAx==1?c.y:c.z); int bi = std::min(NB-1,(int)((v-clo[bestAx])/ext*NB)); return bi<=bestSplit;}); mid = (int)(it-order.begin()); if (mid==a||mid==b) mid=(a+b)/2; }
    int id = (int)nodes.size(); nodes.push_back(BvhNode());
    int l0,c0,l1,c1; B6 b0,b1; build(a, mid, l0,c0,b0); build(mid, b, l1,c1,b1);
    BvhNode n; n.a = make_float4(b0.lo[0],b0.lo[1],b0.lo[2],b0.hi[0]); n.b = make_float4(b0.hi[1],b0.hi[2],b1.lo[0],b1.lo[1]); n.c = make_float4(b1.lo[2],b1.hi[0],b1.hi[1],b1.hi[2]); n.d = make_int4(c0>0?leaf_ref(l0,c0):l0, c1>0?leaf_ref(l1,c1):l1, 0, 0);
    nodes[id] = n; link = id; cnt = 0;
  }
};

// ---- packet experiment: rays of one (32-triangle tile, 64-source chunk) are compacted in slot order (as k_forward's phase A does)
// and traversed 32 at a time with ONE shared stack: a node is entered when any live lane's ray hits its box.  Counts what a
// warp-packet kernel would execute, to compare with the per-lane traversal (lane-slots = node visits x 32 lanes).
static int g_packet = 0;          // 0: off, 1: row-major chunks of 64 sources, 2: 8x8 wall tiles (wall is WxW)
static int g_wall = 64;
extern "C" void bvh_quality_packet(int mode, int wall) { g_packet = mode; g_wall = wall; }
struct PRay { Ray ray; float ts; int prim; };
static void packet_trace(const std::vector<BvhNode>& nodes, const std::vector<float4>& ttris, const PRay* pr, int n,
                         unsigned long long& visits, unsigned long long& lane_box, unsigned long long& leaf_visits, unsigned long long& lane_tri,
                         unsigned long long& useful_box) {
  bool alive[32]; int nalive = n; for (int i = 0; i < n; ++i) alive[i] = true;
  int stack[256]; int sp = 0; int cur = 0;
  auto leaf = [&](int ref, const bool* hit) {
    ++leaf_visits; const int f0 = leaf_first(ref), c0 = leaf_count(ref);
    for (int i = 0; i < n; ++i) if (alive[i] && hit[i]) { lane_tri += c0;
      for (int j = 0; j < c0; ++j) if (tri_occludes(ttris.data(), f0 + j, pr[i].ray, pr[i].ts, pr[i].prim)) { alive[i] = false; --nalive; break; } }
  };
  while (nalive > 0) {
    const BvhNode& nd = nodes[cur]; ++visits; lane_box += 32;
    bool h0[32], h1[32]; bool any0 = false, any1 = false; float t0f = 0, t1f = 0; bool have = false;
    for (int i = 0; i < n; ++i) { h0[i] = h1[i] = false; if (!alive[i]) continue; const float tlim = pr[i].ts * 1.000001f; float t0, t1;
      h0[i] = slab(pr[i].ray, nd.a.x, nd.a.y, nd.a.z, nd.a.w, nd.b.x, nd.b.y, tlim, t0);
      h1[i] = slab(pr[i].ray, nd.b.z, nd.b.w, nd.c.x, nd.c.y, nd.c.z, nd.c.w, tlim, t1);
      if (h0[i] || h1[i]) ++useful_box;
      if (!have && (h0[i] || h1[i])) { have = true; t0f = h0[i] ? t0 : 3e38f; t1f = h1[i] ? t1 : 3e38f; }
      any0 |= h0[i]; any1 |= h1[i]; }
    int r0 = nd.d.x, r1 = nd.d.y;
    if (any0 && r0 < 0) { leaf(r0, h0); any0 = false; }
    if (any1 && r1 < 0) { leaf(r1, h1); any1 = false; }
    if (nalive == 0) break;
    if (any0 && any1) { const bool first0 = t0f <= t1f; stack[sp++] = first0 ? r1 : r0; cur = first0 ? r0 : r1; }
    else if (any0) cur = r0; else if (any1) cur = r1;
    else { if (sp == 0) break; cur = stack[--sp]; }
  }
}

extern "C" {
// mode 0: LBVH (Karras, leaf runs <= leafmax) ; mode 1: binned SAH.  Returns per-ray averages in out[0..3] = {rays, box/ray, tri/ray, nodes}
int bvh_quality(const float* origin, int L, const float* verts, int V, const int* faces, int F, int mode, int leafmax, float padscale, double* out, int* hist /*64 bins of node visits per ray*/) {
  float absmax = 0; for (int i = 0; i < 3*V; ++i) absmax = std::max(absmax, fabsf(verts[i])); for (int i = 0; i < 3*L; ++i) absmax = std::max(absmax, fabsf(origin[i]));
  const float pad = absmax * padscale;
  std::vector<B6> leaf(F); std::vector<f3> cen(F);
  for (int f = 0; f < F; ++f) { f3 a = ldv(verts, faces[3*f]), b = ldv(verts, faces[3*f+1]), c = ldv(verts, faces[3*f+2]);
    B6& bx = leaf[f]; bx.lo[0]=fminf(a.x,fminf(b.x,c.x))-pad; bx.lo[1]=fminf(a.y,fminf(b.y,c.y))-pad; bx.lo[2]=fminf(a.z,fminf(b.z,c.z))-pad; bx.hi[0]=fmaxf(a.x,fmaxf(b.x,c.x))+pad; bx.hi[1]=fmaxf(a.y,fmaxf(b.y,c.y))+pad; bx.hi[2]=fmaxf(a.z,fmaxf(b.z,c.z))+pad;
    cen[f] = mk3(0.5f*(bx.lo[0]+bx.hi[0]), 0.5f*(bx.lo[1]+bx.hi[1]), 0.5f*(bx.lo[2]+bx.hi[2])); }
  std::vector<int> order; std::vector<BvhNode> nodes; int root_count = F <= leafmax ? F : 0;
  if (mode == 1) { SahBuilder sb(leaf, cen, leafmax); int l,c; B6 b; sb.build(0, F, l, c, b); order = sb.order; nodes = sb.nodes; }
  else {
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
  // leafmax>4 cannot be encoded by child_ref (2 bits) -> only occluded() (count field) is used here
  unsigned long long nb=0, nt=0, nr=0; for (int i=0;i<64;++i) hist[i]=0;
  if (g_packet && root_count == 0) {
    unsigned long long visits = 0, lane_box = 0, leaf_visits = 0, lane_tri = 0, useful = 0, npk = 0, ind_box = 0, ind_tri = 0;
    const int W = g_wall; const int nchunk = L / 64;
    for (int ch = 0; ch < nchunk; ++ch) {
      int src[64];
      for (int q = 0; q < 64; ++q) { if (g_packet == 1) src[q] = ch * 64 + q; else { const int tx = ch % (W / 8), ty = ch / (W / 8); src[q] = (ty * 8 + q / 8) * W + tx * 8 + q % 8; } }
      for (int p0 = 0; p0 < F; p0 += 32) {
        std::vector<PRay> q;
        for (int qi = 0; qi < 64; ++qi) { const int s = src[qi]; if (s >= L) continue; f3 o = ldv(origin, s);
          for (int p = p0; p < std::min(F, p0 + 32); ++p) { ShadeTri st; TriRec tr; st.v1=xyz(stris[4*p]);st.A=stris[4*p].w;st.v2=xyz(stris[4*p+1]);st.v3=xyz(stris[4*p+2]);st.nf=mk3(stris[4*p+1].w,stris[4*p+2].w,stris[4*p+3].x);
            tr.v0=xyz(ttris[4*p]);tr.e1=xyz(ttris[4*p+1]);tr.e2=xyz(ttris[4*p+2]);tr.Ng=xyz(ttris[4*p+3]); int prim=f2i(ttris[4*p].w);
            SampleGeom g; if(!sample_self_hit(5489,s,prim,0,o,st,tr,g)) continue;
            float ff=-dot3(st.nf,g.d)*g.d.z; if(!(ff>0)) continue;
            PRay r; r.ray = make_ray(o, g.d); r.ts = g.t; r.prim = prim; q.push_back(r);
            uint32_t cb=0,ct=0; occluded(nodes.data(),ttris.data(),root_count,r.ray,g.t,prim,&cb,&ct); ind_box+=cb; ind_tri+=ct; } }
        for (size_t k = 0; k < q.size(); k += 32) { const int n = (int)std::min<size_t>(32, q.size() - k); packet_trace(nodes, ttris, q.data() + k, n, visits, lane_box, leaf_visits, lane_tri, useful); ++npk; nr += n; }
      }
    }
    out[0]=(double)nr; out[1]=(double)visits/npk; out[2]=(double)leaf_visits/npk; out[3]=(double)nr/npk;
    out[4]=(double)ind_box/2/nr; out[5]=(double)ind_tri/nr; out[6]=(double)useful/(double)(visits?visits:1); out[7]=(double)lane_tri/nr;
    return 0;
  }
  for (int s=0;s<L;++s){ f3 o=ldv(origin,s);
    for (int p=0;p<F;++p){ ShadeTri st; TriRec tr; st.v1=xyz(stris[4*p]);st.A=stris[4*p].w;st.v2=xyz(stris[4*p+1]);st.v3=xyz(stris[4*p+2]);st.nf=mk3(stris[4*p+1].w,stris[4*p+2].w,stris[4*p+3].x);
      tr.v0=xyz(ttris[4*p]);tr.e1=xyz(ttris[4*p+1]);tr.e2=xyz(ttris[4*p+2]);tr.Ng=xyz(ttris[4*p+3]); int prim=f2i(ttris[4*p].w);
      SampleGeom g; if(!sample_self_hit(5489,s,prim,0,o,st,tr,g)) continue;
      float ff=-dot3(st.nf,g.d)*g.d.z; if(!(ff>0)) continue;      // wall normal (0,0,1): only rays the forward kernel traces
      Ray ray=make_ray(o,g.d); uint32_t cb=0,ct=0; occluded(nodes.data(),ttris.data(),root_count,ray,g.t,prim,&cb,&ct); nb+=cb;nt+=ct;nr++; hist[std::min(63,(int)cb/2/2)]++; } }
  out[0]=(double)nr; out[1]=(double)nb/nr; out[2]=(double)nt/nr; out[3]=(double)nodes.size();
  return 0;
}
}
