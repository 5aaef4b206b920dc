// pgrid_emul.cpp — offline experiment / logic check (not product): the per-source perspective grid of the forward pass
// (nlos_core.cuh "per-source perspective grid"), run on the host with the DEVICE functions: pg_make_frame, pg_tri_rect,
// pg_entry, pg_ray, pg_precheck.  Counts entries per triangle, list length / pre-check passes / exact tests per ray and checks
// that the visibility answer equals the BVH any-hit query (== brute force) for every traced ray.
#include "emul_build.h"

extern "C" {
// out: [0] rays, [1] entries/triangle, [2] list length/ray, [3] pre-check passes/ray, [4] exact tests/ray (early exit), [5] mismatches,
//      [6] occluded fraction, [7] max list length, [8] warp max/mean list, [9] sources, [10] sources without grid
int pgrid_emul(const float* origin, const float* onormal, int L, const float* verts, int V, const int* faces, int F, int G,
               int src_stride, double* out) {
  EmulScene S; emul_build(origin, L, verts, V, faces, F, 4, 1.0f / 65536.0f, S);
  float blo[3], bhi[3]; for (int k = 0; k < 3; ++k) { blo[k] = S.vlo[k] - S.pad; bhi[k] = S.vhi[k] + S.pad; }
  const float zmin = pg_zmin(blo, bhi);
  const std::vector<float4>& ttris = S.ttris; const std::vector<float4>& stris = S.stris;
  double n_rays=0,n_entries=0,n_tris=0,n_list=0,n_pre=0,n_exact=0,n_mis=0,n_occ=0,maxl=0,w_num=0,w_den=0,nsrc=0,nogrid=0;
#pragma omp parallel for schedule(dynamic,1) reduction(+:n_rays,n_entries,n_tris,n_list,n_pre,n_exact,n_mis,n_occ,w_num,w_den,nsrc,nogrid) reduction(max:maxl)
  for (int s = 0; s < L; ++s) {
    if (src_stride > 1 && ((s / 64) % src_stride != 0 || (s % 64) % src_stride != 0)) continue;
    nsrc += 1;
    const f3 o = ldv(origin, s), n = ldv(onormal, s);
    PGridFrame g; bool ok; pg_init_frame(o, n, g, ok);
    if (!ok) { nogrid += 1; continue; }
    // pass 0: project the vertices (as the kernel does)
    std::vector<float2> proj(V); float U0=3e38f,U1=-3e38f,V0=3e38f,V1=-3e38f,mm=0; bool zok = true;
    for (int i = 0; i < V; ++i) { float u,v,m,z; pg_project(g, ldv(verts,i), u,v,m,z); zok = zok && z >= zmin; proj[i]=make_float2(u,v); U0=std::min(U0,u);U1=std::max(U1,u);V0=std::min(V0,v);V1=std::max(V1,v);mm=std::max(mm,m); }
    if (!zok || !(mm < 1.0f)) { nogrid += 1; continue; }
    const float pad_u = mm * (1.0f + std::max(fabsf(U0), fabsf(U1))), pad_v = mm * (1.0f + std::max(fabsf(V0), fabsf(V1)));
    pg_set_rect(g, U0 - pad_u, U1 + pad_u, V0 - pad_v, V1 + pad_v, G);
    if (g.G == 0) { nogrid += 1; continue; }
    std::vector<std::vector<uint2>> cells((size_t)G*G);
    for (int p = 0; p < F; ++p) { int a0,a1,b0,b1; const int f = S.order[p];
      const float2 p1 = proj[faces[3*f]], p2 = proj[faces[3*f+1]], p3 = proj[faces[3*f+2]];
      pg_tri_rect(g, p1.x,p1.y,p2.x,p2.y,p3.x,p3.y, pad_u, pad_v, a0,a1,b0,b1);
      for (int cy = b0 >> kPgSub; cy <= (b1 >> kPgSub); ++cy) for (int cx = a0 >> kPgSub; cx <= (a1 >> kPgSub); ++cx) { cells[(size_t)cy*G+cx].push_back(make_uint2(pg_entry(a0,a1,b0,b1,cx,cy), (unsigned)p)); n_entries += 1; } }
    n_tris += F;
    int wn = 0; double wsum = 0, wmax = 0;
    for (int p = 0; p < F; ++p) { ShadeTri st; TriRec tr; st.v1=xyz(stris[4*p]);st.A=stris[4*p].w;st.v2=xyz(stris[4*p+1]);st.v3=xyz(stris[4*p+2]);st.nf=mk3(stris[4*p+1].w,stris[4*p+2].w,stris[4*p+3].x);
      tr.v0=xyz(ttris[4*p]);tr.e1=xyz(ttris[4*p+1]);tr.e2=xyz(ttris[4*p+2]);tr.Ng=xyz(ttris[4*p+3]); int prim=f2i(ttris[4*p].w);
      SampleGeom sg; if(!sample_self_hit(5489,s,prim,0,o,st,tr,sg)) continue;
      const float dn = dot3(n,sg.d); float ff=-dot3(st.nf,sg.d)*dn; if(!(ff>0)) continue;
      int cx, cy; unsigned R; pg_ray(g, sg.d, cx, cy, R);
      const std::vector<uint2>& lst = cells[(size_t)cy*G+cx];
      bool occ=false; int pre=0, ex=0;
      for (const uint2& e : lst) { if (!pg_precheck(e.x, R) || (int)e.y == p) continue; ++pre; ++ex; if (tri_occludes_od(ttris.data(), (int)e.y, o, sg.d, sg.t, prim)) { occ=true; break; } }
      n_rays+=1; n_list+=lst.size(); n_pre+=pre; n_exact+=ex; n_occ+=occ; maxl=std::max(maxl,(double)lst.size());
      wsum += lst.size(); wmax = std::max(wmax,(double)lst.size()); if (++wn==32){ w_num+=wmax*32; w_den+=wsum; wn=0;wsum=0;wmax=0; }
      const bool ref = occluded(S.nodes.data(), ttris.data(), F <= 4 ? F : 0, make_ray(o, sg.d), sg.t, prim);
      if (ref != occ) n_mis += 1;
    }
  }
  out[0]=n_rays; out[1]=n_entries/std::max(1.0,n_tris); out[2]=n_list/std::max(1.0,n_rays); out[3]=n_pre/std::max(1.0,n_rays); out[4]=n_exact/std::max(1.0,n_rays); out[5]=n_mis; out[6]=n_occ/std::max(1.0,n_rays); out[7]=maxl; out[8]=w_num/std::max(1.0,w_den); out[9]=nsrc; out[10]=nogrid;
  return 0;
}
}
