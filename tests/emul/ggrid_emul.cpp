// ggrid_emul.cpp — offline experiment / logic check (not product): the shared perspective grid of a GROUP of wall points
// (nlos_core.cuh "shared perspective grid of a GROUP"), run on the host with the DEVICE functions gg_finish_frame, gg_tri_box,
// gg_tri_edges, gg_ray_setup, gg_ray_slice, pg_precheck, gg_edges_pass.  Counts entries per triangle, lookups / scanned words /
// rectangle passes / edge passes / exact tests per ray and checks that the visibility answer equals the BVH any-hit query for
// every traced ray of the sampled groups.  The wall is taken as a W x W row-major grid; a group is a gx x gy tile of it.
#include "emul_build.h"

static int run(const float* origin, const float* onormal, int L, const float* verts, int V, const int* faces, int F, int G, int K,
               int W, int gx, int gy, int group_stride, double* out) {
  EmulScene S; emul_build(origin, L, verts, V, faces, F, 4, 1.0f / 65536.0f, S);
  float blo[3], bhi[3]; for (int k = 0; k < 3; ++k) { blo[k] = S.vlo[k] - S.pad; bhi[k] = S.vhi[k] + S.pad; }
  const float zmin = pg_zmin(blo, bhi);
  const std::vector<float4>& ttris = S.ttris; const std::vector<float4>& stris = S.stris;
  const int H = L / W, ngx = (W + gx - 1) / gx, ngy = (H + gy - 1) / gy;
  double n_rays=0,n_entries=0,n_tris=0,n_look=0,n_look_ne=0,n_scan=0,n_rect=0,n_edge=0,n_exact=0,n_mis=0,n_occ=0,n_fallback=0,ngroups=0,nogrid=0,n_never=0,max_e=0;
#pragma omp parallel for schedule(dynamic,1) reduction(+:n_rays,n_entries,n_tris,n_look,n_look_ne,n_scan,n_rect,n_edge,n_exact,n_mis,n_occ,n_fallback,ngroups,nogrid,n_never) reduction(max:max_e)
  for (int gi = 0; gi < ngx * ngy; ++gi) {
    if (group_stride > 1 && gi % group_stride != 0) continue;
    const int tx = gi % ngx, ty = gi / ngx;
    std::vector<int> members;
    for (int j = ty * gy; j < std::min(H, (ty + 1) * gy); ++j) for (int i = tx * gx; i < std::min(W, (tx + 1) * gx); ++i) members.push_back(j * W + i);
    ngroups += 1;
    // frame: centre = mean origin, normal of the first member
    f3 c = mk3(0, 0, 0); for (int s : members) c = c + ldv(origin, s); c = c * (1.0f / (float)members.size());
    GGFrame g; bool ok; g.o = c; pg_make_axes(ldv(onormal, members[0]), g.a, g.b, g.n, ok); g.G = 0;
    if (!ok) { nogrid += 1; continue; }
    float mda = 0, mdb = 0, mdn = 0;
    for (int s : members) { const f3 d = ldv(origin, s) - c; mda = std::max(mda, fabsf(dot3(d, g.a))); mdb = std::max(mdb, fabsf(dot3(d, g.b))); mdn = std::max(mdn, fabsf(dot3(d, g.n))); }
    std::vector<float4> proj(V); float U0=3e38f,U1=-3e38f,V0=3e38f,V1=-3e38f,W0=3e38f,W1=-3e38f,mm=0; bool zok = true;
    for (int i = 0; i < V; ++i) { float u,v,w,z,m; gg_project(g.o,g.a,g.b,g.n, ldv(verts,i), u,v,w,z,m); zok = zok && z >= zmin; proj[i]=make_float4(u,v,w,0);
      U0=std::min(U0,u);U1=std::max(U1,u);V0=std::min(V0,v);V1=std::max(V1,v);W0=std::min(W0,w);W1=std::max(W1,w);mm=std::max(mm,m); }
    if (!zok) { nogrid += 1; continue; }
    gg_finish_frame(g, U0, U1, V0, V1, W0, W1, mm, mda, mdb, mdn, G, K);
    if (g.G == 0) { nogrid += 1; continue; }
    max_e = std::max(max_e, (double)std::max(g.eu * g.su, g.ev * g.sv));
    struct Ent { unsigned E, e0, e1, e2, tri; };
    std::vector<std::vector<Ent>> cells((size_t)G * G * K);
    for (int p = 0; p < F; ++p) { const int f = S.order[p];
      const float4 p1 = proj[faces[3*f]], p2 = proj[faces[3*f+1]], p3 = proj[faces[3*f+2]];
      int a0,a1,b0,b1,k0,k1; float wlo, whi; gg_tri_box(g, p1.x,p1.y,p1.z, p2.x,p2.y,p2.z, p3.x,p3.y,p3.z, a0,a1,b0,b1,k0,k1,wlo,whi);
      float hmax = 0; int jm[64]; for (int k = k0; k <= k1; ++k) { float h; jm[k] = gg_tri_fine(g, wlo, whi, k0, k1, k, h); hmax = std::max(hmax, h); }
      for (int cy = b0 >> kPgSub; cy <= (b1 >> kPgSub); ++cy) for (int cx = a0 >> kPgSub; cx <= (a1 >> kPgSub); ++cx) {
        Ent e; e.E = pg_entry(a0,a1,b0,b1,cx,cy);
        if (!gg_tri_edges(g, p1.x,p1.y,p2.x,p2.y,p3.x,p3.y, cx, cy, hmax, e.e0, e.e1, e.e2)) { e.E = 0u; n_never += (k1 - k0 + 1); }
        for (int k = k0; k <= k1; ++k) { e.tri = (unsigned)p | ((unsigned)jm[k] << 27); cells[((size_t)cy*G+cx)*K + k].push_back(e); n_entries += 1; } } }
    n_tris += F;
    for (int s : members) {
      const f3 o = ldv(origin, s), n = ldv(onormal, s);
      const f3 del = o - c; const float da = dot3(del, g.a), db = dot3(del, g.b), dn = dot3(del, g.n);
      for (int p = 0; p < F; ++p) { ShadeTri st; TriRec tr; st.v1=xyz(stris[4*p]);st.A=stris[4*p].w;st.v2=xyz(stris[4*p+1]);st.v3=xyz(stris[4*p+2]);st.nf=mk3(stris[4*p+1].w,stris[4*p+2].w,stris[4*p+3].x);
        tr.v0=xyz(ttris[4*p]);tr.e1=xyz(ttris[4*p+1]);tr.e2=xyz(ttris[4*p+2]);tr.Ng=xyz(ttris[4*p+3]); int prim=f2i(ttris[4*p].w);
        SampleGeom sg; if(!sample_self_hit(5489,s,prim,0,o,st,tr,sg)) continue;
        const float dnn = dot3(n,sg.d); float ff=-dot3(st.nf,sg.d)*dnn; if(!(ff>0)) continue;
        n_rays += 1;
        const bool ref = occluded(S.nodes.data(), ttris.data(), F <= 4 ? F : 0, make_ray(o, sg.d), sg.t, prim);
        GGRay r; bool occ = false;
        if (!gg_ray_setup(g, da, db, dn, sg.d, sg.t, r)) { n_fallback += 1; occ = ref; }
        else {
          const GGWalk wk = gg_ray_walk(g, r);
          for (int k = 0; k <= r.kr && !occ; ++k) {
            int uq, vq; gg_walk_point(g, wk, (float)k, uq, vq);
            struct { unsigned R; int cx, cy; } q; q.R = gg_rect_word(uq, vq); q.cx = uq >> kPgSub; q.cy = vq >> kPgSub;
            const GGFinePoint fp = gg_walk_fine(wk, (float)k, uq, vq);
            const std::vector<Ent>& lst = cells[((size_t)q.cy*G+q.cx)*K + k];
            n_look += 1; if (!lst.empty()) n_look_ne += 1; n_scan += (double)((lst.size() + 3) & ~(size_t)3);
            for (const Ent& e : lst) {
              if (!pg_precheck(e.E, q.R)) continue; n_rect += 1;
              unsigned rw; const bool inr = gg_ray_word(fp, (int)(e.tri >> 27), rw);
              if (inr && !gg_edges_pass(e.e0, e.e1, e.e2, rw)) continue; n_edge += 1;
              const int tj = (int)(e.tri & 0x7ffffffu);
              if (tj == p) continue; n_exact += 1;
              if (tri_occludes_od(ttris.data(), tj, o, sg.d, sg.t, prim)) { occ = true; break; }
            }
          }
        }
        n_occ += occ;
        if (ref != occ) n_mis += 1;
      }
    }
  }
  const double nr = std::max(1.0, n_rays);
  out[0]=n_rays; out[1]=n_entries/std::max(1.0,n_tris); out[2]=n_look/nr; out[3]=n_look_ne/nr; out[4]=n_scan/nr; out[5]=n_rect/nr; out[6]=n_edge/nr; out[7]=n_exact/nr;
  out[8]=n_mis; out[9]=n_occ/nr; out[10]=n_fallback; out[11]=ngroups; out[12]=nogrid; out[13]=n_never/std::max(1.0,n_entries); out[14]=max_e;
  return 0;
}

extern "C" int ggrid_emul(const float* origin, const float* onormal, int L, const float* verts, int V, const int* faces, int F, int G, int K,
                          int W, int gx, int gy, int group_stride, double* out) {
  if (K < 1 || K > 64) return -1;
  return run(origin, onormal, L, verts, V, faces, F, G, K, W, gx, gy, group_stride, out);
}
