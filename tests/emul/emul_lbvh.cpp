// emul_lbvh.cpp — HOST logic check of the device building blocks in csrc/nlos_core.cuh (test-only).
// Builds the LBVH with the same per-node functions the kernels call (lbvh_range, morton30), emits the same
// 64-byte nodes as k_emit_nodes, then compares the any-hit traversal `occluded()` against a brute-force loop
// over all triangles for the renderer's own sample rays.  Never linked into libnlos_b200.so.
#include <algorithm>
#include <cstdio>
#include <vector>
#include <vector_types.h>
#include <vector_functions.h>
#include "../../nlos_surface_optimization_b200/csrc/nlos_core.cuh"

using namespace nlos;

namespace {
struct Built { std::vector<float4> ttris, stris; std::vector<BvhNode> nodes; int root_count; int F; };

struct B6 { float lo[3], hi[3]; };
static B6 uni(const B6& a, const B6& b) { B6 r; for (int k = 0; k < 3; ++k) { r.lo[k] = std::min(a.lo[k], b.lo[k]); r.hi[k] = std::max(a.hi[k], b.hi[k]); } return r; }

static f3 ldv(const float* v, int i) { return mk3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

static void build(const float* verts, int V, const int* faces, int F, const float* origin, int L, Built& out) {
  out.F = F;
  float absmax = 0; for (int i = 0; i < 3 * V; ++i) absmax = std::max(absmax, fabsf(verts[i]));
  for (int i = 0; i < 3 * L; ++i) absmax = std::max(absmax, fabsf(origin[i]));
  const float pad = absmax * (1.0f / 65536.0f);
  std::vector<f3> cen(F); float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
  for (int f = 0; f < F; ++f) {
    f3 a = ldv(verts, faces[3 * f]), b = ldv(verts, faces[3 * f + 1]), c = ldv(verts, faces[3 * f + 2]);
    cen[f] = mk3(0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x))), 0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y))),
                 0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z))));
    lo[0] = std::min(lo[0], cen[f].x); lo[1] = std::min(lo[1], cen[f].y); lo[2] = std::min(lo[2], cen[f].z);
    hi[0] = std::max(hi[0], cen[f].x); hi[1] = std::max(hi[1], cen[f].y); hi[2] = std::max(hi[2], cen[f].z);
  }
  std::vector<uint64_t> keys(F);
  for (int f = 0; f < F; ++f) {
    float ex = fmaxf(hi[0] - lo[0], 1e-30f), ey = fmaxf(hi[1] - lo[1], 1e-30f), ez = fmaxf(hi[2] - lo[2], 1e-30f);
    keys[f] = ((uint64_t)morton30((cen[f].x - lo[0]) / ex, (cen[f].y - lo[1]) / ey, (cen[f].z - lo[2]) / ez) << 32) | (uint32_t)f;
  }
  std::sort(keys.begin(), keys.end());
  out.ttris.resize(4 * (size_t)F); out.stris.resize(4 * (size_t)F);
  std::vector<B6> leaf(F);
  for (int p = 0; p < F; ++p) {
    int f = (int)(uint32_t)keys[p];
    int i1 = faces[3 * f], i2 = faces[3 * f + 1], i3 = faces[3 * f + 2];
    f3 v1 = ldv(verts, i1), v2 = ldv(verts, i2), v3 = ldv(verts, i3);
    TriRec tr = make_tri(v1, v2, v3);
    out.ttris[4 * p] = make_float4(tr.v0.x, tr.v0.y, tr.v0.z, i2f(f));
    out.ttris[4 * p + 1] = make_float4(tr.e1.x, tr.e1.y, tr.e1.z, 0);
    out.ttris[4 * p + 2] = make_float4(tr.e2.x, tr.e2.y, tr.e2.z, 0);
    out.ttris[4 * p + 3] = make_float4(tr.Ng.x, tr.Ng.y, tr.Ng.z, 0);
    f3 N = cross3(v2 - v1, v3 - v1); float A = len3(N) / 2; f3 nf = N / (2 * A);
    out.stris[4 * p] = make_float4(v1.x, v1.y, v1.z, A); out.stris[4 * p + 1] = make_float4(v2.x, v2.y, v2.z, nf.x);
    out.stris[4 * p + 2] = make_float4(v3.x, v3.y, v3.z, nf.y); out.stris[4 * p + 3] = make_float4(nf.z, i2f(i1), i2f(i2), i2f(i3));
    B6& b = leaf[p];
    b.lo[0] = fminf(v1.x, fminf(v2.x, v3.x)) - pad; b.lo[1] = fminf(v1.y, fminf(v2.y, v3.y)) - pad; b.lo[2] = fminf(v1.z, fminf(v2.z, v3.z)) - pad;
    b.hi[0] = fmaxf(v1.x, fmaxf(v2.x, v3.x)) + pad; b.hi[1] = fmaxf(v1.y, fmaxf(v2.y, v3.y)) + pad; b.hi[2] = fmaxf(v1.z, fmaxf(v2.z, v3.z)) + pad;
  }
  out.root_count = F <= kLeafMax ? F : 0;
  if (F < 2) { out.nodes.resize(1); return; }
  const int NI = F - 1;
  std::vector<int> first(NI), last(NI), cl(NI), cr(NI);
  for (int i = 0; i < NI; ++i) {
    int a, b, s; lbvh_range(keys.data(), F, i, a, b, s);
    first[i] = a; last[i] = b; cl[i] = (a == s) ? ~s : s; cr[i] = (b == s + 1) ? ~(s + 1) : s + 1;
  }
  // structural checks: every node except the root is referenced exactly once, ranges nest
  std::vector<int> refs(NI, 0), lrefs(F, 0);
  for (int i = 0; i < NI; ++i) { for (int c : {cl[i], cr[i]}) { if (c < 0) lrefs[~c]++; else refs[c]++; } }
  for (int i = 1; i < NI; ++i) if (refs[i] != 1) { fprintf(stderr, "node %d referenced %d times\n", i, refs[i]); out.F = -1; return; }
  for (int p = 0; p < F; ++p) if (lrefs[p] != 1) { fprintf(stderr, "leaf %d referenced %d times\n", p, lrefs[p]); out.F = -1; return; }
  if (refs[0] != 0 || first[0] != 0 || last[0] != F - 1) { fprintf(stderr, "bad root\n"); out.F = -1; return; }
  std::vector<B6> nb(NI);
  // bottom-up boxes by recursion over ranges (iterative post-order)
  std::vector<int> order; order.reserve(NI); std::vector<int> stack = {0};
  while (!stack.empty()) { int n = stack.back(); stack.pop_back(); order.push_back(n); if (cl[n] >= 0) stack.push_back(cl[n]); if (cr[n] >= 0) stack.push_back(cr[n]); }
  for (int k = (int)order.size() - 1; k >= 0; --k) {
    int n = order[k];
    B6 a = cl[n] < 0 ? leaf[~cl[n]] : nb[cl[n]], b = cr[n] < 0 ? leaf[~cr[n]] : nb[cr[n]];
    nb[n] = uni(a, b);
  }
  out.nodes.resize(NI);
  for (int i = 0; i < NI; ++i) {
    int link[2], cnt[2]; B6 bx[2]; int cc[2] = {cl[i], cr[i]};
    for (int k = 0; k < 2; ++k) {
      int c = cc[k];
      if (c < 0) { link[k] = ~c; cnt[k] = 1; bx[k] = leaf[~c]; }
      else { int size = last[c] - first[c] + 1; bx[k] = nb[c]; if (size <= kLeafMax) { link[k] = first[c]; cnt[k] = size; } else { link[k] = c; cnt[k] = 0; } }
    }
    BvhNode n;
    n.a = make_float4(bx[0].lo[0], bx[0].lo[1], bx[0].lo[2], bx[0].hi[0]);
    n.b = make_float4(bx[0].hi[1], bx[0].hi[2], bx[1].lo[0], bx[1].lo[1]);
    n.c = make_float4(bx[1].lo[2], bx[1].hi[0], bx[1].hi[1], bx[1].hi[2]);
    n.d = make_int4(cnt[0] > 0 ? leaf_ref(link[0], cnt[0]) : link[0], cnt[1] > 0 ? leaf_ref(link[1], cnt[1]) : link[1], 0, 0);
    out.nodes[i] = n;
  }
}
}  // namespace

extern "C" {
// returns #mismatches between BVH any-hit and brute force (or -1 on a structural error); vis_out[L,F,spp] in caller order
long emul_visibility(const float* origin, int L, const float* verts, int V, const int* faces, int F, int num_samples, unsigned long long seed,
                     unsigned char* vis_out, int check_brute, unsigned long long* counters) {
  Built b; build(verts, V, faces, F, origin, L, b);
  if (b.F < 0) return -1;
  const int spp = std::max(1, 1 + (num_samples - 1) / F);
  long mism = 0; unsigned long long nb = 0, nt = 0, nr = 0;
  for (int s = 0; s < L; ++s) {
    f3 o = ldv(origin, s);
    for (int p = 0; p < F; ++p) {
      ShadeTri st; TriRec tr;
      const float4 s0 = b.stris[4 * p], s1 = b.stris[4 * p + 1], s2 = b.stris[4 * p + 2], s3 = b.stris[4 * p + 3];
      st.v1 = xyz(s0); st.A = s0.w; st.v2 = xyz(s1); st.v3 = xyz(s2); st.nf = mk3(s1.w, s2.w, s3.x);
      tr.v0 = xyz(b.ttris[4 * p]); tr.e1 = xyz(b.ttris[4 * p + 1]); tr.e2 = xyz(b.ttris[4 * p + 2]); tr.Ng = xyz(b.ttris[4 * p + 3]);
      const int prim = f2i(b.ttris[4 * p].w);
      for (int k = 0; k < spp; ++k) {
        SampleGeom g; unsigned char bit = 0;
        if (sample_self_hit(seed, s, prim, k, o, st, tr, g)) {
          Ray ray = make_ray(o, g.d);
          uint32_t cb = 0, ct = 0;
          bool occ = occluded(b.nodes.data(), b.ttris.data(), b.root_count, ray, g.t, prim, &cb, &ct);
          if (occluded_ww(b.nodes.data(), b.ttris.data(), b.root_count, ray, g.t, prim) != occ) ++mism;   // both traversal orders agree
          nb += cb; nt += ct; nr++;
          bit = occ ? 0 : 1;
          if (check_brute) {
            bool occ2 = false;
            for (int j = 0; j < F && !occ2; ++j) { occ2 = tri_occludes(b.ttris.data(), j, ray, g.t, prim); if (tri_occludes_fast(b.ttris.data(), j, ray, g.t, prim) != occ2) ++mism; }
            if (occ2 != occ) ++mism;
          }
        }
        vis_out[((size_t)s * F + prim) * spp + k] = bit;
      }
    }
  }
  if (counters) { counters[0] = nr; counters[1] = nb; counters[2] = nt; }
  return mism;
}
}
