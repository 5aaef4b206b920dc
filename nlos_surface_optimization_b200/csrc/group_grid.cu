// group_grid.cu — K1s, part 1: wall-point groups and the per-group binning of the mesh (DESIGN.md "K1s").
//
//   k_wall_keys     one block: bounds of the wall points, wall spacing h estimated from the bounds and the count (exact for a
//                   regular grid), key of every wall point = its tile of (side x h)^3
//   cub radix sort  (key, wall point) pairs
//   k_wall_groups   one block: a group = a run of equal keys, split every side^2 points
//   k_group_bin     one block per group (persistent over the batch): projects the vertices into the group's frame, counts and
//                   fills the (cell, slice) lists of the 3-D perspective grid (nlos_core.cuh "shared perspective grid of a GROUP")
//
// The reference has no counterpart: it rebuilds nothing per wall point because Embree walks one BVH for every ray
// (smoothed_transient/transient_and_gradient.cpp:196-206).  Here the per-ray tree walk is replaced by list scans, and the lists are
// shared by the wall points of a group so that building them costs a sixteenth of what a per-point grid costs.
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include "group_grid.h"

namespace nlos {

namespace {

constexpr int kBinBlock = 1024;

// block-wide inclusive scan of one value per thread (OP: 0 sum, 1 max); wsum: kBinBlock / 32 words of shared memory
template <int OP>
__device__ __forceinline__ int block_scan_incl(int x, int* wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x = OP ? max(x, y) : x + y; }
  __syncthreads();
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kBinBlock / 32 ? wsum[lane] : (OP ? -0x7fffffff : 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w = OP ? max(w, y) : w + y; }
    if (lane < kBinBlock / 32) wsum[lane] = w;
  }
  __syncthreads();
  if (warp > 0) x = OP ? max(x, wsum[warp - 1]) : x + wsum[warp - 1];
  return x;
}

__global__ void __launch_bounds__(kBinBlock) k_wall_keys(const float4* __restrict__ origin, int L, int side, uint64_t* __restrict__ keys, int* __restrict__ idx) {
  __shared__ unsigned lo[3], hi[3];
  __shared__ float s_tile, s_base[3];
  if (threadIdx.x < 3) { lo[threadIdx.x] = f2ord(3.0e38f); hi[threadIdx.x] = f2ord(-3.0e38f); }
  __syncthreads();
  float l[3] = {3.0e38f, 3.0e38f, 3.0e38f}, h[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = threadIdx.x; i < L; i += kBinBlock) {
    const float4 o = __ldg(origin + i);
    l[0] = fminf(l[0], o.x); h[0] = fmaxf(h[0], o.x); l[1] = fminf(l[1], o.y); h[1] = fmaxf(h[1], o.y); l[2] = fminf(l[2], o.z); h[2] = fmaxf(h[2], o.z);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o)); h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&lo[k], f2ord(l[k])); atomicMax(&hi[k], f2ord(h[k])); }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float e[3]; for (int k = 0; k < 3; ++k) { e[k] = ord2f(hi[k]) - ord2f(lo[k]); if (!(e[k] >= 0.0f) || !(e[k] < 3.0e30f)) e[k] = 0.0f; }
    // the two largest extents span the wall; n points spaced h apart on a (ex + h) x (ey + h) patch: (L - 1) h^2 - (ex + ey) h - ex ey = 0
    float ex = fmaxf(e[0], fmaxf(e[1], e[2])), ey = e[0] + e[1] + e[2] - ex - fminf(e[0], fminf(e[1], e[2]));
    float hh = 1.0f;
    if (L > 1 && ex > 0.0f) {
      const float n1 = (float)(L - 1), sum = ex + ey;
      hh = (sum + sqrtf(sum * sum + 4.0f * n1 * ex * ey)) / (2.0f * n1);
      if (!(hh > 0.0f)) hh = ex / n1;
    }
    s_tile = (float)side * hh;
    for (int k = 0; k < 3; ++k) s_base[k] = ord2f(lo[k]) - 0.5f * hh;
  }
  __syncthreads();
  const float inv = 1.0f / s_tile;
  for (int i = threadIdx.x; i < L; i += kBinBlock) {
    const float4 o = __ldg(origin + i);
    uint64_t key = 0;
    const float c[3] = {o.x, o.y, o.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float t = (c[k] - s_base[k]) * inv;
      t = fminf(fmaxf(t, 0.0f), 2097151.0f);                    // NaN -> 0
      key |= (uint64_t)(unsigned)(int)t << (21 * k);
    }
    keys[i] = key; idx[i] = i;
  }
}

// sorted keys -> groups: a run of equal keys, split every 'mmax' points.  start[g] = first sorted position of group g, start[n] = L, *n_groups = n
__global__ void __launch_bounds__(kBinBlock) k_wall_groups(const uint64_t* __restrict__ keys, int L, int mmax, int* __restrict__ group_of, int* __restrict__ start, int* __restrict__ n_groups) {
  __shared__ int wsum[kBinBlock / 32];
  const int per = (L + kBinBlock - 1) / kBinBlock;
  const int b = min(L, (int)threadIdx.x * per), e = min(L, b + per);
  int lastflag = -1;
  for (int i = b; i < e; ++i) if (i == 0 || keys[i] != keys[i - 1]) lastflag = i;
  const int incl = block_scan_incl<1>(lastflag, wsum);
  int carry = __shfl_up_sync(0xffffffffu, incl, 1);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) carry = threadIdx.x ? wsum[(threadIdx.x >> 5) - 1] : -1;
  // carry = last run start before this thread's chunk
  int runstart = carry, cnt = 0;
  for (int i = b; i < e; ++i) { if (i == 0 || keys[i] != keys[i - 1]) runstart = i; if ((i - runstart) % mmax == 0) ++cnt; }
  __syncthreads();
  const int gincl = block_scan_incl<0>(cnt, wsum);
  int g = gincl - cnt - 1;                                     // index of the group open at the start of the chunk
  runstart = carry;
  for (int i = b; i < e; ++i) {
    if (i == 0 || keys[i] != keys[i - 1]) runstart = i;
    if ((i - runstart) % mmax == 0) { ++g; start[g] = i; }
    group_of[i] = g;
  }
  if (threadIdx.x == kBinBlock - 1) { start[gincl] = L; *n_groups = gincl; }
}

struct BinShared {
  GGFrame fr;
  float zmin;
  float mda, mdb, mdn;
  int use_grid;
  unsigned rect[6];                          // ordered-uint min / max of u, v, w of the projected vertices
  unsigned maxm;
  int wsum[kBinBlock / 32];
  unsigned long long total, added;           // entries (padded) of the grid; entries added minus entries counted (non-zero: a 16-bit counter wrapped)
  float rad1;                                // largest L1 distance of a member from the centre
  f3 on;                                     // the members' common wall normal (as given, not normalised)
  int same_normal;
};

// One block per group.  table: (cursor, count) records, entries: per-group slices of 'cap' words.
__global__ void __launch_bounds__(kBinBlock, 1) k_group_bin(const DeviceScene sc, const float4* __restrict__ origin, const float4* __restrict__ onormal,
                                                            const int* __restrict__ order, const int* __restrict__ gstart, int g0, int ng,
                                                            GroupHdr* __restrict__ hdr, uint2* __restrict__ table, unsigned* __restrict__ ent_all,
                                                            float4* __restrict__ proj_all, int* __restrict__ live_all, int cull, unsigned cap, int G0, int K, unsigned tab_stride, int smem_bytes) {
  __shared__ BinShared bs;
  extern __shared__ __align__(16) unsigned cnt16[];          // packed 16-bit (cell, slice) counters
  const int tid = threadIdx.x, lane = tid & 31;
  float4* __restrict__ proj = proj_all + (size_t)blockIdx.x * sc.V;
  const int F = sc.F;
  const float pad = __int_as_float((int)sc.bounds->absmax) * (1.0f / 65536.0f);
  for (int slot = blockIdx.x; slot < ng; slot += gridDim.x) {
    const int g = g0 + slot;
    const int m0 = gstart[g], m1 = gstart[g + 1];
    uint2* __restrict__ tab = table + (size_t)slot * tab_stride;
    unsigned* __restrict__ ent = ent_all + 2 * (size_t)slot * cap;         // blocks of 4 entries: [E0 E1 E2 E3][T0 T1 T2 T3]
    int* __restrict__ live = live_all + (size_t)slot * F;
    // ---------------- frame: centre = mean of the members, axes around the first member's normal
    if (tid == 0) {
      f3 c = mk3(0.f, 0.f, 0.f);
      for (int i = m0; i < m1; ++i) c = c + xyz(__ldg(origin + order[i]));
      c = c * (1.0f / (float)(m1 - m0));
      bool ok; bs.fr.o = c; pg_make_axes(xyz(__ldg(onormal + order[m0])), bs.fr.a, bs.fr.b, bs.fr.n, ok);
      bs.fr.G = 0; bs.fr.K = K;
      float da = 0.f, db = 0.f, dn = 0.f;
      for (int i = m0; i < m1; ++i) {
        const f3 d = xyz(__ldg(origin + order[i])) - c;
        da = fmaxf(da, fabsf(dot3(d, bs.fr.a))); db = fmaxf(db, fabsf(dot3(d, bs.fr.b))); dn = fmaxf(dn, fabsf(dot3(d, bs.fr.n)));
      }
      bs.mda = da; bs.mdb = db; bs.mdn = dn;
      const f3 on0 = xyz(__ldg(onormal + order[m0]));
      float r1 = 0.f; int same = 1;
      for (int i = m0; i < m1; ++i) {
        const f3 d = xyz(__ldg(origin + order[i])) - c; r1 = fmaxf(r1, fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
        const f3 oni = xyz(__ldg(onormal + order[i])); if (!(oni.x == on0.x && oni.y == on0.y && oni.z == on0.z)) same = 0;
      }
      bs.rad1 = r1; bs.on = on0; bs.same_normal = same;
      float blo[3], bhi[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { blo[k] = ord2f(sc.bounds->vlo[k]) - pad; bhi[k] = ord2f(sc.bounds->vhi[k]) + pad; }
      bs.zmin = pg_zmin(blo, bhi);
      bs.use_grid = (ok && G0 > 0 && (da + db + dn) < 3.0e30f) ? 1 : 0;
      for (int k = 0; k < 3; ++k) { bs.rect[2 * k] = f2ord(3.0e38f); bs.rect[2 * k + 1] = f2ord(-3.0e38f); }
      bs.maxm = 0u;
    }
    __syncthreads();
    // ---------------- pass 0: project the vertices
    if (bs.use_grid) {
      const f3 o = bs.fr.o, a = bs.fr.a, b = bs.fr.b, n = bs.fr.n; const float zmin = bs.zmin;
      float U0 = 3.0e38f, U1 = -3.0e38f, V0 = 3.0e38f, V1 = -3.0e38f, W0 = 3.0e38f, W1 = -3.0e38f, mm = 0.f; bool zok = true;
      for (int i = tid; i < sc.V; i += kBinBlock) {
        const f3 x = mk3(__ldg(sc.verts + 3 * (size_t)i), __ldg(sc.verts + 3 * (size_t)i + 1), __ldg(sc.verts + 3 * (size_t)i + 2));
        float u, v, w, z, m; gg_project(o, a, b, n, x, u, v, w, z, m);
        zok = zok && (z >= zmin);
        proj[i] = make_float4(u, v, w, 0.f);
        U0 = fminf(U0, u); U1 = fmaxf(U1, u); V0 = fminf(V0, v); V1 = fmaxf(V1, v); W0 = fminf(W0, w); W1 = fmaxf(W1, w); mm = fmaxf(mm, m);
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        U0 = fminf(U0, __shfl_xor_sync(0xffffffffu, U0, d)); U1 = fmaxf(U1, __shfl_xor_sync(0xffffffffu, U1, d));
        V0 = fminf(V0, __shfl_xor_sync(0xffffffffu, V0, d)); V1 = fmaxf(V1, __shfl_xor_sync(0xffffffffu, V1, d));
        W0 = fminf(W0, __shfl_xor_sync(0xffffffffu, W0, d)); W1 = fmaxf(W1, __shfl_xor_sync(0xffffffffu, W1, d));
        mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, d));
      }
      zok = __all_sync(0xffffffffu, zok);
      if (lane == 0) {
        if (!zok || !(mm < 1.0f)) bs.use_grid = 0;
        atomicMin(&bs.rect[0], f2ord(U0)); atomicMax(&bs.rect[1], f2ord(U1)); atomicMin(&bs.rect[2], f2ord(V0)); atomicMax(&bs.rect[3], f2ord(V1));
        atomicMin(&bs.rect[4], f2ord(W0)); atomicMax(&bs.rect[5], f2ord(W1));
        atomicMax(&bs.maxm, (unsigned)__float_as_int(mm));
      }
    }
    __syncthreads();
    if (tid == 0 && bs.use_grid) {
      gg_finish_frame(bs.fr, ord2f(bs.rect[0]), ord2f(bs.rect[1]), ord2f(bs.rect[2]), ord2f(bs.rect[3]), ord2f(bs.rect[4]), ord2f(bs.rect[5]),
                      __int_as_float((int)bs.maxm), bs.mda, bs.mdb, bs.mdn, G0, K);
      if (bs.fr.G == 0) bs.use_grid = 0;
    }
    __syncthreads();
    // ---------------- live triangles: those the plane-side cull of the forward kernel does not remove for EVERY member.  With
    // w = v - o_m = (v - c) - delta_m:  n.w >= n.(v - c) - |n|_inf |delta_m|_1, and the member's margin is 1e-5 |w|_1 <= 1e-5 (|v - c|_1 + rad1);
    // four margins of slack dwarf the float error of either side.  Order (Morton) is kept: every thread owns a run of triangles.
    int nlive = F;
    {
      const int per = (F + kBinBlock - 1) / kBinBlock;
      const int b = min(F, tid * per), e = min(F, b + per);
      const bool do_cull = cull && bs.same_normal;
      const f3 c = bs.fr.o, on = bs.on; const float rad1 = bs.rad1;
      const float on_inf = fmaxf(fabsf(on.x), fmaxf(fabsf(on.y), fabsf(on.z)));
      int cnt = 0;
      unsigned long long deadbits = 0ull;                      // per <= 64 triangles of the run; longer runs are re-evaluated
      for (int p = b; p < e; ++p) {
        bool dead = false;
        if (do_cull) {
          const float4 s0 = __ldg(sc.stris + p), s1 = __ldg(sc.stris + 1 * (size_t)sc.F + p), s2 = __ldg(sc.stris + 2 * (size_t)sc.F + p), s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
          const f3 nf = mk3(s1.w, s2.w, s3.x);
          const f3 w1 = xyz(s0) - c, w2 = xyz(s1) - c, w3 = xyz(s2) - c;
          const float nf_inf = fmaxf(fabsf(nf.x), fmaxf(fabsf(nf.y), fabsf(nf.z)));
          const float l1 = fabsf(w1.x) + fabsf(w1.y) + fabsf(w1.z), l2 = fabsf(w2.x) + fabsf(w2.y) + fabsf(w2.z), l3 = fabsf(w3.x) + fabsf(w3.y) + fabsf(w3.z);
          dead = dot3(nf, w1) - nf_inf * rad1 > 4e-5f * (l1 + rad1) * (1.0f + nf_inf) &&
                 dot3(on, w1) - on_inf * rad1 > 4e-5f * (l1 + rad1) * (1.0f + on_inf) &&
                 dot3(on, w2) - on_inf * rad1 > 4e-5f * (l2 + rad1) * (1.0f + on_inf) &&
                 dot3(on, w3) - on_inf * rad1 > 4e-5f * (l3 + rad1) * (1.0f + on_inf);
        }
        if (p - b < 64 && dead) deadbits |= 1ull << (p - b);
        if (!dead) ++cnt;
      }
      __syncthreads();
      const int incl = block_scan_incl<0>(cnt, bs.wsum);
      int pos = incl - cnt;
      for (int p = b; p < e; ++p) {
        bool dead;
        if (p - b < 64) dead = (deadbits >> (p - b)) & 1ull;
        else {                                                                   // (runs longer than 64: F > 65536 — evaluate again)
          dead = false;
          if (do_cull) {
            const float4 s0 = __ldg(sc.stris + p), s1 = __ldg(sc.stris + 1 * (size_t)sc.F + p), s2 = __ldg(sc.stris + 2 * (size_t)sc.F + p), s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
            const f3 nf = mk3(s1.w, s2.w, s3.x);
            const f3 w1 = xyz(s0) - c, w2 = xyz(s1) - c, w3 = xyz(s2) - c;
            const float nf_inf = fmaxf(fabsf(nf.x), fmaxf(fabsf(nf.y), fabsf(nf.z)));
            const float l1 = fabsf(w1.x) + fabsf(w1.y) + fabsf(w1.z), l2 = fabsf(w2.x) + fabsf(w2.y) + fabsf(w2.z), l3 = fabsf(w3.x) + fabsf(w3.y) + fabsf(w3.z);
            dead = dot3(nf, w1) - nf_inf * rad1 > 4e-5f * (l1 + rad1) * (1.0f + nf_inf) &&
                   dot3(on, w1) - on_inf * rad1 > 4e-5f * (l1 + rad1) * (1.0f + on_inf) &&
                   dot3(on, w2) - on_inf * rad1 > 4e-5f * (l2 + rad1) * (1.0f + on_inf) &&
                   dot3(on, w3) - on_inf * rad1 > 4e-5f * (l3 + rad1) * (1.0f + on_inf);
          }
        }
        if (!dead) live[pos++] = p;
      }
      if (tid == kBinBlock - 1) bs.total = (unsigned long long)incl;
      __syncthreads();
      nlive = (int)bs.total;
      __syncthreads();
    }
    // ---------------- pass 1: count (coarsen until the entries fit), pass 2: fill.
    // The (cell, slice) counters live in shared memory as packed 16-bit pairs whenever G*G*K of them fit (one global atomic per entry
    // costs an L2 round trip each and bounded this kernel at ~1 ms per group); a count that could wrap (> 65535 entries in one list,
    // detected through the sum of the counts) and grids too large for shared memory take the global counters of the table instead.
    bool smem_counts = (size_t)bs.fr.G * bs.fr.G * K * 2 <= (size_t)smem_bytes;
    while (bs.use_grid) {
      const GGFrame fr = bs.fr;
      const int G = fr.G, ncell = G * G * K;
      if (smem_counts) { for (int i = tid; i < (ncell + 1) / 2; i += kBinBlock) cnt16[i] = 0u; }
      else { for (int i = tid; i < ncell; i += kBinBlock) tab[i] = make_uint2(0u, 0u); }
      if (tid == 0) { bs.total = 0ull; bs.added = 0ull; }
      __syncthreads();
      unsigned mine = 0u;
      for (int p = tid; p < F; p += kBinBlock) {
        const float4 s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
        const float4 p1 = proj[__float_as_int(s3.y)], p2 = proj[__float_as_int(s3.z)], p3 = proj[__float_as_int(s3.w)];
        int a0, a1, b0, b1, k0, k1; float wlo, whi;
        gg_tri_box(fr, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z, a0, a1, b0, b1, k0, k1, wlo, whi);
        const int cx0 = a0 >> kPgSub, cx1 = a1 >> kPgSub, cy0 = b0 >> kPgSub, cy1 = b1 >> kPgSub;
        mine += (unsigned)((cx1 - cx0 + 1) * (cy1 - cy0 + 1) * (k1 - k0 + 1));
        for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) for (int k = k0; k <= k1; ++k) {
          const int idx = (cy * G + cx) * K + k;
          if (smem_counts) atomicAdd(&cnt16[idx >> 1], 1u << (16 * (idx & 1))); else atomicAdd(&tab[idx].y, 1u);
        }
      }
      if (smem_counts) atomicAdd(&bs.added, (unsigned long long)mine);
      __syncthreads();
      // exclusive scan of the counts (each rounded up to a multiple of 4) -> first entry of every list
      unsigned long long total;
      {
        const int per = (ncell + kBinBlock - 1) / kBinBlock;
        const int b = min(ncell, tid * per), e = min(ncell, b + per);
        unsigned sum = 0, raw = 0;
        for (int i = b; i < e; ++i) { const unsigned c = smem_counts ? ((cnt16[i >> 1] >> (16 * (i & 1))) & 0xffffu) : tab[i].y; raw += c; sum += (c + 3u) & ~3u; }
        atomicAdd(&bs.total, (unsigned long long)sum);          // exact total in 64 bits; the 32-bit scan below is exact whenever total <= cap
        if (smem_counts) atomicAdd(&bs.added, 0ull - (unsigned long long)raw);
        const int incl = block_scan_incl<0>((int)sum, bs.wsum);
        unsigned base = (unsigned)incl - sum;
        for (int i = b; i < e; ++i) {
          const unsigned c = smem_counts ? ((cnt16[i >> 1] >> (16 * (i & 1))) & 0xffffu) : tab[i].y;
          tab[i] = make_uint2(base, c); base += (c + 3u) & ~3u;
        }
        __syncthreads();
        total = bs.total;
        if (smem_counts && bs.added != 0ull) { smem_counts = false; __syncthreads(); continue; }     // a 16-bit count wrapped: count again with the global counters
      }
      if (total <= (unsigned long long)cap) {
        if (smem_counts) { __syncthreads(); for (int i = tid; i < (ncell + 1) / 2; i += kBinBlock) cnt16[i] = 0u; __syncthreads(); }
        for (int p = tid; p < F; p += kBinBlock) {
          const float4 s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
          const float4 p1 = proj[__float_as_int(s3.y)], p2 = proj[__float_as_int(s3.z)], p3 = proj[__float_as_int(s3.w)];
          int a0, a1, b0, b1, k0, k1; float wlo, whi;
          gg_tri_box(fr, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z, a0, a1, b0, b1, k0, k1, wlo, whi);
          const int cx0 = a0 >> kPgSub, cx1 = a1 >> kPgSub, cy0 = b0 >> kPgSub, cy1 = b1 >> kPgSub;
          for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) {
            const unsigned E = pg_entry(a0, a1, b0, b1, cx, cy);
            for (int k = k0; k <= k1; ++k) {
              const int idx = (cy * G + cx) * K + k;
              unsigned pos;
              if (smem_counts) pos = tab[idx].x + ((atomicAdd(&cnt16[idx >> 1], 1u << (16 * (idx & 1))) >> (16 * (idx & 1))) & 0xffffu);
              else pos = atomicAdd(&tab[idx].x, 1u);
              ent[(size_t)(pos >> 2) * 8 + (pos & 3u)] = E; ent[(size_t)(pos >> 2) * 8 + 4 + (pos & 3u)] = (unsigned)p;
            }
          }
        }
        __syncthreads();
        // the tail of every list up to its group-of-4 boundary gets the never-matching word 0 (global counters: cursor -> first entry)
        for (int c = tid; c < ncell; c += kBinBlock) {
          uint2 t = tab[c];
          if (!smem_counts) { t.x -= t.y; tab[c] = t; }
          for (unsigned k = t.x + t.y; k < t.x + ((t.y + 3u) & ~3u); ++k) ent[(size_t)(k >> 2) * 8 + (k & 3u)] = 0u;
        }
        break;
      }
      if (G == 1) { if (tid == 0) bs.use_grid = 0; __syncthreads(); break; }       // cannot happen while cap >= F * K + 4 K
      if (tid == 0) gg_coarsen(bs.fr);
      __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
      GroupHdr h; h.fr = bs.fr; if (!bs.use_grid) h.fr.G = 0;
      h.member0 = m0; h.nmember = m1 - m0; h.ent0 = (unsigned long long)slot * cap; h.tab0 = (unsigned long long)slot * tab_stride;
      h.live0 = (unsigned long long)slot * (unsigned long long)F; h.nlive = nlive; h.pad_ = 0;
      hdr[slot] = h;
    }
    __syncthreads();
  }
}

}  // namespace

int make_wall_groups(Ctx& cx, const RenderParams& P, int side) {
  const int L = (int)P.L;
  cudaStream_t st = cx.stream;
  uint64_t* keys_in = cx.buf("wg_keys_in").as<uint64_t>(L);
  uint64_t* keys = cx.buf("wg_keys").as<uint64_t>(L);
  int* idx_in = cx.buf("wg_idx_in").as<int>(L);
  int* order = cx.buf("wg_order").as<int>(L);
  int* group_of = cx.buf("wg_group_of").as<int>(L);
  int* start = cx.buf("wg_start").as<int>((size_t)L + 1);
  int* n_dev = cx.buf("wg_count").as<int>(1);
  if (side < 1) side = 1;
  k_wall_keys<<<1, kBinBlock, 0, st>>>(P.origin, L, side, keys_in, idx_in);
  size_t tmp_bytes = 0;
  NLOS_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys, idx_in, order, L, 0, 63, st));
  void* tmp = cx.buf("wg_sort_tmp").ensure(tmp_bytes);
  NLOS_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys, idx_in, order, L, 0, 63, st));
  k_wall_groups<<<1, kBinBlock, 0, st>>>(keys, L, side * side, group_of, start, n_dev);
  cx.launches += 2;
  NLOS_CUDA_OK(cudaGetLastError());
  int n = 0;
  NLOS_CUDA_OK(cudaMemcpyAsync(&n, n_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  NLOS_CUDA_OK(cudaStreamSynchronize(st));
  return n;
}

void bin_wall_groups(Ctx& cx, const DeviceScene& sc, const RenderParams& P, int g0, int ng, int n_groups, int G, int K, unsigned cap, bool cull, GroupGrid& out) {
  (void)n_groups;
  const int sms = cx.num_sms > 0 ? cx.num_sms : 148;
  const int blocks = std::min(ng, sms);
  const unsigned tab_stride = (unsigned)G * (unsigned)G * (unsigned)K;
  GroupHdr* hdr = cx.buf("gg_hdr").as<GroupHdr>((size_t)ng);
  uint2* table = cx.buf("gg_table").as<uint2>((size_t)ng * tab_stride);
  unsigned* ent = cx.buf("gg_ent").as<unsigned>(2 * (size_t)ng * cap);
  float4* proj = cx.buf("gg_proj").as<float4>((size_t)blocks * sc.V);
  int* live = cx.buf("gg_live").as<int>((size_t)ng * sc.F);
  const int* order = cx.buf("wg_order").as<int>((size_t)P.L);
  const int* group_of = cx.buf("wg_group_of").as<int>((size_t)P.L);
  const int* start = cx.buf("wg_start").as<int>((size_t)P.L + 1);
  // shared memory for the packed 16-bit counters of the (cell, slice) lists, when they fit beside the static part
  int smem_bytes = (int)std::min<size_t>(((size_t)tab_stride * 2 + 15) & ~size_t(15), (size_t)200 * 1024);
  if ((size_t)tab_stride * 2 > (size_t)smem_bytes) smem_bytes = 0;
  NLOS_CUDA_OK(cudaFuncSetAttribute(k_group_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  k_group_bin<<<blocks, kBinBlock, smem_bytes, cx.stream>>>(sc, P.origin, P.onormal, order, start, g0, ng, hdr, table, ent, proj, live, cull ? 1 : 0, cap, G, K, tab_stride, smem_bytes);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
  out.hdr = hdr; out.order = order; out.group_of = group_of; out.table = table; out.ent = ent; out.live = live; out.group0 = g0;
}

}  // namespace nlos
