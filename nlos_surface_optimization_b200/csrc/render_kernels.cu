// render_kernels.cu — K1 (forward), K3 (residual), K4 (vertex gradient), K5 (scalar gradients),
// K6 (per-triangle intensity) and the debug visibility kernel.
//
// Work mapping (DESIGN.md "Kernel mapping"): one THREAD owns one triangle (Morton order, so a warp is 32
// spatially adjacent triangles) and loops over a chunk of relay-wall sources; blockIdx.y selects the chunk.
//   * rays of a warp share their origin and hit a compact patch  -> coherent BVH traversal
//   * triangle data stays in registers for the whole source loop -> no re-gather per (source,triangle) task
//   * the vertex gradient of a triangle accumulates in 9 FP64 registers across the source loop and is
//     flushed with 9 atomics per (triangle, chunk) instead of 9*K atomics per visible sample
//   * visibility is an ANY-HIT query bounded by the self-hit distance (equivalent to the reference's
//     "nearest hit == sampled triangle", TG.cpp:206) and is skipped for samples whose contribution is
//     exactly zero (back-facing / out of range)
//   * the forward pass leaves one visibility bit per sample (shared-memory tile per warp) that the gradient pass
//     reuses, so the gradient pass traces no rays (legitimate: both passes use identical samples)
//   * forward kernel: the four warps of a block own the SAME triangle tile and take different slot chunks (L1 locality);
//     generated samples are compacted into a per-warp ray queue and lanes are refilled from it while others still traverse;
//     histogram updates of finished rays are batched at the refills (see k_forward)
//
// Reference: smoothed_transient/transient_and_gradient.cpp:122-237 (forward task), :843-1007 (gradient task),
// :571-695 (albedo), :22-119 (intensity); ggx/transient_and_gradient.cpp:126-243, 385-512, 648-823.
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include "nlos_ctx.h"
#include "render_kernels.h"
#include "group_grid.h"

namespace nlos {
#ifdef NLOS_EXT_BUILD      // second build of this file (render_kernels_ext.cu): the same kernels with the external-sample test hook compiled in
namespace ext {
#endif

namespace {

constexpr int kBlock = 128;
#ifndef NLOS_FWD_BLOCK
#define NLOS_FWD_BLOCK 128
#endif
constexpr int kFwdBlock = NLOS_FWD_BLOCK;   // threads per block of the forward kernel (all its warps share one triangle tile)

struct TriRegs {
  ShadeTri st; TriRec tr; int prim;
  f3 n1, n2, n3; float a1, a2, a3;
};

template <bool HAS_VN, bool HAS_VA>
__device__ __forceinline__ void load_tri(const DeviceScene& sc, int p, TriRegs& t) {
  const float4 s0 = __ldg(sc.stris + p), s1 = __ldg(sc.stris + 1 * (size_t)sc.F + p), s2 = __ldg(sc.stris + 2 * (size_t)sc.F + p), s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
  t.st.v1 = xyz(s0); t.st.A = s0.w; t.st.v2 = xyz(s1); t.st.v3 = xyz(s2); t.st.nf = mk3(s1.w, s2.w, s3.x);
  t.st.i1 = __float_as_int(s3.y); t.st.i2 = __float_as_int(s3.z); t.st.i3 = __float_as_int(s3.w);
  const float4 q0 = __ldg(sc.ttris + 4 * (size_t)p), q1 = __ldg(sc.ttris + 4 * (size_t)p + 1), q2 = __ldg(sc.ttris + 4 * (size_t)p + 2), q3 = __ldg(sc.ttris + 4 * (size_t)p + 3);
  t.tr.v0 = xyz(q0); t.tr.e1 = xyz(q1); t.tr.e2 = xyz(q2); t.tr.Ng = xyz(q3); t.prim = __ldg(sc.sprim + p);     // (not q0.w: a kernel that does not use t.tr then loads no trace record at all)
  if (HAS_VN) {
    const float* vn = sc.vnormal;
    t.n1 = mk3(__ldg(vn + 3 * (size_t)t.st.i1), __ldg(vn + 3 * (size_t)t.st.i1 + 1), __ldg(vn + 3 * (size_t)t.st.i1 + 2));
    t.n2 = mk3(__ldg(vn + 3 * (size_t)t.st.i2), __ldg(vn + 3 * (size_t)t.st.i2 + 1), __ldg(vn + 3 * (size_t)t.st.i2 + 2));
    t.n3 = mk3(__ldg(vn + 3 * (size_t)t.st.i3), __ldg(vn + 3 * (size_t)t.st.i3 + 1), __ldg(vn + 3 * (size_t)t.st.i3 + 2));
  }
  if (HAS_VA) { t.a1 = __ldg(sc.valbedo + t.st.i1); t.a2 = __ldg(sc.valbedo + t.st.i2); t.a3 = __ldg(sc.valbedo + t.st.i3); }
}

template <bool HAS_VN>
__device__ __forceinline__ f3 shading_normal(const TriRegs& t, const SampleGeom& g) { return HAS_VN ? blend3(g.u, t.n1, g.v, t.n2, g.w, t.n3) : t.st.nf; }
template <bool HAS_VA>
__device__ __forceinline__ float shading_albedo(const TriRegs& t, const SampleGeom& g) { return HAS_VA ? blend1(g.u, t.a1, g.v, t.a2, g.w, t.a3) : 1.0f; }

// The two uniforms of sample k of (global source s, triangle prim): Philox.  In the SECOND build of this file (NLOS_EXT_BUILD,
// namespace nlos::ext, selected by run_job only while nlos_ctx_set_external_samples is active) they come from the external stream
// of the test hook instead — compiled separately because even a never-taken branch costs the production kernel 3-4 % (registers).
__device__ __forceinline__ bool draw_sample(const RenderParams& P, int64_t src_global, int prim, int k, f3 o, const ShadeTri& st, const TriRec& tr, SampleGeom& g) {
#ifdef NLOS_EXT_BUILD
  const int64_t idx = P.ext_base + 2 * ((src_global * (int64_t)P.F + prim) * P.spp + k);
  if (P.ext_samples == nullptr || idx < 0 || idx + 1 >= P.ext_count) return false;
  return sample_self_hit_st(__ldg(P.ext_samples + idx), __ldg(P.ext_samples + idx + 1), o, st, tr, g);
#else
  return sample_self_hit(P.seed, src_global, prim, k, o, st, tr, g);
#endif
}

// ---------------------------------------------------------------------------------------------- K1 forward
// Warp-level two-phase kernel (DESIGN.md "Forward kernel"):
//   phase A (generate): lane <-> triangle.  Every lane draws the sample of the next slot (slot = source*spp + k), runs
//     the self intersection and the shading that needs no visibility; samples that can contribute (front-facing, in
//     range) are COMPACTED into a per-warp ray queue in shared memory (ballot + popc).
//   phase B (trace): idle lanes pop rays — of any triangle of the tile — from the queue (refilled whenever kRefill lanes are
//     idle) and run the any-hit traversal; visible rays add their value to the transient row (FP64 RED) and set their bit in
//     the warp's visibility tile.
// Back-facing / out-of-range samples therefore never occupy a lane during traversal, which is where the time goes.
constexpr int kQCap = 64;        // ray queue capacity per warp (power of two, >= 63)
#ifndef NLOS_TILE
#define NLOS_TILE 256
#endif
constexpr int kTile = NLOS_TILE; // sample slots per warp pass (upper bound of the chunk_forward option) == visibility words in the warp tile
#ifndef NLOS_REFILL
#define NLOS_REFILL 24
#endif
#ifndef NLOS_MINLANES
#define NLOS_MINLANES 6
#endif
#ifndef NLOS_FWD_MINBLOCKS
#define NLOS_FWD_MINBLOCKS 6
#endif
constexpr int kRefill = NLOS_REFILL;   // idle lanes that trigger a refill of the traversal lanes from the queue
constexpr int kDone = (int)0x80000000;   // traversal state: stack exhausted, no occluder found (never a valid leaf ref)
struct WarpShared {
  float dx[kQCap], dy[kQCap], dz[kQCap], ts[kQCap], val[kQCap];
  int bin[kQCap];                // histogram bin of the sample (-1: outside the histogram)
  int meta[kQCap];               // slot_local | tri_lane << 16
  unsigned tile[kTile];          // one visibility word (32 triangles) per slot
};
// (Per-lane stacks stay in local memory: moving them to shared memory [level][lane] was measured 33 % SLOWER — the 32 KB
//  per block it costs shrink the L1 that keeps the 13.5 MB of nodes + triangles hot.)

// MODE 0: transient histogram (+ optional visibility bits);  MODE 1: per-triangle intensity (K6)
template <bool GGX, bool HAS_VN, bool HAS_VA, bool SMOOTH, bool WRITE_VIS, int MODE>
__global__ void __launch_bounds__(kFwdBlock, NLOS_FWD_MINBLOCKS) k_forward(const DeviceScene sc, const RenderParams P, double* __restrict__ out,
                                                    uint32_t* __restrict__ vis, const double* __restrict__ wprefix) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpShared* ws_all = reinterpret_cast<WarpShared*>(smem_raw);
  double* s_w = reinterpret_cast<double*>(smem_raw + (kFwdBlock / 32) * sizeof(WarpShared));   // SMOOTH: tap prefix sums
  if (SMOOTH) { for (int i = threadIdx.x; i <= P.K; i += blockDim.x) s_w[i] = wprefix[i]; __syncthreads(); }
  const int lane = threadIdx.x & 31;
  WarpShared& ws = ws_all[threadIdx.x >> 5];
  // all warps of a block share ONE triangle tile and split the slot chunks among them: they walk the same BVH region at the
  // same time, which keeps it in L1 (measured 26.8 vs 31.4 ms with one tile per warp)
  const int p = blockIdx.x * 32 + lane;
  const bool active = p < sc.F;
  const int warp_global = p >> 5;
  TriRegs t;
  t.prim = 0;
  if (active) load_tri<HAS_VN, HAS_VA>(sc, p, t);
  const float ub_half = P.ub / 2.0f, lb_half = P.lb / 2.0f;
  const int64_t nbf = (int64_t)P.numBins * P.r_fwd;
  const int64_t total_slots = P.L * (int64_t)P.spp;
  const int64_t nchunks = (total_slots + P.chunk - 1) / P.chunk;
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t chunk = (int64_t)blockIdx.y * (kFwdBlock / 32) + (threadIdx.x >> 5); chunk < nchunks; chunk += (int64_t)gridDim.y * (kFwdBlock / 32)) {
    const int64_t slot0 = chunk * P.chunk;
    const int nslots = (int)(total_slots - slot0 < P.chunk ? total_slots - slot0 : P.chunk);
    if (WRITE_VIS) { for (int i = lane; i < nslots; i += 32) ws.tile[i] = 0u; }
    __syncwarp();
    int gen = 0, qhead = 0, qcount = 0;
    // per-lane traversal state of the ray in flight (cur == kSentinel: lane is idle)
    int cur = kSentinel, sp = 0, stack[kStack];
    Ray ray; float ts = 0.f, val = 0.f; int bin = -1, prim = 0, tri_lane = 0, slot_local = 0; int64_t src = 0;
    bool pending = false;                                        // this lane's last ray ended visible and its contribution is not added yet
    ray.o = ray.d = ray.id = ray.oid = mk3(0.f, 0.f, 0.f);
    for (;;) {
      const unsigned idle_mask = __ballot_sync(0xffffffffu, cur == kSentinel);
      const int n_idle = __popc(idle_mask);
      if (n_idle >= kRefill) {
        if (pending) {                                             // rays that ended visible since the last refill
          const double dv = P.spp == 1 ? (double)val : (double)val / (double)P.spp;   // TG.cpp:231-232 (x / 1.0 == x: skips an FP64 division)
          if (MODE == 1) atomicAdd(out + prim, dv);
          else if (bin >= 0) {
            if (!SMOOTH) atomicAdd(out + src * P.numBins + bin, dv);
            else {
              // Gaussian smoothing + downsampling of TG.cpp:348-371 applied per sample: tap i of fine bin m lands in
              // coarse bin floor((m + i - 2rs)/r); the taps of one coarse bin are a contiguous run -> prefix sums
              const int half = 2 * P.r_fwd * P.s_bin;
              const int q0 = floordiv32(bin, P.r_fwd); int b0 = q0 - 2 * P.s_bin, b1 = q0 + 2 * P.s_bin;      // half = 2 r s is a multiple of r: one division (bin < numBins * r_fwd < 2^31)
              if (b0 < 0) b0 = 0; if (b1 > P.numBins - 1) b1 = P.numBins - 1;
              for (int b = b0; b <= b1; ++b) {
                int ilo, ihi; tap_span32(bin, b, P.r_fwd, half, P.K, ilo, ihi);
                if (ihi > ilo) atomicAdd(out + src * P.numBins + b, dv * (s_w[ihi] - s_w[ilo]));
              }
            }
          }
          if (WRITE_VIS) atomicOr(&ws.tile[slot_local], 1u << tri_lane);
          pending = false;
        }
        // ---------------- phase A: generate + compact until the idle lanes can be fed
        while (qcount < n_idle && gen < nslots) {
          const int64_t slot = slot0 + gen;
          const int64_t s = P.spp == 1 ? slot : slot / P.spp;
          const int k = P.spp == 1 ? 0 : (int)(slot - s * P.spp);
          bool need = false;
          float r_dx = 0.f, r_dy = 0.f, r_dz = 0.f, r_ts = 0.f, r_val = 0.f; int r_bin = -1;
          if (active) {
            const f3 o = xyz(__ldg(P.origin + s)), on = xyz(__ldg(P.onormal + s));
            SampleGeom g;
            // Cheap exact-safe cull before any RNG: with face-normal shading, a triangle whose plane has the source clearly on
            // its back side (and which lies clearly in front of the wall point) has n.d > 0 and n_o.d > 0 for EVERY point on
            // it, so ff = -(n.d)(n_o.d)/r^2 < 0 and the sample adds nothing (TG.cpp:224-228).  The margins (1e-5 relative)
            // are far above float round-off; anything closer to edge-on goes through the full path.
            bool culled = false;
            if (!HAS_VN && !P.sr) {
              const f3 w1 = t.st.v1 - o, w2 = t.st.v2 - o, w3 = t.st.v3 - o;
              const float m1 = 1e-5f * (fabsf(w1.x) + fabsf(w1.y) + fabsf(w1.z));
              culled = dot3(t.st.nf, w1) > m1 && dot3(on, w1) > m1 &&
                       dot3(on, w2) > 1e-5f * (fabsf(w2.x) + fabsf(w2.y) + fabsf(w2.z)) &&
                       dot3(on, w3) > 1e-5f * (fabsf(w3.x) + fabsf(w3.y) + fabsf(w3.z));
            }
            // Embree's edge form is rebuilt from the shading vertices (same expressions as k_tri_records -> same bits) instead of
            // living in 9 more registers across the traversal: with it the kernel fits 6 blocks per SM (26.4 -> 24.6 ms)
            const TriRec trr = make_tri(t.st.v1, t.st.v2, t.st.v3);
            if (!culled && draw_sample(P, P.src_offset + s, t.prim, k, o, t.st, trr, g) && g.r <= ub_half && g.r >= lb_half) {
              const f3 n = shading_normal<HAS_VN>(t, g);
              const float ff = -dot3(n, g.d) * dot3(on, g.d) / g.r / g.r;          // TG.cpp:224-227
              if (P.sr ? ff != 0.0f : ff > 0.0f) {                                  // max(0,ff)==0 adds exactly 0 (TG.cpp:228); SR has no clamp
                const float alb = (MODE == 1) ? 1.0f : shading_albedo<HAS_VA>(t, g);
                float v = t.st.A * alb * ff * ff;
                if (GGX) v = v * ggx_eval(P.alpha, dot3(n, -g.d));                  // ggx/TG.cpp:236-238
                if (MODE == 0) {
                  const int64_t b = (int64_t)floorf((2.0f * g.r - P.lb) / P.res_fwd);   // TG.cpp:229
                  r_bin = (b >= 0 && b < nbf) ? (int)b : -1;
                }
                need = true; r_dx = g.d.x; r_dy = g.d.y; r_dz = g.d.z; r_ts = g.t; r_val = v;
              }
            }
          }
          const unsigned mask = __ballot_sync(0xffffffffu, need);
          if (need) {
            const int pos = (qhead + qcount + __popc(mask & lt)) & (kQCap - 1);
            ws.dx[pos] = r_dx; ws.dy[pos] = r_dy; ws.dz[pos] = r_dz; ws.ts[pos] = r_ts; ws.val[pos] = r_val; ws.bin[pos] = r_bin;
            ws.meta[pos] = gen | (lane << 16);
          }
          qcount += __popc(mask); ++gen;
        }
        __syncwarp();
        // ---------------- refill: idle lanes pop rays (of any triangle of the tile)
        const int take = n_idle < qcount ? n_idle : qcount;
        if (take == 0 && n_idle == 32) break;                  // slots exhausted, queue empty, nobody tracing
        const int rank = __popc(idle_mask & lt);
        const bool fetch = cur == kSentinel && rank < take;
        int meta = 0;
        if (fetch) {
          const int pos = (qhead + rank) & (kQCap - 1);
          meta = ws.meta[pos];
          ts = ws.ts[pos]; val = ws.val[pos]; bin = ws.bin[pos];
          slot_local = meta & 0xffff; tri_lane = (meta >> 16) & 31;
          const int64_t slot = slot0 + slot_local;
          src = P.spp == 1 ? slot : slot / P.spp;
          ray = make_ray(xyz(__ldg(P.origin + src)), mk3(ws.dx[pos], ws.dy[pos], ws.dz[pos]));
          stack[0] = kDone; sp = 1; cur = sc.root_count > 0 ? leaf_ref(0, sc.root_count) : 0;
        }
        const int prim_new = __shfl_sync(0xffffffffu, t.prim, (meta >> 16) & 31);
        if (fetch) prim = prim_new;
        qhead = (qhead + take) & (kQCap - 1); qcount -= take;
        __syncwarp();                            // pops complete before the next phase A overwrites the ring
      }
      // ---------------- one traversal round: internal nodes until a leaf run (or until too few lanes are still walking
      // nodes: NLOS_MINLANES keeps a few long node runs from stalling the lanes that already hold a leaf), then that leaf run.
      // (Postponing leaf runs to keep walking nodes — "speculative traversal" — was measured slower here: 42.5 vs 39.4 ms.)
      const float tlim = ts * 1.000001f;
      while ((unsigned)cur < (unsigned)kSentinel) {
        float4 a, b, c, dq;
        ld256(&sc.nodes[cur].a, a, b);
        ld256(&sc.nodes[cur].c, c, dq);
        const int r0 = __float_as_int(dq.x), r1 = __float_as_int(dq.y);
        float t0, t1;
        const bool h0 = slab(ray, a.x, a.y, a.z, a.w, b.x, b.y, tlim, t0);
        const bool h1 = slab(ray, b.z, b.w, c.x, c.y, c.z, c.w, tlim, t1);
        {   // fully predicated child selection: no BSSY/BSYNC pair inside the step (measured 33.3 vs 35.8 ms with if/else).
            // Keeping the stack top in a register (pop without a dependent local load) was measured SLOWER (27.8 vs 26.8 ms).
          const bool both = h0 && h1, any = h0 || h1, first0 = t0 <= t1;
          const int nearr = (h0 && (!h1 || first0)) ? r0 : r1;
          if (both) stack[sp] = first0 ? r1 : r0;
          sp += both ? 1 : 0;
          int popped = kDone;
          if (!any) popped = stack[sp - 1];
          sp -= any ? 0 : 1;
          cur = any ? nearr : popped;
        }
#if NLOS_MINLANES > 0
        if (__popc(__activemask()) < NLOS_MINLANES) break;
#endif
      }
      if (cur < 0 && cur != kDone) {
        const int first = leaf_first(cur), cnt = leaf_count(cur);
        bool occ = false;
        for (int j = 0; j < cnt && !occ; ++j) occ = tri_occludes_fast(sc.ttris, first + j, ray, ts, prim);
        cur = occ ? kSentinel : stack[--sp];                     // occluded: drop the ray
      }
      // a ray that ended visible only raises a flag: its histogram update runs at the next refill, when >= kRefill lanes are idle and
      // most of them have one pending, instead of at 10-14 active lanes right here (forward 24.4 -> 23.8 ms, smoothed forward 28.0 -> 25.1)
      if (cur == kDone) { pending = true; cur = kSentinel; }
    }
    if (WRITE_VIS) {
      __syncwarp();
      if (warp_global * 32 < sc.F) for (int i = lane; i < nslots; i += 32) vis[(size_t)(slot0 + i) * P.words_per_row + warp_global] = ws.tile[i];
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------- K1g forward, perspective grid
// One BLOCK per wall point (persistent blocks stride over the sources).  Seen from the wall point the scene is a 2-D picture
// (nlos_core.cuh "per-source perspective grid"); per source the block
//   pass 0  projects the VERTICES once (picture coordinates to a per-block scratch array, picture rectangle, round-off padding,
//           "everything safely in front" check)
//   pass 1  per triangle: quantised rectangle from its three projected vertices (kept for pass 2); counts, per grid cell
//           (shared memory), the triangles whose rectangle overlaps it;  block scan -> offsets
//           (entry budget exceeded -> halve the resolution and recount: no host round trip, the answer does not depend on G)
//   pass 2  writes the cell lists into this block's slice of a global scratch buffer: 4-byte entry words (the rectangle clipped
//           to the cell, four guarded bytes) and the triangle index, in 32-byte blocks of four entries [E0 E1 E2 E3][T0 T1 T2 T3]
//   pass 3  lane <-> triangle as in k_forward (the warps draw their batches of 32 triangles from a block-wide counter): draw the
//           sample, self intersection, shading.  Every lane scans the list of ITS ray's cell (depth slices 0 .. slice of its own
//           hit are contiguous) with the packed pre-check — one add, one and, one compare per entry, four entries per 16-byte
//           load — and hands the survivors to the warp's pool; the pooled triangles then get the exact test tri_occludes_od with
//           all lanes of the warp in the same loop.  No tree walk, no per-lane stack.
// Visibility is still "no other triangle beats (t_self, prim)" over the exact float test — the grid only selects candidates,
// conservatively (tests/emul/pgrid_emul.cpp checks the selection against the BVH query on every ray of C-bunny).
// A source that does not see the whole mesh safely in front of it (a vertex with depth < zmin along the wall normal) falls
// back to the per-ray BVH query inside this kernel.
#ifndef NLOS_GRID_BLOCK
#define NLOS_GRID_BLOCK 1024
#endif
#ifndef NLOS_GRID_MINBLOCKS
#define NLOS_GRID_MINBLOCKS 1
#endif
constexpr int kGridBlock = NLOS_GRID_BLOCK;
#ifndef NLOS_GRID_K
#define NLOS_GRID_K 4
#endif
static_assert(NLOS_GRID_K == 1 || NLOS_GRID_K == 2 || NLOS_GRID_K == 4, "the slice index is packed into 2 bits");
constexpr int kGridK = NLOS_GRID_K;           // depth slices per picture cell at most (1, 2 or 4; the launcher passes the count Kz): a ray only scans the slices up to its own depth
#ifndef NLOS_GRID_PUSH
#define NLOS_GRID_PUSH 8
#endif
constexpr int kGridPush = NLOS_GRID_PUSH;   // candidates a lane hands to the warp's work pool per round
constexpr int kGridPool = 32 * kGridPush;  // (ray, candidate) work items per round
struct GridWarp {                          // per-warp scratch of pass 3
  float dx[32], dy[32], dz[32], ts[32]; int prim[32];   // the warp's rays, readable by every lane
  unsigned pool[kGridPool];                // work items: ray lane << 27 | candidate triangle (Morton index)
  unsigned occ;                            // bit l: the ray of lane l is occluded
  unsigned pad_[3];
};
#ifndef NLOS_GRID_SADDR
#define NLOS_GRID_SADDR 15
#endif
#if NLOS_GRID_SADDR & 1
// Pass 3 addresses the warp's scratch through ONE 32-bit shared-window address that went through a shuffle: ptxas otherwise
// re-derives it (S2R SR_TID.X, S2R SR_CgaCtaId, shift, multiply-add) inside the exact-test and push loops instead of holding a register
__device__ __forceinline__ unsigned lds_u(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f(unsigned a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f(unsigned a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void reds_or(unsigned a, unsigned v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#define GW_F(field, i) lds_f(gwa + (unsigned)offsetof(GridWarp, field) + 4u * (unsigned)(i))
#define GW_U(field, i) lds_u(gwa + (unsigned)offsetof(GridWarp, field) + 4u * (unsigned)(i))
#define GW_OCC() lds_u(gwa + (unsigned)offsetof(GridWarp, occ))
#define GW_SET_F(field, i, v) sts_f(gwa + (unsigned)offsetof(GridWarp, field) + 4u * (unsigned)(i), (v))
#define GW_SET_U(field, i, v) sts_u(gwa + (unsigned)offsetof(GridWarp, field) + 4u * (unsigned)(i), (unsigned)(v))
#define GW_SET_OCC(v) sts_u(gwa + (unsigned)offsetof(GridWarp, occ), (v))
#define GW_OR_OCC(v) reds_or(gwa + (unsigned)offsetof(GridWarp, occ), (v))
#else
#define GW_F(field, i) gw.field[i]
#define GW_U(field, i) ((unsigned)gw.field[i])
#define GW_OCC() gw.occ
#define GW_SET_F(field, i, v) gw.field[i] = (v)
#define GW_SET_U(field, i, v) gw.field[i] = (v)
#define GW_SET_OCC(v) gw.occ = (v)
#define GW_OR_OCC(v) atomicOr(&gw.occ, (v))
#endif
// a block-uniform global pointer that went through a shuffle: ptxas holds it (in uniform registers) instead of re-deriving
// base + blockIdx.x * stride (S2R, 64-bit multiply-add, LEA pair) at every access of the binning loops
template <class T> __device__ __forceinline__ T* uniform_global_ptr(T* p) {
  unsigned long long a = (unsigned long long)__cvta_generic_to_global(p);
  a = __shfl_sync(0xffffffffu, a, 0);
  return reinterpret_cast<T*>(__cvta_global_to_generic((size_t)a));
}
struct GridShared {
  PGridFrame fr;
  float zmin, pad_u, pad_v;
  float z0, sz, pad_z;                     // depth slice of depth z: clamp(floor((z - z0) * sz), 0, kGridK-1)
  int use_grid;
  unsigned rect[6];                        // ordered-uint min u, max u, min v, max v, min z, max z of the projected vertices
  unsigned maxm;                           // max round-off margin (float bits, m >= 0)
  unsigned next_batch;                     // pass 3: first triangle of the next batch nobody has taken yet
  unsigned wsum[kGridBlock / 32];
};
struct GridScratch { float4* proj; uint2* rect; unsigned* ent;
                     unsigned long long* counters; };   // counters: null, or 6 work counters (option "count_work"): samples generated, rays traced, entry words scanned, cell-level check passes, visible samples, sources without grid

// exclusive prefix sum of the counts a[0..n), each rounded up to a multiple of 4, in place (shared memory); returns the total;
// per-thread runs of odd length avoid bank conflicts
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned* a, int n, unsigned* wsum) {
  const int T = kGridBlock, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = ((n + T - 1) / T) | 1;
  const int b = min(n, tid * per), e = min(n, b + per);
  unsigned sum = 0;
  for (int i = b; i < e; ++i) sum += (a[i] + 3u) & ~3u;               // every list is padded to whole groups of 4 entries
  unsigned x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) wsum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < T / 32 ? wsum[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
    if (lane < T / 32) wsum[lane] = w;
  }
  __syncthreads();
  unsigned base = (warp ? wsum[warp - 1] : 0u) + x - sum;
  const unsigned total = wsum[T / 32 - 1];
  for (int i = b; i < e; ++i) { const unsigned c = (a[i] + 3u) & ~3u; a[i] = base; base += c; }
  return total;
}

// KZT: depth slices as a compile-time constant (kGridK: the common case, 2 % faster than the run-time count) or 0 = the run-time count Kz_arg
template <bool GGX, bool HAS_VN, bool HAS_VA, bool SMOOTH, bool WRITE_VIS, int MODE, bool COUNT, int KZT>
__global__ void __launch_bounds__(kGridBlock, NLOS_GRID_MINBLOCKS) k_forward_grid(const DeviceScene sc, const RenderParams P, double* __restrict__ out,
                                                    uint32_t* __restrict__ vis, const double* __restrict__ wprefix,
                                                    const GridScratch scr, unsigned cap, int G0, int Kz_arg) {
  const int Kz = KZT ? KZT : Kz_arg;
#if NLOS_GRID_SADDR & 2     // experiment: the whole shared window addressed from a shuffled (= warp-uniform, not rematerialisable) base
  extern __shared__ __align__(16) unsigned char smem_raw0[];
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic(__shfl_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(smem_raw0), 0)));
#else
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  GridShared& gs = *reinterpret_cast<GridShared*>(smem_raw);
  double* s_w = reinterpret_cast<double*>(smem_raw + ((sizeof(GridShared) + 15) & ~size_t(15)));                  // SMOOTH: tap prefix sums
  GridWarp* gw_all = reinterpret_cast<GridWarp*>(reinterpret_cast<unsigned char*>(s_w) + (SMOOTH ? (((size_t)(P.K + 1) * sizeof(double) + 15) & ~size_t(15)) : 0));
  unsigned* cells = reinterpret_cast<unsigned*>(gw_all + kGridBlock / 32);
  if (SMOOTH) { for (int i = threadIdx.x; i <= P.K; i += kGridBlock) s_w[i] = wprefix[i]; }
#if NLOS_GRID_SADDR & 4
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform to the compiler as well
#else
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#endif
  GridWarp& gw = gw_all[warp];
#if NLOS_GRID_SADDR & 1
  const unsigned gwa = __shfl_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(&gw), 0);
#endif
#if NLOS_GRID_SADDR & 8
  float4* __restrict__ proj = uniform_global_ptr(scr.proj + (size_t)blockIdx.x * sc.V);
  uint2* __restrict__ trect = uniform_global_ptr(scr.rect + (size_t)blockIdx.x * sc.F);
  unsigned* __restrict__ ent = uniform_global_ptr(scr.ent + 2 * (size_t)blockIdx.x * cap);
#else
  float4* __restrict__ proj = scr.proj + (size_t)blockIdx.x * sc.V;
  uint2* __restrict__ trect = scr.rect + (size_t)blockIdx.x * sc.F;
  unsigned* __restrict__ ent = scr.ent + 2 * (size_t)blockIdx.x * cap;
#endif      // blocks of 4 entries (32 bytes): [E0 E1 E2 E3][T0 T1 T2 T3] — a candidate's triangle word sits in the sector its rectangle word was read from
  const float ub_half = P.ub / 2.0f, lb_half = P.lb / 2.0f;
  const int64_t nbf = (int64_t)P.numBins * P.r_fwd;
  const int F = sc.F;
  const float pad = __int_as_float((int)sc.bounds->absmax) * (1.0f / 65536.0f);
  for (int64_t s = blockIdx.x; s < P.L; s += gridDim.x) {
    const f3 o = xyz(__ldg(P.origin + s)), on = xyz(__ldg(P.onormal + s));
    // ---------------- frame
    if (tid == 0) {
      bool ok; pg_init_frame(o, on, gs.fr, ok);
      float blo[3], bhi[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { blo[k] = ord2f(sc.bounds->vlo[k]) - pad; bhi[k] = ord2f(sc.bounds->vhi[k]) + pad; }
      gs.zmin = pg_zmin(blo, bhi);
      gs.use_grid = (ok && G0 > 0) ? 1 : 0;
      gs.rect[0] = f2ord(3.0e38f); gs.rect[1] = f2ord(-3.0e38f); gs.rect[2] = f2ord(3.0e38f); gs.rect[3] = f2ord(-3.0e38f); gs.rect[4] = f2ord(3.0e38f); gs.rect[5] = f2ord(-3.0e38f); gs.maxm = 0u;
    }
    __syncthreads();
    // ---------------- pass 0: project the vertices
    if (gs.use_grid) {
      const PGridFrame fr = gs.fr; const float zmin = gs.zmin;
      float U0 = 3.0e38f, U1 = -3.0e38f, V0 = 3.0e38f, V1 = -3.0e38f, Z0 = 3.0e38f, Z1 = -3.0e38f, mm = 0.f; bool zok = true;
      for (int i = tid; i < sc.V; i += kGridBlock) {
        const f3 x = mk3(__ldg(sc.verts + 3 * (size_t)i), __ldg(sc.verts + 3 * (size_t)i + 1), __ldg(sc.verts + 3 * (size_t)i + 2));
        float u, v, m, z; pg_project(fr, x, u, v, m, z);
        zok = zok && (z >= zmin);
        proj[i] = make_float4(u, v, z, 0.f);
        U0 = fminf(U0, u); U1 = fmaxf(U1, u); V0 = fminf(V0, v); V1 = fmaxf(V1, v); Z0 = fminf(Z0, z); Z1 = fmaxf(Z1, z); mm = fmaxf(mm, m);
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        U0 = fminf(U0, __shfl_xor_sync(0xffffffffu, U0, d)); U1 = fmaxf(U1, __shfl_xor_sync(0xffffffffu, U1, d));
        V0 = fminf(V0, __shfl_xor_sync(0xffffffffu, V0, d)); V1 = fmaxf(V1, __shfl_xor_sync(0xffffffffu, V1, d));
        Z0 = fminf(Z0, __shfl_xor_sync(0xffffffffu, Z0, d)); Z1 = fmaxf(Z1, __shfl_xor_sync(0xffffffffu, Z1, d));
        mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, d));
      }
      zok = __all_sync(0xffffffffu, zok);
      if (lane == 0) {
        if (!zok || !(mm < 1.0f)) gs.use_grid = 0;
        atomicMin(&gs.rect[0], f2ord(U0)); atomicMax(&gs.rect[1], f2ord(U1)); atomicMin(&gs.rect[2], f2ord(V0)); atomicMax(&gs.rect[3], f2ord(V1));
        atomicMin(&gs.rect[4], f2ord(Z0)); atomicMax(&gs.rect[5], f2ord(Z1));
        atomicMax(&gs.maxm, (unsigned)__float_as_int(mm));
      }
    }
    __syncthreads();
    if (tid == 0 && gs.use_grid) {
      float U0 = ord2f(gs.rect[0]), U1 = ord2f(gs.rect[1]), V0 = ord2f(gs.rect[2]), V1 = ord2f(gs.rect[3]);
      const float mm = __int_as_float((int)gs.maxm);
      gs.pad_u = mm * (1.0f + fmaxf(fabsf(U0), fabsf(U1))); gs.pad_v = mm * (1.0f + fmaxf(fabsf(V0), fabsf(V1)));
      pg_set_rect(gs.fr, U0 - gs.pad_u, U1 + gs.pad_u, V0 - gs.pad_v, V1 + gs.pad_v, G0);
      const float Z0 = ord2f(gs.rect[4]), Z1 = ord2f(gs.rect[5]);
      gs.pad_z = 1.0e-5f * fabsf(Z1); gs.z0 = Z0; gs.sz = (float)Kz / fmaxf(Z1 - Z0, 1.0e-30f);
      if (gs.fr.G == 0) gs.use_grid = 0;
    }
    __syncthreads();
    // ---------------- pass 1: count (coarsen until the entries fit), then pass 2: fill
    while (gs.use_grid) {
      const PGridFrame fr = gs.fr; const float pad_u = gs.pad_u, pad_v = gs.pad_v, z0 = gs.z0, sz = gs.sz, pad_z = gs.pad_z;
      const int G = fr.G, ncell = G * G * Kz;
      for (int i = tid; i < ncell; i += kGridBlock) cells[i] = 0u;
      __syncthreads();
      // two-stage software pipeline: the vertex indices of triangle p + 2*stride and the projected vertices of p + stride are in
      // flight while triangle p is binned (each is a dependent L2 round trip otherwise: 11 % of the kernel's stall samples)
      {
        int p = tid;
        float4 s3n = make_float4(0.f, 0.f, 0.f, 0.f), q1 = s3n, q2 = s3n, q3 = s3n;
        if (p < F) {
          const float4 s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
          q1 = proj[__float_as_int(s3.y)]; q2 = proj[__float_as_int(s3.z)]; q3 = proj[__float_as_int(s3.w)];
          if (p + kGridBlock < F) s3n = __ldg(sc.stris + 3 * (size_t)sc.F + (p + kGridBlock));
        }
        for (; p < F; p += kGridBlock) {
          const float4 p1 = q1, p2 = q2, p3 = q3;
          if (p + kGridBlock < F) {
            q1 = proj[__float_as_int(s3n.y)]; q2 = proj[__float_as_int(s3n.z)]; q3 = proj[__float_as_int(s3n.w)];
            if (p + 2 * kGridBlock < F) s3n = __ldg(sc.stris + 3 * (size_t)sc.F + (p + 2 * kGridBlock));
          }
          int a0, a1, b0, b1;
          pg_tri_rect(fr, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y, pad_u, pad_v, a0, a1, b0, b1);
          // depth slice of the triangle's nearest vertex: it can only occlude rays whose own hit is at least that deep
          const int kz = pg_quant(fminf(p1.z, fminf(p2.z, p3.z)) - pad_z, z0, sz, (float)(Kz - 1));
          trect[p] = make_uint2((unsigned)a0 | ((unsigned)a1 << 16) | ((unsigned)(kz & 1) << 15) | ((unsigned)(kz >> 1) << 31), (unsigned)b0 | ((unsigned)b1 << 16));
          const int cx0 = a0 >> kPgSub, cx1 = a1 >> kPgSub, cy1 = b1 >> kPgSub;
          const int cy0 = b0 >> kPgSub;
          if (cx0 == cx1 && cy0 == cy1) atomicAdd(&cells[(cy0 * G + cx0) * Kz + kz], 1u);       // the common case: one cell
          else for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) atomicAdd(&cells[(cy * G + cx) * Kz + kz], 1u);
        }
      }
      __syncthreads();
      const unsigned total = block_exclusive_scan(cells, ncell, gs.wsum);
      __syncthreads();
      if (total <= cap) {
        uint2 rn = tid < F ? trect[tid] : make_uint2(0u, 0u);                                  // next triangle's rectangle in flight while this one is scattered
        for (int p = tid; p < F; p += kGridBlock) {
          const uint2 r = rn;
          if (p + kGridBlock < F) rn = trect[p + kGridBlock];
          const int a0 = (int)(r.x & 0x7fffu), a1 = (int)((r.x >> 16) & 0x7fffu), b0 = (int)(r.y & 0xffffu), b1 = (int)(r.y >> 16);
          const int kz = (int)((r.x >> 15) & 1u) | (int)((r.x >> 31) << 1);
          const int cx0 = a0 >> kPgSub, cx1 = a1 >> kPgSub, cy1 = b1 >> kPgSub;
          const int cy0 = b0 >> kPgSub;
          if (cx0 == cx1 && cy0 == cy1) {                                                      // the common case: one cell
            const unsigned pos = atomicAdd(&cells[(cy0 * G + cx0) * Kz + kz], 1u);
            const unsigned w = pos + (pos & ~3u);                                              // entry pos -> word (pos / 4) * 8 + pos % 4
            ent[w] = pg_entry(a0, a1, b0, b1, cx0, cy0); ent[w + 4] = (unsigned)p;
          } else for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) {
            const unsigned pos = atomicAdd(&cells[(cy * G + cx) * Kz + kz], 1u);
            const unsigned w = pos + (pos & ~3u);
            ent[w] = pg_entry(a0, a1, b0, b1, cx, cy); ent[w + 4] = (unsigned)p;
          }
        }
        __syncthreads();
        // the tail of every list up to its group-of-4 boundary gets the never-matching word 0 (lists are scanned 4 entries per load)
        for (int c = tid; c < ncell; c += kGridBlock) { const unsigned e = cells[c]; for (unsigned k = e; k < ((e + 3u) & ~3u); ++k) ent[k + (k & ~3u)] = 0u; }
        __syncthreads();
        break;
      }
      if (G == 1) { if (tid == 0) gs.use_grid = 0; __syncthreads(); break; }              // cannot happen (cap >= F), kept as a guard
      if (tid == 0) pg_coarsen(gs.fr);
      __syncthreads();
    }
    // ---------------- pass 3: samples
    if (tid == 0) gs.next_batch = 0u;
    __syncthreads();
    const bool grid = gs.use_grid != 0;
    if (COUNT && tid == 0 && !grid) atomicAdd(scr.counters + 5, 1ull);
    const PGridFrame& fr = gs.fr;                 // read from shared memory where needed (a register copy of its 18 words spills)
    const int G = fr.G;
    // the warps draw their work units — one sample index of 32 triangles — from a block-wide counter: a static stride leaves the block
    // waiting at the barrier below for its slowest warp (7.5 % of the kernel's stall samples at C-bunny), and with whole batches as units
    // the 36 batches x 18 samples of C-arm on 32 warps left 31 % barrier stalls (the triangle is re-read per sample index: 5 % of a sample)
    const unsigned nunits = (unsigned)((F + 31) / 32) * (unsigned)P.spp;
    for (;;) {
      unsigned u = 0u;
      if (lane == 0) u = atomicAdd(&gs.next_batch, 1u);
      u = __shfl_sync(0xffffffffu, u, 0);
      if (u >= nunits) break;
      const int base = (int)(P.spp == 1 ? u : u / (unsigned)P.spp) * 32;
      const int k = P.spp == 1 ? 0 : (int)(u % (unsigned)P.spp);
      const int p = base + lane;
      const bool active = p < F;
      TriRegs t; t.prim = 0;
      bool culled = true;
      if (active) {
        load_tri<HAS_VN, HAS_VA>(sc, p, t);
        culled = false;
        if (!HAS_VN && !P.sr) {                                   // same exact-safe plane-side cull as k_forward
          const f3 w1 = t.st.v1 - o, w2 = t.st.v2 - o, w3 = t.st.v3 - o;
          const float m1 = 1e-5f * (fabsf(w1.x) + fabsf(w1.y) + fabsf(w1.z));
          culled = dot3(t.st.nf, w1) > m1 && dot3(on, w1) > m1 &&
                   dot3(on, w2) > 1e-5f * (fabsf(w2.x) + fabsf(w2.y) + fabsf(w2.z)) &&
                   dot3(on, w3) > 1e-5f * (fabsf(w3.x) + fabsf(w3.y) + fabsf(w3.z));
        }
      }
      if (__all_sync(0xffffffffu, culled)) {
        if (WRITE_VIS && lane == 0) vis[(size_t)(s * P.spp + k) * P.words_per_row + (base >> 5)] = 0u;
        continue;
      }
      {
        bool need = false; float val = 0.f, ts = 0.f; int bin = -1; f3 d = mk3(0.f, 0.f, 1.f);
        if (!culled) {
          SampleGeom g;
          const TriRec trr = make_tri(t.st.v1, t.st.v2, t.st.v3);
          if (draw_sample(P, P.src_offset + s, t.prim, k, o, t.st, trr, g) && g.r <= ub_half && g.r >= lb_half) {
            const f3 n = shading_normal<HAS_VN>(t, g);
            const float ff = -dot3(n, g.d) * dot3(on, g.d) / g.r / g.r;          // TG.cpp:224-227
            if (P.sr ? ff != 0.0f : ff > 0.0f) {
              const float alb = (MODE == 1) ? 1.0f : shading_albedo<HAS_VA>(t, g);
              val = t.st.A * alb * ff * ff;
              if (GGX) val = val * ggx_eval(P.alpha, dot3(n, -g.d));              // ggx/TG.cpp:236-238
              if (MODE == 0) {
                const int64_t b = (int64_t)floorf((2.0f * g.r - P.lb) / P.res_fwd);   // TG.cpp:229
                bin = (b >= 0 && b < nbf) ? (int)b : -1;
              }
              need = true; d = g.d; ts = g.t;
            }
          }
        }
        bool occ = false;
        if (grid) {
          // a ray that does not point into the picture's half space (n_o.d <= 0 can only happen with an unclamped form factor)
          // takes the BVH query
          const bool in_picture = need && dot3(d, fr.n) > 0.0f;
          if (need && !in_picture) occ = occluded(sc.nodes, sc.ttris, sc.root_count, make_ray(o, d), ts, t.prim);
          int cx = 0, cy = 0; unsigned R = 0u;
          if (in_picture) pg_ray(fr, d, cx, cy, R);
          // this lane's list: groups of 4 entry words, 16-byte aligned (cells[] holds the END of every list after pass 2)
          int start = 0, ngrp = 0;
          if (in_picture) {
            // slices 0 .. kr of the ray's picture cell are contiguous: one scan range (the padding words between them never match)
            const int kr = pg_quant(ts * dot3(d, fr.n) + gs.pad_z, gs.z0, gs.sz, (float)(Kz - 1));
            const int c = (cy * G + cx) * Kz; start = c ? (int)((cells[c - 1] + 3u) & ~3u) : 0; ngrp = ((int)cells[c + kr] - start + 3) >> 2;
          }
          const uint4* __restrict__ lst = reinterpret_cast<const uint4*>(ent) + 2 * (size_t)(start >> 2);
          // the warp's rays, readable by every lane: the exact tests below are pooled over the warp
          GW_SET_F(dx, lane, d.x); GW_SET_F(dy, lane, d.y); GW_SET_F(dz, lane, d.z); GW_SET_F(ts, lane, ts); GW_SET_U(prim, lane, t.prim);
          if (lane == 0) GW_SET_OCC(0u);
          __syncwarp();
          const int maxg = __reduce_max_sync(0xffffffffu, ngrp);
          unsigned nhit = 0u;                                                     // work counter: cell-level check passes of this lane
          for (int g0 = 0; g0 < maxg; g0 += 8) {                                  // 32 entries per lane and round
            unsigned mask = 0u;
            if (g0 < ngrp && !((GW_OCC() >> lane) & 1u)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (g0 + i < ngrp) {
                  const uint4 e = lst[2 * (g0 + i)];
                  if (pg_precheck(e.x, R)) mask |= 1u << (4 * i);
                  if (pg_precheck(e.y, R)) mask |= 2u << (4 * i);
                  if (pg_precheck(e.z, R)) mask |= 4u << (4 * i);
                  if (pg_precheck(e.w, R)) mask |= 8u << (4 * i);
                }
              }
            }
            if (COUNT) nhit += __popc(mask);
            // pooled exact tests: every lane hands up to kGridPush candidates to the warp's pool, then all 32 lanes work the pool
            while (__any_sync(0xffffffffu, mask != 0u)) {
              const int nb = __popc(mask);
              const int n = nb < kGridPush ? nb : kGridPush;
              // the pushing lane fetches the candidate's triangle index (up to kGridPush independent loads in flight) so that the exact-test
              // loop below starts with the triangle fetch instead of two dependent round trips
              // (the triangle word of entry b of this round: word 8 g0 + b + (b & ~3) + 4 of the list — same 32-byte sector as its rectangle word)
              const unsigned tag = (unsigned)lane << 27;
              const unsigned* __restrict__ tw = reinterpret_cast<const unsigned*>(lst) + (8 * g0 + 4);
              int incl = n;
#pragma unroll
              for (int o2 = 1; o2 < 32; o2 <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o2); if (lane >= o2) incl += y; }
              const int total = __shfl_sync(0xffffffffu, incl, 31);
              int off = incl - n;
              for (int i = 0; i < n; ++i) { const unsigned bpos = 31u - (unsigned)__clz(mask); mask ^= 1u << bpos; GW_SET_U(pool, off + i, tag | tw[bpos + (bpos & ~3u)]); }
              __syncwarp();
              for (int i = lane; i < total; i += 32) {
                const unsigned item = GW_U(pool, i);
                const int rl = (int)(item >> 27);
                if (!((GW_OCC() >> rl) & 1u)) {
                  const int tj = (int)(item & 0x7ffffffu);
                  if (tj != base + rl && tri_occludes_od(sc.ttris, tj, o, mk3(GW_F(dx, rl), GW_F(dy, rl), GW_F(dz, rl)), GW_F(ts, rl), (int)GW_U(prim, rl))) GW_OR_OCC(1u << rl);
                }
              }
              __syncwarp();
            }
          }
          __syncwarp();
          occ = occ || ((GW_OCC() >> lane) & 1u);
          __syncwarp();                                                             // gw is rewritten by the next sample
          if (COUNT) {                                                              // measurement instantiation only (bench.py roofline.executed)
            const unsigned c1 = __reduce_add_sync(0xffffffffu, need ? 1u : 0u), c2 = __reduce_add_sync(0xffffffffu, (unsigned)(4 * ngrp)), c3 = __reduce_add_sync(0xffffffffu, nhit);
            const unsigned c0 = __reduce_add_sync(0xffffffffu, culled ? 0u : 1u), c4 = __reduce_add_sync(0xffffffffu, (need && !occ) ? 1u : 0u);
            if (lane == 0) { atomicAdd(scr.counters, (unsigned long long)c0); atomicAdd(scr.counters + 1, (unsigned long long)c1); atomicAdd(scr.counters + 2, (unsigned long long)c2);
                             atomicAdd(scr.counters + 3, (unsigned long long)c3); atomicAdd(scr.counters + 4, (unsigned long long)c4); }
          }
        } else if (need) {
          occ = occluded(sc.nodes, sc.ttris, sc.root_count, make_ray(o, d), ts, t.prim);
        }
        const bool visible = need && !occ;
        if (visible) {
          const double dv = P.spp == 1 ? (double)val : (double)val / (double)P.spp;   // TG.cpp:231-232
          if (MODE == 1) atomicAdd(out + t.prim, dv);
          else if (bin >= 0) {
            if (!SMOOTH) atomicAdd(out + s * P.numBins + bin, dv);
            else {
              const int half = 2 * P.r_fwd * P.s_bin;
              const int q0 = floordiv32(bin, P.r_fwd); int b0 = q0 - 2 * P.s_bin, b1 = q0 + 2 * P.s_bin;      // half = 2 r s is a multiple of r: one division
              if (b0 < 0) b0 = 0; if (b1 > P.numBins - 1) b1 = P.numBins - 1;
              for (int b = b0; b <= b1; ++b) {
                int ilo, ihi; tap_span32(bin, b, P.r_fwd, half, P.K, ilo, ihi);
                if (ihi > ilo) atomicAdd(out + s * P.numBins + b, dv * (s_w[ihi] - s_w[ilo]));
              }
            }
          }
        }
        if (WRITE_VIS) {
          const unsigned m = __ballot_sync(0xffffffffu, visible);
          if (lane == 0) vis[(size_t)(s * P.spp + k) * P.words_per_row + (base >> 5)] = m;
        }
      }
    }
    __syncthreads();       // the next source reuses the cell array, the frame and the scratch arrays
  }
}

// ---------------------------------------------------------------------------------------------- K1s forward, shared perspective grid
// The grid of k_forward_grid, built ONCE per group of neighbouring wall points (group_grid.cu, k_group_bin) instead of once per wall
// point: seen from the group's centre the ray of a member wall point is a straight line of the projective space (u, v, 1/Z)
// (nlos_core.cuh "shared perspective grid of a GROUP"), so the lists are 3-D — picture cell x slice of 1/Z — and a ray tests, slice by
// slice up to the slice of its own hit, the list of the cell its line crosses there.  Nothing is built per wall point any more, so
// nothing is shared between the warps of a block either: the unit of work is (wall point, run of triangles) per WARP, without any
// block-level synchronisation.
//   generate  lane <-> triangle as in k_forward_grid (plane-side cull, Philox, self intersection, shading); samples that can contribute
//             are compacted into a per-warp ray queue (ballot + popc)
//   trace     whenever 32 rays are queued: lane <-> ray.  Per slice: cell lookup (one 8-byte record), packed rectangle check of the
//             list (4 entries per 16-byte load), survivors pooled over the warp for the exact test tri_occludes_od.  Visible rays add
//             their value to the transient row (FP64 RED) and set their bit of the visibility word (RED.OR; the buffer is zeroed first).
// The answer is the exact float test's, as before; rays that cannot use the grid (slope bound, wrong half space, group without grid)
// take the per-ray BVH query.
constexpr int kGrpBlock = 256;             // 8 warps
constexpr int kGrpQ = 64;                  // ray queue slots per warp (a trace starts at 32 queued rays; one generate round adds at most 32)
constexpr int kGrpPush = 8;                // candidates a lane hands to the warp's pool per round
constexpr int kGrpFlush = 128;             // the pool is worked off when it holds more than this many candidates
constexpr int kGrpPool = kGrpFlush + 32 * kGrpPush;
struct GrpWarp {
  float dx[kGrpQ], dy[kGrpQ], dz[kGrpQ], ts[kGrpQ], val[kGrpQ];
  int bin[kGrpQ], tri[kGrpQ], prim[kGrpQ], kk[kGrpQ];   // histogram bin (-1: none), triangle (Morton index), triangle (caller's index), sample index of the slot
  unsigned pool[kGrpPool];                 // work items of the exact tests: ray lane << 27 | candidate triangle
  unsigned occ;                            // bit l: the ray of lane l is occluded
  unsigned pad_[3];
  GGFrame fr;                              // the group's frame (copied per work item)
  unsigned pad2_[(128 - sizeof(GGFrame) % 128) / 4];
};
static_assert(sizeof(GrpWarp) % 16 == 0, "per-warp scratch must keep 16-byte alignment");

// exact tests of the pooled candidates, all 32 lanes in the same loop
__device__ __forceinline__ void group_work_pool(const DeviceScene& sc, GrpWarp& gw, f3 o, int qhead, int total, int lane) {
  __syncwarp();
  for (int i = lane; i < total; i += 32) {
    const unsigned item = gw.pool[i];
    const int rl = (int)(item >> 27);
    if (!((gw.occ >> rl) & 1u)) {
      const int rs = (qhead + rl) & (kGrpQ - 1);
      const int tj = (int)(item & 0x7ffffffu);
      if (tj != gw.tri[rs] && tri_occludes_od(sc.ttris, tj, o, mk3(gw.dx[rs], gw.dy[rs], gw.dz[rs]), gw.ts[rs], gw.prim[rs])) atomicOr(&gw.occ, 1u << rl);
    }
  }
  __syncwarp();
}

template <bool GGX, bool HAS_VN, bool HAS_VA, bool SMOOTH, bool WRITE_VIS, int MODE, bool COUNT>
__device__ __forceinline__ void group_trace(const DeviceScene& sc, const RenderParams& P, double* __restrict__ out, uint32_t* __restrict__ vis,
                                            const double* s_w, GrpWarp& gw, const GroupGrid& gg, bool grid, unsigned long long ent0, unsigned long long tab0,
                                            f3 o, float da, float db, float dn, int64_t s, int qhead, int n, int lane, unsigned long long* counters) {
  const int slot = (qhead + lane) & (kGrpQ - 1);
  const bool has = lane < n;
  const f3 d = mk3(gw.dx[slot], gw.dy[slot], gw.dz[slot]);
  const float ts = gw.ts[slot];
  const unsigned mytri = (unsigned)gw.tri[slot];
  bool occ = false; int kr = -1; GGRay r;
  r.ui = r.vi = r.su = r.sv = 0.f; r.kr = 0;
  if (has) {
    if (grid && gg_ray_setup(gw.fr, da, db, dn, d, ts, r)) kr = r.kr;
    else occ = occluded(sc.nodes, sc.ttris, sc.root_count, make_ray(o, d), ts, gw.prim[slot]);
  }
  if (lane == 0) gw.occ = 0u;
  __syncwarp();
  const int G = gw.fr.G, K = gw.fr.K;
  const uint4* __restrict__ ent = reinterpret_cast<const uint4*>(gg.ent) + 2 * (size_t)(ent0 >> 2);      // blocks of 4 entries: [E0 E1 E2 E3][T0 T1 T2 T3]
  const GGWalk wk = gg_ray_walk(gw.fr, r);
  const uint2* __restrict__ tab = gg.table + tab0;
  const float qmax = gw.fr.qmax;
  unsigned nscan = 0u, nhit = 0u;
  int np = 0;                                            // candidates waiting in the pool (warp-uniform)
  int k = 0; float kf = 0.0f;                            // next slice of this lane's ray
  for (;;) {
    // every lane walks its ray to the next slice whose cell has a non-empty list (most lists a ray crosses are empty)
    unsigned cnt = 0u, start = 0u; int uq = 0, vq = 0;
    while (k <= kr && cnt == 0u) {
      const float fu = fminf(fmaxf(fmaf(kf, wk.bu, wk.au), 0.0f), qmax), fv = fminf(fmaxf(fmaf(kf, wk.bv, wk.av), 0.0f), qmax);     // gg_walk_point
      uq = (int)fu; vq = (int)fv;
      const uint2 t = __ldg(tab + (unsigned)(((vq >> kPgSub) * G + (uq >> kPgSub)) * K + k));
      start = t.x >> 2; cnt = t.y; ++k; kf += 1.0f;
    }
    const int ngrp = (int)((cnt + 3u) >> 2);
    const int maxg = __reduce_max_sync(0xffffffffu, ngrp);
    if (maxg == 0) break;
    if (COUNT) nscan += 4u * (unsigned)ngrp;
    const unsigned R = gg_rect_word(uq, vq);
    const uint4* __restrict__ lst = ent + 2 * (size_t)start;
    for (int g0 = 0; g0 < maxg; g0 += 8) {                                  // 32 entries per lane and round
      unsigned mask = 0u;
      if (g0 < ngrp) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (g0 + i < ngrp) {
            const uint4 e = lst[2 * (g0 + i)];
            if (pg_precheck(e.x, R)) mask |= 1u << (4 * i);
            if (pg_precheck(e.y, R)) mask |= 2u << (4 * i);
            if (pg_precheck(e.z, R)) mask |= 4u << (4 * i);
            if (pg_precheck(e.w, R)) mask |= 8u << (4 * i);
          }
        }
      }
      if (COUNT) nhit += __popc(mask);
      // survivors go to the warp's pool (up to kGrpPush per lane and round); the pool is worked off by all 32 lanes when it is full enough
      while (__any_sync(0xffffffffu, mask != 0u)) {
        const int nb = __popc(mask);
        const int nn = nb < kGrpPush ? nb : kGrpPush;
        int incl = nn;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o2); if (lane >= o2) incl += y; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int off = np + incl - nn;
        for (int i = 0; i < nn; ++i) {
          const int bpos = __ffs(mask) - 1; mask &= mask - 1u;
          // the triangle words sit in the same 32-byte sector as the rectangle words just read
          gw.pool[off + i] = ((unsigned)lane << 27) | (reinterpret_cast<const unsigned*>(lst + 2 * (g0 + (bpos >> 2)) + 1)[bpos & 3] & 0x7ffffffu);
        }
        np += total;
        if (np > kGrpFlush) { group_work_pool(sc, gw, o, qhead, np, lane); np = 0; if ((gw.occ >> lane) & 1u) { mask = 0u; kr = -1; } }
      }
    }
  }
  if (np > 0) group_work_pool(sc, gw, o, qhead, np, lane);
  __syncwarp();
  occ = occ || ((gw.occ >> lane) & 1u);
  const bool visible = has && !occ;
  if (visible) {
    const float val = gw.val[slot]; const int bin = gw.bin[slot];
    const double dv = P.spp == 1 ? (double)val : (double)val / (double)P.spp;   // TG.cpp:231-232
    if (MODE == 1) atomicAdd(out + gw.prim[slot], dv);
    else if (bin >= 0) {
      if (!SMOOTH) atomicAdd(out + s * P.numBins + bin, dv);
      else {
        const int half = 2 * P.r_fwd * P.s_bin;
        const int q0 = floordiv32(bin, P.r_fwd); int b0 = q0 - 2 * P.s_bin, b1 = q0 + 2 * P.s_bin;      // half = 2 r s is a multiple of r: one division
        if (b0 < 0) b0 = 0; if (b1 > P.numBins - 1) b1 = P.numBins - 1;
        for (int b = b0; b <= b1; ++b) {
          int ilo, ihi; tap_span32(bin, b, P.r_fwd, half, P.K, ilo, ihi);
          if (ihi > ilo) atomicAdd(out + s * P.numBins + b, dv * (s_w[ihi] - s_w[ilo]));
        }
      }
    }
    if (WRITE_VIS) atomicOr(vis + (size_t)(s * P.spp + gw.kk[slot]) * P.words_per_row + (mytri >> 5), 1u << (mytri & 31u));
  }
  if (COUNT) {                                                              // measurement instantiation only (bench.py roofline.executed)
    const unsigned c1 = __reduce_add_sync(0xffffffffu, has ? 1u : 0u), c2 = __reduce_add_sync(0xffffffffu, nscan), c3 = __reduce_add_sync(0xffffffffu, nhit);
    const unsigned c4 = __reduce_add_sync(0xffffffffu, visible ? 1u : 0u);
    if (lane == 0) { atomicAdd(counters + 1, (unsigned long long)c1); atomicAdd(counters + 2, (unsigned long long)c2); atomicAdd(counters + 3, (unsigned long long)c3); atomicAdd(counters + 4, (unsigned long long)c4); }
  }
  __syncwarp();                                                             // the queue slots and gw.occ are reused by the next round
}

template <bool GGX, bool HAS_VN, bool HAS_VA, bool SMOOTH, bool WRITE_VIS, int MODE, bool COUNT>
__global__ void __launch_bounds__(kGrpBlock, 4) k_forward_group(const DeviceScene sc, const RenderParams P, double* __restrict__ out,
                                                    uint32_t* __restrict__ vis, const double* __restrict__ wprefix,
                                                    const GroupGrid gg, int chunk_batches, unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_w = reinterpret_cast<double*>(smem_raw);                                                  // SMOOTH: tap prefix sums
  GrpWarp* gw_all = reinterpret_cast<GrpWarp*>(smem_raw + (SMOOTH ? (((size_t)(P.K + 1) * sizeof(double) + 15) & ~size_t(15)) : 0));
  if (SMOOTH) { for (int i = threadIdx.x; i <= P.K; i += kGrpBlock) s_w[i] = wprefix[i]; __syncthreads(); }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  GrpWarp& gw = gw_all[warp];
  const float ub_half = P.ub / 2.0f, lb_half = P.lb / 2.0f;
  const int64_t nbf = (int64_t)P.numBins * P.r_fwd;
  const int F = sc.F;
  const unsigned lt = (1u << lane) - 1u;
  const int nchunks = (F + 32 * chunk_batches - 1) / (32 * chunk_batches);
  const int pos0 = __ldg(gg.gstart + gg.group0), npos = __ldg(gg.gstart + gg.group0 + gg.ngroups) - pos0;      // sorted wall positions of this batch
  const int64_t n_items = (int64_t)npos * nchunks;
  for (int64_t item = (int64_t)blockIdx.x * (kGrpBlock / 32) + warp; item < n_items; item += (int64_t)gridDim.x * (kGrpBlock / 32)) {
    const int pos = pos0 + (int)(item / nchunks), chunk = (int)(item % nchunks);
    const int64_t s = __ldg(gg.order + pos);
    const GroupHdr* __restrict__ h = gg.hdr + (__ldg(gg.group_of + pos) - gg.group0);
    __syncwarp();
    if (lane < (int)(sizeof(GGFrame) / 4)) reinterpret_cast<unsigned*>(&gw.fr)[lane] = __ldg(reinterpret_cast<const unsigned*>(&h->fr) + lane);
    const unsigned long long ent0 = h->ent0, tab0 = h->tab0;
    const int nlive = h->nlive;
    const int* __restrict__ live = gg.live + h->live0;
    __syncwarp();
    const bool grid = gw.fr.G > 0;
    const f3 o = xyz(__ldg(P.origin + s)), on = xyz(__ldg(P.onormal + s));
    const f3 del = o - gw.fr.o;
    const float da = dot3(del, gw.fr.a), db = dot3(del, gw.fr.b), dn = dot3(del, gw.fr.n);
    if (COUNT && lane == 0 && chunk == 0 && !grid) atomicAdd(counters + 5, 1ull);
    int qhead = 0, qcount = 0;
    // one generate step = one sample index of one batch of 32 triangles; the trace runs from a single call site whenever 32 rays are
    // queued, and once more for the remainder when the item's triangles are used up
    const int base0 = chunk * chunk_batches * 32;                      // position in the group's list of live triangles
    if (base0 >= nlive) continue;
    const int nb = min(chunk_batches, (nlive - base0 + 31) / 32);
    int bi = 0, k = 0;
    TriRegs t; t.prim = 0; bool culled = true; int p = 0;
    for (;;) {
      const bool more = bi < nb;
      if (more) {
        if (k == 0) {
          const int li = base0 + bi * 32 + lane;
          culled = true;
          if (li < nlive) {
            p = __ldg(live + li);
            load_tri<HAS_VN, HAS_VA>(sc, p, t);
            culled = false;
            if (!HAS_VN && !P.sr) {                                   // same exact-safe plane-side cull as k_forward
              const f3 w1 = t.st.v1 - o, w2 = t.st.v2 - o, w3 = t.st.v3 - o;
              const float m1 = 1e-5f * (fabsf(w1.x) + fabsf(w1.y) + fabsf(w1.z));
              culled = dot3(t.st.nf, w1) > m1 && dot3(on, w1) > m1 &&
                       dot3(on, w2) > 1e-5f * (fabsf(w2.x) + fabsf(w2.y) + fabsf(w2.z)) &&
                       dot3(on, w3) > 1e-5f * (fabsf(w3.x) + fabsf(w3.y) + fabsf(w3.z));
            }
          }
        }
        if (__all_sync(0xffffffffu, culled)) { ++bi; k = 0; continue; }
        bool need = false; float val = 0.f, ts = 0.f; int bin = -1; f3 d = mk3(0.f, 0.f, 1.f);
        if (!culled) {
          SampleGeom g;
          const TriRec trr = make_tri(t.st.v1, t.st.v2, t.st.v3);
          if (draw_sample(P, P.src_offset + s, t.prim, k, o, t.st, trr, g) && g.r <= ub_half && g.r >= lb_half) {
            const f3 n = shading_normal<HAS_VN>(t, g);
            const float ff = -dot3(n, g.d) * dot3(on, g.d) / g.r / g.r;          // TG.cpp:224-227
            if (P.sr ? ff != 0.0f : ff > 0.0f) {
              const float alb = (MODE == 1) ? 1.0f : shading_albedo<HAS_VA>(t, g);
              val = t.st.A * alb * ff * ff;
              if (GGX) val = val * ggx_eval(P.alpha, dot3(n, -g.d));              // ggx/TG.cpp:236-238
              if (MODE == 0) {
                const int64_t b = (int64_t)floorf((2.0f * g.r - P.lb) / P.res_fwd);   // TG.cpp:229
                bin = (b >= 0 && b < nbf) ? (int)b : -1;
              }
              need = true; d = g.d; ts = g.t;
            }
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (COUNT) { const unsigned c0 = __reduce_add_sync(0xffffffffu, culled ? 0u : 1u); if (lane == 0) atomicAdd(counters, (unsigned long long)c0); }
        if (need) {
          const int q = (qhead + qcount + __popc(m & lt)) & (kGrpQ - 1);
          gw.dx[q] = d.x; gw.dy[q] = d.y; gw.dz[q] = d.z; gw.ts[q] = ts; gw.val[q] = val; gw.bin[q] = bin; gw.tri[q] = p; gw.prim[q] = t.prim; gw.kk[q] = k;
        }
        qcount += __popc(m);
        if (++k == P.spp) { k = 0; ++bi; }
        __syncwarp();
      }
      if (qcount >= 32 || (!more && qcount > 0)) {
        const int n = qcount < 32 ? qcount : 32;
        group_trace<GGX, HAS_VN, HAS_VA, SMOOTH, WRITE_VIS, MODE, COUNT>(sc, P, out, vis, s_w, gw, gg, grid, ent0, tab0, o, da, db, dn, s, qhead, n, lane, counters);
        qhead = (qhead + n) & (kGrpQ - 1); qcount -= n;
      } else if (!more) break;
    }
  }
}

// ---------------------------------------------------------------------------------------------- K3 residual
// diff = (data - T) [-> 2 d^3 if loss_flag] * weight      (SSG.cpp:543-550)
__global__ void k_residual(const double* __restrict__ data, const double* __restrict__ weight, const double* __restrict__ T, double* __restrict__ diff, size_t n, int loss_flag) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double d = data[i] - T[i];
    if (loss_flag == 1) d = 2 * d * d * d;
    diff[i] = weight ? d * weight[i] : d;
  }
}

// centred (2w+1)-box mean of every residual row, zero padded and truncated to the row (one pass of SR/SSG.cpp:447-458; run twice)
__global__ void k_box_filter(const double* __restrict__ in, double* __restrict__ out, int B, size_t n, int width) {
  const double h = 1.0 / ((double)2 * width + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % (size_t)B); const double* row = in + (i - b);
    double a = 0;
    for (int j = 0; j <= 2 * width; ++j) { const int m = b + width - j; if (m >= 0 && m < B) a += h * row[m]; }
    out[i] = a;
  }
}

// ---------------------------------------------------------------------------------------------- K4/K5 gradients
#ifndef NLOS_GRAD_MINBLOCKS
#define NLOS_GRAD_MINBLOCKS 7      // <= 72 registers, 28 warps per SM: measured 1 (125 registers) 4.25 ms, 5 4.15, 6 4.11, 7 4.02, 8 4.04 @C-bunny; GGX + shading normals 7.24 -> 5.83
#endif
// KIND 0: vertex gradient (9 FP64 register accumulators per thread), 1: albedo scalar, 2: GGX alpha scalar
template <bool GGX, bool HAS_VN, bool HAS_VA, int KIND, bool USE_VIS>
__global__ void __launch_bounds__(kBlock, NLOS_GRAD_MINBLOCKS) k_gradient(const DeviceScene sc, const RenderParams P, const double* __restrict__ diff,
                                                     const uint32_t* __restrict__ vis, const double* __restrict__ wprefix,
                                                     const double* __restrict__ dprefix, double* __restrict__ out) {
  extern __shared__ double s_tab[];         // [0..K] prefix of w_i, [K+1..2K+1] prefix of w_i*delta_i
  double* s_w = s_tab; double* s_d = s_tab + (P.K + 1);
  for (int i = threadIdx.x; i <= P.K; i += blockDim.x) { s_w[i] = wprefix[i]; s_d[i] = dprefix[i]; }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < sc.F;
  TriRegs t;
  if (active) load_tri<HAS_VN, HAS_VA>(sc, p, t);
  const int64_t s0 = (int64_t)blockIdx.y * P.chunk;
  const int64_t s1 = s0 + P.chunk < P.L ? s0 + P.chunk : P.L;
  const float ub_half = P.ub / 2.0f, lb_half = P.lb / 2.0f;
  const int lane = threadIdx.x & 31;
  const int warp_global = p >> 5;
  const int half = 2 * P.r_grad * P.s_bin;
  double g1x = 0, g1y = 0, g1z = 0, g2x = 0, g2y = 0, g2z = 0, g3x = 0, g3y = 0, g3z = 0, gs = 0;
  f3 e1, e2, e3;
  if (active) { e1 = t.st.v3 - t.st.v2; e2 = t.st.v1 - t.st.v3; e3 = t.st.v2 - t.st.v1; }
  // slots (source*spp + k) are walked in groups of 32: lane l fetches the visibility word of slot base+l in ONE strided load
  // (32 independent L2 accesses in flight) and the words are then broadcast by shuffle, instead of one dependent L2 round trip
  // per slot at the top of the loop body.
  const int64_t slotA = s0 * P.spp, slotB = s1 * P.spp;
  const double inv_spp = 1.0 / (double)P.spp;
  for (int64_t base = slotA; base < slotB; base += 32) {
    unsigned myword = 0u;
    if (USE_VIS) {
      const int64_t sl = base + lane;
      if (sl < slotB && warp_global * 32 < sc.F) myword = __ldg(vis + (size_t)sl * P.words_per_row + warp_global);
    }
    const int cnt = (int)(slotB - base < 32 ? slotB - base : 32);
    // only the slots in which some lane of the warp has a visible sample (lane l holds the word of slot base + l): 4.01 -> 3.91 ms @C-bunny
    unsigned todo = USE_VIS ? __ballot_sync(0xffffffffu, myword != 0u) : (cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u));
    while (todo) {
      const int i = __ffs(todo) - 1; todo &= todo - 1u;
      const int64_t slot = base + i;
      const int64_t s = P.spp == 1 ? slot : slot / P.spp;
      const int k = P.spp == 1 ? 0 : (int)(slot - s * P.spp);
      bool bit = active;
      if (USE_VIS) { const unsigned m = __shfl_sync(0xffffffffu, myword, i); bit = active && ((m >> lane) & 1u); }
      if (!bit) continue;
      const float4 o4 = __ldg(P.origin + s), n4 = __ldg(P.onormal + s);
      const f3 o = xyz(o4), on = xyz(n4);
      SampleGeom g;
      if (!draw_sample(P, P.src_offset + s, t.prim, k, o, t.st, t.tr, g)) continue;
      if (!(g.r <= ub_half && g.r >= lb_half)) continue;
      const f3 n = shading_normal<HAS_VN>(t, g);
      const f3 d = g.d; const float hl = g.r;
      float c2 = dot3(on, d), c3 = dot3(n, -d);                       // TG.cpp:944-947
      if (c2 < 0) c2 = 0; if (c3 < 0) c3 = 0;
      if (!(c2 * c3 > 0.0f)) continue;                                // every term below carries c2*c3
      if (!USE_VIS) { const Ray ray = make_ray(o, d); if (occluded(sc.nodes, sc.ttris, sc.root_count, ray, g.t, t.prim)) continue; }
      const float alb = shading_albedo<HAS_VA>(t, g);
      const float ff = c2 * c3 / hl / hl;
      double At = 0.0, Bt = 0.0;
      if (P.jitter) {
        // jitter/TG.cpp:947-972: taps sit on WHOLE-bin offsets, so both sums depend only on (source, coarse bin) -> tables
        const int64_t b0 = (int64_t)floorf((2.0f * hl - P.lb) / P.res);
        if (b0 < 0 || b0 > P.numBins) continue;
        At = __ldg(P.jA + s * (P.numBins + 1) + b0); Bt = __ldg(P.jB + s * (P.numBins + 1) + b0);
      } else {
        // K-tap sums against the residual row, grouped per coarse bin (DESIGN.md "K-tap restructuring")
        const double x = ((double)(2.0f * hl) - (double)P.lb) * P.inv_res_fine;
        if (!(x > -1.0e9 && x < 1.0e9)) continue;                     // no tap of such a sample lands in [0, numBins); keeps the rest in 32 bits
        const int m0 = (int)floor(x);
        // coarse bins the taps of fine bin m0 reach: floor((m0 -+ half) / r) with half = 2 r s a multiple of r -> ONE division
        const int q0 = floordiv32(m0, P.r_grad);
        int b0 = q0 - 2 * P.s_bin, b1 = q0 + 2 * P.s_bin;
        if (b0 < 0) b0 = 0; if (b1 > P.numBins - 1) b1 = P.numBins - 1;
        const double* drow = diff + s * P.numBins;
        if (b0 <= b1) {
          // the tap run of coarse bin b is [ilo, ihi) = clamp(b r - m0 + half + {0, r}, 0, K) (tap_span32): the upper end of one bin is
          // the lower end of the next, so every prefix-table entry is read once
          int lo = b0 * P.r_grad - m0 + half;
          int i0 = lo < 0 ? 0 : (lo > P.K ? P.K : lo);
          double w0 = s_w[i0], d0 = s_d[i0];
          for (int b = b0; b <= b1; ++b) {
            lo += P.r_grad;
            const int i1 = lo < 0 ? 0 : (lo > P.K ? P.K : lo);
            const double w1 = s_w[i1], d1 = s_d[i1];
            const double df = __ldg(drow + b);
            At += (w1 - w0) * df;
            Bt += (d1 - d0) * df;
            w0 = w1; d0 = d1;
          }
        }
        At *= -2.0; Bt *= -2.0;
      }
      if (KIND == 1) {                                                // TG.cpp:677-688
        const double g0 = (double)(ff * ff);
        gs += (double)t.st.A * (g0 * At) / (double)P.spp;
      } else if (KIND == 2) {                                         // ggx/TG.cpp:492-505
        const double g0 = (double)(alb * ff * ff * ggx_eval_adiff(P.alpha, dot3(n, -d)));
        gs += (double)t.st.A * g0 * At / (double)P.spp;
      } else {
        const float hl2 = hl * hl, hl4 = hl2 * hl2, hl5 = hl4 * hl;
        f3 t1, gn = mk3(0.f, 0.f, 0.f); float inten;
        if (!GGX) {
          inten = alb * ff * ff;                                      // TG.cpp:950
          t1 = (2 * alb * c2 * c3) * (on * c3 - n * c2 + (4 * (-d)) * c2 * c3);   // :953
          t1 = t1 * (1.0f / hl5);                                     // :954 (Embree's Vec3fa / float is a reciprocal multiply too)
          if ((HAS_VN && P.testing_flag == 0) || P.sr) {              // :959-964; SR/SSG.cpp:266-271 always
            gn = ((-2 * alb) * d) * c3 * c2 * c2; gn = gn * (1.0f / hl4);
            const float ct = dot3(gn, n); gn = gn - n * ct;
          }
        } else {                                                      // ggx/TG.cpp:756-780
          const float nw = dot3(n, -d);
          const float brdf = ggx_eval(P.alpha, nw);
          const float S = ggx_eval_xdiff(P.alpha, nw);
          const f3 dn = S * (-d), dw = S * n;
          const f3 dx = -dw + d * dot3(d, dw) / hl;                   // :759 (sic)
          inten = alb * ff * ff * brdf;
          f3 t11 = (2 * c2 * c3) * (on * c3 - n * c2 + (4 * (-d)) * c2 * c3);
          t11 = t11 * (1.0f / hl5); t11 = t11 * brdf;
          t1 = t11 + (ff * ff) * dx;
          if (HAS_VN && P.testing_flag == 0) {
            gn = (-2 * d) * c3 * c2 * c2 * brdf; gn = gn * (1.0f / hl4);
            gn = gn + (ff * ff) * dn;
            const float ct = dot3(gn, n); gn = gn - n * ct;
          }
        }
        f3 t2 = n * inten;                                            // :956
        t2 = (t2 + gn) * (1.0f / (2 * t.st.A));                       // :966
        // sum_i g_k(i) = (t1 b_k + t2 x e_k) At + grad_coef I d b_k Bt   (grad_coef = 2/sigma^2, or -2/res for the jitter kernel)
        const float fa = (float)At;
        const float fb = (float)((double)inten * P.grad_coef * Bt);
        const float sA = t.st.A;
        const f3 gk1 = (t1 * g.u + cross3(t2, e1)) * fa + d * (g.u * fb);
        const f3 gk2 = (t1 * g.v + cross3(t2, e2)) * fa + d * (g.v * fb);
        const f3 gk3 = (t1 * g.w + cross3(t2, e3)) * fa + d * (g.w * fb);
        if (P.spp == 1) {                                               // x * (1.0 / 1) == x: nine FP64 multiplies less per visible sample
          g1x += (double)(sA * gk1.x); g1y += (double)(sA * gk1.y); g1z += (double)(sA * gk1.z);
          g2x += (double)(sA * gk2.x); g2y += (double)(sA * gk2.y); g2z += (double)(sA * gk2.z);
          g3x += (double)(sA * gk3.x); g3y += (double)(sA * gk3.y); g3z += (double)(sA * gk3.z);
        } else {
          g1x += (double)(sA * gk1.x) * inv_spp; g1y += (double)(sA * gk1.y) * inv_spp; g1z += (double)(sA * gk1.z) * inv_spp;
          g2x += (double)(sA * gk2.x) * inv_spp; g2y += (double)(sA * gk2.y) * inv_spp; g2z += (double)(sA * gk2.z) * inv_spp;
          g3x += (double)(sA * gk3.x) * inv_spp; g3y += (double)(sA * gk3.y) * inv_spp; g3z += (double)(sA * gk3.z) * inv_spp;
        }
      }
    }
  }
  if (KIND == 0) {
    if (active) {
      if (g1x != 0.0 || g1y != 0.0 || g1z != 0.0) { atomicAdd(out + 3 * (size_t)t.st.i1, g1x); atomicAdd(out + 3 * (size_t)t.st.i1 + 1, g1y); atomicAdd(out + 3 * (size_t)t.st.i1 + 2, g1z); }
      if (g2x != 0.0 || g2y != 0.0 || g2z != 0.0) { atomicAdd(out + 3 * (size_t)t.st.i2, g2x); atomicAdd(out + 3 * (size_t)t.st.i2 + 1, g2y); atomicAdd(out + 3 * (size_t)t.st.i2 + 2, g2z); }
      if (g3x != 0.0 || g3y != 0.0 || g3z != 0.0) { atomicAdd(out + 3 * (size_t)t.st.i3, g3x); atomicAdd(out + 3 * (size_t)t.st.i3 + 1, g3y); atomicAdd(out + 3 * (size_t)t.st.i3 + 2, g3z); }
    }
  } else {
    // block reduction of the scalar, one atomic per block
    __shared__ double s_red[kBlock / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = gs;
    __syncthreads();
    if (threadIdx.x == 0) { double tot = 0; for (int w = 0; w < kBlock / 32; ++w) tot += s_red[w]; if (tot != 0.0) atomicAdd(out, tot); }
  }
}

// gradient[d] += acc[d] / L        (TG.cpp:561-565: '+=' into the caller's array)
__global__ void k_finalize_gradient(const double* __restrict__ acc, double* __restrict__ gradient, size_t n, double inv_L) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) gradient[i] += acc[i] * inv_L;
}

// ---------------------------------------------------------------------------------------------- debug visibility
// Pure geometry: bit = (nearest hit == sampled triangle), no shading-based skipping. vis_out[L,F,spp] in caller order.
__global__ void __launch_bounds__(kBlock) k_visibility(const DeviceScene sc, const RenderParams P, uint8_t* __restrict__ vis_out, unsigned long long* __restrict__ counters) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= sc.F) return;
  TriRegs t; load_tri<false, false>(sc, p, t);
  const int64_t s0 = (int64_t)blockIdx.y * P.chunk;
  const int64_t s1 = s0 + P.chunk < P.L ? s0 + P.chunk : P.L;
  unsigned long long nb = 0, nt = 0, nr = 0;
  for (int64_t s = s0; s < s1; ++s) {
    const f3 o = xyz(__ldg(P.origin + s));
    for (int k = 0; k < P.spp; ++k) {
      SampleGeom g; uint8_t bit = 0;
      if (draw_sample(P, P.src_offset + s, t.prim, k, o, t.st, t.tr, g)) {
        const Ray ray = make_ray(o, g.d);
        uint32_t cb = 0, ct = 0;
        bit = occluded(sc.nodes, sc.ttris, sc.root_count, ray, g.t, t.prim, &cb, &ct) ? 0 : 1;
        nb += cb; nt += ct; nr += 1;
      }
      vis_out[((size_t)s * sc.F + t.prim) * P.spp + k] = bit;
    }
  }
  if (counters) { atomicAdd(counters, nr); atomicAdd(counters + 1, nb); atomicAdd(counters + 2, nt); }
}

// ---- jitter/ temporal kernel
// forward (jitter/TG.cpp:331-350): T[s,b] = sum_i w[i] * H[s, b + offset - i]
__global__ void k_jitter_conv(const double* __restrict__ H, const double* __restrict__ w, int J, int off, int B, size_t n, double* __restrict__ T) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / B; const int b = (int)(i - s * B);
    double acc = 0.0;
    for (int k = 0; k < J; ++k) { const int m = b + off - k; if (m >= 0 && m < B) acc += w[k] * H[s * B + m]; }
    T[i] = acc;
  }
}
// gradient tables: jA[s,b] = sum_i w[i] (-2) diff[s, b+i-off], jB likewise with the kernel derivative; b in [0,B]
__global__ void k_jitter_tables(const double* __restrict__ diff, const double* __restrict__ w, const double* __restrict__ g, int J, int off, int B, size_t n,
                                double* __restrict__ jA, double* __restrict__ jB) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t s = i / (B + 1); const int b = (int)(i - s * (B + 1));
    double a = 0.0, c = 0.0;
    for (int k = 0; k < J; ++k) { const int m = b + k - off; if (m >= 0 && m < B) { const double d = -2.0 * diff[s * B + m]; a += w[k] * d; c += g[k] * d; } }
    jA[i] = a; jB[i] = c;
  }
}

__global__ void k_pack4(const float* __restrict__ in, float4* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.f);
}
__global__ void k_pathlengths(double* __restrict__ pl, int B, float lb, float res) {   // SST.cpp:126-129
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) pl[i] = (double)(lb + i * res);
}

inline dim3 sample_grid(const DeviceScene& sc, const RenderParams& P) {
  return dim3((unsigned)((sc.F + kBlock - 1) / kBlock), (unsigned)((P.L + P.chunk - 1) / P.chunk), 1);
}

// shared memory of k_forward_grid for a G x G grid
inline size_t grid_smem_bytes(int G, int Kz, bool smooth, int K) {
  return ((sizeof(GridShared) + 15) & ~size_t(15)) + (smooth ? (((size_t)(K + 1) * sizeof(double) + 15) & ~size_t(15)) : 0) +
         (size_t)(kGridBlock / 32) * sizeof(GridWarp) + (size_t)G * G * Kz * sizeof(unsigned);
}
// perspective-grid forward kernel: applies outside the first-generation mode (its unclamped form factor traces rays behind the wall point)
inline bool use_grid_forward(const Ctx& cx, const DeviceScene& sc, const RenderParams& P) {
  if (cx.forward_algo == 1 || P.sr || sc.F < 1 || sc.F >= (1 << 27) || sc.V < 1 || sc.bounds == nullptr || sc.verts == nullptr) return false;   // work items pack the triangle index into 27 bits
  if (cx.forward_algo == 2) return true;
  // auto: one block per wall point needs enough wall points to fill the machine; below that the BVH kernel (finer work items) is used
  return P.L >= (cx.num_sms > 0 ? cx.num_sms : 148);
}
template <bool GGX, bool VN, bool VA, bool SMOOTH, int MODE>
void launch_forward_grid_t(Ctx& cx, const DeviceScene& sc, const RenderParams& P, double* out, uint32_t* vis, const double* wprefix) {
  const int sms = cx.num_sms > 0 ? cx.num_sms : 148;
  const size_t budget = (size_t)(NLOS_GRID_MINBLOCKS >= 2 ? 113 : 226) * 1024;          // dynamic shared memory per block at the wanted residency
  int G = cx.grid_res > 0 ? cx.grid_res : (int)(std::sqrt((double)sc.F * std::min(P.spp, 16)) * 0.25 + 0.5);   // ~16 samples per picture cell (measured optima: C-bunny G = 64, C-arm G = 16..32)
  if (G < 1) G = 1;
  if (G > 256) G = 256;                                                                  // quantised coordinates are 15-bit
  // depth slices: kGridK (4) while the wanted resolution fits shared memory with them, else fewer slices for a finer picture (the cell
  // counters are G*G*Kz words).  Measured on the C-scale mesh (F = 500 000, wanted G = 177): Kz = 4 / G = 105 207 ms, 2 / 148 185 ms, 1 / 177 156 ms
  int Kz = cx.grid_slices > 0 ? std::min(cx.grid_slices, kGridK) : kGridK;
  if (cx.grid_slices <= 0) while (Kz > 1 && grid_smem_bytes(G, Kz, SMOOTH, P.K) > budget) Kz >>= 1;
  while (G > 1 && grid_smem_bytes(G, Kz, SMOOTH, P.K) > budget) --G;
  const size_t smem = grid_smem_bytes(G, Kz, SMOOTH, P.K);
  const int blocks = (int)std::min<int64_t>(P.L, (int64_t)sms * NLOS_GRID_MINBLOCKS);
  const unsigned cap = (unsigned)(std::min<int64_t>((int64_t)6 * sc.F + 4 * (int64_t)G * G * Kz + 1024, 0x7fffff0) & ~(int64_t)3);     // entry positions are 27-bit, lists 16-byte aligned
  unsigned capv = cap;
  if (cx.grid_cap > 0) capv = (unsigned)std::max<int64_t>(std::min<int64_t>(cap, cx.grid_cap), ((int64_t)sc.F + 3 + 4 * Kz) & ~(int64_t)3);   // never below one coarsest-grid fill
  GridScratch scr;
  scr.proj = cx.buf("grid_proj").as<float4>((size_t)blocks * sc.V);
  scr.rect = cx.buf("grid_rect").as<uint2>((size_t)blocks * sc.F);
  scr.ent = cx.buf("grid_ent").as<unsigned>(2 * (size_t)blocks * cap);
  cx.last_forward_algo = 2; cx.last_grid_res = G;
  scr.counters = nullptr;
  // work counters (option "count_work"): a separate instantiation, only for the Lambertian face-normal transient kernels (the headline)
  constexpr bool kCanCount = !GGX && !VN && !VA && MODE == 0;
  const bool count = kCanCount && cx.count_work != 0;
  if (cx.count_work) {
    scr.counters = cx.buf("work_counters").as<unsigned long long>(8);
    NLOS_CUDA_OK(cudaMemsetAsync(scr.counters, 0, 8 * sizeof(unsigned long long), cx.stream));
    cx.work_G = count ? G : 0;
  }
#define NLOS_GRID_LAUNCH(WV, CNT)                                                                                                              \
  do {                                                                                                                                         \
    if (Kz == kGridK) {                                                                                                                        \
      NLOS_CUDA_OK(cudaFuncSetAttribute(k_forward_grid<GGX, VN, VA, SMOOTH, WV, MODE, CNT, kGridK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_forward_grid<GGX, VN, VA, SMOOTH, WV, MODE, CNT, kGridK><<<blocks, kGridBlock, smem, cx.stream>>>(sc, P, out, vis, wprefix, scr, capv, G, Kz);   \
    } else {                                                                                                                                   \
      NLOS_CUDA_OK(cudaFuncSetAttribute(k_forward_grid<GGX, VN, VA, SMOOTH, WV, MODE, CNT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_forward_grid<GGX, VN, VA, SMOOTH, WV, MODE, CNT, 0><<<blocks, kGridBlock, smem, cx.stream>>>(sc, P, out, vis, wprefix, scr, capv, G, Kz);        \
    }                                                                                                                                          \
  } while (0)
  if (count) {
    if constexpr (kCanCount) { if (vis) NLOS_GRID_LAUNCH(true, true); else NLOS_GRID_LAUNCH(false, true); }
  } else {
    if (vis) NLOS_GRID_LAUNCH(true, false); else NLOS_GRID_LAUNCH(false, false);
  }
#undef NLOS_GRID_LAUNCH
  cx.launches += 1;
}

// shared-grid forward kernel: applies outside the first-generation mode, to any number of wall points (the unit of work is a warp)
inline bool use_group_forward(const Ctx& cx, const DeviceScene& sc, const RenderParams& P) {
  if (P.sr || sc.F < 1 || sc.F >= (1 << 27) || sc.V < 1 || P.L < 1 || P.L > 0x7fffffff || sc.bounds == nullptr || sc.verts == nullptr) return false;
  return cx.forward_algo == 3;          // on request only: measured slower than the per-wall-point grid on every BASELINE config
}
template <bool GGX, bool VN, bool VA, bool SMOOTH, int MODE>
void launch_forward_group_t(Ctx& cx, const DeviceScene& sc, const RenderParams& P, double* out, uint32_t* vis, const double* wprefix) {
  const int sms = cx.num_sms > 0 ? cx.num_sms : 148;
  const int side = cx.group_side > 0 ? cx.group_side : 4;
  const int n_groups = make_wall_groups(cx, P, side);
  const int K = cx.grid_slices > 0 ? std::min(cx.grid_slices, 64) : 16;
  int G = cx.grid_res > 0 ? cx.grid_res : (int)(std::sqrt((double)sc.F * std::min(P.spp, 16)) * 0.25 + 0.5);
  G = std::max(1, std::min(G, 256));                                                       // quantised coordinates are 15-bit
  while (G > 1 && (size_t)G * G * K > ((size_t)1 << 24)) --G;
  const size_t tab = (size_t)G * G * K;
  unsigned cap = (unsigned)(std::min<int64_t>((int64_t)10 * sc.F + 4 * (int64_t)tab + 1024, 0x7ffffff0) & ~(int64_t)3);
  if (cx.grid_cap > 0) cap = (unsigned)std::max<int64_t>(std::min<int64_t>(cap, cx.grid_cap), ((int64_t)sc.F * K + 3 + 4 * K) & ~(int64_t)3);   // never below one coarsest-grid fill
  const size_t per_group = (size_t)cap * 8 + tab * 8 + (size_t)sc.F * 4 + sizeof(GroupHdr);
  const size_t budget = (size_t)(cx.grid_budget_mb > 0 ? cx.grid_budget_mb : 6144) << 20;
  const int per_batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_groups, budget / per_group));
  cx.last_forward_algo = 3; cx.last_grid_res = G;
  unsigned long long* counters = nullptr;
  constexpr bool kCanCount = !GGX && !VN && !VA && MODE == 0;
  const bool count = kCanCount && cx.count_work != 0;
  if (cx.count_work) {
    counters = cx.buf("work_counters").as<unsigned long long>(8);
    NLOS_CUDA_OK(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned long long), cx.stream));
    cx.work_G = count ? G : 0;
  }
  if (vis) NLOS_CUDA_OK(cudaMemsetAsync(vis, 0, (size_t)P.L * P.spp * P.words_per_row * sizeof(uint32_t), cx.stream));      // visibility bits are OR-ed in
  // triangles per work item: 64 batches of 32 (the rays left in the queue at the end of an item are traced with idle lanes), fewer when that would leave warps without work
  const int64_t warps = (int64_t)sms * 4 * (kGrpBlock / 32);
  int cb = 64;
  while (cb > 1 && P.L * (((int64_t)sc.F + 32 * cb - 1) / (32 * cb)) < 4 * warps) cb >>= 1;
  const size_t smem = (size_t)(kGrpBlock / 32) * sizeof(GrpWarp) + (SMOOTH ? (((size_t)(P.K + 1) * sizeof(double) + 15) & ~size_t(15)) : 0);
  for (int g0 = 0; g0 < n_groups; g0 += per_batch) {
    const int ng = std::min(per_batch, n_groups - g0);
    GroupGrid gg;
    bin_wall_groups(cx, sc, P, g0, ng, n_groups, G, K, cap, !VN, gg);
    gg.gstart = cx.buf("wg_start").as<int>((size_t)P.L + 1); gg.ngroups = ng;
    const int blocks = sms * 4;
#define NLOS_GROUP_LAUNCH(WV, CNT)                                                                                                              \
  do {                                                                                                                                         \
    NLOS_CUDA_OK(cudaFuncSetAttribute(k_forward_group<GGX, VN, VA, SMOOTH, WV, MODE, CNT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_forward_group<GGX, VN, VA, SMOOTH, WV, MODE, CNT><<<blocks, kGrpBlock, smem, cx.stream>>>(sc, P, out, vis, wprefix, gg, cb, counters);           \
  } while (0)
    if (count) {
      if constexpr (kCanCount) { if (vis) NLOS_GROUP_LAUNCH(true, true); else NLOS_GROUP_LAUNCH(false, true); }
    } else {
      if (vis) NLOS_GROUP_LAUNCH(true, false); else NLOS_GROUP_LAUNCH(false, false);
    }
#undef NLOS_GROUP_LAUNCH
    cx.launches += 1;
  }
}

template <bool GGX, bool VN, bool VA, bool SMOOTH, int MODE>
void launch_forward_t(Ctx& cx, const DeviceScene& sc, const RenderParams& P, double* out, uint32_t* vis, const double* wprefix) {
#ifdef NLOS_QUICK_BUILD      // kernel experiments only (tools/build_variant.sh): the headline instantiation of the per-point grid kernel and nothing else
  launch_forward_grid_t<GGX, VN, VA, SMOOTH, MODE>(cx, sc, P, out, vis, wprefix); return;
#else
  if (use_group_forward(cx, sc, P)) { launch_forward_group_t<GGX, VN, VA, SMOOTH, MODE>(cx, sc, P, out, vis, wprefix); return; }
  if (use_grid_forward(cx, sc, P)) { launch_forward_grid_t<GGX, VN, VA, SMOOTH, MODE>(cx, sc, P, out, vis, wprefix); return; }
  cx.last_forward_algo = 1; cx.last_grid_res = 0;
  const int64_t nchunks = (P.L * (int64_t)P.spp + P.chunk - 1) / P.chunk;
  const dim3 grid((unsigned)((sc.F + 31) / 32), (unsigned)std::min<int64_t>((nchunks + kFwdBlock / 32 - 1) / (kFwdBlock / 32), 65535), 1);
  const size_t smem = (kFwdBlock / 32) * sizeof(WarpShared) + (SMOOTH ? (size_t)(P.K + 1) * sizeof(double) : 0);
  if (smem > 48 * 1024) {      // long tap tables (large refine_scale * sigma_bin) need the opt-in shared-memory limit
    NLOS_CUDA_OK(cudaFuncSetAttribute(k_forward<GGX, VN, VA, SMOOTH, true, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NLOS_CUDA_OK(cudaFuncSetAttribute(k_forward<GGX, VN, VA, SMOOTH, false, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (vis) k_forward<GGX, VN, VA, SMOOTH, true, MODE><<<grid, kFwdBlock, smem, cx.stream>>>(sc, P, out, vis, wprefix);
  else k_forward<GGX, VN, VA, SMOOTH, false, MODE><<<grid, kFwdBlock, smem, cx.stream>>>(sc, P, out, vis, wprefix);
  cx.launches += 1;
#endif
}

template <bool GGX, bool VN, bool VA, int KIND>
void launch_gradient_t(Ctx& cx, const DeviceScene& sc, const RenderParams& P, const double* diff, const uint32_t* vis,
                       const double* wprefix, const double* dprefix, double* out) {
  const dim3 grid = sample_grid(sc, P);
  const size_t smem = 2 * (size_t)(P.K + 1) * sizeof(double);
  if (smem > 48 * 1024) {      // long tap tables (refine_scale * sigma_bin >= ~768) need the opt-in shared-memory limit, as the forward kernel does
    NLOS_CUDA_OK(cudaFuncSetAttribute(k_gradient<GGX, VN, VA, KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NLOS_CUDA_OK(cudaFuncSetAttribute(k_gradient<GGX, VN, VA, KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (vis) k_gradient<GGX, VN, VA, KIND, true><<<grid, kBlock, smem, cx.stream>>>(sc, P, diff, vis, wprefix, dprefix, out);
  else k_gradient<GGX, VN, VA, KIND, false><<<grid, kBlock, smem, cx.stream>>>(sc, P, diff, vis, wprefix, dprefix, out);
  cx.launches += 1;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ launchers
void launch_forward(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* transient, uint32_t* vis, const double* wprefix) {
  if (sc.F <= 0 || P.L <= 0) return;
  const bool vn = sc.vnormal != nullptr, va = sc.valbedo != nullptr, sm = P.r_fwd > 1;
#define NLOS_FWD(G, N, A, S) launch_forward_t<G, N, A, S, 0>(cx, sc, P, transient, vis, wprefix)
#ifdef NLOS_QUICK_BUILD
  if (ggx || vn || va || sm) std::abort();
  NLOS_FWD(false, false, false, false);
#else
  if (!ggx) {
    if (!vn && !va) { if (sm) NLOS_FWD(false, false, false, true); else NLOS_FWD(false, false, false, false); }
    else if (vn && !va) { if (sm) NLOS_FWD(false, true, false, true); else NLOS_FWD(false, true, false, false); }
    else if (!vn && va) { if (sm) NLOS_FWD(false, false, true, true); else NLOS_FWD(false, false, true, false); }
    else { if (sm) NLOS_FWD(false, true, true, true); else NLOS_FWD(false, true, true, false); }
  } else {
    if (!vn && !va) { if (sm) NLOS_FWD(true, false, false, true); else NLOS_FWD(true, false, false, false); }
    else if (vn && !va) { if (sm) NLOS_FWD(true, true, false, true); else NLOS_FWD(true, true, false, false); }
    else if (!vn && va) { if (sm) NLOS_FWD(true, false, true, true); else NLOS_FWD(true, false, true, false); }
    else { if (sm) NLOS_FWD(true, true, true, true); else NLOS_FWD(true, true, true, false); }
  }
#endif
#undef NLOS_FWD
  NLOS_CUDA_OK(cudaGetLastError());
}

int forward_max_chunk() { return kTile; }      // sample slots per warp pass of k_forward == words of its visibility tile

void launch_intensity(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* intensity) {
  if (sc.F <= 0 || P.L <= 0) return;
  const bool vn = sc.vnormal != nullptr;
#ifdef NLOS_QUICK_BUILD
  std::abort();
#else
  if (!ggx) { if (vn) launch_forward_t<false, true, false, false, 1>(cx, sc, P, intensity, nullptr, nullptr); else launch_forward_t<false, false, false, false, 1>(cx, sc, P, intensity, nullptr, nullptr); }
  else { if (vn) launch_forward_t<true, true, false, false, 1>(cx, sc, P, intensity, nullptr, nullptr); else launch_forward_t<true, false, false, false, 1>(cx, sc, P, intensity, nullptr, nullptr); }
#endif
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_residual(Ctx& cx, const double* data, const double* weight, const double* T, double* diff, size_t n, int loss_flag) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  k_residual<<<blocks, 256, 0, cx.stream>>>(data, weight, T, diff, n, loss_flag);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_box_filter(Ctx& cx, const double* in, double* out, int B, int64_t L, int width) {
  const size_t n = (size_t)L * B; if (n == 0) return;
  k_box_filter<<<(int)std::min<size_t>((n + 255) / 256, 148 * 32), 256, 0, cx.stream>>>(in, out, B, n, width);
  cx.launches += 1; NLOS_CUDA_OK(cudaGetLastError());
}

void launch_gradient(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, int kind, const double* diff, const uint32_t* vis,
                     const double* wprefix, const double* dprefix, double* out) {
  if (sc.F <= 0 || P.L <= 0) return;
  const bool vn = sc.vnormal != nullptr, va = sc.valbedo != nullptr;
#define NLOS_GRAD(G, N, A, KD) launch_gradient_t<G, N, A, KD>(cx, sc, P, diff, vis, wprefix, dprefix, out)
#ifdef NLOS_QUICK_BUILD
  if (kind != 0 || ggx || vn || va) std::abort();
  NLOS_GRAD(false, false, false, 0);
#else
  if (kind == 0) {
    if (!ggx) {
      if (!vn && !va) NLOS_GRAD(false, false, false, 0); else if (vn && !va) NLOS_GRAD(false, true, false, 0);
      else if (!vn && va) NLOS_GRAD(false, false, true, 0); else NLOS_GRAD(false, true, true, 0);
    } else {
      if (!vn && !va) NLOS_GRAD(true, false, false, 0); else if (vn && !va) NLOS_GRAD(true, true, false, 0);
      else if (!vn && va) NLOS_GRAD(true, false, true, 0); else NLOS_GRAD(true, true, true, 0);
    }
  } else if (kind == 1) {
    if (!vn && !va) NLOS_GRAD(false, false, false, 1); else if (vn && !va) NLOS_GRAD(false, true, false, 1);
    else if (!vn && va) NLOS_GRAD(false, false, true, 1); else NLOS_GRAD(false, true, true, 1);
  } else {
    if (!vn && !va) NLOS_GRAD(true, false, false, 2); else if (vn && !va) NLOS_GRAD(true, true, false, 2);
    else if (!vn && va) NLOS_GRAD(true, false, true, 2); else NLOS_GRAD(true, true, true, 2);
  }
#endif
#undef NLOS_GRAD
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_finalize_gradient(Ctx& cx, const double* acc, double* gradient, size_t n, double inv_L) {
  if (n == 0) return;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  k_finalize_gradient<<<blocks, 256, 0, cx.stream>>>(acc, gradient, n, inv_L);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_visibility(Ctx& cx, const DeviceScene& sc, const RenderParams& P, uint8_t* vis_out, unsigned long long* counters) {
  if (sc.F <= 0 || P.L <= 0) return;
  k_visibility<<<sample_grid(sc, P), kBlock, 0, cx.stream>>>(sc, P, vis_out, counters);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_jitter_conv(Ctx& cx, const double* H, const double* w, int J, int off, int B, int64_t L, double* T) {
  const size_t n = (size_t)L * B; if (n == 0) return;
  k_jitter_conv<<<(int)std::min<size_t>((n + 255) / 256, 148 * 32), 256, 0, cx.stream>>>(H, w, J, off, B, n, T);
  cx.launches += 1; NLOS_CUDA_OK(cudaGetLastError());
}
void launch_jitter_tables(Ctx& cx, const double* diff, const double* w, const double* g, int J, int off, int B, int64_t L, double* jA, double* jB) {
  const size_t n = (size_t)L * (B + 1); if (n == 0) return;
  k_jitter_tables<<<(int)std::min<size_t>((n + 255) / 256, 148 * 32), 256, 0, cx.stream>>>(diff, w, g, J, off, B, n, jA, jB);
  cx.launches += 1; NLOS_CUDA_OK(cudaGetLastError());
}

void launch_pack4(Ctx& cx, const float* in, float4* out, size_t n) {
  if (n == 0) return;
  k_pack4<<<(int)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, cx.stream>>>(in, out, n);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_pathlengths(Ctx& cx, double* pl, int B, float lb, float res) {
  if (B <= 0) return;
  k_pathlengths<<<(B + 255) / 256, 256, 0, cx.stream>>>(pl, B, lb, res);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

#ifdef NLOS_EXT_BUILD
}  // namespace ext
#endif
}  // namespace nlos
