// group_grid.h — the shared perspective grid of a GROUP of wall points (DESIGN.md "K1s"): data model and host entry points of
// group_grid.cu.  The wall points of a call are sorted into groups of spatial neighbours; the mesh is binned once per group
// (k_group_bin) and the forward sample kernel (render_kernels.cu, k_forward_group) serves every wall point of the group from it.
#pragma once
#include "nlos_ctx.h"

namespace nlos {

struct GroupHdr {             // one per group of the current batch
  GGFrame fr;                 // fr.G == 0: no grid for this group (its rays take the per-ray BVH query)
  int member0, nmember;       // the group's wall points: positions [member0, member0 + nmember) of the sorted wall order
  unsigned long long ent0;    // first entry of the group in entE / entI
  unsigned long long tab0;    // first (cell, slice) record of the group in table
  unsigned long long live0;   // first word of the group's list of live triangles in live
  int nlive;                  // triangles that can contribute to at least one wall point of the group (the others are culled by the plane-side test of every member)
  int pad_;
};

struct GroupGrid {            // passed by value to the forward kernel
  const GroupHdr* hdr;        // [groups of the batch]
  const int* order;           // [L]   sorted position -> wall point (index into RenderParams::origin)
  const int* group_of;        // [L]   sorted position -> group (global index)
  const uint2* table;         // per group G*G*K records (first entry relative to ent0, count), index (cy*G + cx)*K + k
  const unsigned* ent;        // entries in blocks of 4 (32 bytes): [E0 E1 E2 E3][T0 T1 T2 T3], E = rectangle word (pg_entry), T = triangle (Morton
                              // index) | fine depth index << 27; every list starts on a block and is padded with never-matching words (E = 0)
  const int* live;            // per group: Morton indices of its live triangles, ascending
  const int* gstart;          // [groups + 1] first sorted position of every group
  int group0, ngroups;        // the batch: groups [group0, group0 + ngroups), hdr[0] describes group0
};

// Sorts the L wall points of P into groups of at most side*side spatial neighbours (tiles of side x side wall spacings).
// Fills cx.buf("wg_order"), ("wg_group_of"), ("wg_start") and returns the number of groups (one small D2H copy: synchronises cx.stream).
int make_wall_groups(Ctx& cx, const RenderParams& P, int side);

// Bins the mesh for groups [g0, g0 + ng) into the context's scratch buffers (asynchronous on cx.stream) and describes them in 'out'.
// cull: the forward kernel applies the plane-side cull (face-normal shading, second-generation renderer), so triangles that every
// member of a group culls are left out of the group's live list.
void bin_wall_groups(Ctx& cx, const DeviceScene& sc, const RenderParams& P, int g0, int ng, int n_groups, int G, int K, unsigned cap, bool cull, GroupGrid& out);

}  // namespace nlos
