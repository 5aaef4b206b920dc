// render_kernels.h — host-callable launchers of render_kernels.cu (all asynchronous on cx.stream).
#pragma once
#include "nlos_ctx.h"

namespace nlos {

void launch_forward(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* transient, uint32_t* vis, const double* wprefix);
void launch_intensity(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* intensity);
int forward_max_chunk();   // upper bound of RenderParams::chunk for the BVH forward kernel (its shared-memory visibility tile)
void launch_residual(Ctx& cx, const double* data, const double* weight, const double* T, double* diff, size_t n, int loss_flag);
void launch_gradient(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, int kind, const double* diff, const uint32_t* vis,
                     const double* wprefix, const double* dprefix, double* out);
void launch_finalize_gradient(Ctx& cx, const double* acc, double* gradient, size_t n, double inv_L);
void launch_visibility(Ctx& cx, const DeviceScene& sc, const RenderParams& P, uint8_t* vis_out, unsigned long long* counters);
void launch_pack4(Ctx& cx, const float* in, float4* out, size_t n);
void launch_pathlengths(Ctx& cx, double* pl, int B, float lb, float res);
void launch_box_filter(Ctx& cx, const double* in, double* out, int B, int64_t L, int width);
void launch_jitter_conv(Ctx& cx, const double* H, const double* w, int J, int off, int B, int64_t L, double* T);
void launch_jitter_tables(Ctx& cx, const double* diff, const double* w, const double* g, int J, int off, int B, int64_t L, double* jA, double* jB);
// mesh_kernels.cu
void launch_vertex_gradient(Ctx& cx, const DeviceScene& sc, const RenderParams& P, int vertex_num, const double* taps, double sigma2, double* acc);
void launch_ray_query(Ctx& cx, const DeviceScene& sc, int mode, const float* origins, const float* dirs, int64_t N, float* out);
void launch_bary_to_world(Ctx& cx, const float* verts, const int* faces, const float* bary, int64_t N, float* out);
void launch_regulariser(Ctx& cx, int mode, const float* verts, int V, const int* faces, int F, const int* aff, double* grad, double* value);

namespace ext {   // render_kernels_ext.cu: the launchers of the sample kernels built with the external-sample test hook
void launch_forward(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* transient, uint32_t* vis, const double* wprefix);
void launch_intensity(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, double* intensity);
void launch_gradient(Ctx& cx, const DeviceScene& sc, const RenderParams& P, bool ggx, int kind, const double* diff, const uint32_t* vis,
                     const double* wprefix, const double* dprefix, double* out);
}  // namespace ext

}  // namespace nlos
