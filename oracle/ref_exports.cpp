// TEST INFRASTRUCTURE.  extern "C" forwarding stubs over the reference's own (C++-linkage) entry points, so that tests can call the
// reference's unmodified translation units (compiled by `make -C oracle ref` from /root/reference against oracle/ref_shim) via ctypes.
// Compiled once per reference module: -DREF_RENDERER (smoothed_transient/), -DREF_GGX (ggx/), -DREF_JITTER (jitter/),
// -DREF_INTERSECTOR (embree_intersector/), -DREF_SR (stratified_transient_raytracer/, the first-generation renderer).  The reference headers are found through -I<reference module dir>.
#if defined(REF_INTERSECTOR)
#include "c_embree_intersector.h"
extern "C" {
void ref_embree3_tbb_line_intersection(float* o, float* d, int n, float* v, int nv, int* f, int nf, float* out) { embree3_tbb_line_intersection(o, d, n, v, nv, f, nf, out); }
void ref_embree3_tbb_short_line_intersection(float* o, float* d, int n, float* v, int nv, int* f, int nf, float* out) { embree3_tbb_short_line_intersection(o, d, n, v, nv, f, nf, out); }
void ref_barycentric_to_world(float* v, int* f, float* bary, int n, float* world) { barycentric_to_world(v, f, bary, n, world); }
}
#elif defined(REF_SR)
#include "stratifiedTransientRenderer.h"
#include "stratifiedStreamedTransientRenderer.h"
#include "stratifiedStreamedGradientRenderer.h"
extern "C" {
void ref_sr_render_transient(float* o, float* n, float* v, int V, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl) {
    render_transient(o, n, v, V, f, F, S, lb, ub, res, T, pl); }
void ref_sr_streamed_render_transient(float* o, int L, float* n, float* v, int V, float* vn, float* va, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl) {
    streamed_render_transient(o, L, n, v, V, vn, va, f, F, S, lb, ub, res, T, pl); }
void ref_sr_streamed_render_gradient(double* data, float* o, int L, float* n, float* v, int V, int* f, int F, int S, float lb, float ub, float res, int w, double* T, double* pl, double* g) {
    streamed_render_gradient(data, o, L, n, v, V, f, F, S, lb, ub, res, w, T, pl, g); }
void ref_sr_streamed_render_curvature_grad(float* v, int V, int* f, int F, double* g) { streamed_render_curvature_grad(v, V, f, F, g); }
}
#else
#include "stratifiedStreamedTransientRenderer.h"
#include "stratifiedStreamedGradientRenderer.h"
#include <cstddef>
#include <vector>
#include "sampler.h"
extern "C" {
#if defined(REF_RENDERER)
// the (S,T) stream a single-threaded reference run consumes: worker 0 of a default-constructed SamplerSet (sampler.cpp:20-33)
void ref_sampler_stream(int n, float* out) { smp::SamplerSet set(1); for (int i = 0; i < n; ++i) out[i] = set[0](); }
void ref_streamed_render_intensity(float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, int S, float lb, float ub, double* out) {
    streamed_render_intensity(o, L, n, v, V, vn, f, F, S, lb, ub, out); }
void ref_streamed_render_transient(float* o, int L, float* n, float* v, int V, float* vn, float* va, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl, int rs, int sb) {
    streamed_render_transient(o, L, n, v, V, vn, va, f, F, S, lb, ub, res, T, pl, rs, sb); }
void ref_streamed_render_gradient(double* data, double* w, float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl,
                                  double* g, int rs, int sb, int testing, int loss) {
    streamed_render_gradient(data, w, o, L, n, v, V, vn, f, F, S, lb, ub, res, T, pl, g, rs, sb, testing, loss); }
void ref_streamed_render_gradient_w_albedo(double* data, double* w, float* o, int L, float* n, float* v, int V, float* al, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl,
                                           double* g, int rs, int sb, int testing, int loss) {
    streamed_render_gradient_w_albedo(data, w, o, L, n, v, V, al, f, F, S, lb, ub, res, T, pl, g, rs, sb, testing, loss); }
double ref_streamed_render_gradient_albedo(double* data, double* w, float* o, int L, float* n, float* v, int V, float* al, int* f, int F, int S, float lb, float ub, float res, double* T, double* pl,
                                           int rs, int sb, int testing, int loss) {
    return streamed_render_gradient_albedo(data, w, o, L, n, v, V, al, f, F, S, lb, ub, res, T, pl, rs, sb, testing, loss); }
void ref_streamed_render_vertex_gradient(int vertex, float* o, int L, float* n, float* v, int V, int* f, int F, int S, float lb, float ub, float res, double* g, int rs, int sb) {
    streamed_render_vertex_gradient(vertex, o, L, n, v, V, f, F, S, lb, ub, res, g, rs, sb); }
double ref_streamed_render_normal_smoothing(float* v, int V, int* f, int F, int* aff, double* g) { return streamed_render_normal_smoothing(v, V, f, F, aff, g); }
void ref_streamed_render_curvature_grad(float* v, int V, int* f, int F, double* g) { streamed_render_curvature_grad(v, V, f, F, g); }
#elif defined(REF_GGX)
void ref_ggx_streamed_render_intensity(float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, float alpha, int S, float lb, float ub, double* out) {
    streamed_render_intensity(o, L, n, v, V, vn, f, F, alpha, S, lb, ub, out); }
void ref_ggx_streamed_render_transient(float* o, int L, float* n, float* v, int V, float* vn, float* va, int* f, int F, float alpha, int S, float lb, float ub, float res, double* T, double* pl, int rs, int sb) {
    streamed_render_transient(o, L, n, v, V, vn, va, f, F, alpha, S, lb, ub, res, T, pl, rs, sb); }
void ref_ggx_streamed_render_gradient(double* data, double* w, float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, float alpha, int S, float lb, float ub, float res, double* T,
                                      double* pl, double* g, int rs, int sb, int testing) {
    streamed_render_gradient(data, w, o, L, n, v, V, vn, f, F, alpha, S, lb, ub, res, T, pl, g, rs, sb, testing); }
double ref_ggx_streamed_render_gradient_alpha(double* data, double* w, float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, float alpha, int S, float lb, float ub, float res,
                                              double* T, double* pl, int rs, int sb) {
    return streamed_render_gradient_alpha(data, w, o, L, n, v, V, vn, f, F, alpha, S, lb, ub, res, T, pl, rs, sb); }
#elif defined(REF_JITTER)
void ref_jitter_streamed_render_transient(float* o, int L, float* n, float* v, int V, float* vn, float* va, int* f, int F, int S, float lb, float ub, float res, double* jw, int off, int len,
                                          double* T, double* pl) {
    streamed_render_transient(o, L, n, v, V, vn, va, f, F, S, lb, ub, res, jw, off, len, T, pl); }
void ref_jitter_streamed_render_gradient(double* data, double* w, float* o, int L, float* n, float* v, int V, float* vn, int* f, int F, int S, float lb, float ub, float res, double* jw, double* jg,
                                         int off, int len, double* T, double* pl, double* g, int testing) {
    streamed_render_gradient(data, w, o, L, n, v, V, vn, f, F, S, lb, ub, res, jw, jg, off, len, T, pl, g, testing); }
#endif
}
#endif
