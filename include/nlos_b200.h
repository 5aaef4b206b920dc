/* nlos_b200.h — C ABI of libnlos_b200.so, the B200 (sm_100a) drop-in for the reference's differentiable
 * confocal transient renderer.
 *
 * Each nlos_streamed_* entry point replaces one C++ function that the reference's Cython modules
 * `renderer` (smoothed_transient/renderer.pyx) and `ggx` (ggx/ggx.pyx) bind.  Argument order and meaning
 * are the reference's, with three additions: a leading context handle, an `int` status return (0 = ok;
 * the reference returns void and only printf()s on failure), and an explicit trailing `numBins`
 * (the reference recomputes ceil((ub-lb)/res) in float on both sides of the boundary; passing it removes
 * the row-misalignment quirk of SURVEY.md A.6).  Scalar results come back through an out-pointer.
 *
 * Citations are relative to /root/reference/transient_rendering_cython/.
 *
 * Pointers: every array argument may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 * context's GPU, independently per argument (detected with cudaPointerGetAttributes).  Host arrays are
 * staged to HBM inside the call and results copied back before it returns; device arrays are used in
 * place and the call returns after enqueueing on the context stream (use nlos_ctx_synchronize).
 * Layouts are the reference's: C-contiguous, float = f32, int = i32, double = f64;
 * `transient` [L,B] is overwritten, `gradient` [V,3] is accumulated into ("+=", TG.cpp:563).
 *
 * Sampling: the two uniforms of sample k of (source s, triangle f) are
 * Philox4x32-10(key = seed, counter = (f, s_global, k>>1, 0)) -> reference bit trick (rng_sse.h:33-42);
 * the default seed is 5489, the reference's built-in boost::mt19937 default (sampler.cpp:25).
 *
 * There is no CPU fallback: every entry point fails with NLOS_ERR_CUDA if no sm_100-class device is usable.
 */
#ifndef NLOS_B200_H_
#define NLOS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nlos_ctx nlos_ctx;

enum {
  NLOS_OK = 0,
  NLOS_ERR_INVALID = 1,   /* bad argument (null pointer, non-positive size, numBins mismatch ...) */
  NLOS_ERR_CUDA = 2,      /* CUDA runtime error; text in nlos_last_error() */
  NLOS_ERR_NOMEM = 3      /* device allocation failed (reference: printf("memory insufficient"), TG.cpp:302-305) */
};

/* ---- context ------------------------------------------------------------------------------------- */
int nlos_ctx_create(int device, nlos_ctx** out);
void nlos_ctx_destroy(nlos_ctx* ctx);
const char* nlos_last_error(nlos_ctx* ctx);            /* ctx may be NULL: error of the last failed create */
int nlos_ctx_synchronize(nlos_ctx* ctx);
void* nlos_ctx_stream(nlos_ctx* ctx);                  /* the cudaStream_t all work is enqueued on */
int nlos_ctx_set_seed(nlos_ctx* ctx, uint64_t seed);
/* Sharded runs (one process per GPU): this call's sources are the global sources
 * [src_offset, src_offset + numSources); gradients are normalised by num_sources_global (0 = by numSources). */
int nlos_ctx_set_source_window(nlos_ctx* ctx, int64_t src_offset, int64_t num_sources_global);
/* Ordering against a caller's CUDA stream (only matters for DEVICE-pointer arguments, which are used in place on the context's own
 * non-blocking stream): nlos_ctx_wait_stream makes all later work of the context wait for everything enqueued on `stream` so far
 * (inputs written by the caller's kernels are seen); nlos_ctx_signal_stream makes `stream` wait for everything the context has
 * enqueued so far (the caller's later kernels see the outputs).  `stream` is a cudaStream_t (NULL = the legacy default stream).
 * The Python modules call both around every entry point that receives a torch CUDA tensor. */
int nlos_ctx_wait_stream(nlos_ctx* ctx, void* stream);
int nlos_ctx_signal_stream(nlos_ctx* ctx, void* stream);
/* keys: "reuse_visibility" (1), "chunk_forward" (0 = auto), "chunk_gradient" (0 = auto), "timing" (0),
 *       "forward_algo" (0 = auto, 1 = BVH traversal kernel, 2 = per-source perspective-grid kernel, 3 = perspective grid shared by
 *       groups of neighbouring wall points), "grid_res" (0 = auto: cells per axis of the perspective grid), "grid_cap" (0 = none; test
 *       hook: entry budget of the grid per wall point / group), "count_work" (0); shared grid only: "group_side" (0 = 4: a group is a tile
 *       of side x side wall spacings), "grid_slices" (0 = 16 slices of 1/depth), "grid_budget_mb" (0 = 6144: scratch for one batch of groups) */
int nlos_ctx_set_option(nlos_ctx* ctx, const char* key, int64_t value);
/* ms of the last call: {scene build, forward, residual, gradient, total}; needs option "timing" = 1 */
int nlos_ctx_get_timing(nlos_ctx* ctx, float* ms5);
uint64_t nlos_ctx_launch_count(nlos_ctx* ctx);         /* kernels launched through this context so far */
/* MEASUREMENT: with option "count_work" = 1 the perspective-grid forward kernel counts its own work; after the call out8 holds
 * {samples generated (not plane-culled), rays traced, entry words scanned, cell-level check passes (= exact-test candidates incl. self),
 *  visible samples, wall points handled without a grid, grid resolution G of the last forward launch, its kernel (1 = BVH traversal,
 *  2 = perspective grid per wall point, 3 = perspective grid shared by groups of wall points)}; the first six are zero unless the counting instantiation ran (Lambertian, face normals).  bench.py prices
 *  roofline.executed with them. */
int nlos_ctx_get_work_counters(nlos_ctx* ctx, uint64_t* out8);
/* TEST HOOK (no reference counterpart): replace the counter-based generator by an external stream of (S,T) pairs, n floats (host or
 * device pointer, copied); sample k of (global source s, triangle f) reads st[2*((s*numTriangles + f)*spp + k)] — the order in which
 * one worker of the reference consumes its Mersenne-Twister stream (sampler.cpp:20-33, transient_and_gradient.cpp:184-186), so that
 * outputs can be compared with the reference's own sample for sample.  n = 0 restores the generator.  Not honoured by
 * nlos_streamed_render_vertex_gradient (the reference skips non-adjacent triangles there before drawing). */
int nlos_ctx_set_external_samples(nlos_ctx* ctx, const float* st, int64_t n);

/* ---- module `renderer` (smoothed_transient/) ----------------------------------------------------- */

/* smoothed_transient/stratifiedStreamedTransientRenderer.h:5  streamed_render_transient
 * (renderer.pyx:175 renderStreamedTransient, :139 ...Shading, :157 ...wAlbedo) */
int nlos_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                   const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                   const float* vertexAlbedo /*nullable*/, const int* trianglesD, int numTriangles,
                                   int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                   float pathlengthResolution, double* transient, double* pathlengths,
                                   int refine_scale, int sigma_bin, int numBins);

/* smoothed_transient/stratifiedStreamedTransientRenderer.h:3  streamed_render_intensity
 * (renderer.pyx:191 renderStreamedTriangleIntensity); intensity[F] is accumulated into */
int nlos_streamed_render_intensity(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                   const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                   const int* trianglesD, int numTriangles, int numSamples,
                                   float pathlengthLowerBound, float pathlengthUpperBound, double* intensity);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:7  streamed_render_gradient
 * (renderer.pyx:94 renderStreamedGradient, :116 renderStreamedShadingGradient) */
int nlos_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                  int measurement, const float* normalD, const float* verticesD, int numVertices,
                                  const float* vertexNormal /*nullable*/, const int* trianglesD, int numTriangles,
                                  int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                  float pathlengthResolution, double* transient, double* pathlengths, double* gradient,
                                  int refine_scale, int sigma_bin, int testing_flag, int loss_test, int numBins);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:5  streamed_render_gradient_w_albedo
 * (renderer.pyx:55 renderStreamedGradientWithAlbedo) */
int nlos_streamed_render_gradient_w_albedo(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                           int measurement, const float* normalD, const float* verticesD, int numVertices,
                                           const float* albedo, const int* trianglesD, int numTriangles, int numSamples,
                                           float pathlengthLowerBound, float pathlengthUpperBound, float pathlengthResolution,
                                           double* transient, double* pathlengths, double* gradient, int refine_scale,
                                           int sigma_bin, int testing_flag, int loss_test, int numBins);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:3  streamed_render_gradient_albedo
 * (renderer.pyx:35 renderStreamedGradientAlbedo); the reference returns the double */
int nlos_streamed_render_gradient_albedo(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                         int measurement, const float* normalD, const float* verticesD, int numVertices,
                                         const float* albedo, const int* trianglesD, int numTriangles, int numSamples,
                                         float pathlengthLowerBound, float pathlengthUpperBound, float pathlengthResolution,
                                         double* transient, double* pathlengths, int refine_scale, int sigma_bin,
                                         int testing_flag, int loss_test, int numBins, double* result /*host*/);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:9  streamed_render_vertex_gradient
 * (renderer.pyx:78 renderStreamedVertexGradient — the Python side hard-codes measurement = 1, renderer.pyx:88);
 * gradient[B,3] = per-time-bin gradient of vertex `vertex_num`, accumulated into */
int nlos_streamed_render_vertex_gradient(nlos_ctx* ctx, int vertex_num, const float* originD, int measurement, const float* normalD,
                                         const float* verticesD, int numVertices, const int* trianglesD, int numTriangles,
                                         int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                         float pathlengthResolution, double* gradient, int refine_scale, int sigma_bin, int numBins);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:11  streamed_render_normal_smoothing (renderer.pyx:13);
 * the reference returns the regulariser value; curvature_grad[V,3] is overwritten.  Per-vertex writes use '=' in the
 * reference (last adjacent face wins, scheduler dependent); here the adjacent face with the highest index wins. */
int nlos_streamed_render_normal_smoothing(nlos_ctx* ctx, const float* verticesD, int numVertices, const int* trianglesD,
                                          int numTriangles, const int* face_affinity, double* curvature_grad,
                                          double* value_out /*host, or device (then no host synchronisation)*/);

/* smoothed_transient/stratifiedStreamedGradientRenderer.h:13  streamed_render_curvature_grad (renderer.pyx:26) */
int nlos_streamed_render_curvature_grad(nlos_ctx* ctx, const float* verticesD, int numVertices, const int* trianglesD,
                                        int numTriangles, double* curvature_grad);

/* ---- module `ggx` (ggx/) -------------------------------------------------------------------------- */

/* ggx/stratifiedStreamedTransientRenderer.h:4  streamed_render_transient (ggx.pyx:118, :82, :100) */
int nlos_ggx_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                       const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                       const float* vertexAlbedo /*nullable*/, const int* trianglesD, int numTriangles,
                                       float alpha, int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                       float pathlengthResolution, double* transient, double* pathlengths,
                                       int refine_scale, int sigma_bin, int numBins);

/* ggx/stratifiedStreamedTransientRenderer.h:3  streamed_render_intensity (ggx.pyx:134) */
int nlos_ggx_streamed_render_intensity(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                       const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                       const int* trianglesD, int numTriangles, float alpha, int numSamples,
                                       float pathlengthLowerBound, float pathlengthUpperBound, double* intensity);

/* ggx/stratifiedStreamedGradientRenderer.h:6  streamed_render_gradient (ggx.pyx:37, :59); no loss_test */
int nlos_ggx_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                      int measurement, const float* normalD, const float* verticesD, int numVertices,
                                      const float* vertexNormal /*nullable*/, const int* trianglesD, int numTriangles,
                                      float alpha, int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                      float pathlengthResolution, double* transient, double* pathlengths, double* gradient,
                                      int refine_scale, int sigma_bin, int testing_flag, int numBins);

/* ggx/stratifiedStreamedGradientRenderer.h:4  streamed_render_gradient_alpha (ggx.pyx:12) */
int nlos_ggx_streamed_render_gradient_alpha(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                            int measurement, const float* normalD, const float* verticesD, int numVertices,
                                            const float* vertexNormal /*nullable*/, const int* trianglesD, int numTriangles,
                                            float alpha, int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                            float pathlengthResolution, double* transient, double* pathlengths,
                                            int refine_scale, int sigma_bin, int numBins, double* result /*host*/);

/* ---- module `jitter` (jitter/), SURVEY.md 8f row N3: tabulated SPAD-jitter temporal kernel --------------------- */

/* jitter/stratifiedStreamedTransientRenderer.h:3  streamed_render_transient (jitter.pyx:140, :104, :122):
 * T[s,b] = sum_i weight[i] * hist[s, b + weight_offset - i] */
int nlos_jitter_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                          const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                          const float* vertexAlbedo /*nullable*/, const int* trianglesD, int numTriangles,
                                          int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                          float pathlengthResolution, const double* weight, int weight_offset,
                                          int weight_length, double* transient, double* pathlengths, int numBins);
/* jitter/stratifiedStreamedGradientRenderer.h:9  streamed_render_gradient (jitter.pyx:59) */
int nlos_jitter_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD,
                                         int measurement, const float* normalD, const float* verticesD, int numVertices,
                                         const float* vertexNormal /*nullable*/, const int* trianglesD, int numTriangles,
                                         int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                         float pathlengthResolution, const double* jitter_weight, const double* jitter_grad,
                                         int weight_offset, int weight_length, double* transient, double* pathlengths,
                                         double* gradient, int testing_flag, int numBins);

/* ---- first-generation module `renderer` (stratified_transient_raytracer/), SURVEY.md 8f row N4 ------------------ */

/* stratified_transient_raytracer/stratifiedStreamedTransientRenderer.h  streamed_render_transient (renderer.pyx:36-88):
 * raw histogram (no temporal smoothing); the form factor is NOT clamped (stratifiedStreamedTransientRenderer.cpp:130-137),
 * so a visible back-facing sample contributes ff^2 */
int nlos_sr_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD,
                                      const float* verticesD, int numVertices, const float* vertexNormal /*nullable*/,
                                      const float* vertexAlbedo /*nullable*/, const int* trianglesD, int numTriangles,
                                      int numSamples, float pathlengthLowerBound, float pathlengthUpperBound,
                                      float pathlengthResolution, double* transient, double* pathlengths, int numBins);
/* stratified_transient_raytracer/stratifiedTransientRenderer.h  render_transient (renderer.pyx:93-102): one origin,
 * originD[3], normalD[3], transient[numBins] */
int nlos_sr_render_transient(nlos_ctx* ctx, const float* originD, const float* normalD, const float* verticesD,
                             int numVertices, const int* trianglesD, int numTriangles, int numSamples,
                             float pathlengthLowerBound, float pathlengthUpperBound, float pathlengthResolution,
                             double* transient, double* pathlengths, int numBins);
/* stratified_transient_raytracer/stratifiedStreamedGradientRenderer.h  streamed_render_gradient (renderer.pyx:22-34):
 * difference = data - transient, box-filtered twice with half-width w_width (stratifiedStreamedGradientRenderer.cpp:447-458);
 * one tap per sample, normal-variation term always included (:266-271); gradient is OVERWRITTEN (:414), not accumulated.
 * The reference's output-index slips (:278, :290-291) are not reproduced. */
int nlos_sr_streamed_render_gradient(nlos_ctx* ctx, const double* data, const float* originD, int measurement,
                                     const float* normalD, const float* verticesD, int numVertices, const int* trianglesD,
                                     int numTriangles, int numSamples, float pathlengthLowerBound,
                                     float pathlengthUpperBound, float pathlengthResolution, int w_width, double* transient,
                                     double* pathlengths, double* gradient, int numBins);

/* ---- module `embree_intersector` (embree_intersector/), SURVEY.md 8f row N2 ---------------------------- */

/* embree_intersector/c_embree_intersector.h:8  embree3_tbb_line_intersection (embree_intersector.pyx:92):
 * nearest hit per ray, intersect[3i..3i+2] = (primID, u, v) as floats, or intersect[3i] = -1 (u,v untouched) */
int nlos_embree3_tbb_line_intersection(nlos_ctx* ctx, const float* originsD, const float* directionsD, int num_ray,
                                       const float* verticesD, int num_vertices, const int* trianglesD, int num_triangles,
                                       float* intersect);
/* c_embree_intersector.h:9  embree3_tbb_short_line_intersection (embree_intersector.pyx:81): intersect[i] = primID or -1 */
int nlos_embree3_tbb_short_line_intersection(nlos_ctx* ctx, const float* originsD, const float* directionsD, int num_ray,
                                             const float* verticesD, int num_vertices, const int* trianglesD,
                                             int num_triangles, float* intersect);
/* c_embree_intersector.h:4  barycentric_to_world (embree_intersector.pyx:69); the two mesh sizes are additions
 * (needed to stage host arrays); rows with primID < 0 are left untouched */
int nlos_barycentric_to_world(nlos_ctx* ctx, const float* verticesD, int num_vertices, const int* trianglesD, int num_triangles,
                              const float* barycoord, int num_ray, float* intersection_p);

/* ---- test / profiling hooks ------------------------------------------------------------------------ */

/* TEST / DEBUG: the visibility words the forward pass of the LAST gradient call left for its gradient pass (one bit per sample:
 * word [(source*spp + k) * ceil(F/32) + w], bit l = Morton-ordered triangle 32*w + l).  n_available receives the word count;
 * out may be NULL to query it.  Used by the tests to show that the two forward kernels decide every sample identically. */
int nlos_debug_copy_visibility_words(nlos_ctx* ctx, uint32_t* out, int64_t n, int64_t* n_available);

/* Pure-geometry per-sample visibility (nearest hit == sampled triangle, TG.cpp:206) as bytes [L,F,spp];
 * counters (nullable, host) = {rays traced, box tests, triangle tests}. */
int nlos_debug_visibility(nlos_ctx* ctx, const float* originD, int numSources, const float* verticesD, int numVertices,
                          const int* trianglesD, int numTriangles, int numSamples, uint8_t* visibility, uint64_t* counters3);

#ifdef __cplusplus
}
#endif
#endif /* NLOS_B200_H_ */
