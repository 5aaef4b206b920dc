// Second build of the render kernels with the external-sample TEST HOOK compiled in (nlos_ctx_set_external_samples, include/nlos_b200.h):
// identical source, namespace nlos::ext, draw_sample() reads the (S,T) stream instead of Philox.  run_job() dispatches here only while
// a stream is installed, so the production kernels carry no trace of the hook.
#define NLOS_EXT_BUILD 1
#include "render_kernels.cu"
