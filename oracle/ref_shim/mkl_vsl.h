// Stand-in for Intel MKL's mkl_vsl.h (see oracle/ref_shim/README.md): 1-D double convolution task, direct summation.
// z[i] = sum_j x[j] * y[start + i - j], i in [0, nz)  (the "start" is set by vslConvSetStart, default 0).
#pragma once
#define VSL_CONV_MODE_AUTO 0
#define VSL_CONV_MODE_DIRECT 1
#define VSL_STATUS_OK 0
struct refshim_conv_task { int nx, ny, nz, start; };
typedef refshim_conv_task* VSLConvTaskPtr;
#ifndef MKL_INT
#define MKL_INT int
#endif
inline int vsldConvNewTask1D(VSLConvTaskPtr* t, int, int nx, int ny, int nz) { *t = new refshim_conv_task{nx, ny, nz, 0}; return VSL_STATUS_OK; }
inline int vslConvSetStart(VSLConvTaskPtr t, const int* start) { t->start = start[0]; return VSL_STATUS_OK; }
inline int vsldConvExec1D(VSLConvTaskPtr t, const double* x, int xs, const double* y, int ys, double* z, int zs) {
    for (int i = 0; i < t->nz; ++i) {
        double acc = 0.0; const int k = t->start + i;
        for (int j = 0; j < t->nx; ++j) { const int m = k - j; if (m >= 0 && m < t->ny) acc += x[j * xs] * y[m * ys]; }
        z[i * zs] = acc;
    }
    return VSL_STATUS_OK;
}
inline int vslConvDeleteTask(VSLConvTaskPtr* t) { delete *t; *t = nullptr; return VSL_STATUS_OK; }
