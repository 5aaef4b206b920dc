// Stand-in for Embree's tasking/taskscheduler.h (see oracle/ref_shim/README.md): thread count / index from OpenMP.
#pragma once
#include <cstddef>
#include <omp.h>
namespace embree {
struct TaskScheduler {
    static size_t threadCount() { return (size_t)omp_get_max_threads(); }
    static size_t threadIndex() { return (size_t)omp_get_thread_num(); }
};
}
