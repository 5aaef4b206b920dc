// Stand-in for Embree's algorithms/parallel_for.h (see oracle/ref_shim/README.md): blocked parallel_for over OpenMP with a static
// schedule (so a run with a fixed thread count is reproducible).
#pragma once
#include "../tasking/taskscheduler.h"
namespace embree {
template <typename Index> struct range {
    Index b, e; range(Index b_, Index e_) : b(b_), e(e_) {}
    Index begin() const { return b; } Index end() const { return e; } Index size() const { return e - b; }
};
template <typename Index, typename Func> inline void parallel_for(Index first, Index last, const Func& func) {
    const long long n = (long long)last - (long long)first; if (n <= 0) return;
    const long long block = 16, blocks = (n + block - 1) / block;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < blocks; ++i) {
        const Index b = (Index)(first + i * block), e = (Index)(first + ((i + 1) * block < n ? (i + 1) * block : n));
        func(range<Index>(b, e));
    }
}
template <typename Index, typename Func> inline void parallel_for(Index first, Index last, Index, const Func& func) { parallel_for(first, last, func); }
}
