// Stand-in for <boost/random.hpp> (see oracle/ref_shim/README.md): the three names the reference's sampler.cpp uses, on top of <random>.
#pragma once
#include <random>
#include <climits>
#include <iostream>
namespace boost {
typedef std::mt19937 mt19937;     // same algorithm, same default seed (5489) as boost::mt19937
template <typename T = int> struct uniform_int {
    T lo, hi; uniform_int(T a, T b) : lo(a), hi(b) {}
    template <typename E> T operator()(E& e) { return std::uniform_int_distribution<T>(lo, hi)(e); }
};
template <typename EngineRef, typename Distr> struct variate_generator {
    typedef typename std::remove_reference<EngineRef>::type Engine;
    Engine& e; Distr d; variate_generator(Engine& e_, Distr d_) : e(e_), d(d_) {}
    auto operator()() -> decltype(d(e)) { return d(e); }
};
}
