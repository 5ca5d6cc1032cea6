// Stand-in for <boost/container/vector.hpp> (test infrastructure). The
// reference needs the std::vector interface plus resize(n, default_init)
// (reference: src/simulator-mpi/SimulatorMPI.hpp:51, SimulatorMPI.cpp:819,906).
#pragma once
#include <cstddef>
#include <vector>
using std::size_t;
namespace boost { namespace container {
struct default_init_t {};
static const default_init_t default_init{};
template <class T, class A = std::allocator<T>>
struct vector : std::vector<T, A> {
     using std::vector<T, A>::vector;
     using std::vector<T, A>::resize;
     void resize(size_t n, default_init_t) { std::vector<T, A>::resize(n); }
};
}}  // namespace boost::container
