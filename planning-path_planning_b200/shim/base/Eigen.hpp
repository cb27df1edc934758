// Minimal stand-in for the Rock `base-types` vector typedefs (base/Eigen.hpp).
// The DyMu hot path only ever indexes positions with operator[] (see
// /root/reference/src/DyMu.hpp:17-23 for the includes it expects), so a
// zero-initialised fixed array is sufficient.  Written for this repository;
// not derived from Rock sources.
#ifndef DYMU_SHIM_BASE_EIGEN_HPP
#define DYMU_SHIM_BASE_EIGEN_HPP
#include <cstddef>
namespace base
{
template <int N> struct FixedVector
{
    double v[N];
    FixedVector()
    {
        for (int k = 0; k < N; ++k) v[k] = 0.0;
    }
    double& operator[](std::size_t k) { return v[k]; }
    const double& operator[](std::size_t k) const { return v[k]; }
    double& x() { return v[0]; }
    double& y() { return v[1]; }
};
typedef FixedVector<2> Vector2d;
typedef FixedVector<3> Vector3d;
typedef Vector2d Position2D;
typedef Vector3d Position;
}  // namespace base
#endif
