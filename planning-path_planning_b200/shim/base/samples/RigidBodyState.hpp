// Stand-in for base/samples/RigidBodyState.hpp: only `.position` is read
// (/root/reference/src/DyMu_GlobalPathPlanning.cpp:945-946).
#ifndef DYMU_SHIM_BASE_SAMPLES_RBS_HPP
#define DYMU_SHIM_BASE_SAMPLES_RBS_HPP
#include <base/Eigen.hpp>
#include <base/Time.hpp>
namespace base
{
namespace samples
{
struct RigidBodyState
{
    Position position;
};
}  // namespace samples
}  // namespace base
#endif
