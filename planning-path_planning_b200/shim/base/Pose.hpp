// Stand-in for base/Pose.hpp: only base::Pose2D {position[0..1], orientation}.
#ifndef DYMU_SHIM_BASE_POSE_HPP
#define DYMU_SHIM_BASE_POSE_HPP
#include <base/Eigen.hpp>
namespace base
{
struct Pose2D
{
    Position2D position;
    double orientation;
    Pose2D() : orientation(0.0) {}
};
}  // namespace base
#endif
