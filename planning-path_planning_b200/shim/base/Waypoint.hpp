// Stand-in for base/Waypoint.hpp: position[0..2], heading, tolerances.
#ifndef DYMU_SHIM_BASE_WAYPOINT_HPP
#define DYMU_SHIM_BASE_WAYPOINT_HPP
#include <base/Eigen.hpp>
#include <base/Pose.hpp>
namespace base
{
struct Waypoint
{
    Position position;
    double heading;
    double tol_position;
    double tol_heading;
    Waypoint() : heading(0.0), tol_position(0.0), tol_heading(0.0) {}
};
}  // namespace base
#endif
