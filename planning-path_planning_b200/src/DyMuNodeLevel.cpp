// DyMuNodeLevel.cpp -- the node-level public methods of the reference class (H.hpp:497-575).
//
// In the reference these are the inner steps of its sequential loops and work on the pointer
// graph.  Here the loops run on the device, so a caller that names one of them gets a thin forward
// that acts on a VALUE VIEW of the node (getGlobalNode / getLocalNode) and, where the step writes
// a field, stores the result back into the device plane.  The narrow-band vectors stay empty:
// steps whose only effect is on those vectors (minCost*Node, maxRiskNode, setHorizonCost) have
// nothing to do and say so.  INTEGRATION.md lists which is which.
#include "DyMu.hpp"

#include <math.h>

#include <limits>

#include "dymu_cuda.h"

using namespace PathPlanning_lib;

namespace
{
const double kInf = std::numeric_limits<double>::infinity();

// one axis of gradientNode (G.cpp:722-740 / L.cpp:983-1001): central difference when both
// neighbours carry a finite value, one-sided against the node when one is missing or infinite,
// 0 when there is nothing to difference
double axisDifference(bool has_lo, double lo, double mid, bool has_hi, double hi)
{
    if ((!has_lo && !has_hi) || (has_lo && has_hi && lo == kInf && hi == kInf)) return 0.0;
    if (!has_lo || lo == kInf) return hi - mid;
    if (!has_hi || hi == kInf) return mid - lo;
    return (hi - lo) * 0.5;
}
}  // namespace

// The three passes of computeCostMap (G.cpp:186-308) run over whole planes on the device; per
// node there is nothing left to compute.  The forwards bring the view up to date with the planes,
// which already hold what the pass would have produced.
void DyMuPathPlanner::calculateSlope(globalNode* nodeTarget)
{
    if (!nodeTarget || !dev) return;
    readNode((uint)nodeTarget->pose.position[0], (uint)nodeTarget->pose.position[1], *nodeTarget);
}

void DyMuPathPlanner::calculateNominalCost(globalNode* nodeTarget, int, int)
{
    calculateSlope(nodeTarget);
}

void DyMuPathPlanner::smoothCost(globalNode* nodeTarget) { calculateSlope(nodeTarget); }

// reference: G.cpp:500-546, one application of the upwind update to the node, on the values that
// are resident on the device; an improvement is stored back.
void DyMuPathPlanner::propagateGlobalNode(globalNode* nodeTarget)
{
    if (!nodeTarget || !dev) return;
    uint i = (uint)nodeTarget->pose.position[0], j = (uint)nodeTarget->pose.position[1];
    if (i >= num_nodes_X || j >= num_nodes_Y) return;
    readNode(i, j, *nodeTarget);
    uint32_t idx[4];
    bool has[4] = {j > 0, i > 0, i + 1 < num_nodes_X, j + 1 < num_nodes_Y};  // nb4List order
    const long di[4] = {0, -1, 1, 0}, dj[4] = {-1, 0, 0, 1};
    double t[4] = {kInf, kInf, kInf, kInf};
    int n = 0, where[4];
    for (int k = 0; k < 4; ++k)
        if (has[k])
        {
            idx[n] = (uint32_t)((j + dj[k]) * (long)num_nodes_X + (i + di[k]));
            where[n++] = k;
        }
    double v[4];
    if (n && !deviceOk(dymu_read_cells(dev, DYMU_PLANE_TOTAL_COST, 0, idx, n, v), "propagateGlobalNode")) return;
    for (int q = 0; q < n; ++q) t[where[q]] = v[q];
    double Ty = fmin(t[3], t[0]), Tx = fmin(t[1], t[2]);  // a missing neighbour reads as +inf
    double C = global_res * (nodeTarget->cost) * (2 + nodeTarget->hazard_density - nodeTarget->trafficability);
    double T;
    if ((fabs(Tx - Ty) < C) && (Tx < kInf) && (Ty < kInf))
        T = (Tx + Ty + sqrt(2 * pow(C, 2.0) - pow((Tx - Ty), 2.0))) / 2;
    else
        T = fmin(Tx, Ty) + C;
    if (T < nodeTarget->total_cost)
    {
        nodeTarget->total_cost = T;
        deviceOk(dymu_write_rect(dev, DYMU_PLANE_TOTAL_COST, i, j, 1, 1, &T), "propagateGlobalNode");
    }
}

// The narrow band lives in the device solver's work lists; there is no host vector to pop from.
globalNode* DyMuPathPlanner::minCostGlobalNode()
{
    LOG_WARN_S << "PLANNER (B200): minCostGlobalNode has no narrow band to pop from; the march runs inside "
                  "compute*TotalCostMap";
    return NULL;
}

// reference: G.cpp:718-772
void DyMuPathPlanner::gradientNode(globalNode* nodeTarget, double& dnx, double& dny)
{
    dnx = dny = 0;
    if (!nodeTarget || !dev) return;
    uint i = (uint)nodeTarget->pose.position[0], j = (uint)nodeTarget->pose.position[1];
    if (i >= num_nodes_X || j >= num_nodes_Y) return;
    bool has[4] = {j > 0, i > 0, i + 1 < num_nodes_X, j + 1 < num_nodes_Y};
    const long di[4] = {0, -1, 1, 0}, dj[4] = {-1, 0, 0, 1};
    uint32_t idx[5];
    int n = 0, where[5];
    idx[n] = j * num_nodes_X + i;
    where[n++] = 4;
    for (int k = 0; k < 4; ++k)
        if (has[k])
        {
            idx[n] = (uint32_t)((j + dj[k]) * (long)num_nodes_X + (i + di[k]));
            where[n++] = k;
        }
    double v[5], t[5] = {kInf, kInf, kInf, kInf, kInf};
    if (!deviceOk(dymu_read_cells(dev, DYMU_PLANE_TOTAL_COST, 0, idx, n, v), "gradientNode")) return;
    for (int q = 0; q < n; ++q) t[where[q]] = v[q];
    double dx = axisDifference(has[1], t[1], t[4], has[2], t[2]);
    double dy = axisDifference(has[0], t[0], t[4], has[3], t[3]);
    if ((dx == 0) && (dy == 0)) return;
    dnx = dx / sqrt(pow(dx, 2) + pow(dy, 2));
    dny = dy / sqrt(pow(dx, 2) + pow(dy, 2));
}

// reference: G.cpp:666-714.  One step of the descent from wPos; wPos gets its interpolated
// elevation, the returned waypoint its heading, exactly like the kernel that walks the whole path
// (dymu_extract_global_path asked for two waypoints: the start and the one after it).
base::Waypoint DyMuPathPlanner::computeNextGlobalWaypoint(base::Waypoint& wPos, double tau)
{
    base::Waypoint wNext = wPos;
    if (!dev || !goal_set) return wNext;
    double buf[10];
    uint32_t n = 0;
    int status = 0;
    if (!deviceOk(dymu_extract_global_path(dev, 0, wPos.position[0], wPos.position[1], tau, goal_i, goal_j, buf, 2,
                                           &n, &status),
                  "computeNextGlobalWaypoint"))
        return wNext;
    if (n >= 1) wPos.position[2] = buf[2];
    if (n >= 1)
    {
        // the step the kernel took from the start: x1 = x0 - global_res * tau * dCost
        wNext.position[0] = wPos.position[0] - global_res * tau * buf[3];
        wNext.position[1] = wPos.position[1] - global_res * tau * buf[4];
        wNext.heading = atan2(-buf[4], -buf[3]);
    }
    return wNext;
}

// reference: L.cpp:23-145.  The device window holds every local cell; what remains of
// createLocalMap is the hasLocalMap flag of this one node (no neighbours, unlike
// subdivideGlobalNode) and making sure the window reaches it.
void DyMuPathPlanner::createLocalMap(globalNode* gNode)
{
    if (!gNode || !dev) return;
    long i = (long)gNode->pose.position[0], j = (long)gNode->pose.position[1];
    if (i < 0 || j < 0 || i >= (long)num_nodes_X || j >= (long)num_nodes_Y) return;
    has_local[(size_t)j * num_nodes_X + i] = 1;
    gNode->hasLocalMap = true;
    ensureLocalWindow((double)i * global_res, (double)j * global_res, 0.0, 0.0, true);
}

// A local node view remembers the window cell it was filled from; the window may have grown or
// moved since (dymu_local_reshape), so the cell is looked up again from the node's position.
long DyMuPathPlanner::viewCell(localNode* n)
{
    if (!n || !dev || !local_ready) return -1;
    int64_t cell = -1;
    if (!deviceOk(dymu_local_cell_of(dev, n->global_pose.position[0], n->global_pose.position[1], &cell),
                  "local node view"))
        return -1;
    n->window_cell = (long)cell;
    return (long)cell;
}

// reference: L.cpp:979-1023 (no zero test: a flat neighbourhood yields NaN, L.cpp:1021-1022)
void DyMuPathPlanner::gradientNode(localNode* nodeTarget, double& dnx, double& dny)
{
    dnx = dny = std::numeric_limits<double>::quiet_NaN();
    if (viewCell(nodeTarget) < 0) return;
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    const long w = (long)wg * r, X = nodeTarget->window_cell % w, Y = nodeTarget->window_cell / w;
    // a local neighbour exists where the global map does (L.cpp:58-145); inside the window
    const long mx = (long)num_nodes_X * r, my = (long)num_nodes_Y * r;
    const long ax = X + gx0 * (long)r, ay = Y + gy0 * (long)r;  // absolute local coordinates
    const long x0 = X > 0 ? X - 1 : X, x1 = X + 1 < w ? X + 1 : X, y0 = Y > 0 ? Y - 1 : Y, y1 = Y + 1 < w ? Y + 1 : Y;
    double d[9];
    const uint32_t rw = (uint32_t)(x1 - x0 + 1), rh = (uint32_t)(y1 - y0 + 1);
    if (!deviceOk(dymu_local_read_rect(dev, DYMU_LPLANE_DEVIATION, (uint32_t)x0, (uint32_t)y0, rw, rh, d),
                  "gradientNode"))
        return;
    auto at = [&](long x, long y) { return d[(y - y0) * rw + (x - x0)]; };
    const bool hl = X > 0 && ax - 1 >= 0, hr = X + 1 < w && ax + 1 < mx;
    const bool hu = Y > 0 && ay - 1 >= 0, hd = Y + 1 < w && ay + 1 < my;
    const double mid = at(X, Y);
    double dx = axisDifference(hl, hl ? at(X - 1, Y) : kInf, mid, hr, hr ? at(X + 1, Y) : kInf);
    double dy = axisDifference(hu, hu ? at(X, Y - 1) : kInf, mid, hd, hd ? at(X, Y + 1) : kInf);
    dnx = dx / sqrt(pow(dx, 2) + pow(dy, 2));
    dny = dy / sqrt(pow(dx, 2) + pow(dy, 2));
}

// reference: L.cpp:441-471 for one obstacle node, on the device like the batched test of
// computeLocalPlanning.
bool DyMuPathPlanner::isBlockingObstacle(localNode* obNode, uint& maxIndex, uint& minIndex)
{
    if (viewCell(obNode) < 0 || current_path.empty()) return false;
    std::vector<double> xy(current_path.size() * 2);
    for (size_t k = 0; k < current_path.size(); ++k)
    {
        xy[2 * k] = current_path[k].position[0];
        xy[2 * k + 1] = current_path[k].position[1];
    }
    uint32_t cell = (uint32_t)obNode->window_cell, mn = minIndex, mx = maxIndex;
    int blocked = 0;
    if (!deviceOk(dymu_local_blocking(dev, &cell, 1, xy.data(), (uint32_t)current_path.size(), risk_distance, &mn,
                                      &mx, &blocked),
                  "isBlockingObstacle"))
        return false;
    minIndex = mn;
    maxIndex = mx;
    return blocked != 0;
}

// ---- steps of the sequential loops that only make sense on the reference's narrow-band vectors.
// The loops themselves (expandRisk L.cpp:493-523, computeLocalPropagation L.cpp:578-698,
// getLocalPath L.cpp:807-849) run on the device; these entry points exist so that code naming
// them still compiles and links, and they report that there is nothing for them to do.
namespace
{
void notAStep(const char* name)
{
    LOG_WARN_S << "PLANNER (B200): " << name << " is a step of a loop that runs on the device in this build; "
               << "call expandRisk / computeLocalPropagation / getLocalPath instead";
}
}  // namespace

localNode* DyMuPathPlanner::maxRiskNode()
{
    notAStep("maxRiskNode");
    return NULL;
}
void DyMuPathPlanner::propagateRisk(localNode*) { notAStep("propagateRisk"); }
void DyMuPathPlanner::setHorizonCost(localNode*) { notAStep("setHorizonCost"); }
void DyMuPathPlanner::propagateLocalNode(localNode*) { notAStep("propagateLocalNode"); }
localNode* DyMuPathPlanner::minCostLocalNode(double, double)
{
    notAStep("minCostLocalNode");
    return NULL;
}
localNode* DyMuPathPlanner::minCostLocalNode(localNode*)
{
    notAStep("minCostLocalNode");
    return NULL;
}
bool DyMuPathPlanner::computeLocalWaypointGDM(base::Waypoint&, double)
{
    notAStep("computeLocalWaypointGDM");
    return false;
}
base::Waypoint DyMuPathPlanner::computeLocalWaypointDijkstra(localNode* lNode)
{
    notAStep("computeLocalWaypointDijkstra");
    base::Waypoint w;
    if (lNode)
    {
        w.position[0] = lNode->global_pose.position[0];
        w.position[1] = lNode->global_pose.position[1];
    }
    return w;
}
