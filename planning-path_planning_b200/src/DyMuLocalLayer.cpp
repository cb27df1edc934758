// DyMuLocalLayer.cpp -- host side of the drop-in DyMuPathPlanner, local layer.
//
// Mirrors the public behaviour of the reference's src/DyMu_LocalPathRepairing.cpp
// ("L.cpp").  The sequential control logic (which waypoint is blocked, where to overtake
// and reconnect, how the repaired piece is spliced into current_path, the trafficability
// feedback) is host code like in the reference; obstacle ingestion, risk dilation, the
// local wave propagation and the local descent run on the device window.
//
// hasLocalMap (DyMu.hpp:77) is bookkeeping about which global nodes the reference would
// have subdivided.  It is emulated exactly on the host as a byte plane: every place where
// the reference calls subdivideGlobalNode (L.cpp:150-156) marks the node and its eight
// neighbours.
#include "DyMu.hpp"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>

#include "dymu_cuda.h"
#include "DyMuTiming.hpp"

using namespace PathPlanning_lib;

namespace
{
const double kInf = std::numeric_limits<double>::infinity();

double dist2(const base::Waypoint& a, const base::Waypoint& b)
{
    return sqrt(pow(a.position[0] - b.position[0], 2) + pow(a.position[1] - b.position[1], 2));
}
}  // namespace

/**********************SUBDIVIDE THE GLOBAL NODE*******************************/
// reference: L.cpp:150-156 (node + 8 neighbours get a local map)
void DyMuPathPlanner::subdivideIndex(long g)
{
    if (g < 0) return;
    long i = g % num_nodes_X, j = g / num_nodes_X;
    for (long dj = -1; dj <= 1; ++dj)
        for (long di = -1; di <= 1; ++di)
        {
            long ii = i + di, jj = j + dj;
            if (ii < 0 || jj < 0 || ii >= (long)num_nodes_X || jj >= (long)num_nodes_Y) continue;
            has_local[(size_t)jj * num_nodes_X + ii] = 1;
        }
}

void DyMuPathPlanner::subdivideGlobalNode(globalNode* gNode)
{
    if (!gNode) return;
    subdivideIndex((long)gNode->pose.position[1] * num_nodes_X + (long)gNode->pose.position[0]);
}

// Makes sure the device window exists and covers the rectangle centred on (x, y) with the
// given half extents (metres) plus a margin for the risk band and the local wave.  The reference's
// local layer is unbounded and never forgets (L.cpp:150-156, G.cpp:36); the device window follows
// it by growing to the union of what it holds and what is asked for (obstacles and risk move
// along, dymu_local_reshape).  Only beyond kMaxLocalWindow nodes it slides instead, keeping the
// overlap and dropping what falls outside, with a warning.
bool DyMuPathPlanner::ensureLocalWindow(double x, double y, double half_x, double half_y,
                                        bool may_grow)
{
    if (!dev) return false;
    double margin = 3.0 * global_res;
    long lo_x = (long)floor((x - half_x - margin) / global_res), hi_x = (long)ceil((x + half_x + margin) / global_res);
    long lo_y = (long)floor((y - half_y - margin) / global_res), hi_y = (long)ceil((y + half_y + margin) / global_res);
    if (!local_ready)
    {
        long need = std::max(hi_x - lo_x, hi_y - lo_y);
        long wg = std::max((long)local_window_nodes, need);
        int64_t gx0 = (lo_x + hi_x) / 2 - wg / 2, gy0 = (lo_y + hi_y) / 2 - wg / 2;
        // the window allocated by initGlobalLayer is empty: placing it is a clear, not a copy
        if (local_created && wg == (long)local_window_nodes)
        {
            if (!deviceOk(dymu_local_anchor(dev, gx0, gy0), "local window")) return false;
        }
        else if (!deviceOk(dymu_local_reshape(dev, (uint32_t)wg, gx0, gy0), "local window"))
            return false;
        local_created = true;
        local_ready = true;
        return true;
    }
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    bool fits = lo_x >= (long)gx0 && hi_x <= (long)(gx0 + wg) && lo_y >= (long)gy0 && hi_y <= (long)(gy0 + wg);
    if (fits) return true;
    if (!may_grow) return false;
    return growLocalWindow(lo_x, hi_x, lo_y, hi_y);
}

// New window = union of the current one and [lo, hi) (global nodes) with a quarter of slack.
bool DyMuPathPlanner::growLocalWindow(long lo_x, long hi_x, long lo_y, long hi_y)
{
    const long kMaxLocalWindow = 512;  // nodes per edge; whole-window kernels stay cheap below this
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    long ux0 = std::min(lo_x, (long)gx0), ux1 = std::max(hi_x, (long)(gx0 + wg));
    long uy0 = std::min(lo_y, (long)gy0), uy1 = std::max(hi_y, (long)(gy0 + wg));
    long need = std::max(ux1 - ux0, uy1 - uy0);
    long nwg = need + need / 4;
    nwg = (nwg + 7) / 8 * 8;
    int64_t nx0, ny0;
    if (nwg <= kMaxLocalWindow && (uint64_t)nwg * r <= 65535u)
    {
        nx0 = (ux0 + ux1) / 2 - nwg / 2;
        ny0 = (uy0 + uy1) / 2 - nwg / 2;
    }
    else
    {
        // slide: same size (or what the request needs), centred on the request
        nwg = std::max((long)wg, std::max(hi_x - lo_x, hi_y - lo_y));
        nx0 = (lo_x + hi_x) / 2 - nwg / 2;
        ny0 = (lo_y + hi_y) / 2 - nwg / 2;
        LOG_WARN_S << "PLANNER (B200): local window limit reached; local obstacles outside the new window are dropped";
    }
    return deviceOk(dymu_local_reshape(dev, (uint32_t)nwg, nx0, ny0), "local window");
}

// localNode::global_pose of a window cell, L.cpp:35-40
void DyMuPathPlanner::localCellPose(long cell, double& gx, double& gy) const
{
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    long w = (long)wg * r;
    long X = cell % w, Y = cell / w;
    double px = (double)(gx0 + X / r), py = (double)(gy0 + Y / r);
    double lx = (double)(X % r), ly = (double)(Y % r), rr = (double)r;
    gx = px - 0.5 + (0.5 / rr) + lx * (1 / rr);
    gy = py - 0.5 + (0.5 / rr) + ly * (1 / rr);
}

/*************************GET CLOSEST LOCAL NODE*******************************/
// reference: L.cpp:160-189.  Returns a value view of the local node under pos.
localNode* DyMuPathPlanner::getLocalNode(base::Pose2D pos)
{
    base::Waypoint w;
    w.position[0] = pos.position[0];
    w.position[1] = pos.position[1];
    return getLocalNode(w);
}

localNode* DyMuPathPlanner::getLocalNode(base::Waypoint wPos)
{
    long g = nearestIndex(wPos.position[0], wPos.position[1]);
    if (g < 0 || !dev) return NULL;
    subdivideIndex(g);
    if (!ensureLocalWindow(wPos.position[0], wPos.position[1], 0.0, 0.0, true)) return NULL;
    int64_t cell = -1;
    if (!deviceOk(dymu_local_cell_of(dev, wPos.position[0], wPos.position[1], &cell), "getLocalNode"))
        return NULL;
    if (cell < 0) return NULL;
    localNode& v = local_views[(long)cell];
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    uint32_t w = wg * r, X = (uint32_t)(cell % w), Y = (uint32_t)(cell / w);
    v.window_cell = (long)cell;
    v.pose.position[0] = (double)(X % r);
    v.pose.position[1] = (double)(Y % r);
    v.parent_pose.position[0] = (double)(gx0 + X / r);
    v.parent_pose.position[1] = (double)(gy0 + Y / r);
    localCellPose((long)cell, v.global_pose.position[0], v.global_pose.position[1]);
    v.world_pose.position[0] = v.global_pose.position[0] / global_res;
    v.world_pose.position[1] = v.global_pose.position[1] / global_res;
    unsigned char b = 0;
    dymu_local_read_rect(dev, DYMU_LPLANE_RISK, X, Y, 1, 1, &v.risk);
    dymu_local_read_rect(dev, DYMU_LPLANE_DEVIATION, X, Y, 1, 1, &v.deviation);
    dymu_local_read_rect(dev, DYMU_LPLANE_TOTAL_COST, X, Y, 1, 1, &v.total_cost);
    dymu_local_read_rect_u8(dev, DYMU_LPLANE_U8_OBSTACLE, X, Y, 1, 1, &b);
    v.isObstacle = b != 0;
    dymu_local_read_rect_u8(dev, DYMU_LPLANE_U8_STATE, X, Y, 1, 1, &b);
    v.state = b ? CLOSED : OPEN;
    return &v;
}

/*************************LOCAL PATH REPAIRING*********************************/
// reference: L.cpp:193-291
bool DyMuPathPlanner::computeLocalPlanning(base::Waypoint wPos,
                                           base::samples::frame::Frame traversabilityMap,
                                           double res,
                                           std::vector<base::Waypoint>& trajectory,
                                           base::Time& localTime)
{
    if (!dev) return false;
    wPos.position[0] -= global_offset[0];
    wPos.position[1] -= global_offset[1];

    uint height = traversabilityMap.getHeight();
    uint width = traversabilityMap.getWidth();
    if (height == 0 || width == 0) return false;

    // Create new Local Nodes if necessary (L.cpp:210-217)
    uint a = (uint)(fmax(0, ((wPos.position[1] - (double)height / 2 * res) / global_res)));
    uint b = (uint)(fmin((double)num_nodes_Y, ((wPos.position[1] + (double)height / 2 * res) / global_res)));
    uint c = (uint)(fmax(0, ((wPos.position[0] - (double)width / 2 * res) / global_res)));
    uint d = (uint)(fmin((double)num_nodes_X, ((wPos.position[0] + (double)width / 2 * res) / global_res)));
    for (uint j = a; j < b; j++)
        for (uint i = c; i < d; i++) subdivideIndex((long)j * num_nodes_X + i);

    {
        Lap lap("ensureLocalWindow");
        if (!ensureLocalWindow(wPos.position[0], wPos.position[1], res * (double)width / 2,
                               res * (double)height / 2, true))
            return false;
    }
    Lap lap_all("computeLocalPlanning (rest)");

    // getLocalNode(pos) subdivides the nearest node of every in-map pixel (L.cpp:246)
    double offsetX = wPos.position[0] - res * (double)width / 2;
    double offsetY = wPos.position[1] + res * (double)height / 2;
    double globalSizeX = global_res * num_nodes_X - 0.5;
    double globalSizeY = global_res * num_nodes_Y - 0.5;
    for (uint j = 0; j < height; j++)
    {
        double py = offsetY - j * res;
        if (!((py > -0.5) && (py < globalSizeY))) continue;
        for (uint i = 0; i < width; i++)
        {
            double px = offsetX + i * res;
            if ((px > -0.5) && (px < globalSizeX)) subdivideIndex(nearestIndex(px, py));
        }
    }

    // obstacle / risk mask on the device; new obstacle cells come back in raster order
    std::vector<uint32_t> new_cells((size_t)width * height);
    uint32_t n_new = 0;
    Lap* lap_ing = new Lap("ingest");
    if (!deviceOk(dymu_local_ingest(dev, traversabilityMap.image.data(), width, height,
                                    traversabilityMap.getRowSize(), traversabilityMap.getPixelSize(),
                                    res, wPos.position[0], wPos.position[1], new_cells.data(),
                                    (uint32_t)new_cells.size(), &n_new),
                  "computeLocalPlanning"))
    {
        delete lap_ing;
        return false;
    }
    delete lap_ing;

    uint minIndex = current_path.size(), maxIndex = 0;
    bool pathBlocked = false;
    if (n_new > 0)
    {
        pending_risk = true;
        // hazard density feedback, L.cpp:264-274, applied in the reference's order on the
        // small rectangle of global nodes that can be touched
        int64_t gx0, gy0;
        uint32_t wg, r;
        dymu_local_info(dev, &gx0, &gy0, &wg, &r);
        uint32_t w = wg * r;
        long lo_i = num_nodes_X, hi_i = -1, lo_j = num_nodes_Y, hi_j = -1;
        std::vector<long> parents(n_new);
        for (uint32_t k = 0; k < n_new; ++k)
        {
            long X = new_cells[k] % w, Y = new_cells[k] / w;
            // getNearestGlobalNode(lNode->parent_pose), L.cpp:247
            long g = nearestIndex((double)(gx0 + X / r), (double)(gy0 + Y / r));
            parents[k] = g;
            if (g < 0) continue;
            long gi = g % num_nodes_X, gj = g / num_nodes_X;
            lo_i = std::min(lo_i, gi - 1); hi_i = std::max(hi_i, gi + 1);
            lo_j = std::min(lo_j, gj - 1); hi_j = std::max(hi_j, gj + 1);
        }
        lo_i = std::max(lo_i, 0L); lo_j = std::max(lo_j, 0L);
        hi_i = std::min(hi_i, (long)num_nodes_X - 1); hi_j = std::min(hi_j, (long)num_nodes_Y - 1);
        if (hi_i >= lo_i && hi_j >= lo_j)
        {
            uint32_t rw = (uint32_t)(hi_i - lo_i + 1), rh = (uint32_t)(hi_j - lo_j + 1);
            std::vector<double> hz((size_t)rw * rh);
            if (!deviceOk(dymu_read_rect(dev, DYMU_PLANE_HAZARD_DENSITY, (uint32_t)lo_i, (uint32_t)lo_j, rw, rh,
                                         hz.data()),
                          "hazard density"))
                return false;
            static const int dx8[8] = {1, 1, 0, -1, -1, -1, 0, 1};
            static const int dy8[8] = {0, 1, 1, 1, 0, -1, -1, -1};
            for (uint32_t k = 0; k < n_new; ++k)
            {
                if (parents[k] < 0) continue;
                long gi = parents[k] % num_nodes_X, gj = parents[k] / num_nodes_X;
                double& h0 = hz[(size_t)(gj - lo_j) * rw + (gi - lo_i)];
                h0 = std::min(1.0, h0 + 1.0 / (res_ratio * res_ratio));
                for (int q = 0; q < 8; ++q)
                {
                    long ii = gi + dx8[q], jj = gj + dy8[q];
                    if (ii < 0 || jj < 0 || ii >= (long)num_nodes_X || jj >= (long)num_nodes_Y) continue;
                    double& h = hz[(size_t)(jj - lo_j) * rw + (ii - lo_i)];
                    h = std::min(1.0, h + 0.1 / (res_ratio * res_ratio));
                }
            }
            if (!deviceOk(dymu_write_rect(dev, DYMU_PLANE_HAZARD_DENSITY, (uint32_t)lo_i, (uint32_t)lo_j, rw, rh,
                                          hz.data()),
                          "hazard density"))
                return false;
        }
        // path-blocking test, L.cpp:261-263 + 441-471
        if (!current_path.empty())
        {
            std::vector<double> xy(current_path.size() * 2);
            for (size_t k = 0; k < current_path.size(); ++k)
            {
                xy[2 * k] = current_path[k].position[0];
                xy[2 * k + 1] = current_path[k].position[1];
            }
            uint32_t mn = minIndex, mx = maxIndex;
            int blk = 0;
            if (!deviceOk(dymu_local_blocking(dev, new_cells.data(), n_new, xy.data(),
                                              (uint32_t)current_path.size(), risk_distance, &mn, &mx, &blk),
                          "isBlockingObstacle"))
                return false;
            minIndex = mn;
            maxIndex = mx;
            pathBlocked = blk != 0;
        }
    }

    if ((pathBlocked) && (maxIndex > minIndex))
    {
        base::Time tInit = base::Time::now();
        {
            Lap lap("expandRisk");
            expandRisk();
        }
        trajectory.clear();
        {
            Lap lap("repairPath");
            reconnecting_index = repairPath(wPos, maxIndex);
        }
        if (repairing_approach == SWEEPING)
        {
            Lap lap("evaluatePath");
            evaluatePath(reconnecting_index);
        }
        trajectory = current_path;
        localTime = base::Time::now() - tInit;
        return true;
    }
    return false;
}

// Applies subdivideGlobalNode to every global node a device wave looked into.
void DyMuPathPlanner::markEntered()
{
    int64_t gx0, gy0;
    uint32_t wg, r;
    dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    std::vector<unsigned char> e((size_t)wg * wg);
    if (!deviceOk(dymu_local_read_entered(dev, e.data(), 1), "entered nodes")) return;
    for (uint32_t y = 0; y < wg; ++y)
        for (uint32_t x = 0; x < wg; ++x)
            if (e[(size_t)y * wg + x])
            {
                long gi = (long)(gx0 + x), gj = (long)(gy0 + y);
                if (gi < 0 || gj < 0 || gi >= (long)num_nodes_X || gj >= (long)num_nodes_Y) continue;
                subdivideIndex(gj * (long)num_nodes_X + gi);
            }
}

/*******************************RISK DILATION**********************************/
// reference: L.cpp:493-576.  The device iterates propagateRisk to its fixed point.
void DyMuPathPlanner::expandRisk()
{
    if (!dev || !local_ready || !pending_risk) return;
    dymu_solve_stats st;
    if (deviceOk(dymu_local_expand_risk(dev, risk_distance, &st), "expandRisk"))
    {
        pending_risk = false;
        markEntered();
    }
}

// reference: L.cpp:473-491
double DyMuPathPlanner::getTotalCost(localNode* lNode)
{
    if (!lNode || !dev) return kInf;
    uint i = (uint)(lNode->global_pose.position[0]);
    uint j = (uint)(lNode->global_pose.position[1]);
    double a = lNode->global_pose.position[0] - (double)(i);
    double b = lNode->global_pose.position[1] - (double)(j);
    long g = nearestIndex(lNode->parent_pose.position[0], lNode->parent_pose.position[1]);
    if (g < 0) return kInf;
    uint gi = (uint)(g % num_nodes_X), gj = (uint)(g / num_nodes_X);
    double w[4] = {kInf, kInf, kInf, kInf};
    uint32_t idx[4];
    int map[4], n = 0;
    for (int k = 0; k < 4; ++k)
    {
        uint ii = gi + (k & 1), jj = gj + (k >> 1);
        if (ii < num_nodes_X && jj < num_nodes_Y)
        {
            idx[n] = jj * num_nodes_X + ii;
            map[n++] = k;
        }
    }
    double v[4];
    if (!deviceOk(dymu_read_cells(dev, DYMU_PLANE_TOTAL_COST, 0, idx, n, v), "getTotalCost")) return kInf;
    for (int q = 0; q < n; ++q) w[map[q]] = v[q];
    double w00 = w[0], w10 = w[1], w01 = w[2], w11 = w[3];
    return w00 + (w10 - w00) * a + (w01 - w00) * b + (w11 + w00 - w10 - w01) * a * b;
}

/****************************LOCAL PROPAGATION*********************************/
// reference: L.cpp:578-698.  Returns the window cell of the end node or -1 (NULL).
long DyMuPathPlanner::localPropagationCell(base::Waypoint wayp_start, base::Waypoint wOvertake)
{
    if (!dev) return -1;
    // getTotalCost(Waypoint) subtracts the offset from an already offset-free waypoint
    // (L.cpp:583 -> G.cpp:862-863); reproduced as is
    double Tovertake = getTotalCost(wOvertake);
    if (!ensureLocalWindow(wayp_start.position[0], wayp_start.position[1], 0.0, 0.0, true)) return -1;
    subdivideIndex(nearestIndex(wayp_start.position[0], wayp_start.position[1]));  // getLocalNode(wayp_start)
    if (repairing_approach == CONSERVATIVE)
        subdivideIndex(nearestIndex(wOvertake.position[0], wOvertake.position[1]));  // getLocalNode(wOvertake)
    int64_t end_cell = -1;
    uint32_t status = 0;
    uint64_t closed = 0;
    for (int attempt = 0;; ++attempt)
    {
        if (!deviceOk(dymu_local_propagate(dev, repairing_approach == SWEEPING ? 1 : 0, wayp_start.position[0],
                                           wayp_start.position[1], wOvertake.position[0], wOvertake.position[1],
                                           Tovertake, risk_ratio, &end_cell, &status, &closed),
                      "computeLocalPropagation"))
            return -1;
        if (status != DYMU_LOCAL_WINDOW_EXCEEDED || attempt >= 4) break;
        // the wave reached the window border (the reference's layer has none): double the window
        // around its centre -- obstacles and risk move along -- and march again
        int64_t gx0, gy0;
        uint32_t wg, r;
        dymu_local_info(dev, &gx0, &gy0, &wg, &r);
        long half = (long)wg / 2;
        if (!growLocalWindow((long)gx0 - half, (long)(gx0 + wg) + half, (long)gy0 - half, (long)(gy0 + wg) + half))
            return -1;
    }
    markEntered();
    int64_t c = -1;
    dymu_local_cell_of(dev, wayp_start.position[0], wayp_start.position[1], &c);
    local_agent_cell = (long)c;
    if (status != DYMU_LOCAL_OK)
    {
        if (status == DYMU_LOCAL_WINDOW_EXCEEDED)
            LOG_WARN_S << "PLANNER (B200): local wave reached the window border; enlarge the local window";
        return -1;
    }
    return (long)end_cell;
}

localNode* DyMuPathPlanner::computeLocalPropagation(base::Waypoint wInit, base::Waypoint wOvertake)
{
    long cell = localPropagationCell(wInit, wOvertake);
    if (cell < 0) return NULL;
    base::Waypoint w;
    localCellPose(cell, w.position[0], w.position[1]);
    localNode* n = getLocalNode(w);
    if (local_agent_cell >= 0)
    {
        base::Waypoint a;
        localCellPose(local_agent_cell, a.position[0], a.position[1]);
        localNode* ag = getLocalNode(a);
        if (ag)
        {
            agent_view = *ag;
            local_agent = &agent_view;
        }
    }
    return n;
}

/****************************LOCAL PATH****************************************/
// reference: L.cpp:807-849.  Waypoints come back in generation order with the gradient
// (or Dijkstra step) that produced them; the reference inserts each at the FRONT.
std::vector<base::Waypoint> DyMuPathPlanner::localPathFromCell(long cell, base::Waypoint wayp_start)
{
    std::vector<base::Waypoint> trajectory;
    if (!dev || cell < 0) return trajectory;
    uint32_t cap = 1u << 14, n = 0;
    int status = 0;
    std::vector<double> buf;
    for (;;)
    {
        buf.resize((size_t)cap * 6);
        if (!deviceOk(dymu_local_extract_path(dev, cell, wayp_start.position[0], wayp_start.position[1],
                                              global_offset[0], global_offset[1], buf.data(), cap, &n, &status),
                      "getLocalPath"))
            return trajectory;
        if (status != DYMU_PATH_CAPACITY || cap >= (1u << 22)) break;
        cap *= 4;
    }
    trajectory.resize(n);
    for (uint32_t k = 0; k < n; ++k)
    {
        base::Waypoint w;
        w.position[0] = buf[6 * k + 0];
        w.position[1] = buf[6 * k + 1];
        w.position[2] = buf[6 * k + 2];
        int kind = (int)buf[6 * k + 5];
        if (kind == 2) w.heading = 0.0;  // lSetNode->global_pose.orientation, never set (L.cpp:815)
        else w.heading = atan2(buf[6 * k + 4], buf[6 * k + 3]);  // L.cpp:974 / L.cpp:866-867
        trajectory[n - 1 - k] = w;
        subdivideIndex(nearestIndex(w.position[0], w.position[1]));  // getLocalNode side effect
    }
    return trajectory;
}

std::vector<base::Waypoint> DyMuPathPlanner::getLocalPath(localNode* lSetNode, base::Waypoint wInit, double)
{
    if (!lSetNode) return std::vector<base::Waypoint>();
    return localPathFromCell(viewCell(lSetNode), wInit);
}

/*****************************PATH REPAIRING***********************************/
// reference: L.cpp:298-435
int DyMuPathPlanner::repairPath(base::Waypoint wayp_start, uint index)
{
    if (current_path.empty()) return -1;

    uint overtake_index;
    if (repairing_approach == CONSERVATIVE)
    {
        // `reconnecting_index > index` compares int with uint in the reference (L.cpp:317)
        uint ri = (uint)reconnecting_index;
        overtake_index = (ri > index) ? ri : index;
        index = overtake_index;
    }
    else
        overtake_index = index;

    while ((index < current_path.size())
           && (dist2(current_path[index], current_path[overtake_index]) < reconnect_distance))
        index++;

    if (index >= current_path.size() || index == current_path.size() - 1)
    {
        if (index == current_path.size() - 1)
            LOG_WARN_S << "Repairing is not possible, goal is too close to obstacles";
        current_path.clear();
        current_path.push_back(wayp_start);
        return -1;
    }

    long lSet;
    {
        Lap lap("  localPropagationCell");
        lSet = localPropagationCell(wayp_start, current_path[index]);
    }
    if (lSet < 0)
    {
        LOG_WARN_S << "Repairing aborted, the robot is in obstacle area";
        current_path.clear();
        current_path.push_back(wayp_start);
        return -1;
    }

    // Look for the waypoint closest to the rover position (L.cpp:363-378; `proximity`
    // is never updated inside the loop, as in the reference)
    double proximity, candidate_proximity, original_distance = 0, new_distance = 0;
    uint closest_index = 0;
    proximity = dist2(current_path[0], wayp_start);
    for (uint k = 1; k < index; k++)
    {
        candidate_proximity = dist2(current_path[k], wayp_start);
        if (candidate_proximity < proximity) closest_index = k;
    }
    for (uint k = closest_index; k < index; k++)
        original_distance += dist2(current_path[k + 1], current_path[k]);

    std::vector<base::Waypoint> localPath;
    {
        Lap lap("  localPathFromCell");
        localPath = localPathFromCell(lSet, wayp_start);
    }
    double lgx, lgy;
    localCellPose(lSet, lgx, lgy);
    if (localPath.size() > 1)
    {
        for (uint k = 0; k < localPath.size() - 1; k++) new_distance += dist2(localPath[k + 1], localPath[k]);
        // trafficability feedback on the traversed global nodes, L.cpp:388-394
        {
            Lap lap_tr("  trafficability feedback");
            // one read and one write of the bounding rectangle instead of a round trip per waypoint
            std::vector<long> nodes;
            long lo_i = num_nodes_X, hi_i = -1, lo_j = num_nodes_Y, hi_j = -1;
            for (uint k = closest_index; k < index; k++)
            {
                long g = nearestIndex(current_path[k].position[0], current_path[k].position[1]);
                if (g < 0) continue;
                nodes.push_back(g);
                long gi = g % num_nodes_X, gj = g / num_nodes_X;
                lo_i = std::min(lo_i, gi); hi_i = std::max(hi_i, gi);
                lo_j = std::min(lo_j, gj); hi_j = std::max(hi_j, gj);
            }
            if (!nodes.empty())
            {
                uint32_t rw = (uint32_t)(hi_i - lo_i + 1), rh = (uint32_t)(hi_j - lo_j + 1);
                std::vector<double> tr((size_t)rw * rh);
                if (deviceOk(dymu_read_rect(dev, DYMU_PLANE_TRAFFICABILITY, (uint32_t)lo_i, (uint32_t)lo_j, rw, rh,
                                            tr.data()),
                             "trafficability"))
                {
                    for (size_t q = 0; q < nodes.size(); ++q)
                    {
                        double& t = tr[(size_t)(nodes[q] / num_nodes_X - lo_j) * rw + (nodes[q] % num_nodes_X - lo_i)];
                        t = std::min(original_distance / new_distance, t);
                    }
                    deviceOk(dymu_write_rect(dev, DYMU_PLANE_TRAFFICABILITY, (uint32_t)lo_i, (uint32_t)lo_j, rw, rh,
                                             tr.data()),
                             "trafficability");
                }
            }
        }
        if (repairing_approach == CONSERVATIVE)
        {
            current_path.erase(current_path.begin(), current_path.begin() + index);
            localPath.pop_back();
            current_path.insert(current_path.begin(), localPath.begin(), localPath.end());
            return localPath.size();
        }
        else
        {
            base::Waypoint newWaypoint;
            newWaypoint.position[0] = lgx;
            newWaypoint.position[1] = lgy;
            {
                Lap lap("  computeGlobalPath");
                computeGlobalPath(newWaypoint);
            }
            localPath.pop_back();
            current_path.insert(current_path.begin(), localPath.begin(), localPath.end());
            return localPath.size();
        }
    }
    else
    {
        if (repairing_approach == CONSERVATIVE)
        {
            current_path.erase(current_path.begin(), current_path.begin() + index);
            return 0;
        }
        else
        {
            base::Waypoint newWaypoint;
            newWaypoint.position[0] = lgx;
            newWaypoint.position[1] = lgy;
            computeGlobalPath(newWaypoint);
            return 0;
        }
    }
}

/*************EVALUATE IF THE PATH PASSES THROUGH UNDESIRED AREAS**************/
// reference: L.cpp:1027-1109
bool DyMuPathPlanner::evaluatePath(uint starting_index)
{
    uint minIndex = 0, rectifiedIndex = 0;
    bool isBlocked = false;
    std::vector<base::Waypoint> final_path;
    uint index_waypoint = starting_index;
    reconnecting_index = 0;

    // risk under every waypoint, sampled on the device in one call and refreshed whenever a
    // repair rewrites current_path (risk itself is not modified by repairs)
    std::vector<double> risk;
    bool risk_valid = false;
    auto riskAt = [&](uint k) -> double {
        if (!local_ready) return 0.0;
        if (!risk_valid)
        {
            std::vector<double> xy(current_path.size() * 2);
            for (size_t q = 0; q < current_path.size(); ++q)
            {
                xy[2 * q] = current_path[q].position[0];
                xy[2 * q + 1] = current_path[q].position[1];
            }
            risk.assign(current_path.size(), 0.0);
            deviceOk(dymu_local_sample_risk(dev, xy.data(), (uint32_t)current_path.size(), risk.data()),
                     "evaluatePath");
            risk_valid = true;
        }
        return k < risk.size() ? risk[k] : 0.0;
    };

    while (index_waypoint < current_path.size())
    {
        long g = nearestIndex(current_path[index_waypoint].position[0],
                              current_path[index_waypoint].position[1]);
        bool do_repair = false;
        if (g >= 0 && has_local[(size_t)g])
        {
            subdivideIndex(g);  // getLocalNode, L.cpp:1046
            if (riskAt(index_waypoint) > 0.0)
            {
                if (!isBlocked)
                {
                    isBlocked = true;
                    minIndex = index_waypoint;
                }
            }
            else if (isBlocked)
                do_repair = true;
        }
        else if (isBlocked)
            do_repair = true;

        if (do_repair)
        {
            rectifiedIndex = minIndex;
            while (rectifiedIndex > 0)
            {
                if (dist2(current_path[minIndex], current_path[rectifiedIndex]) > 2.0) break;
                rectifiedIndex--;
            }
            final_path.insert(final_path.end(), current_path.begin(), current_path.begin() + rectifiedIndex);
            base::Waypoint ws = current_path[rectifiedIndex];
            index_waypoint = repairPath(ws, index_waypoint);
            risk_valid = false;
            isBlocked = false;
            minIndex = 0;
        }
        if (index_waypoint == (uint)-1)
            return false;
        else
            index_waypoint++;
    }
    if (isBlocked)
        final_path.insert(final_path.end(), current_path.begin(),
                          current_path.begin() + std::min((size_t)minIndex, current_path.size()));
    else
        final_path.insert(final_path.end(), current_path.begin() + std::min((size_t)minIndex, current_path.size()),
                          current_path.end());
    current_path = final_path;
    return true;
}

/**************************LOCAL DEBUG MATRICES********************************/
// reference: L.cpp:1111-1211.  21 x 21 global nodes around the nearest node of the RAW
// rover position (no offset subtraction, L.cpp:1121); nodes without a local map read 0.
namespace
{
std::vector<std::vector<double>> local_matrix(dymu_ctx* dev, const std::vector<unsigned char>& has_local,
                                              uint nx, uint ny, uint res_ratio, long g, int lplane,
                                              bool local_ready)
{
    uint half_num = 10, global_side_num = 2 * half_num + 1, side = global_side_num * res_ratio;
    std::vector<std::vector<double>> m(side, std::vector<double>(side, 0.0));
    if (g < 0) return m;
    long gi = g % nx, gj = g / nx;
    int64_t gx0 = 0, gy0 = 0;
    uint32_t wg = 0, r = res_ratio;
    if (local_ready) dymu_local_info(dev, &gx0, &gy0, &wg, &r);
    std::vector<double> block((size_t)res_ratio * res_ratio);
    for (uint j = 0; j < global_side_num; j++)
        for (uint i = 0; i < global_side_num; i++)
        {
            long cx = gi - half_num + i, cy = gj - half_num + j;
            if (cx < 0 || cy < 0 || cx >= (long)nx || cy >= (long)ny) continue;
            if (!has_local[(size_t)cy * nx + cx]) continue;
            bool in_window = local_ready && cx >= gx0 && cy >= gy0 && cx < gx0 + (long)wg && cy < gy0 + (long)wg;
            if (in_window
                && dymu_local_read_rect(dev, lplane, (uint32_t)(cx - gx0) * r, (uint32_t)(cy - gy0) * r, r, r,
                                        block.data()) == DYMU_OK)
            {
                for (uint l = 0; l < res_ratio; l++)
                    for (uint k = 0; k < res_ratio; k++)
                    {
                        double v = block[(size_t)l * res_ratio + k];
                        if (lplane == DYMU_LPLANE_DEVIATION && v == kInf) v = -1;
                        m[l + j * res_ratio][k + i * res_ratio] = v;
                    }
            }
            else if (lplane == DYMU_LPLANE_DEVIATION)
            {
                // subdivided in the reference but outside the device window: untouched nodes
                for (uint l = 0; l < res_ratio; l++)
                    for (uint k = 0; k < res_ratio; k++) m[l + j * res_ratio][k + i * res_ratio] = -1;
            }
        }
    return m;
}
}  // namespace

std::vector<std::vector<double>> DyMuPathPlanner::getRiskMatrix(base::Waypoint rover_pos)
{
    return local_matrix(dev, has_local, num_nodes_X, num_nodes_Y, res_ratio,
                        nearestIndex(rover_pos.position[0], rover_pos.position[1]), DYMU_LPLANE_RISK,
                        local_ready);
}

std::vector<std::vector<double>> DyMuPathPlanner::getDeviationMatrix(base::Waypoint rover_pos)
{
    return local_matrix(dev, has_local, num_nodes_X, num_nodes_Y, res_ratio,
                        nearestIndex(rover_pos.position[0], rover_pos.position[1]), DYMU_LPLANE_DEVIATION,
                        local_ready);
}

int DyMuPathPlanner::getReconnectingIndex() { return reconnecting_index; }
