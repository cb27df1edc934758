// dymu_path.cu -- gradient-descent waypoint extraction on the device-resident total-cost
// plane (reference: computeGlobalPath / computeNextGlobalWaypoint / gradientNode /
// interpolate, src/DyMu_GlobalPathPlanning.cpp:615-784).
//
// The descent is a sequential chain (each waypoint depends on the previous one), so it is
// latency-bound, not bandwidth-bound: one warp walks one path.  Lanes 0..3 evaluate the
// four cell-corner gradients in parallel (each gathers its own 5-point stencil of T),
// the results are exchanged with warp shuffles and every lane carries the identical
// waypoint state, so there is no divergence; lane 0 stores the waypoint.  Batches of
// queries run one warp per query.
#include "dymu_ctx.cuh"

namespace
{
struct PathArgs
{
    const double* T;
    const double* elev;
    uint32_t pitch, nx, ny;
    double gres, tau;
    double x0, y0;
    uint32_t goal_i, goal_j;
    double* out;       // 5 doubles per waypoint
    uint32_t cap;
    uint32_t* result;  // [0] n_out, [1] status
};

// per-axis rule of gradientNode (G.cpp:722-761); lo/hi missing = NULL neighbour
__device__ __forceinline__ double grad_axis(double f, bool has_lo, double flo, bool has_hi,
                                            double fhi)
{
    const double inf = DYMU_INF;
    if ((!has_lo && !has_hi) || (has_lo && has_hi && flo == inf && fhi == inf)) return 0;
    if (!has_lo || flo == inf)
    {
        if (!has_hi) return 0;  // reference dereferences NULL here
        return fhi - f;
    }
    if (!has_hi || fhi == inf) return f - flo;
    return (fhi - flo) * 0.5;
}

// gradientNode(globalNode*), G.cpp:718-772
__device__ __forceinline__ void grad_node(const PathArgs& a, uint32_t i, uint32_t j, double& dnx,
                                          double& dny)
{
    const double* Tc = a.T + (size_t)j * a.pitch + i;
    double f = __ldcg(Tc);
    bool hl = i > 0, hr = i + 1 < a.nx, hd = j > 0, hu = j + 1 < a.ny;
    double fl = hl ? __ldcg(Tc - 1) : 0, fr = hr ? __ldcg(Tc + 1) : 0;
    double fd = hd ? __ldcg(Tc - a.pitch) : 0, fu = hu ? __ldcg(Tc + a.pitch) : 0;
    double dx = grad_axis(f, hl, fl, hr, fr);
    double dy = grad_axis(f, hd, fd, hu, fu);
    if ((dx == 0) && (dy == 0))
    {
        dnx = 0;
        dny = 0;
    }
    else
    {
        double n = sqrt(dx * dx + dy * dy);
        dnx = dx / n;
        dny = dy / n;
    }
}

// computeNextGlobalWaypoint, G.cpp:666-714.  Returns false if the cell leaves the grid.
__device__ __forceinline__ bool next_waypoint(const PathArgs& a, int lane, double wx, double wy,
                                              double& z, double& dCx, double& dCy, double& nx_,
                                              double& ny_)
{
    double gx = wx / a.gres, gy = wy / a.gres;
    if (!(gx >= 0.0) || !(gy >= 0.0) || !(gx < (double)(a.nx - 1)) || !(gy < (double)(a.ny - 1)))
        return false;
    uint32_t cx = (uint32_t)gx, cy = (uint32_t)gy;
    double da = gx - (double)cx, db = gy - (double)cy;
    // lane c (0..3) owns corner (cx + (c&1), cy + (c>>1)): 0=n00 1=n10 2=n01 3=n11
    int c = lane & 3;
    uint32_t ci = cx + (uint32_t)(c & 1), cj = cy + (uint32_t)(c >> 1);
    double gxc, gyc;
    grad_node(a, ci, cj, gxc, gyc);
    double ec = a.elev[(size_t)cj * a.pitch + ci];
    const unsigned full = 0xffffffffu;
    double gx00 = __shfl_sync(full, gxc, 0), gx10 = __shfl_sync(full, gxc, 1);
    double gx01 = __shfl_sync(full, gxc, 2), gx11 = __shfl_sync(full, gxc, 3);
    double gy00 = __shfl_sync(full, gyc, 0), gy10 = __shfl_sync(full, gyc, 1);
    double gy01 = __shfl_sync(full, gyc, 2), gy11 = __shfl_sync(full, gyc, 3);
    double e00 = __shfl_sync(full, ec, 0), e10 = __shfl_sync(full, ec, 1);
    double e01 = __shfl_sync(full, ec, 2), e11 = __shfl_sync(full, ec, 3);
    dCx = dymu_interp(da, db, gx00, gx01, gx10, gx11);
    dCy = dymu_interp(da, db, gy00, gy01, gy10, gy11);
    // elevation corners are passed as (e00, e10, e01, e11) into (g00, g01, g10, g11):
    // the reference's swapped order, G.cpp:699-704
    z = dymu_interp(da, db, e00, e10, e01, e11);
    nx_ = wx - a.gres * a.tau * dCx;
    ny_ = wy - a.gres * a.tau * dCy;
    return true;
}

__device__ __forceinline__ double dist2d(double ax, double ay, double bx, double by)
{
    return sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by));
}

// computeGlobalPath, G.cpp:615-662
__global__ void __launch_bounds__(32, 1) k_global_path(const PathArgs* args_arr)
{
    const PathArgs a = args_arr[blockIdx.x];
    const int lane = threadIdx.x;
    const double sx = a.gres * (double)a.goal_i, sy = a.gres * (double)a.goal_j;
    uint32_t n = 0, status = DYMU_PATH_OK;
    double wx = a.x0, wy = a.y0, z, dCx, dCy, nx_, ny_;

    auto push = [&](double x, double y, double zz, double dx, double dy) -> bool {
        if (n >= a.cap) return false;
        if (lane == 0)
        {
            double* o = a.out + (size_t)5 * n;
            o[0] = x; o[1] = y; o[2] = zz; o[3] = dx; o[4] = dy;
        }
        n++;
        return true;
    };

    if (!next_waypoint(a, lane, wx, wy, z, dCx, dCy, nx_, ny_)) status = DYMU_PATH_OUTSIDE;
    else if (isnan(nx_) || isnan(ny_)) status = DYMU_PATH_NAN;
    else
    {
        push(wx, wy, z, dCx, dCy);
        wx = nx_;
        wy = ny_;
        while (dist2d(wx, wy, sx, sy) > 2.0 * a.gres)
        {
            if (!next_waypoint(a, lane, wx, wy, z, dCx, dCy, nx_, ny_))
            {
                status = DYMU_PATH_OUTSIDE;
                break;
            }
            if (!push(wx, wy, z, dCx, dCy))
            {
                status = DYMU_PATH_CAPACITY;
                break;
            }
            if (dist2d(wx, wy, nx_, ny_) < 0.01 * a.tau * a.gres)
            {
                status = DYMU_PATH_STALLED;
                break;
            }
            // a NaN step simply ends the reference's while loop (sqrt(NaN) > x is false)
            // and the goal is appended: same here, the loop condition fails next
            wx = nx_;
            wy = ny_;
        }
    }
    if (lane == 0)
    {
        a.result[0] = n;
        a.result[1] = status;
    }
}
}  // namespace

extern "C" {

int dymu_extract_global_path(dymu_ctx* ctx, uint32_t slot, double x0, double y0, double tau,
                             uint32_t goal_i, uint32_t goal_j, double* out, uint32_t cap,
                             uint32_t* n_out, int* status)
{
    if (!ctx || !out || !n_out || !status || cap == 0 || slot >= ctx->n_slots) return DYMU_ERR_ARG;
    if (goal_i >= ctx->nx || goal_j >= ctx->ny) return DYMU_ERR_ARG;
    size_t out_bytes = (size_t)cap * 5 * sizeof(double);
    size_t hdr = 256;
    DYMU_TRY(dymu_internal_scratch(ctx, hdr + out_bytes, hdr + out_bytes));
    PathArgs* h_args = (PathArgs*)ctx->h_pinned;
    PathArgs a;
    a.T = ctx->T + (size_t)slot * ctx->pitch * ctx->rows;
    a.elev = ctx->elev;
    a.pitch = ctx->pitch; a.nx = ctx->nx; a.ny = ctx->ny;
    a.gres = ctx->gres; a.tau = tau; a.x0 = x0; a.y0 = y0;
    a.goal_i = goal_i; a.goal_j = goal_j;
    a.out = (double*)((char*)ctx->d_scratch + hdr);
    a.cap = cap;
    a.result = (uint32_t*)((char*)ctx->d_scratch + sizeof(PathArgs));
    static_assert(sizeof(PathArgs) + 8 <= 256, "path header overflow");
    *h_args = a;
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scratch, h_args, sizeof(PathArgs),
                                       cudaMemcpyHostToDevice, ctx->stream));
    k_global_path<<<1, 32, 0, ctx->stream>>>((const PathArgs*)ctx->d_scratch);
    ctx->launches++;
    DYMU_CUDA_TRY(ctx, cudaGetLastError());
    uint32_t* h_res = (uint32_t*)((char*)ctx->h_pinned + sizeof(PathArgs));
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(h_res, a.result, 8, cudaMemcpyDeviceToHost, ctx->stream));
    DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = h_res[0];
    *status = (int)h_res[1];
    if (h_res[0])
    {
        DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(out, a.out, (size_t)h_res[0] * 5 * sizeof(double),
                                           cudaMemcpyDeviceToHost, ctx->stream));
        DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return DYMU_OK;
}

}  // extern "C"
