// dymu_path.cu -- gradient-descent waypoint extraction on the device-resident total-cost
// plane (reference: computeGlobalPath / computeNextGlobalWaypoint / gradientNode /
// interpolate, src/DyMu_GlobalPathPlanning.cpp:615-784).
//
// The descent is a sequential chain (each waypoint depends on the previous one), so it is
// latency-bound, not bandwidth-bound: one warp walks one path.  The CTA keeps a
// PATCH x PATCH window in shared memory: total cost, elevation and -- computed by all eight
// warps right after a (re)load -- the normalised node gradients of gradientNode.  A step of
// the walker is then twelve shared-memory reads and three bilinear interpolations; the
// square root and the two divisions per corner are off the per-step chain.  The window is
// re-placed only when the 4x4 stencil of the current cell leaves it, shifted ahead in the
// direction of travel.  Every lane of the walker carries the identical waypoint state (no
// divergence); lane 0 stores the waypoint.  Batches of queries run one CTA per query.
#include <limits.h>

#include "dymu_ctx.cuh"

namespace
{
constexpr int PATCH = 64;  // cells per patch edge: 4 x 32 KB of shared memory
constexpr int PATCH_LEAD = 20;  // how far the window is shifted ahead of the walker

struct Patch
{
    double* t;   // PATCH x PATCH total cost
    double* e;   // PATCH x PATCH elevation
    double* gx;  // PATCH x PATCH normalised gradient (valid one cell inside the border)
    double* gy;
    int x0, y0;  // grid coordinates of patch cell (0,0); valid region clipped to the grid
    bool valid;
};

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
__device__ __forceinline__ void grad_node(const PathArgs& a, const Patch& pt, uint32_t i, uint32_t j,
                                          double& dnx, double& dny)
{
    const double* Tc = pt.t + ((int)j - pt.y0) * PATCH + ((int)i - pt.x0);
    double f = *Tc;
    bool hl = i > 0, hr = i + 1 < a.nx, hd = j > 0, hu = j + 1 < a.ny;
    double fl = hl ? Tc[-1] : 0, fr = hr ? Tc[1] : 0;
    double fd = hd ? Tc[-PATCH] : 0, fu = hu ? Tc[PATCH] : 0;
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
constexpr int PATH_THREADS = 256;  // warp 0 walks the path; all 8 warps reload the patch

// Patch reload protocol: the walker (warp 0) publishes the new origin in shared memory and
// meets the helper warps at a block barrier; everybody loads; a second barrier releases
// the walker.  x0 = INT_MIN tells the helpers to exit.
struct PatchCmd
{
    int x0, y0;
};

__device__ __forceinline__ void load_patch(const PathArgs& a, double* pt_t, double* pt_e, int x0, int y0)
{
    for (int k = threadIdx.x; k < PATCH * PATCH; k += PATH_THREADS)
    {
        int r = k / PATCH, c = k % PATCH;
        int gx = x0 + c, gy = y0 + r;
        double tv = DYMU_INF, ev = 0.0;
        if (gx >= 0 && gy >= 0 && gx < (int)a.nx && gy < (int)a.ny)
        {
            tv = __ldcg(a.T + (size_t)gy * a.pitch + gx);
            ev = a.elev[(size_t)gy * a.pitch + gx];
        }
        pt_t[k] = tv;
        pt_e[k] = ev;
    }
}

// second half of a reload (after a barrier): gradientNode for every node whose stencil is
// inside the window
__device__ __forceinline__ void patch_gradients(const PathArgs& a, const Patch& pt)
{
    for (int k = threadIdx.x; k < PATCH * PATCH; k += PATH_THREADS)
    {
        int r = k / PATCH, c = k % PATCH;
        int gx = pt.x0 + c, gy = pt.y0 + r;
        double dnx = 0, dny = 0;
        if (r >= 1 && c >= 1 && r < PATCH - 1 && c < PATCH - 1 && gx >= 0 && gy >= 0 && gx < (int)a.nx
            && gy < (int)a.ny)
            grad_node(a, pt, (uint32_t)gx, (uint32_t)gy, dnx, dny);
        pt.gx[k] = dnx;
        pt.gy[k] = dny;
    }
}

// makes sure cells [cx-1, cx+2] x [cy-1, cy+2] are inside the shared-memory window; a new
// window is placed PATCH_LEAD cells ahead along the last step (dirx, diry in [-1, 1])
__device__ __forceinline__ void ensure_patch(const PathArgs& a, Patch& pt, PatchCmd* cmd, int lane, int cx,
                                             int cy, double dirx, double diry)
{
    if (pt.valid && cx - 1 >= pt.x0 && cy - 1 >= pt.y0 && cx + 2 < pt.x0 + PATCH && cy + 2 < pt.y0 + PATCH)
        return;
    int lx = (dirx == dirx) ? (int)(dirx * PATCH_LEAD) : 0, ly = (diry == diry) ? (int)(diry * PATCH_LEAD) : 0;
    lx = max(-PATCH_LEAD, min(PATCH_LEAD, lx));
    ly = max(-PATCH_LEAD, min(PATCH_LEAD, ly));
    pt.x0 = cx - PATCH / 2 + lx;
    pt.y0 = cy - PATCH / 2 + ly;
    pt.valid = true;
    if (lane == 0)
    {
        cmd->x0 = pt.x0;
        cmd->y0 = pt.y0;
    }
    __syncthreads();  // helpers are parked at the matching barrier
    load_patch(a, pt.t, pt.e, pt.x0, pt.y0);
    __syncthreads();
    patch_gradients(a, pt);
    __syncthreads();
}

template <bool UNIT_RES>
__device__ __forceinline__ bool next_waypoint(const PathArgs& a, Patch& pt, PatchCmd* cmd, int lane,
                                              double wx, double wy, double& z, double& dCx,
                                              double& dCy, double& nx_, double& ny_)
{
    // x / 1.0 == x exactly: the usual global_res = 1 gets its own instantiation without the
    // two fp64 divisions (as a run-time select the compiler still evaluated them every step)
    double gx = UNIT_RES ? wx : wx / a.gres, gy = UNIT_RES ? wy : wy / a.gres;
    if (!(gx >= 0.0) || !(gy >= 0.0) || !(gx < (double)(a.nx - 1)) || !(gy < (double)(a.ny - 1)))
        return false;
    uint32_t cx = (uint32_t)gx, cy = (uint32_t)gy;
    double da = gx - (double)cx, db = gy - (double)cy;
    ensure_patch(a, pt, cmd, lane, (int)cx, (int)cy, -dCx, -dCy);  // dCx/dCy: previous step
    // corners n00 = (cx, cy), n10 = (cx+1, cy), n01 = (cx, cy+1), n11 = (cx+1, cy+1)
    const int k00 = ((int)cy - pt.y0) * PATCH + ((int)cx - pt.x0);
    const int k10 = k00 + 1, k01 = k00 + PATCH, k11 = k00 + PATCH + 1;
    const double gx00 = pt.gx[k00], gx10 = pt.gx[k10], gx01 = pt.gx[k01], gx11 = pt.gx[k11];
    const double gy00 = pt.gy[k00], gy10 = pt.gy[k10], gy01 = pt.gy[k01], gy11 = pt.gy[k11];
    const double e00 = pt.e[k00], e10 = pt.e[k10], e01 = pt.e[k01], e11 = pt.e[k11];
    dCx = dymu_interp(da, db, gx00, gx01, gx10, gx11);
    dCy = dymu_interp(da, db, gy00, gy01, gy10, gy11);
    // elevation corners are passed as (e00, e10, e01, e11) into (g00, g01, g10, g11):
    // the reference's swapped order, G.cpp:699-704
    z = dymu_interp(da, db, e00, e10, e01, e11);
    nx_ = wx - a.gres * a.tau * dCx;
    ny_ = wy - a.gres * a.tau * dCy;
    return true;
}

// The reference compares sqrt(d2) with a constant twice per step (G.cpp:633, 649).  sqrt is
// correctly rounded and monotone, so "sqrt(d2) > c" is "d2 > hi(c)" with hi(c) the largest
// double whose root does not exceed c, and "sqrt(d2) < c" is "d2 < lo(c)" with lo(c) the
// smallest double whose root reaches c.  Both thresholds are found once per path; the
// per-step tests are then exact without a square root on the dependency chain.
__device__ __forceinline__ double dist2(double ax, double ay, double bx, double by)
{
    return (ax - bx) * (ax - bx) + (ay - by) * (ay - by);
}
__device__ __forceinline__ double next_up(double x) { return __longlong_as_double(__double_as_longlong(x) + 1); }
__device__ __forceinline__ double next_down(double x) { return __longlong_as_double(__double_as_longlong(x) - 1); }
__device__ double sqrt_gt_threshold(double c)
{
    if (!(c > 0)) return (c == 0) ? 0.0 : -1.0;  // sqrt(x) > 0 <=> x > 0; any x >= 0 beats a negative c
    double t = c * c;
    while (sqrt(t) > c) t = next_down(t);
    while (sqrt(next_up(t)) <= c) t = next_up(t);
    return t;  // sqrt(x) > c  <=>  x > t
}
__device__ double sqrt_lt_threshold(double c)
{
    if (!(c > 0)) return 0.0;  // sqrt(x) < c never holds
    double t = c * c;
    while (sqrt(t) < c) t = next_up(t);
    while (t > 0 && sqrt(next_down(t)) >= c) t = next_down(t);
    return t;  // sqrt(x) < c  <=>  x < t
}

// computeGlobalPath, G.cpp:615-662
template <bool UNIT_RES>
__global__ void __launch_bounds__(PATH_THREADS, 1) k_global_path(const PathArgs* args_arr)
{
    const PathArgs a = args_arr[blockIdx.x];
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(16) double path_smem[];
    __shared__ PatchCmd cmd;
    Patch pt;
    pt.t = path_smem;
    pt.e = path_smem + PATCH * PATCH;
    pt.gx = path_smem + 2 * PATCH * PATCH;
    pt.gy = path_smem + 3 * PATCH * PATCH;
    pt.x0 = pt.y0 = 0;
    pt.valid = false;
    if (threadIdx.x >= 32)
    {
        // helper warps: serve patch reloads until the walker signals the end
        for (;;)
        {
            __syncthreads();
            const int x0 = cmd.x0, y0 = cmd.y0;
            if (x0 == INT_MIN) return;
            pt.x0 = x0;
            pt.y0 = y0;
            load_patch(a, pt.t, pt.e, x0, y0);
            __syncthreads();
            patch_gradients(a, pt);
            __syncthreads();
        }
    }
    const double sx = a.gres * (double)a.goal_i, sy = a.gres * (double)a.goal_j;
    uint32_t n = 0, status = DYMU_PATH_OK;
    double wx = a.x0, wy = a.y0, z, dCx = 0, dCy = 0, nx_, ny_;

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

    if (!next_waypoint<UNIT_RES>(a, pt, &cmd, lane, wx, wy, z, dCx, dCy, nx_, ny_)) status = DYMU_PATH_OUTSIDE;
    else if (isnan(nx_) || isnan(ny_)) status = DYMU_PATH_NAN;
    else
    {
        push(wx, wy, z, dCx, dCy);
        wx = nx_;
        wy = ny_;
        const double far2 = sqrt_gt_threshold(2.0 * a.gres);            // dist > 2 * global_res
        const double stall2 = sqrt_lt_threshold(0.01 * a.tau * a.gres);  // dist < 0.01 * tau * global_res
        while (dist2(wx, wy, sx, sy) > far2)
        {
            if (!next_waypoint<UNIT_RES>(a, pt, &cmd, lane, wx, wy, z, dCx, dCy, nx_, ny_))
            {
                status = DYMU_PATH_OUTSIDE;
                break;
            }
            if (!push(wx, wy, z, dCx, dCy))
            {
                status = DYMU_PATH_CAPACITY;
                break;
            }
            if (dist2(wx, wy, nx_, ny_) < stall2)
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
        cmd.x0 = INT_MIN;  // release the helper warps
    }
    __syncthreads();
}

int launch_paths(dymu_ctx* ctx, uint32_t n, size_t smem, const PathArgs* d_args)
{
    auto kernel = (ctx->gres == 1.0) ? k_global_path<true> : k_global_path<false>;
    DYMU_CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<n, PATH_THREADS, smem, ctx->stream>>>(d_args);
    return DYMU_OK;
}
}  // namespace

extern "C" {

int dymu_extract_global_path(dymu_ctx* ctx, uint32_t slot, double x0, double y0, double tau,
                             uint32_t goal_i, uint32_t goal_j, double* out, uint32_t cap,
                             uint32_t* n_out, int* status)
{
    DYMU_GUARD(ctx);
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
    const size_t smem = 4 * PATCH * PATCH * sizeof(double);
    DYMU_TRY(launch_paths(ctx, 1, smem, (const PathArgs*)ctx->d_scratch));
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

int dymu_extract_global_path_batch(dymu_ctx* ctx, uint32_t n, const uint32_t* slots,
                                   const double* xy0, double tau, const uint32_t* goal_ij,
                                   double* out, uint32_t cap, uint32_t* n_out, int* status)
{
    DYMU_GUARD(ctx);
    if (!ctx || !slots || !xy0 || !goal_ij || !out || !n_out || !status || n == 0 || cap == 0)
        return DYMU_ERR_ARG;
    for (uint32_t q = 0; q < n; ++q)
        if (slots[q] >= ctx->n_slots || goal_ij[2 * q] >= ctx->nx || goal_ij[2 * q + 1] >= ctx->ny)
            return DYMU_ERR_ARG;
    const size_t hdr = (((size_t)n * (sizeof(PathArgs) + 8)) + 255) & ~(size_t)255;
    const size_t out_bytes = (size_t)n * cap * 5 * sizeof(double);
    DYMU_TRY(dymu_internal_scratch(ctx, hdr + out_bytes, hdr));
    PathArgs* h_args = (PathArgs*)ctx->h_pinned;
    char* d_base = (char*)ctx->d_scratch;
    uint32_t* d_res = (uint32_t*)(d_base + (size_t)n * sizeof(PathArgs));
    for (uint32_t q = 0; q < n; ++q)
    {
        PathArgs a;
        a.T = ctx->T + (size_t)slots[q] * ctx->pitch * ctx->rows;
        a.elev = ctx->elev;
        a.pitch = ctx->pitch; a.nx = ctx->nx; a.ny = ctx->ny;
        a.gres = ctx->gres; a.tau = tau; a.x0 = xy0[2 * q]; a.y0 = xy0[2 * q + 1];
        a.goal_i = goal_ij[2 * q]; a.goal_j = goal_ij[2 * q + 1];
        a.out = (double*)(d_base + hdr) + (size_t)q * cap * 5;
        a.cap = cap;
        a.result = d_res + 2 * q;
        h_args[q] = a;
    }
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(d_base, h_args, (size_t)n * sizeof(PathArgs), cudaMemcpyHostToDevice,
                                       ctx->stream));
    const size_t smem = 4 * PATCH * PATCH * sizeof(double);
    DYMU_TRY(launch_paths(ctx, n, smem, (const PathArgs*)d_base));
    ctx->launches++;
    DYMU_CUDA_TRY(ctx, cudaGetLastError());
    uint32_t* h_res = (uint32_t*)((char*)ctx->h_pinned + (size_t)n * sizeof(PathArgs));
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(h_res, d_res, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t q = 0; q < n; ++q)
    {
        n_out[q] = h_res[2 * q];
        status[q] = (int)h_res[2 * q + 1];
        if (n_out[q])
            DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(out + (size_t)q * cap * 5,
                                               (double*)(d_base + hdr) + (size_t)q * cap * 5,
                                               (size_t)n_out[q] * 5 * sizeof(double), cudaMemcpyDeviceToHost,
                                               ctx->stream));
    }
    DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DYMU_OK;
}

}  // extern "C"
