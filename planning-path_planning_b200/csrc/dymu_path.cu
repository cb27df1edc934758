// dymu_path.cu -- gradient-descent waypoint extraction on the device-resident total-cost
// plane (reference: computeGlobalPath / computeNextGlobalWaypoint / gradientNode /
// interpolate, src/DyMu_GlobalPathPlanning.cpp:615-784).
//
// The descent is a sequential chain (each waypoint depends on the previous one), so it is
// latency-bound, not bandwidth-bound: one warp walks one path.  The warp keeps a
// PATCH x PATCH window of T and of the elevation in shared memory and re-centres it only
// when the 4x4 stencil of the current cell leaves it (a 0.4-cell step stays inside for
// ~100 steps), so the per-step loads are shared-memory hits.  Lanes 0..3 evaluate the four
// cell-corner gradients in parallel, the results are exchanged with warp shuffles and every
// lane carries the identical waypoint state, so there is no divergence; lane 0 stores the
// waypoint.  Batches of queries run one warp (one CTA) per query.
#include <limits.h>

#include "dymu_ctx.cuh"

namespace
{
constexpr int PATCH = 64;  // cells per patch edge: 2 x 32 KB of shared memory

struct Patch
{
    double* t;   // PATCH x PATCH total cost
    double* e;   // PATCH x PATCH elevation
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

// makes sure cells [cx-1, cx+2] x [cy-1, cy+2] are inside the shared-memory window
__device__ __forceinline__ void ensure_patch(const PathArgs& a, Patch& pt, PatchCmd* cmd, int lane, int cx,
                                             int cy)
{
    if (pt.valid && cx - 1 >= pt.x0 && cy - 1 >= pt.y0 && cx + 2 < pt.x0 + PATCH && cy + 2 < pt.y0 + PATCH)
        return;
    pt.x0 = cx - PATCH / 2;
    pt.y0 = cy - PATCH / 2;
    pt.valid = true;
    if (lane == 0)
    {
        cmd->x0 = pt.x0;
        cmd->y0 = pt.y0;
    }
    __syncthreads();  // helpers are parked at the matching barrier
    load_patch(a, pt.t, pt.e, pt.x0, pt.y0);
    __syncthreads();
}

__device__ __forceinline__ bool next_waypoint(const PathArgs& a, Patch& pt, PatchCmd* cmd, int lane,
                                              double wx, double wy, double& z, double& dCx,
                                              double& dCy, double& nx_, double& ny_)
{
    // x / 1.0 == x exactly: skip the two fp64 divisions for the usual global_res = 1
    const bool unit = (a.gres == 1.0);
    double gx = unit ? wx : wx / a.gres, gy = unit ? wy : wy / a.gres;
    if (!(gx >= 0.0) || !(gy >= 0.0) || !(gx < (double)(a.nx - 1)) || !(gy < (double)(a.ny - 1)))
        return false;
    uint32_t cx = (uint32_t)gx, cy = (uint32_t)gy;
    double da = gx - (double)cx, db = gy - (double)cy;
    ensure_patch(a, pt, cmd, lane, (int)cx, (int)cy);
    // lane c (0..3) owns corner (cx + (c&1), cy + (c>>1)): 0=n00 1=n10 2=n01 3=n11
    int c = lane & 3;
    uint32_t ci = cx + (uint32_t)(c & 1), cj = cy + (uint32_t)(c >> 1);
    double gxc, gyc;
    grad_node(a, pt, ci, cj, gxc, gyc);
    double ec = pt.e[((int)cj - pt.y0) * PATCH + ((int)ci - pt.x0)];
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
__global__ void __launch_bounds__(PATH_THREADS, 1) k_global_path(const PathArgs* args_arr)
{
    const PathArgs a = args_arr[blockIdx.x];
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(16) double path_smem[];
    __shared__ PatchCmd cmd;
    Patch pt;
    pt.t = path_smem;
    pt.e = path_smem + PATCH * PATCH;
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
            load_patch(a, pt.t, pt.e, x0, y0);
            __syncthreads();
        }
    }
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

    if (!next_waypoint(a, pt, &cmd, lane, wx, wy, z, dCx, dCy, nx_, ny_)) status = DYMU_PATH_OUTSIDE;
    else if (isnan(nx_) || isnan(ny_)) status = DYMU_PATH_NAN;
    else
    {
        push(wx, wy, z, dCx, dCy);
        wx = nx_;
        wy = ny_;
        while (dist2d(wx, wy, sx, sy) > 2.0 * a.gres)
        {
            if (!next_waypoint(a, pt, &cmd, lane, wx, wy, z, dCx, dCy, nx_, ny_))
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
        cmd.x0 = INT_MIN;  // release the helper warps
    }
    __syncthreads();
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
    const size_t smem = 2 * PATCH * PATCH * sizeof(double);
    DYMU_CUDA_TRY(ctx, cudaFuncSetAttribute(k_global_path, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    k_global_path<<<1, PATH_THREADS, smem, ctx->stream>>>((const PathArgs*)ctx->d_scratch);
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
    const size_t smem = 2 * PATCH * PATCH * sizeof(double);
    DYMU_CUDA_TRY(ctx, cudaFuncSetAttribute(k_global_path, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
    k_global_path<<<n, PATH_THREADS, smem, ctx->stream>>>((const PathArgs*)d_base);
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
