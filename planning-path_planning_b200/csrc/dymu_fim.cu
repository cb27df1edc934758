// dymu_fim.cu -- tiled Fast Iterative Method for the DyMu wave propagations.
//
// Replaces the reference's sequential Fast Marching loop (narrow band held in an
// unsorted std::vector with a linear argmin scan and vector::erase per pop,
// src/DyMu_GlobalPathPlanning.cpp:443-468, 551-568) by a fixed-point iteration of the
// very same node update (propagateGlobalNode, G.cpp:500-546):
//
//     Tx = min(T[i-1], T[i+1]);  Ty = min(T[j-1], T[j+1]);  C = ceff[j][i]
//     Tn = |Tx-Ty| < C && both finite ? (Tx+Ty+sqrt(2C^2-(Tx-Ty)^2))/2 : min(Tx,Ty)+C
//     T  = min(T, Tn)
//
// The FMM result is a fixed point of that update (a CLOSED node's value only depends on
// neighbours with smaller values), and the monotone iteration from T = +inf, T(goal) = 0
// converges to it from above, so both agree to rounding.
//
// Execution model (one persistent cooperative launch per solve):
//   * the plane is cut into TILE x TILE tiles; a global work list holds the ACTIVE tiles;
//   * each CTA repeatedly takes a tile, stages T (+1-cell halo) and the cost term C in
//     shared memory, relaxes it there until nothing changes (or an iteration cap), writes
//     T back and, for every tile edge whose cells changed, appends the neighbouring tile
//     to the next list (deduplicated by an atomic flag word whose bits say WHICH halo of
//     the neighbour went stale);
//   * inside a tile, work is tracked per 8x4-cell warp block: a warp only re-evaluates
//     blocks marked dirty (a neighbouring cell changed in the previous sweep), found with
//     __ballot_sync on the per-lane "improved" predicate; __syncthreads_or detects tile
//     convergence.  A wave crossing a tile therefore costs work proportional to the
//     front length, not the tile area;
//   * phases are separated by a grid-wide barrier; three rotating lists let one be
//     reset while the next is filled.
// The same kernel runs the local layer's risk dilation (propagateRisk,
// src/DyMu_LocalPathRepairing.cpp:550-576) in MODE 1, a max-propagation on risk.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "dymu_ctx.cuh"

namespace
{
template <int TILE> struct Cfg
{
    static constexpr int THREADS = TILE * TILE / 4;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int BX = TILE / 8;   // warp blocks per tile row (8 cells wide)
    static constexpr int BY = TILE / 4;   // warp blocks per tile column (4 cells tall)
    static constexpr int NBLK = BX * BY;
    static constexpr int BPW = NBLK / WARPS;  // = 4
    // row pitch of the shared arrays in doubles; PITCH % 16 == 8 makes the four 8-double
    // rows of a warp block fall into disjoint bank groups (conflict-free 64-bit access)
    static constexpr int PITCH = TILE + 8;
    static constexpr int CPT = TILE * TILE / THREADS;  // cells per thread in load/store = 4
    static constexpr int MIN_CTAS = (TILE == 32) ? 4 : 1;
    static constexpr size_t SMEM = sizeof(double) * ((TILE + 2) * PITCH + TILE * PITCH) + 2 * NBLK;
};

// warp block owned by (warp w, slot s): a bijection that spreads every row, column and
// diagonal of blocks over distinct warps, so a straight wave front keeps all warps busy
template <int TILE> __device__ __forceinline__ void block_of(int w, int s, int& bx, int& by);
template <> __device__ __forceinline__ void block_of<32>(int w, int s, int& bx, int& by)
{
    bx = s;
    by = (w - 2 * s) & 7;
}
template <> __device__ __forceinline__ void block_of<64>(int w, int s, int& bx, int& by)
{
    bx = 2 * s + (w & 1);
    by = ((w >> 1) - 3 * bx) & 15;
}

constexpr uint32_t kLeftLanes = 0x01010101u, kRightLanes = 0x80808080u;
constexpr uint32_t kTopLanes = 0x000000FFu, kBottomLanes = 0xFF000000u;
// activation flag bits: which halo of the tile is stale / whole tile must be re-evaluated
constexpr uint32_t kHaloTop = 1u, kHaloBottom = 2u, kHaloLeft = 4u, kHaloRight = 8u, kFull = 16u;

struct Params
{
    double* T;
    size_t slot_stride;
    const double* C;
    uint32_t pitch, rows, ntx, nty, nprob;
    uint32_t* list0;
    uint32_t* list1;
    uint32_t* list2;
    uint32_t* flag0;
    uint32_t* flag1;
    uint32_t* flag2;
    uint32_t* ctrl;
    unsigned long long* stats;
    int inner_cap, max_outer;
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p)
{
    return *reinterpret_cast<const volatile uint32_t*>(p);
}

// grid-wide barrier on a monotonically increasing arrival counter (all CTAs are
// co-resident: the kernel is launched with cudaLaunchCooperativeKernel)
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t& phase)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        uint32_t target = (phase + 1) * gridDim.x;
        atomicAdd(counter, 1u);
        while (ld_volatile_u32(counter) < target) { }
        __threadfence();
    }
    phase++;
    __syncthreads();
}

template <int MODE> __device__ __forceinline__ double outside_value()
{
    return MODE == 0 ? DYMU_INF : 0.0;  // missing neighbour: skipped (G.cpp:504-523) / risk 0 (L.cpp:555-558)
}

// one node update; returns true when the stored value improves
template <int MODE>
__device__ __forceinline__ bool relax(double tc, double tl, double tr, double tu, double td,
                                      double c, double& out)
{
    if (MODE == 0)
    {
        double Tx = fmin(tl, tr), Ty = fmin(td, tu);
        double Tn = dymu_eikonal(Tx, Ty, c);
        out = Tn;
        return Tn < tc;
    }
    else
    {
        // propagateRisk, L.cpp:550-576; obstacle cells are never targets (c = +inf)
        if (c == DYMU_INF) return false;
        double Ry = fmax(tu, td), Rx = fmax(tl, tr);
        double Sx = 1 - Rx, Sy = 1 - Ry, S;
        double d = Sx - Sy;
        if (fabs(d) < c) S = (Sx + Sy + sqrt(2 * (c * c) - (d * d))) / 2;
        else S = fmin(Sx, Sy) + c;
        double R = fmax(1 - S, 0.0);
        out = R;
        return (R > 0) && (R > tc);
    }
}

template <int TILE, int MODE>
__global__ void __launch_bounds__(Cfg<TILE>::THREADS, Cfg<TILE>::MIN_CTAS) k_fim(Params p)
{
    using K = Cfg<TILE>;
    constexpr int P = K::PITCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Ts = reinterpret_cast<double*>(smem_raw);            // (TILE+2) x P, 1-cell halo
    double* Cs = Ts + (TILE + 2) * P;                            // TILE x P
    uint8_t(*dirty)[K::NBLK] = reinterpret_cast<uint8_t(*)[K::NBLK]>(Cs + TILE * P);
    __shared__ uint32_t edge_changed[4];
    __shared__ uint32_t s_idx;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tiles_per_prob = p.ntx * p.nty;
    auto sel3 = [](uint32_t* a, uint32_t* b, uint32_t* c, int k) { return k == 0 ? a : (k == 1 ? b : c); };
    uint32_t phase = 0;
    unsigned long long n_tiles = 0, n_visits = 0;
    int outer = 0;
    bool converged = false;

    for (; outer < p.max_outer; ++outer)
    {
        const int cur = outer % 3, nxt = (outer + 1) % 3, old = (outer + 2) % 3;
        const uint32_t n_active = ld_volatile_u32(&p.ctrl[cur]);
        if (n_active == 0)
        {
            converged = true;
            break;
        }
        uint32_t* list_cur = sel3(p.list0, p.list1, p.list2, cur);
        uint32_t* list_nxt = sel3(p.list0, p.list1, p.list2, nxt);
        uint32_t* flag_cur = sel3(p.flag0, p.flag1, p.flag2, cur);
        uint32_t* flag_nxt = sel3(p.flag0, p.flag1, p.flag2, nxt);

        for (;;)
        {
            __syncthreads();
            if (tid == 0) s_idx = atomicAdd(&p.ctrl[3 + cur], 1u);
            __syncthreads();
            const uint32_t idx = s_idx;
            if (idx >= n_active) break;
            const uint32_t tile_id = ld_volatile_u32(&list_cur[idx]);
            const uint32_t prob = tile_id / tiles_per_prob;
            const uint32_t trem = tile_id - prob * tiles_per_prob;
            const uint32_t ty = trem / p.ntx, tx = trem - ty * p.ntx;
            const uint32_t why = ld_volatile_u32(&flag_cur[tile_id]);
            double* Tg = p.T + (size_t)prob * p.slot_stride + (size_t)ty * TILE * p.pitch
                         + (size_t)tx * TILE;
            const double* Cg = p.C + (size_t)ty * TILE * p.pitch + (size_t)tx * TILE;

            // ---- stage the tile: T interior + halo (L2-coherent loads: other CTAs write
            // T during this phase), C through the read-only path
            double told[K::CPT];
#pragma unroll
            for (int k = 0; k < K::CPT; ++k)
            {
                int e = tid + k * K::THREADS;
                int y = e / TILE, x = e % TILE;
                double v = __ldcg(&Tg[(size_t)y * p.pitch + x]);
                told[k] = v;
                Ts[(y + 1) * P + x + 1] = v;
                Cs[y * P + x] = __ldg(&Cg[(size_t)y * p.pitch + x]);
            }
            for (int e = tid; e < 4 * TILE; e += K::THREADS)
            {
                int side = e / TILE, q = e % TILE;
                double v = outside_value<MODE>();
                if (side == 0)
                {
                    if (ty > 0) v = __ldcg(&Tg[-(ptrdiff_t)p.pitch + q]);
                    Ts[q + 1] = v;
                }
                else if (side == 1)
                {
                    if (ty + 1 < p.nty) v = __ldcg(&Tg[(size_t)TILE * p.pitch + q]);
                    Ts[(TILE + 1) * P + q + 1] = v;
                }
                else if (side == 2)
                {
                    if (tx > 0) v = __ldcg(&Tg[(size_t)q * p.pitch - 1]);
                    Ts[(q + 1) * P] = v;
                }
                else
                {
                    if (tx + 1 < p.ntx) v = __ldcg(&Tg[(size_t)q * p.pitch + TILE]);
                    Ts[(q + 1) * P + TILE + 1] = v;
                }
            }
            // initial dirty set: only the warp blocks along the stale halos, or everything
            for (int b = tid; b < K::NBLK; b += K::THREADS)
            {
                int bx = b % K::BX, by = b / K::BX;
                bool d = (why & kFull) || ((why & kHaloTop) && by == 0)
                         || ((why & kHaloBottom) && by == K::BY - 1)
                         || ((why & kHaloLeft) && bx == 0) || ((why & kHaloRight) && bx == K::BX - 1);
                dirty[0][b] = d ? 1 : 0;
                dirty[1][b] = 0;
            }
            if (tid < 4) edge_changed[tid] = 0;
            if (tid == 0) flag_cur[tile_id] = 0;  // consumed; writers use flag_nxt this phase
            __syncthreads();

            // ---- relax in shared memory
            int buf = 0, it = 0, more = 1;
            while (more && it < p.inner_cap)
            {
                int any = 0;
#pragma unroll
                for (int s = 0; s < K::BPW; ++s)
                {
                    int bx, by;
                    block_of<TILE>(warp, s, bx, by);
                    const int b = by * K::BX + bx;
                    if (!dirty[buf][b]) continue;  // warp-uniform
                    __syncwarp();
                    if (lane == 0) dirty[buf][b] = 0;
                    const int x = bx * 8 + (lane & 7), y = by * 4 + (lane >> 3);
                    const int o = (y + 1) * P + x + 1;
                    double tn;
                    bool ch = relax<MODE>(Ts[o], Ts[o - 1], Ts[o + 1], Ts[o - P], Ts[o + P],
                                          Cs[y * P + x], tn);
                    if (ch) Ts[o] = tn;
                    const uint32_t m = __ballot_sync(0xffffffffu, ch);
                    n_visits++;
                    if (m)
                    {
                        any = 1;
                        if (lane == 0)
                        {
                            uint8_t* dn = dirty[buf ^ 1];
                            dn[b] = 1;
                            if (m & kLeftLanes) { if (bx > 0) dn[b - 1] = 1; else edge_changed[2] = 1; }
                            if (m & kRightLanes) { if (bx < K::BX - 1) dn[b + 1] = 1; else edge_changed[3] = 1; }
                            if (m & kTopLanes) { if (by > 0) dn[b - K::BX] = 1; else edge_changed[0] = 1; }
                            if (m & kBottomLanes) { if (by < K::BY - 1) dn[b + K::BX] = 1; else edge_changed[1] = 1; }
                        }
                    }
                }
                more = __syncthreads_or(any);
                buf ^= 1;
                ++it;
            }

            // ---- write back changed cells and wake the neighbours whose halo went stale
#pragma unroll
            for (int k = 0; k < K::CPT; ++k)
            {
                int e = tid + k * K::THREADS;
                int y = e / TILE, x = e % TILE;
                double v = Ts[(y + 1) * P + x + 1];
                if (v != told[k]) __stcg(&Tg[(size_t)y * p.pitch + x], v);
            }
            __threadfence();
            if (tid < 5)
            {
                uint32_t target = 0xffffffffu, bit = 0;
                if (tid == 0 && edge_changed[0] && ty > 0) { target = tile_id - p.ntx; bit = kHaloBottom; }
                if (tid == 1 && edge_changed[1] && ty + 1 < p.nty) { target = tile_id + p.ntx; bit = kHaloTop; }
                if (tid == 2 && edge_changed[2] && tx > 0) { target = tile_id - 1; bit = kHaloRight; }
                if (tid == 3 && edge_changed[3] && tx + 1 < p.ntx) { target = tile_id + 1; bit = kHaloLeft; }
                if (tid == 4 && more) { target = tile_id; bit = kFull; }  // cap hit: not converged
                if (target != 0xffffffffu)
                {
                    if (atomicOr(&flag_nxt[target], bit) == 0)
                    {
                        uint32_t pos = atomicAdd(&p.ctrl[nxt], 1u);
                        list_nxt[pos] = target;
                    }
                }
            }
            n_tiles++;
        }
        if (blockIdx.x == 0 && tid == 0)
        {
            p.ctrl[old] = 0;      // count of the list that becomes "next" after this barrier
            p.ctrl[3 + old] = 0;  // and its cursor
        }
        grid_barrier(&p.ctrl[6], phase);
    }

    // ---- statistics
    if (lane == 0 && n_visits) atomicAdd(&p.stats[1], n_visits);
    if (tid == 0)
    {
        if (n_tiles) atomicAdd(&p.stats[0], n_tiles);
        if (blockIdx.x == 0)
        {
            p.stats[2] = (unsigned long long)outer;
            p.stats[3] = converged ? 1ull : 0ull;
        }
    }
}

__global__ void k_seed(double* T, size_t slot_stride, uint32_t pitch, uint32_t ntx, uint32_t nty,
                       int tile, const uint32_t* goal_ij, uint32_t n, uint32_t* list0,
                       uint32_t* flag0, uint32_t* ctrl)
{
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t gi = goal_ij[2 * q], gj = goal_ij[2 * q + 1];
    T[(size_t)q * slot_stride + (size_t)gj * pitch + gi] = 0.0;  // resetGlobalNarrowBand, G.cpp:490-496
    uint32_t tile_id = q * ntx * nty + (gj / tile) * ntx + gi / tile;
    flag0[tile_id] = kFull;
    list0[q] = tile_id;
    if (q == 0) ctrl[0] = n;
}

__global__ void k_seed_rows(uint32_t ntx, uint32_t ty0, uint32_t ty1, uint32_t* list0,
                            uint32_t* flag0, uint32_t* ctrl)
{
    uint32_t n = (ty1 - ty0) * ntx;
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t tile_id = ty0 * ntx + q;
    flag0[tile_id] = kFull;
    list0[q] = tile_id;
    if (q == 0) ctrl[0] = n;
}

template <int TILE, int MODE> int launch_fim(dymu_ctx* ctx, Params& prm, size_t total_tiles)
{
    // persistent grid = every CTA that can be co-resident (required by the grid barrier)
    int per_sm = 0;
    DYMU_CUDA_TRY(ctx, cudaFuncSetAttribute(k_fim<TILE, MODE>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)Cfg<TILE>::SMEM));
    DYMU_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                           &per_sm, k_fim<TILE, MODE>, Cfg<TILE>::THREADS, Cfg<TILE>::SMEM));
    if (per_sm < 1) DYMU_FAIL(ctx, DYMU_ERR_CUDA, "FIM kernel does not fit on an SM");
    if (ctx->fim_grid_per_sm > 0 && ctx->fim_grid_per_sm < per_sm) per_sm = ctx->fim_grid_per_sm;
    size_t grid = (size_t)per_sm * ctx->sm_count;
    if (grid > total_tiles) grid = total_tiles;
    if (grid < 1) grid = 1;
    void* args[] = {&prm};
    DYMU_CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void*)k_fim<TILE, MODE>, dim3((unsigned)grid),
                                                   dim3(Cfg<TILE>::THREADS), args,
                                                   Cfg<TILE>::SMEM, ctx->stream));
    ctx->launches++;
    return DYMU_OK;
}
}  // namespace

int dymu_internal_fim_alloc(dymu_ctx* ctx, dymu_fim_work* w, size_t capacity)
{
    dymu_internal_fim_free(w);
    w->capacity = capacity;
    for (int k = 0; k < 3; ++k)
    {
        DYMU_CUDA_TRY(ctx, cudaMalloc((void**)&w->list[k], capacity * sizeof(uint32_t)));
        DYMU_CUDA_TRY(ctx, cudaMalloc((void**)&w->flag[k], capacity * sizeof(uint32_t)));
        DYMU_CUDA_TRY(ctx, cudaMemsetAsync(w->flag[k], 0, capacity * sizeof(uint32_t), ctx->stream));
    }
    DYMU_CUDA_TRY(ctx, cudaMalloc((void**)&w->ctrl, 8 * sizeof(uint32_t)));
    DYMU_CUDA_TRY(ctx, cudaMalloc((void**)&w->stats, 4 * sizeof(unsigned long long)));
    return DYMU_OK;
}

void dymu_internal_fim_free(dymu_fim_work* w)
{
    for (int k = 0; k < 3; ++k)
    {
        if (w->list[k]) cudaFree(w->list[k]);
        if (w->flag[k]) cudaFree(w->flag[k]);
        w->list[k] = w->flag[k] = nullptr;
    }
    if (w->ctrl) cudaFree(w->ctrl);
    if (w->stats) cudaFree(w->stats);
    w->ctrl = nullptr;
    w->stats = nullptr;
    w->capacity = 0;
}

int dymu_internal_fim_configure(dymu_ctx* ctx)
{
    int dev_coop = 0;
    DYMU_CUDA_TRY(ctx, cudaDeviceGetAttribute(&dev_coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!dev_coop) DYMU_FAIL(ctx, DYMU_ERR_NODEVICE, "device lacks cooperative launch");
    ctx->fim_inner_cap = (int)ctx->tile * 2;
    if (const char* e = getenv("DYMU_FIM_INNER"))
        if (atoi(e) > 0) ctx->fim_inner_cap = atoi(e);
    ctx->fim_grid_per_sm = 0;  // 0 = as many CTAs per SM as fit
    if (const char* e = getenv("DYMU_FIM_GRID_PER_SM"))
        if (atoi(e) > 0) ctx->fim_grid_per_sm = atoi(e);
    ctx->fim_max_outer = 0;  // derived per launch
    return DYMU_OK;
}

// Runs the persistent kernel on a prepared work list and collects statistics.
int dymu_internal_fim_run(dymu_ctx* ctx, const dymu_fim_launch& L, dymu_solve_stats* stats)
{
    dymu_fim_work* w = L.work;
    Params prm;
    prm.T = L.T;
    prm.slot_stride = L.slot_stride;
    prm.C = L.C;
    prm.pitch = L.pitch;
    prm.rows = L.rows;
    prm.ntx = L.ntx;
    prm.nty = L.nty;
    prm.nprob = L.nprob;
    prm.list0 = w->list[0]; prm.list1 = w->list[1]; prm.list2 = w->list[2];
    prm.flag0 = w->flag[0]; prm.flag1 = w->flag[1]; prm.flag2 = w->flag[2];
    prm.ctrl = w->ctrl;
    prm.stats = w->stats;
    prm.inner_cap = ctx->fim_inner_cap;
    // a wave needs at most ~(ntx+nty) tile hops in free space; obstacles lengthen the
    // geodesic, so leave two orders of magnitude of head room before reporting NOCONV
    prm.max_outer = 64 * (int)(L.ntx + L.nty) + 4096;
    if (const char* e = getenv("DYMU_FIM_MAX_OUTER"))
        if (atoi(e) > 0) prm.max_outer = atoi(e);
    size_t total_tiles = (size_t)L.ntx * L.nty * L.nprob;
    DYMU_CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    int rc;
    if (L.tile == 32) rc = (L.mode == 0) ? launch_fim<32, 0>(ctx, prm, total_tiles) : launch_fim<32, 1>(ctx, prm, total_tiles);
    else rc = (L.mode == 0) ? launch_fim<64, 0>(ctx, prm, total_tiles) : launch_fim<64, 1>(ctx, prm, total_tiles);
    DYMU_TRY(rc);
    DYMU_CUDA_TRY(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
    unsigned long long h[4];
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(h, w->stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats)
    {
        stats->tile_activations = h[0];
        stats->cell_updates = h[1] * 32ull;
        stats->outer_iterations = (uint32_t)h[2];
        stats->converged = (uint32_t)h[3];
        DYMU_CUDA_TRY(ctx, cudaEventElapsedTime(&stats->kernel_ms, ctx->ev1, ctx->ev2));
    }
    if (!h[3])
        DYMU_FAIL(ctx, DYMU_ERR_NOCONV, "tile FIM hit the outer-iteration cap (%d) before converging",
                  prm.max_outer);
    return DYMU_OK;
}

int dymu_internal_fill(dymu_ctx* ctx, double* p, double v, size_t n);

static int reset_work(dymu_ctx* ctx, dymu_fim_work* w)
{
    DYMU_CUDA_TRY(ctx, cudaMemsetAsync(w->ctrl, 0, 8 * sizeof(uint32_t), ctx->stream));
    DYMU_CUDA_TRY(ctx, cudaMemsetAsync(w->stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    return DYMU_OK;
}

extern "C" {

int dymu_reserve_slots(dymu_ctx* ctx, uint32_t n_slots)
{
    if (!ctx || n_slots < 1) return DYMU_ERR_ARG;
    if (n_slots == ctx->n_slots) return DYMU_OK;
    size_t n = (size_t)ctx->pitch * ctx->rows;
    double* T = nullptr;
    DYMU_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    DYMU_CUDA_TRY(ctx, cudaMalloc((void**)&T, n * n_slots * sizeof(double)));
    cudaFree(ctx->T);
    ctx->T = T;
    ctx->n_slots = n_slots;
    DYMU_TRY(dymu_internal_fill(ctx, ctx->T, 1.0 / 0.0, n * n_slots));
    DYMU_TRY(dymu_internal_fim_alloc(ctx, &ctx->work, (size_t)ctx->ntx * ctx->nty * n_slots));
    ctx->solved = false;
    return DYMU_OK;
}

int dymu_solve_total_cost(dymu_ctx* ctx, uint32_t n_goals, const uint32_t* goal_i,
                          const uint32_t* goal_j, dymu_solve_stats* stats)
{
    if (!ctx || !goal_i || !goal_j || n_goals < 1 || n_goals > ctx->n_slots) return DYMU_ERR_ARG;
    if (!ctx->have_cost) DYMU_FAIL(ctx, DYMU_ERR_STATE, "no cost map: call dymu_set_cost_map / dymu_compute_cost_map first");
    for (uint32_t q = 0; q < n_goals; ++q)
        if (goal_i[q] >= ctx->nx || goal_j[q] >= ctx->ny) DYMU_FAIL(ctx, DYMU_ERR_ARG, "goal %u outside the grid", q);
    DYMU_TRY(dymu_internal_refresh_ceff(ctx));
    size_t n = (size_t)ctx->pitch * ctx->rows;
    // resetTotalCostMap, G.cpp:473-485 (whole plane instead of the propagated-node list)
    DYMU_CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    DYMU_TRY(dymu_internal_fill(ctx, ctx->T, 1.0 / 0.0, n * n_goals));
    DYMU_TRY(reset_work(ctx, &ctx->work));
    DYMU_TRY(dymu_internal_scratch(ctx, (size_t)n_goals * 8, (size_t)n_goals * 8));
    uint32_t* h_goals = (uint32_t*)ctx->h_pinned;
    for (uint32_t q = 0; q < n_goals; ++q)
    {
        h_goals[2 * q] = goal_i[q];
        h_goals[2 * q + 1] = goal_j[q];
    }
    DYMU_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_scratch, h_goals, (size_t)n_goals * 8,
                                       cudaMemcpyHostToDevice, ctx->stream));
    k_seed<<<dymu_div_up(n_goals, 128), 128, 0, ctx->stream>>>(
        ctx->T, n, ctx->pitch, ctx->ntx, ctx->nty, (int)ctx->tile, (const uint32_t*)ctx->d_scratch,
        n_goals, ctx->work.list[0], ctx->work.flag[0], ctx->work.ctrl);
    ctx->launches++;
    DYMU_CUDA_TRY(ctx, cudaGetLastError());
    dymu_fim_launch L;
    L.T = ctx->T; L.slot_stride = n; L.C = ctx->ceff; L.pitch = ctx->pitch; L.rows = ctx->rows;
    L.ntx = ctx->ntx; L.nty = ctx->nty; L.nprob = n_goals; L.mode = 0; L.tile = (int)ctx->tile;
    L.work = &ctx->work; L.n_initial = n_goals;
    dymu_solve_stats local;
    memset(&local, 0, sizeof(local));
    int rc = dymu_internal_fim_run(ctx, L, &local);
    if (rc == DYMU_OK || rc == DYMU_ERR_NOCONV)
    {
        cudaEventElapsedTime(&local.reset_ms, ctx->ev0, ctx->ev1);
        if (stats) stats[0] = local;
    }
    ctx->solved = (rc == DYMU_OK);
    return rc;
}

int dymu_solve_resume(dymu_ctx* ctx, uint32_t j0, uint32_t j1, dymu_solve_stats* stats)
{
    if (!ctx || j0 >= j1 || j1 > ctx->ny) return DYMU_ERR_ARG;
    if (!ctx->have_cost) DYMU_FAIL(ctx, DYMU_ERR_STATE, "no cost map");
    DYMU_TRY(dymu_internal_refresh_ceff(ctx));
    DYMU_TRY(reset_work(ctx, &ctx->work));
    uint32_t ty0 = j0 / ctx->tile, ty1 = dymu_div_up(j1, ctx->tile);
    uint32_t ntl = (ty1 - ty0) * ctx->ntx;
    k_seed_rows<<<dymu_div_up(ntl, 128), 128, 0, ctx->stream>>>(ctx->ntx, ty0, ty1, ctx->work.list[0],
                                                                ctx->work.flag[0], ctx->work.ctrl);
    ctx->launches++;
    DYMU_CUDA_TRY(ctx, cudaGetLastError());
    dymu_fim_launch L;
    L.T = ctx->T; L.slot_stride = (size_t)ctx->pitch * ctx->rows; L.C = ctx->ceff;
    L.pitch = ctx->pitch; L.rows = ctx->rows; L.ntx = ctx->ntx; L.nty = ctx->nty; L.nprob = 1;
    L.mode = 0; L.tile = (int)ctx->tile; L.work = &ctx->work; L.n_initial = ntl;
    dymu_solve_stats local;
    memset(&local, 0, sizeof(local));
    int rc = dymu_internal_fim_run(ctx, L, &local);
    if (stats && (rc == DYMU_OK || rc == DYMU_ERR_NOCONV)) stats[0] = local;
    return rc;
}

int dymu_stop_threshold(dymu_ctx* ctx, uint32_t slot, uint32_t start_i, uint32_t start_j,
                        double* t_stop)
{
    if (!ctx || !t_stop || slot >= ctx->n_slots) return DYMU_ERR_ARG;
    if (start_i < 1 || start_j < 1 || start_i + 1 >= ctx->nx || start_j + 1 >= ctx->ny)
        return DYMU_ERR_ARG;
    uint32_t idx[5] = {start_j * ctx->nx + start_i, (start_j - 1) * ctx->nx + start_i,
                       start_j * ctx->nx + start_i - 1, start_j * ctx->nx + start_i + 1,
                       (start_j + 1) * ctx->nx + start_i};
    double v[5];
    DYMU_TRY(dymu_read_cells(ctx, DYMU_PLANE_TOTAL_COST, slot, idx, 5, v));
    double m = v[0];
    for (int k = 1; k < 5; ++k)
        if (v[k] > m) m = v[k];
    *t_stop = m;
    return DYMU_OK;
}

}  // extern "C"
