/*
 * dymu_oracle.c -- CPU restatement of the DyMu total-cost propagation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a plain-C,
 * structure-of-arrays restatement of the algorithms in
 *   /root/reference/src/DyMu_GlobalPathPlanning.cpp   ("G.cpp")
 *   /root/reference/src/DyMu_LocalPathRepairing.cpp   ("L.cpp")
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; nothing in the product path links it.
 *
 * Pinning: the reference ships no tests or golden vectors (SURVEY.md section 4,
 * 8c), so this restatement is pinned against the reference ITSELF: the
 * unmodified reference sources are compiled into oracle/_ref/libdymu_ref.so
 * (oracle/Makefile) and tests/test_oracle_vs_reference.py checks every
 * function here against it (bit-exact for the literal restatements), plus
 * against the committed fixtures in tests/golden/ that were generated from
 * oracle/_ref by tests/golden/make_golden.py.
 *
 * Arithmetic is fp64 with the reference's expression order; compile with
 * -ffp-contract=off (no FMA), as the reference build (g++ -O2, baseline
 * x86-64) contains no fused operations.
 *
 * Layout: planes are row-major [j][i], index = j*nx + i.  The local layer is
 * held dense over the whole map: local cell (X, Y) = (gi*r + li, gj*r + lj);
 * a local cell "exists" iff has_local[parent]; a neighbour link is non-NULL
 * iff the neighbour is inside the map and its parent has a local map, which
 * is exactly the state the reference's lazy two-way linking in createLocalMap
 * (L.cpp:23-145) maintains.
 */
#define _GNU_SOURCE 1 /* M_PI, NAN under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF (1.0 / 0.0)
#define ORC_OPEN 0
#define ORC_CLOSED 1
#define ORC_CONSERVATIVE 0
#define ORC_SWEEPING 1

typedef struct
{
    uint32_t* v;
    size_t n, cap;
} u32vec;

static void vec_push(u32vec* a, uint32_t x)
{
    if (a->n == a->cap)
    {
        a->cap = a->cap ? 2 * a->cap : 1024;
        a->v = (uint32_t*)realloc(a->v, a->cap * sizeof(uint32_t));
    }
    a->v[a->n++] = x;
}
/* std::vector::erase(begin()+k): order-preserving */
static void vec_erase(u32vec* a, size_t k)
{
    memmove(a->v + k, a->v + k + 1, (a->n - k - 1) * sizeof(uint32_t));
    a->n--;
}

typedef struct
{
    double* v; /* packed x,y,z,heading */
    size_t n, cap;
} wpvec;

static void wp_reserve(wpvec* a, size_t n)
{
    if (n > a->cap)
    {
        while (a->cap < n) a->cap = a->cap ? 2 * a->cap : 256;
        a->v = (double*)realloc(a->v, a->cap * 4 * sizeof(double));
    }
}
static void wp_push(wpvec* a, const double* w)
{
    wp_reserve(a, a->n + 1);
    memcpy(a->v + 4 * a->n, w, 4 * sizeof(double));
    a->n++;
}
static void wp_insert_front(wpvec* a, const double* w, size_t count)
{
    wp_reserve(a, a->n + count);
    memmove(a->v + 4 * count, a->v, a->n * 4 * sizeof(double));
    memcpy(a->v, w, count * 4 * sizeof(double));
    a->n += count;
}
static void wp_erase_front(wpvec* a, size_t count)
{
    if (count > a->n) count = a->n;
    memmove(a->v, a->v + 4 * count, (a->n - count) * 4 * sizeof(double));
    a->n -= count;
}

typedef struct orc_ctx
{
    /* constructor arguments, G.cpp:22-33 */
    double risk_distance, reconnect_distance, risk_ratio;
    int approach;
    /* initGlobalLayer, G.cpp:39-50 */
    uint32_t nx, ny, r;
    double gres, lres, off[2];
    /* globalNode fields as planes, DyMu.hpp:69-108 */
    double *elev, *slope, *raw, *cost, *haz, *traff, *T;
    uint32_t* terrain;
    uint8_t *obst, *state, *has_local;
    int32_t* locmode; /* -1 = "DONT_CARE" */
    /* computeCostMap inputs kept as members, G.cpp:151-153 */
    double *lut, *slopes;
    int n_lut, n_slopes, n_locs;
    /* goal, G.cpp:354-355 */
    int has_goal;
    uint32_t goal_i, goal_j;
    double goal_heading;
    /* public vectors, DyMu.hpp:446-456 */
    u32vec gnb, gprop;
    u32vec lexp, lnb, lprop;
    wpvec path;
    int reconnecting_index;
    /* localNode fields as dense planes, DyMu.hpp:42-67 */
    uint64_t lnx, lny;
    double *risk, *dev, *ltot;
    uint8_t *lobst, *lstate;
    int64_t local_agent;
    /* instrumentation */
    uint64_t n_pops, n_updates;
} orc_ctx;

/* ------------------------------------------------------------------ */
/* construction                                                        */
/* ------------------------------------------------------------------ */

orc_ctx* orc_create(double risk_distance, double reconnect_distance, double risk_ratio,
                    int approach)
{
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->risk_distance = risk_distance;
    c->reconnect_distance = reconnect_distance;
    c->risk_ratio = risk_ratio;
    c->approach = approach;
    c->local_agent = -1;
    return c;
}

static void free_local(orc_ctx* c)
{
    free(c->risk);
    free(c->dev);
    free(c->ltot);
    free(c->lobst);
    free(c->lstate);
    c->risk = c->dev = c->ltot = NULL;
    c->lobst = c->lstate = NULL;
}

void orc_destroy(orc_ctx* c)
{
    if (!c) return;
    free(c->elev); free(c->slope); free(c->raw); free(c->cost); free(c->haz); free(c->traff);
    free(c->T); free(c->terrain); free(c->obst); free(c->state); free(c->has_local);
    free(c->locmode); free(c->lut); free(c->slopes);
    free(c->gnb.v); free(c->gprop.v); free(c->lexp.v); free(c->lnb.v); free(c->lprop.v);
    free(c->path.v);
    free_local(c);
    free(c);
}

/* initGlobalLayer, G.cpp:39-104 + globalNode ctor DyMu.hpp:88-107.
 * (elevation/slope/terrain are left uninitialised by the reference; 0 here.) */
int orc_init_global_layer(orc_ctx* c, double gres, double lres, uint32_t nx, uint32_t ny,
                          double offx, double offy)
{
    size_t n = (size_t)nx * ny;
    c->gres = gres; c->lres = lres; c->nx = nx; c->ny = ny;
    c->r = (uint32_t)(gres / lres); /* G.cpp:49 */
    c->off[0] = offx; c->off[1] = offy;
    c->elev = (double*)calloc(n, 8); c->slope = (double*)calloc(n, 8);
    c->raw = (double*)calloc(n, 8); c->cost = (double*)calloc(n, 8);
    c->haz = (double*)calloc(n, 8); c->traff = (double*)malloc(n * 8);
    c->T = (double*)malloc(n * 8); c->terrain = (uint32_t*)calloc(n, 4);
    c->obst = (uint8_t*)calloc(n, 1); c->state = (uint8_t*)calloc(n, 1);
    c->has_local = (uint8_t*)calloc(n, 1); c->locmode = (int32_t*)malloc(n * 4);
    for (size_t k = 0; k < n; ++k) { c->traff[k] = 1.0; c->T[k] = ORC_INF; c->locmode[k] = -1; }
    return 1;
}

/* getGlobalNode, G.cpp:313-317: unsigned wrap-around makes (i-1) at i==0 out of range */
static inline int64_t gnode(const orc_ctx* c, uint32_t i, uint32_t j)
{
    if (i >= c->nx || j >= c->ny) return -1;
    return (int64_t)j * c->nx + i;
}
/* nb4List order, G.cpp:76-80: (i,j-1), (i-1,j), (i+1,j), (i,j+1) */
static inline int64_t gnb4(const orc_ctx* c, int64_t k, int d)
{
    uint32_t i = (uint32_t)(k % c->nx), j = (uint32_t)(k / c->nx);
    switch (d)
    {
        case 0: return gnode(c, i, j - 1);
        case 1: return gnode(c, i - 1, j);
        case 2: return gnode(c, i + 1, j);
        default: return gnode(c, i, j + 1);
    }
}
/* nb8List order, G.cpp:92-99: CCW starting at (i+1,j) */
static inline int64_t gnb8(const orc_ctx* c, int64_t k, int d)
{
    static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    static const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    uint32_t i = (uint32_t)(k % c->nx), j = (uint32_t)(k / c->nx);
    return gnode(c, i + (uint32_t)dx[d], j + (uint32_t)dy[d]);
}
/* getNearestGlobalNode, G.cpp:572-584 */
static inline int64_t gnearest(const orc_ctx* c, double x, double y)
{
    return gnode(c, (uint32_t)(x / c->gres + 0.5), (uint32_t)(y / c->gres + 0.5));
}

/* ------------------------------------------------------------------ */
/* cost map                                                            */
/* ------------------------------------------------------------------ */

/* setCostMap, G.cpp:109-126 */
int orc_set_cost_map(orc_ctx* c, const double* cost, uint32_t ny, uint32_t nx)
{
    if (ny != c->ny || nx != c->nx) return 0;
    size_t n = (size_t)nx * ny;
    for (size_t k = 0; k < n; ++k)
    {
        c->cost[k] = cost[k];
        if (cost[k] <= 0)
        {
            c->obst[k] = 1;
            c->traff[k] = 0.0;
            c->haz[k] = 1.0;
        }
    }
    return 1;
}

/* calculateSlope, G.cpp:186-210 */
static void calc_slope(orc_ctx* c, int64_t k)
{
    double dx, dy;
    int64_t n0 = gnb4(c, k, 0), n1 = gnb4(c, k, 1), n2 = gnb4(c, k, 2), n3 = gnb4(c, k, 3);
    if (n1 < 0) dx = (c->elev[n2] - c->elev[k]) / c->gres;
    else if (n2 < 0) dx = (c->elev[k] - c->elev[n1]) / c->gres;
    else dx = (c->elev[n2] - c->elev[n1]) * 0.5 / c->gres;
    if (n0 < 0) dy = (c->elev[n3] - c->elev[k]) / c->gres;
    else if (n3 < 0) dy = (c->elev[k] - c->elev[n0]) / c->gres;
    else dy = (c->elev[n3] - c->elev[n0]) * 0.5 / c->gres;
    c->slope[k] = atan(sqrt(dx * dx + dy * dy)); /* pow(x,2) == x*x exactly */
}

/* calculateNominalCost, G.cpp:217-293 (dead neighbour loops 229-233/255-259 omitted:
 * their guard `!nodeTarget->isObstacle` is false right after isObstacle = true). */
static void calc_nominal_cost(orc_ctx* c, int64_t k, int range, int numLocs, double Cmax)
{
    double Cdef, Ccand, C1, C2;
    uint32_t terr = c->terrain[k];
    if (terr == 0)
    {
        c->raw[k] = Cmax;
        c->obst[k] = 1;
    }
    else if (range == 1)
    {
        Cdef = c->lut[terr * numLocs];
        for (int i = 0; i < numLocs; ++i)
        {
            Ccand = c->lut[terr * numLocs + i];
            if (Ccand < Cdef) Cdef = Ccand;
        }
        c->raw[k] = fmax(c->raw[k], Cdef); /* std::max on non-NaN doubles */
    }
    else
    {
        double slope_index = (c->slope[k]) * 180 / M_PI
                             / (c->slopes[c->n_slopes - 1] - c->slopes[0])
                             * (double)(c->n_slopes - 1);
        if (slope_index > (double)(c->n_slopes - 1))
        {
            c->raw[k] = Cmax;
            c->obst[k] = 1;
        }
        else
        {
            double smin = floor(slope_index), smax = ceil(slope_index);
            Cdef = Cmax;
            if (numLocs > 1)
            {
                for (int i = 1; i < numLocs; ++i) /* starts at 1: G.cpp:268 */
                {
                    C1 = c->lut[terr * range * numLocs + i * range + (int)smin];
                    C2 = c->lut[terr * range * numLocs + i * range + (int)smax];
                    Ccand = C1 + (C2 - C1) * (slope_index - smin);
                    if (Ccand < Cdef)
                    {
                        Cdef = Ccand;
                        c->raw[k] = fmax(c->raw[k], Cdef);
                        c->locmode[k] = i;
                    }
                }
            }
            else
            {
                C1 = c->lut[terr * range + (int)smin];
                C2 = c->lut[terr * range + (int)smax];
                Cdef = C1 + (C2 - C1) * (slope_index - smin);
                c->raw[k] = fmax(c->raw[k], Cdef);
                c->locmode[k] = 0;
            }
        }
    }
}

/* smoothCost, G.cpp:297-308: seeded with the node's CURRENT cost (quirk) */
static void smooth_cost(orc_ctx* c, int64_t k)
{
    double Csum = c->cost[k], n = 5;
    for (int d = 0; d < 4; ++d)
    {
        int64_t nb = gnb4(c, k, d);
        if (nb < 0) n--;
        else Csum += c->raw[nb];
    }
    c->cost[k] = Csum / n;
}

/* computeCostMap, G.cpp:145-181 */
int orc_compute_cost_map(orc_ctx* c, const double* lut, int n_lut, const double* slopes,
                         int n_slopes, int n_locs, const double* elevation,
                         const double* terrain)
{
    size_t n = (size_t)c->nx * c->ny;
    free(c->lut); free(c->slopes);
    c->lut = (double*)malloc(sizeof(double) * n_lut);
    memcpy(c->lut, lut, sizeof(double) * n_lut);
    c->slopes = (double*)malloc(sizeof(double) * n_slopes);
    memcpy(c->slopes, slopes, sizeof(double) * n_slopes);
    c->n_lut = n_lut; c->n_slopes = n_slopes; c->n_locs = n_locs;
    for (uint32_t j = 0; j < c->ny; ++j)
        for (uint32_t i = 0; i < c->nx; ++i)
        {
            size_t k = (size_t)j * c->nx + i;
            c->raw[k] = 0;
            c->elev[k] = elevation[k];
            if (i == 0 || j == 0 || i == c->nx - 1 || j == c->ny - 1) c->terrain[k] = 0;
            else c->terrain[k] = (uint32_t)terrain[k];
        }
    double Cmax = lut[0]; /* std::max_element, G.cpp:221 */
    for (int q = 1; q < n_lut; ++q) if (lut[q] > Cmax) Cmax = lut[q];
    for (size_t k = 0; k < n; ++k)
    {
        calc_slope(c, (int64_t)k);
        calc_nominal_cost(c, (int64_t)k, n_slopes, n_locs, Cmax);
        if (c->obst[k]) { c->traff[k] = 0.0; c->haz[k] = 1.0; }
    }
    for (size_t k = 0; k < n; ++k) smooth_cost(c, (int64_t)k);
    return 1;
}

/* setGoal, G.cpp:322-357 */
int orc_set_goal(orc_ctx* c, double x, double y, double heading)
{
    x = (x - c->off[0]) / c->gres;
    y = (y - c->off[1]) / c->gres;
    if (x < 0 || y < 0) return 0;
    uint32_t sx = (uint32_t)(x + 0.5), sy = (uint32_t)(y + 0.5);
    int64_t g = gnode(c, sx, sy);
    if (g < 0) return 0;
    for (int d = 0; d < 4; ++d) if (gnb4(c, g, d) < 0) return 0;
    if (c->obst[g]) return 0;
    for (int d = 0; d < 4; ++d) if (c->obst[gnb4(c, g, d)]) return 0;
    c->has_goal = 1; c->goal_i = sx; c->goal_j = sy; c->goal_heading = heading;
    return 1;
}

/* ------------------------------------------------------------------ */
/* global Fast Marching                                                */
/* ------------------------------------------------------------------ */

/* the eikonal update shared by propagateGlobalNode (G.cpp:527-535) and
 * propagateLocalNode (L.cpp:734-738) */
static inline double eikonal(double Tx, double Ty, double C)
{
    if ((fabs(Tx - Ty) < C) && (Tx < ORC_INF) && (Ty < ORC_INF))
        return (Tx + Ty + sqrt(2 * (C * C) - ((Tx - Ty) * (Tx - Ty)))) / 2;
    return fmin(Tx, Ty) + C;
}

/* effective cost term of G.cpp:527-528 */
static inline double ceff(const orc_ctx* c, int64_t k)
{
    return c->gres * (c->cost[k]) * (2 + c->haz[k] - c->traff[k]);
}

static inline double axis_min(const double* T, int64_t a, int64_t b)
{
    if (a >= 0 && b >= 0) return fmin(T[b], T[a]);
    if (a < 0) return T[b];
    return T[a];
}

/* propagateGlobalNode, G.cpp:500-546 */
static void propagate_global(orc_ctx* c, int64_t k)
{
    double Ty = axis_min(c->T, gnb4(c, k, 0), gnb4(c, k, 3));
    double Tx = axis_min(c->T, gnb4(c, k, 1), gnb4(c, k, 2));
    double Tn = eikonal(Tx, Ty, ceff(c, k));
    c->n_updates++;
    if (Tn < c->T[k])
    {
        if (c->T[k] == ORC_INF)
        {
            vec_push(&c->gprop, (uint32_t)k);
            vec_push(&c->gnb, (uint32_t)k);
        }
        c->T[k] = Tn;
    }
}

/* minCostGlobalNode, G.cpp:551-568: strict '<' linear scan, order-preserving erase */
static int64_t min_cost_global(orc_ctx* c)
{
    size_t idx = 0;
    double best = c->T[c->gnb.v[0]];
    for (size_t q = 0; q < c->gnb.n; ++q)
        if (c->T[c->gnb.v[q]] < best) { best = c->T[c->gnb.v[q]]; idx = q; }
    int64_t k = c->gnb.v[idx];
    vec_erase(&c->gnb, idx);
    return k;
}

/* resetTotalCostMap G.cpp:473-485 + resetGlobalNarrowBand G.cpp:490-496 */
static void reset_global(orc_ctx* c)
{
    for (size_t q = 0; q < c->gprop.n; ++q)
    {
        c->state[c->gprop.v[q]] = ORC_OPEN;
        c->T[c->gprop.v[q]] = ORC_INF;
    }
    c->gprop.n = 0;
    c->gnb.n = 0;
    int64_t g = gnode(c, c->goal_i, c->goal_j);
    vec_push(&c->gnb, (uint32_t)g);
    vec_push(&c->gprop, (uint32_t)g);
    c->T[g] = 0;
}

/* isSafeNode, G.cpp:410-422 */
static int is_safe(const orc_ctx* c, int64_t k)
{
    if (c->obst[k]) return 0;
    for (int d = 0; d < 8; ++d) if (c->obst[gnb8(c, k, d)]) return 0;
    return 1;
}
/* isFullyClosedNode, G.cpp:424-436 */
static int is_fully_closed(const orc_ctx* c, int64_t k)
{
    if (c->state[k] == ORC_OPEN) return 0;
    for (int d = 0; d < 4; ++d) if (c->state[gnb4(c, k, d)] == ORC_OPEN) return 0;
    return 1;
}

static void close_and_spread(orc_ctx* c, int64_t k)
{
    c->state[k] = ORC_CLOSED;
    c->n_pops++;
    for (int d = 0; d < 4; ++d)
    {
        int64_t nb = gnb4(c, k, d);
        if (nb >= 0 && c->state[nb] == ORC_OPEN && !c->obst[nb]) propagate_global(c, nb);
    }
}

/* computeEntireTotalCostMap, G.cpp:443-468 -- literal narrow-band vector */
int orc_compute_entire_total_cost_map(orc_ctx* c)
{
    if (!c->has_goal || c->obst[gnode(c, c->goal_i, c->goal_j)]) return 0;
    reset_global(c);
    while (c->gnb.n) close_and_spread(c, min_cost_global(c));
    return 1;
}

/* computeTotalCostMap, G.cpp:364-408 -- literal, with the early stop */
int orc_compute_total_cost_map(orc_ctx* c, double x, double y)
{
    x -= c->off[0]; y -= c->off[1];
    if (!c->has_goal || c->obst[gnode(c, c->goal_i, c->goal_j)]) return 0;
    int64_t s = gnearest(c, x, y);
    if (!is_safe(c, s)) return 0;
    reset_global(c);
    while (c->gnb.n && !is_fully_closed(c, s)) close_and_spread(c, min_cost_global(c));
    return c->gnb.n ? 1 : 0;
}

/* ---- heap-ordered variant ("fast oracle") ----
 * Same update expression and acceptance rule; the narrow band is a binary
 * min-heap keyed on (T, insertion sequence) with decrease-key, so pops come
 * in non-decreasing T like the reference's linear argmin.  Ties may pop in a
 * different order than the reference's vector position rule; the converged
 * values agree to rounding (checked against the literal version in tests).
 * stop_i/stop_j < 0 => full solve; otherwise stop when that node and its four
 * neighbours are CLOSED (G.cpp:390). */
typedef struct { double key; uint64_t seq; uint32_t node; } hent;
typedef struct { hent* h; size_t n, cap; int64_t* pos; uint64_t seq; } heap_t;

static int hless(const hent* a, const hent* b)
{
    return a->key < b->key || (a->key == b->key && a->seq < b->seq);
}
static void hswap(heap_t* H, size_t a, size_t b)
{
    hent t = H->h[a]; H->h[a] = H->h[b]; H->h[b] = t;
    H->pos[H->h[a].node] = (int64_t)a; H->pos[H->h[b].node] = (int64_t)b;
}
static void hup(heap_t* H, size_t i)
{
    while (i > 0) { size_t p = (i - 1) / 2; if (!hless(&H->h[i], &H->h[p])) break; hswap(H, i, p); i = p; }
}
static void hdown(heap_t* H, size_t i)
{
    for (;;)
    {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < H->n && hless(&H->h[l], &H->h[m])) m = l;
        if (r < H->n && hless(&H->h[r], &H->h[m])) m = r;
        if (m == i) break;
        hswap(H, i, m); i = m;
    }
}
static void hpush_or_decrease(heap_t* H, uint32_t node, double key)
{
    if (H->pos[node] >= 0)
    {
        size_t i = (size_t)H->pos[node];
        H->h[i].key = key; /* keeps its original seq: position in the vector is unchanged */
        hup(H, i);
        return;
    }
    if (H->n == H->cap) { H->cap = H->cap ? 2 * H->cap : 4096; H->h = (hent*)realloc(H->h, H->cap * sizeof(hent)); }
    H->h[H->n].key = key; H->h[H->n].seq = H->seq++; H->h[H->n].node = node;
    H->pos[node] = (int64_t)H->n; H->n++;
    hup(H, H->n - 1);
}
static uint32_t hpop(heap_t* H)
{
    uint32_t node = H->h[0].node;
    H->pos[node] = -1;
    H->n--;
    if (H->n) { H->h[0] = H->h[H->n]; H->pos[H->h[0].node] = 0; hdown(H, 0); }
    return node;
}

int orc_solve_heap(orc_ctx* c, int64_t stop_i, int64_t stop_j)
{
    size_t n = (size_t)c->nx * c->ny;
    if (!c->has_goal || c->obst[gnode(c, c->goal_i, c->goal_j)]) return 0;
    int64_t s = -1;
    if (stop_i >= 0) { s = gnode(c, (uint32_t)stop_i, (uint32_t)stop_j); if (s < 0 || !is_safe(c, s)) return 0; }
    /* full reset (the port does not need the propagated-node list for this) */
    for (size_t k = 0; k < n; ++k) { c->T[k] = ORC_INF; c->state[k] = ORC_OPEN; }
    c->gprop.n = 0; c->gnb.n = 0;
    heap_t H; memset(&H, 0, sizeof(H));
    H.pos = (int64_t*)malloc(n * sizeof(int64_t));
    for (size_t k = 0; k < n; ++k) H.pos[k] = -1;
    int64_t g = gnode(c, c->goal_i, c->goal_j);
    c->T[g] = 0; hpush_or_decrease(&H, (uint32_t)g, 0.0);
    while (H.n && !(s >= 0 && is_fully_closed(c, s)))
    {
        int64_t k = hpop(&H);
        c->state[k] = ORC_CLOSED; c->n_pops++;
        for (int d = 0; d < 4; ++d)
        {
            int64_t nb = gnb4(c, k, d);
            if (nb < 0 || c->state[nb] != ORC_OPEN || c->obst[nb]) continue;
            double Ty = axis_min(c->T, gnb4(c, nb, 0), gnb4(c, nb, 3));
            double Tx = axis_min(c->T, gnb4(c, nb, 1), gnb4(c, nb, 2));
            double Tn = eikonal(Tx, Ty, ceff(c, nb));
            c->n_updates++;
            if (Tn < c->T[nb]) { c->T[nb] = Tn; hpush_or_decrease(&H, (uint32_t)nb, Tn); }
        }
    }
    int nonempty = H.n > 0;
    free(H.h); free(H.pos);
    return (s >= 0) ? nonempty : 1;
}

/* one Jacobi sweep of the update over all free non-goal cells; returns the
 * number of cells whose value would still decrease (fixed-point residual).
 * Used by property tests: a converged map must return 0. */
uint64_t orc_fixed_point_violations(const orc_ctx* c, const double* T, double rel_tol)
{
    uint64_t bad = 0;
    size_t n = (size_t)c->nx * c->ny;
    int64_t g = c->has_goal ? gnode(c, c->goal_i, c->goal_j) : -1;
    for (size_t k = 0; k < n; ++k)
    {
        if (c->obst[k] || (int64_t)k == g) continue;
        double Ty = axis_min(T, gnb4(c, (int64_t)k, 0), gnb4(c, (int64_t)k, 3));
        double Tx = axis_min(T, gnb4(c, (int64_t)k, 1), gnb4(c, (int64_t)k, 2));
        double Tn = eikonal(Tx, Ty, ceff(c, (int64_t)k));
        if (Tn == ORC_INF && T[k] == ORC_INF) continue;
        if (fabs(Tn - T[k]) > rel_tol * fabs(Tn)) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------ */
/* global path extraction                                              */
/* ------------------------------------------------------------------ */

/* interpolate, G.cpp:776-784 */
static inline double interp(double a, double b, double g00, double g01, double g10, double g11)
{
    return g00 + (g10 - g00) * a + (g01 - g00) * b + (g11 + g00 - g10 - g01) * a * b;
}

/* the per-axis rule of gradientNode (G.cpp:722-741 / L.cpp:983-1001); lo/hi < 0 = NULL */
static inline double grad_axis(const double* F, int64_t k, int64_t lo, int64_t hi)
{
    int lo_inf = (lo < 0) || (F[lo] == ORC_INF);
    int hi_inf = (hi < 0) || (F[hi] == ORC_INF);
    if ((lo < 0 && hi < 0) || (lo >= 0 && hi >= 0 && F[lo] == ORC_INF && F[hi] == ORC_INF)) return 0;
    if (lo_inf)
    {
        if (hi < 0) return 0; /* lo non-NULL & inf, hi NULL: the reference dereferences NULL */
        return F[hi] - F[k];
    }
    if (hi_inf) return F[k] - F[lo];
    return (F[hi] - F[lo]) * 0.5;
}

/* gradientNode(globalNode*), G.cpp:718-772 */
static void grad_global(const orc_ctx* c, int64_t k, double* dnx, double* dny)
{
    double dx = grad_axis(c->T, k, gnb4(c, k, 1), gnb4(c, k, 2));
    double dy = grad_axis(c->T, k, gnb4(c, k, 0), gnb4(c, k, 3));
    if (dx == 0 && dy == 0) { *dnx = 0; *dny = 0; }
    else
    {
        *dnx = dx / sqrt(dx * dx + dy * dy);
        *dny = dy / sqrt(dx * dx + dy * dy);
    }
}

/* computeNextGlobalWaypoint, G.cpp:666-714.  w = {x,y,z,heading}; writes z of
 * the CURRENT waypoint (swapped-corner quirk, G.cpp:699-704) and returns next. */
static void next_global_waypoint(const orc_ctx* c, double* w, double tau, double* next)
{
    double gx = w[0] / c->gres, gy = w[1] / c->gres;
    uint32_t cx = (uint32_t)gx, cy = (uint32_t)gy;
    double a = gx - (double)cx, b = gy - (double)cy;
    int64_t n00 = gnode(c, cx, cy);
    int64_t n10 = gnb4(c, n00, 2), n01 = gnb4(c, n00, 3), n11 = gnb4(c, n10, 3);
    double gx00, gx10, gx01, gx11, gy00, gy10, gy01, gy11;
    grad_global(c, n00, &gx00, &gy00);
    grad_global(c, n10, &gx10, &gy10);
    grad_global(c, n01, &gx01, &gy01);
    grad_global(c, n11, &gx11, &gy11);
    double dCx = interp(a, b, gx00, gx01, gx10, gx11);
    double dCy = interp(a, b, gy00, gy01, gy10, gy11);
    w[2] = interp(a, b, c->elev[n00], c->elev[n10], c->elev[n01], c->elev[n11]);
    next[0] = w[0] - c->gres * tau * dCx;
    next[1] = w[1] - c->gres * tau * dCy;
    next[2] = 0.0;
    next[3] = atan2(-dCy, -dCx);
}

/* computeGlobalPath, G.cpp:615-662.  (x, y) already offset-free.  max_steps
 * guards the reference's unbounded loop (quirk 8); returns 1 ok, 0 failure,
 * -1 if the guard tripped. */
int orc_compute_global_path(orc_ctx* c, double x, double y, double heading_in, int64_t max_steps)
{
    double sink[4], w[4], nx_[4];
    int64_t g = gnode(c, c->goal_i, c->goal_j);
    sink[0] = c->gres * (double)c->goal_i;
    sink[1] = c->gres * (double)c->goal_j;
    sink[2] = c->elev[g];
    sink[3] = c->goal_heading;
    c->path.n = 0;
    double tau = fmin(0.4, c->risk_distance);
    w[0] = x; w[1] = y; w[2] = 0.0; w[3] = heading_in;
    next_global_waypoint(c, w, tau, nx_);
    if (isnan(nx_[0]) || isnan(nx_[1])) return 0;
    wp_push(&c->path, w);
    memcpy(w, nx_, sizeof(w));
    int64_t steps = 0;
    while (sqrt((w[0] - sink[0]) * (w[0] - sink[0]) + (w[1] - sink[1]) * (w[1] - sink[1]))
           > 2.0 * c->gres)
    {
        if (++steps > max_steps) return -1;
        next_global_waypoint(c, w, tau, nx_);
        wp_push(&c->path, w);
        if (sqrt((w[0] - nx_[0]) * (w[0] - nx_[0]) + (w[1] - nx_[1]) * (w[1] - nx_[1]))
            < 0.01 * tau * c->gres)
            return 0;
        memcpy(w, nx_, sizeof(w));
    }
    wp_push(&c->path, sink);
    return 1;
}

/* getTotalCost(Waypoint), G.cpp:860-890 */
double orc_get_total_cost(const orc_ctx* c, double x, double y)
{
    x -= c->off[0]; y -= c->off[1];
    uint32_t i = (uint32_t)(x / c->gres), j = (uint32_t)(y / c->gres);
    double a = x - (double)i, b = y - (double)j;
    int64_t n00 = gnode(c, i, j);
    int64_t n10 = gnb4(c, n00, 2), n01 = gnb4(c, n00, 3);
    int64_t n11 = n10 >= 0 ? gnb4(c, n10, 3) : -1;
    if (n00 < 0 || n10 < 0 || n01 < 0 || n11 < 0 || c->state[n00] == ORC_OPEN
        || c->state[n10] == ORC_OPEN || c->state[n01] == ORC_OPEN || c->state[n11] == ORC_OPEN)
        return c->T[gnearest(c, x, y)];
    double w00 = c->T[n00], w10 = c->T[n10], w01 = c->T[n01], w11 = c->T[n11];
    return w00 + (w10 - w00) * a + (w01 - w00) * b + (w11 + w00 - w10 - w01) * a * b;
}

#include "dymu_oracle_local.inc"

/* ------------------------------------------------------------------ */
/* accessors                                                            */
/* ------------------------------------------------------------------ */
double* orc_plane(orc_ctx* c, int which)
{
    switch (which)
    {
        case 0: return c->elev;
        case 1: return c->slope;
        case 2: return c->raw;
        case 3: return c->cost;
        case 4: return c->haz;
        case 5: return c->traff;
        case 6: return c->T;
        default: return NULL;
    }
}
uint8_t* orc_plane_u8(orc_ctx* c, int which)
{
    switch (which)
    {
        case 0: return c->obst;
        case 1: return c->state;
        case 2: return c->has_local;
        default: return NULL;
    }
}
uint32_t* orc_terrain(orc_ctx* c) { return c->terrain; }
int32_t* orc_locmode(orc_ctx* c) { return c->locmode; }
size_t orc_path_size(const orc_ctx* c) { return c->path.n; }
const double* orc_path_data(const orc_ctx* c) { return c->path.v; }
uint64_t orc_counter(const orc_ctx* c, int which) { return which == 0 ? c->n_pops : c->n_updates; }
int orc_reconnecting_index(const orc_ctx* c) { return c->reconnecting_index; }
size_t orc_propagated_count(const orc_ctx* c) { return c->gprop.n; }
