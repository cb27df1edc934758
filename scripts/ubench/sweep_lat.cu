// Latency anatomy of one tile-FIM sweep: a single CTA (512 threads) relaxes one 32x32 tile in
// shared memory like k_fim does, with one warp's block dirty, and clock64 times 64 sweeps.
// Variants strip one stage at a time to expose its marginal cost on the dependent chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o sweep_lat sweep_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double sqrt_normal(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(y * y), x, 1.0);
    const double t = fma(e, 0.375, 0.5);
    const double y1 = fma(t, y * e, y);
    const double s = y1 * x;
    const double h = y1 * 0.5;
    const double r = fma(-s, s, x);
    return fma(r, h, s);
}
__device__ __forceinline__ double min_nn(double a, double b)
{
    long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return __longlong_as_double(x < y ? x : y);
}

constexpr int P = 40, TILE = 32;

template <int VARIANT>
__global__ void __launch_bounds__(512, 2) k(double* out, long long* cycles, int sweeps, int dirty_warps)
{
    __shared__ double Ts[(TILE + 2) * P];
    __shared__ uint32_t dmask[3];
    __shared__ uint32_t edge_mask;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < (TILE + 2) * P; k += 512) Ts[k] = 1.0 / 0.0;
    if (tid == 0) { dmask[0] = (1u << dirty_warps) - 1; dmask[1] = dmask[2] = 0; edge_mask = 0; }
    __syncthreads();
    if (tid == 0) Ts[1 * P + 1] = 0.0;  // source in the corner
    const int bx = warp & 3, by = warp >> 2, lx = lane & 7, ly = lane >> 3;
    int off = (by * 8 + ly + 1) * P + bx * 8 + lx + 1;
    asm volatile("" : "+r"(off));
    double* a = Ts + off;
    double* b = a + 4 * P;
    const double c = 1.0 + 0.01 * lane, q = 2 * (c * c);
    const uint32_t my_bit = 1u << warp;
    __syncthreads();
    uint32_t m = dmask[0];
    int cur = 0;
    long long t0 = clock64();
    for (int it = 0; it < sweeps; ++it)
    {
        uint32_t* nxt = &dmask[(cur + 1) % 3];
        uint32_t* old = &dmask[(cur + 2) % 3];
        if (tid == 0) *old = 0;
        if (m & my_bit)
        {
            const double tA = a[0], lA = a[-1], rA = a[1], uA = a[-P], dA = a[P];
            const double tB = b[0], lB = b[-1], rB = b[1], uB = b[-P], dB = b[P];
            double nA, nB;
            bool chA, chB;
            {
                const double Tx = min_nn(lA, rA), Ty = min_nn(dA, uA), d = Tx - Ty;
                const bool two = fabs(d) < c;
                const double t2 = (VARIANT == 3) ? (Tx + Ty) * 0.5 : (Tx + Ty + sqrt_normal(q - d * d)) * 0.5;
                const double t1 = min_nn(Tx, Ty) + c;
                nA = two ? t2 : t1;
                chA = nA < tA;
            }
            {
                const double Tx = min_nn(lB, rB), Ty = min_nn(dB, uB), d = Tx - Ty;
                const bool two = fabs(d) < c;
                const double t2 = (VARIANT == 3) ? (Tx + Ty) * 0.5 : (Tx + Ty + sqrt_normal(q - d * d)) * 0.5;
                const double t1 = min_nn(Tx, Ty) + c;
                nB = two ? t2 : t1;
                chB = nB < tB;
            }
            if (chA) a[0] = nA;
            if (chB) b[0] = nB;
            if (VARIANT != 1)
            {
                const uint32_t all = __reduce_or_sync(0xffffffffu, (chA ? my_bit : 0u) | (chB ? my_bit : 0u));
                if (all != 0 && lane == 0) atomicOr(nxt, all | my_bit);
            }
            else if (lane == 0)
                *nxt = my_bit;  // no reduction, no atomic: plain store
        }
        if (VARIANT != 2) __syncthreads(); else __syncwarp();
        m = (VARIANT == 2) ? my_bit * (warp < dirty_warps) : *nxt;
        if (VARIANT == 4) m |= (1u << dirty_warps) - 1;  // keep everything dirty
        cur = (cur + 1) % 3;
    }
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 512 + tid] = a[0] + b[0];
}

template <int V> void run(const char* name, int dirty, int ctas)
{
    double* out; long long* cyc;
    cudaMalloc(&out, ctas * 512 * 8); cudaMalloc(&cyc, ctas * 8);
    const int sweeps = 64;
    k<V><<<ctas, 512>>>(out, cyc, sweeps, dirty);
    k<V><<<ctas, 512>>>(out, cyc, sweeps, dirty);
    long long h[512];
    cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
    printf("%-46s dirty warps %2d, CTAs %3d: %6.0f cycles per sweep (%s)\n", name, dirty, ctas, (double)h[0] / sweeps,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int ctas : {1, 296})
        for (int dirty : {1, 4, 16})
        {
            run<0>("full sweep", dirty, ctas);
            run<1>("  - REDUX and atomicOr (plain store)", dirty, ctas);
            run<2>("  - __syncthreads and mask read (syncwarp)", dirty, ctas);
            run<3>("  - sqrt", dirty, ctas);
            run<4>("full sweep, mask forced dirty", dirty, ctas);
        }
    return 0;
}
