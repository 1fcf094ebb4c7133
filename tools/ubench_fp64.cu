/*
 * ubench_fp64.cu -- FP64-pipe micro-benchmarks for B200 (sm_100a): the guides in this image were measured on
 * B300 (vestigial FP64), so the numbers the kernel design rests on are measured here.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench_fp64.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double rsqrt_fast(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    /* two Newton steps: y <- y*(1.5 - 0.5*x*y*y) */
    double hx = 0.5 * x;
    double t = fma(-hx * y, y, 0.5);
    y = fma(y, t, y);
    t = fma(-hx * y, y, 0.5);
    y = fma(y, t, y);
    return y;
}

template <int ILP>
__global__ void k_dfma(double* out, int iters, long long* cyc)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 123.456) out[0] = s;
    if (cyc && threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

/* OP: 0 rcp (1.0/x), 1 div, 2 sqrt, 3 rcp_fast, 4 rsqrt_fast, 5 dadd, 6 dmul, 7 x*rsqrt_fast(x) */
template <int OP, int ILP>
__global__ void k_op(double* out, int iters, double seed)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = seed + threadIdx.x * 1e-6 + i * 0.01;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) a[i] = 1.0 / a[i];
            if (OP == 1) a[i] = 1.37 / a[i] + 0.5;   /* div + add (keeps the value near 1.4) */
            if (OP == 2) a[i] = sqrt(a[i]) + 0.25;
            if (OP == 3) a[i] = rcp_fast(a[i]);
            if (OP == 4) a[i] = rsqrt_fast(a[i]) + 0.25;
            if (OP == 5) a[i] = a[i] + 1e-9;
            if (OP == 6) a[i] = a[i] * 1.0000001;
            if (OP == 7) a[i] = a[i] * rsqrt_fast(a[i]) + 0.25;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    if (s == 123.456) out[0] = s;
}

/* DFMA mixed with other pipes: MIX 0 none, 1 one LDS per DFMA, 2 one IMAD per DFMA, 3 two IADD/LOP per DFMA */
template <int MIX>
__global__ void k_mix(double* out, int iters)
{
    __shared__ double sm[1024];
    sm[threadIdx.x % 1024] = threadIdx.x;
    __syncthreads();
    double a[4];
    for (int i = 0; i < 4; i++) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    int idx = threadIdx.x, acc = threadIdx.x;
    double ld = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                a[i] = fma(a[i], m, c);
                if (MIX == 1) { ld += sm[(idx + u * 4 + i) & 1023]; }
                if (MIX == 2) { acc = acc * 3 + idx; }
                if (MIX == 3) { acc = (acc ^ idx) + u; idx = (idx + acc) | 1; }
            }
    }
    double s = a[0] + a[1] + a[2] + a[3] + ld + acc + idx;
    if (s == 123.456) out[0] = s;
}

template <class F>
double time_ms(F f, int reps = 3)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, clock %d MHz\n", prop.name, nsm, clk_khz / 1000);
    double* out;
    long long* cyc;
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&cyc, 64));

    /* 1. latency: one warp, dependent chain */
    {
        const int iters = 4096;
        k_dfma<1><<<1, 32>>>(out, iters, cyc);
        CK(cudaDeviceSynchronize());
        k_dfma<1><<<1, 32>>>(out, iters, cyc);
        long long h;
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DFMA dependent-chain latency: %.2f cycles\n", (double)h / (iters * 8.0));
        k_dfma<2><<<1, 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("  1 warp ILP2: %.2f cycles per DFMA\n", (double)h / (iters * 8.0 * 2));
        k_dfma<4><<<1, 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("  1 warp ILP4: %.2f cycles per DFMA\n", (double)h / (iters * 8.0 * 4));
        k_dfma<8><<<1, 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("  1 warp ILP8: %.2f cycles per DFMA\n", (double)h / (iters * 8.0 * 8));
    }
    /* 2. throughput vs warps/SMSP and ILP */
    printf("DFMA throughput, TFLOP/s (2 flop per FMA); rows: warps per SMSP, cols: ILP 1 2 4 8\n");
    const int iters = 2048;
    for (int wps : {1, 2, 4, 8, 16}) {
        const int threads = (wps >= 8) ? 1024 : wps * 128;
        const int bps = (wps >= 8) ? wps / 8 : 1;
        const int blocks = nsm * bps;
        printf("  w/SMSP %2d:", wps);
        double ms;
        ms = time_ms([&] { k_dfma<1><<<blocks, threads>>>(out, iters, nullptr); });
        printf(" %7.2f", 2.0 * 8 * 1 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dfma<2><<<blocks, threads>>>(out, iters, nullptr); });
        printf(" %7.2f", 2.0 * 8 * 2 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dfma<4><<<blocks, threads>>>(out, iters, nullptr); });
        printf(" %7.2f", 2.0 * 8 * 4 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, nullptr); });
        printf(" %7.2f\n", 2.0 * 8 * 8 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    /* 3. op costs in DFMA-issue equivalents (8 warps/SMSP, ILP 4): time relative to a DFMA kernel of the same shape */
    {
        const int blocks = nsm, threads = 1024, it = 4096;
        const double ops = 4.0 * it * (double)blocks * threads;
        const double ms_fma = time_ms([&] { k_op<5, 4><<<blocks, threads>>>(out, it, 1.3); });
        const double per_fma = ms_fma / ops;
        printf("op cost in DADD-issue equivalents (8 w/SMSP, ILP4); DADD rate = %.2f Tinst/s\n", ops / (ms_fma * 1e-3) / 1e12);
        double ms;
        ms = time_ms([&] { k_op<6, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  dmul        %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<0, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  1.0/x       %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<1, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  a/x + c     %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<2, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  sqrt + c    %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<3, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  rcp_fast    %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<4, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  rsqrt_fast+c %.2f\n", ms / ops / per_fma);
        ms = time_ms([&] { k_op<7, 4><<<blocks, threads>>>(out, it, 1.3); }); printf("  x*rsqrt_fast+c %.2f\n", ms / ops / per_fma);
    }
    /* 4. co-issue with other pipes */
    {
        const int blocks = nsm, threads = 1024, it = 2048;
        const double fl = 2.0 * 8 * 4 * it * (double)blocks * threads;
        double ms;
        ms = time_ms([&] { k_mix<0><<<blocks, threads>>>(out, it); }); printf("mix none : %.2f TFLOP/s\n", fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_mix<1><<<blocks, threads>>>(out, it); }); printf("mix 1 LDS per DFMA : %.2f TFLOP/s\n", fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_mix<2><<<blocks, threads>>>(out, it); }); printf("mix 1 IMAD per DFMA: %.2f TFLOP/s\n", fl / (ms * 1e-3) / 1e12);
        ms = time_ms([&] { k_mix<3><<<blocks, threads>>>(out, it); }); printf("mix 4 int per DFMA : %.2f TFLOP/s\n", fl / (ms * 1e-3) / 1e12);
    }
    return 0;
}
