/*
 * fbench.cu -- isolates the FACE phase (characteristic projection + WCNS5-JS + HLLC) of the sweep kernel:
 * every thread evaluates face_midpoint_fast on a synthetic shared-memory window REPS times.  Reports the
 * face rate and the FP64-pipe utilisation implied by the static FP64 instruction count given on the command line.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I hamers_b200/csrc -o tools/bin/fbench tools/fbench.cu
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "hb2_sweep.cuh"

using namespace hb2;
using Tr = Traits<SS, 3, 1>;

#ifndef MINB
#define MINB 2
#endif

template <int DIR>
__global__ void __launch_bounds__(256, MINB) k_face(double* out, int reps, Consts K, int hyb_mask)
{
    using Sh = SweepShape<Tr, DIR>;
    extern __shared__ double smem[];
    /* fill the primitive ring with a smooth positive state */
    for (int i = threadIdx.x; i < Sh::NV * Sh::CSV; i += blockDim.x) {
        const int comp = i / Sh::CSV, r = i % Sh::CSV;
        const double x = 0.01 * (r % 97) + 0.1 * blockIdx.x;
        double v = 1.0 + 0.3 * sin(x + comp);
        if (comp == Tr::NEQ) v = 1.2 + 0.1 * cos(x);
        smem[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pp = (DIR == 0) ? w : lane, o = (DIR == 0) ? lane : w;
    double acc = 0.0;
    for (int r = 0; r < reps; r++) {
        const double* win = smem + Sh::slotv(pp, (o + r) & (Sh::RING - 1));
        double Fm[Tr::NEQ], um;
        face_midpoint_fast<Tr, DIR, Sh::CSV, Sh::MS>(win, (hyb_mask >> (r & 7)) & 1, K, Fm, um);
#pragma unroll
        for (int e = 0; e < Tr::NEQ; e++) acc += Fm[e];
    }
    if (acc == 123.456) out[0] = acc;
    out[1 + (blockIdx.x * 256 + threadIdx.x) % 1024] = acc;
}

int main(int argc, char** argv)
{
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    const double fp64_per_face = argc > 2 ? atof(argv[2]) : 650.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, 8 * 2048);
    Consts K;
    K.gamma[0] = 1.4; K.inv_gm1[0] = 2.5; K.weno_p = 2;
    using Sh = SweepShape<Tr, 1>;
    const size_t smem_min = Sh::NV * Sh::CSV * sizeof(double);
    cudaFuncSetAttribute(k_face<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_face<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_face<1>);
    printf("k_face<1>: %d regs, %zu B local, MINB %d\n", fa.numRegs, fa.localSizeBytes, MINB);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int bps : {1, 2, 3, 4, 6, 8}) {
        /* force `bps` resident blocks per SM through the dynamic shared memory size */
        size_t smem = (size_t)(220 * 1024 / bps) - 1024;
        if (smem < smem_min) smem = smem_min;
        if (smem > 200 * 1024) smem = 200 * 1024;
        int maxb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k_face<1>, 256, smem);
        const int blocks = nsm * maxb * 4;
        k_face<1><<<blocks, 256, smem>>>(out, 10, K, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_face<1><<<blocks, 256, smem>>>(out, reps, K, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double faces = (double)blocks * 256 * reps;
        const double rate = faces / (ms * 1e-3);
        const double pipe = rate * fp64_per_face / (nsm * 64.0 * 1.965e9);
        printf("  resident blocks/SM %d (asked %d): %.3f ms, %.2f Gface/s, FP64 pipe %.1f%% (at %.0f FP64 instr/face, 1965 MHz)\n",
               maxb, bps, ms, rate / 1e9, 100.0 * pipe, fp64_per_face);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    }
    return 0;
}
