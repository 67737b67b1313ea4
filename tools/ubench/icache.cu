// Instruction-fetch microbenchmark (not part of the product): how much does it cost when the warps that share an SM
// sub-partition run DIFFERENT code, compared with all of them running the SAME loop body?
// body<COPY, N>: N straight-line FADD/FMUL instructions (8 independent chains), one copy per template argument, so
// different COPY values occupy different instruction addresses.  Each case runs 20 warps per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int COPY, int N>
__device__ __noinline__ void body(float (&f)[8], float c) {
#pragma unroll
    for (int i = 0; i < N / 8; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) f[j] = (i & 1) ? __fmul_rn(f[j], c) : __fadd_rn(f[j], c + COPY);
    }
}

// mode 0: every warp calls copy 0.  mode 1: warp group g (CTA-resident slot) calls copy g.  mode 2: like 0, with a
// CTA barrier per iteration (lock-step).
template <int N>
__global__ void k(float *out, int iters, int mode, int sm_count, int warps_per_group) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = 1.0f + j + threadIdx.x * 1e-3f;
    const float c = 1.0001f;
    int g = 0;
    if (mode == 1) g = gridDim.x > sm_count ? (blockIdx.x / sm_count) & 3 : ((threadIdx.x >> 5) / warps_per_group) & 3;
    for (int it = 0; it < iters; it++) {
        switch (g) {
            case 0: body<0, N>(f, c); break;
            case 1: body<1, N>(f, c); break;
            case 2: body<2, N>(f, c); break;
            default: body<3, N>(f, c); break;
        }
        if (mode == 2) __syncthreads();
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N>
static void run(const char *name, int sm_count) {
    float *out;
    cudaMalloc(&out, sizeof(float) * 148 * 4 * 640);
    const int iters = 4000000 / N;
    struct Cfg { const char *what; int ctas_per_sm, threads, mode; } cfgs[] = {
        {"1 CTA x 20 warps, same code            ", 1, 640, 0},
        {"1 CTA x 20 warps, same code, barrier/it", 1, 640, 2},
        {"1 CTA x 20 warps, 4 groups x own code  ", 1, 640, 1},
        {"4 CTA x  5 warps, same code            ", 4, 160, 0},
        {"4 CTA x  5 warps, own code per CTA slot", 4, 160, 1},
    };
    for (auto &c : cfgs) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        k<N><<<sm_count * c.ctas_per_sm, c.threads>>>(out, 10, c.mode, sm_count, 5);
        cudaEventRecord(a);
        k<N><<<sm_count * c.ctas_per_sm, c.threads>>>(out, iters, c.mode, sm_count, 5);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        int khz;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        const double warp_instr = (double)iters * N * 20;  // per SM
        printf("%-10s %s : %7.3f ms  ~%.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, c.what, ms, warp_instr / (ms * 1e-3 * khz * 1e3), khz / 1000);
    }
    cudaFree(out);
}

int main() {
    int sm_count;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, 0);
    run<256>("body 4KB", sm_count);
    run<896>("body 14KB", sm_count);
    run<2048>("body 32KB", sm_count);
    return 0;
}
