// Instruction-throughput microbenchmark for design decisions in kernels.cu (not part of the product).
// For each op: every thread runs ILP independent dependency chains for ITERS iterations; blocks record
// clock64() deltas; reported = thread-level ops per clock per SM (128 = full-rate FP32).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

enum Op { FADD, FMUL, FFMA, DADD, DMUL, DFMA, F2D, D2F, F2D_D2F_CHAIN, DSQRT, FSQRT, FDIV, I2F, PRMT, IADD, LOP, IMADW, DP4A, LDS32, LDS64, SHFL, MUFU_RSQ, CMVN_F64, CMVN_F32X };
static const char *names[] = {"FADD", "FMUL", "FFMA", "DADD", "DMUL", "DFMA", "F2F.F64.F32", "F2F.F32.F64", "f->d->f chain", "dsqrt_rn", "fsqrt_rn", "fdiv_rn",
                              "I2F", "PRMT", "IADD3", "LOP3", "IMAD.WIDE", "IDP.4A", "LDS.32", "LDS.64", "SHFL", "MUFU.RSQ", "cmvn term fp64", "cmvn term fp32x"};

template <int OP>
__global__ void k(float *out, long long *cyc, float seed) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    float f[ILP];
    double d[ILP];
    int n[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) {
        f[j] = seed + j + threadIdx.x * 1e-3f;
        d[j] = (double)f[j];
        n[j] = threadIdx.x + j;
    }
    const float c = seed * 0.999f;
    const double dc = (double)c;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (OP == FADD) f[j] = __fadd_rn(f[j], c);
            if (OP == FMUL) f[j] = __fmul_rn(f[j], c);
            if (OP == FFMA) f[j] = __fmaf_rn(f[j], c, c);
            if (OP == DADD) d[j] = __dadd_rn(d[j], dc);
            if (OP == DMUL) d[j] = __dmul_rn(d[j], dc);
            if (OP == DFMA) d[j] = __fma_rn(d[j], dc, dc);
            if (OP == F2D) { d[j] = (double)f[j]; f[j] = __int_as_float(__double2hiint(d[j]) ^ it); }
            if (OP == D2F) { f[j] = (float)d[j]; d[j] = __hiloint2double(__float_as_int(f[j]), it); }
            if (OP == F2D_D2F_CHAIN) f[j] = (float)__dadd_rn((double)f[j], dc);
            if (OP == DSQRT) d[j] = __dsqrt_rn(d[j]) + 1.0;
            if (OP == FSQRT) f[j] = __fsqrt_rn(f[j]) + 1.0f;
            if (OP == FDIV) f[j] = __fdiv_rn(f[j], c) + 1.0f;
            if (OP == I2F) { f[j] = (float)n[j]; n[j] = __float_as_int(f[j]) ^ it; }
            if (OP == PRMT) n[j] = __byte_perm(n[j], it, 0x7610 ^ j);
            if (OP == IADD) n[j] = n[j] + it;
            if (OP == LOP) n[j] = (n[j] ^ it) & 0x7fffffff;
            if (OP == IMADW) { long long w = (long long)n[j] * (long long)(it | 1); n[j] = (int)(w >> 31); }
            if (OP == DP4A) n[j] = __dp4a(n[j], it, n[j]);
            if (OP == LDS32) n[j] = __float_as_int(sm[(n[j] + threadIdx.x) & 1023]);
            if (OP == LDS64) { float2 v = *(float2 *)&sm[((n[j] + threadIdx.x) * 2) & 1022]; n[j] = __float_as_int(v.x) + __float_as_int(v.y); }
            if (OP == SHFL) n[j] = __shfl_xor_sync(0xffffffffu, n[j], 1);
            if (OP == MUFU_RSQ) f[j] = rsqrtf(f[j]) + 1.0f;
            if (OP == CMVN_F64) {  // S = (float)((double)S + d*d), d = x - mean
                double dd = (double)__fsub_rn(c, f[j] * 0.f + (float)it);
                f[j] = (float)__dadd_rn((double)f[j], __dmul_rn(dd, dd));
            }
            if (OP == CMVN_F32X) {  // float-float emulation: qh+ql = d*d exactly, two-sum with S, add tails
                float dd = __fsub_rn(c, (float)it);
                float qh = __fmul_rn(dd, dd), ql = __fmaf_rn(dd, dd, -qh);
                float s1 = __fadd_rn(f[j], qh), bb = __fsub_rn(s1, f[j]);
                float e1 = __fadd_rn(__fsub_rn(f[j], __fsub_rn(s1, bb)), __fsub_rn(qh, bb));
                f[j] = __fadd_rn(s1, __fadd_rn(e1, ql));
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) acc += f[j] + (float)d[j] + (float)n[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, float *out, long long *cyc, int threads, int blocks_per_sm) {
    int blocks = sms * blocks_per_sm;
    k<OP><<<blocks, threads>>>(out, cyc, 1.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(out, cyc, 1.5f);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    long long *h = (long long *)malloc(sizeof(long long) * blocks);
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += h[i];
    mean /= blocks;
    double ops_per_sm = (double)threads * blocks_per_sm * ITERS * ILP;
    printf("%-16s threads/SM %4d : %8.2f ops/clk/SM   (%.3f ms, %.0f cycles)\n", names[OP], threads * blocks_per_sm, ops_per_sm / mean, ms, mean);
    free(h);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148 * 8);
    int sms = p.multiProcessorCount;
#define R(OP) run<OP>(sms, out, cyc, 256, 4);
    R(FADD) R(FMUL) R(FFMA) R(DADD) R(DMUL) R(DFMA) R(F2D) R(D2F) R(F2D_D2F_CHAIN) R(DSQRT) R(FSQRT) R(FDIV) R(I2F) R(PRMT) R(IADD) R(LOP) R(IMADW) R(DP4A)
    R(LDS32) R(LDS64) R(SHFL) R(MUFU_RSQ) R(CMVN_F64) R(CMVN_F32X)
    return 0;
}
