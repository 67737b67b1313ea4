// Does sm_100a's packed FP32 (add/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2) free issue slots?  (not part of the product)
// Every thread runs independent chains; reported = cycles per loop iteration per SM sub-partition (4 schedulers per SM), i.e. how many
// issue cycles one iteration of the instruction mix costs a scheduler, next to the number of warp instructions in the mix.
// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false (the scalar forms are left alone), so an exactly
// rounded packed product is written fma.rn.f32x2(a, b, NZ) with NZ = {-0.0f, -0.0f} taken from a kernel argument the compiler cannot see through.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long u64;
#define ITERS 4096
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }

enum { S_FADD8, P_FADD2x4, P_FADD2x8, P_FMUL2x8, P_FFMA2x8, S_FADD8_LOP8, P_FADD2x4_LOP8, S_FADD8_DFMA4, P_FADD2x4_DFMA4, S_CMUL4, P_CMUL4, S_FADD8_LDS4, P_FADD2x4_LDS4, NT };
static const char *names[] = {"8 FADD", "4 FADD2", "8 FADD2", "8 FMUL2", "8 FFMA2", "8 FADD + 8 LOP3", "4 FADD2 + 8 LOP3", "8 FADD + 4 DFMA", "4 FADD2 + 4 DFMA",
                              "4 complex mul scalar (24)", "4 complex mul packed (8 FFMA2 + 4 FADD2)", "8 FADD + 4 LDS", "4 FADD2 + 4 LDS"};
static const int ninst[] = {8, 4, 8, 8, 8, 16, 12, 12, 8, 24, 12, 12, 8};

template <int OP>
__global__ void k(float *out, long long *cyc, float seed, u64 nz) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    float f[8];
    u64 p[8];
    double d[4];
    int n[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        f[j] = seed + j + threadIdx.x * 1e-3f;
        p[j] = pack(f[j], f[j] * 0.5f);
        n[j] = threadIdx.x + j;
        if (j < 4) d[j] = f[j];
    }
    const float c = seed * 0.999f;
    const u64 pc = pack(c, c * 1.01f);
    const double dc = c;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (OP == S_FADD8 || OP == S_FADD8_LOP8 || OP == S_FADD8_DFMA4 || OP == S_FADD8_LDS4) {
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = __fadd_rn(f[j], c);
        }
        if (OP == P_FADD2x4 || OP == P_FADD2x4_LOP8 || OP == P_FADD2x4_DFMA4 || OP == P_FADD2x4_LDS4) {
#pragma unroll
            for (int j = 0; j < 4; j++) p[j] = add2(p[j], pc);
        }
        if (OP == P_FADD2x8) {
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = add2(p[j], pc);
        }
        if (OP == P_FMUL2x8) {
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = mul2(p[j], pc);
        }
        if (OP == P_FFMA2x8) {
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = fma2(p[j], pc, nz);
        }
        if (OP == S_FADD8_LOP8 || OP == P_FADD2x4_LOP8) {
#pragma unroll
            for (int j = 0; j < 8; j++) n[j] = (n[j] ^ it) & 0x7fffffff;
        }
        if (OP == S_FADD8_DFMA4 || OP == P_FADD2x4_DFMA4) {
#pragma unroll
            for (int j = 0; j < 4; j++) d[j] = __fma_rn(d[j], dc, dc);
        }
        if (OP == S_FADD8_LDS4 || OP == P_FADD2x4_LDS4) {
#pragma unroll
            for (int j = 0; j < 4; j++) n[j] = __float_as_int(sm[(n[j] + threadIdx.x) & 1023]);
        }
        if (OP == S_CMUL4) {  // (f0 + i f1) *= (c + i c'), kissfft's C_MUL: 4 mul, 1 sub, 1 add
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float re = __fsub_rn(__fmul_rn(f[j], c), __fmul_rn(f[j + 1], seed));
                float im = __fadd_rn(__fmul_rn(f[j], seed), __fmul_rn(f[j + 1], c));
                f[j] = re;
                f[j + 1] = im;
            }
        }
        if (OP == P_CMUL4) {  // two butterflies per packed register: {re_a, re_b}, {im_a, im_b}
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                u64 re = add2(fma2(p[j], pc, nz), fma2(p[j + 1], p[j + 2], nz));
                u64 im = add2(fma2(p[j], p[j + 3], nz), fma2(p[j + 1], pc, nz));
                p[j] = re;
                p[j + 1] = im;
            }
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                u64 re = add2(fma2(p[j], pc, nz), fma2(p[j + 1], p[j + 2], nz));
                u64 im = add2(fma2(p[j], p[j + 3], nz), fma2(p[j + 1], pc, nz));
                p[j] = re;
                p[j + 1] = im;
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc += f[j] + lo(p[j]) + hi(p[j]) + (float)n[j] + (float)d[j & 3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// exactness: fma(a, b, -0) == mul(a, b) bit for bit, incl. denormal / zero-sign / inf / NaN-ness
__global__ void exact_check(const float *a, const float *b, int n, u64 nz, int *bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 r = fma2(pack(a[i], b[i]), pack(b[i], a[(i + 1) % n]), nz);
    float m0 = __fmul_rn(a[i], b[i]), m1 = __fmul_rn(b[i], a[(i + 1) % n]);
    bool ok0 = __float_as_uint(lo(r)) == __float_as_uint(m0) || (m0 != m0 && lo(r) != lo(r));
    bool ok1 = __float_as_uint(hi(r)) == __float_as_uint(m1) || (m1 != m1 && hi(r) != hi(r));
    u64 s = add2(pack(a[i], b[i]), pack(b[i], a[(i + 1) % n]));
    float s0 = __fadd_rn(a[i], b[i]), s1 = __fadd_rn(b[i], a[(i + 1) % n]);
    bool ok2 = __float_as_uint(lo(s)) == __float_as_uint(s0) || (s0 != s0 && lo(s) != lo(s));
    bool ok3 = __float_as_uint(hi(s)) == __float_as_uint(s1) || (s1 != s1 && hi(s) != hi(s));
    if (!(ok0 && ok1 && ok2 && ok3)) atomicAdd(bad, 1);
}

template <int OP>
void run(int sms, float *out, long long *cyc, int threads, int blocks_per_sm) {
    int blocks = sms * blocks_per_sm;
    const u64 nz = 0x8000000080000000ull;
    k<OP><<<blocks, threads>>>(out, cyc, 1.5f, nz);
    cudaDeviceSynchronize();
    k<OP><<<blocks, threads>>>(out, cyc, 1.5f, nz);
    cudaDeviceSynchronize();
    long long *h = (long long *)malloc(sizeof(long long) * blocks);
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += h[i];
    mean /= blocks;
    double warps_per_sched = threads * blocks_per_sm / 32.0 / 4.0;
    printf("%-44s threads/SM %4d : %2d warp inst / iteration, %6.2f issue cycles / iteration / warp  (%s)\n", names[OP], threads * blocks_per_sm, ninst[OP],
           mean / ITERS / warps_per_sched, cudaGetErrorString(cudaGetLastError()));
    free(h);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    float *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148 * 8);
    int sms = p.multiProcessorCount;
    for (int tpb = 320; tpb <= 512; tpb += 192) {
#define R(OP) run<OP>(sms, out, cyc, tpb, 2);
        R(S_FADD8) R(P_FADD2x4) R(P_FADD2x8) R(P_FMUL2x8) R(P_FFMA2x8) R(S_FADD8_LOP8) R(P_FADD2x4_LOP8) R(S_FADD8_DFMA4) R(P_FADD2x4_DFMA4) R(S_CMUL4) R(P_CMUL4)
        R(S_FADD8_LDS4) R(P_FADD2x4_LDS4)
    }
    // exactness on special values + random bit patterns
    const int n = 1 << 22;
    float *ha = (float *)malloc(n * 4), *hb = (float *)malloc(n * 4);
    uint32_t s = 12345;
    for (int i = 0; i < n; i++) {
        s = s * 1664525u + 1013904223u; uint32_t x = s;
        s = s * 1664525u + 1013904223u; uint32_t y = s;
        if (i % 4 == 1) { x &= 0x807fffffu; }                          // denormal a
        if (i % 4 == 2) { x = (x & 0x807fffffu) | 0x1f800000u; y = (y & 0x807fffffu) | 0x1f000000u; }  // product near the denormal border
        if (i % 64 == 3) x = 0x80000000u; if (i % 64 == 5) y = 0; if (i % 64 == 7) x = 0x7f800000u;
        memcpy(&ha[i], &x, 4); memcpy(&hb[i], &y, 4);
    }
    float *da, *db; int *dbad, hbad = 0;
    cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dbad, 4);
    cudaMemcpy(da, ha, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 4, cudaMemcpyHostToDevice); cudaMemset(dbad, 0, 4);
    exact_check<<<n / 256, 256>>>(da, db, n, 0x8000000080000000ull, dbad);
    cudaMemcpy(&hbad, dbad, 4, cudaMemcpyDeviceToHost);
    printf("fma2(a,b,-0) vs mul, add2 vs add on %d pairs (denormals, zeros, inf): %d mismatches (%s)\n", n, hbad, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
