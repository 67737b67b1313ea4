// Does packed FP32 (FADD2 / FFMA2) free ISSUE slots for the other pipes on sm_100a?  (not part of the product)
// One loop iteration imitates the instruction mix of the FFT loop in kernels.cu per 32 FP32 operations: 8 DFMA (fp64 pipe), 4 LOP3 (ALU),
// 4 conflict-free LDS.  Variant S issues the FP32 work as 32 scalar FADD/FMUL (48 warp instructions per iteration), variant P as 16 packed
// add.rn.f32x2 / fma.rn.f32x2 (32 warp instructions).  Timed with CUDA events over the whole grid (2 CTAs x 320 threads per SM like the
// product kernels); reported: issue cycles per iteration per scheduler = elapsed cycles / (iterations x warps per scheduler).
#include <cuda_runtime.h>
#include <stdio.h>
typedef unsigned long long u64;
#define ITERS 8192
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pack(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

// MODE 0: scalar FP32 only (32)   1: packed only (16)   2: scalar + others   3: packed + others   4: others only
template <int MODE>
__global__ void __launch_bounds__(320, 2) k(float *out, float seed, u64 nz) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = seed + i;
    __syncthreads();
    float f[16];
    u64 p[8];
    double d[4];
    int n[4];
#pragma unroll
    for (int j = 0; j < 16; j++) f[j] = seed + j + threadIdx.x * 1e-3f;
#pragma unroll
    for (int j = 0; j < 8; j++) p[j] = pack(f[2 * j], f[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 4; j++) { d[j] = f[j]; n[j] = threadIdx.x + j; }
    const float c = seed * 0.999f, c2 = seed * 1.0001f;
    const u64 pc = pack(c, c2);
    const double dc = c;
    const float *ls = sm + threadIdx.x;  // lane-consecutive: conflict-free
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = __fadd_rn(f[j], c);
#pragma unroll
            for (int j = 0; j < 16; j++) f[j] = __fmul_rn(f[j], c2);
        }
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = add2(p[j], pc);
#pragma unroll
            for (int j = 0; j < 8; j++) p[j] = fma2(p[j], pc, nz);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                d[j] = __fma_rn(d[j], dc, dc);
                d[j] = __fma_rn(d[j], dc, dc);
                n[j] = (n[j] ^ it) & 0x7ffffff;
                n[j] += __float_as_int(ls[(j * 320 + (it & 3) * 32) & 1023]);
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) acc += f[j];
#pragma unroll
    for (int j = 0; j < 8; j++) acc += lo(p[j]);
#pragma unroll
    for (int j = 0; j < 4; j++) acc += (float)d[j] + (float)n[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int sms, float *out, int mhz) {
    const u64 nz = 0x8000000080000000ull;
    k<MODE><<<sms * 2, 320>>>(out, 1.5f, nz);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<sms * 2, 320>>>(out, 1.5f, nz);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double cycles = ms * 1e-3 * mhz * 1e6, warps_per_sched = 2 * 320 / 32.0 / 4.0;
    printf("%-58s %7.3f ms  %6.2f cycles / iteration / scheduler-warp  (%s)\n", name, ms, cycles / ITERS / warps_per_sched, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int mhz = pr.clockRate / 1000;
    printf("%s, %d SMs, %d MHz (nominal; cycles assume it)\n", pr.name, pr.multiProcessorCount, mhz);
    float *out;
    cudaMalloc(&out, sizeof(float) * 148 * 2 * 320);
    int sms = pr.multiProcessorCount;
    for (int rep = 0; rep < 2; rep++) {
        run<0>("32 scalar FADD/FMUL                       (32 inst)", sms, out, mhz);
        run<1>("16 packed FADD2/FFMA2                     (16 inst)", sms, out, mhz);
        run<4>("8 DFMA + 4 LOP3 + 4 IADD + 4 LDS           (~20 inst)", sms, out, mhz);
        run<2>("32 scalar + 8 DFMA + 4 LOP3 + 4 IADD + 4 LDS (~52 inst)", sms, out, mhz);
        run<3>("16 packed + 8 DFMA + 4 LOP3 + 4 IADD + 4 LDS (~36 inst)", sms, out, mhz);
    }
    return 0;
}
