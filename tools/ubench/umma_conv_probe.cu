// Probe for the tensor-core lowering of the int8 1xKW convolution (block 1 of the fused classifier):
//   D[channel][position] = sum_k W[channel][k] * X[position*16 + k]     (k < KW*16 bytes, 16-byte channel-padded rows)
// as ONE tcgen05.mma.kind::i8 per 32 bytes of K, with the im2col matrix never materialised: the B operand is a K-major,
// no-swizzle shared-memory descriptor whose 8-row core matrices overlap (row stride 16 B, K-chunk stride 16 B), i.e. the
// sliding windows of the padded feature matrix are addressed in place.
// The probe runs a matrix of descriptor conventions, dumps all of TMEM's 128 lanes x 128 columns and reports which
// convention and which row -> lane mapping reproduce the host result.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static constexpr int kRows = 160;  // padded feature rows available in shared memory (16 B each)
static constexpr int kK = 128;     // bytes of K per output (8 rows of 16 B; the last row multiplies zero weights)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo16, uint32_t sbo16) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)(lbo16 & 0x3fff) << 16;
    d |= (uint64_t)(sbo16 & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// A (weights) in shared memory: [k_chunk 0..7][row 0..M-1][16 B]  => 8-row group stride 128 B, K-chunk stride M*16 B
__global__ void __launch_bounds__(128, 1)
probe(const int8_t *w, const int8_t *x, int32_t *out, int M, int N, int lboA, int sboA, int lboB, int sboB) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sA = smem;                       // 128 rows x 128 B = 16 KB
    uint8_t *sX = smem + 16384;               // kRows x 16 B
    uint64_t *bar = (uint64_t *)(smem + 16384 + kRows * 16);
    uint32_t *tmem_slot = (uint32_t *)(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * kK; i += 128) {
        const int row = i / kK, k = i % kK;
        sA[(k / 16) * (M * 16) + row * 16 + (k % 16)] = row < M ? (uint8_t)w[row * kK + k] : 0;
    }
    for (int i = tid; i < kRows * 16; i += 128) sX[i] = (uint8_t)x[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < kK / 32; kb++) {
            const uint64_t da = make_desc(smem_u32(sA) + kb * 2 * (M * 16), lboA, sboA);
            const uint64_t db = make_desc(smem_u32(sX) + kb * 32, lboB, sboB);
            const uint32_t acc = kb > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // everyone waits for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(
            smem_u32(bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; j++) out[(warp * 32 + lane) * 128 + c0 + j] = (int32_t)v[j];
    }
    // unaligned column starts (the epilogue reads 7-column pool groups): x16 from column 21 and from column 91 -> out rows 128..255
    for (int t = 0; t < 2; t++) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (t ? 91 : 21);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) out[(128 + warp * 32 + lane) * 128 + t * 16 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    int8_t *hw = (int8_t *)malloc(128 * kK), *hx = (int8_t *)malloc(kRows * 16);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)(s >> 24) - 128; };
    for (int r = 0; r < 128; r++)
        for (int k = 0; k < kK; k++) hw[r * kK + k] = (k < 112 && (k % 16) < 13) ? (int8_t)rnd() : 0;  // 7 taps x 13 channels
    for (int i = 0; i < kRows * 16; i++) hx[i] = (int8_t)rnd();
    int8_t *dw, *dx;
    int32_t *dout, *hout = (int32_t *)malloc(256 * 128 * 4);
    cudaMalloc(&dw, 128 * kK);
    cudaMalloc(&dx, kRows * 16);
    cudaMalloc(&dout, 256 * 128 * 4);
    cudaMemcpy(dw, hw, 128 * kK, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, hx, kRows * 16, cudaMemcpyHostToDevice);
    const int smem_bytes = 16384 + kRows * 16 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    const int Ms[2] = {128, 64};
    const int N = 112;
    for (int mi = 0; mi < 2; mi++) {
        const int M = Ms[mi];
        for (int variant = 0; variant < 1; variant++) {  // variants 1..3 (LBO/SBO swapped) fault: the CUTLASS reading is the right one
            // convention 0: LBO = K-chunk stride, SBO = 8-row-group stride (CUTLASS's reading); convention 1: swapped
            const int kA = M, rA = 8, kB = 1, rB = 8;  // in 16-byte units
            const int lboA = (variant & 1) ? rA : kA, sboA = (variant & 1) ? kA : rA;
            const int lboB = (variant & 2) ? rB : kB, sboB = (variant & 2) ? kB : rB;
            cudaMemset(dout, 0xff, 256 * 128 * 4);
            probe<<<1, 128, smem_bytes>>>(dw, dx, dout, M, N, lboA, sboA, lboB, sboB);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("M=%d variant %d: CUDA error %s\n", M, variant, cudaGetErrorString(e));
                return 1;
            }
            cudaMemcpy(hout, dout, 256 * 128 * 4, cudaMemcpyDeviceToHost);
            // expected D[row][n] = sum_k w[row][k] * x[16 n + k]
            int ok_direct = 0, total = 0;
            int lane_of_row[128];
            for (int r = 0; r < M; r++) {
                int32_t want[128];
                for (int n = 0; n < N; n++) {
                    int32_t acc = 0;
                    for (int k = 0; k < kK; k++) acc += (int32_t)hw[r * kK + k] * (int32_t)hx[16 * n + k];
                    want[n] = acc;
                    total++;
                    if (hout[r * 128 + n] == acc) ok_direct++;
                }
                lane_of_row[r] = -1;
                for (int l = 0; l < 128; l++)
                    if (memcmp(&hout[l * 128], want, N * 4) == 0) { lane_of_row[r] = l; break; }
            }
            printf("M=%3d N=%d A(lbo=%d,sbo=%d) B(lbo=%d,sbo=%d): row->lane direct %d/%d\n  lane of row:", M, N, lboA, sboA, lboB, sboB, ok_direct, total);
            for (int r = 0; r < M; r++) printf(" %d", lane_of_row[r]);
            int un_ok = 0;
            for (int l = 0; l < 128; l++)
                for (int j = 0; j < 16; j++)
                    un_ok += (hout[(128 + l) * 128 + j] == hout[l * 128 + 21 + j]) + (hout[(128 + l) * 128 + 16 + j] == hout[l * 128 + 91 + j]);
            printf("\n  unaligned x16 loads (columns 21.., 91..) equal to the aligned dump: %d/%d\n", un_ok, 128 * 32);
        }
    }
    return 0;
}
