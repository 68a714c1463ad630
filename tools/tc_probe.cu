// Standalone hardware probe for the tcgen05 building blocks in neuralplda_b200/csrc/tc_ptx.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

constexpr int M = 128, N = 176, K = 64;
// canonical K-major no-swizzle images: element (r, k) at (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2
constexpr int A_SBO = 128, A_LBO = (M / 8) * 128 + 64;   // padded, as the real kernel does
constexpr int B_SBO = 128, B_LBO = (N / 8) * 128;
constexpr int A_BYTES = (K / 8) * A_LBO, B_BYTES = (K / 8) * B_LBO;

__global__ void __launch_bounds__(128) probe(const uint8_t *Aimg, const uint8_t *Bimg, const __nv_bfloat16 *Arow,
                                             float *D, uint32_t *dump, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *As = smem, *Bs = smem + ((A_BYTES + 1023) / 1024) * 1024;
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_load, A_BYTES + B_BYTES);
        bulk_g2s(As, Aimg, A_BYTES, &bar_load);
        bulk_g2s(Bs, Bimg, B_BYTES, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t a_tmem = tmem + 256;
    if (mode == 2) {   // A operand to TMEM: lane = row, column c holds k = 2c (low half), 2c+1 (high half)
        const int row = tid;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                uint32_t lo16 = __bfloat16_as_ushort(Arow[row * K + 2 * (c0 + j)]);
                uint32_t hi16 = __bfloat16_as_ushort(Arow[row * K + 2 * (c0 + j) + 1]);
                r[j] = lo16 | (hi16 << 16);
            }
            tmem_st8(a_tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (mode == 3) {   // discover the 16x256b store mapping
        uint32_t r0 = tid * 16 + 0, r1 = tid * 16 + 1, r2 = tid * 16 + 2, r3 = tid * 16 + 3;
        // zero the region first
        uint32_t z[8] = {0xdead, 0xdead, 0xdead, 0xdead, 0xdead, 0xdead, 0xdead, 0xdead};
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 300, z);
        tmem_st_wait();
        uint32_t r4 = tid * 16 + 4, r5 = tid * 16 + 5, r6 = tid * 16 + 6, r7 = tid * 16 + 7;
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 308, z);
        tmem_st_wait();
        asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                         tmem + ((uint32_t)(warp * 32 + 16) << 16) + 300),
                     "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7)
                     : "memory");
        tmem_st_wait();
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + 300, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) dump[tid * 16 + j] = v[j];
    } else {
        if (tid == 0) {
            for (int ks = 0; ks < K / 16; ++ks) {
                uint32_t lboA = A_LBO, sboA = A_SBO, lboB = B_LBO, sboB = B_SBO;
                if (mode == 1) { lboA = A_SBO; sboA = A_LBO; lboB = B_SBO; sboB = B_LBO; }
                uint64_t ad = make_smem_desc(smem_addr(As) + ks * 2 * A_LBO, lboA, sboA);
                uint64_t bd = make_smem_desc(smem_addr(Bs) + ks * 2 * B_LBO, lboB, sboB);
                if (mode == 2) mma_ts(tmem, a_tmem + ks * 8, bd, idesc, ks > 0);
                else mma_ss(tmem, ad, bd, idesc, ks > 0);
            }
            mma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, 0);
        tc_fence_after();
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    std::vector<__nv_bfloat16> A(M * K), B(N * K);
    std::vector<float> Af(M * K), Bf(N * K), ref(M * N);
    srand(1);
    for (int i = 0; i < M * K; ++i) { A[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.f); Af[i] = __bfloat162float(A[i]); }
    for (int i = 0; i < N * K; ++i) { B[i] = __float2bfloat16((rand() % 2001 - 1000) / 700.f); Bf[i] = __bfloat162float(B[i]); }
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)Af[m * K + k] * Bf[n * K + k]; ref[m * N + n] = (float)s; }
    std::vector<uint8_t> Ai(A_BYTES, 0), Bi(B_BYTES, 0);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) *(__nv_bfloat16 *)&Ai[(k / 8) * A_LBO + (r / 8) * A_SBO + (r % 8) * 16 + (k % 8) * 2] = A[r * K + k];
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) *(__nv_bfloat16 *)&Bi[(k / 8) * B_LBO + (r / 8) * B_SBO + (r % 8) * 16 + (k % 8) * 2] = B[r * K + k];
    uint8_t *dA, *dB; __nv_bfloat16 *dArow; float *dD; uint32_t *dump;
    cudaMalloc(&dA, A_BYTES); cudaMalloc(&dB, B_BYTES); cudaMalloc(&dArow, M * K * 2); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dump, 128 * 16 * 4);
    cudaMemcpy(dA, Ai.data(), A_BYTES, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bi.data(), B_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(dArow, A.data(), M * K * 2, cudaMemcpyHostToDevice);
    size_t smem = ((A_BYTES + 1023) / 1024) * 1024 + B_BYTES + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 4; ++mode) { if (mode != 3) continue;
        cudaMemset(dD, 0, M * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dArow, dD, dump, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        if (mode == 3) {
            std::vector<uint32_t> h(128 * 16);
            cudaMemcpy(h.data(), dump, h.size() * 4, cudaMemcpyDeviceToHost);
            printf("mode 3 (16x256b.x1 store, read back 32x32b): lane -> 8 columns as (src_thread,src_reg)\n");
            for (int l = 0; l < 32; ++l) {
                printf("  lane %2d:", l);
                for (int j = 0; j < 16; ++j) { uint32_t v = h[l * 16 + j]; if (v == 0xdead) printf("   --  "); else printf(" t%02d.r%d", v / 16, v % 16); }
                printf("\n");
            }
            continue;
        }
        std::vector<float> D(M * N);
        cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) { double er = fabs(D[i] - ref[i]); if (er > maxerr) maxerr = er; if (er > 1e-2) ++bad; }
        printf("mode %d: max abs err %.3e, mismatches %d / %d   (D[0]=%f ref[0]=%f, D[last]=%f ref[last]=%f)\n", mode, maxerr, bad, M * N, D[0], ref[0], D[M * N - 1], ref[M * N - 1]);
    }
    return 0;
}
