// Micro-benchmark: cycles per tcgen05.mma M128 N176 K16 (bf16) with the A operand in shared memory under a 128-byte
// swizzle (SS) against A in tensor memory (TS), alone and with concurrent writes into shared memory: bulk weight copies
// (22.5 KB per 6 MMAs in the score kernels) and cp.async row gathers.  Question: does operand traffic through shared
// memory (reads by the tensor core + fills by TMA / LDGSTS) cap the MMA rate below its 88-cycle floor?
// The MMA warp runs converged and issues from an elected lane with precomputed descriptors (a divergent lane-0 loop
// is issue-bound at ~178 cycles per MMA and hides the effect).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int A_BYTES = 4 * 16384, B_OFF = A_BYTES, B_STAGE = 22528, B_BYTES = 3 * B_STAGE;
constexpr int W_OFF = B_OFF + B_BYTES, W_BYTES = 2 * B_STAGE;        // scratch target of the weight-copy traffic
constexpr int BG_OFF = W_OFF + W_BYTES, BG_BYTES = 32768;
constexpr int SMEM = BG_OFF + BG_BYTES + 1024;

// variant bit 0: A from smem (SS), bit 1: weight bulk-copy traffic, bit 2: cp.async gather traffic, bit 3: B reads skipped
// (A-only MMAs are impossible, so bit 3 instead uses N = 16 to show the A-side cost)
__global__ void __launch_bounds__(224) k(int variant, int iters, const uint8_t *src, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar, wbar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    __shared__ long long cnt[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < SMEM / 4 - 256; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&wbar, 1); mbar_fence_init(); stop = 0; }
    if (tid < 8) cnt[tid] = 0;
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const bool ss = variant & 1;
    if (warp == 0) {
        const uint32_t N = (variant & 8) ? 16 : 176;
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t sbase = smem_addr(smem);
        uint64_t ad[4], bd[3];
        for (int s = 0; s < 4; ++s) ad[s] = desc_sw128(sbase + s * 16384);
        for (int s = 0; s < 3; ++s) bd[s] = make_smem_desc(sbase + B_OFF + s * B_STAGE, 2816, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 12) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    // stage j / 3 of the A ring, K step (j % 2) inside the swizzle row, hi/lo halves; B stage j % 3
                    if (ss) mma_ss(tmem, ad[j >> 2] + ((j & 3) * 2), bd[j % 3] + ((j & 1) * (5632 >> 4)), idesc, 1);
                    else mma_ts(tmem, tmem + 400 + (j & 3) * 8, bd[j % 3] + ((j & 1) * (5632 >> 4)), idesc, 1);
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            out[blockIdx.x * 4] = clock64() - t0;
            stop = 1;
        }
    } else if (warp == 1) {
        if (lane == 0 && (variant & 2)) {     // free-running weight copies, one in flight... two in flight
            long long n = 0;
            uint32_t ph = 0;
            mbar_arrive_expect_tx(&wbar, B_STAGE);
            bulk_g2s(smem + W_OFF, src, B_STAGE, &wbar);
            while (!stop) {
                mbar_wait(&wbar, ph); ph ^= 1;
                ++n;
                mbar_arrive_expect_tx(&wbar, B_STAGE);
                bulk_g2s(smem + W_OFF + (n & 1) * B_STAGE, src + (n % 32) * B_STAGE, B_STAGE, &wbar);
            }
            mbar_wait(&wbar, ph);
            cnt[0] = n;
        }
    } else if (warp < 6 && (variant & 4)) {
        // 128 threads: row-gather pattern, 8 lanes cover one 128-byte row segment, 16 rows per instruction
        const int t = tid - 64;
        long long n = 0;
        uint32_t rng = 1234567u + blockIdx.x * 977u + t / 8;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                rng = rng * 1664525u + 1013904223u;
                const uint32_t row = (rng >> 8) % 6500u;
                const int r = j * 16 + t / 8, c = t % 8;
                cp16(smem_addr(smem + BG_OFF + (n & 1) * 16384 + r * 128 + ((c ^ (r & 7)) << 4)), src + (size_t)row * 2048 + ((n & 15) * 128) + c * 16);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            ++n;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (lane == 0) cnt[warp] = n;
    }
    tc_fence_before(); __syncthreads();
    if (tid == 0) { out[blockIdx.x * 4 + 1] = cnt[0]; out[blockIdx.x * 4 + 2] = cnt[2] + cnt[3] + cnt[4] + cnt[5]; }
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long *out; cudaMalloc(&out, 148 * 32);
    uint8_t *src; cudaMalloc(&src, 6500 * 2048); cudaMemset(src, 0x3c, 6500 * 2048);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const int iters = 24000;
    for (int rep = 0; rep < 2; ++rep)
    for (int variant : {0, 1, 2, 3, 4, 5, 6, 7, 8, 9}) {
        k<<<148, 224, SMEM>>>(variant, iters, src, out);
        long long h[148 * 4]; cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc = 0, w = 0, gth = 0;
        for (int b = 0; b < 148; ++b) { cyc += h[4 * b]; w += h[4 * b + 1]; gth += h[4 * b + 2]; }
        cyc /= 148; w /= 148; gth /= 148;
        if (rep == 0) continue;    // first pass warms the clocks up
        printf("%s%s%s%s: %.1f cycles/MMA; weight copies %.1f B/cycle/SM, gather %.1f B/cycle/SM  [%s]\n",
               (variant & 1) ? "SS(A smem SW128)" : "TS(A tmem)      ", (variant & 2) ? " +weightcopies" : "", (variant & 4) ? " +gather" : "",
               (variant & 8) ? " N=16" : "", cyc / iters, w * 22528.0 / cyc, gth * 4096.0 / cyc, cudaGetErrorString(e));
    }
    return 0;
}
