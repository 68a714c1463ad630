// Micro-benchmark of the MMA-issuing warp's per-stage instruction sequence (compile-time variants), one CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/issue_cost tools/issue_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;
constexpr int N = 176, LBO_B = (N / 8) * 128;
__device__ __forceinline__ void mma_f8_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
enum { FENCE = 1, ELECT = 2, SYNCW = 4, WAITS = 8, COMMIT2 = 16, COMMIT1 = 32, MMA6 = 64, MMA4 = 128, WAIT1 = 256, LANE0 = 512 };
template <int W>
__global__ void __launch_bounds__(128) k(int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bw[2], bc[2];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bw[0], 1); mbar_init(&bw[1], 1); mbar_init(&bc[0], 1); mbar_init(&bc[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot, a = tmem + 400;
    if (warp == 1) {
        const uint32_t id = make_idesc_bf16(128, N), id8 = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t bd = make_smem_desc(smem_addr(smem), LBO_B, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (W & FENCE) tc_fence_after();
            const bool e1 = (W & LANE0) ? (tid & 31) == 0 : ((W & ELECT) ? elect_one() : true);
            if (e1) {
                if (W & MMA6) { mma_ts(tmem, a, bd, id, 1); mma_ts(tmem, a + 8, bd, id, 1); mma_ts(tmem, a, bd + ((2 * LBO_B) >> 4), id, 1); }
                if (W & MMA4) { mma_ts(tmem, a, bd, id, 1); mma_ts(tmem, a + 8, bd + ((2 * LBO_B) >> 4), id, 1); }
            }
            if (W & SYNCW) __syncwarp();
            if (W & WAITS) { mbar_wait(&bw[0], 1); mbar_wait(&bw[1], 1); }
            if (W & WAIT1) { mbar_wait(&bw[0], 1); }
            const bool e2 = (W & LANE0) ? (tid & 31) == 0 : ((W & ELECT) ? elect_one() : true);
            if (e2) {
                if (W & MMA6) { mma_ts(tmem, a + 8, bd, id, 1); mma_ts(tmem, a, bd, id, 1); mma_ts(tmem, a + 8, bd + ((2 * LBO_B) >> 4), id, 1); }
                if (W & MMA4) { mma_f8_ts(tmem, a + 16, bd, id8, 1); mma_f8_ts(tmem, a + 24, bd, id8, 1); }
                if (W & COMMIT2) { mma_commit(&bc[0]); mma_commit(&bc[1]); }
                if (W & COMMIT1) { mma_commit(&bc[0]); }
            }
            if (W & SYNCW) __syncwarp();
        }
        if ((tid & 31) == 0) { mma_commit(&bar); mbar_wait(&bar, 0); out[0] = clock64() - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int W> void run(const char *name, long long *out) {
    const int iters = 2000;
    cudaFuncSetAttribute(k<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    k<W><<<1, 128, 40 * 1024>>>(iters, out);
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-58s %7.1f cycles/iteration [%s]\n", name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    long long *out; cudaMalloc(&out, 8);
    run<0>("empty loop", out);
    run<FENCE>("fence::after_thread_sync", out);
    run<ELECT>("2 x elect.sync", out);
    run<SYNCW>("2 x syncwarp", out);
    run<WAIT1>("1 satisfied mbarrier wait (all lanes)", out);
    run<WAITS>("2 satisfied mbarrier waits (all lanes)", out);
    run<ELECT | COMMIT1>("elect + 1 commit", out);
    run<ELECT | COMMIT2>("elect + 2 commits", out);
    run<LANE0 | COMMIT2>("lane0 + 2 commits", out);
    run<ELECT | MMA6>("elect + 6 MMAs", out);
    run<ELECT | MMA4>("elect + 2 f16 + 2 f8 MMAs", out);
    run<ELECT | MMA6 | COMMIT2>("elect + 6 MMAs + 2 commits", out);
    run<ELECT | MMA4 | COMMIT2>("elect + 4 MMAs + 2 commits", out);
    run<ELECT | MMA4 | COMMIT1>("elect + 4 MMAs + 1 commit", out);
    run<FENCE | ELECT | SYNCW | WAITS | COMMIT2>("kernel pattern without MMAs", out);
    run<FENCE | ELECT | SYNCW | WAITS | COMMIT2 | MMA6>("kernel pattern, 6 MMAs", out);
    run<FENCE | ELECT | SYNCW | WAITS | COMMIT2 | MMA4>("kernel pattern, 2 f16 + 2 f8 MMAs", out);
    run<FENCE | ELECT | SYNCW | WAITS | COMMIT1 | MMA4>("kernel pattern, 2+2 MMAs, 1 commit", out);
    run<ELECT | SYNCW | WAITS | COMMIT1 | MMA4>("no fence, 2+2 MMAs, 1 commit", out);
    return 0;
}
