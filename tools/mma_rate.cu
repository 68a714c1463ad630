// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) for N in {176, 256}, A from smem (SS) or TMEM (TS),
// accumulating into ONE accumulator back to back or alternating between TWO accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

__global__ void __launch_bounds__(128) k(int N, int ts, int nacc, int iters, long long *out, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t sbase = smem_addr(smem);
        const uint64_t ad = make_smem_desc(sbase, 2048, 128);
        const uint64_t bd = make_smem_desc(sbase + 8192, (N / 8) * 128, 128);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (nacc == 2 ? (it & 1) * 256 : 0);
            if (ts) mma_ts(d, tmem + 480, bd, idesc, 1); else mma_ss(d, ad, bd, idesc, 1);
            if (it % 6 == 5) {
                if (mode & 1) mma_commit(&bar2);                 // a commit per 6 MMAs (nobody waits)
                if (mode & 2) { mbar_wait(&bar3, 1); mbar_wait(&bar3, 1); }   // two waits that are already satisfied
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long *out; cudaMalloc(&out, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int grid : {1, 16, 74, 148})
    for (int N : {176})
        for (int ts = 0; ts < 2; ++ts)
            for (int mode = 0; mode < 4; mode += 3) {
                const int iters = 1998;
                k<<<grid, 128, 64 * 1024>>>(N, ts, 1, iters, out, mode);
                long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                printf("grid %3d N=%3d %s mode=%d (1: commit/6 MMAs, 2: two satisfied waits/6 MMAs): %.1f cycles/MMA, %.0f cycles per group of 6  [%s]\n", grid, N, ts ? "TS" : "SS", mode, (double)h / iters, 6.0 * h / iters, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
