// Micro-benchmark of the synchronisation primitives the tensor-core score kernel is built from (one CTA per SM,
// results of CTA 0): cost of a satisfied mbarrier wait, warp-to-warp hand-off latency through mbarriers
// (try_wait vs test_wait polling), tcgen05.commit -> mbarrier latency, tcgen05.st / tcgen05.ld round trips.
#include <cstdio>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_poll(uint64_t *bar, uint32_t parity) { while (!mbar_test(bar, parity)) {} }

__global__ void __launch_bounds__(128) k(int iters, long long *out) {
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    long long t0, t1;
    // 1. satisfied try_wait: bars[7] phase 0 never completes, so waiting for parity 1 returns at once
    if (tid == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) mbar_wait(&bars[7], 1);
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) mbar_poll(&bars[7], 1);
        t1 = clock64();
        if (blockIdx.x == 0) out[1] = (t1 - t0);
    }
    __syncthreads();
    // 2. ping-pong warp 0 <-> warp 1 (lane 0 each), try_wait
    if (lane == 0 && warp < 2) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ph = i & 1;
            if (warp == 0) { mbar_arrive(&bars[0]); mbar_wait(&bars[1], ph); }
            else { mbar_wait(&bars[0], ph); mbar_arrive(&bars[1]); }
        }
        t1 = clock64();
        if (blockIdx.x == 0 && warp == 0) out[2] = (t1 - t0);
    }
    __syncthreads();
    // 3. same with test_wait polling (barriers 2, 3)
    if (lane == 0 && warp < 2) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ph = i & 1;
            if (warp == 0) { mbar_arrive(&bars[2]); mbar_poll(&bars[3], ph); }
            else { mbar_poll(&bars[2], ph); mbar_arrive(&bars[3]); }
        }
        t1 = clock64();
        if (blockIdx.x == 0 && warp == 0) out[3] = (t1 - t0);
    }
    __syncthreads();
    // 4. tcgen05.commit with nothing in flight -> own wait
    if (tid == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { mma_commit(&bars[4]); mbar_wait(&bars[4], i & 1); }
        t1 = clock64();
        if (blockIdx.x == 0) out[4] = (t1 - t0);
    }
    __syncthreads();
    // 5. commit issue cost alone (nobody waits; barrier 5 just flips)
    if (tid == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) mma_commit(&bars[5]);
        t1 = clock64();
        if (blockIdx.x == 0) out[5] = (t1 - t0);
        mbar_wait(&bars[5], (iters - 1) & 1);
    }
    __syncthreads();
    // 6. tcgen05.st x2 (16x256b.x2 twice) + wait::st, whole warp
    {
        uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, (uint32_t)tid};
        const uint32_t a = tmem + ((uint32_t)(warp * 32) << 16) + 352;
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a + 16), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            tmem_st_wait();
        }
        t1 = clock64();
        if (blockIdx.x == 0 && tid == 0) out[6] = (t1 - t0);
        // 7. tcgen05.ld 16x256b.x2 twice + wait::ld
        uint32_t v[8], w[8], accv = 0;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(a + 16) : "memory");
            tmem_ld_wait();
            accv += v[0] + w[7];
        }
        t1 = clock64();
        if (blockIdx.x == 0 && tid == 0) { out[7] = (t1 - t0); out[15] = accv; }
    }
    __syncthreads();
    // 8. ping-pong with 8 arrivals per side (8 lanes of each warp arrive; models 8 converter warps): skip
    // 9. commit -> OTHER warp waits -> arrives back (A-ring hand-off shape): warp 0 commits bars[0'] ...
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
    if (lane == 0 && warp < 2) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t ph = i & 1;
            if (warp == 0) { mma_commit(&bars[0]); mbar_wait(&bars[1], ph); }
            else { mbar_wait(&bars[0], ph); mbar_arrive(&bars[1]); }
        }
        t1 = clock64();
        if (blockIdx.x == 0 && warp == 0) out[8] = (t1 - t0);
    }
    __syncthreads();
    // 10. satisfied try_wait executed by ALL 32 lanes of a warp
    if (warp == 0) {
        __syncwarp();
        t0 = clock64();
        for (int i = 0; i < iters; ++i) mbar_wait(&bars[7], 1);
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[9] = (t1 - t0);
        // 11. lane-0 wait + __syncwarp
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { if (lane == 0) mbar_wait(&bars[7], 1); __syncwarp(); }
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[10] = (t1 - t0);
        // 12. tcgen05.fence::after_thread_sync
        t0 = clock64();
        for (int i = 0; i < iters; ++i) tc_fence_after();
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) out[11] = (t1 - t0);
        // 13. elect_one + __syncwarp
        uint32_t e = 0;
        t0 = clock64();
        for (int i = 0; i < iters; ++i) { e += elect_one(); __syncwarp(); }
        t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[12] = (t1 - t0); out[15] += e; }
        // 14. two satisfied waits overlapped: both try_waits issued before either branch (single lane)
        if (lane == 0) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                uint32_t ok;
                do {
                    asm volatile("{\n.reg .pred p, q;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %3;\n"
                                 "mbarrier.try_wait.parity.shared::cta.b64 q, [%2], %3;\nand.pred p, p, q;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(ok) : "r"(smem_addr(&bars[7])), "r"(smem_addr(&bars[6])), "r"(1u) : "memory");
                } while (!ok);
            }
            t1 = clock64();
            if (blockIdx.x == 0) out[13] = (t1 - t0);
            // 15. clock64 read cost
            long long acc = 0;
            t0 = clock64();
            for (int i = 0; i < iters; ++i) acc += clock64();
            t1 = clock64();
            if (blockIdx.x == 0) { out[14] = (t1 - t0); out[15] += acc; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long *out; cudaMalloc(&out, 16 * 8); cudaMemset(out, 0, 128);
    const int iters = 2000;
    k<<<148, 128>>>(iters, out);
    long long h[16]; cudaMemcpy(h, out, 128, cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
    const char *names[15] = {"satisfied try_wait", "satisfied test_wait poll", "ping-pong try_wait (round trip)", "ping-pong test_wait poll (round trip)",
                            "commit (idle pipe) + own wait", "commit issue only", "2x tcgen05.st.16x256b.x2 + wait::st", "2x tcgen05.ld.16x256b.x2 + wait::ld",
                            "commit -> other warp try_wait -> arrive -> try_wait (round trip)",
                            "satisfied try_wait, all 32 lanes", "lane-0 satisfied try_wait + __syncwarp", "tcgen05.fence::after_thread_sync (warp)",
                            "elect_one + __syncwarp", "two satisfied try_waits overlapped (1 lane)", "clock64 read"};
    for (int i = 0; i < 15; ++i) printf("%-70s %8.1f cycles\n", names[i], (double)h[i] / iters);
    return 0;
}
