// Hardware probe for mixing tcgen05.mma kinds on ONE fp32 accumulator in tensor memory:
//   D  = A16[128 x 32] * B16[N x 32]^T      kind::f16 (fp16 operands, two K = 16 instructions)
//   D += A8 [128 x 32] * B8 [N x 32]^T      kind::f8f6f4 (e4m3 or e5m2 operands, ONE K = 32 instruction)
// with A from tensor memory (32-bit cell c of a row holds K elements 2c, 2c+1 as fp16 / 4c..4c+3 as fp8) or
// from shared memory, B from shared memory in the K-major no-swizzle core-matrix layout (8 rows x 16 bytes).
// Also times both instruction kinds back to back.  N = 176.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/build/fp8_probe tools/fp8_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

constexpr int M = 128, N = 176, K = 32;
constexpr int LBO_B = (N / 8) * 128, LBO_A = (M / 8) * 128;
constexpr int B16_BYTES = (K / 8) * LBO_B;    // fp16: core matrix = 8 rows x 8 k
constexpr int B8_BYTES = (K / 16) * LBO_B;    // fp8 : core matrix = 8 rows x 16 k
constexpr int A16_BYTES = (K / 8) * LBO_A, A8_BYTES = (K / 16) * LBO_A;

__host__ __device__ constexpr uint32_t make_idesc(int fa, int fb, int m, int n) {
    return (1u << 4) | ((uint32_t)fa << 7) | ((uint32_t)fb << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f8_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f8_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

struct Imgs { const uint8_t *a16, *a8, *b16, *b8; const __half *a16row; const uint8_t *a8row; };

__global__ void __launch_bounds__(128) probe(Imgs im, float *D, long long *cyc, int ts, int fmt8, int which, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *B16 = smem, *B8 = B16 + B16_BYTES, *A16 = B8 + B8_BYTES, *A8 = A16 + A16_BYTES;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < B16_BYTES; i += 128) B16[i] = im.b16[i];
    for (int i = tid; i < B8_BYTES; i += 128) B8[i] = im.b8[i];
    for (int i = tid; i < A16_BYTES; i += 128) A16[i] = im.a16[i];
    for (int i = tid; i < A8_BYTES; i += 128) A8[i] = im.a8[i];
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot, a16t = tmem + 256, a8t = tmem + 288;
    {   // A operands in tensor memory: lane = row
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t r[8];
        for (int c0 = 0; c0 < 16; c0 += 8) {
            for (int j = 0; j < 8; ++j) {
                const uint32_t lo = __half_as_ushort(im.a16row[tid * K + 2 * (c0 + j)]), hi = __half_as_ushort(im.a16row[tid * K + 2 * (c0 + j) + 1]);
                r[j] = lo | (hi << 16);
            }
            tmem_st8(a16t + lane_base + c0, r);
        }
        for (int j = 0; j < 8; ++j) r[j] = *reinterpret_cast<const uint32_t *>(im.a8row + tid * K + 4 * j);
        tmem_st8(a8t + lane_base, r);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0 && which < 10) {
        const uint32_t id16 = make_idesc(0, 0, M, N);            // kind::f16: 0 = F16
        const uint32_t id8 = make_idesc(fmt8, fmt8, M, N);       // kind::f8f6f4: 0 = E4M3, 1 = E5M2
        const uint64_t b16d = make_smem_desc(smem_addr(B16), LBO_B, 128), b8d = make_smem_desc(smem_addr(B8), LBO_B, 128);
        const uint64_t a16d = make_smem_desc(smem_addr(A16), LBO_A, 128), a8d = make_smem_desc(smem_addr(A8), LBO_A, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t acc = it > 0;
            if (which & 1) {
                if (ts) { mma_ts(tmem, a16t, b16d, id16, acc); mma_ts(tmem, a16t + 8, b16d + ((2 * LBO_B) >> 4), id16, 1); }
                else { mma_ss(tmem, a16d, b16d, id16, acc); mma_ss(tmem, a16d + ((2 * LBO_A) >> 4), b16d + ((2 * LBO_B) >> 4), id16, 1); }
                acc = 1;
            }
            if (which & 2) {
                if (ts) mma_f8_ts(tmem, a8t, b8d, id8, acc); else mma_f8_ss(tmem, a8d, b8d, id8, acc);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        cyc[0] = clock64() - t0;
    }
    if (warp == 1 && which >= 10) {
        // "stage" pattern of the score kernel's MMA warp: converged warp, elected issue, two satisfied barrier
        // waits in the middle, two commits at the end.  which = 10: 6 bf16-style MMAs, 11: 2 f16 + 2 f8 MMAs,
        // 12: no MMAs, 13: 6 MMAs without waits/commits, 14: 2+2 without waits/commits
        __shared__ uint64_t bw[2], bc[2];
        if ((tid & 31) == 0) { mbar_init(&bw[0], 1); mbar_init(&bw[1], 1); mbar_init(&bc[0], 1); mbar_init(&bc[1], 1); mbar_fence_init(); }
        __syncwarp();
        const uint32_t id16 = make_idesc(0, 0, M, N), id8 = make_idesc(0, 0, M, N);
        const uint64_t b16d = make_smem_desc(smem_addr(B16), LBO_B, 128), b8d = make_smem_desc(smem_addr(B8), LBO_B, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            tc_fence_after();
            if (elect_one()) {
                if (which == 10 || which == 13) { mma_ts(tmem, a16t, b16d, id16, 1); mma_ts(tmem, a16t + 8, b16d, id16, 1); mma_ts(tmem, a16t, b16d + ((2 * LBO_B) >> 4), id16, 1); }
                if (which == 11 || which == 14) { mma_ts(tmem, a16t, b16d, id16, 1); mma_ts(tmem, a16t + 8, b16d + ((2 * LBO_B) >> 4), id16, 1); }
            }
            __syncwarp();
            if (which <= 12 || which == 15) { mbar_wait(&bw[0], 1); mbar_wait(&bw[1], 1); }
            if (elect_one()) {
                if (which == 10 || which == 13) { mma_ts(tmem, a16t + 8, b16d, id16, 1); mma_ts(tmem, a16t, b16d, id16, 1); mma_ts(tmem, a16t + 8, b16d + ((2 * LBO_B) >> 4), id16, 1); }
                if (which == 11 || which == 14) { mma_f8_ts(tmem, a8t, b8d, id8, 1); mma_f8_ts(tmem, a8t, b8d, id8, 1); }
                if (which <= 12 || which == 16) { mma_commit(&bc[0]); mma_commit(&bc[1]); }
                if (which == 17 || which == 18) mma_commit(&bc[0]);
                if (which == 19) { mma_ts(tmem, a16t, b16d, id16, 1); mma_commit(&bc[0]); mma_ts(tmem, a16t, b16d, id16, 1); mma_commit(&bc[1]); }
                if (which == 20) { mma_ts(tmem, a16t, b16d, id16, 1); mma_ts(tmem, a16t, b16d, id16, 1); mma_commit(&bc[0]); }
            }
            __syncwarp();
            if (which == 18) { mbar_wait(&bc[0], it & 1); }      // wait for the commit's arrival every iteration
        }
        if (elect_one()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if ((tid & 31) == 0) cyc[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float f8_to_float(uint8_t v, int fmt) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(v, fmt ? __NV_E5M2 : __NV_E4M3);
    return __half2float(*reinterpret_cast<__half *>(&h));
}

int main() {
    int bad_total = 0;
    for (int fmt8 = 0; fmt8 < 2; ++fmt8) {
        std::vector<__half> A16(M * K), B16(N * K);
        std::vector<uint8_t> A8(M * K), B8(N * K);
        std::vector<float> A16f(M * K), B16f(N * K), A8f(M * K), B8f(N * K);
        srand(11 + fmt8);
        auto rnd = [] { return (rand() % 2001 - 1000) / 500.f; };
        for (int i = 0; i < M * K; ++i) {
            A16[i] = __float2half(rnd()); A16f[i] = __half2float(A16[i]);
            A8[i] = __nv_cvt_float_to_fp8(rnd(), __NV_SATFINITE, fmt8 ? __NV_E5M2 : __NV_E4M3); A8f[i] = f8_to_float(A8[i], fmt8);
        }
        for (int i = 0; i < N * K; ++i) {
            B16[i] = __float2half(rnd()); B16f[i] = __half2float(B16[i]);
            B8[i] = __nv_cvt_float_to_fp8(rnd(), __NV_SATFINITE, fmt8 ? __NV_E5M2 : __NV_E4M3); B8f[i] = f8_to_float(B8[i], fmt8);
        }
        std::vector<uint8_t> b16i(B16_BYTES), b8i(B8_BYTES), a16i(A16_BYTES), a8i(A8_BYTES);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
            *(__half *)&b16i[(k / 8) * LBO_B + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2] = B16[n * K + k];
            b8i[(k / 16) * LBO_B + (n / 8) * 128 + (n % 8) * 16 + (k % 16)] = B8[n * K + k];
        }
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            *(__half *)&a16i[(k / 8) * LBO_A + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2] = A16[m * K + k];
            a8i[(k / 16) * LBO_A + (m / 8) * 128 + (m % 8) * 16 + (k % 16)] = A8[m * K + k];
        }
        Imgs im; uint8_t *p; float *dD; long long *dc;
        auto up = [&](const void *h, size_t b) { cudaMalloc(&p, b); cudaMemcpy(p, h, b, cudaMemcpyHostToDevice); return p; };
        im.a16 = up(a16i.data(), a16i.size()); im.a8 = up(a8i.data(), a8i.size());
        im.b16 = up(b16i.data(), b16i.size()); im.b8 = up(b8i.data(), b8i.size());
        im.a16row = (const __half *)up(A16.data(), A16.size() * 2); im.a8row = up(A8.data(), A8.size());
        cudaMalloc(&dD, M * N * 4); cudaMalloc(&dc, 8);
        const int smem = B16_BYTES + B8_BYTES + A16_BYTES + A8_BYTES + 1024;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int ts = 0; ts < 2; ++ts)
            for (int which = 1; which <= 3; ++which) {
                cudaMemset(dD, 0xff, M * N * 4);
                probe<<<1, 128, smem>>>(im, dD, dc, ts, fmt8, which, 1);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> D(M * N);
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0; double worst = 0;
                for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) {
                        if (which & 1) ref += (double)A16f[m * K + k] * B16f[n * K + k];
                        if (which & 2) ref += (double)A8f[m * K + k] * B8f[n * K + k];
                    }
                    const double d = fabs(ref - D[m * N + n]);
                    if (!(d <= 2e-3)) { if (bad < 3) printf("   mismatch m %d n %d got %g ref %g\n", m, n, D[m * N + n], ref); ++bad; }
                    if (d > worst) worst = d;
                }
                printf("%s A=%s which=%d (1 f16, 2 f8, 3 both on one accumulator): %s, %d mismatches, worst %.3g\n", fmt8 ? "e5m2" : "e4m3",
                       ts ? "tmem" : "smem", which, cudaGetErrorString(e), bad, worst);
                bad_total += bad;
                if (e != cudaSuccess) return 2;
            }
        if (fmt8 == 0)
            for (int which = 10; which <= 20; ++which) {
                const int iters = 2000;
                probe<<<1, 128, smem>>>(im, dD, dc, 1, 0, which, iters);
                long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
                printf("stage pattern which=%d (10: 6 MMAs+2 waits+2 commits, 11: 2 f16+2 f8 +waits+commits, 12: waits+commits only, 13: 6 MMAs only, 14: 2+2 only, 15: 2 waits only, 16: 2 commits only, 17: 1 commit, 18: 1 commit + wait for it, 19: mma,commit,mma,commit, 20: 2 mma + 1 commit): %.0f cycles per stage [%s]\n", which, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
            }
        if (fmt8 == 0)
            for (int ts = 0; ts < 2; ++ts)
                for (int which = 1; which <= 3; ++which) {
                    const int iters = 2000;
                    probe<<<1, 128, smem>>>(im, dD, dc, ts, 0, which, iters);
                    long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
                    const int per = (which & 1 ? 2 : 0) + (which & 2 ? 1 : 0);
                    printf("timing A=%s which=%d: %.1f cycles per MMA instruction (%d per iteration)\n", ts ? "tmem" : "smem", which, (double)h / iters / per, per);
                }
    }
    printf(bad_total ? "FP8 PROBE FAILED\n" : "FP8 PROBE OK\n");
    return bad_total != 0;
}
