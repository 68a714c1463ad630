// Hardware probe for the CTA-pair (cta_group::2) building blocks of the score kernel:
//   * tcgen05.alloc / dealloc .cta_group::2 by one warp of each CTA of a 2-CTA cluster
//   * B operand halves loaded by EACH CTA with cp.async.bulk.tensor.2d.cta_group::2, completing on the LEADER's mbarrier
//   * tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA) with A from tensor memory (TS) or shared memory (SS)
//   * tcgen05.commit.cta_group::2 ... multicast::cluster arriving on the same barrier offset in both CTAs
//   * remote mbarrier arrive (mapa + mbarrier.arrive.release.cluster) from the peer to the leader
// D = A[256 x K] * B[N x K]^T is compared with the host result for N in {176, 192}, TS and SS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/build/pair_probe tools/pair_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
#include "../neuralplda_b200/csrc/tc_pair_ptx.cuh"
using namespace nplda::tc;

constexpr int K = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe(const __grid_constant__ CUtensorMap bmap, const __nv_bfloat16 *A, float *D, int N, int ts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_b, bar_mma, bar_ready;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int nh = N / 2;                                  // B rows held by this CTA
    const int b_lbo = (nh / 8) * 128;                      // k-chunk stride
    const int b_bytes = (K / 8) * b_lbo;
    uint8_t *Bs = smem, *As = smem + 16384;
    if (tid == 0) { mbar_init(&bar_b, 1); mbar_init(&bar_mma, 1); mbar_init(&bar_ready, 2); mbar_fence_init(); }
    if (warp == 0) tmem_alloc2(&tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // B half of this CTA: rows [rank * b_bytes / 128, +b_bytes / 128) of the [rows x 32 u32] image
    if (tid == 0) {
        if (rank == 0) mbar_arrive_expect_tx(&bar_b, 2 * b_bytes);
        tma_load_2d_pair(Bs, &bmap, 0, (int)rank * (b_bytes / 128), &bar_b);
    }
    // A rows of this CTA: 128 rows [rank * 128, +128)
    const int row = rank * 128 + tid;
    const uint32_t a_tmem = tmem + 256;
    if (ts) {   // lane = row, column c holds k = 2c (low half), 2c + 1 (high half)
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) {
                uint32_t lo16 = __bfloat16_as_ushort(A[row * K + 2 * (c0 + j)]);
                uint32_t hi16 = __bfloat16_as_ushort(A[row * K + 2 * (c0 + j) + 1]);
                r[j] = lo16 | (hi16 << 16);
            }
            tmem_st8(a_tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        }
        tmem_st_wait();
        tc_fence_before();
    } else {    // canonical K-major no-swizzle image: (r, k) at (k/8)*2048 + (r/8)*128 + (r%8)*16 + (k%8)*2
        for (int k = 0; k < K; ++k)
            *reinterpret_cast<__nv_bfloat16 *>(As + (k / 8) * 2048 + (tid / 8) * 128 + (tid % 8) * 16 + (k % 8) * 2) = A[row * K + k];
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) mbar_arrive_cluster(&bar_ready, 0);      // both CTAs -> leader: operands of this CTA are in place
    if (rank == 0 && tid == 0) {
        mbar_wait(&bar_ready, 0);
        mbar_wait(&bar_b, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(256, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t bd = make_smem_desc(smem_addr(Bs) + ks * 2 * b_lbo, b_lbo, 128);
            const uint64_t ad = make_smem_desc(smem_addr(As) + ks * 2 * 2048, 2048, 128);
            if (ts) mma2_ts(tmem, a_tmem + ks * 8, bd, idesc, ks > 0);
            else mma2_ss(tmem, ad, bd, idesc, ks > 0);
        }
        mma2_commit_mc(&bar_mma, 3);
    }
    mbar_wait(&bar_mma, 0);                                 // both CTAs: multicast commit
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    int bad_total = 0;
    for (int N : {176, 192}) {
        const int nh = N / 2, b_lbo = (nh / 8) * 128, b_bytes = (K / 8) * b_lbo;
        std::vector<__nv_bfloat16> A(256 * K), B(N * K);
        std::vector<float> Af(256 * K), Bf(N * K);
        srand(7 + N);
        for (int i = 0; i < 256 * K; ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2bfloat16(v); Af[i] = v; }
        for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 4.f; B[i] = __float2bfloat16(v); Bf[i] = v; }
        std::vector<uint8_t> img(2 * b_bytes, 0);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const int r = n / nh, nn = n % nh;
                size_t off = (size_t)r * b_bytes + (k / 8) * b_lbo + (nn / 8) * 128 + (nn % 8) * 16 + (k % 8) * 2;
                *reinterpret_cast<__nv_bfloat16 *>(&img[off]) = B[n * K + k];
            }
        __nv_bfloat16 *dA; uint8_t *dimg; float *dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dimg, img.size()); cudaMalloc(&dD, 256 * N * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
        CUtensorMap m;
        cuuint64_t dims[2] = {32, (cuuint64_t)(img.size() / 128)};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {32, (cuuint32_t)(b_bytes / 128)}, es[2] = {1, 1};
        CUresult cr = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dimg, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
        for (int ts = 0; ts < 2; ++ts) {
            cudaMemset(dD, 0xff, 256 * N * 4);
            probe<<<2, 128, 48 * 1024>>>(m, dA, dD, N, ts);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(256 * N);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0; double worst = 0;
            for (int r = 0; r < 256; ++r)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)Af[r * K + k] * Bf[n * K + k];
                    double d = fabs(ref - D[r * N + n]);
                    if (!(d <= 1e-3)) { if (bad < 4) printf("  mismatch r %d n %d got %g ref %g\n", r, n, D[r * N + n], ref); ++bad; }
                    if (d > worst) worst = d;
                }
            printf("N=%d %s: %s, %d mismatches of %d, worst %.3g\n", N, ts ? "TS" : "SS", cudaGetErrorString(e), bad, 256 * N, worst);
            bad_total += bad;
            if (e != cudaSuccess) return 2;
        }
        cudaFree(dA); cudaFree(dimg); cudaFree(dD);
    }
    printf(bad_total ? "PAIR PROBE FAILED\n" : "PAIR PROBE OK\n");
    return bad_total != 0;
}
