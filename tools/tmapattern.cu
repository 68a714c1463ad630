// Micro-benchmark: HBM bandwidth of row-sliced tile loads through the TMA engine.
// A CTA streams tiles of 64 rows from x1 and 64 rows from x2 (2 KB rows); per "visit" it fetches a
// [64 rows x W floats] box from each with cp.async.bulk.tensor.2d into a ring of NS stages; consumer
// warps just wait and release.  Also a variant with per-row 1-D bulk copies.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}

template <int W, int NS, bool ROWWISE>
__global__ void __launch_bounds__(128) pat(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                                           const float *x1, const float *x2, long n, float *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[NS], empty[NS];
    constexpr int BOX = 64 * W * 4;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 96); } mbar_fence_init(); }
    __syncthreads();
    const long ntiles = n / 64;
    const long mytiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr int VIS = 512 / W;
    const long total = mytiles * VIS;
    if (tid < 32) {
        if (tid == 0 || ROWWISE) {
            uint32_t st = 0, ph = 0;
            for (long it = 0; it < total; ++it) {
                const long tile = blockIdx.x + (it / VIS) * gridDim.x; const int vis = it % VIS;
                mbar_wait(&empty[st], ph ^ 1);
                uint8_t *dst = smem + st * 2 * BOX;
                if (!ROWWISE) {
                    mbar_arrive_expect_tx(&full[st], 2 * BOX);
                    tma_2d(dst, &m1, vis * W, (int)(tile * 64), &full[st]);
                    tma_2d(dst + BOX, &m2, vis * W, (int)(tile * 64), &full[st]);
                } else {
                    if (tid == 0) mbar_arrive_expect_tx(&full[st], 2 * BOX);
                    __syncwarp();
                    for (int r = tid; r < 64; r += 32) {
                        bulk_g2s(dst + r * W * 4, x1 + (tile * 64 + r) * 512 + vis * W, W * 4, &full[st]);
                        bulk_g2s(dst + BOX + r * W * 4, x2 + (tile * 64 + r) * 512 + vis * W, W * 4, &full[st]);
                    }
                }
                if (++st == NS) { st = 0; ph ^= 1; }
            }
        }
    } else {
        uint32_t st = 0, ph = 0; float acc = 0;
        for (long it = 0; it < total; ++it) {
            mbar_wait(&full[st], ph);
            acc += reinterpret_cast<float *>(smem + st * 2 * BOX)[tid];
            mbar_arrive(&empty[st]);
            if (++st == NS) { st = 0; ph ^= 1; }
        }
        if (acc == 123.456f) out[0] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int NS, bool ROWWISE>
void run(EncodeFn enc, float *x1, float *x2, long n, float *out) {
    CUtensorMap m1, m2;
    cuuint64_t dims[2] = {512, (cuuint64_t)n}, strides[1] = {2048};
    cuuint32_t box[2] = {W, 64}, es[2] = {1, 1};
    enc(&m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x1, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x2, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    size_t smem = (size_t)NS * 2 * 64 * W * 4 + 1024;
    auto k = pat<W, NS, ROWWISE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<148, 128, smem>>>(m1, m2, x1, x2, n, out);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k<<<148, 128, smem>>>(m1, m2, x1, x2, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("%s W=%3d floats (%4d B/row/visit) stages=%d (%3zu KB smem): %.3f ms  %.0f GB/s  [%s]\n", ROWWISE ? "1D-per-row" : "2D-tensor ", W, W * 4, NS, smem / 1024, ms, n * 4096.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    void *h = dlopen("libcuda.so.1", RTLD_NOW);
    EncodeFn enc = (EncodeFn)dlsym(h, "cuTensorMapEncodeTiled");
    const long n = 1000000 / 64 * 64;
    float *x1, *x2, *out;
    cudaMalloc(&x1, n * 2048); cudaMalloc(&x2, n * 2048); cudaMalloc(&out, 4);
    cudaMemset(x1, 0, n * 2048); cudaMemset(x2, 0, n * 2048);
    run<32, 2, false>(enc, x1, x2, n, out);
    run<32, 4, false>(enc, x1, x2, n, out);
    run<32, 8, false>(enc, x1, x2, n, out);
    run<64, 2, false>(enc, x1, x2, n, out);
    run<64, 3, false>(enc, x1, x2, n, out);
    run<64, 4, false>(enc, x1, x2, n, out);
    run<128, 2, false>(enc, x1, x2, n, out);
    run<128, 3, false>(enc, x1, x2, n, out);
    run<32, 4, true>(enc, x1, x2, n, out);
    run<64, 3, true>(enc, x1, x2, n, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
