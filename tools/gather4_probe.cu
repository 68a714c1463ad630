// Hardware probe: semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a (no public docs in this image).
//   gather4_probe <box_rows 1|4> <swizzle 0 none | 1 128B | 2 64B> <box_cols>
// A [64 rows x 512] bf16 table holds value = row * 512 + col (mod 65536); one thread gathers rows {5, 17, 3, 60} at
// column 64 into shared memory and the CTA dumps what landed.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
using namespace nplda::tc;

__global__ void probe(const __grid_constant__ CUtensorMap m, int col, int r0, int r1, int r2, int r3, uint32_t tx,
                      uint16_t *out, int *status) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    __shared__ uint64_t bar;
    uint8_t *sm = sm_raw + ((1024u - (smem_addr(sm_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 4096 / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(sm)[i] = 0xFFFF;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, tx);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(smem_addr(sm)), "l"(&m), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(&bar)) : "memory");
        int ok = 0;
        for (int spin = 0; spin < 2000000 && !ok; ++spin) {
            uint32_t p;
            asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                         : "=r"(p) : "r"(smem_addr(&bar)), "r"(0) : "memory");
            ok = p;
        }
        status[0] = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4096 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(sm)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1, swz = argc > 2 ? atoi(argv[2]) : 0, box_cols = argc > 3 ? atoi(argv[3]) : 64;
    const int U = 64, D = 512;
    std::vector<uint16_t> h(U * D);
    for (int r = 0; r < U; ++r) for (int c = 0; c < D; ++c) h[r * D + c] = (uint16_t)(r * 512 + c);
    uint16_t *tab, *out; int *status;
    cudaMalloc(&tab, h.size() * 2); cudaMemcpy(tab, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 4096); cudaMalloc(&status, 4); cudaMemset(status, 0xff, 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)U}, strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUtensorMapSwizzle sw = swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult rc = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, tab, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box_rows %d swizzle %d box_cols %d: encode rc %d\n", box_rows, swz, box_cols, (int)rc);
    if (rc != CUDA_SUCCESS) return 0;
    const uint32_t tx = 4u * box_cols * 2;
    probe<<<1, 128, 8192>>>(m, 64, 5, 17, 3, 60, tx, out, status);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<uint16_t> o(2048); int st;
    cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost); cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    printf("  barrier completed with expect_tx %u: %d\n", tx, st);
    for (int row = 0; row < 8; ++row) {      // 128-byte lines of shared memory
        printf("  smem line %d:", row);
        for (int ch = 0; ch < 8; ++ch) {     // first element of each 16-byte chunk
            const uint16_t v = o[row * 64 + ch * 8];
            if (v == 0xFFFF) printf("  [----]"); else printf("  [r%2d c%3d]", v / 512, v % 512);
        }
        printf("\n");
    }
    return 0;
}
