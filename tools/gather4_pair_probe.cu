// Hardware probe: cp.async.bulk.tensor.2d ... tile::gather4 with .cta_group::2 -- both CTAs of a 2-CTA cluster gather four
// table rows into their OWN shared memory while the transaction bytes are posted on the LEADER's mbarrier (what the
// indexed pair kernel needs: the leader issues tcgen05.mma.cta_group::2 once both halves of the A operand have landed).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuralplda_b200/csrc/tc_ptx.cuh"
#include "../neuralplda_b200/csrc/tc_pair_ptx.cuh"
using namespace nplda::tc;

__global__ void __cluster_dims__(2, 1, 1) probe(const __grid_constant__ CUtensorMap m, uint16_t *out, int *status) {
    extern __shared__ __align__(1024) uint8_t sm_raw[];
    __shared__ uint64_t bar;
    uint8_t *sm = sm_raw + ((1024u - (smem_addr(sm_raw) & 1023u)) & 1023u);
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 2048 / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(sm)[i] = 0xFFFF;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        if (rank == 0) mbar_arrive_expect_tx(&bar, 1024);
        const int r0 = rank ? 40 : 5, r1 = rank ? 41 : 17, r2 = rank ? 2 : 3, r3 = rank ? 63 : 60;
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(smem_addr(sm)), "l"(&m), "r"(64), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(&bar) & 0xFEFFFFFFu) : "memory");
        if (rank == 0) {
            int ok = 0;
            for (int spin = 0; spin < 2000000 && !ok; ++spin) {
                uint32_t p;
                asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                             : "=r"(p) : "r"(smem_addr(&bar)), "r"(0) : "memory");
                ok = p;
            }
            status[0] = ok;
        }
    }
    cluster_sync_all();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[rank * 1024 + i] = reinterpret_cast<uint16_t *>(sm)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int U = 64, D = 1024;
    std::vector<uint16_t> h(U * D);
    for (int r = 0; r < U; ++r) for (int c = 0; c < D; ++c) h[r * D + c] = (uint16_t)(r * 1024 + c);
    uint16_t *tab, *out; int *status;
    cudaMalloc(&tab, h.size() * 2); cudaMemcpy(tab, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 4096); cudaMalloc(&status, 4); cudaMemset(status, 0xff, 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)U}, strides[1] = {(cuuint64_t)D * 2};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    CUresult rc = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, tab, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d\n", (int)rc);
    probe<<<2, 128, 4096>>>(m, out, status);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    std::vector<uint16_t> o(2048); int st;
    cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost); cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    printf("leader barrier completed with expect_tx 1024 (2 x 512 from both CTAs): %d\n", st);
    for (int rank = 0; rank < 2; ++rank)
        for (int row = 0; row < 4; ++row) {
            printf("  cta %d smem line %d:", rank, row);
            for (int ch = 0; ch < 8; ++ch) {
                const uint16_t v = o[rank * 1024 + row * 64 + ch * 8];
                if (v == 0xFFFF) printf("  [----]"); else printf("  [r%2d c%3d]", v / 1024, v % 1024);
            }
            printf("\n");
        }
    return 0;
}
