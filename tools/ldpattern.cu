// Micro-benchmark: HBM bandwidth of the converter access pattern of score_tc.cu.
// Each CTA streams tiles of 128 rows x 2 KB (64 rows from x1, 64 from x2); a warp instruction
// reads 8 rows x 64 B; per "visit" a thread reads V consecutive 64-B pieces of its row pair; PF
// visits are kept in flight.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_stream(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int V, int PF>
__global__ void __launch_bounds__(256) pat(const float *x1, const float *x2, long n, float *out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pl = warp * 8 + (lane >> 2), kq = (lane & 3) * 4;
    const long ntiles = n / 64;
    float acc = 0.f;
    constexpr int VIS = 512 / (16 * V);      // visits per tile
    float4 buf[PF][2 * V];
    long it_tile = blockIdx.x; int vis = 0;
    auto issue = [&](float4 (&b)[2 * V]) {
        if (it_tile < ntiles) {
            const float *r0 = x1 + (it_tile * 64 + pl) * 512 + kq + vis * 16 * V;
            const float *r1 = x2 + (it_tile * 64 + pl) * 512 + kq + vis * 16 * V;
#pragma unroll
            for (int j = 0; j < V; ++j) { b[2 * j] = ldg_stream(r0 + 16 * j); b[2 * j + 1] = ldg_stream(r1 + 16 * j); }
            if (++vis == VIS) { vis = 0; it_tile += gridDim.x; }
        }
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) issue(buf[u]);
    const long mytiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long total = mytiles * VIS;
    for (long it = 0; it < total; it += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            if (it + u < total) {
#pragma unroll
                for (int j = 0; j < 2 * V; ++j) acc += buf[u][j].x + buf[u][j].y + buf[u][j].z + buf[u][j].w;
                issue(buf[u]);
            }
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
template <int V, int PF>
void run(const float *x1, const float *x2, long n, float *out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int grid : {148, 296}) {
        pat<V, PF><<<grid, 256>>>(x1, x2, n, out);
        cudaEventRecord(a);
        for (int i = 0; i < 5; ++i) pat<V, PF><<<grid, 256>>>(x1, x2, n, out);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        printf("V=%d (%4d B/row/visit) PF=%d grid=%d: %.3f ms  %.0f GB/s  (%d KB in flight per CTA)\n", V, 64 * V, PF, grid, ms, n * 4096.0 / ms / 1e6, 256 * PF * 2 * V * 16 / 1024);
    }
}
int main() {
    const long n = 1000000 / 64 * 64;
    float *x1, *x2, *out;
    cudaMalloc(&x1, n * 2048); cudaMalloc(&x2, n * 2048); cudaMalloc(&out, 4);
    cudaMemset(x1, 0, n * 2048); cudaMemset(x2, 0, n * 2048);
    run<2, 4>(x1, x2, n, out);
    run<2, 8>(x1, x2, n, out);
    run<4, 2>(x1, x2, n, out);
    run<4, 4>(x1, x2, n, out);
    run<8, 2>(x1, x2, n, out);
    run<8, 3>(x1, x2, n, out);
    run<16, 1>(x1, x2, n, out);
    run<16, 2>(x1, x2, n, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
