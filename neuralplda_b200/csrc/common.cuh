// Shared definitions for libnplda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nplda.h"

namespace nplda {

// ---- geometry shared by the pack kernel and the score kernels ----------------
constexpr int NP = 192;        // padded layer width: 16 column groups x 12 columns
constexpr int KC = 32;         // k-chunk of the SIMT kernels
constexpr int TILE_PAIRS = 64; // trial pairs per SIMT tile (128 rows: both sides)

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Byte offsets of the packed-weight workspace (see nplda_pack_weights).
struct PackLayout {
    int d_in, d1, d2, k1p, k2p;
    int64_t w1t;   // [k1p][NP] f32   W1^T, zero padded
    int64_t b1;    // [NP]
    int64_t w2t;   // [k2p][NP] f32   NeuralPlda: W2^T      DPlda: Ww^T
    int64_t w3t;   // [k2p][NP] f32   NeuralPlda: unused    DPlda: Wb^T
    int64_t b2;    // [NP]            NeuralPlda: b2        DPlda: ws
    int64_t p;     // [NP]            P = P_sqrt^2
    int64_t q;     // [NP]
    int64_t c;     // [4]             DPlda constant (logistic_regres.bias)
    int64_t fp;    // 2 x u64         content fingerprint of the packed parameters, slot = pack epoch parity (pack.cu)
    int64_t tc;    // tensor-core images (bf16 hi/lo, tcgen05 smem layout), see score_tc.cu
    int64_t tc_bytes;
    int64_t tcx;   // CTA-pair weight images of the pre-split-table kernel, see score_tcx.cu (built with NPLDA_PACK_PAIR)
    int64_t tcx_bytes;
    int64_t tcp;   // bf16 CTA-pair weight images of the materialised-pairs kernel, see score_tcp.cu (always built)
    int64_t tcp_bytes;
    int64_t total;
};

int64_t tc_image_bytes(int d_in, int d1, int d2);   // score_tc.cu
int64_t tcx_image_bytes(int d_in, int d1, int d2);  // score_tcx.cu
int64_t tcp_image_bytes(int d_in, int d1, int d2);  // score_tcp.cu

inline PackLayout make_pack_layout(int d_in, int d1, int d2) {
    PackLayout L;
    L.d_in = d_in; L.d1 = d1; L.d2 = d2;
    L.k1p = round_up(d_in, KC);
    L.k2p = round_up(d1, KC);
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    L.w1t = take((int64_t)L.k1p * NP * 4);
    L.b1 = take(NP * 4);
    L.w2t = take((int64_t)L.k2p * NP * 4);
    L.w3t = take((int64_t)L.k2p * NP * 4);
    L.b2 = take(NP * 4);
    L.p = take(NP * 4);
    L.q = take(NP * 4);
    L.c = take(16);
    L.fp = take(16);
    L.tc_bytes = tc_image_bytes(d_in, d1, d2);
    L.tc = take(L.tc_bytes);
    L.tcx_bytes = tcx_image_bytes(d_in, d1, d2);
    L.tcx = take(L.tcx_bytes);
    L.tcp_bytes = tcp_image_bytes(d_in, d1, d2);
    L.tcp = take(L.tcp_bytes);
    L.total = o;
    return L;
}

inline bool dims_supported(int d_in, int d1, int d2) {
    return d_in >= 1 && d1 >= 1 && d2 >= 1 && d1 <= NP && d2 <= NP;
}

// ---- row table of the trial-list / grid paths (pairs.cu, grid.cu, grid_tc.cu) ----------------
// [n_rows][2 * 176] fp32 rows {A | B}  |  256-byte trailer {u64 fingerprint the rows were built from, at +64: grid
// operand header floats}  |  (1024-aligned) grid operands [n_rows][1536 B] fp16 hi/lo (grid_tc.cu)
constexpr int ROWTAB_LD = 176, ROWTAB_FLOATS = 2 * ROWTAB_LD;
inline int64_t rowtab_trailer_offset(int64_t n_rows) { return n_rows * ROWTAB_FLOATS * (int64_t)sizeof(float); }
inline int64_t rowtab_gtab_offset(int64_t n_rows) { return (rowtab_trailer_offset(n_rows) + 256 + 1023) / 1024 * 1024; }
int64_t gtab_bytes(int64_t n_rows);      // grid_tc.cu

// ---- bookkeeping --------------------------------------------------------------
void count_launch(int n = 1);            // api.cu
int sm_count();                          // api.cu (cached per process)

#define NPLDA_CUDA_TRY(expr)                         \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) return (int)_e;       \
    } while (0)

#define NPLDA_LAUNCH_CHECK()                         \
    do {                                             \
        cudaError_t _e = cudaGetLastError();         \
        if (_e != cudaSuccess) return (int)_e;       \
        ::nplda::count_launch();                     \
    } while (0)

// ---- small device helpers -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async copy global -> shared, zero-filling bytes past src_bytes.
__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float half_warp_sum(float v) {   // over the 16 lanes sharing lane>>4
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
    v = half_warp_sum(v);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace nplda
