// Register-tiled fp32 SGEMM building blocks shared by the SIMT score kernels
// (forward: score_simt.cu, backward: score_bwd.cu).
#pragma once
#include "common.cuh"

namespace nplda {
namespace simt {

constexpr int TM = 2 * TILE_PAIRS;   // 128 rows per tile
constexpr int LDA = KC + 4;          // 36 floats: conflict-free float4 row reads
constexpr int LDU = NP + 4;          // 196
constexpr int NTHREADS = 256;
constexpr int A_STAGE = TM * LDA;    // floats
constexpr int W_STAGE = KC * NP;
constexpr int SMEM_FLOATS = 2 * A_STAGE + 2 * W_STAGE + TM * LDU;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;   // 186,368


// tile row m (0..127)  <->  pair (m & 15) + 16 * (m >> 5), side (m >> 4) & 1
__device__ __forceinline__ int row_pair(int m) { return (m & 15) + ((m >> 5) << 4); }
__device__ __forceinline__ int row_side(int m) { return (m >> 4) & 1; }

template <bool VEC>
__device__ __forceinline__ void load_a_chunk(float *As, const float *const (&rowp)[4], int kc0,
                                             int d_in, int tid) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int idx = tid + NTHREADS * r;
        int m = idx >> 3, seg = idx & 7;
        int k = kc0 + seg * 4;
        float *dst = As + m * LDA + seg * 4;
        if (VEC) {
            int bytes = min(max((d_in - k) * 4, 0), 16);
            cp_async16(dst, rowp[r] + (bytes > 0 ? k : 0), bytes);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool ok = (k + e) < d_in;
                cp_async4(dst + e, rowp[r] + (ok ? k + e : 0), ok ? 4 : 0);
            }
        }
    }
}

__device__ __forceinline__ void load_w_chunk(float *Ws, const float *wt, int kc0, int tid) {
    const float *src = wt + (int64_t)kc0 * NP;
#pragma unroll
    for (int r = 0; r < (W_STAGE / 4) / NTHREADS; ++r) {   // 6
        int idx = tid + NTHREADS * r;
        cp_async16(Ws + idx * 4, src + idx * 4, 16);
    }
}

// acc[i][2j+h] holds columns 4*tx + 64*j + 2h, +1 of row ty + 16*i
// SWAP reads the A row of the partner side (row i ^ 1) instead of the thread's own.
template <bool SWAP = false>
__device__ __forceinline__ void mma_chunk(float2 (&acc)[8][6], const float *A, int lda,
                                          const float *Ws, int tx, int ty) {
#pragma unroll
    for (int k4 = 0; k4 < KC / 4; ++k4) {
        float4 a4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            a4[i] = *reinterpret_cast<const float4 *>(A + (ty + 16 * (SWAP ? (i ^ 1) : i)) * lda + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float4 w[3];
#pragma unroll
            for (int j = 0; j < 3; ++j)
                w[j] = *reinterpret_cast<const float4 *>(Ws + (k4 * 4 + kk) * NP + 4 * tx + 64 * j);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
                float2 aa = make_float2(a, a);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    acc[i][2 * j] = __ffma2_rn(aa, make_float2(w[j].x, w[j].y), acc[i][2 * j]);
                    acc[i][2 * j + 1] = __ffma2_rn(aa, make_float2(w[j].z, w[j].w), acc[i][2 * j + 1]);
                }
            }
        }
    }
}

__device__ __forceinline__ void zero_acc(float2 (&acc)[8][6]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = make_float2(0.f, 0.f);
}

// layer-2 style product: A operand is the resident u tile; W^T chunks stream.
// PAIRED: the weight image has 2*k2p rows; the second k2p rows multiply the
// partner side's row (used by the DPlda backward: du1 = R u1 + P u2).
template <bool PAIRED = false>
__device__ __forceinline__ void layer2_gemm(float2 (&acc)[8][6], const float *Us, float *Ws,
                                            const float *wt, int k2p, int tx, int ty, int tid) {
    zero_acc(acc);
    const int nhalf = k2p / KC;
    const int nch = PAIRED ? 2 * nhalf : nhalf;
    load_w_chunk(Ws, wt, 0, tid);
    cp_async_commit();
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) {
            load_w_chunk(Ws + ((c + 1) & 1) * W_STAGE, wt, (c + 1) * KC, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (PAIRED && c >= nhalf)
            mma_chunk<true>(acc, Us + (c - nhalf) * KC, LDU, Ws + (c & 1) * W_STAGE, tx, ty);
        else
            mma_chunk<false>(acc, Us + c * KC, LDU, Ws + (c & 1) * W_STAGE, tx, ty);
        __syncthreads();
    }
}


}  // namespace simt
}  // namespace nplda
