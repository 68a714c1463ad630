// Gradient of DPlda's logistic_regres (autograd through models.py:483-489) from the normalised rows u kept by the
// training forward (dplda_score_fwd_train_u) -- the whole backward when the LDA is frozen, which is how the reference's
// driver trains (xvector_DPlda_pytorch.py:140-147).
//
// With s = u1 + u2, d = u1 - u2 and g = dL/dS:
//     u1 u1^T + u2 u2^T = (s s^T + d d^T) / 2        u1 u2^T + u2 u1^T = (s s^T - d d^T) / 2
// so   dWw = (Ps + Pd) / 2,  dWb = (Ps - Pd) / 2,  Ps = sum_p g_p s_p s_p^T,  Pd = sum_p g_p d_p d_p^T
// -- two batch contractions instead of four (and dws = sum_p g_p s_p, dc = sum_p g_p).  Per chunk of pairs one
// elementwise kernel writes the rows S, D, gS, gD (and reduces dws, dc), the two contractions run on tcgen05
// (gemm_tc.cu, bf16x3), and a small kernel folds Ps, Pd into the gradient.
#include <algorithm>

#include "common.cuh"

namespace nplda {

int gemm_tn_auto(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C, int ldc,
                 cudaStream_t st);   // score_bwd.cu: C[m][n] += sum_r A[r][m] B[r][n]

namespace dlr {

constexpr int LD = 176;                   // floats per row (EMIT_LD of score_tc.cu)
constexpr int64_t CHUNK_PAIRS = 524288;

__global__ void __launch_bounds__(256) sd_rows_kernel(const float *__restrict__ U0, const float *__restrict__ U1,
                                                      const float *__restrict__ g, int64_t nc, float *__restrict__ S,
                                                      float *__restrict__ D, float *__restrict__ GS, float *__restrict__ GD,
                                                      float *__restrict__ dws, int d1, float *__restrict__ dc) {
    __shared__ float4 col[8][LD / 4];
    __shared__ float gsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    float gacc = 0.f;
    for (int64_t p = w0; p < nc; p += nw) {
        const float gg = g[p];
        gacc += gg;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = lane + 32 * h;
            if (k >= LD / 4) continue;
            const float4 a = reinterpret_cast<const float4 *>(U0 + p * LD)[k], b = reinterpret_cast<const float4 *>(U1 + p * LD)[k];
            const float4 s = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
            const float4 gs = make_float4(gg * s.x, gg * s.y, gg * s.z, gg * s.w);
            reinterpret_cast<float4 *>(S + p * LD)[k] = s;
            reinterpret_cast<float4 *>(D + p * LD)[k] = d;
            reinterpret_cast<float4 *>(GS + p * LD)[k] = gs;
            reinterpret_cast<float4 *>(GD + p * LD)[k] = make_float4(gg * d.x, gg * d.y, gg * d.z, gg * d.w);
            acc[h].x += gs.x; acc[h].y += gs.y; acc[h].z += gs.z; acc[h].w += gs.w;
        }
    }
    col[warp][lane] = acc[0];
    if (lane + 32 < LD / 4) col[warp][lane + 32] = acc[1];
    if (lane == 0) gsum[warp] = gacc;
    __syncthreads();
    if (dws != nullptr)
        for (int c = threadIdx.x; c < d1; c += blockDim.x) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += reinterpret_cast<const float *>(col[w])[c];
            atomicAdd(dws + c, v);
        }
    if (dc != nullptr && threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += gsum[w];
        atomicAdd(dc, v);
    }
}

__global__ void fold_kernel(const float *__restrict__ Ps, const float *__restrict__ Pd, int d1, float *__restrict__ dWb,
                            float *__restrict__ dWw) {
    const int total = d1 * d1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const float a = Ps[e], b = Pd[e];
        dWw[e] += 0.5f * (a + b);
        dWb[e] += 0.5f * (a - b);
    }
}

static int64_t cap_for(int64_t n) { return std::min(n, CHUNK_PAIRS); }
static int64_t p_floats(int d1) { return ((int64_t)d1 * d1 + 63) / 64 * 64; }

}  // namespace dlr

int64_t dplda_lr_workspace_bytes(int64_t n, int d1) { return (4 * dlr::cap_for(n) * dlr::LD + 2 * dlr::p_floats(d1)) * 4; }

// urows: [2 cap][176], side 1 `cap` rows after side 0
int dplda_lr_grad(const float *urows, int64_t n, int64_t cap, int d1, const float *dscores, float *dw_lr, float *dc_lr,
                  void *workspace, int64_t workspace_bytes, cudaStream_t st) {
    if (!workspace || ((uintptr_t)workspace & 255) != 0) return NPLDA_ERR_BAD_ARG;
    if (workspace_bytes < dplda_lr_workspace_bytes(n, d1)) return NPLDA_ERR_WORKSPACE;
    const int64_t wc = dlr::cap_for(n);
    float *Ps = (float *)workspace, *Pd = Ps + dlr::p_floats(d1);
    float *S = Pd + dlr::p_floats(d1), *D = S + wc * dlr::LD, *GS = D + wc * dlr::LD, *GD = GS + wc * dlr::LD;
    float *dWb = dw_lr, *dWw = dw_lr ? dw_lr + (int64_t)d1 * d1 : nullptr, *dws = dw_lr ? dw_lr + 2 * (int64_t)d1 * d1 : nullptr;
    if (dw_lr) NPLDA_CUDA_TRY(cudaMemsetAsync(Ps, 0, 2 * dlr::p_floats(d1) * 4, st));
    for (int64_t c0 = 0; c0 < n; c0 += dlr::CHUNK_PAIRS) {
        const int64_t nc = std::min(dlr::CHUNK_PAIRS, n - c0);
        const int grid = (int)std::min<int64_t>((nc + 7) / 8, 8 * (int64_t)sm_count());
        dlr::sd_rows_kernel<<<grid, 256, 0, st>>>(urows + c0 * dlr::LD, urows + (cap + c0) * dlr::LD, dscores + c0, nc, S, D, GS, GD,
                                                  dws, d1, dc_lr);
        NPLDA_LAUNCH_CHECK();
        if (!dw_lr) continue;
        int rc = gemm_tn_auto(GS, dlr::LD, d1, S, dlr::LD, d1, nc, Ps, d1, st);
        if (rc != NPLDA_OK) return rc;
        rc = gemm_tn_auto(GD, dlr::LD, d1, D, dlr::LD, d1, nc, Pd, d1, st);
        if (rc != NPLDA_OK) return rc;
    }
    if (dw_lr) {
        dlr::fold_kernel<<<32, 256, 0, st>>>(Ps, Pd, d1, dWb, dWw);
        NPLDA_LAUNCH_CHECK();
    }
    return NPLDA_OK;
}

}  // namespace nplda
