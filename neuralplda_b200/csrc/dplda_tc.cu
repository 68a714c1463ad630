// DPlda forward on the tensor cores (models.py:478-495 in the closed form of SURVEY.md 8 a-6):
//     u = normalize(W1 x + b1)
//     S = u1^T Pm u2 + u1^T Ww u1 + u2^T Ww u2 + ws.(u1 + u2) + c ,   Pm = Wb + Wb^T
// The reference materialises a 57 970-wide feature vector per trial (models.py:483-489).  Scoring (and training with the
// LDA frozen, the reference driver's case) is ONE pass of the tcgen05 score kernel in its DPL form (score_tc.cu): x read
// once, both square products per tile, no workspace.  Training with a trainable LDA keeps the two-pass form below: two
// EMIT passes leave, per row, a = W1 x + b1, R u and Z = Pm u for the backward, and one warp per pair finishes
// S = (a1.Z2 + a1.V1 / 2 + ws.a1) / |a1| + (a2.V2 / 2 + ws.a2) / |a2| + c.
#include <algorithm>

#include "common.cuh"

namespace nplda {

bool tc_dplda_ok(const PackLayout &L);   // score_tc.cu
int score_tc_dplda_emit(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, int which,
                        float *aout, float *yout, int64_t emit_cap, cudaStream_t st);   // score_tc.cu
int score_tc_dplda_fused(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                         float *uout, int64_t emit_cap, cudaStream_t st);                // score_tc.cu

namespace dtc {

constexpr int64_t CHUNK_PAIRS = 131072;
constexpr int LD = 176;                  // floats per emitted row (EMIT_LD of score_tc.cu)

__global__ void __launch_bounds__(256) dplda_finish_kernel(const float *__restrict__ A, const float *__restrict__ V,
                                                           const float *__restrict__ Z, int64_t cap, int64_t nc,
                                                           const float *__restrict__ ws, const float *__restrict__ c,
                                                           float *__restrict__ scores, float vscale) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = w0; p < nc; p += nw) {
        const float4 *a1 = reinterpret_cast<const float4 *>(A + p * LD), *a2 = reinterpret_cast<const float4 *>(A + (cap + p) * LD);
        const float4 *v1 = reinterpret_cast<const float4 *>(V + p * LD), *v2 = reinterpret_cast<const float4 *>(V + (cap + p) * LD);
        const float4 *z2 = reinterpret_cast<const float4 *>(Z + (cap + p) * LD);
        const float4 *w4 = reinterpret_cast<const float4 *>(ws);
        float n1 = 0.f, n2 = 0.f, t1 = 0.f, t2 = 0.f;
        for (int k = lane; k < 176 / 4; k += 32) {
            const float4 x1 = a1[k], x2 = a2[k], y1 = v1[k], y2 = v2[k], zz = z2[k], w = w4[k];
            n1 += x1.x * x1.x + x1.y * x1.y + x1.z * x1.z + x1.w * x1.w;
            n2 += x2.x * x2.x + x2.y * x2.y + x2.z * x2.z + x2.w * x2.w;
            // vscale = 1 with V = Ww u;  1/2 with V = (Ww + Ww^T) u: a quadratic form only sees the symmetric part
            t1 += x1.x * (zz.x + vscale * y1.x + w.x) + x1.y * (zz.y + vscale * y1.y + w.y) + x1.z * (zz.z + vscale * y1.z + w.z) +
                  x1.w * (zz.w + vscale * y1.w + w.w);
            t2 += x2.x * (vscale * y2.x + w.x) + x2.y * (vscale * y2.y + w.y) + x2.z * (vscale * y2.z + w.z) + x2.w * (vscale * y2.w + w.w);
        }
        n1 = warp_sum(n1); n2 = warp_sum(n2); t1 = warp_sum(t1); t2 = warp_sum(t2);
        if (lane == 0) scores[p] = t1 / fmaxf(sqrtf(n1), 1e-12f) + t2 / fmaxf(sqrtf(n2), 1e-12f) + c[0];
    }
}

static int64_t cap_for(int64_t n) { return (std::min(n, CHUNK_PAIRS) + TILE_PAIRS - 1) / TILE_PAIRS * TILE_PAIRS; }

}  // namespace dtc

int dplda_score_tc(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                   cudaStream_t st) {
    return score_tc_dplda_fused(x1, x2, n, L, pack, scores, nullptr, 0, st);
}

// Training forward with the LDA frozen (xvector_DPlda_pytorch.py:140-147): scores and the normalised rows u, [2 n][176],
// side 1 n rows after side 0 -- the gradient of logistic_regres (dplda_lr_bwd, gemm_tc.cu) needs nothing else.
int dplda_score_tc_train_u(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                           float *urows, cudaStream_t st) {
    return score_tc_dplda_fused(x1, x2, n, L, pack, scores, urows, n, st);
}

// Training forward: the rows the backward needs -- a, R u (R = Ww + Ww^T) and Pm u, each [2 n][176], side 1 n rows
// after side 0 (176 floats per row) -- are produced for the whole batch and kept by the caller; the score uses u^T Ww u = u^T R u / 2.
int dplda_score_tc_train(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                         float *act, cudaStream_t st) {
    if (!tc_dplda_ok(L)) return NPLDA_ERR_UNSUPPORTED_DIM;
    float *A = act, *RU = act + 2 * n * dtc::LD, *Z = act + 4 * n * dtc::LD;
    int rc = score_tc_dplda_emit(x1, x2, n, L, pack, 2, A, RU, n, st);
    if (rc != NPLDA_OK) return rc;
    rc = score_tc_dplda_emit(x1, x2, n, L, pack, 1, nullptr, Z, n, st);
    if (rc != NPLDA_OK) return rc;
    const int grid = (int)std::min<int64_t>((n + 7) / 8, 16 * (int64_t)sm_count());
    dtc::dplda_finish_kernel<<<grid, 256, 0, st>>>(A, RU, Z, n, n, (const float *)(pack + L.b2), (const float *)(pack + L.c),
                                                    scores, 0.5f);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda

using namespace nplda;

extern "C" int64_t dplda_fwd_workspace_bytes(int64_t n, int d_in, int d1) {
    if (n < 0) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d1)) return NPLDA_ERR_UNSUPPORTED_DIM;
    return 0;            // the fused kernel needs none (kept in the ABI: callers size their workspace with it)
}
