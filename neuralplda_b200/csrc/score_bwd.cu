// K3: backward of the score kernels (replaces autograd through
// utils/models.py:366-376 for NeuralPlda and 478-489 for DPlda).
//
// Nothing is saved by the forward.  Per chunk of trial pairs:
//   1. a fused tile kernel recomputes a, u (and y) from x exactly as the
//      forward does, turns dL/dS into dL/dy -> dL/du -> dL/da in registers /
//      shared memory, reduces the bias / P / Q gradients, and writes the three
//      row matrices U, DY (or g*U for DPlda) and DA to the workspace;
//   2. the weight gradients are dense contractions over the rows,
//      dW2 += DY^T U,  dW1 += DA^T X   (DPlda: dWw += GU^T U, dWb += GU1^T U2 + GU2^T U1),
//      done by a split-row SGEMM (gemm_tn) that adds its tiles atomically.
// All gradient outputs are ADDED INTO.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "common.cuh"
#include "simt_tile.cuh"

namespace nplda {

int gemm_tn_tc(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C, int ldc,
               cudaStream_t st);   // gemm_tc.cu: C[n][m] += sum_r A[r][m] B[r][n], N <= 176

bool tc_shape_ok(bool dplda, const PackLayout &L, bool indexed);   // score_tc.cu
int score_tc(bool dplda, const float *x1, const float *x2, const int64_t *i1, const int64_t *i2, int64_t n_rows,
             int32_t *bad_flag, int64_t n, const PackLayout &L, const char *pack, float *scores, int mode, cudaStream_t st,
             float *aout, float *yout, int64_t emit_cap);   // score_tc.cu (EMIT mode when aout != nullptr)

bool tc_dplda_ok(const PackLayout &L);   // score_tc.cu
int score_tc_dplda_emit(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, int which,
                        float *aout, float *yout, int64_t emit_cap, cudaStream_t st);   // score_tc.cu

int64_t tc_rows_image_bytes(int d_in);   // score_tc.cu
int tc_rows_image_pack(const float *Wkn, int64_t ldk, int N, int K, int d_in, uint8_t *img, cudaStream_t st);
int score_tc_rows_emit(const float *xa, const float *xb, int64_t n, int row_width, int d_in, const uint8_t *w1img,
                       const uint8_t *w2img_any, int ksteps2, const float *zeros, float *aout, int64_t emit_cap, cudaStream_t st);
int score_tc_bwd_mid(const float *yrows, const float *arows, int64_t pre_cap, int64_t n, int rw, int d_k, int d1, int d2,
                     const uint8_t *w2t_img, const float *p, const float *q, const float *psq2, const float *dscores,
                     float *U, float *G, float *DA, int64_t out_cap, float *db1, float *db2, float *dq, float *dpsqrt,
                     cudaStream_t st);   // score_tc.cu (BWD instantiation)

namespace bwd {

using namespace simt;

constexpr int64_t CHUNK_PAIRS = 524288;
constexpr int COLACC = 4 * NP;                       // floats of per-CTA column accumulators
constexpr int BWD_SMEM_BYTES = SMEM_BYTES + COLACC * 4 + 16;

struct Pack {   // float offsets inside the backward's private weight image
    int k1p, k2p, k2q;
    int64_t w1t, b1, w2t, w2n, b2, p, q, psq2, rp, ws, total;
};

static Pack make_pack(int d_in, int d1, int d2) {
    Pack P;
    P.k1p = round_up(d_in, KC); P.k2p = round_up(d1, KC); P.k2q = round_up(d2, KC);
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += (n + 63) / 64 * 64; return r; };
    P.w1t = take((int64_t)P.k1p * NP); P.b1 = take(NP);
    P.w2t = take((int64_t)P.k2p * NP); P.w2n = take((int64_t)P.k2q * NP);
    P.b2 = take(NP); P.p = take(NP); P.q = take(NP); P.psq2 = take(NP);
    P.rp = take(2 * (int64_t)P.k2p * NP); P.ws = take(NP);
    P.total = o;
    return P;
}

// ---- weight images ------------------------------------------------------------------
__global__ void pack_bwd_nplda(const float *__restrict__ W1, const float *__restrict__ b1,
                               const float *__restrict__ W2, const float *__restrict__ b2,
                               const float *__restrict__ ps, const float *__restrict__ q, int d_in,
                               int d1, int d2, Pack P, float *__restrict__ out) {
    const int64_t n1 = (int64_t)P.k1p * NP, n2 = (int64_t)P.k2p * NP, n3 = (int64_t)P.k2q * NP;
    const int64_t total = n1 + n2 + n3 + 5 * NP;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        if (e < n1) {
            int k = (int)(e / NP), n = (int)(e % NP);
            out[P.w1t + e] = (n < d1 && k < d_in) ? W1[(int64_t)n * d_in + k] : 0.f;
        } else if (e < n1 + n2) {
            int64_t r = e - n1; int k = (int)(r / NP), n = (int)(r % NP);       // W2^T: [k=b][n=a]
            out[P.w2t + r] = (n < d2 && k < d1) ? W2[(int64_t)n * d1 + k] : 0.f;
        } else if (e < n1 + n2 + n3) {
            int64_t r = e - n1 - n2; int k = (int)(r / NP), n = (int)(r % NP);  // W2:   [k=a][n=b]
            out[P.w2n + r] = (k < d2 && n < d1) ? W2[(int64_t)k * d1 + n] : 0.f;
        } else {
            int r = (int)(e - n1 - n2 - n3), which = r / NP, n = r % NP;
            if (which == 0) out[P.b1 + n] = n < d1 ? b1[n] : 0.f;
            if (which == 1) out[P.b2 + n] = n < d2 ? b2[n] : 0.f;
            if (which == 2) out[P.p + n] = n < d2 ? ps[n] * ps[n] : 0.f;
            if (which == 3) out[P.q + n] = n < d2 ? q[n] : 0.f;
            if (which == 4) out[P.psq2 + n] = n < d2 ? 2.f * ps[n] : 0.f;
        }
    }
}

// DPlda: R = Ww + Ww^T and Pm = Wb + Wb^T stacked along k (both symmetric).
__global__ void pack_bwd_dplda(const float *__restrict__ W1, const float *__restrict__ b1,
                               const float *__restrict__ w_lr, int d_in, int d1, Pack P,
                               float *__restrict__ out) {
    const int64_t n1 = (int64_t)P.k1p * NP, n2 = (int64_t)P.k2p * NP;
    const int64_t total = n1 + 2 * n2 + 2 * NP;
    const float *Wb = w_lr, *Ww = w_lr + (int64_t)d1 * d1, *ws = w_lr + 2 * (int64_t)d1 * d1;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        if (e < n1) {
            int k = (int)(e / NP), n = (int)(e % NP);
            out[P.w1t + e] = (n < d1 && k < d_in) ? W1[(int64_t)n * d_in + k] : 0.f;
        } else if (e < n1 + 2 * n2) {
            int64_t r = e - n1;
            const float *M = r < n2 ? Ww : Wb;
            int64_t rr = r < n2 ? r : r - n2;
            int k = (int)(rr / NP), n = (int)(rr % NP);
            out[P.rp + r] = (k < d1 && n < d1) ? M[(int64_t)k * d1 + n] + M[(int64_t)n * d1 + k] : 0.f;
        } else {
            int r = (int)(e - n1 - 2 * n2), which = r / NP, n = r % NP;
            if (which == 0) out[P.b1 + n] = n < d1 ? b1[n] : 0.f;
            if (which == 1) out[P.ws + n] = n < d1 ? ws[n] : 0.f;
        }
    }
}

// ---- fused tile kernel ----------------------------------------------------------------
struct Args {
    const float *x1, *x2;      // chunk base pointers
    const float *ds;           // chunk base
    int64_t nc;                // pairs in this chunk
    int64_t cap;               // workspace rows per side
    int rw;                    // floats per workspace / activation row: 176 (what the tensor-core kernels emit) when the
                               // layer widths fit, else NP; rows stay 16-byte aligned, columns >= rw are never touched
    int d_in, d1, d2, k1p, k2p, k2q;
    const float *w1t, *b1, *w2t, *w2n, *b2, *p, *q, *psq2, *rp, *ws;
    float *U, *G, *DA;         // [2*cap][NP]; G = DY (NeuralPlda) or g*U (DPlda)
    float *PMU;                // workspace rows: dL/du from the tensor-core pass (NeuralPlda PHASE 2)
    const float *Apre, *Ypre, *Zpre;   // PRE: rows of a, y (DPlda: R u) and (DPlda) Pm u -- areas DA / G / PMU of the
    int64_t pre_cap;                   //      workspace (EMIT passes in the backward) or activations saved by a training
                                       //      forward; side 1 of pair p sits pre_cap rows after side 0
    float *db1, *db2, *dq, *dpsqrt, *dws, *dc;   // may be null
};

__device__ __forceinline__ void col_add(float *colacc, int which, int tx, int j, int e, float v, int lane) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);          // the two row groups of this warp
    if (lane < 16) atomicAdd(colacc + which * NP + 4 * tx + 64 * j + e, v);
}

// PRE: a = W1 x + b1 and y = W2 u + b2 (DPlda: R u, and Pm u in PMU) of every row were written by the tensor-core forward kernel
// in its EMIT mode (score_tc.cu) into the DA and G areas of the workspace; the two recomputations (80 % of this
// kernel's flops) are replaced by loads.  Each thread later overwrites exactly the elements it loaded here with
// dL/da and dL/dy, so the aliasing is race-free.
// PHASE (with PRE, NeuralPlda): 0 = everything in one launch; 1 = up to dL/dy (U, DY rows and the b2 / P / Q sums), the
// product dL/du = dL/dy . W2 then runs on the tensor cores (score_tc_rows_emit) into the PMU area; 2 = the rest
// (dL/da from dL/du, b1 sums).
template <bool DPLDA, bool VEC, bool PRE, int PHASE = 0>
__global__ void __launch_bounds__(NTHREADS, 1) bwd_tile_kernel(Args g) {
    extern __shared__ __align__(16) float smem[];
    float *As = smem;
    float *Ws = smem + 2 * A_STAGE;
    float *Us = Ws + 2 * W_STAGE;
    float *colacc = Us + TM * LDU;       // [4][NP]: db1, db2|dws, dq, dp
    float *gsum = colacc + COLACC;       // [1]: sum of dS (DPlda dc)

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31;
    const int64_t ntiles = (g.nc + TILE_PAIRS - 1) / TILE_PAIRS;
    const int nch1 = g.k1p / KC;
    for (int i = tid; i < COLACC + 1; i += NTHREADS) colacc[i] = 0.f;
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t pair0 = tile * TILE_PAIRS;
        const float *rowp[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int m = (tid + NTHREADS * r) >> 3;
            int64_t pr = min(pair0 + row_pair(m), g.nc - 1);
            rowp[r] = (row_side(m) ? g.x2 : g.x1) + pr * g.d_in;
        }
        float gs[4];
        bool live[4];
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            int64_t pc = pair0 + ty + 16 * jp;
            live[jp] = pc < g.nc;
            gs[jp] = live[jp] ? g.ds[pc] : 0.f;
        }

        // ---- recompute layer 1 ----
        float2 acc[8][6];
        if (PRE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float *arow = g.Apre + ((int64_t)(i & 1) * g.pre_cap + min(pair0 + ty + 16 * (i >> 1), g.nc - 1)) * g.rw + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 v = 4 * tx + 64 * j < g.rw ? *reinterpret_cast<const float4 *>(arow + 64 * j)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[i][2 * j] = make_float2(v.x, v.y);
                    acc[i][2 * j + 1] = make_float2(v.z, v.w);
                }
            }
        } else {
        zero_acc(acc);
        load_a_chunk<VEC>(As, rowp, 0, g.d_in, tid);
        load_w_chunk(Ws, g.w1t, 0, tid);
        cp_async_commit();
        for (int c = 0; c < nch1; ++c) {
            if (c + 1 < nch1) {
                int s = (c + 1) & 1;
                load_a_chunk<VEC>(As + s * A_STAGE, rowp, (c + 1) * KC, g.d_in, tid);
                load_w_chunk(Ws + s * W_STAGE, g.w1t, (c + 1) * KC, tid);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            mma_chunk(acc, As + (c & 1) * A_STAGE, LDA, Ws + (c & 1) * W_STAGE, tx, ty);
            __syncthreads();
        }
        }

        // ---- u = a / max(|a|, eps); keep 1/den per row; U -> smem + workspace ----
        float rden[8];
        {
            float2 bb[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 b = *reinterpret_cast<const float4 *>(g.b1 + 4 * tx + 64 * j);
                bb[2 * j] = make_float2(b.x, b.y);
                bb[2 * j + 1] = make_float2(b.z, b.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    if (!PRE) {                         // the emitted a already carries the bias
                        acc[i][j].x += bb[j].x;
                        acc[i][j].y += bb[j].y;
                    }
                    ss = fmaf(acc[i][j].x, acc[i][j].x, ss);
                    ss = fmaf(acc[i][j].y, acc[i][j].y, ss);
                }
                ss = half_warp_sum(ss);
                float nrm = sqrtf(ss);
                float den = fmaxf(nrm, 1e-12f);
                // below the clamp F.normalize is the linear map a / eps
                rden[i] = nrm > 1e-12f ? 1.f / den : -1e12f;   // sign flags the clamped branch
                const int jp = i >> 1, side = i & 1;
                float *urow = Us + (ty + 16 * i) * LDU + 4 * tx;
                float *grow = g.U + ((int64_t)side * g.cap + pair0 + ty + 16 * jp) * g.rw + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 u;
                    u.x = acc[i][2 * j].x / den;
                    u.y = acc[i][2 * j].y / den;
                    u.z = acc[i][2 * j + 1].x / den;
                    u.w = acc[i][2 * j + 1].y / den;
                    if (PHASE != 2) {
                        *reinterpret_cast<float4 *>(urow + 64 * j) = u;
                        if (live[jp] && 4 * tx + 64 * j < g.rw) *reinterpret_cast<float4 *>(grow + 64 * j) = u;
                    }
                }
            }
        }
        __syncthreads();

        if (PHASE == 2) {
            // dL/du rows from the tensor-core pass
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float *drow = g.PMU + ((int64_t)(i & 1) * g.cap + pair0 + ty + 16 * (i >> 1)) * g.rw + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4 v = 4 * tx + 64 * j < g.rw ? *reinterpret_cast<const float4 *>(drow + 64 * j)
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[i][2 * j] = make_float2(v.x, v.y);
                    acc[i][2 * j + 1] = make_float2(v.z, v.w);
                }
            }
        } else if (!DPLDA) {
            // ---- recompute layer 2, form dL/dy in place ----
            if (PRE) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float *yrow = g.Ypre + ((int64_t)(i & 1) * g.pre_cap + min(pair0 + ty + 16 * (i >> 1), g.nc - 1)) * g.rw + 4 * tx;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const float4 v = 4 * tx + 64 * j < g.rw ? *reinterpret_cast<const float4 *>(yrow + 64 * j)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                        acc[i][2 * j] = make_float2(v.x, v.y);
                        acc[i][2 * j + 1] = make_float2(v.z, v.w);
                    }
                }
            } else {
                layer2_gemm(acc, Us, Ws, g.w2t, g.k2p, tx, ty, tid);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 b2 = PRE ? make_float4(0.f, 0.f, 0.f, 0.f)          // the emitted y already carries b2
                                : *reinterpret_cast<const float4 *>(g.b2 + 4 * tx + 64 * j);
                float4 P = *reinterpret_cast<const float4 *>(g.p + 4 * tx + 64 * j);
                float4 Q = *reinterpret_cast<const float4 *>(g.q + 4 * tx + 64 * j);
                const float bv[4] = {b2.x, b2.y, b2.z, b2.w};
                const float pv[4] = {P.x, P.y, P.z, P.w};
                const float qv[4] = {Q.x, Q.y, Q.z, Q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float sq = 0.f, sp = 0.f, sb = 0.f;
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        float2 &r1 = acc[2 * jp][2 * j + (e >> 1)];
                        float2 &r2 = acc[2 * jp + 1][2 * j + (e >> 1)];
                        float &v1 = (e & 1) ? r1.y : r1.x;
                        float &v2 = (e & 1) ? r2.y : r2.x;
                        const float y1 = v1 + bv[e], y2 = v2 + bv[e], gg = gs[jp];
                        sq += gg * (y1 * y1 + y2 * y2);
                        sp += gg * 2.f * y1 * y2;
                        const float d1v = 2.f * gg * (qv[e] * y1 + pv[e] * y2);
                        const float d2v = 2.f * gg * (qv[e] * y2 + pv[e] * y1);
                        sb += d1v + d2v;
                        v1 = d1v;
                        v2 = d2v;
                    }
                    col_add(colacc, 1, tx, j, e, sb, lane);
                    col_add(colacc, 2, tx, j, e, sq, lane);
                    col_add(colacc, 3, tx, j, e, sp, lane);
                }
            }
            // dy replaces u in shared memory (A operand of du = dy W2) and goes to the workspace
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int jp = i >> 1, side = i & 1;
                float *urow = Us + (ty + 16 * i) * LDU + 4 * tx;
                float *grow = g.G + ((int64_t)side * g.cap + pair0 + ty + 16 * jp) * g.rw + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 v = make_float4(acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y);
                    *reinterpret_cast<float4 *>(urow + 64 * j) = v;
                    if (live[jp] && 4 * tx + 64 * j < g.rw) *reinterpret_cast<float4 *>(grow + 64 * j) = v;
                }
            }
            if (PHASE == 1) continue;                                 // du = dy W2 runs on the tensor cores
            __syncthreads();
            layer2_gemm(acc, Us, Ws, g.w2n, g.k2q, tx, ty, tid);      // du = dy W2
        } else {
            // du = g * (R u_self + Pm u_other + ws);  g*u -> workspace
            if (PRE) {                       // R u rows sit in the G area, Pm u rows in PMU (EMIT passes of score_tc.cu)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t prow = min(pair0 + ty + 16 * (i >> 1), g.nc - 1);
                    const int64_t rself = ((int64_t)(i & 1) * g.pre_cap + prow) * g.rw + 4 * tx;
                    const int64_t roth = ((int64_t)((i & 1) ^ 1) * g.pre_cap + prow) * g.rw + 4 * tx;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (4 * tx + 64 * j < g.rw) {
                            v = *reinterpret_cast<const float4 *>(g.Ypre + rself + 64 * j);
                            const float4 w = *reinterpret_cast<const float4 *>(g.Zpre + roth + 64 * j);
                            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                        }
                        acc[i][2 * j] = make_float2(v.x, v.y);
                        acc[i][2 * j + 1] = make_float2(v.z, v.w);
                    }
                }
            } else {
                layer2_gemm<true>(acc, Us, Ws, g.rp, g.k2p, tx, ty, tid);
            }
            float gtot = 0.f;
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) gtot += gs[jp];
            if (tx == 0) atomicAdd(gsum, gtot);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 w4 = *reinterpret_cast<const float4 *>(g.ws + 4 * tx + 64 * j);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
                float sw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int jp = i >> 1, side = i & 1;
                    const float gg = gs[jp];
                    float4 u = *reinterpret_cast<const float4 *>(Us + (ty + 16 * i) * LDU + 4 * tx + 64 * j);
                    float4 gu = make_float4(gg * u.x, gg * u.y, gg * u.z, gg * u.w);
                    if (live[jp] && 4 * tx + 64 * j < g.rw)
                        *reinterpret_cast<float4 *>(g.G + ((int64_t)side * g.cap + pair0 + ty + 16 * jp) * g.rw +
                                                    4 * tx + 64 * j) = gu;
                    sw[0] += gu.x; sw[1] += gu.y; sw[2] += gu.z; sw[3] += gu.w;
                    acc[i][2 * j].x = gg * (acc[i][2 * j].x + wv[0]);
                    acc[i][2 * j].y = gg * (acc[i][2 * j].y + wv[1]);
                    acc[i][2 * j + 1].x = gg * (acc[i][2 * j + 1].x + wv[2]);
                    acc[i][2 * j + 1].y = gg * (acc[i][2 * j + 1].y + wv[3]);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) col_add(colacc, 1, tx, j, e, sw[e], lane);
            }
        }

        // ---- dL/da = (du - u (u.du)) / |a|   (length-norm backward), DA -> workspace ----
        float sb1[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) sb1[c] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int jp = i >> 1, side = i & 1;
            const int64_t grow = ((int64_t)side * g.cap + pair0 + ty + 16 * jp) * g.rw + 4 * tx;
            float4 u[3];
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (DPLDA) u[j] = *reinterpret_cast<const float4 *>(Us + (ty + 16 * i) * LDU + 4 * tx + 64 * j);
                else u[j] = (live[jp] && 4 * tx + 64 * j < g.rw) ? *reinterpret_cast<const float4 *>(g.U + grow + 64 * j)
                                                                 : make_float4(0, 0, 0, 0);
                dot += u[j].x * acc[i][2 * j].x + u[j].y * acc[i][2 * j].y + u[j].z * acc[i][2 * j + 1].x +
                       u[j].w * acc[i][2 * j + 1].y;
            }
            dot = half_warp_sum(dot);
            const bool clamped = rden[i] < 0.f;
            const float r = clamped ? 1e12f : rden[i];
            if (clamped) dot = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 d;
                d.x = (acc[i][2 * j].x - u[j].x * dot) * r;
                d.y = (acc[i][2 * j].y - u[j].y * dot) * r;
                d.z = (acc[i][2 * j + 1].x - u[j].z * dot) * r;
                d.w = (acc[i][2 * j + 1].y - u[j].w * dot) * r;
                if (live[jp] && 4 * tx + 64 * j < g.rw) {
                    *reinterpret_cast<float4 *>(g.DA + grow + 64 * j) = d;
                    sb1[4 * j + 0] += d.x; sb1[4 * j + 1] += d.y; sb1[4 * j + 2] += d.z; sb1[4 * j + 3] += d.w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) col_add(colacc, 0, tx, j, e, sb1[4 * j + e], lane);
        __syncthreads();
    }

    // ---- flush the per-CTA column sums ----
    __syncthreads();
    for (int c = tid; c < NP; c += NTHREADS) {
        if (g.db1 && c < g.d1) atomicAdd(g.db1 + c, colacc[0 * NP + c]);
        if (!DPLDA) {
            if (g.db2 && c < g.d2) atomicAdd(g.db2 + c, colacc[1 * NP + c]);
            if (g.dq && c < g.d2) atomicAdd(g.dq + c, colacc[2 * NP + c]);
            if (g.dpsqrt && c < g.d2) atomicAdd(g.dpsqrt + c, colacc[3 * NP + c] * g.psq2[c]);   // dP * 2 P_sqrt
        } else {
            if (g.dws && c < g.d1) atomicAdd(g.dws + c, colacc[1 * NP + c]);
        }
    }
    if (DPLDA && tid == 0 && g.dc) atomicAdd(g.dc, gsum[0]);
}

// ---- C[M,N] += A^T B over R rows (split across CTAs, atomically added) -------------------
constexpr int GT_K = 16, GT_M = NP, GT_N = 128;

template <bool VEC>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float *__restrict__ A, int lda, int M,
                                                      const float *__restrict__ B, int ldb, int N,
                                                      int64_t R, int64_t rows_per_cta,
                                                      float *__restrict__ C, int ldc) {
    __shared__ __align__(16) float As[2][GT_K][GT_M];
    __shared__ __align__(16) float Bs[2][GT_K][GT_N];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * GT_N;
    const int64_t r0 = blockIdx.y * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    if (r0 >= r1) return;
    const int nch = (int)((r1 - r0 + GT_K - 1) / GT_K);

    auto load = [&](int s, int c) {
        const int64_t rb = r0 + (int64_t)c * GT_K;
#pragma unroll
        for (int q = 0; q < 3; ++q) {          // A: 16 x 192 floats = 768 float4
            int idx = tid + 256 * q, k = idx / (GT_M / 4), m = (idx % (GT_M / 4)) * 4;
            int64_t r = rb + k;
            bool rok = r < r1;
            const float *src = A + (rok ? r : r0) * lda;
            if (VEC) {
                int bytes = rok ? min(max((M - m) * 4, 0), 16) : 0;
                cp_async16(&As[s][k][m], src + (bytes > 0 ? m : 0), bytes);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bool ok = rok && (m + e) < M;
                    cp_async4(&As[s][k][m + e], src + (ok ? m + e : 0), ok ? 4 : 0);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {          // B: 16 x 128 floats = 512 float4
            int idx = tid + 256 * q, k = idx / (GT_N / 4), n = (idx % (GT_N / 4)) * 4;
            int64_t r = rb + k;
            bool rok = r < r1;
            const float *src = B + (rok ? r : r0) * ldb;
            if (VEC) {
                int bytes = rok ? min(max((N - n0 - n) * 4, 0), 16) : 0;
                cp_async16(&Bs[s][k][n], src + (bytes > 0 ? n0 + n : 0), bytes);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    bool ok = rok && (n0 + n + e) < N;
                    cp_async4(&Bs[s][k][n + e], src + (ok ? n0 + n + e : 0), ok ? 4 : 0);
                }
            }
        }
    };

    float2 acc[12][4];
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

    load(0, 0);
    cp_async_commit();
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) {
            load((c + 1) & 1, c + 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int s = c & 1;
#pragma unroll
        for (int k = 0; k < GT_K; ++k) {
            float4 a[3], b[2];
#pragma unroll
            for (int j = 0; j < 3; ++j) a[j] = *reinterpret_cast<const float4 *>(&As[s][k][4 * ty + 64 * j]);
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = *reinterpret_cast<const float4 *>(&Bs[s][k][4 * tx + 64 * j]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float av[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 aa = make_float2(av[e], av[e]);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        acc[4 * j + e][2 * jj] = __ffma2_rn(aa, make_float2(b[jj].x, b[jj].y), acc[4 * j + e][2 * jj]);
                        acc[4 * j + e][2 * jj + 1] = __ffma2_rn(aa, make_float2(b[jj].z, b[jj].w), acc[4 * j + e][2 * jj + 1]);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = 4 * ty + 64 * j + e;
            if (m >= M) continue;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const int n = n0 + 4 * tx + 64 * jj;
                float *crow = C + (int64_t)m * ldc + n;
                const float2 v0 = acc[4 * j + e][2 * jj], v1 = acc[4 * j + e][2 * jj + 1];
                if (n + 0 < N) atomicAdd(crow + 0, v0.x);
                if (n + 1 < N) atomicAdd(crow + 1, v0.y);
                if (n + 2 < N) atomicAdd(crow + 2, v1.x);
                if (n + 3 < N) atomicAdd(crow + 3, v1.y);
            }
        }
}

static int gemm_tn(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C,
                   int ldc, cudaStream_t st) {
    if (R <= 0 || M <= 0 || N <= 0) return NPLDA_OK;
    if (M > GT_M) return NPLDA_ERR_UNSUPPORTED_DIM;
    const int ntn = (N + GT_N - 1) / GT_N;
    int splits = std::max(1, 2 * sm_count() / ntn);
    int64_t rows = (R + splits - 1) / splits;
    rows = std::max<int64_t>((rows + GT_K - 1) / GT_K * GT_K, 2 * GT_K);   // small batches: more CTAs, fewer chunks each
    splits = (int)((R + rows - 1) / rows);
    const bool vec = lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0;
    dim3 grid(ntn, splits);
    if (vec) gemm_tn_kernel<true><<<grid, 256, 0, st>>>(A, lda, M, B, ldb, N, R, rows, C, ldc);
    else gemm_tn_kernel<false><<<grid, 256, 0, st>>>(A, lda, M, B, ldb, N, R, rows, C, ldc);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

// dx[r, :] = DA[r, :d1] . W1   (rows x d_in); only when the inputs require grad (rare)
__global__ void dx_kernel(const float *__restrict__ DA, int ldda, const float *__restrict__ W1, int d1, int d_in,
                          int64_t R, float *__restrict__ dx) {
    extern __shared__ float da[];
    for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
        for (int a = threadIdx.x; a < d1; a += blockDim.x) da[a] = DA[r * ldda + a];
        __syncthreads();
        for (int k = threadIdx.x; k < d_in; k += blockDim.x) {
            float s = 0.f;
            for (int a = 0; a < d1; ++a) s = fmaf(da[a], W1[(int64_t)a * d_in + k], s);
            dx[r * d_in + k] += s;
        }
        __syncthreads();
    }
}

// room for the forward pack (tensor-core weight images) at the end of the workspace, used by the EMIT path
static int64_t fwd_pack_room(int d_in, int d1, int d2) {
    return make_pack_layout(d_in, d1, d2).total + 512 + tc_rows_image_bytes(NP) + 256 + 1024;   // + dL/du image + zeros
}

static int64_t workspace_bytes(int64_t n, int d_in, int d1, int d2) {
    const int64_t cap = (std::min(n, CHUNK_PAIRS) + TILE_PAIRS - 1) / TILE_PAIRS * TILE_PAIRS;
    return make_pack(d_in, d1, d2).total * 4 + 4 * 2 * cap * NP * 4 + 1024 + fwd_pack_room(d_in, d1, d2);
}

// Path selection of the backward, for A/B tests (nplda_debug_backward_paths): 0 = automatic, 1 = the fp32 piece,
// 2 = the tensor-core piece.  Process-wide; the production default is automatic and nothing reads the environment.
static std::atomic<int> g_force_gemm{0}, g_force_emit{0}, g_force_du{0};

// C[M,N] += A^T B: the tcgen05 bf16x3 kernel for batches worth its launch
}  // namespace bwd
int gemm_tn_auto(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C, int ldc, cudaStream_t st);
namespace bwd {
using nplda::gemm_tn_auto;
}  // namespace bwd
int gemm_tn_auto(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C,
                 int ldc, cudaStream_t st) {
    using namespace bwd;
    const int forced = g_force_gemm.load(std::memory_order_relaxed);
    const bool tc_ok = M <= 176;
    if (tc_ok && (forced == 2 || (forced == 0 && R >= 8192))) return gemm_tn_tc(B, ldb, N, A, lda, M, R, C, ldc, st);
    return gemm_tn(A, lda, M, B, ldb, N, R, C, ldc, st);
}
namespace bwd {

template <bool DPLDA>
static int run(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2, const float *W1,
               const float *b1, const float *W2, const float *b2, const float *ps, const float *q,
               const float *w_lr, const float *dscores, float *dW1, float *db1, float *dW2, float *db2,
               float *dps, float *dq, float *dw_lr, float *dc, float *dx1, float *dx2, void *workspace,
               int64_t workspace_bytes_, cudaStream_t st, const float *act = nullptr) {
    if (n < 0 || !workspace) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    if (!x1 || !x2 || !W1 || !b1 || !dscores) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (workspace_bytes_ < workspace_bytes(n, d_in, d1, d2)) return NPLDA_ERR_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return NPLDA_ERR_BAD_ARG;

    const Pack P = make_pack(d_in, d1, d2);
    float *pk = (float *)workspace;
    const int64_t cap = (std::min(n, CHUNK_PAIRS) + TILE_PAIRS - 1) / TILE_PAIRS * TILE_PAIRS;
    float *U = pk + (P.total + 63) / 64 * 64;
    const int rw = std::max(d1, d2) <= 176 ? 176 : NP;      // row pitch of U / G / DA / PMU and of saved activations
    float *G = U + 2 * cap * NP;
    float *DA = G + 2 * cap * NP;

    if (DPLDA) pack_bwd_dplda<<<2 * sm_count(), 256, 0, st>>>(W1, b1, w_lr, d_in, d1, P, pk);
    else pack_bwd_nplda<<<2 * sm_count(), 256, 0, st>>>(W1, b1, W2, b2, ps, q, d_in, d1, d2, P, pk);
    NPLDA_LAUNCH_CHECK();

    const bool vec = (d_in % 4 == 0) && (((uintptr_t)x1 & 15) == 0) && (((uintptr_t)x2 & 15) == 0);
    // NeuralPlda at the tensor-core kernel's shapes and batches worth the extra launches: layer 1 and layer 2 are
    // not recomputed in fp32 here but emitted by the tcgen05 forward kernel (nplda_debug_backward_paths forces a side)
    const PackLayout FL = make_pack_layout(d_in, d1, d2);
    bool pre = false;
    if (vec && (DPLDA ? tc_dplda_ok(FL) : tc_shape_ok(false, FL, false)) && FL.total <= fwd_pack_room(d_in, d1, d2)) {
        const int f = g_force_emit.load(std::memory_order_relaxed);
        pre = f ? f == 2 : n >= 64;
        if (act) pre = true;               // activations saved by the training forward: nothing to emit here
    }
    if (act && !pre) return NPLDA_ERR_UNSUPPORTED_DIM;
    float *PMU = DA + 2 * cap * NP;                        // DPlda EMIT only
    char *fpack = (char *)(PMU + 2 * cap * NP);
    if (pre && (!act || !DPLDA)) {         // (NeuralPlda keeps the W2 image for the dL/du pass)
        int rc = DPLDA ? dplda_pack_weights(W1, b1, w_lr, nullptr, d_in, d1, fpack, FL.total, 0, st)
                       : nplda_pack_weights(W1, b1, W2, b2, ps, q, d_in, d1, d2, fpack, FL.total, 0, st);
        if (rc != NPLDA_OK) return rc;
    }
    auto kern = pre ? bwd_tile_kernel<DPLDA, true, true>
                    : (vec ? bwd_tile_kernel<DPLDA, true, false> : bwd_tile_kernel<DPLDA, false, false>);
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
    // NeuralPlda with EMIT: dL/du = dL/dy . W2 also on the tensor cores (nplda_debug_backward_paths forces a side): the tile
    // kernel runs as two elementwise phases around a rows-in / rows-out pass of the tcgen05 kernel over the DY rows
    bool du_tc = false;
    uint8_t *duimg = (uint8_t *)(fpack + (FL.total + 255) / 256 * 256);
    float *zeros = (float *)(duimg + (tc_rows_image_bytes(NP) + 255) / 256 * 256);
    if (pre && !DPLDA && d2 <= NP) {
        const int f = g_force_du.load(std::memory_order_relaxed);
        du_tc = f ? f == 2 : true;
    }
    auto kern1 = bwd_tile_kernel<false, true, true, 1>;
    auto kern2 = bwd_tile_kernel<false, true, true, 2>;
    if (du_tc) {
        NPLDA_CUDA_TRY(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
        NPLDA_CUDA_TRY(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES));
        NPLDA_CUDA_TRY(cudaMemsetAsync(zeros, 0, 1024, st));
        // image of M[n = layer-1 index][k = layer-2 index] = W2[k][n], read from the k-major fp32 copy w2n
        int rc = tc_rows_image_pack(pk + P.w2n, NP, d1, d2, NP, duimg, st);
        if (rc != NPLDA_OK) return rc;
    }

    for (int64_t c0 = 0; c0 < n; c0 += CHUNK_PAIRS) {
        const int64_t nc = std::min(CHUNK_PAIRS, n - c0);
        Args a;
        a.x1 = x1 + c0 * d_in; a.x2 = x2 + c0 * d_in; a.ds = dscores + c0; a.nc = nc; a.cap = cap;
        a.rw = rw; a.d_in = d_in; a.d1 = d1; a.d2 = d2; a.k1p = P.k1p; a.k2p = P.k2p; a.k2q = P.k2q;
        a.w1t = pk + P.w1t; a.b1 = pk + P.b1; a.w2t = pk + P.w2t; a.w2n = pk + P.w2n; a.b2 = pk + P.b2;
        a.p = pk + P.p; a.q = pk + P.q; a.psq2 = pk + P.psq2; a.rp = pk + P.rp; a.ws = pk + P.ws;
        a.U = U; a.G = G; a.DA = DA;
        a.db1 = db1; a.db2 = db2; a.dq = dq; a.dpsqrt = dps;
        a.dws = dw_lr ? dw_lr + 2 * (int64_t)d1 * d1 : nullptr; a.dc = dc;
        const int64_t ntiles = (nc + TILE_PAIRS - 1) / TILE_PAIRS;
        a.PMU = PMU;
        if (act) {      // [a | y or R u | Pm u], each [2 n][NP], side 1 n rows after side 0
            a.Apre = act + c0 * rw; a.Ypre = act + (2 * n + c0) * rw; a.Zpre = act + (4 * n + c0) * rw; a.pre_cap = n;
        } else {
            a.Apre = DA; a.Ypre = G; a.Zpre = PMU; a.pre_cap = cap;
        }
        if (pre && !act && !DPLDA) {
            int rc = score_tc(false, a.x1, a.x2, nullptr, nullptr, 0, nullptr, nc, FL, fpack, nullptr, 2, st, DA, G, cap);   // fp16x3 activations
            if (rc != NPLDA_OK) return rc;
        }
        if (pre && !act && DPLDA) {         // a and R u (image 2), then Pm u (image 1)
            int rc = score_tc_dplda_emit(a.x1, a.x2, nc, FL, fpack, 2, DA, G, cap, st);
            if (rc != NPLDA_OK) return rc;
            rc = score_tc_dplda_emit(a.x1, a.x2, nc, FL, fpack, 1, nullptr, PMU, cap, st);
            if (rc != NPLDA_OK) return rc;
        }
        // default: the two elementwise phases folded into the rows pass (one launch); forcing du = 2 keeps the three-launch form
        if (du_tc && rw == 176 && g_force_du.load(std::memory_order_relaxed) == 0) {
            int rc = score_tc_bwd_mid(a.Ypre, a.Apre, a.pre_cap, nc, rw, NP, d1, d2, duimg, pk + P.p, pk + P.q, pk + P.psq2, a.ds,
                                      U, G, DA, cap, db1, db2, dq, dps, st);
            if (rc != NPLDA_OK) return rc;
        } else if (du_tc) {
            const int grid = (int)std::min<int64_t>(ntiles, sm_count());
            kern1<<<grid, NTHREADS, BWD_SMEM_BYTES, st>>>(a);
            NPLDA_LAUNCH_CHECK();
            const uint8_t *img2 = (const uint8_t *)fpack + FL.tc + (tc_rows_image_bytes(d_in) + 255) / 256 * 256;   // W2 image of the forward pack
            int rc = score_tc_rows_emit(G, G + cap * rw, nc, rw, NP, duimg, img2, round_up(d1, 16) / 16, zeros, PMU, cap, st);
            if (rc != NPLDA_OK) return rc;
            kern2<<<grid, NTHREADS, BWD_SMEM_BYTES, st>>>(a);
            NPLDA_LAUNCH_CHECK();
        } else {
            kern<<<(int)std::min<int64_t>(ntiles, sm_count()), NTHREADS, BWD_SMEM_BYTES, st>>>(a);
            NPLDA_LAUNCH_CHECK();
        }

        // Row ranges in the workspace: side 0 rows [0, nc), side 1 rows [cap, cap + nc).
        int rc = NPLDA_OK;
        const float *U0 = U, *U1 = U + cap * rw, *G0 = G, *G1 = G + cap * rw, *DA0 = DA, *DA1 = DA + cap * rw;
        if (dW1) {
            if ((rc = gemm_tn_auto(DA0, rw, d1, a.x1, d_in, d_in, nc, dW1, d_in, st)) != NPLDA_OK) return rc;
            if ((rc = gemm_tn_auto(DA1, rw, d1, a.x2, d_in, d_in, nc, dW1, d_in, st)) != NPLDA_OK) return rc;
        }
        if (!DPLDA && dW2) {
            if ((rc = gemm_tn_auto(G0, rw, d2, U0, rw, d1, nc, dW2, d1, st)) != NPLDA_OK) return rc;
            if ((rc = gemm_tn_auto(G1, rw, d2, U1, rw, d1, nc, dW2, d1, st)) != NPLDA_OK) return rc;
        }
        if (DPLDA && dw_lr) {
            float *dWb = dw_lr, *dWw = dw_lr + (int64_t)d1 * d1;
            if ((rc = gemm_tn_auto(G0, rw, d1, U0, rw, d1, nc, dWw, d1, st)) != NPLDA_OK) return rc;
            if ((rc = gemm_tn_auto(G1, rw, d1, U1, rw, d1, nc, dWw, d1, st)) != NPLDA_OK) return rc;
            if ((rc = gemm_tn_auto(G0, rw, d1, U1, rw, d1, nc, dWb, d1, st)) != NPLDA_OK) return rc;
            if ((rc = gemm_tn_auto(G1, rw, d1, U0, rw, d1, nc, dWb, d1, st)) != NPLDA_OK) return rc;
        }
        if (dx1) {
            dx_kernel<<<(int)std::min<int64_t>(nc, 8 * sm_count()), 256, d1 * 4, st>>>(DA0, rw, W1, d1, d_in, nc, dx1 + c0 * d_in);
            NPLDA_LAUNCH_CHECK();
        }
        if (dx2) {
            dx_kernel<<<(int)std::min<int64_t>(nc, 8 * sm_count()), 256, d1 * 4, st>>>(DA1, rw, W1, d1, d_in, nc, dx2 + c0 * d_in);
            NPLDA_LAUNCH_CHECK();
        }
    }
    return NPLDA_OK;
}

}  // namespace bwd
}  // namespace nplda

using namespace nplda;

extern "C" int64_t nplda_bwd_workspace_bytes(int64_t n, int d_in, int d1, int d2) {
    if (n < 0) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    return bwd::workspace_bytes(n, d_in, d1, d2);
}

extern "C" int nplda_score_bwd(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                               const float *W1, const float *b1, const float *W2, const float *b2,
                               const float *p_sqrt, const float *q, const float *dscores, float *dW1,
                               float *db1, float *dW2, float *db2, float *dp_sqrt, float *dq, float *dx1,
                               float *dx2, void *workspace, int64_t workspace_bytes, void *stream) {
    if (n > 0 && (!W2 || !b2 || !p_sqrt || !q)) return NPLDA_ERR_BAD_ARG;
    return bwd::run<false>(x1, x2, n, d_in, d1, d2, W1, b1, W2, b2, p_sqrt, q, nullptr, dscores, dW1, db1, dW2,
                           db2, dp_sqrt, dq, nullptr, nullptr, dx1, dx2, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

extern "C" int nplda_score_bwd_act(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                                   const float *W1, const float *b1, const float *W2, const float *b2,
                                   const float *p_sqrt, const float *q, const float *dscores, float *dW1,
                                   float *db1, float *dW2, float *db2, float *dp_sqrt, float *dq, float *dx1,
                                   float *dx2, const float *act, void *workspace, int64_t workspace_bytes, void *stream) {
    if (n > 0 && (!W2 || !b2 || !p_sqrt || !q)) return NPLDA_ERR_BAD_ARG;
    return bwd::run<false>(x1, x2, n, d_in, d1, d2, W1, b1, W2, b2, p_sqrt, q, nullptr, dscores, dW1, db1, dW2,
                           db2, dp_sqrt, dq, nullptr, nullptr, dx1, dx2, workspace, workspace_bytes,
                           (cudaStream_t)stream, act);
}

extern "C" int dplda_score_bwd_act(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                                   const float *W1, const float *b1, const float *w_lr, const float *dscores,
                                   float *dW1, float *db1, float *dw_lr, float *dc_lr, float *dx1, float *dx2,
                                   const float *act, void *workspace, int64_t workspace_bytes, void *stream) {
    if (n > 0 && !w_lr) return NPLDA_ERR_BAD_ARG;
    return bwd::run<true>(x1, x2, n, d_in, d1, d1, W1, b1, nullptr, nullptr, nullptr, nullptr, w_lr, dscores,
                          dW1, db1, nullptr, nullptr, nullptr, nullptr, dw_lr, dc_lr, dx1, dx2, workspace,
                          workspace_bytes, (cudaStream_t)stream, act);
}

extern "C" int dplda_score_bwd(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                               const float *W1, const float *b1, const float *w_lr, const float *dscores,
                               float *dW1, float *db1, float *dw_lr, float *dc_lr, float *dx1, float *dx2,
                               void *workspace, int64_t workspace_bytes, void *stream) {
    if (n > 0 && !w_lr) return NPLDA_ERR_BAD_ARG;
    return bwd::run<true>(x1, x2, n, d_in, d1, d1, W1, b1, nullptr, nullptr, nullptr, nullptr, w_lr, dscores,
                          dW1, db1, nullptr, nullptr, nullptr, nullptr, dw_lr, dc_lr, dx1, dx2, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

extern "C" void nplda_debug_backward_paths(int gemm, int emit, int du) {
    bwd::g_force_gemm.store(gemm < 0 || gemm > 2 ? 0 : gemm, std::memory_order_relaxed);
    bwd::g_force_emit.store(emit < 0 || emit > 2 ? 0 : emit, std::memory_order_relaxed);
    bwd::g_force_du.store(du < 0 || du > 2 ? 0 : du, std::memory_order_relaxed);
}
