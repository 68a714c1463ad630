// K1 (SIMT): fused pairwise score kernel in fp32 on the CUDA cores.
//
// One persistent CTA per SM walks tiles of 64 trial pairs (128 rows: the two
// sides of every pair live in the same thread so the pair term needs no
// exchange).  Per tile:
//   layer 1   a = x W1^T + b1      register-tiled SGEMM, 8 rows x 12 cols per
//                                  thread, packed FFMA2 (fma.rn.f32x2), x and
//                                  W1^T chunks double-buffered with cp.async
//   norm      u = a / max(|a|,eps) half-warp shuffle reduction, u -> smem
//   layer 2   y = u W2^T + b2      same SGEMM loop, A operand from smem
//   score     S = sum Q y1^2 + Q y2^2 + 2 P y1 y2   (NeuralPlda, models.py:372-376)
//             or the DPlda bilinear form (models.py:483-489) with two layer-2
//             passes (Ww, Wb) and u re-read from smem
// Nothing but the 4-byte score per pair is written to HBM.
#include "common.cuh"
#include "simt_tile.cuh"

namespace nplda {
namespace simt {

struct Args {
    const float *x1, *x2;          // materialised pairs, or both = table when indexed
    const int64_t *i1, *i2;        // indexed layout
    int64_t n_rows;                // table rows (indexed)
    int32_t *bad_flag;
    int64_t n;
    int d_in, k1p, k2p, d_out;     // d_out: embedding width written by the embed modes
    int64_t ld_out;                // row stride of the embedding output (floats)
    const float *w1t, *b1, *w2t, *w3t, *b2, *p, *q, *c;
    float *scores;
    const unsigned long long *fp_cur, *fp_built;   // embed modes: skip the launch when *fp_cur == *fp_built (pairs.cu)
};

// MODE: 0 NeuralPlda score, 1 DPlda score, 2 NeuralPlda embeddings y[n,d2], 3 DPlda embeddings u[n,d1],
//       4 DPlda score from embeddings (x1, x2 hold u rows of width d_in = d1)
template <int MODE, bool INDEXED, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 1) score_kernel(Args g) {
    constexpr bool DPLDA = MODE == 1 || MODE == 3 || MODE == 4;
    constexpr bool EMBED = MODE == 2 || MODE == 3;
    constexpr bool FROM_EMB = MODE == 4;
    constexpr int ROWS_PER_UNIT = EMBED ? 1 : 2;             // an embed tile is 128 input rows, a score tile 64 pairs
    if (EMBED && g.fp_cur != nullptr && *g.fp_cur == *g.fp_built) return;   // row table still valid for these parameters
    extern __shared__ __align__(16) float smem[];
    float *As = smem;
    float *Ws = smem + 2 * A_STAGE;
    float *Us = Ws + 2 * W_STAGE;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    constexpr int UNITS = TM / ROWS_PER_UNIT;
    const int64_t ntiles = (g.n + UNITS - 1) / UNITS;
    const int nch1 = g.k1p / KC;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t pair0 = tile * UNITS;

        // source row pointers of the 4 tile rows this thread copies
        const float *rowp[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int m = (tid + NTHREADS * r) >> 3;
            int64_t pr = min(pair0 + (EMBED ? m : row_pair(m)), g.n - 1);   // tail rows re-read the last one
            int side = EMBED ? 0 : row_side(m);
            if (INDEXED) {
                int64_t ix = side ? g.i2[pr] : g.i1[pr];
                if (ix < 0 || ix >= g.n_rows) { *g.bad_flag = 1; ix = 0; }
                rowp[r] = g.x1 + ix * g.d_in;
            } else {
                rowp[r] = (side ? g.x2 : g.x1) + pr * g.d_in;
            }
        }

        float2 acc[8][6];
        if (FROM_EMB) {
            // the inputs already are length-normalised embeddings: tile rows -> Us, zero padded
            for (int e = tid; e < TM * NP; e += NTHREADS) {
                int m = e / NP, k = e % NP;
                int64_t pr = min(pair0 + row_pair(m), g.n - 1);
                const float *src = (row_side(m) ? g.x2 : g.x1) + pr * g.d_in;
                Us[m * LDU + k] = k < g.d_in ? src[k] : 0.f;
            }
            __syncthreads();
        } else {
        // ---------------- layer 1 ------------------------------------------------
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

        // ---------------- bias + length norm (models.py:367-368), u -> smem -------
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
                    acc[i][j].x += bb[j].x;
                    acc[i][j].y += bb[j].y;
                    ss = fmaf(acc[i][j].x, acc[i][j].x, ss);
                    ss = fmaf(acc[i][j].y, acc[i][j].y, ss);
                }
                ss = half_warp_sum(ss);
                float den = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize eps
                float *urow = Us + (ty + 16 * i) * LDU + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 u;
                    u.x = acc[i][2 * j].x / den;
                    u.y = acc[i][2 * j].y / den;
                    u.z = acc[i][2 * j + 1].x / den;
                    u.w = acc[i][2 * j + 1].y / den;
                    *reinterpret_cast<float4 *>(urow + 64 * j) = u;
                    if (MODE == 3) {                             // DPlda embeddings: u is the output
                        int64_t row = pair0 + ty + 16 * i;
                        const float uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            int c = 4 * tx + 64 * j + e;
                            if (row < g.n && c < g.d_out) g.scores[row * g.ld_out + c] = uv[e];
                        }
                    }
                }
            }
        }
        __syncthreads();
        }   // !FROM_EMB
        if (MODE == 3) continue;

        // ---------------- layer 2 + pair score ------------------------------------
        float part[4];
        if (MODE == 2) {                                         // NeuralPlda embeddings: y = W2 u + b2
            layer2_gemm(acc, Us, Ws, g.w2t, g.k2p, tx, ty, tid);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 b2 = *reinterpret_cast<const float4 *>(g.b2 + 4 * tx + 64 * j);
                const float bv[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int64_t row = pair0 + ty + 16 * i;
                    const float yv[4] = {acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        int c = 4 * tx + 64 * j + e;
                        if (row < g.n && c < g.d_out) g.scores[row * g.ld_out + c] = yv[e] + bv[e];
                    }
                }
            }
            __syncthreads();
            continue;
        }
        if (!DPLDA) {
            layer2_gemm(acc, Us, Ws, g.w2t, g.k2p, tx, ty, tid);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) part[jp] = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float4 b2 = *reinterpret_cast<const float4 *>(g.b2 + 4 * tx + 64 * j);
                float4 P = *reinterpret_cast<const float4 *>(g.p + 4 * tx + 64 * j);
                float4 Q = *reinterpret_cast<const float4 *>(g.q + 4 * tx + 64 * j);
                const float bv[4] = {b2.x, b2.y, b2.z, b2.w};
                const float pv[4] = {P.x, P.y, P.z, P.w};
                const float qv[4] = {Q.x, Q.y, Q.z, Q.w};
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {
                    const float y1v[4] = {acc[2 * jp][2 * j].x, acc[2 * jp][2 * j].y,
                                          acc[2 * jp][2 * j + 1].x, acc[2 * jp][2 * j + 1].y};
                    const float y2v[4] = {acc[2 * jp + 1][2 * j].x, acc[2 * jp + 1][2 * j].y,
                                          acc[2 * jp + 1][2 * j + 1].x, acc[2 * jp + 1][2 * j + 1].y};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float y1 = y1v[e] + bv[e], y2 = y2v[e] + bv[e];
                        part[jp] += qv[e] * (y1 * y1 + y2 * y2) + 2.f * pv[e] * (y1 * y2);
                    }
                }
            }
        } else {
            // within-speaker term: sum_a u_a (Ww u)_a for each of the 8 rows
            layer2_gemm(acc, Us, Ws, g.w2t, g.k2p, tx, ty, tid);
            float within[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                within[i] = 0.f;
                const float *urow = Us + (ty + 16 * i) * LDU + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 u = *reinterpret_cast<const float4 *>(urow + 64 * j);
                    within[i] += u.x * acc[i][2 * j].x + u.y * acc[i][2 * j].y +
                                 u.z * acc[i][2 * j + 1].x + u.w * acc[i][2 * j + 1].y;
                }
            }
            // between-speaker term: u1.(Wb u2) + u2.(Wb u1), and the linear term ws.(u1+u2)
            layer2_gemm(acc, Us, Ws, g.w3t, g.k2p, tx, ty, tid);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                part[jp] = within[2 * jp] + within[2 * jp + 1];
                const float *u1r = Us + (ty + 16 * (2 * jp)) * LDU + 4 * tx;
                const float *u2r = Us + (ty + 16 * (2 * jp + 1)) * LDU + 4 * tx;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 u1 = *reinterpret_cast<const float4 *>(u1r + 64 * j);
                    float4 u2 = *reinterpret_cast<const float4 *>(u2r + 64 * j);
                    float4 ws = *reinterpret_cast<const float4 *>(g.b2 + 4 * tx + 64 * j);
                    part[jp] += u1.x * acc[2 * jp + 1][2 * j].x + u1.y * acc[2 * jp + 1][2 * j].y +
                                u1.z * acc[2 * jp + 1][2 * j + 1].x + u1.w * acc[2 * jp + 1][2 * j + 1].y;
                    part[jp] += u2.x * acc[2 * jp][2 * j].x + u2.y * acc[2 * jp][2 * j].y +
                                u2.z * acc[2 * jp][2 * j + 1].x + u2.w * acc[2 * jp][2 * j + 1].y;
                    part[jp] += ws.x * (u1.x + u2.x) + ws.y * (u1.y + u2.y) + ws.z * (u1.z + u2.z) +
                                ws.w * (u1.w + u2.w);
                }
            }
        }
        const float cst = DPLDA ? g.c[0] : 0.f;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
            float s = half_warp_sum(part[jp]);
            int64_t pr = pair0 + ty + 16 * jp;
            if (tx == 0 && pr < g.n) g.scores[pr] = s + cst;
        }
        __syncthreads();   // Us / Ws are rewritten by the next tile
    }
}

template <int MODE, bool INDEXED>
static int launch(const Args &a, bool vec, cudaStream_t st) {
    const int units = (MODE == 2 || MODE == 3) ? TM : TILE_PAIRS;
    int64_t ntiles = (a.n + units - 1) / units;
    int grid = (int)std::min<int64_t>(ntiles, sm_count());
    auto kern = vec ? score_kernel<MODE, INDEXED, true> : score_kernel<MODE, INDEXED, false>;
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    kern<<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace simt

// Shared by the materialised / indexed and NeuralPlda / DPlda entry points.
int score_simt(bool dplda, const float *x1, const float *x2, const int64_t *i1, const int64_t *i2,
               int64_t n_rows, int32_t *bad_flag, int64_t n, const PackLayout &L, const char *pack,
               float *scores, cudaStream_t st) {
    simt::Args a;
    a.fp_cur = a.fp_built = nullptr;
    a.x1 = x1; a.x2 = x2; a.i1 = i1; a.i2 = i2; a.n_rows = n_rows; a.bad_flag = bad_flag;
    a.n = n; a.d_in = L.d_in; a.k1p = L.k1p; a.k2p = L.k2p;
    a.w1t = (const float *)(pack + L.w1t); a.b1 = (const float *)(pack + L.b1);
    a.w2t = (const float *)(pack + L.w2t); a.w3t = (const float *)(pack + L.w3t);
    a.b2 = (const float *)(pack + L.b2); a.p = (const float *)(pack + L.p);
    a.q = (const float *)(pack + L.q); a.c = (const float *)(pack + L.c);
    a.scores = scores;
    const bool indexed = i1 != nullptr;
    bool vec = (L.d_in % 4 == 0) && (((uintptr_t)x1 & 15) == 0) && (indexed || ((uintptr_t)x2 & 15) == 0);
    a.d_out = 0; a.ld_out = 0;
    if (dplda) return indexed ? simt::launch<1, true>(a, vec, st) : simt::launch<1, false>(a, vec, st);
    return indexed ? simt::launch<0, true>(a, vec, st) : simt::launch<0, false>(a, vec, st);
}

// mode 2: NeuralPlda embeddings, 3: DPlda embeddings, 4: DPlda score from embeddings (x1, x2 = u rows)
int simt_aux(int mode, const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack,
             float *out, int64_t ld_out, cudaStream_t st, const unsigned long long *fp_cur,
             const unsigned long long *fp_built) {
    simt::Args a;
    a.fp_cur = fp_cur; a.fp_built = fp_built;
    a.x1 = x1; a.x2 = x2 ? x2 : x1; a.i1 = a.i2 = nullptr; a.n_rows = 0; a.bad_flag = nullptr;
    a.n = n; a.d_in = mode == 4 ? L.d1 : L.d_in; a.k1p = L.k1p; a.k2p = L.k2p;
    a.w1t = (const float *)(pack + L.w1t); a.b1 = (const float *)(pack + L.b1);
    a.w2t = (const float *)(pack + L.w2t); a.w3t = (const float *)(pack + L.w3t);
    a.b2 = (const float *)(pack + L.b2); a.p = (const float *)(pack + L.p);
    a.q = (const float *)(pack + L.q); a.c = (const float *)(pack + L.c);
    a.scores = out;
    a.d_out = mode == 2 ? L.d2 : L.d1;
    a.ld_out = ld_out > 0 ? ld_out : a.d_out;
    bool vec = (a.d_in % 4 == 0) && (((uintptr_t)x1 & 15) == 0) && (((uintptr_t)a.x2 & 15) == 0);
    if (mode == 2) return simt::launch<2, false>(a, vec, st);
    if (mode == 3) return simt::launch<3, false>(a, vec, st);
    return simt::launch<4, false>(a, false, st);
}

}  // namespace nplda
