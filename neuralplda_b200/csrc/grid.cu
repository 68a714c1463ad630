// Enrol x test GRID scoring over the row table.
//
// Trial lists that are full grids -- every enrol id against every test segment (BASELINE.json configs[2] and
// [3]: 2500 x 4000 and 5000 x 10000), every id against the cohort (utils/adaptive_score_normalization.py reads
// exactly such a matrix, :32) -- do not need an index pair per trial.  With the per-utterance rows of
// nplda_table_prepare (pairs.cu)
//     S[i, j] = r[e_i] + r[t_j] + A[e_i] . B[t_j]
// is one [E,176] x [176,T] product.  The reference scores such lists like any other: gather [B,512] pairs and
// push both sides through both layers per trial (scorefile_generator.py:29-36).
//
// Kernel: fp32 register-tiled SGEMM (packed FFMA2), 128 x 128 scores per CTA, 8 x 8 per thread, K in 11 chunks
// of 16.  A rows are gathered with 16-byte cp.async into [row][k]; B rows are gathered through registers and
// stored transposed [k][col] so that the inner loop reads both operands with conflict-free LDS.128.  HBM
// traffic per trial: the 4-byte score (streaming store); the row table (1408 B per utterance) is L2 resident.
// fp32 throughout, same rounding class as nplda_score_pairs.
#include <algorithm>

#include "common.cuh"

namespace nplda {

int score_grid_tc(const uint8_t *gtab, const float *hdr, int64_t n_rows, const int64_t *er, int64_t E, const int64_t *tr, int64_t T,
                  float *scores, int64_t ld, int32_t *bad_flag, cudaStream_t st);   // grid_tc.cu

namespace grid {

constexpr int ROW_LD = 176, ROW_FLOATS = 2 * ROW_LD;     // row table geometry (pairs.cu)
constexpr int TM = 128, TN = 128, KG = 16, NCH = ROW_LD / KG;
constexpr int LDA = KG + 4;                              // 20 floats: 16-byte aligned rows, broadcast reads
constexpr int NT = 256;
constexpr int A_STAGE = TM * LDA, B_STAGE = KG * TN;

struct Args {
    const float *rowtab;
    int64_t n_rows;
    const int64_t *erow, *trow;
    int64_t ne, nt;
    float *out;
    int64_t ld;
    int32_t *bad_flag;
    int vec_store;
};

__global__ void __launch_bounds__(NT, 2) score_grid_kernel(Args g) {
    __shared__ __align__(16) float As[2 * A_STAGE];
    __shared__ __align__(16) float Bs[2 * B_STAGE];
    __shared__ float re[TM], rt[TN];                     // r of the tile's rows / columns
    __shared__ unsigned char bade[TM], badt[TN];         // index outside the table
    __shared__ int64_t eoff[TM], toff[TN];               // float offsets of the rows in the table

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t tiles_n = (g.nt + TN - 1) / TN;
    const int64_t i0 = (blockIdx.x / tiles_n) * TM, j0 = (blockIdx.x % tiles_n) * TN;

    if (tid < TM) {
        const int64_t i = min(i0 + tid, g.ne - 1);
        int64_t r = g.erow[i];
        const bool bad = r < 0 || r >= g.n_rows;
        if (bad) { *g.bad_flag = 1; r = 0; }
        eoff[tid] = r * ROW_FLOATS;
        bade[tid] = bad;
        re[tid] = g.rowtab[r * ROW_FLOATS + ROW_LD - 1];
    } else {
        const int t = tid - TM;
        const int64_t j = min(j0 + t, g.nt - 1);
        int64_t r = g.trow[j];
        const bool bad = r < 0 || r >= g.n_rows;
        if (bad) { *g.bad_flag = 1; r = 0; }
        toff[t] = r * ROW_FLOATS;
        badt[t] = bad;
        rt[t] = g.rowtab[r * ROW_FLOATS + ROW_LD - 1];
    }
    __syncthreads();

    // copy roles: A -- 2 x 16 B per thread (row = idx >> 2, segment = idx & 3); B -- 32 B of one column
    const float *asrc[2];
    float *adst[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int idx = tid + NT * r, m = idx >> 2, seg = idx & 3;
        asrc[r] = g.rowtab + eoff[m] + seg * 4;
        adst[r] = As + m * LDA + seg * 4;
    }
    const int bn = tid & 127, bk = (tid >> 7) * 8;
    const float *bsrc = g.rowtab + toff[bn] + ROW_LD + bk;

    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

    auto load_a = [&](int c, int s) {
#pragma unroll
        for (int r = 0; r < 2; ++r) cp_async16(adst[r] + s * A_STAGE, asrc[r] + c * KG, 16);
        cp_async_commit();
    };
    auto store_b = [&](const float4 &b0, const float4 &b1, int s) {
        float *d = Bs + s * B_STAGE + bk * TN + bn;
        d[0 * TN] = b0.x; d[1 * TN] = b0.y; d[2 * TN] = b0.z; d[3 * TN] = b0.w;
        d[4 * TN] = b1.x; d[5 * TN] = b1.y; d[6 * TN] = b1.z; d[7 * TN] = b1.w;
    };

    float4 b0 = *reinterpret_cast<const float4 *>(bsrc), b1 = *reinterpret_cast<const float4 *>(bsrc + 4);
    load_a(0, 0);
    store_b(b0, b1, 0);
    for (int c = 0; c < NCH; ++c) {
        const int s = c & 1;
        cp_async_wait<0>();
        __syncthreads();                                 // chunk c has landed; everyone is done reading stage s ^ 1
        if (c + 1 < NCH) {                               // refill stage s ^ 1 behind the barrier, under this chunk's math
            load_a(c + 1, s ^ 1);
            b0 = *reinterpret_cast<const float4 *>(bsrc + (c + 1) * KG);
            b1 = *reinterpret_cast<const float4 *>(bsrc + (c + 1) * KG + 4);
        }
        if (c == NCH - 1) {                              // A[ROW_LD - 1] holds r, not a factor of the dot product
            if (tid < TM) As[s * A_STAGE + tid * LDA + KG - 1] = 0.f;
            __syncthreads();
        }
        const float *A = As + s * A_STAGE, *B = Bs + s * B_STAGE;
#pragma unroll
        for (int k4 = 0; k4 < KG / 4; ++k4) {
            float4 a4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4 *>(A + (ty + 16 * i) * LDA + k4 * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float4 w[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) w[j] = *reinterpret_cast<const float4 *>(B + (k4 * 4 + kk) * TN + 4 * tx + 64 * j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
                    const float2 aa = make_float2(a, a);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        acc[i][2 * j] = __ffma2_rn(aa, make_float2(w[j].x, w[j].y), acc[i][2 * j]);
                        acc[i][2 * j + 1] = __ffma2_rn(aa, make_float2(w[j].z, w[j].w), acc[i][2 * j + 1]);
                    }
                }
            }
        }
        if (c + 1 < NCH) store_b(b0, b1, s ^ 1);         // stage s ^ 1 was last read before this chunk's barrier
    }

    // S = A.B + r_e + r_t; a bad row or column scores 0 (the flag is set), like nplda_score_pairs
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = ty + 16 * i;
        const int64_t gi = i0 + m;
        if (gi >= g.ne) continue;
        const float r1 = re[m];
        const bool b1 = bade[m];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = 4 * tx + 64 * j;
            const int64_t gj = j0 + n;
            float v[4] = {acc[i][2 * j].x, acc[i][2 * j].y, acc[i][2 * j + 1].x, acc[i][2 * j + 1].y};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                v[e] += r1 + rt[n + e];
                if (b1 || badt[n + e]) v[e] = 0.f;
            }
            float *dst = g.out + gi * g.ld + gj;
            if (g.vec_store && gj + 3 < g.nt) {
                __stcs(reinterpret_cast<float4 *>(dst), make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (gj + e < g.nt) __stcs(dst + e, v[e]);
            }
        }
    }
}

}  // namespace grid
}  // namespace nplda

using namespace nplda;

extern "C" int nplda_score_grid_impl(const float *rowtab, int64_t n_rows, const int64_t *enrol_rows, int64_t n_enrol,
                                     const int64_t *test_rows, int64_t n_test, float *scores, int64_t ld_scores,
                                     int32_t *bad_index_flag, int impl, void *stream) {
    if (n_enrol < 0 || n_test < 0 || n_rows < 0 || ld_scores < n_test) return NPLDA_ERR_BAD_ARG;
    if (n_enrol == 0 || n_test == 0) return NPLDA_OK;
    if (!rowtab || !enrol_rows || !test_rows || !scores || !bad_index_flag || n_rows == 0) return NPLDA_ERR_BAD_ARG;
    if (impl != NPLDA_IMPL_SIMT)      // the tensor-core kernel over the fp16 hi/lo operands nplda_table_prepare left behind the rows
        return score_grid_tc((const uint8_t *)rowtab + rowtab_gtab_offset(n_rows),
                             (const float *)((const char *)rowtab + rowtab_trailer_offset(n_rows) + 64), n_rows, enrol_rows, n_enrol,
                             test_rows, n_test, scores, ld_scores, bad_index_flag, (cudaStream_t)stream);
    const int64_t tiles = ((n_enrol + grid::TM - 1) / grid::TM) * ((n_test + grid::TN - 1) / grid::TN);
    if (tiles > 0x7fffffff) return NPLDA_ERR_BAD_ARG;
    grid::Args a{rowtab, n_rows, enrol_rows, test_rows, n_enrol, n_test, scores, ld_scores, bad_index_flag,
                 (ld_scores % 4 == 0 && ((uintptr_t)scores & 15) == 0) ? 1 : 0};
    grid::score_grid_kernel<<<(unsigned)tiles, grid::NT, 0, (cudaStream_t)stream>>>(a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_score_grid(const float *rowtab, int64_t n_rows, const int64_t *enrol_rows, int64_t n_enrol,
                                const int64_t *test_rows, int64_t n_test, float *scores, int64_t ld_scores,
                                int32_t *bad_index_flag, void *stream) {
    return nplda_score_grid_impl(rowtab, n_rows, enrol_rows, n_enrol, test_rows, n_test, scores, ld_scores, bad_index_flag,
                                 NPLDA_IMPL_AUTO, stream);
}
