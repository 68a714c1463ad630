// Trial-list scoring with every utterance transformed ONCE.
//
// The reference gathers [B,512] x-vector pairs per batch and pushes both sides through the two affine
// layers again for every trial (sv_trials_loaders.py:418-437 + models.py:378-382).  A trial list over U
// unique utterances only needs U transforms: both models' scores split into per-utterance terms plus one
// 170-long dot product per trial,
//     S(i, j) = r[i] + r[j] + A[i] . B[j]
//   NeuralPlda (models.py:372-376):  y = W2 normalize(W1 x + b1) + b2,  A = y,  B = 2 P_sqrt^2 * y,  r = sum Q y^2
//   DPlda      (models.py:483-489):  u = normalize(W1 x + b1),          A = u,  B = (Wb + Wb^T) u,
//                                    r = u^T Ww u + ws . u + c / 2
// nplda_table_prepare builds the row table [n_rows][2 * ROW_LD] = { A (ROW_LD floats, r stored in A[ROW_LD-1]) |
// B (ROW_LD floats) } with the fp32 SIMT embed kernel; nplda_score_pairs streams the index pairs (16 B per
// trial from HBM) and gathers two 704-byte rows per trial, which stay in L2 for tables up to ~80 k utterances.
//
// Validity of a cached row table is decided ON THE DEVICE: the table carries the fingerprint of the packed
// parameters it was built from (pack.cu); with NPLDA_PREPARE_IF_CHANGED the prepare kernels return at once when
// it equals the fingerprint of the current pack, so a caller may re-pack and "prepare" before every scoring call
// at the cost of three empty launches, and never scores with rows of stale parameters -- however the parameters
// were changed (`.data.copy_()`, fused optimisers: neither bumps tensor._version).
#include <algorithm>

#include "common.cuh"

namespace nplda {

int simt_aux(int mode, const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack,
             float *out, int64_t ld_out, cudaStream_t st, const unsigned long long *fp_cur = nullptr,
             const unsigned long long *fp_built = nullptr);   // score_simt.cu
int gtab_prepare(const float *rowtab, int64_t n_rows, int row_floats, int row_ld, int d, float *hdr, uint8_t *gtab,
                 const unsigned long long *fp_cur, const unsigned long long *fp_built, cudaStream_t st);   // grid_tc.cu

constexpr int ROW_LD = 176;              // floats per half row (width <= 175; the last float of A holds r)
constexpr int ROW_FLOATS = 2 * ROW_LD;

// one warp per utterance: B = 2 P y, r = sum Q y^2 (y already sits in the A half)
__global__ void __launch_bounds__(256) rowtab_nplda_kernel(float *__restrict__ rowtab, int64_t n_rows, int d,
                                                           const float *__restrict__ p, const float *__restrict__ q,
                                                           const unsigned long long *fp_cur, const unsigned long long *fp_built) {
    if (fp_cur != nullptr && *fp_cur == *fp_built) return;
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w0; r < n_rows; r += nw) {
        float *A = rowtab + r * ROW_FLOATS, *B = A + ROW_LD;
        float acc = 0.f;
        for (int k = lane; k < ROW_LD; k += 32) {
            const float y = k < d ? A[k] : 0.f;
            if (k >= d && k < ROW_LD - 1) A[k] = 0.f;
            B[k] = k < d ? 2.f * p[k] * y : 0.f;
            acc = fmaf(k < d ? q[k] : 0.f, y * y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) A[ROW_LD - 1] = acc;
    }
}

// one CTA per utterance: B = (Wb + Wb^T) u, r = u^T Ww u + ws . u + c / 2.  wwt[k][n] = Ww[n][k], wbt[k][n] = Wb[n][k].
__global__ void __launch_bounds__(NP) rowtab_dplda_kernel(float *__restrict__ rowtab, int64_t n_rows, int d,
                                                          const float *__restrict__ wwt, const float *__restrict__ wbt,
                                                          const float *__restrict__ ws, const float *__restrict__ c,
                                                          const unsigned long long *fp_cur, const unsigned long long *fp_built) {
    if (fp_cur != nullptr && *fp_cur == *fp_built) return;
    __shared__ float u[NP];
    __shared__ float red[NP / 32];
    const int n = threadIdx.x;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        float *A = rowtab + r * ROW_FLOATS, *B = A + ROW_LD;
        __syncthreads();
        u[n] = n < d ? A[n] : 0.f;
        __syncthreads();
        float pm = 0.f, ww = 0.f;
        if (n < d) {
            for (int k = 0; k < d; ++k) {
                pm = fmaf(wbt[(int64_t)k * NP + n] + wbt[(int64_t)n * NP + k], u[k], pm);
                ww = fmaf(wwt[(int64_t)k * NP + n], u[k], ww);
            }
        }
        float part = n < d ? u[n] * (ww + ws[n]) : 0.f;          // u_n (Ww u)_n + ws_n u_n
        part = warp_sum(part);
        if ((n & 31) == 0) red[n >> 5] = part;
        if (n < ROW_LD) B[n] = n < d ? pm : 0.f;
        if (n >= d && n < ROW_LD - 1) A[n] = 0.f;
        __syncthreads();
        if (n == 0) {
            float tot = 0.5f * c[0];
            for (int w = 0; w < NP / 32; ++w) tot += red[w];
            A[ROW_LD - 1] = tot;
        }
    }
}

__global__ void rowtab_commit_kernel(const unsigned long long *fp_cur, unsigned long long *fp_built) { *fp_built = *fp_cur; }

// one warp per trial: S = r[i] + r[j] + A[i] . B[j]; 44 float4 per half row = lanes 0..31 + lanes 0..11.
// A warp takes BLOCKS of 16 consecutive trials: trial lists are written enrol-major (the reference's key files, the
// grids of BASELINE.json), so consecutive trials mostly share idx1 and the A row (and r[i]) stay in registers --
// the row-gather traffic from L2, which bounds this kernel, drops from 1408 to ~704 bytes per trial.
constexpr int PAIR_BLOCK = 16;

__global__ void __launch_bounds__(256) score_pairs_kernel(const float *__restrict__ rowtab, int64_t n_rows,
                                                          const int64_t *__restrict__ i1, const int64_t *__restrict__ i2,
                                                          int64_t n, float *__restrict__ scores, int32_t *bad_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nblocks = (n + PAIR_BLOCK - 1) / PAIR_BLOCK;
    for (int64_t blk = w0; blk < nblocks; blk += nw) {
        const int64_t t0 = blk * PAIR_BLOCK, t1 = min(n, t0 + PAIR_BLOCK);
        int64_t cur = -1;                                            // row whose A half is in registers
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f), x2 = x;
        float ra = 0.f;
        // the block's indices: one coalesced load per list, handed round by shuffles
        const int64_t ia = (t0 + lane < t1) ? i1[t0 + lane] : 0, ib = (t0 + lane < t1) ? i2[t0 + lane] : 0;
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t a = __shfl_sync(0xffffffffu, ia, (int)(t - t0)), b = __shfl_sync(0xffffffffu, ib, (int)(t - t0));
            if (a < 0 || a >= n_rows || b < 0 || b >= n_rows) {       // reported, never a fault
                if (lane == 0) { *bad_flag = 1; scores[t] = 0.f; }
                continue;
            }
            if (a != cur) {
                const float4 *A = reinterpret_cast<const float4 *>(rowtab + a * ROW_FLOATS);
                x = A[lane];
                x2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (lane < ROW_LD / 4 - 32) {
                    x2 = A[32 + lane];
                    if (lane == ROW_LD / 4 - 33) { ra = x2.w; x2.w = 0.f; }   // A[ROW_LD - 1] is r, not part of the dot product
                }
                ra = __shfl_sync(0xffffffffu, ra, ROW_LD / 4 - 33);
                cur = a;
            }
            const float4 *B = reinterpret_cast<const float4 *>(rowtab + b * ROW_FLOATS + ROW_LD);
            const float4 y = B[lane];
            float acc = x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
            if (lane < ROW_LD / 4 - 32) {
                const float4 y2 = B[32 + lane];
                acc += x2.x * y2.x + x2.y * y2.y + x2.z * y2.z + x2.w * y2.w;
            }
            acc = warp_sum(acc);
            if (lane == 0) scores[t] = acc + ra + rowtab[b * ROW_FLOATS + ROW_LD - 1];
        }
    }
}

// ---- dense trial lists: sub-grid + gather ------------------------------------------------------------------------
// A trial list that covers a sizeable fraction of (its enrol rows) x (its test rows) -- key files are written enrol-major,
// every model against the same segments -- is cheaper as ONE grid product over those rows on the tensor cores
// (grid_tc.cu: ~2.7 ps per grid entry) followed by a 4-byte gather per trial than as one 704-byte row gather per trial:
//   1. trial_rows_mark:    flags[0][i1[t]] = flags[1][i2[t]] = 1              (one pass over the index lists)
//   2. trial_rows_compact: pos[s][row] = rank of the row among the marked ones, list[s][rank] = row, counts[s]
//   3. nplda_score_grid over list[0] x list[1]   (the caller reads the two counts to size the grid)
//   4. trial_grid_gather:  scores[t] = grid[pos[0][i1[t]]][pos[1][i2[t]]]
__global__ void __launch_bounds__(256) trial_rows_mark_kernel(const int64_t *__restrict__ i1, const int64_t *__restrict__ i2, int64_t n,
                                                              int64_t n_rows, int32_t *__restrict__ flags, int32_t *bad_flag) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = i1[t], b = i2[t];
        if (a < 0 || a >= n_rows || b < 0 || b >= n_rows) { *bad_flag = 1; continue; }
        if (flags[a] == 0) flags[a] = 1;                       // benign race: every writer stores 1
        if (flags[n_rows + b] == 0) flags[n_rows + b] = 1;
    }
}

// one CTA per side: every thread owns a contiguous segment of rows
__global__ void __launch_bounds__(1024) trial_rows_compact_kernel(const int32_t *__restrict__ flags, int64_t n_rows,
                                                                  int32_t *__restrict__ pos, int64_t *__restrict__ list,
                                                                  int32_t *__restrict__ counts) {
    __shared__ int32_t part[1024];
    const int side = blockIdx.x, tid = threadIdx.x;
    const int32_t *f = flags + side * n_rows;
    const int64_t seg = (n_rows + 1023) / 1024, r0 = min(n_rows, tid * seg), r1 = min(n_rows, r0 + seg);
    int32_t c = 0;
    for (int64_t r = r0; r < r1; ++r) c += f[r] != 0;
    part[tid] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                       // inclusive scan
        const int32_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int32_t k = part[tid] - c;
    for (int64_t r = r0; r < r1; ++r) {
        if (f[r] != 0) { pos[side * n_rows + r] = k; list[side * n_rows + k] = r; ++k; }
        else pos[side * n_rows + r] = -1;
    }
    if (tid == 1023) counts[side] = part[1023];
}

__global__ void __launch_bounds__(256) trial_grid_gather_kernel(const float *__restrict__ grid, int64_t ld, const int32_t *__restrict__ pos,
                                                                int64_t n_rows, const int64_t *__restrict__ i1,
                                                                const int64_t *__restrict__ i2, int64_t n, float *__restrict__ scores) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = i1[t], b = i2[t];
        float v = 0.f;                                          // rows outside the table: reported by the mark pass
        if (a >= 0 && a < n_rows && b >= 0 && b < n_rows) v = grid[(int64_t)pos[a] * ld + pos[n_rows + b]];
        scores[t] = v;
    }
}

}  // namespace nplda

using namespace nplda;

extern "C" int nplda_trial_rows(const int64_t *idx1, const int64_t *idx2, int64_t n, int64_t n_rows, int32_t *flags, int32_t *pos,
                                int64_t *list, int32_t *counts, int32_t *bad_index_flag, void *stream) {
    if (n < 0 || n_rows <= 0 || n_rows >= ((int64_t)1 << 31) || !flags || !pos || !list || !counts || !bad_index_flag ||
        (n > 0 && (!idx1 || !idx2)))
        return NPLDA_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    NPLDA_CUDA_TRY(cudaMemsetAsync(flags, 0, 2 * n_rows * sizeof(int32_t), st));
    if (n > 0) {
        const int grid = (int)std::min<int64_t>((n + 255) / 256, 16 * (int64_t)sm_count());
        trial_rows_mark_kernel<<<grid, 256, 0, st>>>(idx1, idx2, n, n_rows, flags, bad_index_flag);
        NPLDA_LAUNCH_CHECK();
    }
    trial_rows_compact_kernel<<<2, 1024, 0, st>>>(flags, n_rows, pos, list, counts);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_trial_grid_gather(const float *grid, int64_t ld, const int32_t *pos, int64_t n_rows, const int64_t *idx1,
                                       const int64_t *idx2, int64_t n, float *scores, void *stream) {
    if (n < 0 || n_rows <= 0 || (n > 0 && (!grid || !pos || !idx1 || !idx2 || !scores))) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    const int g = (int)std::min<int64_t>((n + 255) / 256, 16 * (int64_t)sm_count());
    trial_grid_gather_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(grid, ld, pos, n_rows, idx1, idx2, n, scores);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int64_t nplda_rowtab_bytes(int64_t n_rows) {
    if (n_rows < 0) return NPLDA_ERR_BAD_ARG;
    return rowtab_gtab_offset(n_rows) + gtab_bytes(n_rows);        // rows + trailer (fingerprint, grid header) + grid operands
}

extern "C" int nplda_table_prepare(const float *table, int64_t n_rows, int d_in, int d1, int d2, const void *pack,
                                   int is_dplda, float *rowtab, int flags, void *stream) {
    if (n_rows < 0 || !pack || (n_rows > 0 && (!table || !rowtab))) return NPLDA_ERR_BAD_ARG;
    if (is_dplda) d2 = d1;
    if (!dims_supported(d_in, d1, d2) || d1 >= ROW_LD || d2 >= ROW_LD) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n_rows == 0) return NPLDA_OK;
    PackLayout L = make_pack_layout(d_in, d1, d2);
    cudaStream_t st = (cudaStream_t)stream;
    const char *pk = (const char *)pack;
    const unsigned long long *fp_cur = (const unsigned long long *)(pk + L.fp) + ((flags >> 1) & 1);
    unsigned long long *fp_built = (unsigned long long *)(rowtab + n_rows * ROW_FLOATS);
    const unsigned long long *g_cur = (flags & NPLDA_PREPARE_IF_CHANGED) ? fp_cur : nullptr;
    int rc = simt_aux(is_dplda ? 3 : 2, table, nullptr, n_rows, L, pk, rowtab, ROW_FLOATS, st, g_cur, fp_built);
    if (rc != NPLDA_OK) return rc;
    if (is_dplda) {
        const int grid = (int)std::min<int64_t>(n_rows, 8 * (int64_t)sm_count());
        rowtab_dplda_kernel<<<grid, NP, 0, st>>>(rowtab, n_rows, d1, (const float *)(pk + L.w2t), (const float *)(pk + L.w3t),
                                                 (const float *)(pk + L.b2), (const float *)(pk + L.c), g_cur, fp_built);
    } else {
        const int grid = (int)std::min<int64_t>((n_rows + 7) / 8, 8 * (int64_t)sm_count());
        rowtab_nplda_kernel<<<grid, 256, 0, st>>>(rowtab, n_rows, d2, (const float *)(pk + L.p), (const float *)(pk + L.q),
                                                 g_cur, fp_built);
    }
    NPLDA_LAUNCH_CHECK();
    // fp16 hi/lo operands of the tensor-core grid kernel (grid_tc.cu), rebuilt together with the rows
    rc = gtab_prepare(rowtab, n_rows, ROW_FLOATS, ROW_LD, is_dplda ? d1 : d2, (float *)((char *)rowtab + rowtab_trailer_offset(n_rows) + 64),
                      (uint8_t *)rowtab + rowtab_gtab_offset(n_rows), g_cur, fp_built, st);
    if (rc != NPLDA_OK) return rc;
    rowtab_commit_kernel<<<1, 1, 0, st>>>(fp_cur, fp_built);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_score_pairs(const float *rowtab, int64_t n_rows, const int64_t *idx1, const int64_t *idx2, int64_t n,
                                 float *scores, int32_t *bad_index_flag, void *stream) {
    if (n < 0 || n_rows < 0 || (n > 0 && (!rowtab || !idx1 || !idx2 || !scores || !bad_index_flag))) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    if (n_rows == 0) return NPLDA_ERR_BAD_ARG;
    const int grid = (int)std::min<int64_t>(((n + PAIR_BLOCK - 1) / PAIR_BLOCK + 7) / 8, 16 * (int64_t)sm_count());
    score_pairs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rowtab, n_rows, idx1, idx2, n, scores, bad_index_flag);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}
