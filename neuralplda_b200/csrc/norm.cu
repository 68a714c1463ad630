// Cohort score normalisation: the statistics and the per-trial arithmetic of the
// reference's utils/adaptive_score_normalization.py on the device.
//   :32-37  per enrol/test id: cohort scores sorted ascending, mean / std over all, mean / std over the FIRST
//           ASnorm_topN of the ascending sort (i.e. the N LOWEST scores -- kept as the reference computes it)
//   :62-66  z = (s - mean[e]) / std[e], t = (s - mean[t]) / std[t], snorm = (z + t) / 2,
//           asnorm1 = ((s - mean_top[e]) / std_top[e] + (s - mean_top[t]) / std_top[t]) / 2       (all float64)
// The cohort score matrix itself is an id x cohort grid scored by nplda_score_pairs.
// Row statistics: one CTA per row.  The N smallest values are found without sorting: a 4-pass radix select on
// the order-preserving integer image of the floats gives the N-th smallest value v*; the selected multiset is
// {x < v*} plus (N - #{x < v*}) copies of v*, exactly sort(row)[:N].  Sums are two-pass (mean, then centred
// squares) in fp64 with a fixed reduction tree, so results are deterministic.
#include <algorithm>

#include "common.cuh"

namespace nplda {

__device__ __forceinline__ uint32_t f32_key(float f) {          // monotone: a < b  <=>  key(a) < key(b)
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ double block_sum(double v, double *red) {            // all threads get the sum; fixed tree
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    return tot;
}

__global__ void __launch_bounds__(256) cohort_stats_kernel(const float *__restrict__ scores, int64_t m, int64_t c, int64_t top_n,
                                                           double *__restrict__ stats) {
    __shared__ unsigned int hist[256];
    __shared__ double red[8];
    __shared__ uint32_t sel_prefix;
    __shared__ unsigned long long sel_below;
    for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
        const float *x = scores + row * c;
        // ---- all scores: mean and population std (np.mean / np.std) ----
        double s = 0.0;
        for (int64_t i = threadIdx.x; i < c; i += blockDim.x) s += (double)x[i];
        const double mean = block_sum(s, red) / (double)c;
        double q = 0.0;
        for (int64_t i = threadIdx.x; i < c; i += blockDim.x) { const double d = (double)x[i] - mean; q += d * d; }
        const double var = block_sum(q, red) / (double)c;
        // ---- the N lowest scores ----
        const int64_t n = top_n < c ? top_n : c;
        double mean_top = mean, var_top = var;
        if (n < c) {
            // radix select of the n-th smallest key, most significant byte first
            uint32_t prefix = 0;
            unsigned long long below = 0;                       // elements with key < current prefix range
            for (int shift = 24; shift >= 0; shift -= 8) {
                hist[threadIdx.x] = 0;
                __syncthreads();
                const uint32_t mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
                for (int64_t i = threadIdx.x; i < c; i += blockDim.x) {
                    const uint32_t k = f32_key(x[i]);
                    if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    unsigned long long acc = below;
                    int b = 0;
                    for (; b < 256; ++b) {
                        if (acc + hist[b] >= (unsigned long long)n) break;
                        acc += hist[b];
                    }
                    sel_prefix = prefix | ((uint32_t)b << shift);
                    sel_below = acc;
                }
                __syncthreads();
                prefix = sel_prefix;
                below = sel_below;
                __syncthreads();
            }
            // prefix = key of the n-th smallest value v*, below = #{key < prefix}
            const double copies = (double)((unsigned long long)n - below);
            float vstar = 0.f;
            {
                const uint32_t u = (prefix & 0x80000000u) ? (prefix & 0x7FFFFFFFu) : ~prefix;
                vstar = __uint_as_float(u);
            }
            double st = 0.0;
            for (int64_t i = threadIdx.x; i < c; i += blockDim.x)
                if (f32_key(x[i]) < prefix) st += (double)x[i];
            mean_top = (block_sum(st, red) + copies * (double)vstar) / (double)n;
            double qt = 0.0;
            for (int64_t i = threadIdx.x; i < c; i += blockDim.x)
                if (f32_key(x[i]) < prefix) { const double d = (double)x[i] - mean_top; qt += d * d; }
            const double dv = (double)vstar - mean_top;
            var_top = (block_sum(qt, red) + copies * dv * dv) / (double)n;
        }
        if (threadIdx.x == 0) {
            stats[row * 4 + 0] = mean;
            stats[row * 4 + 1] = sqrt(var);
            stats[row * 4 + 2] = mean_top;
            stats[row * 4 + 3] = sqrt(var_top);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) score_norm_kernel(const float *__restrict__ raw, const int64_t *__restrict__ er,
                                                         const int64_t *__restrict__ tr, int64_t n, const double *__restrict__ stats,
                                                         int64_t m, double *__restrict__ out, int32_t *bad_flag) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = er[i], t = tr[i];
        if (e < 0 || e >= m || t < 0 || t >= m) {
            *bad_flag = 1;
            out[i] = out[n + i] = out[2 * n + i] = out[3 * n + i] = 0.0;
            continue;
        }
        const double s = (double)raw[i];
        const double z = (s - stats[e * 4 + 0]) / stats[e * 4 + 1];
        const double tt = (s - stats[t * 4 + 0]) / stats[t * 4 + 1];
        out[i] = z;
        out[n + i] = tt;
        out[2 * n + i] = (z + tt) / 2;
        out[3 * n + i] = ((s - stats[e * 4 + 2]) / stats[e * 4 + 3] + (s - stats[t * 4 + 2]) / stats[t * 4 + 3]) / 2;
    }
}

}  // namespace nplda

using namespace nplda;

extern "C" int nplda_cohort_stats(const float *scores, int64_t m, int64_t c, int64_t top_n, double *stats, void *stream) {
    if (m < 0 || c <= 0 || top_n <= 0 || (m > 0 && (!scores || !stats))) return NPLDA_ERR_BAD_ARG;
    if (m == 0) return NPLDA_OK;
    const int grid = (int)std::min<int64_t>(m, 16 * (int64_t)sm_count());
    cohort_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scores, m, c, top_n, stats);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_score_norm(const float *raw, const int64_t *enrol_row, const int64_t *test_row, int64_t n,
                                const double *stats, int64_t m, double *out, int32_t *bad_index_flag, void *stream) {
    if (n < 0 || m < 0 || (n > 0 && (!raw || !enrol_row || !test_row || !stats || !out || !bad_index_flag))) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 8 * (int64_t)sm_count());
    score_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(raw, enrol_row, test_row, n, stats, m, out, bad_index_flag);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}
