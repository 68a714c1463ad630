// Batch gather from the device-resident x-vector table: the device half of
// load_xvec_trials_from_numbatch / _from_idbatch (sv_trials_loaders.py:418-437), which the reference does with a
// Python loop over dict lookups + np.asarray + a host->device copy of 4 KB per trial every batch.
// One warp per output row, 16-byte loads/stores; both sides of the batch in one launch.  Rows outside the table
// are reported through *bad_index_flag (which may live in pinned host memory, so that the host can look at it
// without a synchronising copy) and come back filled with NaN -- never a fault, and never a silently wrong score:
// whatever is computed from such a row (scores, losses, gradients) is NaN until the host has seen the flag.
#include <algorithm>

#include "common.cuh"

namespace nplda {

template <bool VEC>
__global__ void __launch_bounds__(256) gather_pairs_kernel(const float *__restrict__ table, int64_t n_rows, int d,
                                                           const int64_t *__restrict__ i1, const int64_t *__restrict__ i2,
                                                           int64_t n, float *__restrict__ x1, float *__restrict__ x2,
                                                           int32_t *bad_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = w0; r < 2 * n; r += nw) {
        const bool second = r >= n;
        const int64_t t = second ? r - n : r;
        const int64_t row = second ? i2[t] : i1[t];
        float *dst = (second ? x2 : x1) + t * d;
        const bool bad = row < 0 || row >= n_rows;
        if (bad && lane == 0) *bad_flag = 1;
        const float *src = table + (bad ? 0 : row) * (int64_t)d;
        if (VEC) {
            for (int k = lane; k < d / 4; k += 32) {
                const float qn = __int_as_float(0x7fc00000);
                const float4 v = bad ? make_float4(qn, qn, qn, qn) : reinterpret_cast<const float4 *>(src)[k];
                reinterpret_cast<float4 *>(dst)[k] = v;
            }
        } else {
            for (int k = lane; k < d; k += 32) dst[k] = bad ? __int_as_float(0x7fc00000) : src[k];
        }
    }
}

}  // namespace nplda

using namespace nplda;

extern "C" int nplda_gather_pairs(const float *table, int64_t n_rows, int d, const int64_t *idx1, const int64_t *idx2,
                                  int64_t n, float *x1, float *x2, int32_t *bad_index_flag, void *stream) {
    if (n < 0 || n_rows < 0 || d <= 0 || (n > 0 && (!table || !idx1 || !idx2 || !x1 || !x2 || !bad_index_flag))) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    if (n_rows == 0) return NPLDA_ERR_BAD_ARG;
    const int grid = (int)std::min<int64_t>((2 * n + 7) / 8, 16 * (int64_t)sm_count());
    const bool vec = d % 4 == 0 && (((uintptr_t)table | (uintptr_t)x1 | (uintptr_t)x2) & 15) == 0;
    if (vec) gather_pairs_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(table, n_rows, d, idx1, idx2, n, x1, x2, bad_index_flag);
    else gather_pairs_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(table, n_rows, d, idx1, idx2, n, x1, x2, bad_index_flag);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}
