// Score populations for the minC threshold sweep (models.py:406-409: `torch.sort(output[target > 0.5])`,
// `torch.sort(output[target < 0.5])`): a label split and an ascending sort, hand-written.
//
//   nplda_split_by_label   scores with label > 0.5 -> tgt, label < 0.5 -> non (a label of exactly 0.5 belongs to neither,
//                          as in the reference's two masks); warp-aggregated atomic counters, order irrelevant (sorted next)
//   nplda_sort_f32         in-place ascending bitonic sort in the "flip + disperse" form: every compare-exchange puts the
//                          minimum at the lower index, so a length that is not a power of two is handled by treating the
//                          missing tail as +inf (those comparators are no-ops) instead of padding the buffer.  Steps up to
//                          2048 elements run fused in shared memory, the larger ones one launch each: 54 launches for 2^20
//                          keys.  The sweep is not on the trial-pairs/s path; what matters is that the sorted multiset is
//                          exactly torch.sort's.  NaN keys are placed last (like torch.sort).
#include <algorithm>

#include "common.cuh"

namespace nplda {
namespace srt {

constexpr int TILE = 2048, THREADS = 1024;

// total order with NaN last: a "less than or equal" that never reports a NaN as smaller
__device__ __forceinline__ bool out_of_order(float lo, float hi) { return (lo > hi) || (lo != lo && hi == hi); }

__global__ void __launch_bounds__(256) split_kernel(const float *__restrict__ s, const float *__restrict__ t, int64_t n,
                                                    float *__restrict__ tgt, float *__restrict__ non,
                                                    unsigned long long *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;                      // whole warps take part in the ballots
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < n;
        const float v = in ? s[i] : 0.f, lab = in ? t[i] : 0.5f;
        const bool is_t = lab > 0.5f, is_n = lab < 0.5f;
        const unsigned mt = __ballot_sync(0xffffffffu, is_t), mn = __ballot_sync(0xffffffffu, is_n);
        unsigned long long bt = 0, bn = 0;
        if (lane == 0) {
            if (mt) bt = atomicAdd(counts, (unsigned long long)__popc(mt));
            if (mn) bn = atomicAdd(counts + 1, (unsigned long long)__popc(mn));
        }
        bt = __shfl_sync(0xffffffffu, bt, 0);
        bn = __shfl_sync(0xffffffffu, bn, 0);
        const unsigned below = (1u << lane) - 1u;
        if (is_t) tgt[bt + __popc(mt & below)] = v;
        if (is_n) non[bn + __popc(mn & below)] = v;
    }
}

// all steps with block height <= TILE of the merges [h_first, h_last] on one tile in shared memory;
// flip_first: the first step of height h_first is a flip (start of a merge), otherwise a disperse
__global__ void __launch_bounds__(THREADS) tile_steps_kernel(float *__restrict__ a, int64_t n, int h_first, int h_last,
                                                             int first_is_local_sort) {
    __shared__ float sh[TILE];
    const int64_t base = (int64_t)blockIdx.x * TILE;
    const float inf = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < TILE; i += THREADS) sh[i] = base + i < n ? a[base + i] : inf;
    __syncthreads();
    auto ce = [&](int i, int l) {
        const float x = sh[i], y = sh[l];
        if (out_of_order(x, y)) { sh[i] = y; sh[l] = x; }
    };
    if (first_is_local_sort) {
        for (int h = 2; h <= h_last; h <<= 1) {
            for (int t = threadIdx.x; t < TILE / 2; t += THREADS) {          // flip(h)
                const int blk = t / (h / 2), off = t % (h / 2);
                ce(blk * h + off, blk * h + h - 1 - off);
            }
            __syncthreads();
            for (int hh = h >> 1; hh >= 2; hh >>= 1) {                       // disperse(hh)
                for (int t = threadIdx.x; t < TILE / 2; t += THREADS) {
                    const int blk = t / (hh / 2), off = t % (hh / 2);
                    ce(blk * hh + off, blk * hh + off + hh / 2);
                }
                __syncthreads();
            }
        }
    } else {
        for (int hh = h_first; hh >= 2; hh >>= 1) {                          // tail of a large merge: disperse(TILE .. 2)
            for (int t = threadIdx.x; t < TILE / 2; t += THREADS) {
                const int blk = t / (hh / 2), off = t % (hh / 2);
                ce(blk * hh + off, blk * hh + off + hh / 2);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < TILE; i += THREADS)
        if (base + i < n) a[base + i] = sh[i];
}

// one compare-exchange per thread: flip (partner = mirror inside the block of height h) or disperse (partner h / 2 above)
__global__ void __launch_bounds__(256) global_step_kernel(float *__restrict__ a, int64_t n, int64_t npairs, int64_t h, int flip) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= npairs) return;
    const int64_t half = h >> 1, blk = t / half, off = t % half;
    const int64_t i = blk * h + off, l = flip ? blk * h + h - 1 - off : i + half;
    if (l >= n) return;                                                       // partner is the virtual +inf tail
    const float x = a[i], y = a[l];
    if (out_of_order(x, y)) { a[i] = y; a[l] = x; }
}

}  // namespace srt
}  // namespace nplda

using namespace nplda;

extern "C" int nplda_split_by_label(const float *scores, const float *labels, int64_t n, float *tgt, float *non,
                                    unsigned long long *counts, void *stream) {
    if (n < 0 || !counts || (n > 0 && (!scores || !labels || !tgt || !non))) return NPLDA_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    NPLDA_CUDA_TRY(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), st));
    if (n == 0) return NPLDA_OK;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 8 * (int64_t)sm_count());
    srt::split_kernel<<<grid, 256, 0, st>>>(scores, labels, n, tgt, non, counts);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_sort_f32(float *keys, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && !keys)) return NPLDA_ERR_BAD_ARG;
    if (n < 2) return NPLDA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t N = 1;
    while (N < n) N <<= 1;
    const int tiles = (int)((n + srt::TILE - 1) / srt::TILE);
    srt::tile_steps_kernel<<<tiles, srt::THREADS, 0, st>>>(keys, n, 2, (int)std::min<int64_t>(N, srt::TILE), 1);
    NPLDA_LAUNCH_CHECK();
    const int64_t npairs = N / 2;
    const unsigned gblocks = (unsigned)((npairs + 255) / 256);
    for (int64_t h = 2 * (int64_t)srt::TILE; h <= N; h <<= 1) {
        srt::global_step_kernel<<<gblocks, 256, 0, st>>>(keys, n, npairs, h, 1);
        NPLDA_LAUNCH_CHECK();
        for (int64_t hh = h >> 1; hh > srt::TILE; hh >>= 1) {
            srt::global_step_kernel<<<gblocks, 256, 0, st>>>(keys, n, npairs, hh, 0);
            NPLDA_LAUNCH_CHECK();
        }
        srt::tile_steps_kernel<<<tiles, srt::THREADS, 0, st>>>(keys, n, srt::TILE, srt::TILE, 0);
        NPLDA_LAUNCH_CHECK();
    }
    return NPLDA_OK;
}
