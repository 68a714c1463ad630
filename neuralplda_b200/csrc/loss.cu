// K2: detection-cost / cross-entropy accumulators, their finalisation, the
// loss backward (dL/ds, dL/dthreshold) and the minC threshold sweep.
// Reference: utils/models.py softcdet 384-388, crossentropy 390-393 (DPlda
// 503-506), cdet 401-404, minc 406-421 + arr2val 23-27.
#include "common.cuh"

namespace nplda {

constexpr int LOSS_THREADS = 256;
constexpr int MAXK = NPLDA_MAX_BETAS;

struct Thr {
    float v[MAXK];
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

// PyTorch binary_cross_entropy on p = sigmoid(z): logs clamped at -100.
__device__ __forceinline__ float bce_of_logit(float z, float t) {
    float p = sigmoidf_(z);
    float lp = fmaxf(logf(p), -100.f);
    float l1p = fmaxf(logf(1.f - p), -100.f);
    return -(t * lp + (1.f - t) * l1p);
}

__device__ __forceinline__ void block_add(double v, double *dst, double *sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double r = lane < LOSS_THREADS / 32 ? sh[lane] : 0.0;
        r = warp_sum(r);
        if (lane == 0 && r != 0.0) atomicAdd(dst, r);
    }
}

template <int K>
__global__ void __launch_bounds__(LOSS_THREADS) loss_accum_kernel(
    const float *__restrict__ s, const float *__restrict__ t, int64_t n,
    const float *__restrict__ thresholds, float alpha, const float *__restrict__ th_xent,
    double *__restrict__ acc, int kdyn) {
    __shared__ double sh[LOSS_THREADS / 32];
    const int KK = K > 0 ? K : kdyn;
    float th[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) th[k] = k < KK ? thresholds[k] : 0.f;
    const float thx = th_xent ? th_xent[0] : 0.f;

    // fp32 partials over a short per-thread run, flushed into fp64
    double A[MAXK], B[MAXK], M[MAXK], F[MAXK], nt = 0, nn = 0, bce = 0;
#pragma unroll
    for (int k = 0; k < MAXK; ++k) A[k] = B[k] = M[k] = F[k] = 0.0;

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float si = s[i], ti = t[i], ui = 1.f - ti;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < KK) {
                A[k] += (double)(ti * sigmoidf_(alpha * (th[k] - si)));
                B[k] += (double)(ui * sigmoidf_(alpha * (si - th[k])));
                M[k] += (double)(si < th[k] ? ti : 0.f);
                F[k] += (double)(si > th[k] ? ui : 0.f);
            }
        }
        nt += (double)ti;
        nn += (double)ui;
        bce += (double)bce_of_logit(si - thx, ti);
    }
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        if (k < KK) {
            block_add(A[k], acc + 4 * k + 0, sh);
            block_add(B[k], acc + 4 * k + 1, sh);
            block_add(M[k], acc + 4 * k + 2, sh);
            block_add(F[k], acc + 4 * k + 3, sh);
        }
    }
    block_add(nt, acc + 4 * KK + 0, sh);
    block_add(nn, acc + 4 * KK + 1, sh);
    block_add(bce, acc + 4 * KK + 2, sh);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(acc + 4 * KK + 3, (double)n);
}

struct Betas {
    double v[MAXK];
};

__global__ void loss_finalize_kernel(const double *__restrict__ acc, Betas betas, int K,
                                     float *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double nt = acc[4 * K + 0], nn = acc[4 * K + 1], sb = acc[4 * K + 2], n = acc[4 * K + 3];
    double soft = 0, hard = 0;
    for (int k = 0; k < K; ++k) {
        soft += acc[4 * k + 0] / nt + betas.v[k] * acc[4 * k + 1] / nn;
        hard += acc[4 * k + 2] / nt + betas.v[k] * acc[4 * k + 3] / nn;
    }
    out[0] = K > 0 ? (float)(soft / K) : 0.f;
    out[1] = (float)(sb / n);
    out[2] = K > 0 ? (float)(hard / K) : 0.f;
    out[3] = 0.f;
}

__global__ void __launch_bounds__(LOSS_THREADS) loss_bwd_kernel(
    const float *__restrict__ s, const float *__restrict__ t, int64_t n,
    const float *__restrict__ thresholds, Betas betas, int K, float alpha,
    const float *__restrict__ th_xent, const double *__restrict__ acc, int loss_id,
    const float *__restrict__ grad_out, float *__restrict__ ds, double *__restrict__ dth) {
    __shared__ double sh[LOSS_THREADS / 32];
    const float g = grad_out ? grad_out[0] : 1.f;
    const double nt = acc[4 * K + 0], nn = acc[4 * K + 1];
    float th[MAXK], wt[MAXK], wn[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        th[k] = k < K ? thresholds[k] : 0.f;
        // d/ds of  A_k/nt + beta_k B_k/nn  =  alpha * sig' * ( -t/nt + beta_k (1-t)/nn ), mean over K
        wt[k] = k < K ? (float)(-(double)alpha / (K * nt)) : 0.f;
        wn[k] = k < K ? (float)((double)alpha * betas.v[k] / (K * nn)) : 0.f;
    }
    const float thx = th_xent ? th_xent[0] : 0.f;
    // mean over ALL trials of the loss: acc[4K+3] is the trial count the accumulators were built from -- the global one
    // after the all-reduce when the trial list is sharded across ranks (n is only this rank's share)
    const double n_all = acc[4 * K + 3];
    const float inv_n = (float)(1.0 / (n_all > 0.0 ? n_all : (double)n));
    double dk[MAXK], dx = 0.0;
#pragma unroll
    for (int k = 0; k < MAXK; ++k) dk[k] = 0.0;

    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float si = s[i], ti = t[i];
        float d = 0.f;
        if (loss_id == NPLDA_LOSS_SOFTCDET) {
#pragma unroll
            for (int k = 0; k < MAXK; ++k) {
                if (k < K) {
                    float sg = sigmoidf_(alpha * (si - th[k]));
                    float dk_i = sg * (1.f - sg) * (ti * wt[k] + (1.f - ti) * wn[k]);
                    d += dk_i;
                    dk[k] -= (double)dk_i;
                }
            }
        } else {
            // binary_cross_entropy backward: (p - t) / max(p(1-p), 1e-12) / N, then sigmoid': p(1-p)
            float p = sigmoidf_(si - thx);
            float pq = p * (1.f - p);
            d = (p - ti) / fmaxf(pq, 1e-12f) * pq * inv_n;
            dx -= (double)d;
        }
        ds[i] = g * d;
    }
    if (dth) {
        if (loss_id == NPLDA_LOSS_SOFTCDET) {
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
                if (k < K) block_add(dk[k] * (double)g, dth + k, sh);
        } else {
            block_add(dx * (double)g, dth + K, sh);
        }
    }
}

// ---- minC sweep ----------------------------------------------------------------
__device__ __forceinline__ int64_t lower_bound_f(const float *a, int64_t n, float v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct BetasF {
    float v[MAXK];
};

// One block-level (min, first index) per beta; blocks combine through a packed
// 64-bit atomicMin on (orderable float bits << 32 | ... ) is not enough for an
// int64 index, so: pass 1 writes per-block minima, pass 2 (one block) reduces.
__global__ void __launch_bounds__(256) minc_sweep_kernel(
    const float *__restrict__ tg, int64_t n_t, const float *__restrict__ nt_sorted, int64_t n_n,
    float sum_t, float sum_n, BetasF betas, int K, float *__restrict__ blk_min,
    int64_t *__restrict__ blk_arg) {
    __shared__ float smin[256];
    __shared__ int64_t sarg[256];
    float best[MAXK];
    int64_t arg[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) { best[k] = INFINITY; arg[k] = INT64_MAX; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_t; j += stride) {
        const float v = tg[j];
        int64_t below = lower_bound_f(tg, n_t, v);                 // targets strictly below
        int64_t ge = n_n - lower_bound_f(nt_sorted, n_n, v);       // non-targets >= v
        float pm = (below > 0 ? (float)(below - 1) : 1.f) / sum_t;  // arr2val quirk (models.py:23-27)
        float pf = (ge > 0 ? (float)(ge - 1) : 1.f) / sum_n;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            if (k < K) {
                float c = __fadd_rn(pm, __fmul_rn(betas.v[k], pf));   // no FMA contraction: match torch
                if (c < best[k] || (c == best[k] && j < arg[k])) { best[k] = c; arg[k] = j; }
            }
        }
    }
    for (int k = 0; k < K; ++k) {
        smin[threadIdx.x] = best[k];
        sarg[threadIdx.x] = arg[k];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                float c = smin[threadIdx.x + o];
                int64_t a = sarg[threadIdx.x + o];
                if (c < smin[threadIdx.x] || (c == smin[threadIdx.x] && a < sarg[threadIdx.x])) {
                    smin[threadIdx.x] = c;
                    sarg[threadIdx.x] = a;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            blk_min[(int64_t)k * gridDim.x + blockIdx.x] = smin[0];
            blk_arg[(int64_t)k * gridDim.x + blockIdx.x] = sarg[0];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) minc_final_kernel(const float *__restrict__ blk_min,
                                                         const int64_t *__restrict__ blk_arg,
                                                         int nblk, int K, float *__restrict__ out_min,
                                                         int64_t *__restrict__ out_arg) {
    __shared__ float smin[256];
    __shared__ int64_t sarg[256];
    for (int k = 0; k < K; ++k) {
        float best = INFINITY;
        int64_t arg = INT64_MAX;
        for (int b = threadIdx.x; b < nblk; b += 256) {
            float c = blk_min[(int64_t)k * nblk + b];
            int64_t a = blk_arg[(int64_t)k * nblk + b];
            if (c < best || (c == best && a < arg)) { best = c; arg = a; }
        }
        smin[threadIdx.x] = best;
        sarg[threadIdx.x] = arg;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                float c = smin[threadIdx.x + o];
                int64_t a = sarg[threadIdx.x + o];
                if (c < smin[threadIdx.x] || (c == smin[threadIdx.x] && a < sarg[threadIdx.x])) {
                    smin[threadIdx.x] = c;
                    sarg[threadIdx.x] = a;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) { out_min[k] = smin[0]; out_arg[k] = sarg[0]; }
        __syncthreads();
    }
}

}  // namespace nplda

using namespace nplda;

static int loss_grid(int64_t n) {
    int64_t blocks = (n + LOSS_THREADS - 1) / LOSS_THREADS;
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, 4 * (int64_t)sm_count()));
}

extern "C" int nplda_loss_accum(const float *scores, const float *labels, int64_t n,
                                const float *thresholds, int K, float alpha, const float *th_xent,
                                double *acc, void *stream) {
    if (n < 0 || K < 0 || K > MAXK || !acc || (n > 0 && (!scores || !labels)) || (K > 0 && !thresholds))
        return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = loss_grid(n);
    if (K == 1)
        loss_accum_kernel<1><<<grid, LOSS_THREADS, 0, st>>>(scores, labels, n, thresholds, alpha, th_xent, acc, K);
    else if (K == 2)
        loss_accum_kernel<2><<<grid, LOSS_THREADS, 0, st>>>(scores, labels, n, thresholds, alpha, th_xent, acc, K);
    else
        loss_accum_kernel<0><<<grid, LOSS_THREADS, 0, st>>>(scores, labels, n, thresholds, alpha, th_xent, acc, K);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_loss_finalize(const double *acc, const double *betas_host, int K, float *out,
                                   void *stream) {
    if (!acc || !out || K < 0 || K > MAXK || (K > 0 && !betas_host)) return NPLDA_ERR_BAD_ARG;
    Betas b;
    for (int k = 0; k < MAXK; ++k) b.v[k] = k < K ? betas_host[k] : 0.0;
    loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, b, K, out);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_loss_bwd(const float *scores, const float *labels, int64_t n,
                              const float *thresholds, const double *betas_host, int K, float alpha,
                              const float *th_xent, const double *acc, int loss_id,
                              const float *grad_out, float *dscores, double *dth, void *stream) {
    if (n < 0 || K < 0 || K > MAXK || !acc || (n > 0 && (!scores || !labels || !dscores)) ||
        (K > 0 && (!thresholds || !betas_host)) ||
        (loss_id != NPLDA_LOSS_SOFTCDET && loss_id != NPLDA_LOSS_CROSSENTROPY))
        return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    Betas b;
    for (int k = 0; k < MAXK; ++k) b.v[k] = k < K ? betas_host[k] : 0.0;
    loss_bwd_kernel<<<loss_grid(n), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
        scores, labels, n, thresholds, b, K, alpha, th_xent, acc, loss_id, grad_out, dscores, dth);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int nplda_minc_sweep(const float *tgt_sorted, int64_t n_t, const float *non_sorted,
                                int64_t n_n, float sum_t, float sum_n, const double *betas_host, int K,
                                float *out_min, int64_t *out_arg, void *stream) {
    if (n_t <= 0 || n_n < 0 || K <= 0 || K > MAXK || !tgt_sorted || (n_n > 0 && !non_sorted) ||
        !betas_host || !out_min || !out_arg)
        return NPLDA_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    BetasF b;
    for (int k = 0; k < MAXK; ++k) b.v[k] = k < K ? (float)betas_host[k] : 0.f;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_t + 255) / 256, 2 * (int64_t)sm_count()));
    float *blk_min = nullptr;
    int64_t *blk_arg = nullptr;
    // stream-ordered scratch: freed in stream order, no synchronisation
    NPLDA_CUDA_TRY(cudaMallocAsync(&blk_min, sizeof(float) * (size_t)K * grid, st));
    NPLDA_CUDA_TRY(cudaMallocAsync(&blk_arg, sizeof(int64_t) * (size_t)K * grid, st));
    minc_sweep_kernel<<<grid, 256, 0, st>>>(tgt_sorted, n_t, non_sorted, n_n, sum_t, sum_n, b, K, blk_min, blk_arg);
    NPLDA_LAUNCH_CHECK();
    minc_final_kernel<<<1, 256, 0, st>>>(blk_min, blk_arg, grid, K, out_min, out_arg);
    NPLDA_LAUNCH_CHECK();
    NPLDA_CUDA_TRY(cudaFreeAsync(blk_min, st));
    NPLDA_CUDA_TRY(cudaFreeAsync(blk_arg, st));
    return NPLDA_OK;
}
