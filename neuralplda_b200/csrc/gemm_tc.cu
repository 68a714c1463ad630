// Weight-gradient contraction on the tensor cores:  C[n][m] += sum_r A[r][m] * B[r][n]  over the batch rows r.
//
// The backward of the affine layers (autograd through models.py:366-370 / 478-481) needs
//     dW1 = DA^T X   [d1, d_in]     dW2 = DY^T U   [d2, d1]     (DPlda: dWw, dWb = G^T U)
// -- dense products whose reduction dimension is the batch.  Both factors are stored one batch row per line
// (fp32), i.e. K is the SLOW index of both operands.  tcgen05 wants K-major operands, so converter warps read
// 8 consecutive batch rows of one feature column (coalesced across the warp: 32 columns of one row = 128 B),
// split the values into bf16 hi/lo, and store the 16-byte K-runs straight into the canonical no-swizzle
// core-matrix layout in shared memory (8 features x 8 batch rows per 128-byte core matrix); one elected thread
// issues hi*hi + lo*hi + hi*lo (the "bf16x3" of score_tc.cu) into fp32 accumulators in tensor memory.
//
// One CTA owns up to 256 A-features (two M=128 tiles) x all B-features (N = 176, zero padded) and a contiguous
// range of batch rows; stages of 32 rows flow through a 3-deep ring; at the end the accumulators are added to C
// with coalesced fp32 reductions (C is [Nfeat][ldc], m contiguous).  Per 1 M pairs the dW1 product streams X
// once (4.1 GB) and DA twice: it is bound by HBM, not by the 6 MMAs per 16 rows.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nplda {
namespace gtc {

using namespace tc;

constexpr int NPADN = 176;                     // MMA N: B-side features, zero padded
constexpr int MT = 128;                        // A-side features per M tile
constexpr int KS = 32;                         // batch rows per stage
constexpr int NCHUNK = KS / 8;                 // K-chunks (8 rows) per stage
constexpr int KCH_A = (MT / 8) * 128;          // 2048 B: one K-chunk of an M tile
constexpr int KCH_B = (NPADN / 8) * 128;       // 2816 B
constexpr int A_HALF = NCHUNK * KCH_A;         // hi (or lo) image of one M tile
constexpr int A_TILE = 2 * A_HALF;
constexpr int B_HALF = NCHUNK * KCH_B;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_HALF;   // 55,296
constexpr int NSTAGE = 3;
constexpr int CONV_WARPS = 15;                 // + the MMA warp = 16 warps: 128 registers per thread
constexpr int NTHREADS = (CONV_WARPS + 1) * 32;
constexpr int HDR_BYTES = 1024;
constexpr int SMEM_BYTES = HDR_BYTES + NSTAGE * STAGE_BYTES;

struct Ring {
    uint32_t stage = 0, phase = 0;
    int n;
    __device__ explicit Ring(int n_) : n(n_) {}
    __device__ void advance() { if (++stage == (uint32_t)n) { stage = 0; phase ^= 1; } }
};

struct Args {
    const float *A; int lda; int Mfeat;
    const float *B; int ldb; int Nfeat;
    int64_t R, rows_per_cta;
    float *C; int ldc;
};

// bounded wait: a protocol bug traps after ~2 s instead of hanging the device
__device__ __forceinline__ void wait_or_trap(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_addr(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > NPLDA_WAIT_TRAP_CYCLES) __trap();
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) gemm_tn_tc_kernel(Args g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);            // [NSTAGE] converters -> MMA
    uint64_t *empty = full + NSTAGE;                                // [NSTAGE] MMA -> converters
    uint64_t *done = empty + NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    uint8_t *stages = smem + HDR_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 2 * MT;
    const int mtiles = (g.Mfeat - m0 > MT) ? 2 : 1;
    const int64_t r0 = blockIdx.y * g.rows_per_cta, r1 = min(g.R, r0 + g.rows_per_cta);
    if (r0 >= r1) return;
    const int nit = (int)((r1 - r0 + KS - 1) / KS);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], CONV_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == CONV_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = make_idesc_bf16(MT, NPADN);

    if (warp < CONV_WARPS) {
        // ---------------- converters: global fp32 -> bf16 hi/lo K-major core matrices ----------------
        // A warp-task = 32 consecutive columns of one operand x one 8-row group of the stage; warp w owns tasks
        // w, w + 15, w + 30, w + 45.  What a task touches is fixed for the whole kernel (column, stage offset), only
        // the rows advance: everything but the loads themselves is hoisted out of the loop, and the loads of stage
        // it + 1 are issued before stage it is converted (two register sets), so that ~8 KB per warp are in flight.
        const int ntask = 16 * mtiles + 24;
        const float *src[4];                          // first row of the task in stage 0 (nullptr: column is padding)
        int64_t ldq[4];
        int soff[4], grp8[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = warp + CONV_WARPS * q;
            soff[q] = -1; src[q] = nullptr; ldq[q] = 0; grp8[q] = 0;
            if (t >= ntask) continue;
            const bool isA = t < 16 * mtiles;
            const int u = isA ? t : t - 16 * mtiles;
            const int grp = u & 3, cw = u >> 2;                   // 8-row group, column group (A: mt * 4 + cw4)
            const int c = (isA ? (cw & 3) : cw) * 32 + lane;      // column inside the M tile / inside B
            const int mt = isA ? (cw >> 2) : 0;
            const int col = isA ? m0 + mt * MT + c : c;           // column in the source matrix
            if (!isA && c >= NPADN) continue;                     // the B image has 176 columns
            soff[q] = (isA ? mt * A_TILE + grp * KCH_A : 2 * A_TILE + grp * KCH_B) + (c >> 3) * 128 + (c & 7) * 16;
            grp8[q] = grp * 8;
            ldq[q] = isA ? g.lda : g.ldb;
            if (isA ? col < g.Mfeat : col < g.Nfeat) src[q] = (isA ? g.A : g.B) + col + (r0 + grp * 8) * ldq[q];
        }
        auto load_stage = [&](int it, float (&v)[4][8]) {
            const int64_t rb = r0 + (int64_t)it * KS;
            const bool full_stage = rb + KS <= r1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float *p = src[q] + (int64_t)it * KS * ldq[q];
                if (src[q] == nullptr) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[q][j] = 0.f;
                } else if (full_stage) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[q][j] = __ldg(p + j * ldq[q]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[q][j] = (rb + grp8[q] + j < r1) ? __ldg(p + j * ldq[q]) : 0.f;
                }
            }
        };
        auto store_stage = [&](const float (&v)[4][8], uint8_t *st) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (soff[q] < 0) continue;
                uint4 hi, lo;
                split_bf16x2(v[q][0], v[q][1], hi.x, lo.x);
                split_bf16x2(v[q][2], v[q][3], hi.y, lo.y);
                split_bf16x2(v[q][4], v[q][5], hi.z, lo.z);
                split_bf16x2(v[q][6], v[q][7], hi.w, lo.w);
                const bool isA = soff[q] < 2 * A_TILE;
                *reinterpret_cast<uint4 *>(st + soff[q]) = hi;
                *reinterpret_cast<uint4 *>(st + soff[q] + (isA ? A_HALF : B_HALF)) = lo;
            }
        };
        Ring ring(NSTAGE);
        float va[4][8], vb[4][8];
        load_stage(0, va);
        for (int it = 0; it < nit; it += 2) {
            if (it + 1 < nit) load_stage(it + 1, vb);
            wait_or_trap(&empty[ring.stage], ring.phase ^ 1);         // the MMAs that read this stage have completed
            store_stage(va, stages + ring.stage * STAGE_BYTES);
            fence_proxy_async();                                      // generic-proxy stores -> tcgen05.mma operand reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[ring.stage]);
            ring.advance();
            if (it + 1 >= nit) break;
            if (it + 2 < nit) load_stage(it + 2, va);
            wait_or_trap(&empty[ring.stage], ring.phase ^ 1);
            store_stage(vb, stages + ring.stage * STAGE_BYTES);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[ring.stage]);
            ring.advance();
        }
        // ---------------- epilogue: C[n][m] += D[m][n] ----------------
        wait_or_trap(done, 0);
        tc_fence_after();
        const int q4 = warp & 3;                                      // a warp reaches the TMEM lanes of quadrant warp % 4
        const int nwq = (CONV_WARPS - q4 + 3) / 4;                    // converter warps sharing this quadrant
        for (int mt = 0; mt < mtiles; ++mt) {
            const int m = m0 + mt * MT + 32 * q4 + lane;
            for (int blk = warp >> 2; blk < NPADN / 16; blk += nwq) {
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(32 * q4) << 16) + mt * NPADN + blk * 16, r);
                tmem_ld_wait();
                if (m < g.Mfeat) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = blk * 16 + e;
                        if (n < g.Nfeat) atomicAdd(g.C + (int64_t)n * g.ldc + m, __uint_as_float(r[e]));
                    }
                }
            }
        }
    } else {
        // ---------------- MMA warp ----------------
        Ring ring(NSTAGE);
        const uint32_t sbase = smem_addr(stages);
        for (int it = 0; it < nit; ++it, ring.advance()) {
            wait_or_trap(&full[ring.stage], ring.phase);
            tc_fence_after();
            const uint32_t st = sbase + ring.stage * STAGE_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KS / 16; ++ks) {
                    const uint64_t bhi = make_smem_desc(st + 2 * A_TILE + 2 * ks * KCH_B, KCH_B, 128);
                    const uint64_t blo = make_smem_desc(st + 2 * A_TILE + B_HALF + 2 * ks * KCH_B, KCH_B, 128);
                    for (int mt = 0; mt < mtiles; ++mt) {
                        const uint64_t ahi = make_smem_desc(st + mt * A_TILE + 2 * ks * KCH_A, KCH_A, 128);
                        const uint64_t alo = make_smem_desc(st + mt * A_TILE + A_HALF + 2 * ks * KCH_A, KCH_A, 128);
                        const uint32_t d = tmem + mt * NPADN;
                        mma_ss(d, ahi, bhi, IDESC, (it | ks) != 0);
                        mma_ss(d, alo, bhi, IDESC, 1);
                        mma_ss(d, ahi, blo, IDESC, 1);
                    }
                }
                mma_commit(&empty[ring.stage]);
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CONV_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace gtc

// Returns NPLDA_ERR_UNSUPPORTED_DIM when the shape is not one this kernel takes (caller falls back to SIMT).
int gemm_tn_tc(const float *A, int lda, int M, const float *B, int ldb, int N, int64_t R, float *C, int ldc,
               cudaStream_t st) {
    if (R <= 0 || M <= 0 || N <= 0) return NPLDA_OK;
    if (N > gtc::NPADN) return NPLDA_ERR_UNSUPPORTED_DIM;
    const int mpairs = (M + 2 * gtc::MT - 1) / (2 * gtc::MT);
    int splits = std::max(1, sm_count() / mpairs);
    int64_t rows = (R + splits - 1) / splits;
    rows = std::max<int64_t>((rows + gtc::KS - 1) / gtc::KS * gtc::KS, 8 * gtc::KS);
    splits = (int)((R + rows - 1) / rows);
    gtc::Args a{A, lda, M, B, ldb, N, R, rows, C, ldc};
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(gtc::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::SMEM_BYTES));
    gtc::gemm_tn_tc_kernel<<<dim3(mpairs, splits), gtc::NTHREADS, gtc::SMEM_BYTES, st>>>(a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda
