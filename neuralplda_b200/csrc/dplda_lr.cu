// Gradient of DPlda's logistic_regres (autograd through models.py:483-489) from the normalised rows u kept by the
// training forward (dplda_score_fwd_train_u) -- the whole backward when the LDA is frozen, which is how the reference's
// driver trains (xvector_DPlda_pytorch.py:140-147).
//
// With s = u1 + u2, d = u1 - u2 and g = dL/dS:
//     u1 u1^T + u2 u2^T = (s s^T + d d^T) / 2        u1 u2^T + u2 u1^T = (s s^T - d d^T) / 2
// so   dWw = (Ps + Pd) / 2,  dWb = (Ps - Pd) / 2,  Ps = sum_p g_p s_p s_p^T,  Pd = sum_p g_p d_p d_p^T
// -- two SYMMETRIC batch contractions instead of four (and dws = sum_p g_p s_p, dc = sum_p g_p).
//
// One kernel (wgrad_kernel), one pass over the rows: a CTA owns a contiguous range of pairs; 24 converter warps read
// u1, u2 (each element once), form s, d, g s, g d, split them into bf16 hi/lo and store the 16-byte K-runs in the K-major
// core-matrix layout (K = pairs); one elected thread issues hi*hi + lo*hi + hi*lo into four accumulators in tensor memory:
// rows [0, 128) x all 176 columns of Ps and Pd, plus their [128, 176) x [128, 176) corner blocks -- the rest follows from
// symmetry.  The epilogue adds (Ps +- Pd) / 2 to the gradient with fp32 reductions.
#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nplda {

namespace dlr {

using namespace tc;

constexpr int LD = 176;                   // floats per row (EMIT_LD of score_tc.cu) = MMA N of the big blocks
constexpr int MT = 128;                   // feature rows of the big blocks; the corner block covers [MT, LD)
constexpr int NC = LD - MT;               // 48
constexpr int KS = 32;                    // pairs per stage
constexpr int KCH = (LD / 8) * 128;       // 2816 B: one K-chunk (8 pairs) of an image: 22 core matrices of 8 features x 8 pairs
constexpr int IMG = (KS / 8) * KCH;       // 11264 B: hi (or lo) image of one of s, d, g s, g d
constexpr int STAGE_BYTES = 8 * IMG;      // [s hi][s lo][d hi][d lo][gs hi][gs lo][gd hi][gd lo]
constexpr int NSTAGE = 2;
constexpr int CONV_WARPS = 24;            // warp w: 8-pair group w & 3, 32-column group w >> 2
constexpr int NTHREADS = (CONV_WARPS + 1) * 32;
constexpr int HDR_BYTES = 1024;
// the corner blocks' A operand (M = 128) starts at feature 128 and reads 80 feature rows past the image: finite bf16
// values of the following chunks / images, rows whose results are never read; the slack keeps the last one in bounds
constexpr int SMEM_BYTES = HDR_BYTES + NSTAGE * STAGE_BYTES + 16 * 128;
// TMEM columns: Ps[0:128][:] | Ps corner | Pd[0:128][:] | Pd corner
constexpr int COL_S0 = 0, COL_S1 = LD, COL_D0 = LD + NC, COL_D1 = 2 * LD + NC;

struct Args {
    const float *U0, *U1, *g;      // side 0 / side 1 rows of this launch, dL/dS
    int64_t n, rows_per_cta;
    int d1;
    float *dWb, *dWw, *dws, *dc;   // dWb / dWw: [d1][d1]; any may be null (dWb and dWw together)
};

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

__global__ void __launch_bounds__(NTHREADS, 1) wgrad_kernel(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);            // [NSTAGE] converters -> MMA
    uint64_t *empty = full + NSTAGE;                                // [NSTAGE] MMA -> converters
    uint64_t *done = empty + NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    uint8_t *stages = smem + HDR_BYTES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = blockIdx.x * a.rows_per_cta, r1 = min(a.n, r0 + a.rows_per_cta);
    if (r0 >= r1) return;
    const int nit = (int)((r1 - r0 + KS - 1) / KS);

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], CONV_WARPS); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    // the slack behind the last stage is read (never used): keep it finite
    for (int i = tid; i < 16 * 128 / 4; i += NTHREADS) reinterpret_cast<uint32_t *>(stages + NSTAGE * STAGE_BYTES)[i] = 0u;
    fence_proxy_async();
    if (warp == CONV_WARPS) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC_BIG = make_idesc_bf16(MT, LD), IDESC_CORNER = make_idesc_bf16(MT, NC);

    if (warp < CONV_WARPS) {
        // ---------------- converters ----------------
        const int grp = warp & 3, c = (warp >> 2) * 32 + lane;       // 8-pair group of the stage, column
        const bool col_ok = c < LD;                                  // the sixth column group is half empty
        const bool col_live = c < a.d1;
        const int soff = grp * KCH + (c >> 3) * 128 + (c & 7) * 16;  // this thread's 16-byte K-run inside an image
        auto load_stage = [&](int it, float (&u0)[8], float (&u1)[8], float (&gg)[8]) {
            const int64_t rb = r0 + (int64_t)it * KS + grp * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = col_live && rb + j < r1;
                u0[j] = ok ? __ldg(a.U0 + (rb + j) * LD + c) : 0.f;
                u1[j] = ok ? __ldg(a.U1 + (rb + j) * LD + c) : 0.f;
                gg[j] = rb + j < r1 ? __ldg(a.g + rb + j) : 0.f;
            }
        };
        float wsum = 0.f, gsum = 0.f;
        auto store_stage = [&](const float (&u0)[8], const float (&u1)[8], const float (&gg)[8], uint8_t *st) {
            float s[8], d[8], gs[8], gd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] = u0[j] + u1[j]; d[j] = u0[j] - u1[j];
                gs[j] = gg[j] * s[j]; gd[j] = gg[j] * d[j];
                wsum += gs[j]; gsum += gg[j];
            }
            if (!col_ok) return;
            auto put = [&](const float (&v)[8], int img) {
                uint4 hi, lo;
                split_bf16x2(v[0], v[1], hi.x, lo.x); split_bf16x2(v[2], v[3], hi.y, lo.y);
                split_bf16x2(v[4], v[5], hi.z, lo.z); split_bf16x2(v[6], v[7], hi.w, lo.w);
                *reinterpret_cast<uint4 *>(st + (2 * img) * IMG + soff) = hi;
                *reinterpret_cast<uint4 *>(st + (2 * img + 1) * IMG + soff) = lo;
            };
            put(s, 0); put(d, 1); put(gs, 2); put(gd, 3);
        };
        uint32_t stage = 0, phase = 0;
        float a0[8], a1[8], ag[8], b0[8], b1[8], bg[8];
        load_stage(0, a0, a1, ag);
        for (int it = 0; it < nit; it += 2) {
            if (it + 1 < nit) load_stage(it + 1, b0, b1, bg);
            wait_or_trap(&empty[stage], phase ^ 1);
            store_stage(a0, a1, ag, stages + stage * STAGE_BYTES);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            if (it + 1 >= nit) break;
            if (it + 2 < nit) load_stage(it + 2, a0, a1, ag);
            wait_or_trap(&empty[stage], phase ^ 1);
            store_stage(b0, b1, bg, stages + stage * STAGE_BYTES);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (a.dws && col_live) atomicAdd(a.dws + c, wsum);
        // the four row groups of column group 0 cover every pair once, all lanes with the same sum: lane 0 adds it
        if (a.dc && (warp >> 2) == 0 && lane == 0) atomicAdd(a.dc, gsum);
        // ---------------- epilogue: dWw += (Ps + Pd) / 2, dWb += (Ps - Pd) / 2 ----------------
        wait_or_trap(done, 0);
        tc_fence_after();
        if (a.dWw != nullptr) {
            const int q4 = warp & 3, sub = warp >> 2;                 // TMEM quadrant; six warps share it
            const int m = 32 * q4 + lane;                             // feature row of the big blocks / m - 128 of the corner
            const uint32_t lane_base = tmem + ((uint32_t)(32 * q4) << 16);
            for (int blk = sub; blk < LD / 16; blk += CONV_WARPS / 4) {
                uint32_t ps[16], pd[16];
                tmem_ld16(lane_base + COL_S0 + blk * 16, ps);
                tmem_ld16(lane_base + COL_D0 + blk * 16, pd);
                tmem_ld_wait();
                if (m < a.d1) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = blk * 16 + e;
                        if (n >= a.d1) continue;
                        const float vs = __uint_as_float(ps[e]), vd = __uint_as_float(pd[e]);
                        const float ww = 0.5f * (vs + vd), wb = 0.5f * (vs - vd);
                        atomicAdd(a.dWw + (int64_t)n * a.d1 + m, ww);         // [n][m]: lanes write consecutive addresses
                        atomicAdd(a.dWb + (int64_t)n * a.d1 + m, wb);
                        if (n >= MT) {                                       // the mirror entries [m][n] of rows >= 128
                            atomicAdd(a.dWw + (int64_t)m * a.d1 + n, ww);
                            atomicAdd(a.dWb + (int64_t)m * a.d1 + n, wb);
                        }
                    }
                }
            }
            // corner blocks: rows / columns [128, 176) (lanes 0..47 of the accumulator)
            if (q4 < 2) {                                             // warp-uniform: tcgen05.ld is a whole-warp instruction
                for (int blk = sub; blk < NC / 16; blk += CONV_WARPS / 4) {
                    uint32_t ps[16], pd[16];
                    tmem_ld16(lane_base + COL_S1 + blk * 16, ps);
                    tmem_ld16(lane_base + COL_D1 + blk * 16, pd);
                    tmem_ld_wait();
                    if (m >= NC || MT + m >= a.d1) continue;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = MT + blk * 16 + e;
                        if (n >= a.d1) continue;
                        const float vs = __uint_as_float(ps[e]), vd = __uint_as_float(pd[e]);
                        atomicAdd(a.dWw + (int64_t)n * a.d1 + MT + m, 0.5f * (vs + vd));
                        atomicAdd(a.dWb + (int64_t)n * a.d1 + MT + m, 0.5f * (vs - vd));
                    }
                }
            }
        }
    } else {
        // ---------------- MMA warp ----------------
        uint32_t stage = 0, phase = 0;
        const uint32_t sbase = smem_addr(stages);
        for (int it = 0; it < nit; ++it) {
            wait_or_trap(&full[stage], phase);
            tc_fence_after();
            const uint32_t st = sbase + stage * STAGE_BYTES;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KS / 16; ++ks) {
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {                  // 0: s (images 0, 2), 1: d (images 1, 3)
                        const uint32_t x = st + (2 * pr) * IMG + 2 * ks * KCH, gx = st + (2 * (pr + 2)) * IMG + 2 * ks * KCH;
                        const uint64_t bhi = make_smem_desc(x, KCH, 128), blo = make_smem_desc(x + IMG, KCH, 128);
                        const uint64_t ahi = make_smem_desc(gx, KCH, 128), alo = make_smem_desc(gx + IMG, KCH, 128);
                        const uint32_t dbig = tmem + (pr ? COL_D0 : COL_S0), dcor = tmem + (pr ? COL_D1 : COL_S1);
                        const uint32_t acc = (it | ks) != 0;
                        mma_ss(dbig, ahi, bhi, IDESC_BIG, acc);
                        mma_ss(dbig, alo, bhi, IDESC_BIG, 1);
                        mma_ss(dbig, ahi, blo, IDESC_BIG, 1);
                        constexpr uint32_t off = (MT / 8) * 128 >> 4;          // feature 128: 16 core matrices into a chunk
                        mma_ss(dcor, ahi + off, bhi + off, IDESC_CORNER, acc);
                        mma_ss(dcor, alo + off, bhi + off, IDESC_CORNER, 1);
                        mma_ss(dcor, ahi + off, blo + off, IDESC_CORNER, 1);
                    }
                }
                mma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CONV_WARPS) tmem_dealloc(tmem, 512);
}

}  // namespace dlr

int64_t dplda_lr_workspace_bytes(int64_t, int) { return 256; }      // none needed (kept in the ABI for callers that size one)

// urows: [2 cap][176], side 1 `cap` rows after side 0
int dplda_lr_grad(const float *urows, int64_t n, int64_t cap, int d1, const float *dscores, float *dw_lr, float *dc_lr,
                  void *workspace, int64_t workspace_bytes, cudaStream_t st) {
    (void)workspace; (void)workspace_bytes;
    if (d1 > dlr::LD || d1 < 1) return NPLDA_ERR_UNSUPPORTED_DIM;
    dlr::Args a;
    a.U0 = urows; a.U1 = urows + cap * dlr::LD; a.g = dscores; a.n = n; a.d1 = d1;
    a.dWb = dw_lr; a.dWw = dw_lr ? dw_lr + (int64_t)d1 * d1 : nullptr; a.dws = dw_lr ? dw_lr + 2 * (int64_t)d1 * d1 : nullptr;
    a.dc = dc_lr;
    int64_t rows = (n + sm_count() - 1) / sm_count();
    rows = std::max<int64_t>((rows + dlr::KS - 1) / dlr::KS * dlr::KS, 4 * dlr::KS);
    a.rows_per_cta = rows;
    const int grid = (int)((n + rows - 1) / rows);
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(dlr::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dlr::SMEM_BYTES));
    dlr::wgrad_kernel<<<grid, dlr::NTHREADS, dlr::SMEM_BYTES, st>>>(a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda
