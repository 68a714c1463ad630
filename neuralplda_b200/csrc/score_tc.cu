// K1 (tensor cores): fused pairwise score kernel on tcgen05 / TMEM.
//
// Precision: every fp32 operand is split x = hi + lo into two bf16 values and each
// product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM
// ("bf16x3"; the dropped lo*lo term and the split residuals are O(2^-17)), which keeps
// the scores within ~1e-5 of the fp32 reference (numerics).
//
// One persistent CTA per SM, warp-specialised, tiles of 64 trial pairs = 128 rows with
// in every 16-row group rows 0-7 are side 0 and rows 8-15 side 1 of the same 8 pairs:
//   converters (8 warps)  x rows: global -> registers (16 B per lane, 64 B per row per
//                         request) -> bf16 hi/lo -> tcgen05.st.16x256b into a 5-stage
//                         ring of A operands in TENSOR MEMORY (no shared-memory traffic)
//   B loader   (1 thread) weight images (bf16 hi/lo, already in the tcgen05 K-major
//                         core-matrix layout, packed once per parameter update) ->
//                         5-stage shared-memory ring with 1-D bulk async copies (TMA
//                         engine) completing on mbarriers
//   MMA issuer (1 thread) layer 1:  D[128x176] += A(tmem) * W1^T   (3 MMAs per K=16 step)
//                         layer 2:  Y[128x176]  = U(smem) * W2^T   (Y overwrites D)
//                         two D buffers in TMEM: layer 1 of tile t+1 runs under the
//                         epilogue of tile t
//   epilogue   (4 warps)  thread = row: D -> +b1 -> |a| -> u = a/|a| -> bf16 hi/lo ->
//                         shared memory (A operand of layer 2); then Y -> +b2 ->
//                         S = sum Q y1^2 + Q y2^2 + 2 P y1 y2 with the partner row
//                         fetched by warp shuffle; one 4-byte store per pair
#include <algorithm>
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace nplda {
namespace tcg {

using namespace tc;

constexpr int TP = 64;               // pairs per tile
constexpr int NPAD = 176;            // MMA N (both layers): 170 padded to a multiple of 16
constexpr int KST = 32;              // K per pipeline stage (two MMA K-steps)
constexpr int KCH_B = (NPAD / 8) * 128;        // 2816 B: one 8-wide k-chunk of B (22 core matrices)
constexpr int B_STAGE = 2 * 4 * KCH_B;         // 22528 B: hi + lo, K = 32
constexpr int NB = 5;                          // B ring stages
constexpr int KCH_U = (128 / 8) * 128;         // 2048 B: one k-chunk of U (16 core matrices)
constexpr int U_HALF = (NPAD / 8) * KCH_U;     // 45056 B (hi or lo), K = 176
constexpr int NA = 5;                          // A ring stages in TMEM
constexpr int A_COL0 = 2 * NPAD;               // TMEM columns: D0 [0,176) D1 [176,352) A ring [352,512)
constexpr int A_STAGE_COLS = 32;               // per stage: 16 columns hi + 16 columns lo (K = 32)

constexpr int EPI_WARPS = 4, CONV_WARPS = 8;
constexpr int WARP_MMA = EPI_WARPS + CONV_WARPS, WARP_LOAD = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_LOAD + 1) * 32;   // 448

// shared-memory map (bytes)
constexpr int SM_B = 0;
constexpr int SM_U = SM_B + NB * B_STAGE;                // 112640
constexpr int SM_PAR = SM_U + 2 * U_HALF;                // 202752: b1, b2, P, Q (NPAD floats each)
constexpr int SM_BAR = SM_PAR + 4 * NPAD * 4;            // 205568
constexpr int N_BARS = 2 * NA + 2 * NB + 2 + 2 + 2 + 2;  // a_full/empty, b_full/empty, d_full, d_empty, y_full, u_full/u_empty
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16;

struct Args {
    const float *x1, *x2;
    const int64_t *i1, *i2;
    int64_t n_rows;
    int32_t *bad_flag;
    int64_t n;
    int d_in;               // multiple of 32
    int nst1;               // layer-1 stages  = d_in / 32
    int ksteps2;            // layer-2 K steps = round_up(d1, 16) / 16
    const uint8_t *w1img, *w2img;
    const float *b1, *b2, *p, *q;   // padded to >= NPAD floats
    float *scores;
    int dbg;                // bottleneck experiments (env NPLDA_TC_DEBUG): 1 cached x, 2 no weight copies, 4 no MMAs
};

struct Ring {
    uint32_t stage = 0, phase = 0;
    int n;
    __device__ explicit Ring(int n_) : n(n_) {}
    __device__ void advance() { if (++stage == (uint32_t)n) { stage = 0; phase ^= 1; } }
};

__device__ __forceinline__ float4 ldg_stream(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

template <bool INDEXED>
__global__ void __launch_bounds__(NTHREADS, 1) score_tc_kernel(Args g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *Bs = smem + SM_B;
    uint8_t *Us = smem + SM_U;
    float *par = reinterpret_cast<float *>(smem + SM_PAR);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *a_full = bars, *a_empty = bars + NA, *b_full = bars + 2 * NA, *b_empty = bars + 2 * NA + NB;
    uint64_t *d_full = bars + 2 * NA + 2 * NB, *d_empty = d_full + 2, *y_full = d_full + 4;
    uint64_t *u_full = d_full + 6, *u_empty = d_full + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (g.n + TP - 1) / TP;
    const int64_t T = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- one-time setup ----
    for (int i = tid; i < NPAD; i += NTHREADS) {
        par[i] = g.b1[i]; par[NPAD + i] = g.b2[i]; par[2 * NPAD + i] = g.p[i]; par[3 * NPAD + i] = g.q[i];
    }
    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], CONV_WARPS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&d_full[d], 1); mbar_init(&d_empty[d], EPI_WARPS * 32); mbar_init(&y_full[d], 1); }
        mbar_init(u_full, EPI_WARPS * 32);
        mbar_init(u_empty, 1);
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = make_idesc_bf16(128, NPAD);

    if (warp < EPI_WARPS) {
        // =============================== EPILOGUE ===============================
        // Row map inside each 16-lane group: lanes 0-7 = side 0, lanes 8-15 = side 1 of pairs 0-7, so a
        // tcgen05.ld.16x256b hands thread t both sides of pair (t >> 2) for columns {2(t&3), 2(t&3)+1} of
        // every 8-column block: no cross-lane traffic per element, two shuffles per row sum at the end.
        const int rsub = lane >> 2, cq = lane & 3;
        const float2 *b1s = reinterpret_cast<const float2 *>(par);
        const float2 *b2s = reinterpret_cast<const float2 *>(par + NPAD);
        const float2 *ps = reinterpret_cast<const float2 *>(par + 2 * NPAD);
        const float2 *qs = reinterpret_cast<const float2 *>(par + 3 * NPAD);
        for (int64_t i = 0; i < T; ++i) {
            const int d = (int)(i & 1);
            const uint32_t par_d = (uint32_t)((i >> 1) & 1);
            float rinv[2][2];
            // ---- layer-1 accumulator: a = D + b1, |a|, bf16 hi/lo of a -> U (normalised after layer 2) ----
            mbar_wait(&d_full[d], par_d);
            tc_fence_after();
            mbar_wait(u_empty, (uint32_t)((i & 1) ^ 1));           // layer 2 of the previous tile has read U
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int mrow = warp * 32 + h * 16 + rsub;          // side-0 row; side 1 is mrow + 8
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32 + h * 16) << 16) + d * NPAD;
                uint8_t *u0 = Us + (mrow >> 3) * 128 + (mrow & 7) * 16 + cq * 4;   // + kchunk * KCH_U
                uint8_t *u1 = u0 + 128;                                               // row + 8: next core matrix
                float ss0a = 0.f, ss0b = 0.f, ss1a = 0.f, ss1b = 0.f;
#pragma unroll 2
                for (int c0 = 0; c0 < NPAD; c0 += 16) {
                    uint32_t v[8];
                    tmem_ld_16x256b_x2(taddr + c0, v);
                    tmem_ld_wait();
                    const float2 ba = b1s[(c0 >> 1) + cq], bb = b1s[(c0 >> 1) + 4 + cq];
                    const float a00 = __uint_as_float(v[0]) + ba.x, a01 = __uint_as_float(v[1]) + ba.y;
                    const float a10 = __uint_as_float(v[2]) + ba.x, a11 = __uint_as_float(v[3]) + ba.y;
                    const float a02 = __uint_as_float(v[4]) + bb.x, a03 = __uint_as_float(v[5]) + bb.y;
                    const float a12 = __uint_as_float(v[6]) + bb.x, a13 = __uint_as_float(v[7]) + bb.y;
                    ss0a = fmaf(a00, a00, ss0a); ss0b = fmaf(a01, a01, ss0b);
                    ss0a = fmaf(a02, a02, ss0a); ss0b = fmaf(a03, a03, ss0b);
                    ss1a = fmaf(a10, a10, ss1a); ss1b = fmaf(a11, a11, ss1b);
                    ss1a = fmaf(a12, a12, ss1a); ss1b = fmaf(a13, a13, ss1b);
                    uint32_t hi, lo;
                    const int kc = c0 >> 3;
                    split_bf16x2(a00, a01, hi, lo);
                    *reinterpret_cast<uint32_t *>(u0 + kc * KCH_U) = hi;
                    *reinterpret_cast<uint32_t *>(u0 + U_HALF + kc * KCH_U) = lo;
                    split_bf16x2(a02, a03, hi, lo);
                    *reinterpret_cast<uint32_t *>(u0 + (kc + 1) * KCH_U) = hi;
                    *reinterpret_cast<uint32_t *>(u0 + U_HALF + (kc + 1) * KCH_U) = lo;
                    split_bf16x2(a10, a11, hi, lo);
                    *reinterpret_cast<uint32_t *>(u1 + kc * KCH_U) = hi;
                    *reinterpret_cast<uint32_t *>(u1 + U_HALF + kc * KCH_U) = lo;
                    split_bf16x2(a12, a13, hi, lo);
                    *reinterpret_cast<uint32_t *>(u1 + (kc + 1) * KCH_U) = hi;
                    *reinterpret_cast<uint32_t *>(u1 + U_HALF + (kc + 1) * KCH_U) = lo;
                }
                float ss0 = ss0a + ss0b, ss1 = ss1a + ss1b;
                ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
                ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
                rinv[h][0] = 1.f / fmaxf(sqrtf(ss0), 1e-12f);      // F.normalize eps (models.py:368)
                rinv[h][1] = 1.f / fmaxf(sqrtf(ss1), 1e-12f);
            }
            fence_proxy_async();      // U is read by tcgen05.mma (async proxy)
            tc_fence_before();        // our TMEM reads of D are done before Y overwrites it
            mbar_arrive(u_full);
            // ---- layer-2 accumulator: y = Y / |a| + b2, pair score ----
            mbar_wait(&y_full[d], par_d);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32 + h * 16) << 16) + d * NPAD;
                const float r0 = rinv[h][0], r1 = rinv[h][1];
                float sa = 0.f, sb = 0.f;
#pragma unroll 2
                for (int c0 = 0; c0 < NPAD; c0 += 16) {
                    uint32_t v[8];
                    tmem_ld_16x256b_x2(taddr + c0, v);
                    tmem_ld_wait();
                    const int ci = (c0 >> 1) + cq;
                    const float2 ba = b2s[ci], bb = b2s[ci + 4], pa = ps[ci], pb = ps[ci + 4], qa = qs[ci], qb = qs[ci + 4];
                    const float y00 = fmaf(__uint_as_float(v[0]), r0, ba.x), y01 = fmaf(__uint_as_float(v[1]), r0, ba.y);
                    const float y10 = fmaf(__uint_as_float(v[2]), r1, ba.x), y11 = fmaf(__uint_as_float(v[3]), r1, ba.y);
                    const float y02 = fmaf(__uint_as_float(v[4]), r0, bb.x), y03 = fmaf(__uint_as_float(v[5]), r0, bb.y);
                    const float y12 = fmaf(__uint_as_float(v[6]), r1, bb.x), y13 = fmaf(__uint_as_float(v[7]), r1, bb.y);
                    sa = fmaf(qa.x, fmaf(y00, y00, y10 * y10), sa); sa = fmaf(2.f * pa.x, y00 * y10, sa);
                    sb = fmaf(qa.y, fmaf(y01, y01, y11 * y11), sb); sb = fmaf(2.f * pa.y, y01 * y11, sb);
                    sa = fmaf(qb.x, fmaf(y02, y02, y12 * y12), sa); sa = fmaf(2.f * pb.x, y02 * y12, sa);
                    sb = fmaf(qb.y, fmaf(y03, y03, y13 * y13), sb); sb = fmaf(2.f * pb.y, y03 * y13, sb);
                }
                float sc = sa + sb;
                sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                const int64_t pr = (blockIdx.x + i * gridDim.x) * TP + warp * 16 + h * 8 + rsub;
                if (cq == 0 && pr < g.n) g.scores[pr] = sc;
            }
            tc_fence_before();
            mbar_arrive(&d_empty[d]);
        }
    } else if (warp < EPI_WARPS + CONV_WARPS) {
        // =============================== CONVERTERS ===============================
        const int cw = warp - EPI_WARPS;
        const int q = warp & 3, h = cw >> 2;                  // TMEM quadrant of this warp, 16-row half
        const int pl = q * 16 + h * 8 + (lane >> 2);          // pair (within the tile) of TMEM rows m and m + 8
        const int kq = (lane & 3) * 4;                        // k offset inside a 16-wide K step
        const uint32_t st_addr = tmem + ((uint32_t)(q * 32 + h * 16) << 16) + A_COL0;
        Ring ra(NA);
        const float *r0 = nullptr, *r1 = nullptr;
        auto set_rows = [&](int64_t i) {                       // r0: side 0 (x1), r1: side 1 (x2) of pair pl
            const int64_t pr = min((blockIdx.x + i * gridDim.x) * TP + pl, g.n - 1);
            if (INDEXED) {
                int64_t ia = g.i1[pr], ib = g.i2[pr];
                if (ia < 0 || ia >= g.n_rows) { *g.bad_flag = 1; ia = 0; }
                if (ib < 0 || ib >= g.n_rows) { *g.bad_flag = 1; ib = 0; }
                r0 = g.x1 + ia * g.d_in + kq;
                r1 = g.x1 + ib * g.d_in + kq;
            } else {
                r0 = g.x1 + pr * g.d_in + kq;
                r1 = g.x2 + pr * g.d_in + kq;
            }
        };
        // Software pipeline: PF stages (PF * 64 B per thread) of x are in flight while one is
        // converted -- 8 warps x 32 lanes x PF x 64 B = 64 KB per SM.
        constexpr int PF = 4;
        int64_t ip = 0;
        int sp = 0;
        auto issue = [&](float4 (&v)[4]) {
            if (ip < T) {
                const int k = (g.dbg & 1) ? 0 : sp * KST;           // dbg 1: every load hits the same (cached) line
                v[0] = ldg_stream(r0 + k); v[1] = ldg_stream(r0 + k + 16);
                v[2] = ldg_stream(r1 + k); v[3] = ldg_stream(r1 + k + 16);
                if (++sp == g.nst1) { sp = 0; if (++ip < T) set_rows(ip); }
            }
        };
        float4 buf[PF][4];
        if (T > 0) set_rows(0);
#pragma unroll
        for (int u = 0; u < PF; ++u) issue(buf[u]);
        const int64_t total = T * g.nst1;
        for (int64_t it = 0; it < total; it += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                if (it + u < total) {
                    // registers of tcgen05.st.16x256b.x2: r0,r1 -> (row, cols 2j,2j+1)  r2,r3 -> (row+8, same
                    // cols); r4..r7 the same for the next 8 columns (k + 16)
                    const float4(&c)[4] = buf[u];
                    uint32_t hi[8], lo[8];
                    split_bf16x2(c[0].x, c[0].y, hi[0], lo[0]); split_bf16x2(c[0].z, c[0].w, hi[1], lo[1]);
                    split_bf16x2(c[2].x, c[2].y, hi[2], lo[2]); split_bf16x2(c[2].z, c[2].w, hi[3], lo[3]);
                    split_bf16x2(c[1].x, c[1].y, hi[4], lo[4]); split_bf16x2(c[1].z, c[1].w, hi[5], lo[5]);
                    split_bf16x2(c[3].x, c[3].y, hi[6], lo[6]); split_bf16x2(c[3].z, c[3].w, hi[7], lo[7]);
                    issue(buf[u]);                                   // refill this slot: PF stages ahead
                    mbar_wait(&a_empty[ra.stage], ra.phase ^ 1);
                    tc_fence_after();
                    const uint32_t col = st_addr + ra.stage * A_STAGE_COLS;
                    tmem_st_16x256b_x2(col, hi);
                    tmem_st_16x256b_x2(col + 16, lo);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[ra.stage]);
                    ra.advance();
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA ISSUER ===============================
        if (lane == 0) {
            Ring ra(NA), rb(NB);
            const uint32_t b_base = smem_addr(Bs), u_base = smem_addr(Us);
            const int half = g.nst1 / 2;
            const bool skip_mma = (g.dbg & 4) != 0;
            auto layer1 = [&](uint32_t dcol, int s_begin, int s_end) {
                for (int s = s_begin; s < s_end; ++s) {
                    mbar_wait(&a_full[ra.stage], ra.phase);
                    mbar_wait(&b_full[rb.stage], rb.phase);
                    tc_fence_after();
                    const uint32_t acol = tmem + A_COL0 + ra.stage * A_STAGE_COLS;
                    const uint32_t bs = b_base + rb.stage * B_STAGE;
                    if (!skip_mma) {
#pragma unroll
                        for (int st = 0; st < 2; ++st) {
                            const uint64_t bhi = make_smem_desc(bs + st * 2 * KCH_B, KCH_B, 128);
                            const uint64_t blo = make_smem_desc(bs + 4 * KCH_B + st * 2 * KCH_B, KCH_B, 128);
                            mma_ts(dcol, acol + st * 8, bhi, IDESC, (s | st) != 0);
                            mma_ts(dcol, acol + 16 + st * 8, bhi, IDESC, 1);
                            mma_ts(dcol, acol + st * 8, blo, IDESC, 1);
                        }
                    }
                    mma_commit(&a_empty[ra.stage]);
                    mma_commit(&b_empty[rb.stage]);
                    ra.advance();
                    rb.advance();
                }
            };
            // Order per iteration: first half of layer 1 (tile i), layer 2 (tile i-1), second half of layer 1.
            // Layer 2 lands mid-way so that the final epilogue of tile i-1 runs under the second half and
            // the D buffer it frees is ready when layer 1 of tile i+1 starts.  The B loader follows the
            // same order.
            for (int64_t i = 0; i <= T; ++i) {
                const uint32_t dcol_i = tmem + (uint32_t)(i & 1) * NPAD;
                if (i < T) {
                    mbar_wait(&d_empty[i & 1], (uint32_t)(((i >> 1) & 1) ^ 1));
                    tc_fence_after();
                    layer1(dcol_i, 0, half);
                }
                if (i >= 1) {      // layer 2 of tile i - 1
                    const int64_t j = i - 1;
                    const int d = (int)(j & 1);
                    const uint32_t dcol = tmem + d * NPAD;
                    mbar_wait(u_full, (uint32_t)(j & 1));
                    tc_fence_after();
                    for (int ks = 0; ks < g.ksteps2; ks += 2) {
                        const int nst = min(2, g.ksteps2 - ks);
                        mbar_wait(&b_full[rb.stage], rb.phase);
                        tc_fence_after();
                        const uint32_t bs = b_base + rb.stage * B_STAGE;
                        for (int st = 0; st < nst && !skip_mma; ++st) {
                            const uint64_t bhi = make_smem_desc(bs + st * 2 * KCH_B, KCH_B, 128);
                            const uint64_t blo = make_smem_desc(bs + nst * 2 * KCH_B + st * 2 * KCH_B, KCH_B, 128);
                            const uint64_t uhi = make_smem_desc(u_base + (ks + st) * 2 * KCH_U, KCH_U, 128);
                            const uint64_t ulo = make_smem_desc(u_base + U_HALF + (ks + st) * 2 * KCH_U, KCH_U, 128);
                            mma_ss(dcol, uhi, bhi, IDESC, (ks | st) != 0);
                            mma_ss(dcol, ulo, bhi, IDESC, 1);
                            mma_ss(dcol, uhi, blo, IDESC, 1);
                        }
                        mma_commit(&b_empty[rb.stage]);
                        rb.advance();
                    }
                    mma_commit(&y_full[d]);
                    mma_commit(u_empty);
                }
                if (i < T) {
                    layer1(dcol_i, half, g.nst1);
                    mma_commit(&d_full[i & 1]);
                }
            }
        }
    } else {
        // =============================== B LOADER ===============================
        if (lane == 0) {
            Ring rb(NB);
            const int nst2 = (g.ksteps2 + 1) / 2;
            const int half = g.nst1 / 2;
            const bool skip_b = (g.dbg & 2) != 0;                  // dbg 2: arrive without copying the weights
            auto put = [&](const uint8_t *src, uint32_t bytes) {
                mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
                if (skip_b) {
                    mbar_arrive(&b_full[rb.stage]);
                } else {
                    mbar_arrive_expect_tx(&b_full[rb.stage], bytes);
                    bulk_g2s(Bs + rb.stage * B_STAGE, src, bytes, &b_full[rb.stage]);
                }
                rb.advance();
            };
            for (int64_t i = 0; i <= T; ++i) {
                if (i < T)
                    for (int s = 0; s < half; ++s) put(g.w1img + (size_t)s * B_STAGE, B_STAGE);
                if (i >= 1)
                    for (int s = 0; s < nst2; ++s) put(g.w2img + (size_t)s * B_STAGE, min(2, g.ksteps2 - 2 * s) * (B_STAGE / 2));
                if (i < T)
                    for (int s = half; s < g.nst1; ++s) put(g.w1img + (size_t)s * B_STAGE, B_STAGE);
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem, 512);
}

// ---- weight images --------------------------------------------------------------------------
// Stage s of an image covers K = [32 s, 32 s + 32) (the last layer-2 stage may cover 16):
//   [hi: nkch k-chunks][lo: nkch k-chunks], k-chunk = 22 core matrices of 8 rows x 8 k (128 B each);
//   element (row n, k) of chunk c sits at c*2816 + (n/8)*128 + (n%8)*16 + (k%8)*2.
__global__ void tc_pack_kernel(const float *__restrict__ W, int N, int K, int ksteps, uint8_t *__restrict__ img) {
    const int nstages = (ksteps + 1) / 2;
    const int64_t total = (int64_t)nstages * 4 * NPAD * 8;       // one thread per (stage, chunk, row, 8 k) pair-of-chunks unit
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e & 7);
        const int n = (int)((e >> 3) % NPAD);
        const int c = (int)(((e >> 3) / NPAD) & 3);
        const int s = (int)(((e >> 3) / NPAD) >> 2);
        const int nkch = 2 * min(2, ksteps - 2 * s);
        if (c >= nkch) continue;
        const int k = s * KST + c * 8 + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        uint8_t *st = img + (size_t)s * B_STAGE;
        const size_t off = (size_t)c * KCH_B + (n >> 3) * 128 + (n & 7) * 16 + kk * 2;
        *reinterpret_cast<__nv_bfloat16 *>(st + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(st + (size_t)nkch * KCH_B + off) = lo;
    }
}

static int64_t image_bytes(int ksteps) { return (int64_t)((ksteps + 1) / 2) * B_STAGE; }

}  // namespace tcg

// ---- interface used by pack.cu / api.cu -------------------------------------------------------
static bool tc_dims_ok(int d_in, int d1, int d2) {
    return d_in % tcg::KST == 0 && d_in >= tcg::KST && d1 <= tcg::NPAD && d2 <= tcg::NPAD && d1 >= 1 && d2 >= 1;
}

int64_t tc_image_bytes(int d_in, int d1, int d2) {
    if (!tc_dims_ok(d_in, d1, d2)) return 0;
    return tcg::image_bytes(d_in / 16) + tcg::image_bytes(round_up(d1, 16) / 16) + 512;
}

bool tc_shape_ok(bool dplda, const PackLayout &L, bool indexed) {
    (void)indexed;
    return !dplda && tc_dims_ok(L.d_in, L.d1, L.d2) && L.tc_bytes > 0;
}

int tc_pack_nplda(const float *W1, const float *b1, const float *W2, const float *b2, const float *p_sqrt,
                  const float *q, const PackLayout &L, char *pack, cudaStream_t st) {
    (void)b1; (void)b2; (void)p_sqrt; (void)q;   // the fp32 padded vectors of the SIMT pack are shared
    if (!tc_dims_ok(L.d_in, L.d1, L.d2)) return NPLDA_OK;
    uint8_t *img1 = (uint8_t *)pack + L.tc;
    uint8_t *img2 = img1 + (tcg::image_bytes(L.d_in / 16) + 255) / 256 * 256;
    tcg::tc_pack_kernel<<<2 * sm_count(), 256, 0, st>>>(W1, L.d1, L.d_in, L.d_in / 16, img1);
    NPLDA_LAUNCH_CHECK();
    tcg::tc_pack_kernel<<<sm_count(), 256, 0, st>>>(W2, L.d2, L.d1, round_up(L.d1, 16) / 16, img2);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

int tc_pack_dplda(const float *, const float *, const float *, const float *, const PackLayout &, char *,
                  cudaStream_t) {
    return NPLDA_OK;   // DPlda scores run on the SIMT kernel (tc_shape_ok is false for it)
}

int score_tc(bool dplda, const float *x1, const float *x2, const int64_t *i1, const int64_t *i2, int64_t n_rows,
             int32_t *bad_flag, int64_t n, const PackLayout &L, const char *pack, float *scores, cudaStream_t st) {
    if (dplda || !tc_dims_ok(L.d_in, L.d1, L.d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    tcg::Args a;
    a.x1 = x1; a.x2 = x2; a.i1 = i1; a.i2 = i2; a.n_rows = n_rows; a.bad_flag = bad_flag; a.n = n;
    a.d_in = L.d_in; a.nst1 = L.d_in / tcg::KST; a.ksteps2 = round_up(L.d1, 16) / 16;
    a.w1img = (const uint8_t *)pack + L.tc;
    a.w2img = a.w1img + (tcg::image_bytes(L.d_in / 16) + 255) / 256 * 256;
    a.b1 = (const float *)(pack + L.b1); a.b2 = (const float *)(pack + L.b2);
    a.p = (const float *)(pack + L.p); a.q = (const float *)(pack + L.q);
    a.scores = scores;
    {
        const char *e = getenv("NPLDA_TC_DEBUG");
        a.dbg = e ? atoi(e) : 0;
    }
    const int64_t ntiles = (n + tcg::TP - 1) / tcg::TP;
    const int grid = (int)std::min<int64_t>(ntiles, sm_count());
    auto kern = i1 ? tcg::score_tc_kernel<true> : tcg::score_tc_kernel<false>;
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::SMEM_BYTES));
    kern<<<grid, tcg::NTHREADS, tcg::SMEM_BYTES, st>>>(a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda
