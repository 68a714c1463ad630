// K1x (tensor cores, CTA pairs): fused pairwise score kernel over a PRE-SPLIT x-vector table.
//
// The reference's callers never hold materialised [N,512] pairs: they hold a static table of x-vectors (the pickled
// dict of xvector_NeuralPlda_pytorch.py:117) and index batches (sv_trials_loaders.py:418-426).  The table is split
// ONCE into the tensor cores' operand format and this kernel scores trials (table[i1[k]], table[i2[k]]) with no
// conversion work at all.
//
// Numerics ("fp16x3"): every fp32 operand v is scaled by a power of two into fp16's normal range and split
// v 2^k = hi + lo with hi = fp16(v 2^k), lo = fp16(v 2^k - hi): two 11-bit significands, residual O(2^-22 |v|) -- against
// O(2^-17) for the bf16 split of score_tc.cu's MODE 0.  Products are evaluated as hi*hi + lo*hi + hi*lo with fp32
// accumulation in TMEM (the dropped lo*lo term is O(2^-22)), the same three MMAs per K step at the same tensor rate.
// Measured / emulated on the golden parameters: scores within 1e-6 of the fp64 value (bf16x3: 8e-6), and the training
// activations accurate enough for gradients at fp32-autograd level (2e-6 against 1e-4, tools/grad_diag.py).  The
// scales: the table by 2^kx from its absolute maximum (nplda_table_split, two passes, once per table), W1 / W2 by
// 2^gw from theirs (pack time), the layer-1 output a by 2^-ka from the bound max_j ||W1_j||_1 max|x| + max|b1| -- all
// exact powers of two undone in the epilogue, so nothing can overflow and there is no fallback path.
//
//
//   * two CTAs of a cluster (the two SMs of a TPC) execute M = 256 MMAs (tcgen05.mma.cta_group::2): each CTA holds
//     the 128 rows (64 pairs, both sides) of its own tile in shared memory and HALF of the 176 weight rows, so the
//     weight stream L2 -> shared memory and the tensor cores' operand reads of it are halved.  Measured on B200
//     (tools/ss_probe.cu): the tensor core reads shared-memory operands at ~75 B/cycle; a 1-CTA SS MMA M128 N176 K16
//     (A 4 KB + B 5.5 KB) takes 131 cycles against the 88-98 of the A-in-TMEM form, the pair form (4 + 2.75 KB) ~92.
//   * A loaders (4 warps) gather table rows with TMA tile::gather4 (cp.async.bulk.tensor.2d...gather4, four rows of
//     128 bytes per instruction) straight into the 128-byte-swizzled A operand: a row's stage is [hi k0..31 | lo k0..31],
//     one swizzle row, so the four A descriptors of a stage are the tile base + {0, 32, 64, 96} bytes.  No converter
//     warps, no tensor-memory stores.
//   * both CTAs' TMA loads complete on the LEADER's mbarriers (.cta_group::2); the leader's MMA warp issues for the
//     pair and releases stages / publishes accumulators with multicast commits; the peer's epilogue warps arrive on
//     the leader's barriers through the cluster's shared-memory window.
//   * epilogue (8 warps per CTA, its own 128 rows): identical arithmetic to score_tc.cu.
#include <algorithm>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"

namespace nplda {
namespace tcx {

using namespace tc;

constexpr int TP = 64;                          // pairs per CTA tile (128 rows); a cluster scores 128 pairs per step
constexpr int NPAD = 176;                       // MMA N: 170 padded to a multiple of 16
constexpr int NH = NPAD / 2;                    // weight rows held by one CTA of the pair
constexpr int KST = 32;                         // K per stage
constexpr int A_STAGE = 128 * 128;              // 16384 B: [128 rows][hi 32 bf16 | lo 32 bf16], 128-byte swizzle
#ifndef TCX_NA
#define TCX_NA 5
#endif
#ifndef TCX_NB
#define TCX_NB 4
#endif
constexpr int NA = TCX_NA;                      // A ring stages
constexpr int KCH_BH = (NH / 8) * 128;          // 1408 B: one 8-wide k-chunk of 88 weight rows (11 core matrices)
constexpr int B_HALF = 8 * KCH_BH;              // 11264 B: hi chunks 0-3, lo chunks 0-3 of one stage (K = 32)
constexpr int B_LINES = B_HALF / 128;           // 88 rows of the [lines x 128 B] view the weight TMA uses
constexpr int NB = TCX_NB;                      // B ring stages
constexpr int KCH_U = (128 / 8) * 128;          // 2048 B: one k-chunk of U (16 core matrices)
constexpr int U_HALF = (NPAD / 8) * KCH_U;      // 45056 B (hi or lo), K = 176

constexpr int EPI_WARPS = 8, ALOAD_WARPS = 4;
constexpr int WARP_MMA = EPI_WARPS, WARP_BLOAD = WARP_MMA + 1, WARP_ALOAD = WARP_MMA + 2;
constexpr int NTHREADS = (WARP_ALOAD + ALOAD_WARPS) * 32;   // 448

constexpr int SM_A = 0;                                   // 1024-byte aligned (swizzle atoms)
constexpr int SM_B = SM_A + NA * A_STAGE;
constexpr int SM_U = SM_B + NB * B_HALF;
constexpr int SM_PAR = SM_U + 2 * U_HALF;                 // b1, b2, P, Q (NPAD floats each)
constexpr int SM_BAR = SM_PAR + 4 * NPAD * 4;
constexpr int N_BARS = 2 * NA + 2 * NB + 8;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(SM_B % 1024 == 0 && SM_U % 128 == 0, "alignment");

struct Args {
    const int64_t *i1, *i2;
    int64_t n, n_rows;
    int nst1;               // layer-1 stages  = d_in / 32
    int ksteps2;            // layer-2 K steps = round_up(d1, 16) / 16
    const float *b1, *b2, *p, *q;   // padded to >= NPAD floats
    const float *thdr;      // table header: [0] max|x|, [1] 2^kx (the table is stored as x 2^kx), [2] 2^-kx
    const float *whdr;      // pair-image header: [0] 2^gw1, [1] 2^-gw1, [2] 2^gw2, [3] 2^-gw2, [4] max_j ||W1_j||_1, [5] max|b1|
    float *scores;
    int32_t *bad_flag;
    int variant;            // experiments (DEBUG builds, NPLDA_TCX_VARIANT): 1 = barrier waits not interleaved with the MMAs
};

struct Ring {
    uint32_t stage = 0, phase = 0;
    int n;
    __device__ explicit Ring(int n_) : n(n_) {}
    __device__ void advance() { if (++stage == (uint32_t)n) { stage = 0; phase ^= 1; } }
    __device__ void advance_by(int64_t cnt) {
        const int64_t tot = (int64_t)stage + cnt;
        stage = (uint32_t)(tot % n);
        phase ^= (uint32_t)((tot / n) & 1);
    }
};

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
score_tcx_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW1,
                 const __grid_constant__ CUtensorMap mapW2, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t *As = smem + SM_A;
    uint8_t *Bs = smem + SM_B;
    uint8_t *Us = smem + SM_U;
    float *par = reinterpret_cast<float *>(smem + SM_PAR);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *a_full = bars, *a_empty = a_full + NA, *b_full = a_empty + NA, *b_empty = b_full + NB;
    uint64_t *d_full = b_empty + NB, *d_empty = d_full + 2, *y_full = d_full + 4;
    uint64_t *u_full = d_full + 6, *u_empty = d_full + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int64_t nsuper = (g.n + 2 * TP - 1) / (2 * TP);
    const int64_t T = nsuper > cid ? (nsuper - cid + ncl - 1) / ncl : 0;       // the same in both CTAs of a pair
    auto tile_base = [&](int64_t i) { return ((cid + i * ncl) * 2 + rank) * TP; };

    // ---- one-time setup ----
    for (int i = tid; i < NPAD; i += NTHREADS) {
        par[i] = g.b1[i]; par[NPAD + i] = g.b2[i]; par[2 * NPAD + i] = g.p[i]; par[3 * NPAD + i] = g.q[i];
    }
    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&d_full[d], 1); mbar_init(&d_empty[d], 2 * EPI_WARPS); mbar_init(&y_full[d], 1); }
        mbar_init(u_full, 2 * EPI_WARPS);
        mbar_init(u_empty, 1);
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc2(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                      // barriers of both CTAs initialised before any remote arrive / TMA completion
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = make_idesc_f16(256, NPAD);
    const int nst2 = (g.ksteps2 + 1) / 2;
    // scales (exact powers of two): layer-1 accumulator = a 2^(kx + gw1); U holds a 2^-ka with |a 2^-ka| < 16384 for
    // every input the bound covers; layer-2 accumulator = (W2 a) 2^(gw2 - ka)
    const float s1 = g.thdr[2] * g.whdr[1];
    const float a_bound = g.whdr[4] * g.thdr[0] + g.whdr[5];
    const float sa = a_bound < 16384.f ? 1.f : pow2_scale_to_2p13(a_bound);
    const float c2 = g.whdr[3] / sa;

    if (warp < EPI_WARPS) {
        // =============================== EPILOGUE (this CTA's 128 rows) ===============================
        const int q = warp & 3, h = warp >> 2;                  // TMEM quadrant, 16-lane half
        const int rsub = lane >> 2, cq = lane & 3;
        const int mrow = q * 32 + h * 16 + rsub;                 // side-0 row; side 1 is mrow + 8
        const int pl = q * 16 + h * 8 + rsub;                    // pair within the tile
        const uint32_t tbase = tmem + ((uint32_t)(q * 32 + h * 16) << 16);
        uint8_t *u0 = Us + (mrow >> 3) * 128 + (mrow & 7) * 16 + cq * 4;   // + kchunk * KCH_U
        uint8_t *u1 = u0 + 128;                                            // row + 8: next core matrix
        const float2 *b1s = reinterpret_cast<const float2 *>(par);
        const float2 *b2s = reinterpret_cast<const float2 *>(par + NPAD);
        const float2 *ps = reinterpret_cast<const float2 *>(par + 2 * NPAD);
        const float2 *qs = reinterpret_cast<const float2 *>(par + 3 * NPAD);

        auto pass1 = [&](const uint32_t (&v)[8], int c0, float (&ss)[4]) {
            const float2 ba = b1s[(c0 >> 1) + cq], bb = b1s[(c0 >> 1) + 4 + cq];
            const float a00 = fmaf(__uint_as_float(v[0]), s1, ba.x), a01 = fmaf(__uint_as_float(v[1]), s1, ba.y);
            const float a10 = fmaf(__uint_as_float(v[2]), s1, ba.x), a11 = fmaf(__uint_as_float(v[3]), s1, ba.y);
            const float a02 = fmaf(__uint_as_float(v[4]), s1, bb.x), a03 = fmaf(__uint_as_float(v[5]), s1, bb.y);
            const float a12 = fmaf(__uint_as_float(v[6]), s1, bb.x), a13 = fmaf(__uint_as_float(v[7]), s1, bb.y);
            ss[0] = fmaf(a00, a00, ss[0]); ss[1] = fmaf(a01, a01, ss[1]);
            ss[0] = fmaf(a02, a02, ss[0]); ss[1] = fmaf(a03, a03, ss[1]);
            ss[2] = fmaf(a10, a10, ss[2]); ss[3] = fmaf(a11, a11, ss[3]);
            ss[2] = fmaf(a12, a12, ss[2]); ss[3] = fmaf(a13, a13, ss[3]);
            uint32_t hi, lo;
            const int kc = c0 >> 3;
            split_f16x2(a00 * sa, a01 * sa, hi, lo);
            *reinterpret_cast<uint32_t *>(u0 + kc * KCH_U) = hi;
            *reinterpret_cast<uint32_t *>(u0 + U_HALF + kc * KCH_U) = lo;
            split_f16x2(a02 * sa, a03 * sa, hi, lo);
            *reinterpret_cast<uint32_t *>(u0 + (kc + 1) * KCH_U) = hi;
            *reinterpret_cast<uint32_t *>(u0 + U_HALF + (kc + 1) * KCH_U) = lo;
            split_f16x2(a10 * sa, a11 * sa, hi, lo);
            *reinterpret_cast<uint32_t *>(u1 + kc * KCH_U) = hi;
            *reinterpret_cast<uint32_t *>(u1 + U_HALF + kc * KCH_U) = lo;
            split_f16x2(a12 * sa, a13 * sa, hi, lo);
            *reinterpret_cast<uint32_t *>(u1 + (kc + 1) * KCH_U) = hi;
            *reinterpret_cast<uint32_t *>(u1 + U_HALF + (kc + 1) * KCH_U) = lo;
        };
        auto pass2 = [&](const uint32_t (&v)[8], int c0, float r0, float r1, float (&sc)[2]) {
            const int ci = (c0 >> 1) + cq;
            const float2 ba = b2s[ci], bb = b2s[ci + 4], pa = ps[ci], pb = ps[ci + 4], qa = qs[ci], qb = qs[ci + 4];
            const float y00 = fmaf(__uint_as_float(v[0]), r0, ba.x), y01 = fmaf(__uint_as_float(v[1]), r0, ba.y);
            const float y10 = fmaf(__uint_as_float(v[2]), r1, ba.x), y11 = fmaf(__uint_as_float(v[3]), r1, ba.y);
            const float y02 = fmaf(__uint_as_float(v[4]), r0, bb.x), y03 = fmaf(__uint_as_float(v[5]), r0, bb.y);
            const float y12 = fmaf(__uint_as_float(v[6]), r1, bb.x), y13 = fmaf(__uint_as_float(v[7]), r1, bb.y);
            sc[0] = fmaf(qa.x, fmaf(y00, y00, y10 * y10), sc[0]); sc[0] = fmaf(2.f * pa.x, y00 * y10, sc[0]);
            sc[1] = fmaf(qa.y, fmaf(y01, y01, y11 * y11), sc[1]); sc[1] = fmaf(2.f * pa.y, y01 * y11, sc[1]);
            sc[0] = fmaf(qb.x, fmaf(y02, y02, y12 * y12), sc[0]); sc[0] = fmaf(2.f * pb.x, y02 * y12, sc[0]);
            sc[1] = fmaf(qb.y, fmaf(y03, y03, y13 * y13), sc[1]); sc[1] = fmaf(2.f * pb.y, y03 * y13, sc[1]);
        };

        for (int64_t i = 0; i < T; ++i) {
            const int d = (int)(i & 1);
            const uint32_t par_d = (uint32_t)((i >> 1) & 1);
            const uint32_t taddr = tbase + d * NPAD;
            // ---- layer-1 accumulator: a = D + b1, |a|, bf16 hi/lo of a -> U (normalised after layer 2) ----
            mbar_wait(&d_full[d], par_d);
            tc_fence_after();
            mbar_wait(u_empty, (uint32_t)((i & 1) ^ 1));           // layer 2 of the previous tile has read U
            float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < NPAD - 16; c0 += 32) {           // two 16-column loads per wait
                uint32_t va[8], vb[8];
                tmem_ld_16x256b_x2(taddr + c0, va);
                tmem_ld_16x256b_x2(taddr + c0 + 16, vb);
                tmem_ld_wait();
                pass1(va, c0, ss);
                pass1(vb, c0 + 16, ss);
            }
            {
                uint32_t va[8];
                tmem_ld_16x256b_x2(taddr + NPAD - 16, va);
                tmem_ld_wait();
                pass1(va, NPAD - 16, ss);
            }
            float ss0 = ss[0] + ss[1], ss1 = ss[2] + ss[3];
            ss0 += __shfl_xor_sync(0xffffffffu, ss0, 1); ss0 += __shfl_xor_sync(0xffffffffu, ss0, 2);
            ss1 += __shfl_xor_sync(0xffffffffu, ss1, 1); ss1 += __shfl_xor_sync(0xffffffffu, ss1, 2);
            const float r0 = c2 / fmaxf(sqrtf(ss0), 1e-12f);       // F.normalize eps (models.py:368), times the layer-2 scale
            const float r1 = c2 / fmaxf(sqrtf(ss1), 1e-12f);
            fence_proxy_async();      // U is read by tcgen05.mma (async proxy)
            tc_fence_before();        // our TMEM reads of D are done before Y overwrites it
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(u_full, 0);          // one arrival per warp, on the leader's barrier
            // ---- layer-2 accumulator: y = Y / |a| + b2, pair score ----
            mbar_wait(&y_full[d], par_d);
            tc_fence_after();
            float sc[2] = {0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < NPAD - 16; c0 += 32) {
                uint32_t va[8], vb[8];
                tmem_ld_16x256b_x2(taddr + c0, va);
                tmem_ld_16x256b_x2(taddr + c0 + 16, vb);
                tmem_ld_wait();
                pass2(va, c0, r0, r1, sc);
                pass2(vb, c0 + 16, r0, r1, sc);
            }
            {
                uint32_t va[8];
                tmem_ld_16x256b_x2(taddr + NPAD - 16, va);
                tmem_ld_wait();
                pass2(va, NPAD - 16, r0, r1, sc);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&d_empty[d], 0);     // D buffer free before the shuffles / store
            float s = sc[0] + sc[1];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const int64_t pr = tile_base(i) + pl;
            if (cq == 0 && pr < g.n) g.scores[pr] = s;
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA ISSUER (leader CTA) ===============================
        Ring ra(NA), rb(NB);
        if (rank == 0) {
            const uint32_t a_base = smem_addr(As), b_base = smem_addr(Bs), u_base = smem_addr(Us);
            const int half = g.nst1 / 2;
            // One stage: K = 32 as two K = 16 steps, each hi*Whi + lo*Whi + hi*Wlo.  A descriptors: tile base + 0 / 32 B
            // (hi, steps 0 / 1) and + 64 / 96 B (lo); B: chunks 0-3 hi, 4-7 lo of this CTA's half, two chunks per step.
            // The MMA queue is shallow: an issue stalls until the pipe accepts it, so whatever the warp executes
            // between two issues runs under the previous MMA.  The barrier waits of stage s + 1 (~100 cycles each even
            // when already satisfied) therefore sit between the two K steps of stage s.
            auto layer1 = [&](uint32_t dcol, int s_begin, int s_end) {
                if (s_begin >= s_end) return;
                mbar_wait(&a_full[ra.stage], ra.phase);
                mbar_wait(&b_full[rb.stage], rb.phase);
                for (int s = s_begin; s < s_end; ++s) {
                    tc_fence_after();
                    const uint64_t ad = make_smem_desc_sw128(a_base + ra.stage * A_STAGE);
                    const uint64_t bd = make_smem_desc(b_base + rb.stage * B_HALF, KCH_BH, 128);
                    Ring na = ra, nb = rb;
                    na.advance();
                    nb.advance();
                    // probes of the next stage's barriers: issued before this stage's MMAs, looked at after the first K step
                    // (an already-complete try_wait still takes ~90 cycles to answer)
                    uint32_t ok_a = 1, ok_b = 1;
                    if (s + 1 < s_end && !(g.variant & 1)) { ok_a = mbar_try(&a_full[na.stage], na.phase); ok_b = mbar_try(&b_full[nb.stage], nb.phase); }
                    // one issue block per stage (the probes made the mid-stage waits unnecessary)
                    if (elect_one()) {
                        mma2_ss(dcol, ad, bd, IDESC, s != 0);
                        mma2_ss(dcol, ad + 4, bd, IDESC, 1);
                        mma2_ss(dcol, ad, bd + ((4 * KCH_BH) >> 4), IDESC, 1);
                        mma2_ss(dcol, ad + 2, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                        mma2_ss(dcol, ad + 6, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                        mma2_ss(dcol, ad + 2, bd + ((6 * KCH_BH) >> 4), IDESC, 1);
                        mma2_commit_mc(&a_empty[ra.stage], 3);
                        mma2_commit_mc(&b_empty[rb.stage], 3);
                    }
                    __syncwarp();
                    if (!ok_a) mbar_wait(&a_full[na.stage], na.phase);
                    if (!ok_b) mbar_wait(&b_full[nb.stage], nb.phase);
                    if (s + 1 < s_end && (g.variant & 1)) {
                        mbar_wait(&a_full[na.stage], na.phase);
                        mbar_wait(&b_full[nb.stage], nb.phase);
                    }
                    ra = na;
                    rb = nb;
                }
            };
            for (int64_t i = 0; i <= T; ++i) {
                const uint32_t dcol_i = tmem + (uint32_t)(i & 1) * NPAD;
                if (i < T) {
                    mbar_wait(&d_empty[i & 1], (uint32_t)(((i >> 1) & 1) ^ 1));
                    tc_fence_after();
                    layer1(dcol_i, 0, half);
                }
                if (i >= 1) {      // layer 2 of tile i - 1: Y = U W2^T into the accumulator that held D
                    const int64_t j = i - 1;
                    const int d = (int)(j & 1);
                    const uint32_t dcol = tmem + d * NPAD;
                    mbar_wait(u_full, (uint32_t)(j & 1));
                    tc_fence_after();
                    mbar_wait(&b_full[rb.stage], rb.phase);
                    for (int ks = 0; ks < g.ksteps2; ks += 2) {
                        tc_fence_after();
                        const uint64_t bd = make_smem_desc(b_base + rb.stage * B_HALF, KCH_BH, 128);
                        const uint64_t uhi = make_smem_desc(u_base + ks * 2 * KCH_U, KCH_U, 128);
                        const uint64_t ulo = make_smem_desc(u_base + U_HALF + ks * 2 * KCH_U, KCH_U, 128);
                        Ring nb = rb;
                        nb.advance();
                        const uint32_t ok_b = ks + 2 < g.ksteps2 ? mbar_try(&b_full[nb.stage], nb.phase) : 1u;
                        if (elect_one()) {
                            mma2_ss(dcol, uhi, bd, IDESC, ks != 0);
                            mma2_ss(dcol, ulo, bd, IDESC, 1);
                            mma2_ss(dcol, uhi, bd + ((4 * KCH_BH) >> 4), IDESC, 1);
                            if (ks + 1 < g.ksteps2) {
                                const uint64_t uhi1 = uhi + ((2 * KCH_U) >> 4), ulo1 = ulo + ((2 * KCH_U) >> 4);
                                mma2_ss(dcol, uhi1, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                                mma2_ss(dcol, ulo1, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                                mma2_ss(dcol, uhi1, bd + ((6 * KCH_BH) >> 4), IDESC, 1);
                            }
                            mma2_commit_mc(&b_empty[rb.stage], 3);
                        }
                        __syncwarp();
                        if (!ok_b) mbar_wait(&b_full[nb.stage], nb.phase);
                        __syncwarp();
                        rb = nb;
                    }
                    if (elect_one()) {
                        mma2_commit_mc(&y_full[d], 3);
                        mma2_commit_mc(u_empty, 3);
                    }
                    __syncwarp();
                }
                if (i < T) {
                    layer1(dcol_i, half, g.nst1);
                    if (elect_one()) mma2_commit_mc(&d_full[i & 1], 3);
                    __syncwarp();
                }
            }
        } else {
            ra.advance_by(T * g.nst1);
            rb.advance_by(T * (g.nst1 + nst2));
        }
        // Drain (both CTAs): the multicast arrivals of the last commits on a_empty / b_empty / u_empty are not waited
        // for by any producer; they must land before a CTA exits and its shared memory is handed to the next one.
        if (T > 0) {
            for (int k = 0; k < NA && k < T * g.nst1; ++k) { mbar_wait(&a_empty[ra.stage], ra.phase ^ 1); ra.advance(); }
            for (int k = 0; k < NB; ++k) { mbar_wait(&b_empty[rb.stage], rb.phase ^ 1); rb.advance(); }
            mbar_wait(u_empty, (uint32_t)((T & 1) ^ 1));
        }
    } else if (warp == WARP_BLOAD) {
        // =============================== B LOADER (this CTA's half of every weight stage) ===============================
        if (lane == 0) {
            Ring rb(NB);
            const int half = g.nst1 / 2;
            auto put = [&](const CUtensorMap *map, int stage_idx) {
                mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
                if (rank == 0) mbar_arrive_expect_tx(&b_full[rb.stage], 2 * B_HALF);     // both halves post here
                tma_load_2d_pair(Bs + rb.stage * B_HALF, map, 0, (stage_idx * 2 + (int)rank) * B_LINES, &b_full[rb.stage]);
                rb.advance();
            };
            for (int64_t i = 0; i <= T; ++i) {
                if (i < T)
                    for (int s = 0; s < half; ++s) put(&mapW1, s);
                if (i >= 1)
                    for (int s = 0; s < nst2; ++s) put(&mapW2, s);
                if (i < T)
                    for (int s = half; s < g.nst1; ++s) put(&mapW1, s);
            }
        }
    } else {
        // =============================== A LOADERS (TMA row gather) ===============================
        // Warp w covers tile rows [32 w, 32 w + 32): lane = row.  In every 16-row group rows 0-7 are side 0 and rows
        // 8-15 side 1 of the same 8 pairs (the epilogue's tcgen05.ld.16x256b gives a thread rows r and r + 8).
        const int w = warp - WARP_ALOAD;
        const int R = 32 * w + lane;
        const int pair_in_tile = (R >> 4) * 8 + (R & 7);
        const bool side1 = (R >> 3) & 1;
        Ring ra(NA);
        for (int64_t i = 0; i < T; ++i) {
            const int64_t p = tile_base(i) + pair_in_tile;
            int64_t rid = 0;
            if (p < g.n) {
                rid = side1 ? g.i2[p] : g.i1[p];
                if (rid < 0 || rid >= g.n_rows) { *g.bad_flag = 1; rid = 0; }      // reported; the score is garbage
            }
            const int q4 = lane & ~3;
            const int r0 = __shfl_sync(0xffffffffu, (int)rid, q4), r1 = __shfl_sync(0xffffffffu, (int)rid, q4 + 1);
            const int r2 = __shfl_sync(0xffffffffu, (int)rid, q4 + 2), r3 = __shfl_sync(0xffffffffu, (int)rid, q4 + 3);
            for (int s = 0; s < g.nst1; ++s) {
                mbar_wait(&a_empty[ra.stage], ra.phase ^ 1);
                if (rank == 0 && w == 0 && lane == 0) mbar_arrive_expect_tx(&a_full[ra.stage], 2 * A_STAGE);
                if ((lane & 3) == 0)
                    tma_gather4_pair(As + ra.stage * A_STAGE + R * 128, &mapA, s * 64, r0, r1, r2, r3, &a_full[ra.stage]);
                __syncwarp();
                ra.advance();
            }
        }
    }

    // ---- teardown ----
    __syncwarp();                 // single-lane roles: the whole warp arrives at the (aligned) cluster barrier together
    tc_fence_before();
    cluster_sync_all();           // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (warp == WARP_MMA) tmem_dealloc2(tmem, 512);
}

// ---- table split ------------------------------------------------------------------------------
// split[row][stage s][hi k = 32 s .. 32 s + 31 (fp16) | lo (fp16)] of x 2^kx: d_in * 4 bytes per row, as many as the fp32
// row; then a 128-byte header {max|x|, 2^kx, 2^-kx}.  Two passes, once per table: absolute maximum, then the split.
__global__ void __launch_bounds__(256) table_absmax_kernel(const float *__restrict__ table, int64_t count, uint32_t *__restrict__ hdr) {
    float m = 0.f;
    for (int64_t e = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; e < count; e += (int64_t)gridDim.x * blockDim.x * 4) {
        const float4 v = *reinterpret_cast<const float4 *>(table + e);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(hdr, __float_as_uint(m));     // non-negative floats order like their bit patterns
}

__global__ void __launch_bounds__(256) table_split_kernel(const float *__restrict__ table, int64_t n_rows, int d_in,
                                                          uint8_t *__restrict__ out, float *__restrict__ hdr) {
    const float scale = pow2_scale_to_2p13(hdr[0]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[1] = scale; hdr[2] = 1.f / scale; }
    const int per_row = d_in / 8;                                   // 8 consecutive elements per thread
    const int64_t total = n_rows * per_row;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / per_row;
        const int k0 = (int)(e % per_row) * 8;
        const float4 a = *reinterpret_cast<const float4 *>(table + row * d_in + k0);
        const float4 b = *reinterpret_cast<const float4 *>(table + row * d_in + k0 + 4);
        uint4 hi, lo;
        split_f16x2(a.x * scale, a.y * scale, hi.x, lo.x); split_f16x2(a.z * scale, a.w * scale, hi.y, lo.y);
        split_f16x2(b.x * scale, b.y * scale, hi.z, lo.z); split_f16x2(b.z * scale, b.w * scale, hi.w, lo.w);
        uint8_t *st = out + row * (int64_t)d_in * 4 + (k0 / KST) * 128 + (k0 % KST) * 2;
        *reinterpret_cast<uint4 *>(st) = hi;
        *reinterpret_cast<uint4 *>(st + 64) = lo;
    }
}

// ---- pair weight images -----------------------------------------------------------------------
// Stage s (K = [32 s, 32 s + 32)) = [half 0][half 1], a half = weight rows [88 h, 88 h + 88) as 8 chunks
// [hi k 0-7][hi 8-15][hi 16-23][hi 24-31][lo x 4], a chunk = 11 core matrices of 8 rows x 8 k (128 B each);
// fp16 hi / lo of W 2^gw (*scale).
__global__ void pair_pack_kernel(const float *__restrict__ W, int N, int K, int nstages, const float *__restrict__ scale,
                                 uint8_t *__restrict__ img) {
    const float sc = scale[0];
    const int64_t total = (int64_t)nstages * NPAD * KST;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % KST);
        const int n = (int)((e / KST) % NPAD);
        const int s = (int)(e / ((int64_t)KST * NPAD));
        const int k = s * KST + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * K + k] * sc : 0.f;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int hsel = n / NH, nn = n % NH;
        uint8_t *st = img + ((size_t)s * 2 + hsel) * B_HALF;
        const size_t off = (size_t)(kk >> 3) * KCH_BH + (nn >> 3) * 128 + (nn & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__half *>(st + off) = hi;
        *reinterpret_cast<__half *>(st + 4 * KCH_BH + off) = lo;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// split table as [n_rows][d_in * 2] u16 (hi / lo interleaved per stage); box = one stage of one row, gather4 takes four
static bool make_table_map(CUtensorMap *m, const void *split, int64_t n_rows, int d_in) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)d_in * 2, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_in * 4};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void *)split, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// weight image as [lines][32 u32] (128-byte lines); box = one CTA's half of a stage
static bool make_image_map(CUtensorMap *m, const void *img, int nstages) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {32, (cuuint64_t)nstages * 2 * B_LINES};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, B_LINES}, es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *)img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tcx

// ---- interface used by pack.cu / api.cu -------------------------------------------------------
static bool tcx_dims_ok(int d_in, int d1, int d2) {
    return d_in % tcx::KST == 0 && d_in >= 2 * tcx::KST && d1 <= tcx::NPAD && d2 <= tcx::NPAD && d1 >= 1 && d2 >= 1;
}
static int64_t pair_image_bytes(int nstages) { return (int64_t)nstages * 2 * tcx::B_HALF; }
static int tcx_nst2(int d1) { return (round_up(d1, 16) / 16 + 1) / 2; }

int64_t tcx_image_bytes(int d_in, int d1, int d2) {
    if (!tcx_dims_ok(d_in, d1, d2)) return 0;
    return (pair_image_bytes(d_in / tcx::KST) + 255) / 256 * 256 + (pair_image_bytes(tcx_nst2(d1)) + 255) / 256 * 256 + 512;
}

// pair images of a NeuralPlda pack (flag NPLDA_PACK_PAIR): [W1: d_in / 32 stages][W2: ceil(ksteps2 / 2) stages]
const float *tc_hdr16(const PackLayout &L, const char *pack);   // score_tc.cu: weight scales + layer-1 bound (tc_scales_kernel)

int tcx_pack_nplda(const float *W1, const float *b1, const float *W2, const PackLayout &L, char *pack, cudaStream_t st) {
    if (!tcx_dims_ok(L.d_in, L.d1, L.d2) || L.tcx_bytes <= 0) return NPLDA_OK;
    uint8_t *img1 = (uint8_t *)pack + L.tcx;
    uint8_t *img2 = img1 + (pair_image_bytes(L.d_in / tcx::KST) + 255) / 256 * 256;
    (void)b1;
    const float *hdr = tc_hdr16(L, pack);            // filled by the tensor-core pack that ran just before on this stream
    tcx::pair_pack_kernel<<<sm_count(), 256, 0, st>>>(W1, L.d1, L.d_in, L.d_in / tcx::KST, hdr, img1);
    NPLDA_LAUNCH_CHECK();
    tcx::pair_pack_kernel<<<sm_count() / 2, 256, 0, st>>>(W2, L.d2, L.d1, tcx_nst2(L.d1), hdr + 2, img2);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

int table_split(const float *table, int64_t n_rows, int d_in, void *split, cudaStream_t st) {
    if (d_in % tcx::KST != 0 || d_in < tcx::KST) return NPLDA_ERR_UNSUPPORTED_DIM;
    const int64_t total = n_rows * (d_in / 8);
    const int grid = (int)std::min<int64_t>((total + 255) / 256, 16 * (int64_t)sm_count());
    float *hdr = (float *)((uint8_t *)split + n_rows * (int64_t)d_in * 4);
    NPLDA_CUDA_TRY(cudaMemsetAsync(hdr, 0, 128, st));
    tcx::table_absmax_kernel<<<grid, 256, 0, st>>>(table, n_rows * (int64_t)d_in, (uint32_t *)hdr);
    NPLDA_LAUNCH_CHECK();
    tcx::table_split_kernel<<<grid, 256, 0, st>>>(table, n_rows, d_in, (uint8_t *)split, hdr);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

int score_tcx(const void *split, int64_t n_rows, const int64_t *i1, const int64_t *i2, int64_t n, const PackLayout &L,
              const char *pack, float *scores, int32_t *bad_flag, cudaStream_t st) {
    if (!tcx_dims_ok(L.d_in, L.d1, L.d2) || L.tcx_bytes <= 0) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n_rows >= (int64_t)1 << 31 || n_rows <= 0) return NPLDA_ERR_BAD_ARG;
    const uint8_t *img1 = (const uint8_t *)pack + L.tcx;
    const uint8_t *img2 = img1 + (pair_image_bytes(L.d_in / tcx::KST) + 255) / 256 * 256;
    CUtensorMap mA, mW1, mW2;
    if (!tcx::make_table_map(&mA, split, n_rows, L.d_in) || !tcx::make_image_map(&mW1, img1, L.d_in / tcx::KST) ||
        !tcx::make_image_map(&mW2, img2, tcx_nst2(L.d1)))
        return NPLDA_ERR_NO_DEVICE;
    tcx::Args a;
    a.i1 = i1; a.i2 = i2; a.n = n; a.n_rows = n_rows;
    a.nst1 = L.d_in / tcx::KST; a.ksteps2 = round_up(L.d1, 16) / 16;
    a.b1 = (const float *)(pack + L.b1); a.b2 = (const float *)(pack + L.b2);
    a.p = (const float *)(pack + L.p); a.q = (const float *)(pack + L.q);
    a.thdr = (const float *)((const uint8_t *)split + n_rows * (int64_t)L.d_in * 4);
    a.whdr = tc_hdr16(L, pack);
    a.scores = scores; a.bad_flag = bad_flag;
    a.variant = 0;
#ifdef NPLDA_DEBUG_SWITCHES
    { static const int v = getenv("NPLDA_TCX_VARIANT") ? atoi(getenv("NPLDA_TCX_VARIANT")) : 0; a.variant = v; }
#endif
    const int64_t nsuper = (n + 2 * tcx::TP - 1) / (2 * tcx::TP);
    const int grid = 2 * (int)std::min<int64_t>(nsuper, sm_count() / 2);
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(tcx::score_tcx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcx::SMEM_BYTES));
    tcx::score_tcx_kernel<<<grid, tcx::NTHREADS, tcx::SMEM_BYTES, st>>>(mA, mW1, mW2, a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda
