// K1p (tensor cores, CTA pairs): the fused pairwise score kernel over MATERIALISED fp32 pairs -- NeuralPlda.forward
// (models.py:378-382 -> :366-376) for x1, x2 of shape [N, 512], the layout BASELINE configs[1] is quoted on.
//
// Same arithmetic as score_tc.cu's MODE 0 ("bf16x3": every fp32 operand split x = hi + lo into two bf16 halves, products
// as hi*hi + lo*hi + hi*lo with fp32 accumulation in tensor memory, any fp32 range), same tile and epilogue: the scores
// are bit-identical.  What changes is who shares what.  score_tc.cu runs one CTA per SM and every CTA streams the whole
// weight image (484 KB per 64-pair tile).  Here two CTAs of a cluster (the two SMs of a TPC) execute M = 256 MMAs
// (tcgen05.mma.cta_group::2): each CTA converts its own 128 rows into ITS tensor memory and holds HALF of the 176 weight
// rows, so
//   * the weight stream L2 -> shared memory (7.5 -> 3.75 GB per 1 M pairs) and the tensor cores' operand reads of it are
//     halved,
//   * a weight stage is 11 KB instead of 22.5 KB: five of them fit next to four 16 KB x stages where three did.
// Measured: 1.12-1.20 ms per 1 M pairs against 1.27-1.36 ms for the one-CTA kernel on the same boxes.
// Roles per CTA (864 threads): X loader (TMA boxes of fp32 x into a swizzled ring, local barriers), 2 x 8 converter
// warps (shared memory -> registers -> bf16 hi/lo -> tcgen05.st into a 5-stage A ring in tensor memory; they arrive on
// the LEADER's a_full barriers through the cluster's shared-memory window), B loader (this CTA's half of every weight
// stage, completing on the leader's barrier), MMA warp (leader only: issues for the pair, releases stages and
// publishes accumulators with multicast commits), 8 epilogue warps (this CTA's 128 rows, as in score_tc.cu).
#include <algorithm>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"

namespace nplda {
namespace tcp {

using namespace tc;

constexpr int TP = 64;                          // pairs per CTA tile (128 rows); a cluster scores 128 pairs per step
constexpr int NPAD = 176;                       // MMA N: 170 padded to a multiple of 16
constexpr int NH = NPAD / 2;                    // weight rows held by one CTA of the pair
constexpr int KST = 32;                         // K per stage (two MMA K steps)
constexpr int X_BOX = TP * KST * 4;             // 8192 B: [64 rows x 32 floats], 128-byte swizzle
constexpr int X_STAGE = 2 * X_BOX;              // x1 box + x2 box
#ifndef TCP_NX
#define TCP_NX 4
#endif
#ifndef TCP_NB
#define TCP_NB 5
#endif
constexpr int NX = TCP_NX;                      // x ring stages
constexpr int NB = TCP_NB;                      // B ring stages
constexpr int KCH_BH = (NH / 8) * 128;          // 1408 B: one 8-wide k-chunk of 88 weight rows (11 core matrices)
constexpr int B_HALF = 8 * KCH_BH;              // 11264 B: hi chunks 0-3, lo chunks 0-3 of one stage (K = 32)
constexpr int B_LINES = B_HALF / 128;           // 88 rows of the [lines x 128 B] view the weight TMA uses
constexpr int KCH_U = (128 / 8) * 128;          // 2048 B: one k-chunk of U (16 core matrices)
constexpr int U_HALF = (NPAD / 8) * KCH_U;      // 45056 B (hi or lo), K = 176
constexpr int NA = 5;                           // A ring stages in tensor memory
constexpr int A_COL0 = 2 * NPAD;                // TMEM columns: D0 [0,176) D1 [176,352) A ring [352,512)
constexpr int A_STAGE_COLS = 32;                // per stage: hi k0-15 | hi k16-31 | lo k0-15 | lo k16-31, 8 columns each

constexpr int EPI_WARPS = 8, CONV_WARPS = 8, CONV_SETS = 2;   // converter sets alternate stages
constexpr int WARP_MMA = EPI_WARPS + CONV_SETS * CONV_WARPS, WARP_BLOAD = WARP_MMA + 1, WARP_XLOAD = WARP_MMA + 2;
constexpr int NTHREADS = (WARP_XLOAD + 1) * 32;  // 864

constexpr int SM_X = 0;                                   // 1024-byte aligned (swizzle atoms)
constexpr int SM_B = SM_X + NX * X_STAGE;
constexpr int SM_U = SM_B + NB * B_HALF;
constexpr int SM_PAR = SM_U + 2 * U_HALF;                 // b1, b2, P, Q (NPAD floats each)
constexpr int SM_BAR = SM_PAR + 4 * NPAD * 4;
constexpr int N_BARS = 2 * NX + 2 * NA + 2 * NB + 8;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(SM_B % 1024 == 0 && SM_U % 128 == 0, "alignment");
static_assert(NX % 2 == 0 && NA % 2 == 1 && NA >= 3, "each converter set owns the x slots of its parity; see the converters");

struct Args {
    int64_t n;
    int nst1;               // layer-1 stages  = d_in / 32
    int ksteps2;            // layer-2 K steps = round_up(d1, 16) / 16
    const float *b1, *b2, *p, *q;   // padded to >= NPAD floats
    float *scores;
    const float *hdrm;      // MODE 1: hdrm[2] != 0 when this pack built the mixed images (flag NPLDA_PACK_MIXED); else every tile is flagged
    const float *hdr16;     // MODE 1: {2^gw1, 2^-gw1, ...} of the pack (tc_scales_kernel): max|W1| 2^gw1 is in [8192, 16384)
    int *guard;             // MODE 1: guard[0] is set when an input leaves the range the mode covers; MODE 0 with guard != nullptr:
                            //         run only if guard[0] is set (fallback pass behind a MODE 1 launch), then clear it
    int dbg;                // TCP_DBG builds (bottleneck experiments, env NPLDA_TCP_DEBUG): 4 no MMAs, 8 no TMEM stores, 16 no LDS / conversion, 32 idle epilogue
    int *trace;             // TCP_TRACE builds: host-mapped progress table [2 CTAs][32 warps] (post-mortem of a protocol hang)
};

#ifdef TCP_DBG
#define DBG(bit) ((g.dbg & (bit)) != 0)
#else
#define DBG(bit) false
#endif

#ifdef TCP_TRACE
#define TWAIT(bar, par, code) do { if (g.trace && lane == 0 && blockIdx.x < 2) { ((volatile int *)g.trace)[blockIdx.x * 32 + warp] = (code); __threadfence_system(); } \
                                   mbar_wait(bar, par); \
                                   if (g.trace && lane == 0 && blockIdx.x < 2) { ((volatile int *)g.trace)[blockIdx.x * 32 + warp] = (code) | 0x10000; } } while (0)
#define TMARK(code) do { if (g.trace && lane == 0 && blockIdx.x < 2) { ((volatile int *)g.trace)[blockIdx.x * 32 + warp] = (code); __threadfence_system(); } } while (0)
#else
#define TWAIT(bar, par, code) mbar_wait(bar, par)
#define TMARK(code) do { } while (0)
#endif

// Cycle accounting (TCP_PROF builds): every warp charges the time since its previous mark to a bucket; the warps of
// cluster 0 write [CTA][warp][8 buckets] (long long) to the host-mapped trace buffer when they finish.
#ifdef TCP_PROF
#define PMARK(b) do { const long long _t = clock64(); pacc[b] += _t - ptime; ptime = _t; } while (0)
#define PFLUSH() do { if (g.trace && lane == 0 && blockIdx.x < 2) for (int _b = 0; _b < 8; ++_b) ((long long *)g.trace)[(blockIdx.x * 32 + warp) * 8 + _b] = pacc[_b]; } while (0)
#else
#define PMARK(b) do { } while (0)
#define PFLUSH() do { } while (0)
#endif

__device__ __forceinline__ int pair_of_row(int rsub) { return ((rsub & 1) << 2) | (rsub >> 1); }

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

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
        : "memory");
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

// kind::f8f6f4 pair MMA with the A operand in tensor memory (four e4m3 per 32-bit cell), K = 32 per instruction
__device__ __forceinline__ void mma2_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {   // a -> low half
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {   // a -> byte 0
    uint32_t r;
    asm("{\n.reg .b16 lo, hi;\ncvt.rn.satfinite.e4m3x2.f32 lo, %2, %1;\ncvt.rn.satfinite.e4m3x2.f32 hi, %4, %3;\n"
        "mov.b32 %0, {lo, hi};\n}\n" : "=r"(r) : "f"(a), "f"(b), "f"(c), "f"(d));
    return r;
}
// D fp32, A/B format code 0 (kind::f16: fp16; kind::f8f6f4: e4m3), both K-major
__host__ __device__ constexpr uint32_t make_idesc_fmt0(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MODE 0: bf16x3 on both layers.  MODE 1 ("mixed"): layer 1 as fp16(2^9 x) fp16(W') + e4m3(residual) e4m3(W'h) + e4m3(x)
// e4m3(2^9 W'l) on one accumulator -- four MMAs per stage instead of six (an e4m3 MMA covers K = 32) -- with W' = W1 2^gw,
// max|W'| in [32, 64); layer 2 stays bf16x3.  Inputs outside the range the e4m3 terms cover raise the guard and the
// MODE 0 launch that follows on the stream recomputes the call (see score_tc.cu's MODE 1 for the error analysis).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
score_tcp_kernel(const __grid_constant__ CUtensorMap mapX1, const __grid_constant__ CUtensorMap mapX2,
                 const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 1 KB aligned
    uint8_t *Xs = smem + SM_X;
    uint8_t *Bs = smem + SM_B;
    uint8_t *Us = smem + SM_U;
    float *par = reinterpret_cast<float *>(smem + SM_PAR);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *x_full = bars, *x_empty = x_full + NX, *a_full = x_empty + NX, *a_empty = a_full + NA;
    uint64_t *b_full = a_empty + NA, *b_empty = b_full + NB;
    uint64_t *d_full = b_empty + NB, *d_empty = d_full + 2, *y_full = d_full + 4;
    uint64_t *u_full = d_full + 6, *u_empty = d_full + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int64_t cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
    const int64_t nsuper = (g.n + 2 * TP - 1) / (2 * TP);
    const int64_t T = nsuper > cid ? (nsuper - cid + ncl - 1) / ncl : 0;       // the same in both CTAs of a pair
    auto tile_base = [&](int64_t i) { return ((cid + i * ncl) * 2 + rank) * TP; };
    // fallback pass behind a MODE 1 launch: nothing to do unless its range guard fired (every CTA reads the flag here;
    // the last CTA to finish clears it, so no CTA can see it cleared)
    if (MODE == 0 && g.guard != nullptr && *reinterpret_cast<volatile int *>(g.guard) == 0) return;

    // ---- one-time setup ----
    for (int i = tid; i < NPAD; i += NTHREADS) {
        par[i] = g.b1[i]; par[NPAD + i] = g.b2[i]; par[2 * NPAD + i] = g.p[i]; par[3 * NPAD + i] = g.q[i];
    }
    if (tid == 0) {
        for (int s = 0; s < NX; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], CONV_WARPS); }
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 2 * CONV_WARPS); mbar_init(&a_empty[s], 1); }   // a_full: both CTAs' converters (leader's copy is used)
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
    constexpr uint32_t IDESC = make_idesc_bf16(256, NPAD);        // layer 2, and layer 1 of MODE 0
    constexpr uint32_t IDESC0 = make_idesc_fmt0(256, NPAD);       // MODE 1 layer 1: fp16 x fp16 and e4m3 x e4m3
    // MODE 1: the accumulator holds (2^9 x) (2^gm W) with 2^gm = 2^gw1 / 256 (max|W| 2^gm in [32, 64))
    const float s1 = MODE == 1 ? g.hdr16[1] * 0.5f : 1.f;
    const bool img_ok = MODE != 1 || g.hdrm[2] != 0.f;
#ifdef TCP_PROF
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ptime = clock64();
#endif

    if (warp < EPI_WARPS) {
        // =============================== EPILOGUE (this CTA's 128 rows) ===============================
        const int q = warp & 3, h = warp >> 2;                  // TMEM quadrant, 16-lane half
        const int rsub = lane >> 2, cq = lane & 3;
        const int mrow = q * 32 + h * 16 + rsub;                 // side-0 row; side 1 is mrow + 8
        const int pl = q * 16 + h * 8 + pair_of_row(rsub);       // pair within the tile (the converters' row permutation)
        const uint32_t tbase = tmem + ((uint32_t)(q * 32 + h * 16) << 16);
        uint8_t *u0 = Us + (mrow >> 3) * 128 + (mrow & 7) * 16 + cq * 4;   // + kchunk * KCH_U
        uint8_t *u1 = u0 + 128;                                            // row + 8: next core matrix
        const float2 *b1s = reinterpret_cast<const float2 *>(par);
        const float2 *b2s = reinterpret_cast<const float2 *>(par + NPAD);
        const float2 *ps = reinterpret_cast<const float2 *>(par + 2 * NPAD);
        const float2 *qs = reinterpret_cast<const float2 *>(par + 3 * NPAD);

        auto pass1 = [&](const uint32_t (&v)[8], int c0, float (&ss)[4]) {
            const float2 ba = b1s[(c0 >> 1) + cq], bb = b1s[(c0 >> 1) + 4 + cq];
            const float a00 = fmaf(__uint_as_float(v[0]), s1, ba.x), a01 = fmaf(__uint_as_float(v[1]), s1, ba.y);   // MODE 0: s1 = 1
            const float a10 = fmaf(__uint_as_float(v[2]), s1, ba.x), a11 = fmaf(__uint_as_float(v[3]), s1, ba.y);
            const float a02 = fmaf(__uint_as_float(v[4]), s1, bb.x), a03 = fmaf(__uint_as_float(v[5]), s1, bb.y);
            const float a12 = fmaf(__uint_as_float(v[6]), s1, bb.x), a13 = fmaf(__uint_as_float(v[7]), s1, bb.y);
            ss[0] = fmaf(a00, a00, ss[0]); ss[1] = fmaf(a01, a01, ss[1]);
            ss[0] = fmaf(a02, a02, ss[0]); ss[1] = fmaf(a03, a03, ss[1]);
            ss[2] = fmaf(a10, a10, ss[2]); ss[3] = fmaf(a11, a11, ss[3]);
            ss[2] = fmaf(a12, a12, ss[2]); ss[3] = fmaf(a13, a13, ss[3]);
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
            PMARK(5);
            TWAIT(&d_full[d], par_d, 0x101);
            tc_fence_after();
            PMARK(0);
            TWAIT(u_empty, (uint32_t)((i & 1) ^ 1), 0x102);
            PMARK(1);           // layer 2 of the previous tile has read U
            float ss[4] = {0.f, 0.f, 0.f, 0.f};
            const int cend = DBG(32) ? 0 : NPAD - 16;
#pragma unroll 1
            for (int c0 = 0; c0 < cend; c0 += 32) {           // two 16-column loads per wait
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
            const float r0 = 1.f / fmaxf(sqrtf(ss0), 1e-12f);      // F.normalize eps (models.py:368)
            const float r1 = 1.f / fmaxf(sqrtf(ss1), 1e-12f);
            if (MODE == 1) {
                // upper range guard: an input with |x| >= 128 overflowed fp16(2^9 x) to inf and arrives here as inf / NaN
                // (a NaN input too: the fallback pass then gives the NaN score the reference gives)
                if (!(ss0 < 1.0e30f && ss1 < 1.0e30f) && tile_base(i) + pl < g.n) *reinterpret_cast<volatile int *>(g.guard) = 1;
            }
            fence_proxy_async();      // U is read by tcgen05.mma (async proxy)
            tc_fence_before();        // our TMEM reads of D are done before Y overwrites it
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(u_full, 0);          // one arrival per warp, on the leader's barrier
            PMARK(2);
            // ---- layer-2 accumulator: y = Y / |a| + b2, pair score ----
            TWAIT(&y_full[d], par_d, 0x103);
            tc_fence_after();
            PMARK(3);
            float sc[2] = {0.f, 0.f};
#pragma unroll 1
            for (int c0 = 0; c0 < cend; c0 += 32) {
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
            PMARK(4);
            float s = sc[0] + sc[1];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            const int64_t pr = tile_base(i) + pl;
            if (cq == 0 && pr < g.n) g.scores[pr] = s;
        }
        PFLUSH();
    } else if (warp < WARP_MMA) {
        // =============================== CONVERTERS (this CTA's 128 rows) ===============================
        // Two sets of 8 warps; set s handles stages it = s, s + 2, ...: one warp's chain of waits, LDS, conversion, TMEM
        // store and store-wait then spans two MMA stage times.
        const int cset = (warp - EPI_WARPS) >> 3;
        const int q = warp & 3, h = ((warp - EPI_WARPS) >> 2) & 1;
        const int rsub = lane >> 2, cq = lane & 3;
        // MMA row rsub of a 16-row group holds pair pair_of_row(rsub) = 0, 4, 1, 5, 2, 6, 3, 7 of the group's eight: the two
        // rows a quarter-warp reads with one LDS.128 (rsub = 2j, 2j + 1) then differ in bit 2, and under the 128-byte swizzle
        // of the x boxes (16-byte chunk c of row r sits at chunk c ^ (r & 7)) one of them reads chunks 0-3 and the other
        // chunks 4-7 -- conflict-free without the register swap that fetching the chunks in opposite order needed (a fifth
        // of the converters' instructions).  The epilogue stores a row's score to the pair it belongs to.
        const int pl = q * 16 + h * 8 + pair_of_row(rsub);       // pair within the tile = row of both x boxes
        const uint32_t st_addr = tmem + ((uint32_t)(q * 32 + h * 16) << 16) + A_COL0;
        const int off0 = pl * 128 + ((cq ^ (pl & 7)) << 4);      // k-step 0: chunks 0-3
        const int off1 = pl * 128 + (((4 + cq) ^ (pl & 7)) << 4);   // k-step 1: chunks 4-7
        const int64_t total = T * g.nst1;
        // Set s owns stages it = s, s + 2, ...; NX is even, so it owns the x slots of its own parity and is the ONLY waiter
        // of their barriers: it observes every phase (a parity wait is ambiguous for a waiter that can fall a whole
        // phase behind, which a set merely skipping the other set's stages could).  The A slots (NA odd) alternate
        // between the sets; a set reaches stage `it` only after the MMA warp has consumed stage it - 2 - NA, so the
        // barrier is at most one phase behind the one waited for.
        Ring rx(NX), ra(NA);
        if (cset) { rx.stage = 1; ra.stage = 1; }
        float amax = 0.f;                 // MODE 1 lower range guard: largest sampled |x| of this thread's pair in the current tile
        int sit = cset;                   // stage within the tile
        int64_t tile_i = 0;
        auto advance2 = [](Ring &r) { r.stage += 2; if (r.stage >= (uint32_t)r.n) { r.stage -= (uint32_t)r.n; r.phase ^= 1; } };
        for (int64_t it = cset; it < total; it += 2, advance2(rx), advance2(ra)) {
            PMARK(5);
            TWAIT(&x_full[rx.stage], rx.phase, 0x104 | ((int)it << 20));
            PMARK(0);
            const uint8_t *xs = Xs + rx.stage * X_STAGE;
            float4 a0, a1, b0, b1;
            if (!DBG(16)) {
                a0 = *reinterpret_cast<const float4 *>(xs + off0);
                a1 = *reinterpret_cast<const float4 *>(xs + off1);
                b0 = *reinterpret_cast<const float4 *>(xs + X_BOX + off0);
                b1 = *reinterpret_cast<const float4 *>(xs + X_BOX + off1);
            } else {
                a0 = a1 = b0 = b1 = make_float4(1.f, 2.f, 3.f, 4.f);
            }
            // registers of tcgen05.st.16x256b.x2: r0,r1 -> (row, cols 2j,2j+1)  r2,r3 -> (row+8, same cols);
            // r4..r7 the same for the next 8 columns (k + 16)
            uint32_t hi[8], lo[8];
            if (DBG(16)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { hi[j] = 0x3c003c00u; lo[j] = 0u; }
            } else if (MODE == 1) {
                // x' = 2^9 x = h + r:  hi[0..7] = h = fp16(x') (same register layout as the bf16 path); lo[0..3] = e4m3(r)
                // (|r| <= |x| / 4); lo[4..7] = e4m3(x).  One e4m3 register = K slots 8 cq + {0..3} (k-step 0 chunk) or
                // + {4..7} (k-step 1 chunk): the weight image uses the same slot permutation.
                const float v[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t hp[8];
                float r[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s0 = v[2 * j] * 512.f, s1v = v[2 * j + 1] * 512.f;
                    hp[j] = pack_f16x2(s0, s1v);
                    float h0, h1;
                    asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}\n" : "=f"(h0), "=f"(h1) : "r"(hp[j]));
                    r[2 * j] = s0 - h0; r[2 * j + 1] = s1v - h1;
                }
                // lower range guard on a sample of the values (the upper one is the epilogue's)
                amax = fmaxf(fmaxf(amax, fabsf(v[0])), fmaxf(fabsf(v[5]), fmaxf(fabsf(v[10]), fabsf(v[15]))));
                hi[0] = hp[0]; hi[1] = hp[1];       // side 0, k-step 0
                hi[2] = hp[4]; hi[3] = hp[5];       // side 1, k-step 0
                hi[4] = hp[2]; hi[5] = hp[3];       // side 0, k-step 1
                hi[6] = hp[6]; hi[7] = hp[7];       // side 1, k-step 1
                lo[0] = pack_e4m3x4(r[0], r[1], r[2], r[3]); lo[1] = pack_e4m3x4(r[4], r[5], r[6], r[7]);
                lo[2] = pack_e4m3x4(r[8], r[9], r[10], r[11]); lo[3] = pack_e4m3x4(r[12], r[13], r[14], r[15]);
                lo[4] = pack_e4m3x4(v[0], v[1], v[2], v[3]); lo[5] = pack_e4m3x4(v[4], v[5], v[6], v[7]);
                lo[6] = pack_e4m3x4(v[8], v[9], v[10], v[11]); lo[7] = pack_e4m3x4(v[12], v[13], v[14], v[15]);
            } else {
                split_bf16x2(a0.x, a0.y, hi[0], lo[0]); split_bf16x2(a0.z, a0.w, hi[1], lo[1]);
                split_bf16x2(b0.x, b0.y, hi[2], lo[2]); split_bf16x2(b0.z, b0.w, hi[3], lo[3]);
                split_bf16x2(a1.x, a1.y, hi[4], lo[4]); split_bf16x2(a1.z, a1.w, hi[5], lo[5]);
                split_bf16x2(b1.x, b1.y, hi[6], lo[6]); split_bf16x2(b1.z, b1.w, hi[7], lo[7]);
            }
#ifdef TCP_PROF
            if (hi[0] == 0x12345678u && lo[7] == 0x9abcdefu) g.scores[0] = 0.f;   // pin the conversion before the mark
#endif
            PMARK(1);
            TWAIT(&a_empty[ra.stage], ra.phase ^ 1, 0x105 | ((int)it << 20));
            tc_fence_after();
            PMARK(2);
            const uint32_t col = st_addr + ra.stage * A_STAGE_COLS;
            if (DBG(8)) {
                if (hi[0] == 0x12345678u && lo[7] == 0x9abcdefu) g.scores[0] = 0.f;   // keep the conversion alive
            } else {
            tmem_st_16x256b_x2(col, hi);
            if (MODE == 1) {
                const uint32_t e0[4] = {lo[0], lo[1], lo[2], lo[3]}, e1[4] = {lo[4], lo[5], lo[6], lo[7]};
                tmem_st_16x256b_x1(col + 16, e0);
                tmem_st_16x256b_x1(col + 24, e1);
            } else {
                tmem_st_16x256b_x2(col + 16, lo);
            }
            }
            // Release the x slot only now: the stores above consume every register the four LDS wrote, so the
            // shared-memory reads have completed.
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[rx.stage]);
            PMARK(3);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&a_full[ra.stage], 0);
            if (MODE == 1) {
                // after this set's last stage of a tile: the e4m3 terms need typical |x| of at least ~2^-2 (each set checks
                // the stages it converted); nst1 is even, so the sets keep their stage parity in every tile
                sit += 2;
                if (sit >= g.nst1) {
                    if (tile_base(tile_i) + pl < g.n && (amax < 0.25f || !img_ok)) *reinterpret_cast<volatile int *>(g.guard) = 1;
                    amax = 0.f; sit -= g.nst1; ++tile_i;
                }
            }
            PMARK(4);
        }
        PFLUSH();
    } else if (warp == WARP_MMA) {
        // =============================== MMA ISSUER (leader CTA) ===============================
        // The whole warp runs the loop converged; only the tcgen05 instructions are issued by one elected lane
        // (warp-uniform control flow keeps ring indices and descriptors in uniform registers).
        Ring ra(NA), rb(NB);
        const int nst2 = (g.ksteps2 + 1) / 2;
        if (rank == 0) {
            const uint32_t b_base = smem_addr(Bs), u_base = smem_addr(Us);
            const int half = g.nst1 / 2;
            // One stage: K = 32 as two K = 16 steps, each hi*Whi + lo*Whi + hi*Wlo.  A: TMEM columns of the stage (same
            // address in both CTAs); B: chunks 0-3 hi, 4-7 lo of this CTA's half, two chunks per step.  The barriers of
            // stage s + 1 are probed before the MMAs of stage s are issued (the answers arrive under them).
            auto layer1 = [&](uint32_t dcol, int s_begin, int s_end) {
                if (s_begin >= s_end) return;
                PMARK(6);
                TWAIT(&a_full[ra.stage], ra.phase, 0x106 | (s_begin << 20));
                PMARK(1);
                TWAIT(&b_full[rb.stage], rb.phase, 0x107);
                PMARK(2);
                for (int s = s_begin; s < s_end; ++s) {
                    tc_fence_after();
                    const uint32_t acol = tmem + A_COL0 + ra.stage * A_STAGE_COLS;
                    const uint64_t bd = make_smem_desc(b_base + rb.stage * B_HALF, KCH_BH, 128);
                    Ring na = ra, nb = rb;
                    na.advance();
                    nb.advance();
                    // probes of the next stage's barriers: issued before this stage's MMAs, answered under them
                    uint32_t ok_a = 1, ok_b = 1;
                    if (s + 1 < s_end) { ok_a = mbar_try(&a_full[na.stage], na.phase); ok_b = mbar_try(&b_full[nb.stage], nb.phase); }
                    // one issue block per stage: the probes made the mid-stage waits unnecessary, and every elected block
                    // costs its own vote, branch and register -> uniform-register moves
                    if (elect_one()) {
                        if (DBG(4)) {
                        } else if (MODE == 1) {
                            mma2_ts(dcol, acol, bd, IDESC0, s != 0);                                      // fp16 x fp16, K steps 0 and 1 (chunks 0-1, 2-3)
                            mma2_ts(dcol, acol + 8, bd + ((2 * KCH_BH) >> 4), IDESC0, 1);
                            mma2_f8_ts(dcol, acol + 16, bd + ((4 * KCH_BH) >> 4), IDESC0, 1);          // e4m3(residual) x e4m3(W'h), K = 32 (chunks 4-5)
                            mma2_f8_ts(dcol, acol + 24, bd + ((6 * KCH_BH) >> 4), IDESC0, 1);          // e4m3(x) x e4m3(2^9 W'l) (chunks 6-7)
                        } else {
                            mma2_ts(dcol, acol, bd, IDESC, s != 0);
                            mma2_ts(dcol, acol + 16, bd, IDESC, 1);
                            mma2_ts(dcol, acol, bd + ((4 * KCH_BH) >> 4), IDESC, 1);
                            mma2_ts(dcol, acol + 8, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                            mma2_ts(dcol, acol + 24, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                            mma2_ts(dcol, acol + 8, bd + ((6 * KCH_BH) >> 4), IDESC, 1);
                        }
                        mma2_commit_mc(&a_empty[ra.stage], 3);
                        mma2_commit_mc(&b_empty[rb.stage], 3);
                    }
                    __syncwarp();
                    PMARK(3);
                    if (!ok_a) TWAIT(&a_full[na.stage], na.phase, 0x108 | ((s + 1) << 20));
                    PMARK(1);
                    if (!ok_b) TWAIT(&b_full[nb.stage], nb.phase, 0x109 | ((s + 1) << 20));
                    PMARK(2);
                    __syncwarp();
                    ra = na;
                    rb = nb;
                    PMARK(3);
                }
            };
            // Order per iteration: first half of layer 1 (tile i), layer 2 (tile i - 1), second half of layer 1: the final
            // epilogue of tile i - 1 runs under the second half and the D buffer it frees is ready for tile i + 1.
            for (int64_t i = 0; i <= T; ++i) {
                const uint32_t dcol_i = tmem + (uint32_t)(i & 1) * NPAD;
                if (i < T) {
                    PMARK(6);
                    TWAIT(&d_empty[i & 1], (uint32_t)(((i >> 1) & 1) ^ 1), 0x10a);
                    tc_fence_after();
                    PMARK(0);
                    layer1(dcol_i, 0, half);
                }
                if (i >= 1) {      // layer 2 of tile i - 1: Y = U W2^T into the accumulator that held D
                    const int64_t j = i - 1;
                    const int d = (int)(j & 1);
                    const uint32_t dcol = tmem + d * NPAD;
                    PMARK(6);
                    TWAIT(u_full, (uint32_t)(j & 1), 0x10b);
                    tc_fence_after();
                    PMARK(4);
                    TWAIT(&b_full[rb.stage], rb.phase, 0x10c);
                    PMARK(2);
                    for (int ks = 0; ks < g.ksteps2; ks += 2) {
                        tc_fence_after();
                        const uint64_t bd = make_smem_desc(b_base + rb.stage * B_HALF, KCH_BH, 128);
                        const uint64_t uhi = make_smem_desc(u_base + ks * 2 * KCH_U, KCH_U, 128);
                        const uint64_t ulo = make_smem_desc(u_base + U_HALF + ks * 2 * KCH_U, KCH_U, 128);
                        Ring nb = rb;
                        nb.advance();
                        const uint32_t ok_b = ks + 2 < g.ksteps2 ? mbar_try(&b_full[nb.stage], nb.phase) : 1u;
                        if (elect_one()) {
                            if (!DBG(4)) {
                                mma2_ss(dcol, uhi, bd, IDESC, ks != 0);
                                mma2_ss(dcol, ulo, bd, IDESC, 1);
                                mma2_ss(dcol, uhi, bd + ((4 * KCH_BH) >> 4), IDESC, 1);
                                if (ks + 1 < g.ksteps2) {
                                    const uint64_t uhi1 = uhi + ((2 * KCH_U) >> 4), ulo1 = ulo + ((2 * KCH_U) >> 4);
                                    mma2_ss(dcol, uhi1, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                                    mma2_ss(dcol, ulo1, bd + ((2 * KCH_BH) >> 4), IDESC, 1);
                                    mma2_ss(dcol, uhi1, bd + ((6 * KCH_BH) >> 4), IDESC, 1);
                                }
                            }
                            mma2_commit_mc(&b_empty[rb.stage], 3);
                        }
                        __syncwarp();
                        PMARK(5);
                        if (!ok_b) TWAIT(&b_full[nb.stage], nb.phase, 0x10d);
                        PMARK(2);
                        __syncwarp();
                        rb = nb;
                    }
                    if (elect_one()) {
                        mma2_commit_mc(&y_full[d], 3);
                        mma2_commit_mc(u_empty, 3);
                    }
                    __syncwarp();
                    PMARK(5);
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
            for (int k = 0; k < NA && k < T * g.nst1; ++k) { TWAIT(&a_empty[ra.stage], ra.phase ^ 1, 0x10e); ra.advance(); }
            for (int k = 0; k < NB; ++k) { TWAIT(&b_empty[rb.stage], rb.phase ^ 1, 0x10f); rb.advance(); }
            TWAIT(u_empty, (uint32_t)((T & 1) ^ 1), 0x110);
        }
        PMARK(7);
        PFLUSH();
    } else if (warp == WARP_BLOAD) {
        // =============================== B LOADER (this CTA's half of every weight stage) ===============================
        if (lane == 0) {
            Ring rb(NB);
            const int half = g.nst1 / 2;
            const int nst2 = (g.ksteps2 + 1) / 2;
            auto put = [&](const CUtensorMap *map, int stage_idx) {
                PMARK(1);
                TWAIT(&b_empty[rb.stage], rb.phase ^ 1, 0x111 | (stage_idx << 20));
                PMARK(0);
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
            PMARK(1);
            PFLUSH();
        }
    } else {
        // =============================== X LOADER (this CTA's tile) ===============================
        if (lane == 0) {
            Ring rx(NX);
            for (int64_t i = 0; i < T; ++i) {
                const int row0 = (int)tile_base(i);
                for (int s = 0; s < g.nst1; ++s) {
                    PMARK(1);
                    TWAIT(&x_empty[rx.stage], rx.phase ^ 1, 0x112 | ((int)(i * g.nst1 + s) << 20));
                    PMARK(0);
                    mbar_arrive_expect_tx(&x_full[rx.stage], X_STAGE);
                    uint8_t *dst = Xs + rx.stage * X_STAGE;
                    tma_load_2d(dst, &mapX1, s * KST, row0, &x_full[rx.stage]);      // rows past n are zero-filled
                    tma_load_2d(dst + X_BOX, &mapX2, s * KST, row0, &x_full[rx.stage]);
                    rx.advance();
                }
            }
            PMARK(1);
            PFLUSH();
        }
    }

    // ---- teardown ----
    TMARK(0x1f0);
    __syncwarp();                 // single-lane roles: the whole warp arrives at the (aligned) cluster barrier together
    tc_fence_before();
    cluster_sync_all();           // the leader's MMAs read the peer's tensor / shared memory: nobody leaves early
    TMARK(0x1ff);
    if (warp == WARP_MMA) tmem_dealloc2(tmem, 512);
    if (MODE == 0 && g.guard != nullptr && tid == 0) {
        __threadfence();
        if (atomicAdd(g.guard + 1, 1) == (int)gridDim.x - 1) { g.guard[1] = 0; g.guard[0] = 0; }
    }
}

// ---- pair weight images (bf16) ------------------------------------------------------------------
// Stage s (K = [32 s, 32 s + 32)) = [half 0][half 1], a half = weight rows [88 h, 88 h + 88) as 8 chunks
// [hi k 0-7][hi 8-15][hi 16-23][hi 24-31][lo x 4], a chunk = 11 core matrices of 8 rows x 8 k (128 B each).
// One launch packs both layers: stages [0, nst1) from W1 [N1][K1], then [nst1, nst1 + nst2) from W2 [N2][K2] into img2.
__global__ void pair_pack_bf16_kernel(const float *__restrict__ W1, int N1, int K1, int nst1, uint8_t *__restrict__ img1,
                                      const float *__restrict__ W2, int N2, int K2, int nst2, uint8_t *__restrict__ img2) {
    const int64_t per_stage = (int64_t)KST * NPAD;
    const int64_t total = (int64_t)(nst1 + nst2) * per_stage;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % KST);
        const int n = (int)((e / KST) % NPAD);
        int s = (int)(e / per_stage);
        const bool second = s >= nst1;
        if (second) s -= nst1;
        const float *W = second ? W2 : W1;
        const int N = second ? N2 : N1, K = second ? K2 : K1;
        uint8_t *img = second ? img2 : img1;
        const int k = s * KST + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * K + k] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const int hsel = n / NH, nn = n % NH;
        uint8_t *st = img + ((size_t)s * 2 + hsel) * B_HALF;
        const size_t off = (size_t)(kk >> 3) * KCH_BH + (nn >> 3) * 128 + (nn & 7) * 16 + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(st + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(st + 4 * KCH_BH + off) = lo;
    }
}

// MODE 1 image of W1 (see score_tc.cu's tc_pack_mixed_kernel for the scaling): W' = W 2^gm with max|W'| in [32, 64),
// W' = Wh + Wl, Wh = fp16(W').  Stage s, half h: 8 chunks of 11 core matrices
//   [fp16 Wh k 0-7][k 8-15][k 16-23][k 24-31][e4m3 Wh slots 0-15][slots 16-31][e4m3 2^9 Wl slots 0-15][slots 16-31]
// e4m3 slot t holds k = 4 (t >> 3) + (t & 3) + 16 ((t >> 2) & 1): the order in which a converter thread's two 16-byte
// loads land in one tcgen05.st.16x256b register pair.  hdr16[0] = 2^gw1 with max|W| 2^gw1 in [8192, 16384): 2^gm = 2^gw1 / 256.
__global__ void pair_pack_mixed_kernel(const float *__restrict__ W, int N, int K, int nstages, const float *__restrict__ hdr16,
                                       uint8_t *__restrict__ img) {
    const float up = hdr16[0] * (1.f / 256.f);
    const int64_t total = (int64_t)nstages * NPAD * KST;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(e % KST);
        const int n = (int)((e / KST) % NPAD);
        const int s = (int)(e / ((int64_t)KST * NPAD));
        const int k = s * KST + kk;
        const float w = (n < N && k < K) ? W[(int64_t)n * K + k] * up : 0.f;
        const __half wh = __float2half_rn(w);
        const float whf = __half2float(wh);
        const int hsel = n / NH, nn = n % NH;
        uint8_t *st = img + ((size_t)s * 2 + hsel) * B_HALF;
        const size_t row = (size_t)(nn >> 3) * 128 + (nn & 7) * 16;
        *reinterpret_cast<__half *>(st + (size_t)(kk >> 3) * KCH_BH + row + (kk & 7) * 2) = wh;
        const int t = ((kk & 15) >> 2) * 8 + ((kk >> 4) & 1) * 4 + (kk & 3);      // slot of k within the stage
        const size_t off8 = (size_t)(t >> 4) * KCH_BH + row + (t & 15);
        st[4 * KCH_BH + off8] = (uint8_t)__nv_cvt_float_to_fp8(whf, __NV_SATFINITE, __NV_E4M3);                  // pairs with e4m3(residual)
        st[6 * KCH_BH + off8] = (uint8_t)__nv_cvt_float_to_fp8((w - whf) * 512.f, __NV_SATFINITE, __NV_E4M3);    // pairs with e4m3(x)
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
// rows of d_in floats; box = [64 rows x 32 floats], 128-byte swizzle; rows past n read as zeros
static bool make_x_map(CUtensorMap *m, const float *x, int64_t n, int d_in) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)d_in, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)d_in * 4};
    cuuint32_t box[2] = {KST, TP}, es[2] = {1, 1};
    // 256-byte L2 promotion: the second half of a promoted sector is the same row's next K = 32 stage, read 1-2 us later
    // (1 % faster than 128-byte promotion or none, A/B on one box; DRAM bytes unchanged)
#ifndef TCP_X_PROMOTION
#define TCP_X_PROMOTION CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, TCP_X_PROMOTION, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

}  // namespace tcp

// ---- interface used by pack.cu / api.cu -------------------------------------------------------
static bool tcp_dims_ok(int d_in, int d1, int d2) {
    return d_in % tcp::KST == 0 && d_in >= 2 * tcp::KST && d1 <= tcp::NPAD && d2 <= tcp::NPAD && d1 >= 1 && d2 >= 1;
}
static int64_t tcp_pair_image_bytes(int nstages) { return ((int64_t)nstages * 2 * tcp::B_HALF + 255) / 256 * 256; }
static int tcp_nst2(int d1) { return (round_up(d1, 16) / 16 + 1) / 2; }

int64_t tcp_image_bytes(int d_in, int d1, int d2) {
    if (!tcp_dims_ok(d_in, d1, d2)) return 0;
    return 2 * tcp_pair_image_bytes(d_in / tcp::KST) + tcp_pair_image_bytes(tcp_nst2(d1));     // bf16 W1, bf16 W2, mixed W1
}

bool tcp_shape_ok(const PackLayout &L) { return tcp_dims_ok(L.d_in, L.d1, L.d2) && L.tcp_bytes > 0; }

const float *tc_hdr16(const PackLayout &L, const char *pack);   // score_tc.cu: weight scales (tc_scales_kernel)
const float *tc_hdr_mixed(const PackLayout &L, const char *pack);   // score_tc.cu: [2] != 0 when the pack built the mixed images
int *tc_guard_slot();                                           // score_tc.cu: range-guard slots of the mixed-precision paths

// pair images of a NeuralPlda pack: [bf16 W1: d_in / 32 stages][bf16 W2: ceil(ksteps2 / 2) stages] in one launch; with
// NPLDA_PACK_MIXED also [mixed W1] (it reads the scales the tensor-core pack left on this stream)
int tcp_pack_nplda(const float *W1, const float *W2, const PackLayout &L, char *pack, int flags, cudaStream_t st) {
    if (!tcp_shape_ok(L)) return NPLDA_OK;
    uint8_t *img1 = (uint8_t *)pack + L.tcp;
    uint8_t *img2 = img1 + tcp_pair_image_bytes(L.d_in / tcp::KST);
    uint8_t *img1m = img2 + tcp_pair_image_bytes(tcp_nst2(L.d1));
    tcp::pair_pack_bf16_kernel<<<sm_count(), 256, 0, st>>>(W1, L.d1, L.d_in, L.d_in / tcp::KST, img1, W2, L.d2, L.d1,
                                                          tcp_nst2(L.d1), img2);
    NPLDA_LAUNCH_CHECK();
    if (!(flags & NPLDA_PACK_MIXED)) return NPLDA_OK;      // the mixed images are opt-in (tc_pack_nplda marks them valid / invalid)
    tcp::pair_pack_mixed_kernel<<<sm_count(), 256, 0, st>>>(W1, L.d1, L.d_in, L.d_in / tcp::KST, tc_hdr16(L, pack), img1m);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

#if defined(TCP_TRACE) || defined(TCP_PROF)
static int *g_tcp_trace = nullptr;
extern "C" void nplda_debug_set_tcp_trace(void *p) { g_tcp_trace = (int *)p; }
#endif

template <int MODE>
static int launch_tcp(const CUtensorMap &mX1, const CUtensorMap &mX2, const CUtensorMap &mW1, const CUtensorMap &mW2,
                      const tcp::Args &a, int grid, cudaStream_t st) {
    NPLDA_CUDA_TRY(cudaFuncSetAttribute(tcp::score_tcp_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::SMEM_BYTES));
    tcp::score_tcp_kernel<MODE><<<grid, tcp::NTHREADS, tcp::SMEM_BYTES, st>>>(mX1, mX2, mW1, mW2, a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

// mixed = false: bf16x3 (any fp32 range).  mixed = true: MODE 1 for layer 1, then the bf16x3 kernel as a guarded fallback
// pass on the same stream (its CTAs return at once unless the range guard fired).
int score_tcp(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores, bool mixed,
              cudaStream_t st) {
    if (!tcp_shape_ok(L)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n >= ((int64_t)1 << 31) - 4 * tcp::TP) return NPLDA_ERR_UNSUPPORTED_DIM;   // TMA row coordinates are int32
    const uint8_t *img1 = (const uint8_t *)pack + L.tcp;
    const uint8_t *img2 = img1 + tcp_pair_image_bytes(L.d_in / tcp::KST);
    const uint8_t *img1m = img2 + tcp_pair_image_bytes(tcp_nst2(L.d1));
    CUtensorMap mX1, mX2, mW1, mW2, mW1m;
    if (!tcp::make_x_map(&mX1, x1, n, L.d_in) || !tcp::make_x_map(&mX2, x2, n, L.d_in) ||
        !tcp::make_image_map(&mW1, img1, L.d_in / tcp::KST) || !tcp::make_image_map(&mW2, img2, tcp_nst2(L.d1)) ||
        !tcp::make_image_map(&mW1m, img1m, L.d_in / tcp::KST))
        return NPLDA_ERR_NO_DEVICE;
    tcp::Args a;
    a.n = n;
    a.nst1 = L.d_in / tcp::KST; a.ksteps2 = round_up(L.d1, 16) / 16;
    a.b1 = (const float *)(pack + L.b1); a.b2 = (const float *)(pack + L.b2);
    a.p = (const float *)(pack + L.p); a.q = (const float *)(pack + L.q);
    a.scores = scores;
    a.hdr16 = tc_hdr16(L, pack);
    a.hdrm = tc_hdr_mixed(L, pack);
    a.guard = nullptr;
    a.dbg = 0;
#ifdef TCP_DBG
    { static const int v = getenv("NPLDA_TCP_DEBUG") ? atoi(getenv("NPLDA_TCP_DEBUG")) : 0; a.dbg = v; }
#endif
    a.trace = nullptr;
#if defined(TCP_TRACE) || defined(TCP_PROF)
    a.trace = g_tcp_trace;
#endif
    const int64_t nsuper = (n + 2 * tcp::TP - 1) / (2 * tcp::TP);
    const int grid = 2 * (int)std::min<int64_t>(nsuper, sm_count() / 2);
    if (!mixed) return launch_tcp<0>(mX1, mX2, mW1, mW2, a, grid, st);
    int *slot = tc_guard_slot();
    if (!slot) return NPLDA_ERR_NO_DEVICE;
    a.guard = slot;
    const int rc = launch_tcp<1>(mX1, mX2, mW1m, mW2, a, grid, st);
    if (rc != NPLDA_OK) return rc;
    a.trace = nullptr;
    return launch_tcp<0>(mX1, mX2, mW1, mW2, a, grid, st);
}

}  // namespace nplda
