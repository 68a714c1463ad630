// K5-TC: enrol x test GRID scoring on the tensor cores.
//
// With the per-utterance rows of nplda_table_prepare a grid score is  S[i][j] = r[e_i] + r[t_j] + A[e_i] . B[t_j]
// (pairs.cu).  The two r terms ride inside the contraction: A' = [A (d) | r | 1 | 0...], B' = [B (d) | 1 | r | 0...],
// so S = A' . B' over K = d + 2 <= 176, padded to 192 -- one [E, 192] x [192, T] product and nothing else.
//
// Operands ("fp16x3", as score_tcx.cu): the prepare step leaves every utterance's A' and B' as fp16 hi/lo of
// A' 2^ka and B' 2^kb (scales from the absolute maxima over the table), K-major, 768 bytes each:
// [hi k 0-63][hi 64-127][hi 128-191][lo 0-63][lo 64-127][lo 128-191], i.e. six 128-byte swizzle rows.  Rows are
// gathered by index with TMA tile::gather4 straight into 128-byte-swizzled shared-memory operands; the product is
// hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM, the epilogue scales by 2^-(ka+kb) and streams the scores out
// (the only HBM traffic: 4 bytes per trial).
//
// One CTA = one tile of 128 enrol rows (A' resident in shared memory, 96 KB) x a range of 128-column test tiles whose
// B' chunks (K = 64, hi + lo, 32 KB) stream through a 3-stage ring; two 128-column accumulators in TMEM so that the
// epilogue of one test tile runs under the MMAs of the next.  Warp roles: 4 gather warps, 1 MMA warp, 4 epilogue warps.
#include <algorithm>
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_pair_ptx.cuh"

namespace nplda {
namespace gtc {

using namespace tc;

constexpr int TM = 128, TN = 128;               // enrol rows x test columns per accumulator
constexpr int KP = 192, KCH = 64, NCH = KP / KCH;   // padded K, K per chunk, chunks
constexpr int SUB = 128 * 128;                  // 16384 B: [128 rows][64 fp16], 128-byte swizzle
constexpr int A_BYTES = 2 * NCH * SUB;          // 98304: hi chunks 0-2, lo chunks 0-2
constexpr int B_STAGE = 2 * SUB;                // 32768: one chunk, hi then lo
constexpr int NB = 3;
constexpr int OP_BYTES = 2 * KP * 2;            // 768 B per operand row (hi + lo)
constexpr int ROW_BYTES = 2 * OP_BYTES;         // 1536 B per utterance: A' then B'

constexpr int LOAD_WARPS = 4, EPI_WARPS = 4;
constexpr int WARP_MMA = LOAD_WARPS, WARP_EPI = WARP_MMA + 1;
constexpr int NTHREADS = (WARP_EPI + EPI_WARPS) * 32;    // 288

constexpr int SM_A = 0, SM_B = SM_A + A_BYTES, SM_BAR = SM_B + NB * B_STAGE;
constexpr int N_BARS = 1 + 2 * NB + 4;
constexpr int SM_TMEM = SM_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct Args {
    const int64_t *er, *tr;
    int64_t E, T, n_rows, ld;
    const float *hdr;           // [2] 2^-ka, [3] 2^-kb (grid operand header, see gtab_build_kernel)
    float *scores;
    int32_t *bad_flag;
    int tiles_per_cta;          // test tiles per CTA (blockIdx.x selects the range)
};

struct Ring {
    uint32_t stage = 0, phase = 0;
    __device__ void advance() { if (++stage == NB) { stage = 0; phase ^= 1; } }
};

__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(bar)) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) score_grid_tc_kernel(const __grid_constant__ CUtensorMap map, Args g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    uint8_t *As = smem + SM_A, *Bs = smem + SM_B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = b_full + NB, *d_full = b_empty + NB, *d_empty = d_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int64_t e0 = (int64_t)blockIdx.y * TM;
    const int64_t ntt = (g.T + TN - 1) / TN;
    const int64_t tt0 = (int64_t)blockIdx.x * g.tiles_per_cta;
    const int ntile = (int)max((int64_t)0, min((int64_t)g.tiles_per_cta, ntt - tt0));

    if (tid == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&d_full[d], 1); mbar_init(&d_empty[d], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = make_idesc_f16(TM, TN);

    if (warp < LOAD_WARPS) {
        // =============================== GATHER (lane = row of the tile) ===============================
        const int R = 32 * warp + lane;
        auto row_of = [&](const int64_t *idx, int64_t pos, int64_t count) {
            int64_t rid = g.n_rows;                   // the all-zero operand row behind the table: score 0
            if (pos < count) {
                rid = idx[pos];
                if (rid < 0 || rid >= g.n_rows) { *g.bad_flag = 1; rid = g.n_rows; }   // reported; its row / column reads 0
            }
            return (int)rid;
        };
        const int q4 = lane & ~3;
        const bool issuer = (lane & 3) == 0;
        {   // the enrol tile: A' of 128 rows, six 16 KB sub-tiles, resident for the whole CTA
            const int rid = row_of(g.er, e0 + R, g.E);
            const int r0 = __shfl_sync(0xffffffffu, rid, q4), r1 = __shfl_sync(0xffffffffu, rid, q4 + 1);
            const int r2 = __shfl_sync(0xffffffffu, rid, q4 + 2), r3 = __shfl_sync(0xffffffffu, rid, q4 + 3);
            if (warp == 0 && lane == 0) mbar_arrive_expect_tx(a_full, A_BYTES);
            __syncwarp();
            if (issuer)
                for (int c = 0; c < 2 * NCH; ++c)
                    tma_gather4(As + c * SUB + R * 128, &map, c * KCH, r0, r1, r2, r3, a_full);
        }
        Ring rb;
        for (int t = 0; t < ntile; ++t) {
            const int rid = row_of(g.tr, (tt0 + t) * TN + R, g.T);
            const int r0 = __shfl_sync(0xffffffffu, rid, q4), r1 = __shfl_sync(0xffffffffu, rid, q4 + 1);
            const int r2 = __shfl_sync(0xffffffffu, rid, q4 + 2), r3 = __shfl_sync(0xffffffffu, rid, q4 + 3);
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(&b_empty[rb.stage], rb.phase ^ 1);
                if (warp == 0 && lane == 0) mbar_arrive_expect_tx(&b_full[rb.stage], B_STAGE);
                __syncwarp();
                if (issuer) {
                    uint8_t *dst = Bs + rb.stage * B_STAGE + R * 128;
                    tma_gather4(dst, &map, 2 * KP + c * KCH, r0, r1, r2, r3, &b_full[rb.stage]);                 // B' hi chunk c
                    tma_gather4(dst + SUB, &map, 2 * KP + KP + c * KCH, r0, r1, r2, r3, &b_full[rb.stage]);      // B' lo chunk c
                }
                __syncwarp();
                rb.advance();
            }
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA ISSUER ===============================
        const uint32_t a_base = smem_addr(As), b_base = smem_addr(Bs);
        Ring rb;
        mbar_wait(a_full, 0);
        for (int t = 0; t < ntile; ++t) {
            const int d = t & 1;
            mbar_wait(&d_empty[d], (uint32_t)(((t >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t dcol = tmem + d * TN;
            for (int c = 0; c < NCH; ++c) {
                mbar_wait(&b_full[rb.stage], rb.phase);
                tc_fence_after();
                const uint64_t ahi = make_smem_desc_sw128(a_base + c * SUB), alo = make_smem_desc_sw128(a_base + (NCH + c) * SUB);
                const uint64_t bhi = make_smem_desc_sw128(b_base + rb.stage * B_STAGE), blo = bhi + (SUB >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < KCH / 16; ++k) {          // 32 bytes per K step inside the swizzle row
                        mma_ss(dcol, ahi + 2 * k, bhi + 2 * k, IDESC, (c | k) != 0);
                        mma_ss(dcol, alo + 2 * k, bhi + 2 * k, IDESC, 1);
                        mma_ss(dcol, ahi + 2 * k, blo + 2 * k, IDESC, 1);
                    }
                    mma_commit(&b_empty[rb.stage]);
                }
                __syncwarp();
                rb.advance();
            }
            if (elect_one()) mma_commit(&d_full[d]);
            __syncwarp();
        }
        // drain: the arrivals of the last commits on b_empty must land before the CTA exits
        if (ntile > 0)
            for (int k = 0; k < NB; ++k) { mbar_wait(&b_empty[rb.stage], rb.phase ^ 1); rb.advance(); }
    } else {
        // =============================== EPILOGUE (thread = enrol row) ===============================
        const int q = warp & 3;                                   // TMEM lane quadrant of this warp
        const int row = q * 32 + lane;
        const int64_t e = e0 + row;
        const float s = g.hdr[2] * g.hdr[3];
        for (int t = 0; t < ntile; ++t) {
            const int d = t & 1;
            mbar_wait(&d_full[d], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const int64_t c0 = (tt0 + t) * TN;
            float *out = g.scores + e * g.ld + c0;
#pragma unroll 1
            for (int cc = 0; cc < TN; cc += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + d * TN + cc, v);
                tmem_ld_wait();
                if (e < g.E) {
                    if (c0 + cc + 32 <= g.T && ((reinterpret_cast<uintptr_t>(out + cc) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            __stcs(reinterpret_cast<float4 *>(out + cc + j),
                                   make_float4(__uint_as_float(v[j]) * s, __uint_as_float(v[j + 1]) * s,
                                               __uint_as_float(v[j + 2]) * s, __uint_as_float(v[j + 3]) * s));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + cc + j < g.T) __stcs(out + cc + j, __uint_as_float(v[j]) * s);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[d]);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tmem, 256);
}

// ---- grid operands (part of nplda_table_prepare) --------------------------------------------------------------
// hdr (floats, inside the row table's 256-byte trailer): [0], [1] bit patterns of max|A'|, max|B'| (atomicMax slots,
// zeroed by the caller), [2] 2^-ka, [3] 2^-kb.
__global__ void __launch_bounds__(256) gtab_absmax_kernel(const float *__restrict__ rowtab, int64_t n_rows, int row_floats, int row_ld,
                                                          uint32_t *__restrict__ hdr, const unsigned long long *fp_cur,
                                                          const unsigned long long *fp_built) {
    if (fp_cur != nullptr && *fp_cur == *fp_built) return;
    float ma = 1.f, mb = 1.f;                                       // the constant 1 entries
    const int64_t total = n_rows * row_floats;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % row_floats);
        const float v = fabsf(rowtab[e]);
        if (k < row_ld) {
            ma = fmaxf(ma, v);
            if (k == row_ld - 1) mb = fmaxf(mb, v);                  // r also sits in B'
        } else {
            mb = fmaxf(mb, v);
        }
    }
    for (int o = 16; o > 0; o >>= 1) { ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(hdr, __float_as_uint(ma)); atomicMax(hdr + 1, __float_as_uint(mb)); }
}

// one thread per (utterance, 8 consecutive k) of A' and of B'
__global__ void __launch_bounds__(256) gtab_build_kernel(const float *__restrict__ rowtab, int64_t n_rows, int row_floats, int row_ld, int d,
                                                         float *__restrict__ hdr, uint8_t *__restrict__ gtab,
                                                         const unsigned long long *fp_cur, const unsigned long long *fp_built) {
    if (fp_cur != nullptr && *fp_cur == *fp_built) return;
    const float sa = pow2_scale_to_2p13(hdr[0]), sb = pow2_scale_to_2p13(hdr[1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[2] = 1.f / sa; hdr[3] = 1.f / sb; }
    constexpr int PER_OP = KP / 8;                                   // 24 groups of 8 per operand
    const int64_t total = (n_rows + 1) * 2 * PER_OP;                 // + one all-zero row (bad indices, tile tails)
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / (2 * PER_OP);
        const int op = (int)((e / PER_OP) & 1), k0 = (int)(e % PER_OP) * 8;
        if (row == n_rows) {
            uint8_t *z = gtab + row * ROW_BYTES + op * OP_BYTES + k0 * 2;
            *reinterpret_cast<uint4 *>(z) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(z + KP * 2) = make_uint4(0, 0, 0, 0);
            continue;
        }
        const float *src = rowtab + row * row_floats + op * row_ld;
        const float r = rowtab[row * row_floats + row_ld - 1];
        const float sc = op ? sb : sa;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j;
            float x = k < d ? src[k] : 0.f;
            if (k == d) x = op ? 1.f : r;                            // A' = [A | r | 1],  B' = [B | 1 | r]
            if (k == d + 1) x = op ? r : 1.f;
            v[j] = x * sc;
        }
        uint4 hi, lo;
        split_f16x2(v[0], v[1], hi.x, lo.x); split_f16x2(v[2], v[3], hi.y, lo.y);
        split_f16x2(v[4], v[5], hi.z, lo.z); split_f16x2(v[6], v[7], hi.w, lo.w);
        uint8_t *dst = gtab + row * ROW_BYTES + op * OP_BYTES + k0 * 2;
        *reinterpret_cast<uint4 *>(dst) = hi;
        *reinterpret_cast<uint4 *>(dst + KP * 2) = lo;
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

}  // namespace gtc

int64_t gtab_bytes(int64_t n_rows) { return (n_rows + 1) * gtc::ROW_BYTES; }     // + the all-zero row

// called by nplda_table_prepare between the row kernels and the fingerprint commit
int gtab_prepare(const float *rowtab, int64_t n_rows, int row_floats, int row_ld, int d, float *hdr, uint8_t *gtab,
                 const unsigned long long *fp_cur, const unsigned long long *fp_built, cudaStream_t st) {
    if (d + 2 > gtc::KP) return NPLDA_ERR_UNSUPPORTED_DIM;
    // the atomicMax slots restart from zero on every call; the scales in hdr[2], hdr[3] survive an unchanged table
    NPLDA_CUDA_TRY(cudaMemsetAsync(hdr, 0, 8, st));
    const int grid = (int)std::min<int64_t>((n_rows * row_floats + 255) / 256, 8 * (int64_t)sm_count());
    gtc::gtab_absmax_kernel<<<grid, 256, 0, st>>>(rowtab, n_rows, row_floats, row_ld, (uint32_t *)hdr, fp_cur, fp_built);
    NPLDA_LAUNCH_CHECK();
    const int grid2 = (int)std::min<int64_t>((n_rows * 48 + 255) / 256, 8 * (int64_t)sm_count());
    gtc::gtab_build_kernel<<<grid2, 256, 0, st>>>(rowtab, n_rows, row_floats, row_ld, d, hdr, gtab, fp_cur, fp_built);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

int score_grid_tc(const uint8_t *gtab, const float *hdr, int64_t n_rows, const int64_t *er, int64_t E, const int64_t *tr, int64_t T,
                  float *scores, int64_t ld, int32_t *bad_flag, cudaStream_t st) {
    if (n_rows >= ((int64_t)1 << 31) - 1) return NPLDA_ERR_UNSUPPORTED_DIM;
    gtc::EncodeTiledFn enc = gtc::encode_fn();
    if (!enc) return NPLDA_ERR_NO_DEVICE;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)gtc::ROW_BYTES / 2, (cuuint64_t)n_rows + 1};
    cuuint64_t strides[1] = {(cuuint64_t)gtc::ROW_BYTES};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void *)gtab, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return NPLDA_ERR_NO_DEVICE;
    gtc::Args a;
    a.er = er; a.tr = tr; a.E = E; a.T = T; a.n_rows = n_rows; a.ld = ld;
    a.hdr = hdr; a.scores = scores; a.bad_flag = bad_flag;
    const int64_t net = (E + gtc::TM - 1) / gtc::TM, ntt = (T + gtc::TN - 1) / gtc::TN;
    if (net > 65535) return NPLDA_ERR_UNSUPPORTED_DIM;
    // at most one CTA per SM in total (one wave: a second, mostly empty wave would double the time): every enrol tile
    // is split over floor(SMs / enrol tiles) CTAs along the test tiles
    int64_t nx = std::max<int64_t>(1, std::min<int64_t>(ntt, sm_count() / net));
    a.tiles_per_cta = (int)((ntt + nx - 1) / nx);
    nx = (ntt + a.tiles_per_cta - 1) / a.tiles_per_cta;
    static thread_local int attr_dev = -1;             // the attribute is per device: set once per (thread, device)
    int dev = 0;
    NPLDA_CUDA_TRY(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        NPLDA_CUDA_TRY(cudaFuncSetAttribute(gtc::score_grid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::SMEM_BYTES));
        attr_dev = dev;
    }
    gtc::score_grid_tc_kernel<<<dim3((unsigned)nx, (unsigned)net), gtc::NTHREADS, gtc::SMEM_BYTES, st>>>(map, a);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

}  // namespace nplda
