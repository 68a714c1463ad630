// C-ABI entry points that are not tied to one kernel file: bookkeeping, the
// score-forward dispatch (SIMT vs tcgen05) and the host-buffer entry.
#include <atomic>
#include <algorithm>

#include "common.cuh"

namespace nplda {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

int score_simt(bool dplda, const float *x1, const float *x2, const int64_t *i1, const int64_t *i2,
               int64_t n_rows, int32_t *bad_flag, int64_t n, const PackLayout &L, const char *pack,
               float *scores, cudaStream_t st);   // score_simt.cu
bool tc_shape_ok(bool dplda, const PackLayout &L, bool indexed);   // score_tc.cu
int score_tc(bool dplda, const float *x1, const float *x2, const int64_t *i1, const int64_t *i2,
             int64_t n_rows, int32_t *bad_flag, int64_t n, const PackLayout &L, const char *pack,
             float *scores, int mode, cudaStream_t st, float *aout = nullptr, float *yout = nullptr,
             int64_t emit_cap = 0);   // score_tc.cu

bool tcp_shape_ok(const PackLayout &L);   // score_tcp.cu
int score_tcp(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores, bool mixed,
              cudaStream_t st);   // score_tcp.cu

int table_split(const float *table, int64_t n_rows, int d_in, void *split, cudaStream_t st);   // score_tcx.cu
int score_tcx(const void *split, int64_t n_rows, const int64_t *i1, const int64_t *i2, int64_t n, const PackLayout &L,
              const char *pack, float *scores, int32_t *bad_flag, cudaStream_t st);   // score_tcx.cu

bool tc_dplda_ok(const PackLayout &L);   // score_tc.cu
int dplda_score_tc(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                   cudaStream_t st);   // dplda_tc.cu
int dplda_score_tc_train_u(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                           float *urows, cudaStream_t st);   // dplda_tc.cu
int64_t dplda_lr_workspace_bytes(int64_t n, int d1);   // dplda_lr.cu
int dplda_lr_grad(const float *urows, int64_t n, int64_t cap, int d1, const float *dscores, float *dw_lr, float *dc_lr,
                  void *workspace, int64_t workspace_bytes, cudaStream_t st);   // dplda_lr.cu

int dplda_score_tc_train(const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack, float *scores,
                         float *act, cudaStream_t st);   // dplda_tc.cu

int simt_aux(int mode, const float *x1, const float *x2, int64_t n, const PackLayout &L, const char *pack,
             float *out, int64_t ld_out, cudaStream_t st, const unsigned long long *fp_cur = nullptr,
             const unsigned long long *fp_built = nullptr);   // score_simt.cu

// S = sum_k Q y1^2 + Q y2^2 + 2 P_sqrt^2 y1 y2 from materialised embeddings (models.py:372-376): one warp per pair
__global__ void __launch_bounds__(256) score_from_emb_kernel(const float *__restrict__ y1, const float *__restrict__ y2,
                                                             int64_t n, int d, const float *__restrict__ p_sqrt,
                                                             const float *__restrict__ q, float *__restrict__ s) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = w0; i < n; i += nw) {
        float acc = 0.f;
        for (int k = lane; k < d; k += 32) {
            const float a = y1[i * d + k], b = y2[i * d + k], ps = p_sqrt[k];
            acc += q[k] * (a * a + b * b) + 2.f * (ps * ps) * (a * b);
        }
        acc = warp_sum(acc);
        if (lane == 0) s[i] = acc;
    }
}

static int score_dispatch(bool dplda, const float *x1, const float *x2, const int64_t *i1,
                          const int64_t *i2, int64_t n_rows, int32_t *bad_flag, int64_t n, int d_in,
                          int d1, int d2, const void *pack, float *scores, int impl, void *stream) {
    const bool indexed = i1 != nullptr || i2 != nullptr;
    if (n < 0 || !pack || (n > 0 && (!x1 || !scores))) return NPLDA_ERR_BAD_ARG;
    if (indexed && (!i1 || !i2 || !bad_flag || n_rows <= 0)) return NPLDA_ERR_BAD_ARG;
    if (!indexed && n > 0 && !x2) return NPLDA_ERR_BAD_ARG;
    if (impl != NPLDA_IMPL_AUTO && impl != NPLDA_IMPL_SIMT && impl != NPLDA_IMPL_TC && impl != NPLDA_IMPL_TC_F8 &&
        impl != NPLDA_IMPL_TC_BF16 && impl != NPLDA_IMPL_TC_PAIR && impl != NPLDA_IMPL_TC_PAIR_F8)
        return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    PackLayout L = make_pack_layout(d_in, d1, d2);
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)x1 & 15) == 0) && (indexed || ((uintptr_t)x2 & 15) == 0);
    const bool tc_ok = tc_shape_ok(dplda, L, indexed) && aligned;
    const bool want_tc = impl == NPLDA_IMPL_TC || impl == NPLDA_IMPL_TC_F8 || impl == NPLDA_IMPL_TC_BF16;
    // CTA-pair form of the bf16x3 kernel: from one 64-pair tile per SM on (below that the one-CTA form spreads over more SMs)
    const bool pair_ok = tc_ok && tcp_shape_ok(L);
    if ((impl == NPLDA_IMPL_TC_PAIR || impl == NPLDA_IMPL_TC_PAIR_F8) && !pair_ok) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (impl == NPLDA_IMPL_TC_PAIR || impl == NPLDA_IMPL_TC_PAIR_F8 ||
        (impl == NPLDA_IMPL_AUTO && pair_ok && n >= (int64_t)TILE_PAIRS * sm_count()))
        return score_tcp(x1, x2, n, L, (const char *)pack, scores, impl == NPLDA_IMPL_TC_PAIR_F8, st);
    if (want_tc && !tc_ok) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (want_tc || (impl == NPLDA_IMPL_AUTO && tc_ok))
        return score_tc(dplda, x1, x2, i1, i2, n_rows, bad_flag, n, L, (const char *)pack, scores,
                        impl == NPLDA_IMPL_TC_F8 ? 1 : (impl == NPLDA_IMPL_TC ? 2 : 0), st);   // AUTO: bf16x3 (fastest)
    return score_simt(dplda, x1, x2, i1, i2, n_rows, bad_flag, n, L, (const char *)pack, scores, st);
}

}  // namespace nplda

using namespace nplda;

extern "C" int nplda_version(void) { return 100; }

extern "C" int64_t nplda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char *nplda_error_string(int code) {
    switch (code) {
        case NPLDA_OK: return "ok";
        case NPLDA_ERR_BAD_ARG: return "nplda: bad argument (null pointer, negative size or misaligned buffer)";
        case NPLDA_ERR_UNSUPPORTED_DIM: return "nplda: unsupported layer dimensions for this kernel";
        case NPLDA_ERR_WORKSPACE: return "nplda: workspace too small";
        case NPLDA_ERR_NO_DEVICE: return "nplda: no usable sm_100 device";
        case NPLDA_ERR_IO: return "nplda: file cannot be opened, read or written";
        case NPLDA_ERR_FORMAT: return "nplda: trial file rows have differing numbers of fields";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "nplda: unknown error";
}

extern "C" int nplda_score_fwd(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                               const void *pack, float *scores, int impl, void *stream) {
    return score_dispatch(false, x1, x2, nullptr, nullptr, 0, nullptr, n, d_in, d1, d2, pack, scores, impl, stream);
}

extern "C" int dplda_score_fwd(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                               const void *pack, float *scores, int impl, void *stream) {
    return score_dispatch(true, x1, x2, nullptr, nullptr, 0, nullptr, n, d_in, d1, d1, pack, scores, impl, stream);
}

// Training forwards: scores plus the activations the backward needs, kept by the caller (act), so that the backward
// does not run the tensor-core kernel again.  NPLDA_ERR_UNSUPPORTED_DIM for shapes the tcgen05 kernel does not take.
constexpr int ACT_LD = 176;   // floats per saved activation row (EMIT_LD of score_tc.cu, the backward's row pitch)
// kind: 0 NeuralPlda (a, y), 1 DPlda with a trainable LDA (a, R u, Pm u), 2 DPlda with the LDA frozen (u)
extern "C" int64_t nplda_act_floats(int64_t n, int kind) {
    return n < 0 ? NPLDA_ERR_BAD_ARG : (kind == 1 ? 6 : (kind == 2 ? 2 : 4)) * n * ACT_LD;
}

extern "C" int nplda_score_fwd_train(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2, const void *pack,
                                     float *scores, float *act, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!x1 || !x2 || !scores || !act))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    const PackLayout L = make_pack_layout(d_in, d1, d2);
    if (!tc_shape_ok(false, L, false) || (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)act) & 15) != 0) return NPLDA_ERR_UNSUPPORTED_DIM;
    return score_tc(false, x1, x2, nullptr, nullptr, 0, nullptr, n, L, (const char *)pack, scores, 2, (cudaStream_t)stream,
                    act, act + 2 * n * ACT_LD, n);
}

extern "C" int dplda_score_fwd_train(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                                     float *scores, float *act, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!x1 || !x2 || !scores || !act))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d1)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    const PackLayout L = make_pack_layout(d_in, d1, d1);
    if (!tc_dplda_ok(L) || (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)act) & 15) != 0) return NPLDA_ERR_UNSUPPORTED_DIM;
    return dplda_score_tc_train(x1, x2, n, L, (const char *)pack, scores, act, (cudaStream_t)stream);
}

// Training forward of DPlda with the LDA frozen: scores + normalised rows u ([2 n][176] fp32, nplda_act_floats(n, 2)).
extern "C" int dplda_score_fwd_train_u(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                                       float *scores, float *urows, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!x1 || !x2 || !scores || !urows))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d1)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    const PackLayout L = make_pack_layout(d_in, d1, d1);
    if (!tc_dplda_ok(L) || (((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)urows) & 15) != 0) return NPLDA_ERR_UNSUPPORTED_DIM;
    return dplda_score_tc_train_u(x1, x2, n, L, (const char *)pack, scores, urows, (cudaStream_t)stream);
}

// Gradient of logistic_regres (weight [2 d1^2 + d1] = Wb | Ww | ws, bias) from the rows kept by dplda_score_fwd_train_u and
// dL/dS; ADDED into dw_lr / dc_lr (either may be null).
extern "C" int64_t dplda_lr_bwd_workspace_bytes(int64_t n, int d1) {
    if (n < 0 || d1 < 1 || d1 > 176) return NPLDA_ERR_BAD_ARG;
    return dplda_lr_workspace_bytes(n, d1);
}

extern "C" int dplda_lr_bwd(const float *urows, int64_t n, int d1, const float *dscores, float *dw_lr, float *dc_lr,
                            void *workspace, int64_t workspace_bytes, void *stream) {
    if (n < 0 || d1 < 1 || d1 > 176 || (n > 0 && (!urows || !dscores))) return NPLDA_ERR_BAD_ARG;
    if (n == 0 || (!dw_lr && !dc_lr)) return NPLDA_OK;
    return dplda_lr_grad(urows, n, n, d1, dscores, dw_lr, dc_lr, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dplda_score_fwd_ws(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                                  float *scores, int impl, void *workspace, int64_t workspace_bytes, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!x1 || !x2 || !scores))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d1)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    const PackLayout L = make_pack_layout(d_in, d1, d1);
    const bool aligned = ((((uintptr_t)x1) | ((uintptr_t)x2)) & 15) == 0;
    (void)workspace; (void)workspace_bytes;               // the fused tensor-core kernel needs no workspace
    const bool tc_ok = tc_dplda_ok(L) && aligned;
    if (impl == NPLDA_IMPL_TC && !tc_ok) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (impl == NPLDA_IMPL_TC || (impl == NPLDA_IMPL_AUTO && tc_ok && n >= 1024))
        return dplda_score_tc(x1, x2, n, L, (const char *)pack, scores, (cudaStream_t)stream);
    return score_dispatch(true, x1, x2, nullptr, nullptr, 0, nullptr, n, d_in, d1, d1, pack, scores,
                          impl == NPLDA_IMPL_SIMT ? NPLDA_IMPL_SIMT : NPLDA_IMPL_AUTO, stream);
}

extern "C" int nplda_score_fwd_indexed(const float *table, int64_t n_rows, const int64_t *idx1,
                                       const int64_t *idx2, int64_t n, int d_in, int d1, int d2,
                                       const void *pack, float *scores, int32_t *bad_index_flag,
                                       int impl, void *stream) {
    if (!idx1 || !idx2) return n == 0 ? NPLDA_OK : NPLDA_ERR_BAD_ARG;
    return score_dispatch(false, table, table, idx1, idx2, n_rows, bad_index_flag, n, d_in, d1, d2, pack, scores, impl, stream);
}

extern "C" int dplda_score_fwd_indexed(const float *table, int64_t n_rows, const int64_t *idx1,
                                       const int64_t *idx2, int64_t n, int d_in, int d1,
                                       const void *pack, float *scores, int32_t *bad_index_flag,
                                       int impl, void *stream) {
    if (!idx1 || !idx2) return n == 0 ? NPLDA_OK : NPLDA_ERR_BAD_ARG;
    return score_dispatch(true, table, table, idx1, idx2, n_rows, bad_index_flag, n, d_in, d1, d1, pack, scores, impl, stream);
}

extern "C" int64_t nplda_split_bytes(int64_t n_rows, int d_in) {
    if (n_rows < 0 || d_in <= 0) return NPLDA_ERR_BAD_ARG;
    if (d_in % 32 != 0) return NPLDA_ERR_UNSUPPORTED_DIM;
    return n_rows * (int64_t)d_in * 4 + 128;     // rows + header {max|x|, 2^kx, 2^-kx}
}

extern "C" int nplda_table_split(const float *table, int64_t n_rows, int d_in, void *split, void *stream) {
    if (n_rows < 0 || (n_rows > 0 && (!table || !split))) return NPLDA_ERR_BAD_ARG;
    if ((((uintptr_t)table) & 15) != 0 || (((uintptr_t)split) & 127) != 0) return NPLDA_ERR_BAD_ARG;
    if (n_rows == 0) return NPLDA_OK;
    return table_split(table, n_rows, d_in, split, (cudaStream_t)stream);
}

extern "C" int nplda_score_fwd_split(const void *split, int64_t n_rows, const int64_t *idx1, const int64_t *idx2, int64_t n,
                                     int d_in, int d1, int d2, const void *pack, float *scores, int32_t *bad_index_flag,
                                     void *stream) {
    if (n < 0 || !pack || (n > 0 && (!split || !idx1 || !idx2 || !scores || !bad_index_flag || n_rows <= 0))) return NPLDA_ERR_BAD_ARG;
    if ((((uintptr_t)split) & 127) != 0) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    const PackLayout L = make_pack_layout(d_in, d1, d2);
    return score_tcx(split, n_rows, idx1, idx2, n, L, (const char *)pack, scores, bad_index_flag, (cudaStream_t)stream);
}

extern "C" int nplda_embed_fwd(const float *x, int64_t n, int d_in, int d1, int d2, const void *pack,
                               float *emb, int is_dplda, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!x || !emb))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    PackLayout L = make_pack_layout(d_in, d1, d2);
    return simt_aux(is_dplda ? 3 : 2, x, nullptr, n, L, (const char *)pack, emb, 0, (cudaStream_t)stream);
}

extern "C" int nplda_score_from_embeddings(const float *y1, const float *y2, int64_t n, int d2,
                                           const float *p_sqrt, const float *q, float *scores, void *stream) {
    if (n < 0 || d2 <= 0 || (n > 0 && (!y1 || !y2 || !p_sqrt || !q || !scores))) return NPLDA_ERR_BAD_ARG;
    if (n == 0) return NPLDA_OK;
    const int grid = (int)std::min<int64_t>((n + 7) / 8, 8 * (int64_t)sm_count());
    score_from_emb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y1, y2, n, d2, p_sqrt, q, scores);
    NPLDA_LAUNCH_CHECK();
    return NPLDA_OK;
}

extern "C" int dplda_score_from_embeddings(const float *u1, const float *u2, int64_t n, int d_in, int d1,
                                           const void *pack, float *scores, void *stream) {
    if (n < 0 || !pack || (n > 0 && (!u1 || !u2 || !scores))) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d1)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (n == 0) return NPLDA_OK;
    PackLayout L = make_pack_layout(d_in, d1, d1);
    return simt_aux(4, u1, u2, n, L, (const char *)pack, scores, 0, (cudaStream_t)stream);
}

// ---- host-buffer entry -----------------------------------------------------------
// Two-deep pipeline: while chunk c is scored on stream `comp`, chunk c+1 is in
// flight host->device on stream `copy`.  Device scratch holds two chunks of
// both sides plus two score buffers.
extern "C" int64_t nplda_host_scratch_bytes(int64_t chunk_pairs, int d_in) {
    if (chunk_pairs <= 0 || d_in <= 0) return NPLDA_ERR_BAD_ARG;
    int64_t side = (chunk_pairs * d_in * 4 + 255) / 256 * 256;
    int64_t sc = (chunk_pairs * 4 + 255) / 256 * 256;
    return 2 * (2 * side + sc);
}

extern "C" int nplda_score_fwd_host(const float *x1_host, const float *x2_host, int64_t n, int d_in,
                                    int d1, int d2, const void *pack, float *scores_host,
                                    int64_t chunk_pairs, void *dev_scratch, int64_t dev_scratch_bytes,
                                    int is_dplda, int impl) {
    if (n < 0 || chunk_pairs <= 0 || !pack || !dev_scratch) return NPLDA_ERR_BAD_ARG;
    if (n > 0 && (!x1_host || !x2_host || !scores_host)) return NPLDA_ERR_BAD_ARG;
    if (!dims_supported(d_in, d1, d2)) return NPLDA_ERR_UNSUPPORTED_DIM;
    if (dev_scratch_bytes < nplda_host_scratch_bytes(chunk_pairs, d_in)) return NPLDA_ERR_WORKSPACE;
    if (n == 0) return NPLDA_OK;
    const int64_t side = (chunk_pairs * d_in * 4 + 255) / 256 * 256;
    const int64_t sc = (chunk_pairs * 4 + 255) / 256 * 256;
    char *base = (char *)dev_scratch;
    float *dx1[2], *dx2[2], *ds[2];
    for (int b = 0; b < 2; ++b) {
        char *p = base + b * (2 * side + sc);
        dx1[b] = (float *)p; dx2[b] = (float *)(p + side); ds[b] = (float *)(p + 2 * side);
    }
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    int rc = NPLDA_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == NPLDA_OK) rc = (int)e; return e != cudaSuccess; };
    if (fail(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking)) ||
        fail(cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking))) goto done;
    for (int b = 0; b < 2; ++b)
        if (fail(cudaEventCreateWithFlags(&ready[b], cudaEventDisableTiming)) ||
            fail(cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming))) goto done;
    {
        const int64_t nchunks = (n + chunk_pairs - 1) / chunk_pairs;
        for (int64_t c = 0; c < nchunks && rc == NPLDA_OK; ++c) {
            const int b = (int)(c & 1);
            const int64_t p0 = c * chunk_pairs, m = std::min(chunk_pairs, n - p0);
            if (c >= 2 && fail(cudaStreamWaitEvent(copy, freed[b], 0))) break;   // buffer b drained
            if (fail(cudaMemcpyAsync(dx1[b], x1_host + p0 * d_in, (size_t)m * d_in * 4, cudaMemcpyHostToDevice, copy)) ||
                fail(cudaMemcpyAsync(dx2[b], x2_host + p0 * d_in, (size_t)m * d_in * 4, cudaMemcpyHostToDevice, copy)) ||
                fail(cudaEventRecord(ready[b], copy)) || fail(cudaStreamWaitEvent(comp, ready[b], 0)))
                break;
            int r = score_dispatch(is_dplda != 0, dx1[b], dx2[b], nullptr, nullptr, 0, nullptr, m, d_in, d1,
                                   d2, pack, ds[b], impl, comp);
            if (r != NPLDA_OK) { rc = r; break; }
            if (fail(cudaMemcpyAsync(scores_host + p0, ds[b], (size_t)m * 4, cudaMemcpyDeviceToHost, comp)) ||
                fail(cudaEventRecord(freed[b], comp)))
                break;
        }
        fail(cudaStreamSynchronize(copy));
        fail(cudaStreamSynchronize(comp));
    }
done:
    for (int b = 0; b < 2; ++b) {
        if (ready[b]) cudaEventDestroy(ready[b]);
        if (freed[b]) cudaEventDestroy(freed[b]);
    }
    if (copy) cudaStreamDestroy(copy);
    if (comp) cudaStreamDestroy(comp);
    return rc;
}
