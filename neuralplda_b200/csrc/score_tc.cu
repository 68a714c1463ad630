// K1 (tensor cores): tcgen05 split-bf16 score kernel.  Placeholder until the
// kernel lands: reports "shape unsupported" so NPLDA_IMPL_AUTO uses the SIMT path.
#include "common.cuh"

namespace nplda {

int64_t tc_image_bytes(int, int, int) { return 0; }
bool tc_shape_ok(bool, const PackLayout &, bool) { return false; }

int tc_pack_nplda(const float *, const float *, const float *, const float *, const float *,
                  const float *, const PackLayout &, char *, cudaStream_t) { return NPLDA_OK; }
int tc_pack_dplda(const float *, const float *, const float *, const float *, const PackLayout &,
                  char *, cudaStream_t) { return NPLDA_OK; }

int score_tc(bool, const float *, const float *, const int64_t *, const int64_t *, int64_t, int32_t *,
             int64_t, const PackLayout &, const char *, float *, cudaStream_t) {
    return NPLDA_ERR_UNSUPPORTED_DIM;
}

}  // namespace nplda
