// azg_gemm_tc.cu -- placeholder until the tcgen05 path lands (next commit).
#include "azg_common.cuh"

size_t azg_tc_scratch_bytes(int, int64_t, int) { return 0; }

int azg_tc_output_transform(const void*, int, int, const float*, const float*, const float*, float*, int64_t, void*,
                            size_t, cudaStream_t) {
  azg_set_error("tcgen05 path not built");
  return AZG_ERR_INVALID;
}

extern "C" {
size_t azg_c4_packed_bytes(int, int) { return 0; }
int azg_c4_pack_gnn(const float*, const float*, int, int, void*, size_t, azg_stream) {
  azg_set_error("tcgen05 path not built");
  return AZG_ERR_INVALID;
}
}
