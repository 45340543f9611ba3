// tcgen05 (sm_100a) bf16 implicit-GEMM kernel of the cubed-sphere convolution -- placeholder until the kernel lands.
#include "cs_common.cuh"

namespace dlwpcs {

bool tc_supported(const dlwpcs_conv_desc *, const Geometry &, const char **why) {
  *why = "tensor-core kernel not built";
  return false;
}
int64_t tc_packed_weight_bytes(const dlwpcs_conv_desc *, const Geometry &) { return -1; }
int tc_pack_weights(const dlwpcs_conv_desc *, const Geometry &, const dlwpcs_conv_weights *, void *, cudaStream_t) {
  set_error("tensor-core kernel not built");
  return 1;
}
int tc_conv_fwd(const dlwpcs_conv_desc *, const Geometry &, const void *, const void *, const void *, void *,
                cudaStream_t) {
  set_error("tensor-core kernel not built");
  return 1;
}

}  // namespace dlwpcs
