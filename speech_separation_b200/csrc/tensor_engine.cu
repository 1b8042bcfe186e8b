#include "tensor_engine.cuh"

namespace vatss {

bool tensor_engine_supports(const vatss_model_desc* d) { (void)d; return false; }
size_t tensor_engine_packed_bytes(const vatss_model_desc* d) { (void)d; return 0; }
size_t tensor_engine_workspace_bytes(const vatss_model_desc*, int, int, int, int, int) { return 0; }
int tensor_engine_pack(const vatss_model_desc*, const float* const*, void*, cudaStream_t) { return 0; }
int tensor_engine_forward(const vatss_model_desc*, const float* const*, const void*, const float*, const float*,
                          const float*, int, int, int, int, int, float*, float*, void*, cudaStream_t) {
  set_error("tensor engine not built");
  return -1;
}

}  // namespace vatss
