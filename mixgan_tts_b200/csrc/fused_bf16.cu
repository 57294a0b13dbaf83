// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace mgb {
size_t bf16_packed_bytes(const mgb_model_dims&) { return 0; }
int bf16_pack(const mgb_model_dims&, const float*, void*, cudaStream_t) {
  set_error("bf16 path not built"); return MGB_E_UNSUPPORTED; }
size_t bf16_workspace_bytes(const mgb_model_dims&, int, int, int) { return 0; }
int bf16_denoiser(const mgb_model_dims&, const void*, const float*, const int64_t*, const float*, const float*,
                  const float*, const float*, int, int, float*, float*, int, int, void*, bool, cudaStream_t) {
  set_error("bf16 path not built"); return MGB_E_UNSUPPORTED; }
}
