// placeholder until the tcgen05 path lands
#include "bp_bf16.h"
namespace bp {
int bf16_prepare_net(const std::vector<std::vector<Layer>*>&, int, void** state) {
  *state = nullptr;
  set_error("bf16 path not built yet");
  return BP_E_UNSUPPORTED;
}
void bf16_free_net(void*) {}
void bf16_free_layer(Layer*) {}
int bf16_run_stack(void*, int, std::vector<Layer>&, const float*, long long, float*, long long, int, const float*,
                   float, float, int, cudaStream_t, float**, float* const*, std::vector<float*>*, int) {
  set_error("bf16 path not built yet");
  return BP_E_UNSUPPORTED;
}
}  // namespace bp
