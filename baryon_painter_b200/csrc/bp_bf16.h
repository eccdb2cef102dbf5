// bf16 tensor-core (tcgen05) execution path -- interface used by bp_net.cu.
#pragma once

#include <vector>

#include "bp_common.h"

namespace bp {

// Build the bf16 execution state of a network: packed bf16 weights per layer, activation
// workspaces for `chunk` samples.
int bf16_prepare_net(const std::vector<std::vector<Layer>*>& stacks, int chunk, void** state);
void bf16_free_net(void* state);
void bf16_free_layer(Layer* l);

// Run sub-network `sidx` in bf16.  Same contract as run_stack() in bp_net.cu: fp32 NCHW input with
// per-sample stride in_bs; the last layer writes fp32 NCHW to final_out (stride final_bs) when
// given, otherwise to a pool buffer returned in *result.
int bf16_run_stack(void* state, int sidx, std::vector<Layer>& layers, const float* in, long long in_bs,
                   float* final_out, long long final_bs, int post, const float* post_sigma, float post_k,
                   float post_shift, int nb, cudaStream_t s, float** result, float* const* pool,
                   std::vector<float*>* dbg, int chunk);

}  // namespace bp
