// Tensor-core (tcgen05) convolution path -- interface used by bp_net.cu.
#pragma once

#include "bp_common.h"

namespace bp {

// 16-bit operand formats of the tensor-core path (tcgen05 .kind::f16 a/b format field)
enum { TC_FMT_F16 = 0, TC_FMT_BF16 = 1 };

// can this layer run on the tensor-core kernel?  `first_in_sequence`: its input is packed from fp32
// by launch_pack_c8, so a channel count that is not a multiple of 8 is zero-padded.
bool tc_layer_eligible(const bp_layer_desc& d, bool first_in_sequence);
int tc_pack_layer(Layer* l, int fmt);
void tc_free_layer(Layer* l);

// `in`/`skip`/`out16`: 16-bit c8 tensors [nb][C/8][H][W][8]; exactly one of out16 / out32 is non-null
// (out32: fp32 NCHW with per-sample stride out32_bs).
int launch_conv_tc(const Layer& l, const void* in, void* out16, float* out32, long long out32_bs, const void* skip,
                   int nb, cudaStream_t s);
// windowed kernel (bp_win.cu): stride-1 convolutions and transposed-convolution phases without gathers
bool win_layer_eligible(const bp_layer_desc& d, bool first_in_sequence);
int win_pack_layer(Layer* l, int fmt);
void win_free_layer(Layer* l);
int launch_conv_win(const Layer& l, const void* in, void* out16, float* out32, long long out32_bs, const void* skip,
                    int nb, cudaStream_t s);
int launch_pack_c8(const float* in, long long in_bs, int C, int hw, void* out, int nb, int fmt, cudaStream_t s);
int launch_unpack_c8(const void* in, int C, int hw, float* out, long long out_bs, int nb, int fmt, cudaStream_t s);

}  // namespace bp
