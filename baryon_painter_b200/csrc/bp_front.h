// Front (transform + p_z_in + NHWC packing) and tail (last stencil + inverse transform) kernels of the
// 16-bit path -- bp_front.cu.
#pragma once

#include <algorithm>
#include <vector>

#include "bp_wconv.h"

namespace bp {

// p_z_in: up to 4 single-channel transposed convolutions (k = 2s, p = s/2, k <= 8) with folded BN + activation
struct PzParams {
  int nl;
  int k[4], s[4], p[4], act[4];
  float scale[4], shift[4], act_param[4];
  float w[4][64];
};

// last layer: single-channel k x k convolution + activation + optional inverse transform
struct TailParams {
  int k;
  float w[49];
  float scale, shift;
  int act;
  float act_param;
  int post;
  float post_k, post_shift;
  int precise;               // fp32-accurate path: library expf / log1pf instead of the fast intrinsics
};

// prior_network.0 fused into the front pass: k4 s2 p1 convolution of the 2-channel [y, z] input to <= 8 channels
struct FrontConvParams {
  int cout, act;
  float act_param;
  float w[8][2][16];          // [co][ci][r*4 + q], BN scale folded in
  float shift[8];
  float zsum[3][3][8];        // sum of the z-channel taps inside the image, by (row, column) border class
};

int launch_front_prior_conv(const float* tiles, const ActDesc& out, const float* sigma, const float* aux,
                            const FrontConvParams& fc, float k_in, float shift_in, int do_transform, int H, int W, int nb,
                            int fmt, cudaStream_t s);
int launch_front_prior(const float* tiles, const ActDesc& out, const float* sigma, const float* aux, float k_in,
                       float shift_in, int do_transform, int nb, int fmt, cudaStream_t s);
int launch_front_latent(const float* tiles, const float* latent, const ActDesc& out, const float* sigma, const float* aux,
                        const PzParams& pz, float k_in, float shift_in, int do_transform, int lh, int lw, int nb, int fmt,
                        cudaStream_t s);
int launch_tail_stencil(const float* in, float* out, long long out_bs, const TailParams& tp, const float* post_sigma, int H,
                        int W, int nb, cudaStream_t s);

}  // namespace bp
