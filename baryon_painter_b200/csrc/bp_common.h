// Shared declarations of the baryon_painter_b200 CUDA library (internal; the public ABI is
// include/baryon_painter_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/baryon_painter_b200.h"

namespace bp {

// 16-bit operand formats of the tensor-core path (tcgen05 .kind::f16 a/b format field)
enum { TC_FMT_F16 = 0, TC_FMT_BF16 = 1 };

// Developer switches (BP_V2_*, BP_WIN_TIMING, ...) change formulations, disable work or print timings; they exist
// only in -DBP_DEVELOPER builds (`python -m baryon_painter_b200.build --developer`).  The release library reads
// three documented runtime knobs, none of which changes a painted value: BP_CHUNK, BP_HOST_STEP, BP_HOST_EDGE.
#ifdef BP_DEVELOPER
inline const char* dev_env(const char* name) { return getenv(name); }
#else
inline const char* dev_env(const char*) { return nullptr; }
#endif

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int64_t& launch_counter();

#define BP_CUDA_TRY(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      bp::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return BP_E_CUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define BP_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      bp::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// ---- one dense GEMM view of a (phase of a) convolution -----------------------------------------
// A convolution is executed as an implicit GEMM  D[m][n] = sum_k A[m][k] * B[k][n]  where
//   m  enumerates (sample, i, j) over the logical output grid OH x OW of one phase,
//   n  is the output channel,
//   k  enumerates the (input channel, tap) pairs listed in `ktab`.
// Input pixel of (i, j, k):  (i*istride + dr_k, j*istride + ds_k), zero outside the image.
// Output pixel of (i, j):    (i*os + ph, j*os + pw) in the full OHF x OWF grid.
// A strided convolution has istride = stride, os = 1 and one phase; a transposed convolution
// is split into stride^2 output phases with istride = 1, os = stride (sub-pixel decomposition).
struct PhaseDev {
  int k_begin;  // first row of this phase in ktab / wmat
  int K;        // rows
  int ph, pw;   // output phase offsets
};

constexpr int kMaxPhases = 16;

enum { POST_NONE = 0, POST_INV_SHIFT_LOG = 1 };

struct ConvArgs {
  const float* in;
  float* out;
  const float* skip;
  long long in_bs, out_bs, skip_bs;  // per-sample strides (elements); channel stride is H*W
  const int4* ktab;                  // {c*H*W + dr*W + ds, dr, ds, c}
  const float* wmat;                 // [rows][npad] fp32
  const float* scale;                // [cout]
  const float* shift;                // [cout]
  const float* post_sigma;           // [n] per-sample sigma of the fused inverse transform
  int H, W, OH, OW, OHF, OWF;
  int istride, os;
  int cout, npad;
  int nphase;
  PhaseDev phase[kMaxPhases];
  int nb;  // samples in this launch
  int act;
  float act_param;
  int post;
  float post_k, post_shift;
};

// ---- host-side packed layer ------------------------------------------------------------------
struct Layer {
  bp_layer_desc d;              // weight/scale/shift pointers are NOT retained
  int H = 0, W = 0;             // input spatial dims
  int OHF = 0, OWF = 0;         // output spatial dims
  int OH = 0, OW = 0;           // per-phase logical output dims
  int istride = 1, os = 1;
  int npad = 0;                 // padded GEMM N of the fp32 weight matrix
  int nphase = 1;
  PhaseDev phase[kMaxPhases];
  int Kmax = 0;
  // device copies
  int4* ktab = nullptr;
  float* wmat = nullptr;
  float* scale = nullptr;
  float* shift = nullptr;
  bool v2 = false;               // runs on the window-GEMM engine (bp_wconv.cu)
  std::vector<float> host_weight;  // PyTorch-layout copy kept for the tensor-core packing
  std::vector<float> host_scale, host_shift;
  double flops = 0;             // 2*MACs per sample (SURVEY App. A counting)
};

int pack_layer(const bp_layer_desc& d, int H, int W, Layer* out);
void free_layer(Layer* l);

// ---- kernels (bp_f32.cu) ------------------------------------------------------------------------
int launch_conv_f32(const Layer& l, ConvArgs& a, cudaStream_t s);
int launch_prepare(const float* tiles, float* dst, long long dst_bs, int y_channel, int aux_channel,
                   const float* sigma_in, const float* aux, float k_in, float shift_in, int do_transform,
                   int nb, int hw, cudaStream_t s);
int launch_sample_z(const float* prior_out, const float* eps, float* latent, float* mu_out, float* lv_out,
                    float min_z_var, int nb, int hw, int mode, uint64_t seed, uint64_t offset, cudaStream_t s);
int launch_welford(const float* x, double* mean, double* m2, int count, size_t n, cudaStream_t s);
int launch_var_finalize(const double* mean, const double* m2, float* mean_out, float* var_out, int count, size_t n,
                        cudaStream_t s);
int launch_rng_normal(float* out, uint64_t seed, uint64_t offset, size_t n, cudaStream_t s);
int launch_kl_sum(const float* q, const float* prior, int nb, int hw, double* dst, cudaStream_t s);
int launch_sqdiff_sum(const float* a, const float* b, size_t n, double* dst, cudaStream_t s);

}  // namespace bp
