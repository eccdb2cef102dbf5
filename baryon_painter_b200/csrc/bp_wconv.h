// Window-GEMM convolution engine of the 16-bit path (bp_wconv.cu) -- interface used by bp_v2.cu.
//
// Every convolution of the paint path is lowered on the host to one "window GEMM":
//   D[(R, U)][n] = sum over taps (dl, du), unit elements e:  X[line R*Jy + dl][unit U + du][e] * B[tap][e][n]
// where X is a 16-bit NHWC activation whose lines are cut into *units* of G stored pixels
// (unit = G * Cs channels, contiguous in memory), (R, U) enumerates output block-lines / units and
// column n enumerates (output row jy, output pixel jx, channel) of the Jy x G pixel block the
// M row produces.  Plain stride-1 convolutions, k = 2s strided convolutions (on the shifted
// space-to-depth layout their producer writes) and the sub-pixel phases of transposed convolutions
// all take this form; the Toeplitz expansion of the weights for G > 1 / Jy > 1 happens at pack time.
#pragma once

#include <functional>
#include <vector>

#include "bp_common.h"

namespace bp {

// a 16-bit NHWC activation tensor in device memory: [n][Hs][Ws][b*b][Cp]
//   b == 1: plain (Hs = H, Ws = W)
//   b  > 1: shifted space-to-depth for a k = 2b, stride b, pad b/2 consumer:
//           pixel (y, x) lives in block ((y + b/2) / b, (x + b/2) / b), sub-pixel ((y + b/2) % b, (x + b/2) % b);
//           Hs = H / b + 1; never-written border sub-pixels stay zero (they are the padding)
// or (f32) a dense fp32 NCHW tensor [n][C][H][W].
// split: every value is stored as two fp16 numbers, hi = fp16(v) and lo = fp16(v - hi) (22 significant bits); a pixel
// holds [Cp hi | Cp lo] channels.  This is the operand format of the fp32-accurate tensor-core path.
struct ActDesc {
  void* ptr = nullptr;
  int C = 0, Cp = 0, H = 0, W = 0, b = 1;
  bool f32 = false;
  bool split = false;
  int sexp = 0;                          // split only: stored value = true value * 2^sexp (keeps the lo halves of ordinary
                                         // activations out of fp16's subnormal range; exact, undone by the consumer)
  long long f32_bs = 0;                  // f32 only: per-sample stride in elements when the tensor is a view (0: dense)
  int Hs() const { return b > 1 ? H / b + 1 : H; }
  int Ws() const { return b > 1 ? W / b + 1 : W; }
  int Cpix() const { return split ? 2 * Cp : Cp; }          // stored 16-bit channels per pixel
  size_t elems_per_sample() const {
    return f32 ? (f32_bs ? (size_t)f32_bs : (size_t)C * H * W) : (size_t)Hs() * Ws() * b * b * Cpix();
  }
  size_t bytes_per_sample() const { return elems_per_sample() * (f32 ? 4 : 2); }
};

struct WTap { int dl, du; };
struct WSegOff { int oy, ox; };

// M-tile shapes.  W_FLAT: 128 consecutive positions of the (halo-padded) flat line domain; W_LINE: one line of <= 128
// units; W_BLOCK: 8 units x 16 block-lines -- the sixteen 8-row groups of the UMMA A operand sit on sixteen different
// patch lines (descriptor SBO = one block-line of the patch), so a wide-halo packed layer (conv7 16->8 with G = 4)
// needs a (8 T_r + halo) x (16 Jy + halo) patch instead of (T_r Jy + halo) full-width lines
enum { W_FLAT = 0, W_LINE = 1, W_BLOCK = 2 };

struct WSpec {
  // input view: lines of `Wu` units, each unit `unit_elems` 16-bit elements
  int G = 1, Jy = 1, mode = W_FLAT;
  int OHl = 0, OWl = 0;                  // M domain: block lines, units per line
  int nphase = 1;
  std::vector<WTap> taps[kMaxPhases];
  std::vector<WSegOff> segs[kMaxPhases];
  int N = 0;                             // GEMM N (multiple of 16)
  int seg_len = 0, seg_valid = 0;        // columns per segment / stored channels per segment
  int ry = 1, rx = 1;                    // output pixel of M row (R, U), segment (oy, ox): (R*ry + oy, U*rx + ox)
  // fp32 weight of (phase, tap, element of the unit, column), BN scale already folded in; *part = 0 / 1 when the
  // element is the hi / lo half of a split-precision value.  split: the builder stores w as (hi, lo) fp16 pairs and
  // runs two passes over the input slices: pass 0 meets w_hi with both halves, pass 1 meets w_lo with the hi half
  bool split = false;
  std::function<float(int, int, int, int, int*)> weight;
  std::vector<float> shift;              // [N] per column
  int act = 0;
  float act_param = 0.f;
  int fmt = 0;
};

struct WLayer;   // opaque (bp_wconv.cu)

// tiling of a formulation: strip width (units), M-tiles per region, tap lines per weight stage, weight-ring depth
struct WTiling { int Wt = 0, T_r = 0, gl = 0, nbst = 0; };

// builds the packed weights / k-step tables / TMA tensor map for `spec` reading from `in`
// (whose device pointer must already be final) with capacity for `nb_max` samples
// rank: 0 = the tiling the cost model likes best, 1 = its second choice, ... (BP_E_UNSUPPORTED past the last);
// forced: build exactly this tiling (BP_E_UNSUPPORTED if it does not fit)
int wconv_build(const WSpec& spec, const ActDesc& in, int nb_max, WLayer** out, int rank = 0, const WTiling* forced = nullptr);
void wconv_tiling(const WLayer* w, WTiling* t);
void wconv_free(WLayer* w);
int wconv_launch(const WLayer* w, const ActDesc& out, const void* skip, int nb, cudaStream_t s, bool reverse = false);
int wconv_mma_count(const WLayer* w, int nb, double* cycles_floor);

// lowering (bp_v2.cu)
int v2_padc(int c);
bool v2_eligible(const Layer& l, int* need_b, bool split = false);
int v2_make_spec(const Layer& l, int fmt, bool split, WSpec* sp);
int v2_candidates(const Layer& l, int fmt, int Cp, bool split, std::vector<WSpec>* out);

// layout conversion kernels
int launch_nchw32_to_nhwc16(const float* in, long long in_bs, const ActDesc& out, int nb, int fmt, cudaStream_t s);
int launch_nhwc16_to_nchw32(const ActDesc& in, float* out, long long out_bs, int nb, int fmt, cudaStream_t s);

}  // namespace bp
