// Elementwise / stencil ends of the 16-bit paint path (HBM-bound passes, coalesced and vectorised):
//
//   front_prior_kernel    raw DM tile -> forward transform y = ln(x/sigma + 1)/k - shift  -> [y, z] written in the
//                         shifted space-to-depth NHWC layout prior_network.0 (k4 s2) reads
//                         (reference data_transforms.py:66-76 + models/utils.py:159-182 merge_aux_label)
//   front_latent_kernel   latent (h x w) -> p_z_in (three 1->1 transposed convolutions + BN + ReLU, fp32, whole
//                         pyramid recomputed per CTA in shared memory) and, fused, the forward transform again:
//                         writes the decoder input pixel [p_z_in(latent), y, z, 0] as one 8-byte NHWC store
//                         (reference cvae.py:104-109: merge_aux_label + p_z_in + torch.cat)
//   tail_stencil_kernel   last 1->1 convolution + activation (Softplus) + inverse transform
//                         (exp((x + shift)*k) - 1)*sigma  -> fp32 tile  (reference data_transforms.py:88-98)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "bp_front.h"

namespace bp {

__device__ __forceinline__ uint16_t f_to16(float v, int fmt) {
  if (fmt == 0) {
    __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// forward transform y = ln(x/sigma + 1)/k - shift = lg2(fma(x, 1/sigma, 1)) * (ln 2 / k) - shift.  These kernels only feed
// the 16-bit path, so the approximate base-2 logarithm (2^-22 absolute) is ample; `ikl2` = ln 2 / k.  The argument is
// >= 1 for every physical (non-negative) density, so the subnormal handling of __logf is not needed.
__device__ __forceinline__ float f_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float f_transform(float x, float inv_sigma, float ikl2, float shift, int on) {
  return on ? fmaf(f_lg2(fmaf(x, inv_sigma, 1.f)), ikl2, -shift) : x;
}
constexpr float kLn2 = 0.69314718055994531f, kLog2e = 1.44269504088896341f;
__device__ __forceinline__ float f_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float f_act(float v, int act, float p) {
  switch (act) {
    case BP_ACT_RELU: return fmaxf(v, 0.f);
    case BP_ACT_LEAKY:
    case BP_ACT_PRELU: return v >= 0.f ? v : v * p;
    case BP_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));
    case BP_ACT_TANH: return tanhf(v);
    case BP_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// ---- prior input: [y, z, 0, 0] per pixel, space-to-depth block 2 (pixel (y,x) -> block ((y+1)/2, (x+1)/2)) ----
__global__ void front_prior_kernel(const float* __restrict__ tiles, uint2* __restrict__ out, const float* __restrict__ sigma,
                                   const float* __restrict__ aux, float k_in, float shift_in, int do_t, int H, int W, int b,
                                   int fmt) {
  const int n = blockIdx.z;
  const float sg = do_t ? 1.f / sigma[n] : 1.f, ik = kLn2 / k_in;
  const uint16_t z16 = f_to16(aux[n], fmt);
  const int hw = H * W;
  const int Hs = b > 1 ? H / b + 1 : H, Ws = b > 1 ? W / b + 1 : W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const int y = p / W, x = p - y * W;
    const float v = f_transform(tiles[(size_t)n * hw + p], sg, ik, shift_in, do_t);
    size_t o;
    if (b == 1) {
      o = (size_t)n * hw + p;
    } else {
      const int yy = y + (b >> 1), xx = x + (b >> 1);
      const int by = yy / b, sy = yy - by * b, bx = xx / b, sx = xx - bx * b;
      o = ((((size_t)n * Hs + by) * Ws + bx) * b + sy) * b + sx;
    }
    out[o] = make_uint2((uint32_t)f_to16(v, fmt) | ((uint32_t)z16 << 16), 0u);
  }
}

int launch_front_prior(const float* tiles, const ActDesc& out, const float* sigma, const float* aux, float k_in,
                       float shift_in, int do_transform, int nb, int fmt, cudaStream_t s) {
  BP_REQUIRE(out.Cp == 4 && !out.f32, BP_E_INVALID, "front_prior: output must be a 4-channel NHWC tensor");
  const int hw = out.H * out.W;
  const int bx = std::min(256, (hw + 255) / 256);
  front_prior_kernel<<<dim3(bx, 1, nb), 256, 0, s>>>(tiles, static_cast<uint2*>(out.ptr), sigma, aux, k_in, shift_in,
                                                     do_transform, out.H, out.W, out.b, fmt);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ---- prior input fused with prior_network.0 (k4 s2 p1, 2 -> <= 8 channels, BN + activation) -----------------
// CTA = 32 x 32 output pixels: the 66 x 66 window of raw densities is transformed once into shared memory.
// A thread owns 4 vertically adjacent outputs x 8 channels: its 10 x 4 input window sits in registers as
// column pairs (64-bit shared loads), the weights come as matching pairs from shared memory and every
// fma.rn.f32x2 advances one (pixel, channel) accumulator by two taps; the two halves are summed at the end.
// The constant z plane contributes z * (sum of the in-image taps' weights): a per-channel constant away from
// the image border.
__device__ __forceinline__ unsigned long long f_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long f_pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
// two fp32 -> one packed 16-bit pair (lo in the low half), saturating to the finite range
__device__ __forceinline__ uint32_t f_pack16(float lo, float hi, int fmt) {
  uint32_t r;
  if (fmt == 0) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f_sum2(unsigned long long v) {
  return __uint_as_float((uint32_t)v) + __uint_as_float((uint32_t)(v >> 32));
}

__global__ void __launch_bounds__(256, 2) front_prior_conv_kernel(const float* __restrict__ tiles, uint16_t* __restrict__ out,
                                                                  const float* __restrict__ sigma, const float* __restrict__ aux,
                                                                  const FrontConvParams fc, float k_in, float shift_in, int do_t,
                                                                  int H, int W, int ob, int oCp, int fmt) {
  __shared__ __align__(16) float sy[66][68];
  __shared__ __align__(16) unsigned long long sw[4][2][8];     // [r][column pair][co] = {w(r, 2qp), w(r, 2qp + 1)}
  __shared__ float szs[3][3][8];          // z tap sums by (row class, column class): first / interior / last
  const int n = blockIdx.z;
  const int OH = H >> 1, OW = W >> 1;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const float sg = do_t ? 1.f / sigma[n] : 1.f, ik = kLn2 / k_in;
  const float z = aux[n];
  const float* src = tiles + (size_t)n * H * W;
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  if (tid < 64) {
    const int r = tid >> 4, qp = (tid >> 3) & 1, co = tid & 7;
    sw[r][qp][co] = f_pack2(fc.w[co][0][r * 4 + 2 * qp], fc.w[co][0][r * 4 + 2 * qp + 1]);
  } else if (tid < 64 + 72) {
    const int t = tid - 64;
    (&szs[0][0][0])[t] = (&fc.zsum[0][0][0])[t];
  }
  if ((W & 3) == 0 && 2 * j0 + 64 <= W) {
    // window = columns 2 j0 - 1 .. 2 j0 + 64: the 64 interior columns come as aligned float4 (16 per row), the two
    // halo columns as scalars; all of a thread's loads are issued before the first is consumed (two CTAs per SM:
    // the latency has to be covered inside the thread)
    constexpr int NV = (66 * 16 + 255) / 256;
    float4 raw[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int t = tid + k * 256, ly = t >> 4, q = t & 15;
      const int y = 2 * i0 - 1 + ly;
      raw[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ly < 66 && y >= 0 && y < H) raw[k] = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * W + 2 * j0) + q);
    }
    float hv = 0.f;
    const int hly = tid >> 1, hlx = (tid & 1) * 65;
    const int hy = 2 * i0 - 1 + hly, hx = 2 * j0 - 1 + hlx;
    const bool hok = hly < 66 && hy >= 0 && hy < H && hx >= 0 && hx < W;
    if (hok) hv = __ldg(src + (size_t)hy * W + hx);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int t = tid + k * 256, ly = t >> 4, q = t & 15;
      const int y = 2 * i0 - 1 + ly;
      if (ly < 66) {
        const bool ok = y >= 0 && y < H;
        float* d = &sy[ly][1 + 4 * q];
        d[0] = ok ? f_transform(raw[k].x, sg, ik, shift_in, do_t) : 0.f;
        d[1] = ok ? f_transform(raw[k].y, sg, ik, shift_in, do_t) : 0.f;
        d[2] = ok ? f_transform(raw[k].z, sg, ik, shift_in, do_t) : 0.f;
        d[3] = ok ? f_transform(raw[k].w, sg, ik, shift_in, do_t) : 0.f;
      }
    }
    if (hly < 66) sy[hly][hlx] = hok ? f_transform(hv, sg, ik, shift_in, do_t) : 0.f;
  } else {
    for (int t = tid; t < 66 * 66; t += 256) {
      const int ly = t / 66, lx = t - ly * 66;
      const int y = 2 * i0 - 1 + ly, x = 2 * j0 - 1 + lx;
      sy[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? f_transform(__ldg(src + (size_t)y * W + x), sg, ik, shift_in, do_t) : 0.f;
    }
  }
  __syncthreads();
  const int lj = lane, li0 = wrp * 4;
  const int j = j0 + lj;
  if (j >= OW || i0 + li0 >= OH) return;
  unsigned long long yv[10][2];
#pragma unroll
  for (int rr = 0; rr < 10; ++rr) {
    yv[rr][0] = *reinterpret_cast<const unsigned long long*>(&sy[2 * li0 + rr][2 * lj]);
    yv[rr][1] = *reinterpret_cast<const unsigned long long*>(&sy[2 * li0 + rr][2 * lj + 2]);
  }
  // accumulators start at z * (sum of the in-image z taps) + shift (BN folded)
  unsigned long long acc[4][8];
  {
    const int cc = j == 0 ? 0 : (j == OW - 1 ? 2 : 1);
    float zc[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) zc[co] = fmaf(z, szs[1][cc][co], fc.shift[co]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int i = i0 + li0 + p;
      if (i == 0 || i == OH - 1) {                        // warp-uniform
        const int rc = i == 0 ? 0 : 2;
#pragma unroll
        for (int co = 0; co < 8; ++co) acc[p][co] = f_pack2(fmaf(z, szs[rc][cc][co], fc.shift[co]), 0.f);
      } else {
#pragma unroll
        for (int co = 0; co < 8; ++co) acc[p][co] = f_pack2(zc[co], 0.f);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int qp = 0; qp < 2; ++qp) {
      unsigned long long w2[8];
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(&sw[r][qp][2 * c2]);
        w2[2 * c2] = t.x; w2[2 * c2 + 1] = t.y;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int co = 0; co < 8; ++co) acc[p][co] = f_fma2(yv[2 * p + r][qp], w2[co], acc[p][co]);
    }
  // ReLU / LeakyReLU / PReLU with a slope in [0, 1] are max(a, slope * a); channels >= cout have zero weights, tap
  // sums and shift, so they come out as 0 without a mask
  const bool lin_act = fc.act == BP_ACT_RELU || fc.act == BP_ACT_LEAKY || fc.act == BP_ACT_PRELU || fc.act == BP_ACT_NONE;
  const float slope = fc.act == BP_ACT_RELU ? 0.f : (fc.act == BP_ACT_NONE ? 1.f : fc.act_param);
  const bool max_act = lin_act && slope >= 0.f && slope <= 1.f;
  // store geometry: the x part of the (shifted space-to-depth) address is the same for the four rows
  const int obs = (ob & (ob - 1)) == 0 ? 31 - __clz(ob) : -1;
  const int half = ob >> 1, Hs = OH / ob + 1, Ws = OW / ob + 1;
  size_t xpart, rowpitch;
  if (ob == 1) {
    xpart = (size_t)j; rowpitch = (size_t)OW;
  } else {
    const int xx = j + half;
    const int bx = obs >= 0 ? xx >> obs : xx / ob, sx = xx - bx * ob;
    xpart = (size_t)bx * ob * ob + sx; rowpitch = (size_t)Ws * ob * ob;
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = i0 + li0 + p;
    if (i >= OH) break;
    uint32_t pk[4];
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      float a0 = f_sum2(acc[p][2 * c2]), a1 = f_sum2(acc[p][2 * c2 + 1]);
      if (max_act) {
        a0 = fmaxf(a0, a0 * slope);
        a1 = fmaxf(a1, a1 * slope);
      } else if (lin_act) {
        a0 = a0 >= 0.f ? a0 : a0 * slope;
        a1 = a1 >= 0.f ? a1 : a1 * slope;
      } else {
        a0 = 2 * c2 < fc.cout ? f_act(a0, fc.act, fc.act_param) : 0.f;
        a1 = 2 * c2 + 1 < fc.cout ? f_act(a1, fc.act, fc.act_param) : 0.f;
      }
      pk[c2] = f_pack16(a0, a1, fmt);
    }
    size_t o;
    if (ob == 1) {
      o = (((size_t)n * OH + i) * rowpitch + xpart) * (size_t)oCp;
    } else {
      const int yy = i + half;
      const int by = obs >= 0 ? yy >> obs : yy / ob, sy_ = yy - by * ob;
      o = (((size_t)n * Hs + by) * rowpitch + (size_t)sy_ * ob + xpart) * (size_t)oCp;
    }
    *reinterpret_cast<uint4*>(out + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

int launch_front_prior_conv(const float* tiles, const ActDesc& out, const float* sigma, const float* aux,
                            const FrontConvParams& fc, float k_in, float shift_in, int do_transform, int H, int W, int nb,
                            int fmt, cudaStream_t s) {
  BP_REQUIRE(!out.f32 && out.Cp == 8 && out.H == H / 2 && out.W == W / 2 && (H % 2) == 0 && (W % 2) == 0, BP_E_INVALID,
             "front_prior_conv: bad output tensor");
  const dim3 grid((out.W + 31) / 32, (out.H + 31) / 32, nb);
  front_prior_conv_kernel<<<grid, 256, 0, s>>>(tiles, static_cast<uint16_t*>(out.ptr), sigma, aux, fc, k_in, shift_in,
                                               do_transform, H, W, out.b, out.Cp, fmt);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ---- decoder input ---------------------------------------------------------------------------------
// 1 -> 1 channel transposed convolution with k = 2s, p = s/2: two taps per dimension
// (`ls` = log2(s) when the stride is a power of two, else -1: the divisions become shifts)
__device__ __forceinline__ int f_div(int x, int s, int ls) { return ls >= 0 ? x >> ls : x / s; }
__device__ __forceinline__ float up_at(const float* in, int ih, int iw, const float* w, int k, int s, int ls, int p, int oy,
                                       int ox) {
  const int qh = f_div(oy + p, s, ls), r0 = oy + p - qh * s, qw = f_div(ox + p, s, ls), c0 = ox + p - qw * s;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int iy = qh - a;
    if (iy < 0 || iy >= ih) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ix = qw - b;
      if (ix < 0 || ix >= iw) continue;
      acc = fmaf(in[iy * iw + ix], w[(r0 + s * a) * k + (c0 + s * b)], acc);
    }
  }
  return acc;
}

// one CTA = one band of `rows` output rows of one tile.  Levels 0 .. nl-2 of the pyramid are recomputed per CTA
// for the rows the band needs (the whole pyramid of a tile is 2 MFLOP); the last level is fused with the store.
__global__ void __launch_bounds__(256) front_latent_kernel(const float* __restrict__ tiles, const float* __restrict__ latent,
                                                           uint2* __restrict__ out, const float* __restrict__ sigma,
                                                           const float* __restrict__ aux, const PzParams pz, float k_in,
                                                           float shift_in, int do_t, int H, int W, int lh, int lw, int rows,
                                                           int fmt) {
  extern __shared__ float sm[];
  const int n = blockIdx.y;
  const int y0 = blockIdx.x * rows;
  // level geometry
  int lvh[5], lvw[5];
  lvh[0] = lh; lvw[0] = lw;
  for (int i = 0; i < pz.nl; ++i) { lvh[i + 1] = lvh[i] * pz.s[i]; lvw[i + 1] = lvw[i] * pz.s[i]; }
  // row ranges needed at every level (inclusive), from the band downwards
  int lo[5], hi[5];
  lo[pz.nl] = y0; hi[pz.nl] = min(H, y0 + rows) - 1;
  int lgs[4];
  for (int i = 0; i < pz.nl; ++i) lgs[i] = (pz.s[i] & (pz.s[i] - 1)) == 0 ? 31 - __clz(pz.s[i]) : -1;
  for (int i = pz.nl - 1; i >= 0; --i) {
    lo[i] = max(0, f_div(lo[i + 1] + pz.p[i], pz.s[i], lgs[i]) - 1);
    hi[i] = min(lvh[i] - 1, f_div(hi[i + 1] + pz.p[i], pz.s[i], lgs[i]));
  }
  // shared buffers: level i rows [lo[i], hi[i]] x lvw[i]
  float* buf[5];
  {
    float* q = sm;
    for (int i = 0; i < pz.nl; ++i) { buf[i] = q; q += (hi[i] - lo[i] + 1) * lvw[i]; }
  }
  for (int i = threadIdx.x; i < (hi[0] - lo[0] + 1) * lvw[0]; i += blockDim.x)
    buf[0][i] = latent[(size_t)n * lh * lw + (size_t)lo[0] * lw + i];
  __syncthreads();
  for (int l = 0; l + 1 < pz.nl; ++l) {
    const int nr = hi[l + 1] - lo[l + 1] + 1, w1 = lvw[l + 1];
    const int lw1 = (w1 & (w1 - 1)) == 0 ? 31 - __clz(w1) : -1;
    for (int i = threadIdx.x; i < nr * w1; i += blockDim.x) {
      const int ry = f_div(i, w1, lw1), oy = lo[l + 1] + ry, ox = i - ry * w1;
      // rows of the source buffer are offset by lo[l]
      const float v = up_at(buf[l] - (size_t)lo[l] * lvw[l], hi[l] + 1, lvw[l], pz.w[l], pz.k[l], pz.s[l], lgs[l], pz.p[l], oy,
                            ox);
      buf[l + 1][i] = f_act(fmaf(v, pz.scale[l], pz.shift[l]), pz.act[l], pz.act_param[l]);
    }
    __syncthreads();
  }
  const int L = pz.nl - 1;
  const bool fast4 = pz.s[L] == 4 && pz.k[L] == 8 && pz.p[L] == 2 && (W & 3) == 0 && ((W >> 2) & ((W >> 2) - 1)) == 0 &&
                     lvw[L] * 4 == W;
  const float sg = do_t ? 1.f / sigma[n] : 1.f, ik = kLn2 / k_in;
  const uint32_t z16 = f_to16(aux[n], fmt);
  const int nr = hi[pz.nl] - lo[pz.nl] + 1;
  const float* src = buf[L] - (size_t)lo[L] * lvw[L];
  if (fast4) {
    // last level k8 s4 p2, four pixels (one input column period) per thread: the quad reads a 2 x 3 input window
    // and two weight rows chosen by oy mod 4, with no per-pixel index arithmetic
    //   ox + e: e = 0, 1 -> columns (2, 3) of in[xq], (6, 7) of in[xq - 1];  e = 2, 3 -> (0, 1) of in[xq + 1], (4, 5) of in[xq]
    __shared__ __align__(16) float swz[64];
    if (threadIdx.x < 64) swz[threadIdx.x] = pz.w[L][threadIdx.x];
    __syncthreads();
    const int W4 = W >> 2, w4s = 31 - __clz(W4), iw = lvw[L], ih = hi[L] + 1;
    const float bsc = pz.scale[L], bsh = pz.shift[L], ap = pz.act_param[L];
    const int actL = pz.act[L];
    for (int i = threadIdx.x; i < nr * W4; i += blockDim.x) {
      const int oy = y0 + (i >> w4s), xq = i & (W4 - 1), ox = xq * 4;
      const size_t p = ((size_t)n * H + oy) * W + ox;
      const float4 t = __ldg(reinterpret_cast<const float4*>(tiles + p));
      const int ty = oy + 2, r0 = ty & 3, qh = ty >> 2;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int iy = qh - a;
        if (iy < 0 || iy >= ih) continue;           // uniform over the row
        const float* row = src + iy * iw;
        const float vm = xq > 0 ? row[xq - 1] : 0.f, v0 = row[xq], vp = xq + 1 < iw ? row[xq + 1] : 0.f;
        const float4 wa = *reinterpret_cast<const float4*>(&swz[(r0 + 4 * a) * 8]);
        const float4 wb = *reinterpret_cast<const float4*>(&swz[(r0 + 4 * a) * 8 + 4]);
        acc[0] = fmaf(v0, wa.z, fmaf(vm, wb.z, acc[0]));
        acc[1] = fmaf(v0, wa.w, fmaf(vm, wb.w, acc[1]));
        acc[2] = fmaf(vp, wa.x, fmaf(v0, wb.x, acc[2]));
        acc[3] = fmaf(vp, wa.y, fmaf(v0, wb.y, acc[3]));
      }
      const float tv[4] = {t.x, t.y, t.z, t.w};
      uint32_t lo16[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = fmaf(acc[e], bsc, bsh);
        v = actL == BP_ACT_RELU ? fmaxf(v, 0.f) : f_act(v, actL, ap);
        lo16[e] = f_pack16(v, f_transform(tv[e], sg, ik, shift_in, do_t), fmt);      // one saturating cvt for the pair
      }
      uint4* o = reinterpret_cast<uint4*>(out + p);
      o[0] = make_uint4(lo16[0], z16, lo16[1], z16);
      o[1] = make_uint4(lo16[2], z16, lo16[3], z16);
    }
    return;
  }
  if ((W & 3) == 0) {
    // four pixels per thread: one 16-byte tile load, two 16-byte NHWC stores
    const int W4 = W >> 2;
    for (int i = threadIdx.x; i < nr * W4; i += blockDim.x) {
      const int oy = y0 + i / W4, ox = (i % W4) * 4;
      const size_t p = ((size_t)n * H + oy) * W + ox;
      const float4 t = __ldg(reinterpret_cast<const float4*>(tiles + p));
      const float tv[4] = {t.x, t.y, t.z, t.w};
      uint32_t lo16[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = up_at(src, hi[L] + 1, lvw[L], pz.w[L], pz.k[L], pz.s[L], lgs[L], pz.p[L], oy, ox + e);
        v = f_act(fmaf(v, pz.scale[L], pz.shift[L]), pz.act[L], pz.act_param[L]);
        lo16[e] = f_pack16(v, f_transform(tv[e], sg, ik, shift_in, do_t), fmt);
      }
      uint4* o = reinterpret_cast<uint4*>(out + p);
      o[0] = make_uint4(lo16[0], z16, lo16[1], z16);
      o[1] = make_uint4(lo16[2], z16, lo16[3], z16);
    }
    return;
  }
  for (int i = threadIdx.x; i < nr * W; i += blockDim.x) {
    const int oy = y0 + i / W, ox = i % W;
    float v = up_at(src, hi[L] + 1, lvw[L], pz.w[L], pz.k[L], pz.s[L], lgs[L], pz.p[L], oy, ox);
    v = f_act(fmaf(v, pz.scale[L], pz.shift[L]), pz.act[L], pz.act_param[L]);
    const size_t p = ((size_t)n * H + oy) * W + ox;
    const float yv = f_transform(tiles[p], sg, ik, shift_in, do_t);
    out[p] = make_uint2(f_pack16(v, yv, fmt), z16);
  }
}

int launch_front_latent(const float* tiles, const float* latent, const ActDesc& out, const float* sigma, const float* aux,
                        const PzParams& pz, float k_in, float shift_in, int do_transform, int lh, int lw, int nb, int fmt,
                        cudaStream_t s) {
  BP_REQUIRE(out.Cp == 4 && out.b == 1 && !out.f32, BP_E_INVALID, "front_latent: output must be a plain 4-channel NHWC tensor");
  const int rows = 16;
  // shared memory: every level's needed rows (upper bound: rows/stride products + 2 per level)
  size_t fl = 0;
  {
    int h = lh, w = lw, need = rows;
    std::vector<int> hs{h}, ws{w};
    for (int i = 0; i < pz.nl; ++i) { h *= pz.s[i]; w *= pz.s[i]; hs.push_back(h); ws.push_back(w); }
    for (int i = pz.nl - 1; i >= 0; --i) {
      need = std::min(hs[i], need / pz.s[i] + 3);
      fl += (size_t)need * ws[i];
    }
  }
  const size_t smem = fl * sizeof(float);
  BP_REQUIRE(smem <= 96 * 1024, BP_E_UNSUPPORTED, "front_latent: pyramid band needs %zu bytes of shared memory", smem);
  static bool attr = false;
  if (!attr) {
    BP_CUDA_TRY(cudaFuncSetAttribute(front_latent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  front_latent_kernel<<<dim3((out.H + rows - 1) / rows, nb), 256, smem, s>>>(
      tiles, latent, static_cast<uint2*>(out.ptr), sigma, aux, pz, k_in, shift_in, do_transform, out.H, out.W, lh, lw, rows,
      fmt);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ---- tail: 1 -> 1 channel k x k convolution (fp32) + activation + inverse transform ------------------
// CTA = 64 x 64 outputs (four 16-row passes over one shared window: a quarter of the CTAs, halo rows and barriers of
// a 64 x 16 tile), 4 along x per thread.  The shared tile keeps the image column x0 at a 16-byte aligned offset
// (4 floats in), so a thread's window is one float4 plus R scalars either side per row.
template <int K>
__global__ void __launch_bounds__(256) tail_stencil_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           long long out_bs, const TailParams tp,
                                                           const float* __restrict__ post_sigma, int H, int W) {
  constexpr int R = K / 2;
  constexpr int TX = 64, TY = 64, SW = TX + 8;      // row stride 72 floats: offset 4 + [-R, TX + R)
  static_assert(R <= 4, "halo must fit the 4-float margins");
  __shared__ __align__(16) float sh[TY + 2 * R][SW];
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const float* src = in + (size_t)n * H * W;
  if ((W & 3) == 0 && x0 + TX <= W) {
    // interior columns as aligned float4 (16 per row), the 2R halo columns as scalars
    for (int i = threadIdx.x; i < (TY + 2 * R) * (TX / 4); i += 256) {
      const int ly = i / (TX / 4), q = i - ly * (TX / 4);
      const int y = y0 + ly - R;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < H) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)y * W + x0) + q);
      *reinterpret_cast<float4*>(&sh[ly][4 + 4 * q]) = v;
    }
    for (int i = threadIdx.x; i < (TY + 2 * R) * 2 * R; i += 256) {
      const int ly = i / (R > 0 ? 2 * R : 1), h = i - ly * (2 * R);
      const int lx = h < R ? h : TX + h;                 // [0, R) left of the block, [TX + R, TX + 2R) right of it
      const int y = y0 + ly - R, x = x0 + lx - R;
      sh[ly][4 - R + lx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (size_t)y * W + x) : 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < (TY + 2 * R) * (TX + 2 * R); i += 256) {
      const int ly = i / (TX + 2 * R), lx = i - ly * (TX + 2 * R);
      const int y = y0 + ly - R, x = x0 + lx - R;
      sh[ly][4 - R + lx] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(src + (size_t)y * W + x) : 0.f;
    }
  }
  __syncthreads();
  const int tx = (threadIdx.x % 16) * 4;
  const float sg = tp.post ? post_sigma[n] : 1.f;
  const bool fast_tail = !tp.precise && tp.act == BP_ACT_SOFTPLUS && tp.post;
  const float post_a = tp.post_k * kLog2e, post_b = tp.post_shift * tp.post_k * kLog2e;
#pragma unroll 1
  for (int ty = threadIdx.x / 16; ty < TY; ty += 16) {
    const int y = y0 + ty;
    if (y >= H) break;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < K; ++r) {
      float v[4 + 2 * R];
      const float4 c4 = *reinterpret_cast<const float4*>(&sh[ty + r][4 + tx]);
      v[R] = c4.x; v[R + 1] = c4.y; v[R + 2] = c4.z; v[R + 3] = c4.w;
#pragma unroll
      for (int q = 0; q < R; ++q) {
        v[q] = sh[ty + r][4 + tx - R + q];
        v[R + 4 + q] = sh[ty + r][4 + tx + 4 + q];
      }
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const float w = tp.w[r * K + q];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(v[e + q], w, acc[e]);
      }
    }
    float o[4];
    if (fast_tail) {
      // 16-bit path, Softplus + inverse transform (the shipped networks): base-2 exp / log on the special-function
      // unit, constants folded -- softplus(v) = ln 2 * lg2(1 + 2^(v log2 e)),
      // (exp((x + shift) k) - 1) sigma = 2^(x k log2 e + shift k log2 e) * sigma - sigma
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = fmaf(acc[e], tp.scale, tp.shift);
        const float sp = v > 20.f ? v : kLn2 * f_lg2(1.f + f_ex2(v * kLog2e));
        o[e] = fmaf(f_ex2(fmaf(sp, post_a, post_b)), sg, -sg);
      }
    } else
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = fmaf(acc[e], tp.scale, tp.shift);
      if (tp.precise) {
        // fp32-accurate path: the library functions (<= 2 ulp), as the reference's torch / numpy float32 arithmetic
        v = f_act(v, tp.act, tp.act_param);
        if (tp.post) v = (expf((v + tp.post_shift) * tp.post_k) - 1.f) * sg;
      } else {
        // 16-bit path: fast exp / log (a few ulp) are far below its tolerance
        if (tp.act == BP_ACT_SOFTPLUS) v = v > 20.f ? v : __logf(1.f + __expf(v));
        else v = f_act(v, tp.act, tp.act_param);
        if (tp.post) v = (__expf((v + tp.post_shift) * tp.post_k) - 1.f) * sg;
      }
      o[e] = v;
    }
    float* dst = out + (size_t)n * out_bs + (size_t)y * W + x0 + tx;
    if (x0 + tx + 3 < W && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (x0 + tx + e < W) dst[e] = o[e];
    }
  }
}

int launch_tail_stencil(const float* in, float* out, long long out_bs, const TailParams& tp, const float* post_sigma, int H,
                        int W, int nb, cudaStream_t s) {
  const dim3 grid((W + 63) / 64, (H + 63) / 64, nb);
  switch (tp.k) {
    case 1: tail_stencil_kernel<1><<<grid, 256, 0, s>>>(in, out, out_bs, tp, post_sigma, H, W); break;
    case 3: tail_stencil_kernel<3><<<grid, 256, 0, s>>>(in, out, out_bs, tp, post_sigma, H, W); break;
    case 5: tail_stencil_kernel<5><<<grid, 256, 0, s>>>(in, out, out_bs, tp, post_sigma, H, W); break;
    case 7: tail_stencil_kernel<7><<<grid, 256, 0, s>>>(in, out, out_bs, tp, post_sigma, H, W); break;
    default:
      set_error("tail stencil: kernel size %d", tp.k);
      return BP_E_UNSUPPORTED;
  }
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

}  // namespace bp
