// fp32 kernels of the paint path (sm_100a): implicit-GEMM convolution on the FP32 pipes with the
// fused epilogue (folded batch-norm, residual add, activation, inverse data transform), the
// direct kernel for 1-2 output channels, the fused forward data transform, latent sampling and
// the variance-map accumulators.
//
// Replaces, per reference call site:
//   torch.nn.Conv2d / ConvTranspose2d / BatchNorm2d / ReLU / PReLU / Softplus modules built by
//   build_sequential                    baryon_painter/models/utils.py:128-147
//   ResidualBlock.forward               baryon_painter/models/utils.py:35-38
//   shift-log forward / inverse         baryon_painter/utils/data_transforms.py:76, 98
//   merge_aux_label                     baryon_painter/models/utils.py:159-182
//   CVAE.sample_z                       baryon_painter/models/cvae.py:63-66
#include <math.h>

#include <algorithm>

#include "bp_common.h"

namespace bp {

// ------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float apply_act(float v, int act, float p) {
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

__device__ __forceinline__ float apply_post(float v, int post, float k, float shift, float sigma) {
  if (post == POST_INV_SHIFT_LOG) return (expf((v + shift) * k) - 1.f) * sigma;
  return v;
}

// ------------------------------------------------------------------------------------------
// implicit GEMM, fp32.  Block tile 128 (pixels) x BN (channels) x 16 (k); 256 threads; each thread
// owns TM x TN accumulators.  A is gathered through the per-layer (channel, tap) table with zero
// fill at the image border; B is the packed [k][n] weight matrix.  Register-staged double buffer.
// ------------------------------------------------------------------------------------------
template <int BN, int TM, int TN>
__global__ void __launch_bounds__(256) igemm_f32_kernel(const ConvArgs a) {
  constexpr int BM = 128, BK = 16;
  constexpr int MG = BM / TM, NG = BN / TN;
  static_assert(MG * NG == 256, "thread tiling must cover the block tile");
  constexpr int BPT = (BK * BN + 255) / 256;  // B elements per thread per k-tile

  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const PhaseDev ph = a.phase[blockIdx.z];
  const int K = ph.K;
  const int4* __restrict__ ktab = a.ktab + ph.k_begin;
  const float* __restrict__ wmat = a.wmat + (size_t)ph.k_begin * a.npad + blockIdx.y * BN;
  const int ohw = a.OH * a.OW;
  const int M = a.nb * ohw;

  // ---- A-gather coordinates of this thread's pixel
  const int m_l = t & (BM - 1);
  const int kk0 = t >> 7;  // 0..1
  const int m = blockIdx.x * BM + m_l;
  const bool m_ok = m < M;
  int ih0 = 0, iw0 = 0;
  const float* inb = a.in;
  if (m_ok) {
    const int n = m / ohw, rem = m - n * ohw;
    const int i = rem / a.OW, j = rem - i * a.OW;
    ih0 = i * a.istride;
    iw0 = j * a.istride;
    inb = a.in + (size_t)n * a.in_bs + (size_t)ih0 * a.W + iw0;
  }

  float ra[8];
  float rb[BPT];

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kg = k0 + kk0 + 2 * q;
      float v = 0.f;
      if (m_ok && kg < K) {
        const int4 e = __ldg(ktab + kg);
        const int ih = ih0 + e.y, iw = iw0 + e.z;
        if ((unsigned)ih < (unsigned)a.H && (unsigned)iw < (unsigned)a.W) v = __ldg(inb + e.x);
      }
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      const int idx = t + 256 * q;
      const int kk = idx / BN, nn = idx - kk * BN;
      float v = 0.f;
      if (idx < BK * BN && k0 + kk < K) v = __ldg(wmat + (size_t)(k0 + kk) * a.npad + nn);
      rb[q] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) As[buf][kk0 + 2 * q][m_l] = ra[q];
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      const int idx = t + 256 * q;
      if (idx < BK * BN) {
        const int kk = idx / BN, nn = idx - kk * BN;
        Bs[buf][kk][nn] = rb[q];
      }
    }
  };

  const int tx = t % MG, ty = t / MG;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nkt = (K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
      {
        const float4 v0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
        av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
      }
      if constexpr (TM == 8) {
        const float4 v1 = *reinterpret_cast<const float4*>(&As[buf][kk][BM / 2 + tx * 4]);
        av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
      }
      if constexpr (TN == 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
        const float4 v1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + ty * 4]);
        bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
        bv[4] = v1.x; bv[5] = v1.y; bv[6] = v1.z; bv[7] = v1.w;
      } else if constexpr (TN == 4) {
        const float4 v0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
        bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
      } else if constexpr (TN == 2) {
        const float2 v0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][ty * 2]);
        bv[0] = v0.x; bv[1] = v0.y;
      } else {
        bv[0] = Bs[buf][kk][ty];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nkt) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
  const size_t ohwf = (size_t)a.OHF * a.OWF;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int ml = (TM == 8) ? ((i < 4) ? tx * 4 + i : BM / 2 + tx * 4 + (i - 4)) : tx * 4 + i;
    const int mm = blockIdx.x * BM + ml;
    if (mm >= M) continue;
    const int n = mm / ohw, rem = mm - n * ohw;
    const int oi = rem / a.OW, oj = rem - oi * a.OW;
    const size_t pix = (size_t)(oi * a.os + ph.ph) * a.OWF + (oj * a.os + ph.pw);
    const float sig = a.post ? a.post_sigma[n] : 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int nl;
      if (TN == 8) nl = (j < 4) ? ty * 4 + j : BN / 2 + ty * 4 + (j - 4);
      else nl = ty * TN + j;
      const int co = blockIdx.y * BN + nl;
      if (co >= a.cout) continue;
      float v = fmaf(acc[i][j], a.scale[co], a.shift[co]);
      if (a.skip) v += a.skip[(size_t)n * a.skip_bs + (size_t)co * ohwf + pix];
      v = apply_act(v, a.act, a.act_param);
      v = apply_post(v, a.post, a.post_k, a.post_shift, sig);
      a.out[(size_t)n * a.out_bs + (size_t)co * ohwf + pix] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// direct kernel for 1-2 output channels (latent upsampler p_z_in, prior head, conv5 8->1,
// conv3 1->1): one thread per output pixel over the FULL output grid; the phase (for transposed
// convolutions) is looked up per thread.  Memory/L1 bound by construction.
// ------------------------------------------------------------------------------------------
template <int CO>
__global__ void __launch_bounds__(256) direct_small_kernel(const ConvArgs a, int total_rows) {
  extern __shared__ float wsm[];  // [total_rows][CO]
  for (int i = threadIdx.x; i < total_rows * CO; i += blockDim.x) {
    const int r = i / CO, c = i - r * CO;
    wsm[i] = a.wmat[(size_t)r * a.npad + c];
  }
  __syncthreads();
  const size_t ohwf = (size_t)a.OHF * a.OWF;
  const size_t total = (size_t)a.nb * ohwf;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(g / ohwf);
    const int rem = (int)(g - (size_t)n * ohwf);
    const int oy = rem / a.OWF, ox = rem - oy * a.OWF;
    const int py = oy % a.os, px = ox % a.os;
    const PhaseDev ph = a.phase[py * a.os + px];
    const int ih0 = (oy / a.os) * a.istride, iw0 = (ox / a.os) * a.istride;
    const float* inb = a.in + (size_t)n * a.in_bs + (size_t)ih0 * a.W + iw0;
    const int4* kt = a.ktab + ph.k_begin;
    const float* w = wsm + ph.k_begin * CO;
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    for (int k = 0; k < ph.K; ++k) {
      const int4 e = __ldg(kt + k);
      const int ih = ih0 + e.y, iw = iw0 + e.z;
      if ((unsigned)ih < (unsigned)a.H && (unsigned)iw < (unsigned)a.W) {
        const float x = __ldg(inb + e.x);
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] = fmaf(x, w[k * CO + c], acc[c]);
      }
    }
    const float sig = a.post ? a.post_sigma[n] : 0.f;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      if (c >= a.cout) break;
      float v = fmaf(acc[c], a.scale[c], a.shift[c]);
      if (a.skip) v += a.skip[(size_t)n * a.skip_bs + (size_t)c * ohwf + rem];
      v = apply_act(v, a.act, a.act_param);
      v = apply_post(v, a.post, a.post_k, a.post_shift, sig);
      a.out[(size_t)n * a.out_bs + (size_t)c * ohwf + rem] = v;
    }
  }
}

template <int BN, int TM, int TN>
static int launch_igemm(const Layer& l, const ConvArgs& a, cudaStream_t s) {
  const long long M = (long long)a.nb * a.OH * a.OW;
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)((a.cout + BN - 1) / BN), (unsigned)a.nphase);
  igemm_f32_kernel<BN, TM, TN><<<grid, 256, 0, s>>>(a);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

int launch_conv_f32(const Layer& l, ConvArgs& a, cudaStream_t s) {
  a.ktab = l.ktab;
  a.wmat = l.wmat;
  a.scale = l.scale;
  a.shift = l.shift;
  a.H = l.H; a.W = l.W; a.OH = l.OH; a.OW = l.OW; a.OHF = l.OHF; a.OWF = l.OWF;
  a.istride = l.istride; a.os = l.os;
  a.cout = l.d.cout; a.npad = l.npad;
  a.nphase = l.nphase;
  for (int i = 0; i < l.nphase; ++i) a.phase[i] = l.phase[i];
  a.act = l.d.act; a.act_param = l.d.act_param;
  const int cout = l.d.cout;
  if (cout <= 2) {
    int rows = 0;
    for (int i = 0; i < l.nphase; ++i) rows += l.phase[i].K;
    const size_t total = (size_t)a.nb * l.OHF * l.OWF;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    const size_t smem = (size_t)rows * (cout == 1 ? 1 : 2) * sizeof(float);
    BP_REQUIRE(smem <= 48 * 1024, BP_E_UNSUPPORTED, "direct kernel: %zu B of weights exceed 48 KB", smem);
    if (cout == 1) direct_small_kernel<1><<<blocks, 256, smem, s>>>(a, rows);
    else direct_small_kernel<2><<<blocks, 256, smem, s>>>(a, rows);
    launch_counter()++;
    BP_CUDA_TRY(cudaGetLastError());
    return BP_OK;
  }
  if (cout <= 8) return launch_igemm<8, 4, 1>(l, a, s);
  if (cout <= 16) return launch_igemm<16, 4, 2>(l, a, s);
  if (cout <= 32) return launch_igemm<32, 4, 4>(l, a, s);
  if (cout <= 64) return launch_igemm<64, 8, 4>(l, a, s);
  return launch_igemm<128, 8, 8>(l, a, s);
}

// ------------------------------------------------------------------------------------------
// fused forward transform + aux plane (reference data_transforms.py:76 and merge_aux_label):
// dst[n][y_channel] = ln(x/sigma_n + 1)/k - shift ; dst[n][aux_channel] = aux_n
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ tiles, float* __restrict__ dst,
                                                      long long dst_bs, int y_channel, int aux_channel,
                                                      const float* __restrict__ sigma_in,
                                                      const float* __restrict__ aux, float inv_k, float shift_in,
                                                      int do_transform, int hw4) {
  const int n = blockIdx.y;
  const float sigma = do_transform ? sigma_in[n] : 1.f;
  const float av = aux[n];
  const float4* src = reinterpret_cast<const float4*>(tiles + (size_t)n * hw4 * 4);
  float4* dy = reinterpret_cast<float4*>(dst + (size_t)n * dst_bs + (size_t)y_channel * hw4 * 4);
  float4* da = aux_channel >= 0 ? reinterpret_cast<float4*>(dst + (size_t)n * dst_bs + (size_t)aux_channel * hw4 * 4)
                                : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += gridDim.x * blockDim.x) {
    float4 v = __ldg(src + i);
    if (do_transform) {
      // same operation order as the reference: x/std, +1, log, /k
      v.x = logf(v.x / sigma + 1.f) * inv_k - shift_in;
      v.y = logf(v.y / sigma + 1.f) * inv_k - shift_in;
      v.z = logf(v.z / sigma + 1.f) * inv_k - shift_in;
      v.w = logf(v.w / sigma + 1.f) * inv_k - shift_in;
    }
    dy[i] = v;
    if (da) da[i] = make_float4(av, av, av, av);
  }
}

int launch_prepare(const float* tiles, float* dst, long long dst_bs, int y_channel, int aux_channel,
                   const float* sigma_in, const float* aux, float k_in, float shift_in, int do_transform,
                   int nb, int hw, cudaStream_t s) {
  BP_REQUIRE(hw % 4 == 0, BP_E_INVALID, "tile area must be a multiple of 4");
  const int hw4 = hw / 4;
  int bx = (hw4 + 255) / 256;
  if (bx > 64) bx = 64;
  prepare_kernel<<<dim3(bx, nb), 256, 0, s>>>(tiles, dst, dst_bs, y_channel, aux_channel, sigma_in, aux,
                                             1.f / k_in, shift_in, do_transform, hw4);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// latent sampling (reference cvae.py:63-66): z = mu + eps*(exp(lv/2) + min_z_var)
// prior_out is [nb][2][hw] = (z_mu, z_log_var).  mode: BP_LATENT_EPS reads eps, BP_LATENT_SEED
// draws eps from a counter-based generator (splitmix64 -> Box-Muller) keyed by (seed, offset+idx).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ float counter_normal(uint64_t seed, uint64_t counter) {
  const uint64_t r = splitmix64(seed ^ splitmix64(counter));
  const float u1 = ((float)(uint32_t)(r >> 40) + 1.f) * (1.f / 16777216.f);  // (0,1]
  const float u2 = (float)(uint32_t)((r >> 8) & 0xFFFFFFu) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void rng_normal_kernel(float* __restrict__ out, uint64_t seed, uint64_t offset, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = counter_normal(seed, offset + i);
}
int launch_rng_normal(float* out, uint64_t seed, uint64_t offset, size_t n, cudaStream_t s) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  rng_normal_kernel<<<blocks > 0 ? blocks : 1, 256, 0, s>>>(out, seed, offset, n);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

__global__ void sample_z_kernel(const float* __restrict__ prior_out, const float* __restrict__ eps,
                                float* __restrict__ latent, float* __restrict__ mu_out,
                                float* __restrict__ lv_out, float min_z_var, int nb, int hw, int mode,
                                uint64_t seed, uint64_t offset) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nb * hw) return;
  const int n = g / hw, p = g - n * hw;
  const float mu = prior_out[(size_t)n * 2 * hw + p];
  const float lv = prior_out[(size_t)n * 2 * hw + hw + p];
  float e;
  if (mode == BP_LATENT_EPS) {
    e = eps[g];
  } else {
    e = counter_normal(seed, offset + (uint64_t)g);
  }
  latent[g] = mu + e * (expf(lv * 0.5f) + min_z_var);
  if (mu_out) { mu_out[g] = mu; lv_out[g] = lv; }
}

int launch_sample_z(const float* prior_out, const float* eps, float* latent, float* mu_out, float* lv_out,
                    float min_z_var, int nb, int hw, int mode, uint64_t seed, uint64_t offset, cudaStream_t s) {
  const int total = nb * hw;
  sample_z_kernel<<<(total + 255) / 256, 256, 0, s>>>(prior_out, eps, latent, mu_out, lv_out, min_z_var, nb, hw,
                                                     mode, seed, offset);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// evidence lower bound pieces (reference cvae.py:122-147): float64 block reductions + one atomic per block
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_to(double v, double* dst) {
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
    if (threadIdx.x == 0) atomicAdd(dst, v);
  }
}
// sum over (n, p) of (pm - zm)^2 / pv + exp(zlv) / pv + plv - zlv - 1, with (zm, zlv) = q[n][0/1][p] and
// (pm, plv) = prior[n][0/1][p], pv = exp(plv)
__global__ void kl_sum_kernel(const float* __restrict__ q, const float* __restrict__ prior, int nb, int hw, double* dst) {
  double acc = 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < nb * hw; g += gridDim.x * blockDim.x) {
    const int n = g / hw, p = g - n * hw;
    const float zm = q[(size_t)n * 2 * hw + p], zlv = q[(size_t)n * 2 * hw + hw + p];
    const float pm = prior[(size_t)n * 2 * hw + p], plv = prior[(size_t)n * 2 * hw + hw + p];
    const float pv = expf(plv);
    acc += (double)((pm - zm) * (pm - zm) / pv + expf(zlv) / pv + plv - zlv - 1.f);
  }
  block_sum_to(acc, dst);
}
__global__ void sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* dst) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += (double)(d * d);
  }
  block_sum_to(acc, dst);
}
int launch_kl_sum(const float* q, const float* prior, int nb, int hw, double* dst, cudaStream_t s) {
  kl_sum_kernel<<<std::max(1, std::min(148, (nb * hw + 255) / 256)), 256, 0, s>>>(q, prior, nb, hw, dst);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}
int launch_sqdiff_sum(const float* a, const float* b, size_t n, double* dst, cudaStream_t s) {
  sqdiff_sum_kernel<<<(int)std::max<size_t>(1, std::min<size_t>(148 * 8, (n + 255) / 256)), 256, 0, s>>>(a, b, n, dst);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// running mean / M2 over latent draws (variance maps, BASELINE config 4)
// ------------------------------------------------------------------------------------------
// The running moments are float64: painted pressure spans ~6 decades within a tile and the M2 update subtracts
// nearly equal numbers; the HBM cost (16 B per pixel and draw) is noise next to the painting.
__global__ void welford_kernel(const float* __restrict__ x, double* __restrict__ mean, double* __restrict__ m2,
                               double inv_count, int first, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)x[i];
    if (first) { mean[i] = v; m2[i] = 0.0; continue; }
    const double mu = mean[i], d = v - mu, mu2 = mu + d * inv_count;
    mean[i] = mu2;
    m2[i] += d * (v - mu2);
  }
}
// population variance (numpy.var default, what the oracle computes over its sample_P draws) and mean, as float32
__global__ void var_finalize_kernel(const double* __restrict__ mean, const double* __restrict__ m2, float* __restrict__ mean_out,
                                    float* __restrict__ var_out, double inv_count, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    mean_out[i] = (float)mean[i];
    var_out[i] = (float)(m2[i] * inv_count);
  }
}

int launch_welford(const float* x, double* mean, double* m2, int count, size_t n, cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  welford_kernel<<<blocks, 256, 0, s>>>(x, mean, m2, 1.0 / (double)count, count == 1, n);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}
int launch_var_finalize(const double* mean, const double* m2, float* mean_out, float* var_out, int count, size_t n,
                        cudaStream_t s) {
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  var_finalize_kernel<<<blocks, 256, 0, s>>>(mean, m2, mean_out, var_out, 1.0 / (double)count, n);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

}  // namespace bp
