// Tensor-core convolution path (sm_100a): implicit-GEMM convolution / transposed convolution on
// tcgen05.mma with the accumulator in TMEM.
//
//   D[m][n] = sum_k A[m][k] * B[k][n]     m = output pixel (128 per CTA tile, one TMEM lane each)
//                                          n = output channel (N = cout padded to 16, <= 128)
//                                          k = (filter tap, input channel), 64 per pipeline stage
//
// Activations live in HBM as 16-bit "c8" tensors  [sample][channel/8][H][W][8]  so that the 8
// channels of one pixel are one 16-byte unit: that unit is exactly one row of a UMMA core matrix
// (8 rows x 16 bytes, K-major, no swizzle), consecutive pixels are consecutive rows, and the
// epilogue (one pixel per thread) stores coalesced 16-byte units.
//
// Warp roles (288 threads, persistent CTAs, one per SM):
//   warps 0-3  producers: gather the A tile with 16-byte cp.async (zero fill at the image border
//              and for K padding) straight into the UMMA layout; thread 0 also fetches the packed
//              weight block of the stage with one bulk async copy (TMA engine, mbarrier completion)
//   warp  8    MMA issuer: waits for a full stage, issues 4 x tcgen05.mma (K = 16 each), commits the
//              stage back to the producers and, after the last k-block, the accumulator to the epilogue
//   warps 4-7  epilogue: tcgen05.ld the fp32 accumulator (double-buffered in TMEM so the next tile's
//              MMAs overlap), folded batch-norm scale/shift, residual add, activation, 16-bit pack, store
//
// Replaces torch.nn.Conv2d/ConvTranspose2d(+BatchNorm2d+ReLU/PReLU) and ResidualBlock of the reference
// (baryon_painter/models/utils.py:22-38, 128-147) for the layers that carry 99 % of the FLOPs.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include <algorithm>

#include "bp_tc.h"
#include "bp_tc.cuh"

namespace bp {

using namespace tc;

constexpr int TC_STAGES = 4;
constexpr int TC_LAG = 2;            // cp.async groups in flight per producer thread
constexpr int TC_BM = 128;
constexpr int TC_STAGE_BYTES = 16384;  // A: 8 chunks x 128 rows x 16 B;  B: 8 chunks x N(<=128) x 16 B
constexpr int TC_THREADS = 288;
constexpr int TC_MAX_CHUNKS = 512;
constexpr int TC_SMEM = 2 * TC_STAGES * TC_STAGE_BYTES + TC_MAX_CHUNKS * 8 + 1024 + 256;

struct TcPhase {
  int q_begin;  // first chunk of the phase in qtab (multiple of 8)
  int kb_begin; // first k-block of the phase in wpack
  int nkb;      // k-blocks (8 chunks each)
  int ph, pw;
};

struct TcArgs {
  const uint4* in;
  uint4* out16;
  float* out32;
  long long out32_bs;
  const uint4* skip;
  const uint4* wpack;
  const int2* qtab;
  const float* scale;
  const float* shift;
  int H, W, cg_in, OH, OW, OHF, OWF, istride, os, N, cout, cg_out;
  int nphase, total_chunks;
  TcPhase phase[kMaxPhases];
  int nb, tiles_per_phase, total_tiles;
  int act;
  float act_param;
  int fmt;  // 0 = f16, 1 = bf16
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;
  uint32_t tmem_cols;
};

__device__ __forceinline__ float tc_act(float v, int act, float p) {
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

__device__ __forceinline__ uint32_t pack16(float a, float b, int fmt) {
  if (fmt == 0) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int fmt) {
  if (fmt == 0) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + TC_STAGES * TC_STAGE_BYTES;
  int2* s_qtab = reinterpret_cast<int2*>(smem + 2 * TC_STAGES * TC_STAGE_BYTES);
  float* s_scale = reinterpret_cast<float*>(smem + 2 * TC_STAGES * TC_STAGE_BYTES + TC_MAX_CHUNKS * 8);
  float* s_shift = s_scale + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 128);
  uint64_t* full = bars;                  // [TC_STAGES]  producers (128 arrivals) + weight bytes -> MMA
  uint64_t* empty = bars + TC_STAGES;     // [TC_STAGES]  MMA commit -> producers
  uint64_t* tfull = bars + 2 * TC_STAGES; // [2]          MMA commit -> epilogue
  uint64_t* tempty = tfull + 2;           // [2]          epilogue (128 arrivals) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N;

  for (int i = tid; i < a.total_chunks; i += TC_THREADS) s_qtab[i] = a.qtab[i];
  for (int i = tid; i < 128; i += TC_THREADS) {
    s_scale[i] = i < a.cout ? a.scale[i] : 0.f;
    s_shift[i] = i < a.cout ? a.shift[i] : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 129);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ohw = a.OH * a.OW;
  const int M = a.nb * ohw;

  if (warp < 4) {
    // ===================== producers =====================
    const int p = tid;
    const uint32_t b_bytes = (uint32_t)N * 128u;
    uint32_t it = 0, pending = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int pi = tile / a.tiles_per_phase, mt = tile - pi * a.tiles_per_phase;
      const TcPhase P = a.phase[pi];
      const int m = mt * TC_BM + p;
      const bool ok = m < M;
      int ih0 = 0, iw0 = 0;
      long long base = 0;
      if (ok) {
        const int n = m / ohw, rem = m - n * ohw;
        const int i = rem / a.OW, j = rem - i * a.OW;
        ih0 = i * a.istride;
        iw0 = j * a.istride;
        base = ((long long)n * a.cg_in * a.H + ih0) * a.W + iw0;
      }
      for (int kb = 0; kb < P.nkb; ++kb, ++it) {
        const uint32_t s = it % TC_STAGES, par = (it / TC_STAGES) & 1u;
        mbar_wait(&empty[s], par ^ 1u);
        const int2* qt = s_qtab + P.q_begin + kb * 8;
        const uint32_t dst = smem_u32(sA + s * TC_STAGE_BYTES) + (uint32_t)p * 16u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int2 e = qt[c];
          const int dr = (int)(short)(e.y & 0xFFFF), ds = e.y >> 16;
          const bool v = ok && (unsigned)(ih0 + dr) < (unsigned)a.H && (unsigned)(iw0 + ds) < (unsigned)a.W;
          const uint4* src = v ? (a.in + (base + e.x)) : a.in;
          cp_async16(dst + (uint32_t)c * 2048u, src, v ? 16u : 0u);
        }
        cp_async_commit();
        if (p == 0) {
          mbar_arrive_expect_tx(&full[s], b_bytes);
          bulk_g2s(sB + s * TC_STAGE_BYTES, a.wpack + (size_t)(P.kb_begin + kb) * 8 * N, b_bytes, &full[s]);
        }
        ++pending;
        if (pending > TC_LAG) {
          cp_async_wait<TC_LAG>();
          fence_proxy_async();
          mbar_arrive(&full[(it - TC_LAG) % TC_STAGES]);
          --pending;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (pending) {
      mbar_arrive(&full[(it - pending) % TC_STAGES]);
      --pending;
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_f16(a.fmt, N);
    const uint64_t a_tmpl = make_smem_desc(0, a.lbo_a, a.sbo_a), b_tmpl = make_smem_desc(0, a.lbo_b, a.sbo_b);
    const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const int pi = tile / a.tiles_per_phase;
      const int nkb = a.phase[pi].nkb;
      const uint32_t as = tcount & 1u, apar = (tcount >> 1) & 1u;
      mbar_wait(&tempty[as], apar ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (uint32_t)N;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % TC_STAGES, par = (it / TC_STAGES) & 1u;
        mbar_wait(&full[s], par);
        tc_fence_after();
        if (lane == 0) {
          // descriptor templates + stage base in 16-byte units: two adds per MMA on the issuing thread
          uint64_t da = a_tmpl + (uint64_t)(sA16 + s * (TC_STAGE_BYTES >> 4));
          uint64_t db = b_tmpl + (uint64_t)(sB16 + s * (TC_STAGE_BYTES >> 4));
          const uint32_t acc0 = kb != 0 ? 1u : 0u;
          umma_f16(d_tmem, da, db, idesc, acc0);
#pragma unroll
          for (int j = 1; j < 4; ++j) {
            da += 256u;                       // 2 chunks x 128 rows x 16 B
            db += 2u * (uint32_t)N;           // 2 chunks x N rows x 16 B
            umma_f16(d_tmem, da, db, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (kb == nkb - 1) umma_commit(&tfull[as]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t tcount = 0;
    const size_t ohwf = (size_t)a.OHF * a.OWF;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++tcount) {
      const int pi = tile / a.tiles_per_phase, mt = tile - pi * a.tiles_per_phase;
      const TcPhase P = a.phase[pi];
      const uint32_t as = tcount & 1u, apar = (tcount >> 1) & 1u;
      mbar_wait(&tfull[as], apar);
      tc_fence_after();
      const int m = mt * TC_BM + ew * 32 + lane;
      const bool ok = m < M;
      int n = 0;
      size_t pix = 0;
      if (ok) {
        n = m / ohw;
        const int rem = m - n * ohw;
        const int oi = rem / a.OW, oj = rem - oi * a.OW;
        pix = (size_t)(oi * a.os + P.ph) * a.OWF + (oj * a.os + P.pw);
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * (uint32_t)N;
      for (int g = 0; g < N / 8; ++g) {
        uint32_t v[8];
        tmem_ld8(taddr + (uint32_t)g * 8u, v);
        tmem_ld_wait();
        if (ok && g < a.cg_out) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = fmaf(__uint_as_float(v[e]), s_scale[g * 8 + e], s_shift[g * 8 + e]);
          const size_t o16 = ((size_t)n * a.cg_out + g) * ohwf + pix;
          if (a.skip) {
            const uint4 sk = __ldg(a.skip + o16);
            const float2 s0 = unpack16(sk.x, a.fmt), s1 = unpack16(sk.y, a.fmt), s2 = unpack16(sk.z, a.fmt),
                         s3 = unpack16(sk.w, a.fmt);
            x[0] += s0.x; x[1] += s0.y; x[2] += s1.x; x[3] += s1.y;
            x[4] += s2.x; x[5] += s2.y; x[6] += s3.x; x[7] += s3.y;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = tc_act(x[e], a.act, a.act_param);
          if (a.out16) {
            uint4 o;
            o.x = pack16(x[0], x[1], a.fmt); o.y = pack16(x[2], x[3], a.fmt);
            o.z = pack16(x[4], x[5], a.fmt); o.w = pack16(x[6], x[7], a.fmt);
            a.out16[o16] = o;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = g * 8 + e;
              if (c < a.cout) a.out32[(size_t)n * a.out32_bs + (size_t)c * ohwf + pix] = x[e];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// layout conversion: fp32 NCHW <-> 16-bit c8
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_c8_kernel(const float* __restrict__ in, long long in_bs, int C, int hw,
                                                      uint4* __restrict__ out, int cg, int fmt) {
  const int n = blockIdx.z, g = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      x[e] = c < C ? __ldg(in + (size_t)n * in_bs + (size_t)c * hw + p) : 0.f;
    }
    uint4 o;
    o.x = pack16(x[0], x[1], fmt); o.y = pack16(x[2], x[3], fmt);
    o.z = pack16(x[4], x[5], fmt); o.w = pack16(x[6], x[7], fmt);
    out[((size_t)n * cg + g) * hw + p] = o;
  }
}

__global__ void __launch_bounds__(256) unpack_c8_kernel(const uint4* __restrict__ in, int cg, int C, int hw,
                                                        float* __restrict__ out, long long out_bs, int fmt) {
  const int n = blockIdx.z, g = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += gridDim.x * blockDim.x) {
    const uint4 v = __ldg(in + ((size_t)n * cg + g) * hw + p);
    const float2 a0 = unpack16(v.x, fmt), a1 = unpack16(v.y, fmt), a2 = unpack16(v.z, fmt), a3 = unpack16(v.w, fmt);
    const float x[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      if (c < C) out[(size_t)n * out_bs + (size_t)c * hw + p] = x[e];
    }
  }
}

int launch_pack_c8(const float* in, long long in_bs, int C, int hw, void* out, int nb, int fmt, cudaStream_t s) {
  const int cg = (C + 7) / 8;
  int bx = (hw + 255) / 256;
  if (bx > 128) bx = 128;
  pack_c8_kernel<<<dim3(bx, cg, nb), 256, 0, s>>>(in, in_bs, C, hw, reinterpret_cast<uint4*>(out), cg, fmt);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

int launch_unpack_c8(const void* in, int C, int hw, float* out, long long out_bs, int nb, int fmt, cudaStream_t s) {
  const int cg = (C + 7) / 8;
  int bx = (hw + 255) / 256;
  if (bx > 128) bx = 128;
  unpack_c8_kernel<<<dim3(bx, cg, nb), 256, 0, s>>>(reinterpret_cast<const uint4*>(in), cg, C, hw, out, out_bs, fmt);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

// ------------------------------------------------------------------------------------------
// host: per-layer packing and launch
// ------------------------------------------------------------------------------------------
struct TcLayer {
  int fmt = 0;
  int N = 0, cg_in = 0, cg_out = 0;
  int nphase = 0, total_chunks = 0, total_kb = 0;
  TcPhase phase[kMaxPhases];
  int2* qtab = nullptr;
  uint4* wpack = nullptr;
};

static uint16_t to16(float v, int fmt) {
  if (fmt == 0) {
    __half h = __float2half_rn(v);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

bool tc_layer_eligible(const bp_layer_desc& d, bool first_in_sequence) {
  if (d.cout < 2 || d.cout > 128) return false;
  if (d.cin < 8 && !(first_in_sequence && d.cin >= 2)) return false;   // 1-channel stencils stay on the FP32 pipes
  if (d.cin % 8 != 0 && !first_in_sequence) return false;
  const int cg_in = (d.cin + 7) / 8;
  int taps = d.kernel * d.kernel;
  if (d.kind == BP_CONVT) {
    const int per = (d.kernel + d.stride - 1) / d.stride;
    taps = per * per;
  }
  const int chunks = ((taps * cg_in + 7) / 8) * 8;
  if (d.kind == BP_CONVT) return chunks * d.stride * d.stride <= TC_MAX_CHUNKS;
  return chunks <= TC_MAX_CHUNKS;
}

int tc_pack_layer(Layer* l, int fmt) {
  const bp_layer_desc& d = l->d;
  TcLayer* t = new TcLayer();
  t->fmt = fmt;
  t->N = ((d.cout + 15) / 16) * 16;
  t->cg_in = (d.cin + 7) / 8;
  t->cg_out = (d.cout + 7) / 8;
  t->nphase = l->nphase;
  const int k = d.kernel, s = d.stride, p = d.pad, H = l->H, W = l->W;
  std::vector<int2> qtab;
  std::vector<uint16_t> wp;
  int kb_total = 0;
  for (int pi = 0; pi < l->nphase; ++pi) {
    struct Tap { int dr, ds, r, q; };
    std::vector<Tap> taps;
    if (d.kind == BP_CONV) {
      for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q) taps.push_back({r - p, q - p, r, q});
    } else {
      const int ph = l->phase[pi].ph, pw = l->phase[pi].pw;
      const int r0 = (ph + p) % s, qh = (ph + p) / s, c0 = (pw + p) % s, qw = (pw + p) / s;
      for (int a = 0; r0 + s * a < k; ++a)
        for (int b = 0; c0 + s * b < k; ++b) taps.push_back({qh - a, qw - b, r0 + s * a, c0 + s * b});
    }
    TcPhase& P = t->phase[pi];
    P.q_begin = (int)qtab.size();
    P.kb_begin = kb_total;
    P.ph = l->phase[pi].ph;
    P.pw = l->phase[pi].pw;
    const int chunks = (int)taps.size() * t->cg_in;
    P.nkb = (chunks + 7) / 8;
    if (P.nkb == 0) P.nkb = 1;
    wp.resize((size_t)(kb_total + P.nkb) * 8 * t->N * 8, 0);
    for (int q = 0; q < P.nkb * 8; ++q) {
      if (q >= chunks) {
        qtab.push_back(make_int2(0, 0x7FFF));  // dr = 32767 -> always out of bounds -> zero fill
        continue;
      }
      const Tap& tp = taps[q / t->cg_in];
      const int g = q % t->cg_in;
      BP_REQUIRE(tp.dr > -32768 && tp.dr < 32767 && tp.ds > -32768 && tp.ds < 32767, BP_E_UNSUPPORTED, "tap offset");
      qtab.push_back(make_int2((g * H + tp.dr) * W + tp.ds, (tp.dr & 0xFFFF) | (tp.ds << 16)));
      const int kb = P.kb_begin + q / 8, c = q % 8;
      for (int n = 0; n < d.cout; ++n)
        for (int e = 0; e < 8; ++e) {
          const int ch = g * 8 + e;
          if (ch >= d.cin) continue;
          const float w = d.kind == BP_CONV
                              ? l->host_weight[(((size_t)n * d.cin + ch) * k + tp.r) * k + tp.q]
                              : l->host_weight[(((size_t)ch * d.cout + n) * k + tp.r) * k + tp.q];
          wp[(((size_t)kb * 8 + c) * t->N + n) * 8 + e] = to16(w, fmt);
        }
    }
    kb_total += P.nkb;
  }
  t->total_chunks = (int)qtab.size();
  t->total_kb = kb_total;
  BP_REQUIRE(t->total_chunks <= TC_MAX_CHUNKS, BP_E_UNSUPPORTED, "layer needs %d k-chunks (> %d)", t->total_chunks,
             TC_MAX_CHUNKS);
  BP_CUDA_TRY(cudaMalloc(&t->qtab, qtab.size() * sizeof(int2)));
  BP_CUDA_TRY(cudaMalloc(&t->wpack, wp.size() * sizeof(uint16_t)));
  BP_CUDA_TRY(cudaMemcpy(t->qtab, qtab.data(), qtab.size() * sizeof(int2), cudaMemcpyHostToDevice));
  BP_CUDA_TRY(cudaMemcpy(t->wpack, wp.data(), wp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  l->tc = t;
  return BP_OK;
}

void tc_free_layer(Layer* l) {
  if (!l->tc) return;
  TcLayer* t = static_cast<TcLayer*>(l->tc);
  cudaFree(t->qtab);
  cudaFree(t->wpack);
  delete t;
  l->tc = nullptr;
}

int tc_layer_cg_out(const Layer& l) { return static_cast<const TcLayer*>(l.tc)->cg_out; }

static int g_num_sms = 0;
static bool g_attr_set = false;

int launch_conv_tc(const Layer& l, const void* in, void* out16, float* out32, long long out32_bs, const void* skip,
                   int nb, cudaStream_t s) {
  const TcLayer* t = static_cast<const TcLayer*>(l.tc);
  BP_REQUIRE(t != nullptr, BP_E_INVALID, "layer has no tensor-core packing");
  if (!g_attr_set) {
    int dev = 0;
    BP_CUDA_TRY(cudaGetDevice(&dev));
    BP_CUDA_TRY(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    BP_CUDA_TRY(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    g_attr_set = true;
  }
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.in = static_cast<const uint4*>(in);
  a.out16 = static_cast<uint4*>(out16);
  a.out32 = out32;
  a.out32_bs = out32_bs;
  a.skip = static_cast<const uint4*>(skip);
  a.wpack = t->wpack;
  a.qtab = t->qtab;
  a.scale = l.scale;
  a.shift = l.shift;
  a.H = l.H; a.W = l.W; a.cg_in = t->cg_in; a.OH = l.OH; a.OW = l.OW; a.OHF = l.OHF; a.OWF = l.OWF;
  a.istride = l.istride; a.os = l.os; a.N = t->N; a.cout = l.d.cout; a.cg_out = t->cg_out;
  a.nphase = t->nphase; a.total_chunks = t->total_chunks;
  for (int i = 0; i < t->nphase; ++i) a.phase[i] = t->phase[i];
  a.nb = nb;
  const long long M = (long long)nb * l.OH * l.OW;
  a.tiles_per_phase = (int)((M + TC_BM - 1) / TC_BM);
  a.total_tiles = a.tiles_per_phase * t->nphase;
  a.act = l.d.act; a.act_param = l.d.act_param; a.fmt = t->fmt;
  // K-major, no swizzle: rows at 16 B pitch inside 128 B core matrices; chunk (8 k-elements) stride
  // is 128 rows x 16 B for A and N rows x 16 B for B
  a.lbo_a = 2048; a.sbo_a = 128; a.lbo_b = (uint32_t)t->N * 16u; a.sbo_b = 128;
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)t->N) cols <<= 1;
  a.tmem_cols = cols;
  const int grid = std::min(a.total_tiles, g_num_sms > 0 ? g_num_sms : 148);
  conv_tc_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(a);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  return BP_OK;
}

}  // namespace bp
