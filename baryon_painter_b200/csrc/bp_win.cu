// Windowed tensor-core convolution (sm_100a): stride-1 convolutions and the sub-pixel phases of
// transposed convolutions as an implicit GEMM whose A operand is *never gathered*.
//
// The input window ("patch") of a CTA's output region is brought into shared memory once with bulk
// async copies (TMA engine; one contiguous row segment of 16-byte c8 units per copy) in the layout
//     patch[channel group][padded row][padded column]        16 B per entry (8 channels of a pixel)
// i.e. pixels of a padded row are consecutive 16-byte units.  That is already the K-major, no-swizzle
// UMMA operand layout (8-row core matrices of 16-byte rows at 16-byte pitch), so for filter tap
// (dy, dx) the A tile of 128 consecutive output positions is the same patch viewed through a shared
// memory descriptor whose start address is shifted by (dy*PW + dx)*16 bytes:
//     D[m][n] += sum_c patch[g(c)][m + shift(tap)][c%8] * W[tap][c][n]
// Output positions are enumerated in the padded-width flat domain (the PW-Wt halo columns of every
// row are computed and discarded), which keeps the shift uniform over a 128-row MMA tile that
// straddles image rows.  Every input element is read from L2 ~(1 + halo) times instead of k*k times.
//
// K pairing.  One tcgen05.mma (.kind::f16) consumes K = 16 = two 8-channel chunks whose distance is
// the descriptor's leading-byte-offset: two consecutive channel groups (LBO = group stride) for
// >= 16 input channels, or two different taps of a single 8-channel group (LBO = tap distance).
//
// Row packing (J > 1).  For layers with few output channels (conv7 16->8) the GEMM N is widened to
// J*cout by computing J vertically adjacent output rows from one A row: the tap set grows to
// (k+J-1) x k with zero-padded weights, N = J*cout, and only every J-th row is an M row.
//
// Warp roles (288 threads, persistent CTAs): warps 0-2 load/zero-fill patches (double buffered),
// warp 3 streams packed weight blocks (4-stage ring), warp 8 issues tcgen05.mma for all M tiles of
// the region per weight block, warps 4-7 run the epilogue from the double-buffered TMEM accumulator.
//
// Replaces the same reference modules as bp_tc.cu (baryon_painter/models/utils.py:22-38, 128-147).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <string.h>

#include <algorithm>
#include <map>

#include "bp_tc.h"
#include "bp_tc.cuh"

namespace bp {

using namespace tc;

constexpr int WIN_BSTAGES = 4;
constexpr int WIN_BSTAGE_BYTES = 16384;
constexpr int WIN_THREADS = 288;
constexpr int WIN_MAX_KS = 640;
constexpr int WIN_SMEM_LIMIT = 227 * 1024;
constexpr int WIN_FIXED_SMEM = WIN_BSTAGES * WIN_BSTAGE_BYTES + WIN_MAX_KS * 8 + 1024 + 256;

struct WinPhase {
  int ks_begin;     // first k-step of the phase in kstab
  int nks;          // k-steps, padded to a multiple of 4
  int stage_begin;  // first weight stage (4 k-steps) of the phase in wpack
  int ph, pw;       // output phase offsets
};

struct WinArgs {
  const uint4* in;
  uint4* out16;
  float* out32;
  long long out32_bs;
  const uint4* skip;
  const uint4* wpack;
  const uint64_t* kstab;  // per k-step: A descriptor template (start field = offset of the tap/group in 16 B units)
  const float* scale;
  const float* shift;
  int H, W, cg_in;
  int OH, OW, OHF, OWF, os;
  int N, coutp, cout, cg_out, J;
  int top, tb, left, PW, PP;     // halo rows above / above+below, halo cols left, patch pitch, units per group
  int Wt, nstrips, regs_per_strip, T_r, packed;
  int gb, nblk, ks_per_blk;
  int nphase, total_ks;
  WinPhase phase[kMaxPhases];
  int nb, total_regions;
  int act;
  float act_param;
  int fmt;
  uint32_t tmem_cols, patch_stage_bytes;
  long long* timing;  // optional [grid][10] cycle counters (BP_WIN_TIMING=1), else null
};

__device__ __forceinline__ float win_act(float v, int act, float p) {
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
__device__ __forceinline__ uint32_t win_pack16(float a, float b, int fmt) {
  if (fmt == 0) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 win_unpack16(uint32_t u, int fmt) {
  if (fmt == 0) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

#define TWAIT(slot, stmt)                                    \
  do {                                                       \
    if (a.timing) {                                          \
      const long long _t0 = clock64();                       \
      stmt;                                                  \
      tacc[slot] += clock64() - _t0;                         \
    } else {                                                 \
      stmt;                                                  \
    }                                                        \
  } while (0)

// region id -> coordinates (identical in every role)
struct Region {
  int pi, n, strip, rr;
  int r_first;   // first output row (logical grid) whose window the patch holds
  int f0;        // plain mode: first flat output position (strip-local, pitch PW)
  int rows;      // patch rows of this region
};
__device__ __forceinline__ Region decode_region(const WinArgs& a, int reg) {
  Region R;
  const int per_img = a.nstrips * a.regs_per_strip;
  const int per_phase = per_img * a.nb;
  R.pi = reg / per_phase;
  int rem = reg - R.pi * per_phase;
  R.n = rem / per_img;
  rem -= R.n * per_img;
  R.strip = rem / a.regs_per_strip;
  R.rr = rem - R.strip * a.regs_per_strip;
  if (a.packed) {
    R.f0 = 0;
    R.r_first = R.rr * a.T_r * a.J;
    R.rows = a.T_r * a.J + a.tb;
  } else {
    R.f0 = R.rr * a.T_r * 128;
    R.r_first = R.f0 / a.PW;
    const int r_last = (R.f0 + a.T_r * 128 - 1) / a.PW;
    R.rows = r_last - R.r_first + 1 + a.tb;
  }
  return R;
}

__global__ void __launch_bounds__(WIN_THREADS, 1) conv_win_kernel(const WinArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sP = smem;                                          // 2 patch stages
  uint8_t* sB = smem + 2 * a.patch_stage_bytes;                // WIN_BSTAGES weight stages
  uint64_t* s_ks = reinterpret_cast<uint64_t*>(sB + WIN_BSTAGES * WIN_BSTAGE_BYTES);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_ks) + WIN_MAX_KS * 8);
  float* s_shift = s_scale + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 128);
  uint64_t* full_p = bars;            // [2]  96 producer arrivals + 1 expect_tx arrival + copy bytes
  uint64_t* empty_p = bars + 2;       // [2]  MMA commit
  uint64_t* full_b = bars + 4;        // [4]  1 expect_tx arrival + weight bytes
  uint64_t* empty_b = bars + 8;       // [4]  MMA commit
  uint64_t* tfull = bars + 12;        // [2]  MMA commit
  uint64_t* tempty = bars + 14;       // [2]  128 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N;
  long long tacc[4] = {0, 0, 0, 0};
  const long long t_start = clock64();

  for (int i = tid; i < a.total_ks; i += WIN_THREADS) s_ks[i] = a.kstab[i];
  for (int i = tid; i < 128; i += WIN_THREADS) {
    const int co = i % a.coutp;
    s_scale[i] = (i < N && co < a.cout) ? a.scale[co] : 0.f;
    s_shift[i] = (i < N && co < a.cout) ? a.shift[co] : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_p[s], 97);
      mbar_init(&empty_p[s], 1);
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    for (int s = 0; s < WIN_BSTAGES; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
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

  if (warp < 3) {
    // ===================== patch producers (96 threads) =====================
    uint32_t pit = 0;
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int reg = blockIdx.x; reg < a.total_regions; reg += gridDim.x) {
      const Region R = decode_region(a, reg);
      const int x0 = R.strip * a.Wt - a.left;            // input column of patch column 0
      const int xs = max(0, x0), xe = min(a.W, x0 + a.PW);
      const int ncopy = max(0, xe - xs);
      // input rows of the patch: y = r_first - top + pr
      const int y_lo = R.r_first - a.top;
      const int v_lo = max(0, -y_lo), v_hi = min(R.rows, a.H - y_lo);   // patch rows [v_lo, v_hi) are in-image
      const int nvalid = max(0, v_hi - v_lo);
      for (int blk = 0; blk < a.nblk; ++blk, ++pit) {
        const uint32_t st = pit & 1u, par = (pit >> 1) & 1u;
        TWAIT(0, mbar_wait(&empty_p[st], par ^ 1u));
        uint4* stage = reinterpret_cast<uint4*>(sP + st * a.patch_stage_bytes);
        if (tid == 0) mbar_arrive_expect_tx(&full_p[st], (uint32_t)(a.gb * nvalid * ncopy) * 16u);
        const int items = a.gb * R.rows;
        for (int it = tid; it < items; it += 96) {
          const int gl = it / R.rows, pr = it - gl * R.rows;
          uint4* row = stage + (size_t)gl * a.PP + (size_t)pr * a.PW;
          const int y = y_lo + pr;
          if (y >= 0 && y < a.H && ncopy > 0) {
            const uint4* src = a.in + (((size_t)R.n * a.cg_in + (size_t)(blk * a.gb + gl)) * a.H + y) * a.W + xs;
            bulk_g2s(row + (xs - x0), src, (uint32_t)ncopy * 16u, &full_p[st]);
            for (int c = 0; c < xs - x0; ++c) row[c] = zero4;
            for (int c = xe - x0; c < a.PW; ++c) row[c] = zero4;
          } else {
            for (int c = 0; c < a.PW; ++c) row[c] = zero4;
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_p[st]);
      }
    }
  } else if (warp == 3) {
    // ===================== weight producer =====================
    if (lane == 0) {
      const uint32_t b_bytes = (uint32_t)N * 128u;
      uint32_t bit = 0;
      for (int reg = blockIdx.x; reg < a.total_regions; reg += gridDim.x) {
        const Region R = decode_region(a, reg);
        const WinPhase P = a.phase[R.pi];
        const int nst = P.nks / 4;
        for (int sg = 0; sg < nst; ++sg, ++bit) {
          const uint32_t st = bit % WIN_BSTAGES, par = (bit / WIN_BSTAGES) & 1u;
          TWAIT(0, mbar_wait(&empty_b[st], par ^ 1u));
          mbar_arrive_expect_tx(&full_b[st], b_bytes);
          bulk_g2s(sB + st * WIN_BSTAGE_BYTES, a.wpack + (size_t)(P.stage_begin + sg) * 8 * N, b_bytes, &full_b[st]);
        }
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    // One lane issues every tcgen05.mma of the CTA, so the per-MMA instruction count of this loop is
    // the kernel's critical path (a single thread retires a dependent instruction every ~4-6 cycles;
    // the tensor pipe needs a new M=128 x N=128 MMA every 64).  Descriptors are therefore not built
    // per MMA: the k-step table holds finished A-descriptor templates (LBO/SBO/version bits and the
    // tap offset in the start-address field) and the loop only adds the patch/tile base (in 16-byte
    // units; start-address fields never carry out of their 14 bits for addresses < 256 KB).
    const uint32_t idesc = make_idesc_f16(a.fmt, N);
    const uint64_t* s_tmpl = reinterpret_cast<const uint64_t*>(s_ks);
    const uint64_t db_tmpl = make_smem_desc(0, (uint32_t)N * 16u, 128u);
    const uint32_t sB16 = smem_u32(sB) >> 4, sP16 = smem_u32(sP) >> 4;
    const uint32_t pstage16 = a.patch_stage_bytes >> 4, bstage16 = WIN_BSTAGE_BYTES >> 4;
    const uint32_t bstep16 = 2u * (uint32_t)N;           // one k-step of weights = 2 chunks x N rows x 16 B
    const uint32_t tile_step = a.packed ? (uint32_t)(a.J * a.PW) : 128u;
    const int T_r = a.T_r;
    uint32_t pit = 0, bit = 0, rcount = 0;
    for (int reg = blockIdx.x; reg < a.total_regions; reg += gridDim.x, ++rcount) {
      const Region R = decode_region(a, reg);
      const WinPhase P = a.phase[R.pi];
      const uint32_t as = rcount & 1u, apar = (rcount >> 1) & 1u;
      TWAIT(0, mbar_wait(&tempty[as], apar ^ 1u));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * (uint32_t)(T_r * N);
      const uint32_t tile0 = a.packed ? 0u : (uint32_t)(R.f0 - R.r_first * a.PW);
      const uint64_t* tmpl = s_tmpl + P.ks_begin;
      int ks = 0;
      for (int blk = 0; blk < a.nblk; ++blk, ++pit) {
        const uint32_t pst = pit & 1u, ppar = (pit >> 1) & 1u;
        TWAIT(1, mbar_wait(&full_p[pst], ppar));
        tc_fence_after();
        const uint32_t a_add = sP16 + pst * pstage16 + tile0;
        const int ks_end = (blk == a.nblk - 1) ? P.nks : (blk + 1) * a.ks_per_blk;
        while (ks < ks_end) {
          // one weight stage = 4 k-steps (block boundaries are multiples of 4 except at the padded tail)
          const uint32_t bst = bit % WIN_BSTAGES, bpar = (bit / WIN_BSTAGES) & 1u;
          const int sub0 = ks & 3;
          if (sub0 == 0) {
            TWAIT(2, mbar_wait(&full_b[bst], bpar));
            tc_fence_after();
          }
          const int nsub = min(4 - sub0, ks_end - ks);
          if (lane == 0) {
            const uint64_t db0 = db_tmpl + (uint64_t)(sB16 + bst * bstage16);
            uint64_t t_next = tmpl[ks];
            for (int q = 0; q < nsub; ++q) {
              const uint64_t da0 = t_next + a_add;
              if (q + 1 < nsub) t_next = tmpl[ks + q + 1];
              const uint64_t db = db0 + (uint64_t)((uint32_t)(sub0 + q) * bstep16);
              const uint32_t acc = (ks + q) != 0 ? 1u : 0u;
              uint64_t da = da0;
              uint32_t dt = d_tmem;
              for (int mt = 0; mt < T_r; ++mt) {
                umma_f16(dt, da, db, idesc, acc);
                da += tile_step;
                dt += (uint32_t)N;
              }
            }
            if (sub0 + nsub == 4) umma_commit(&empty_b[bst]);
          }
          __syncwarp();
          ks += nsub;
          if (sub0 + nsub == 4) ++bit;
        }
        if (lane == 0) umma_commit(&empty_p[pst]);
        __syncwarp();
      }
      if (lane == 0) umma_commit(&tfull[as]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    uint32_t rcount = 0;
    const size_t ohwf = (size_t)a.OHF * a.OWF;
    for (int reg = blockIdx.x; reg < a.total_regions; reg += gridDim.x, ++rcount) {
      const Region R = decode_region(a, reg);
      const WinPhase P = a.phase[R.pi];
      const uint32_t as = rcount & 1u, apar = (rcount >> 1) & 1u;
      TWAIT(0, mbar_wait(&tfull[as], apar));
      tc_fence_after();
      const int m = ew * 32 + lane;
      for (int mt = 0; mt < a.T_r; ++mt) {
        int r, c;
        if (a.packed) {
          r = (R.rr * a.T_r + mt) * a.J;
          c = m;
        } else {
          const int f = R.f0 + mt * 128 + m;
          r = f / a.PW;
          c = f - r * a.PW;
        }
        const int xl = R.strip * a.Wt + c;
        const bool okc = c < a.Wt && xl < a.OW;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + as * (uint32_t)(a.T_r * N) + (uint32_t)(mt * N);
        for (int g = 0; g < N / 8; ++g) {
          uint32_t v[8];
          tmem_ld8(taddr + (uint32_t)g * 8u, v);
          tmem_ld_wait();
          const int n0 = g * 8;
          const int q = n0 / a.coutp, co0 = n0 - q * a.coutp;
          const int yl = r + q;
          if (okc && q < a.J && yl < a.OH && co0 < a.cout) {
            const size_t pix = (size_t)(yl * a.os + P.ph) * a.OWF + (size_t)(xl * a.os + P.pw);
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = fmaf(__uint_as_float(v[e]), s_scale[n0 + e], s_shift[n0 + e]);
            const size_t o16 = ((size_t)R.n * a.cg_out + (co0 >> 3)) * ohwf + pix;
            if (a.skip) {
              const uint4 sk = __ldg(a.skip + o16);
              const float2 s0 = win_unpack16(sk.x, a.fmt), s1 = win_unpack16(sk.y, a.fmt),
                           s2 = win_unpack16(sk.z, a.fmt), s3 = win_unpack16(sk.w, a.fmt);
              x[0] += s0.x; x[1] += s0.y; x[2] += s1.x; x[3] += s1.y;
              x[4] += s2.x; x[5] += s2.y; x[6] += s3.x; x[7] += s3.y;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = win_act(x[e], a.act, a.act_param);
            if (a.out16) {
              uint4 o;
              o.x = win_pack16(x[0], x[1], a.fmt); o.y = win_pack16(x[2], x[3], a.fmt);
              o.z = win_pack16(x[4], x[5], a.fmt); o.w = win_pack16(x[6], x[7], a.fmt);
              a.out16[o16] = o;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = co0 + e;
                if (ch < a.cout) a.out32[(size_t)R.n * a.out32_bs + (size_t)ch * ohwf + pix] = x[e];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }
  if (a.timing && lane == 0) {
    // per CTA: [0] total, [1] patch producers wait empty_p, [2] weight producer wait empty_b,
    // [3] mma wait tempty, [4] mma wait full_p, [5] mma wait full_b, [6] epilogue wait tfull
    long long* t = a.timing + (size_t)blockIdx.x * 10;
    if (warp == 0) { t[0] = clock64() - t_start; t[1] = tacc[0]; }
    if (warp == 3) t[2] = tacc[0];
    if (warp == 8) { t[3] = tacc[0]; t[4] = tacc[1]; t[5] = tacc[2]; }
    if (warp == 4) t[6] = tacc[0];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host: layer analysis, packing, launch
// ------------------------------------------------------------------------------------------
struct WinLayer {
  WinArgs proto;            // geometry / tables filled at pack time; pointers and nb at launch
  uint64_t* kstab = nullptr;
  uint4* wpack = nullptr;
  size_t smem = 0;
};

static uint16_t win_to16(float v, int fmt) {
  uint16_t u;
  if (fmt == 0) {
    __half h = __float2half_rn(v);
    memcpy(&u, &h, 2);
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    memcpy(&u, &h, 2);
  }
  return u;
}

bool win_layer_eligible(const bp_layer_desc& d, bool first_in_sequence) {
  if (getenv("BP_NO_WINDOW")) return false;
  if (d.kind == BP_CONV && d.stride != 1) return false;
  if (d.kind == BP_CONVT && (d.kernel % d.stride) != 0) return false;
  if (d.cout < 8 || d.cout > 128 || (d.cout % 8) != 0) return false;
  const int cg = (d.cin + 7) / 8;
  if (cg == 1) return d.cin == 8 || (first_in_sequence && d.cin >= 2);
  return (d.cin % 16) == 0;
}

int win_pack_layer(Layer* l, int fmt) {
  const bp_layer_desc& d = l->d;
  WinLayer* wl = new WinLayer();
  WinArgs& a = wl->proto;
  memset(&a, 0, sizeof(a));
  const int k = d.kernel, s = d.stride, p = d.pad;
  a.H = l->H; a.W = l->W; a.cg_in = (d.cin + 7) / 8;
  a.OH = l->OH; a.OW = l->OW; a.OHF = l->OHF; a.OWF = l->OWF; a.os = l->os;
  a.cout = d.cout; a.coutp = ((d.cout + 7) / 8) * 8; a.cg_out = a.coutp / 8;
  a.act = d.act; a.act_param = d.act_param; a.fmt = fmt;
  a.nphase = l->nphase;
  // row packing for narrow outputs of plain convolutions
  a.J = 1;
  if (d.kind == BP_CONV && a.coutp == 8 && !getenv("BP_NO_ROWPACK")) a.J = 4;
  a.packed = a.J > 1;
  a.N = ((a.J * a.coutp + 15) / 16) * 16;

  struct Tap { int dy, dx, r, q; };
  // per phase: original taps (input offset relative to the logical output pixel)
  std::vector<std::vector<Tap>> taps(l->nphase);
  int dy_min = 0, dy_max = 0, dx_min = 0, dx_max = 0;
  for (int pi = 0; pi < l->nphase; ++pi) {
    if (d.kind == BP_CONV) {
      for (int r = 0; r < k; ++r)
        for (int q = 0; q < k; ++q) taps[pi].push_back({r - p, q - p, r, q});
    } else {
      const int ph = l->phase[pi].ph, pw = l->phase[pi].pw;
      const int r0 = (ph + p) % s, qh = (ph + p) / s, c0 = (pw + p) % s, qw = (pw + p) / s;
      for (int aa = 0; r0 + s * aa < k; ++aa)
        for (int bb = 0; c0 + s * bb < k; ++bb) taps[pi].push_back({qh - aa, qw - bb, r0 + s * aa, c0 + s * bb});
    }
    for (const Tap& t : taps[pi]) {
      dy_min = std::min(dy_min, t.dy); dy_max = std::max(dy_max, t.dy + a.J - 1);
      dx_min = std::min(dx_min, t.dx); dx_max = std::max(dx_max, t.dx);
    }
  }
  a.top = -dy_min; a.tb = a.top + dy_max; a.left = -dx_min;
  const int right = dx_max;
  a.Wt = (a.packed || a.OW > 128) ? 128 : a.OW;
  a.PW = a.Wt + a.left + right;
  a.nstrips = (a.OW + a.Wt - 1) / a.Wt;
  // region size and patch stage
  a.T_r = std::min(8, std::max(1, 256 / a.N));
  if (a.packed) a.T_r = std::min(a.T_r, 2);
  a.gb = a.cg_in == 1 ? 1 : a.cg_in;
  int rows_max = 0;
  for (;;) {
    if (a.packed) rows_max = a.T_r * a.J + a.tb;
    else rows_max = (a.T_r * 128 + a.PW - 1) / a.PW + 1 + a.tb;
    a.PP = rows_max * a.PW + 136;                       // slack: discarded positions may read past the last row
    a.patch_stage_bytes = (uint32_t)(((size_t)a.gb * a.PP * 16 + 1023) / 1024 * 1024);
    if (2 * (size_t)a.patch_stage_bytes + WIN_FIXED_SMEM <= (size_t)WIN_SMEM_LIMIT) break;
    if (a.gb >= 4 && (a.gb / 2) % 2 == 0) a.gb /= 2;
    else if (a.T_r > 1) a.T_r /= 2;
    else {
      delete wl;
      set_error("window kernel: patch does not fit in shared memory");
      return BP_E_UNSUPPORTED;
    }
  }
  a.nblk = a.cg_in / a.gb;
  a.regs_per_strip = a.packed ? (a.OH + a.T_r * a.J - 1) / (a.T_r * a.J)
                              : (a.OH * a.PW + a.T_r * 128 - 1) / (a.T_r * 128);
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)(a.T_r * a.N)) cols <<= 1;
  a.tmem_cols = cols;

  // k-step tables and packed weights
  std::vector<uint64_t> kstab;
  std::vector<uint16_t> wp;
  int stage_total = 0;
  for (int pi = 0; pi < l->nphase; ++pi) {
    // extended taps (row packing): every (dy', dx) with dy' = dy + j
    struct XTap { int dy, dx; };
    std::vector<XTap> xt;
    {
      std::map<std::pair<int, int>, int> seen;
      for (const Tap& t : taps[pi])
        for (int j = 0; j < a.J; ++j)
          if (!seen.count({t.dy + j, t.dx})) {
            seen[{t.dy + j, t.dx}] = 1;
            xt.push_back({t.dy + j, t.dx});
          }
      std::sort(xt.begin(), xt.end(), [](const XTap& x, const XTap& y) {
        return x.dy != y.dy ? x.dy < y.dy : x.dx < y.dx;
      });
    }
    auto weight = [&](int dy, int dx, int j, int ch, int n) -> float {
      // original tap (dy - j, dx) of output row j, input channel ch, output channel n
      for (const Tap& t : taps[pi])
        if (t.dy == dy - j && t.dx == dx) {
          if (ch >= d.cin || n >= d.cout) return 0.f;
          return d.kind == BP_CONV ? l->host_weight[(((size_t)n * d.cin + ch) * k + t.r) * k + t.q]
                                   : l->host_weight[(((size_t)ch * d.cout + n) * k + t.r) * k + t.q];
        }
      return 0.f;
    };
    WinPhase& P = a.phase[pi];
    P.ks_begin = (int)kstab.size();
    P.stage_begin = stage_total;
    P.ph = l->phase[pi].ph; P.pw = l->phase[pi].pw;
    // chunk list per k-step: (tap index, group) x 2
    struct Chunk { int tap, g; };                       // tap < 0: zero chunk
    std::vector<std::pair<Chunk, Chunk>> steps;
    if (a.cg_in == 1) {
      for (size_t t = 0; t < xt.size(); t += 2)
        steps.push_back({{(int)t, 0}, {t + 1 < xt.size() ? (int)t + 1 : -1, 0}});
      a.ks_per_blk = (int)steps.size();
    } else {
      for (int blk = 0; blk < a.nblk; ++blk)
        for (size_t t = 0; t < xt.size(); ++t)
          for (int gp = 0; gp < a.gb; gp += 2) steps.push_back({{(int)t, blk * a.gb + gp}, {(int)t, blk * a.gb + gp + 1}});
      a.ks_per_blk = (int)xt.size() * (a.gb / 2);
    }
    const int real = (int)steps.size();
    P.nks = (real + 3) / 4 * 4;
    wp.resize((size_t)(stage_total + P.nks / 4) * 8 * a.N * 8, 0);
    for (int ks = 0; ks < P.nks; ++ks) {
      if (ks >= real) {
        kstab.push_back(make_smem_desc(0, 16, 128));    // zero weights; any in-bounds A address
        continue;
      }
      const Chunk c0 = steps[ks].first, c1 = steps[ks].second;
      const int sh0 = (a.top + xt[c0.tap].dy) * a.PW + (a.left + xt[c0.tap].dx);
      int a_off, lbo;
      if (a.cg_in == 1) {
        a_off = sh0;
        lbo = c1.tap >= 0 ? ((a.top + xt[c1.tap].dy) * a.PW + (a.left + xt[c1.tap].dx) - sh0) * 16 : 16;
        BP_REQUIRE(lbo > 0 && lbo < (1 << 18), BP_E_UNSUPPORTED, "window kernel: tap distance out of range");
      } else {
        a_off = (c0.g % a.gb) * a.PP + sh0;
        lbo = a.PP * 16;
        BP_REQUIRE(lbo < (1 << 18), BP_E_UNSUPPORTED, "window kernel: group stride out of range");
      }
      BP_REQUIRE(a_off >= 0 && a_off < 8192, BP_E_UNSUPPORTED, "window kernel: patch offset out of range");
      kstab.push_back(make_smem_desc(0, (uint32_t)lbo, 128) + (uint64_t)a_off);
      const int stage = stage_total + ks / 4, kin = ks % 4;
      for (int half = 0; half < 2; ++half) {
        const Chunk c = half ? c1 : c0;
        if (c.tap < 0) continue;
        for (int n = 0; n < a.J * a.coutp; ++n) {
          const int j = n / a.coutp, co = n % a.coutp;
          for (int e = 0; e < 8; ++e) {
            const float w = weight(xt[c.tap].dy, xt[c.tap].dx, j, c.g * 8 + e, co);
            if (w != 0.f) wp[((((size_t)stage * 4 + kin) * 2 + half) * a.N + n) * 8 + e] = win_to16(w, fmt);
          }
        }
      }
    }
    stage_total += P.nks / 4;
  }
  a.total_ks = (int)kstab.size();
  if (a.total_ks > WIN_MAX_KS) {
    delete wl;
    set_error("window kernel: %d k-steps exceed the table (%d)", a.total_ks, WIN_MAX_KS);
    return BP_E_UNSUPPORTED;
  }
  wl->smem = 2 * (size_t)a.patch_stage_bytes + WIN_FIXED_SMEM;
  BP_CUDA_TRY(cudaMalloc(&wl->kstab, kstab.size() * sizeof(uint64_t)));
  BP_CUDA_TRY(cudaMalloc(&wl->wpack, wp.size() * sizeof(uint16_t)));
  BP_CUDA_TRY(cudaMemcpy(wl->kstab, kstab.data(), kstab.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  BP_CUDA_TRY(cudaMemcpy(wl->wpack, wp.data(), wp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  l->win = wl;
  return BP_OK;
}

void win_free_layer(Layer* l) {
  if (!l->win) return;
  WinLayer* wl = static_cast<WinLayer*>(l->win);
  cudaFree(wl->kstab);
  cudaFree(wl->wpack);
  delete wl;
  l->win = nullptr;
}

static int g_win_sms = 0;
static bool g_win_attr = false;

int launch_conv_win(const Layer& l, const void* in, void* out16, float* out32, long long out32_bs, const void* skip,
                    int nb, cudaStream_t s) {
  const WinLayer* wl = static_cast<const WinLayer*>(l.win);
  BP_REQUIRE(wl != nullptr, BP_E_INVALID, "layer has no window packing");
  if (!g_win_attr) {
    int dev = 0;
    BP_CUDA_TRY(cudaGetDevice(&dev));
    BP_CUDA_TRY(cudaDeviceGetAttribute(&g_win_sms, cudaDevAttrMultiProcessorCount, dev));
    BP_CUDA_TRY(cudaFuncSetAttribute(conv_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_LIMIT));
    g_win_attr = true;
  }
  WinArgs a = wl->proto;
  a.in = static_cast<const uint4*>(in);
  a.out16 = static_cast<uint4*>(out16);
  a.out32 = out32;
  a.out32_bs = out32_bs;
  a.skip = static_cast<const uint4*>(skip);
  a.wpack = wl->wpack;
  a.kstab = wl->kstab;
  a.scale = l.scale;
  a.shift = l.shift;
  a.nb = nb;
  a.total_regions = a.nphase * nb * a.nstrips * a.regs_per_strip;
  const int grid = std::min(a.total_regions, g_win_sms > 0 ? g_win_sms : 148);
  static const bool timing = getenv("BP_WIN_TIMING") != nullptr;
  static long long* d_timing = nullptr;
  if (timing) {
    if (!d_timing) BP_CUDA_TRY(cudaMalloc(&d_timing, sizeof(long long) * 10 * 256));
    BP_CUDA_TRY(cudaMemsetAsync(d_timing, 0, sizeof(long long) * 10 * 256, s));
    a.timing = d_timing;
  }
  conv_win_kernel<<<grid, WIN_THREADS, wl->smem, s>>>(a);
  launch_counter()++;
  BP_CUDA_TRY(cudaGetLastError());
  if (timing) {
    BP_CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<long long> h(10 * 256);
    BP_CUDA_TRY(cudaMemcpy(h.data(), d_timing, sizeof(long long) * 10 * 256, cudaMemcpyDeviceToHost));
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < grid; ++c)
      for (int q = 0; q < 7; ++q) acc[q] += (double)h[c * 10 + q] / grid;
    fprintf(stderr,
            "[win] cin=%d cout=%d k=%d H=%d N=%d T_r=%d regions/cta=%.1f total=%.0f cyc | waits: patch-prod %.0f "
            "w-prod %.0f mma:tempty %.0f mma:full_p %.0f mma:full_b %.0f epi:tfull %.0f\n",
            l.d.cin, l.d.cout, l.d.kernel, l.H, a.N, a.T_r, (double)a.total_regions / grid, acc[0], acc[1], acc[2],
            acc[3], acc[4], acc[5], acc[6]);
  }
  return BP_OK;
}

}  // namespace bp
